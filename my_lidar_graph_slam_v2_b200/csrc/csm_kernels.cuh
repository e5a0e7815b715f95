/* csm_kernels.cuh -- CUDA kernels of the correlative scan matching hot path.
 * sm_100a only; compiled into libcsm_b200.so by csm_b200.cu.
 *
 * Kernel inventory (reference unit each one replaces, paths relative to the
 * reference repository):
 *   k_pyramid_stream2 PrecomputeGridMaps, all levels in one pass over the map, row rings
 *                     in registers (grid_map_builder.cpp:987-1012, util.hpp:369-424);
 *   k_pyramid_stream  the same with the rings in shared memory (rows % 32 != 0);
 *   k_pyramid_level   level h from level h-1 (small batches, odd shapes)
 *   k_sliding_max_tile / k_sliding_max
 *                     PrecomputeGridMap(map, win): separable in shared-memory tiles
 *                     (win <= 32) / generic (grid_map_builder.cpp:1044-1065)
 *   k_project         ScanData::HitPoint + PositionToIndex for every
 *                     (query, angle, beam) (sensor_data.hpp:190-203,
 *                     grid_map_geometry.cpp:113-122)
 *   k_rt_blocks       ComputeScore on the coarse map + EvaluateHighResolutionMap
 *                     for every coarse cell (scan_matcher_correlative.cpp:301-368)
 *   rt_replay         (in k_finalize) the sequential accept/skip decisions of
 *                     scan_matcher_correlative.cpp:161-197 replayed on keys
 *   k_bb_init / k_bb_score
 *                     branch-and-bound as level-synchronous frontier expansion
 *                     (scan_matcher_branch_bound.cpp:156-231)
 *   k_grid_window     exhaustive (dy, dx, dtheta) search, integer-shift path
 *   k_grid_general    same, per-candidate FP64 projection (arbitrary steps)
 *                     (scan_matcher_grid_search.cpp:118-142)
 *   k_finalize        winner decode, integer score + reference-order double score at
 *                     the winning pose, packed best word for the NCCL argmax
 *   k_setup           one-launch staging of a batch's descriptors and counters
 *   k_scatter_blocks  block-sparse upload -> dense level 0
 *   (csm_refine.cuh)  k_refine: ScanMatcherLinearSolver + CostSquareError on every found
 *                     pose / cost and covariance of a single-scan match; k_block_alloc
 *   (csm_window_tma.cuh) k_window_tma: grid-search window scoring from TMA-staged tiles
 */
#pragma once

#include "csm_device.cuh"
#include "csm_bounds.cuh"
#include "csm_b200.h"

namespace csm {

/* Pulls a small staged block from page-locked host memory (zero-copy read over
 * PCIe by the SMs). Used for every small descriptor / scan upload instead of
 * cudaMemcpyAsync so that they never queue behind the bulk grid uploads on the
 * host-to-device copy engine. */
__global__ void __launch_bounds__(256)
k_pull(uint4* __restrict__ dst, const uint4* __restrict__ src_host, unsigned int n16)
{
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
        dst[i] = src_host[i];
}

/* One launch that stages everything small a batch needs: pulls a contiguous
 * block of descriptors (queries, candidate angles, incumbents, root offsets,
 * per-call scan) from page-locked host memory, zeroes the per-batch counters,
 * optionally clears the packed best word, and fills the per-beam trig table of
 * a scan that arrives with this call. */
struct SetupArgs
{
    uint4* dst;
    const uint4* src_host;
    unsigned int n16;
    uint4* zero;
    unsigned int z16;
    unsigned long long* best_key;      /* cleared when non-null */
    const double* angles_host;         /* scan arriving with this call (else null) */
    double2* trig;
    int n_beams;
};

__global__ void __launch_bounds__(256)
k_setup(SetupArgs A)
{
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nth = gridDim.x * blockDim.x;
    for (unsigned int i = tid; i < A.n16; i += nth)
        A.dst[i] = A.src_host[i];
    for (unsigned int i = tid; i < A.z16; i += nth)
        A.zero[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0 && A.best_key != nullptr)
        *A.best_key = 0ull;
    if (A.angles_host != nullptr)
        for (unsigned int i = tid; i < (unsigned int)A.n_beams; i += nth) {
            double sn, cs;
            sincos(A.angles_host[i], &sn, &cs);
            A.trig[i] = make_double2(cs, sn);
        }
}

/* Block-sparse upload: expand the allocated blocks of a batch of maps into
 * their dense level-0 grids (zero-filled beforehand). The staging layout is
 * the reference's storage: one 2^k x 2^k block of u16 per allocated block,
 * row-major inside the block (grid_map.cpp:262-266, grid_binary_bayes.hpp),
 * block b of map m at data[prefix[m] + b], its position index[prefix[m] + b]
 * = block_row * block_cols + block_col. One thread moves 8 cells (16 bytes). */
struct ScatterArgs
{
    const uint4* data;
    const int* index;
    const int* prefix;
    uint16_t* dense;
    unsigned char* alloc;      /* per map nb_map bytes (zeroed beforehand): 1 = block allocated; or null */
    int nb_map;
    int log2bs, block_cols, cols;
    size_t map_cells;
    int saturated_unknown;
};

/* A cell at 65535 (ValueMax) reads as unknown. The reference's value -> probability tables hold
 * ValueMax - ValueMin + 1 = 65535 entries (grid_values.cpp:32-35, 72-74), so every read of a cell at 65535
 * -- the score (scan_matcher_correlative.cpp:313, score_function_pixel_accurate.cpp:34) as well as the
 * cost function's interpolation -- runs one element past the table, into the zero tail of the table's own
 * mmap'd chunk under glibc: probability 0.0, which is the unknown probability. The compiled reference
 * therefore treats such a cell exactly like an unknown one, and so does the matchers' view of every map
 * here (option "saturated_unknown", default on). The map builder keeps the true value next to it. */
__device__ __forceinline__ unsigned int saturated_as_unknown(unsigned int w)
{
    return w & ~__vcmpeq2(w, 0xffffffffu);
}
__device__ __forceinline__ uint4 saturated_as_unknown(uint4 v)
{
    v.x = saturated_as_unknown(v.x); v.y = saturated_as_unknown(v.y);
    v.z = saturated_as_unknown(v.z); v.w = saturated_as_unknown(v.w);
    return v;
}

/* the same for a dense upload, in place: 16 bytes per step, the words past the last whole 16 bytes one by one
 * (a map has an even number of columns, so its bytes are a multiple of 4) */
__global__ void __launch_bounds__(256)
k_saturated_unknown(unsigned int* __restrict__ words, size_t n_words)
{
    const size_t n16 = n_words / 4;
    uint4* __restrict__ cells = reinterpret_cast<uint4*>(words);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = cells[i];
        const uint4 w = saturated_as_unknown(v);
        if (v.x != w.x || v.y != w.y || v.z != w.z || v.w != w.w)
            cells[i] = w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n_words & 3)) {
        const size_t i = n16 * 4 + threadIdx.x;
        words[i] = saturated_as_unknown(words[i]);
    }
}

__global__ void __launch_bounds__(256)
k_scatter_blocks(ScatterArgs A)
{
    const int m = blockIdx.y;
    const int first = A.prefix[m];
    const int count = A.prefix[m + 1] - first;
    const int k = A.log2bs;
    const int chunks_per_row = 1 << (k - 3);
    const int chunks_per_block = chunks_per_row << k;
    const long long work = (long long)count * chunks_per_block;
    uint16_t* __restrict__ dense = A.dense + (size_t)m * A.map_cells;
    for (long long e = blockIdx.x * blockDim.x + threadIdx.x; e < work; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / chunks_per_block);
        const int c = (int)(e - (long long)b * chunks_per_block);
        const int r_in = c >> (k - 3), c_in = (c & (chunks_per_row - 1)) << 3;
        const int bi = __ldg(A.index + first + b);
        const int brow = bi / A.block_cols, bcol = bi - brow * A.block_cols;
        uint4 v = __ldg(A.data + (size_t)(first + b) * chunks_per_block + c);
        if (A.saturated_unknown)
            v = saturated_as_unknown(v);
        *reinterpret_cast<uint4*>(dense + (size_t)((brow << k) + r_in) * A.cols + (bcol << k) + c_in) = v;
        if (c == 0 && A.alloc != nullptr)
            A.alloc[(size_t)m * A.nb_map + bi] = 1;
    }
}

/* ------------------------------------------------------------------------ */
/* Precomputation                                                            */
/* ------------------------------------------------------------------------ */

struct PyrJob
{
    const uint16_t* base;   /* level 0, row-major */
    uint16_t*       levels; /* levels 1..hmax, contiguous, rows*cols each */
    int rows, cols;
};

/* Level h (window w = 2^h) from level h-1 with four taps.
 * out_h[r][c] = P_h[min(r, R-w)][min(c, C-w)] (far edge clamped),
 * P_h = max of P_{h-1} at (+0/+w/2, +0/+w/2); out_{h-1} equals P_{h-1} on
 * every index this reads. One thread per 2 cells. */
__global__ void __launch_bounds__(256)
k_pyramid_level(const PyrJob* __restrict__ jobs, int h)
{
    const PyrJob job = jobs[blockIdx.z];
    const int rows = job.rows, cols = job.cols;
    const size_t cells = (size_t)rows * cols;
    const uint16_t* __restrict__ src = (h == 1) ? job.base : job.levels + (size_t)(h - 2) * cells;
    uint16_t* __restrict__ dst = job.levels + (size_t)(h - 1) * cells;
    const int w = 1 << h, half = w >> 1;

    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2;
    const int r = blockIdx.y;
    if (c0 >= cols || r >= rows)
        return;
    const int rc = max(min(r, rows - w), 0);
    unsigned int out[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int c = c0 + k;
        const int cc = max(min(c, cols - w), 0);
        unsigned int v = ld_cell(src, rows, cols, rc, cc);
        v = max(v, ld_cell(src, rows, cols, rc + half, cc));
        v = max(v, ld_cell(src, rows, cols, rc, cc + half));
        v = max(v, ld_cell(src, rows, cols, rc + half, cc + half));
        out[k] = v;
    }
    const size_t o = (size_t)r * cols + c0;
    if (c0 + 1 < cols) {
        *reinterpret_cast<unsigned int*>(dst + o) = out[0] | (out[1] << 16);
    } else {
        dst[o] = (uint16_t)out[0];
    }
}

/* ---- streaming pyramid builder ------------------------------------------------
 * One CTA per map walks the rows bottom-up in blocks of 4 and produces every
 * level in one pass: level 0 is read from HBM exactly once (cp.async ring,
 * 8 blocks ahead), each level is written exactly once, nothing is re-read from
 * L2. Thread j owns cells (2j, 2j+1) of every row as one packed u16x2 word
 * (VIMNMX.U16x2). Per level h (w = 2^h, half = w/2):
 *     T_h[r]  = max(out_{h-1}[r][c], out_{h-1}[r][c + half])      (row buffer in smem)
 *     P_h[r]  = max(T_h[r], T_h[r + half])                         (ring of `half` rows of T_h;
 *                                                                   the slot read is the slot
 *                                                                   overwritten, thread-private)
 *     out_h[r][c] = P_h[min(r, R-w)][min(c, C-w)]                  (far edge clamped)
 * Rows below the map read as 0 (ring starts zeroed), like the reference's
 * ValueOr beyond the map. Requires cols <= 512, cols % 8 == 0, rows % 4 == 0,
 * hmax <= 6. */
constexpr int kPsThreads = 256;
constexpr int kPsRows = 4;
constexpr int kPsStages = 8;
constexpr int kPsInStride = 272;     /* words per staged input row (256 + zero pad) */

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, int src_bytes)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" :: "r"(d), "l"(gsrc), "r"(src_bytes));
}

__device__ __forceinline__ unsigned int splat_lo(unsigned int v) { return __byte_perm(v, v, 0x1010); }

/* One level of one 4-row block (H compile-time so that ring offsets, shifts
 * and the tap distance are immediates). a[] holds this thread's words of
 * out_{H-1} for the 4 rows and is replaced by its words of out_H. */
template <int H>
__device__ __forceinline__ void ps_level(unsigned int (&a)[kPsRows], const unsigned int* __restrict__ prev,
                                         int prev_stride, unsigned int* __restrict__ cur,
                                         unsigned int* __restrict__ ring, const PyrJob& job,
                                         int b, int j, bool in_map, size_t cells)
{
    /* in_map is false for the warm-up blocks of a row segment: they only fill the rings */
    constexpr int half = 1 << (H - 1);
    constexpr int w = 2 * half;
    constexpr int ring_off = half - 1;               /* 1 + 2 + ... + half/2 rows precede */
    const int R = job.rows, C = job.cols;
    const int csw_prev = (H >= 2) ? (max(C - half, 0) >> 1) : 0x3fffffff;   /* clamp word of level H-1 */
    const int csw = max(C - w, 0) >> 1;                                      /* clamp word of level H */
    const int rsrc = max(R - w, 0);
    unsigned int p[kPsRows];
#pragma unroll
    for (int rr = kPsRows - 1; rr >= 0; --rr) {
        const int r = b * kPsRows + rr;
        unsigned int t;
        if (H == 1) {
            t = __byte_perm(a[rr], prev[rr * prev_stride + j + 1], 0x5432);
        } else {
            const int k = j + (half >> 1);
            t = prev[rr * prev_stride + min(k, csw_prev)];
            if (k >= csw_prev) t = splat_lo(t);
        }
        t = __vmaxu2(a[rr], t);
        unsigned int* slot = ring + (ring_off + (r & (half - 1))) * 256 + j;
        const unsigned int old = *slot;
        *slot = t;
        p[rr] = __vmaxu2(t, old);
        cur[rr * 256 + j] = p[rr];
    }
    __syncthreads();
    /* out_H = P_H with the far edge clamped; only the clamped threads re-read */
    if (j >= csw) {
#pragma unroll
        for (int rr = 0; rr < kPsRows; ++rr) p[rr] = splat_lo(cur[rr * 256 + csw]);
    }
    const int r0 = b * kPsRows;
    if (in_map && r0 <= rsrc) {
        unsigned int* dst = reinterpret_cast<unsigned int*>(job.levels + (size_t)(H - 1) * cells + (size_t)r0 * C) + j;
        const int cw = C >> 1;
        if (r0 + kPsRows - 1 < rsrc) {
#pragma unroll
            for (int rr = 0; rr < kPsRows; ++rr) dst[rr * cw] = p[rr];
        } else {
            for (int rr = 0; rr < kPsRows; ++rr) {
                const int r = r0 + rr;
                if (r < rsrc) dst[rr * cw] = p[rr];
                else if (r == rsrc)
                    for (int r2 = r; r2 < R; ++r2) dst[(r2 - r0) * cw] = p[rr];
            }
        }
    }
#pragma unroll
    for (int rr = 0; rr < kPsRows; ++rr) a[rr] = p[rr];
}

/* `segs` CTAs share a map: segment s owns the 4-row blocks [s * nbs, (s + 1) * nbs) and first
 * walks the 16 blocks (64 rows >= 2^hmax - 1) above them without writing, which rebuilds the
 * ring state those rows depend on. Small batches thereby still fill the GPU. */
__global__ void __launch_bounds__(kPsThreads, 2)
k_pyramid_stream(const PyrJob* __restrict__ jobs, int hmax, int segs)
{
    extern __shared__ __align__(16) unsigned int ps_smem[];
    unsigned int* in_ring = ps_smem;                                         /* [stages][4][272] */
    unsigned int* rowbuf = in_ring + kPsStages * kPsRows * kPsInStride;      /* [2][4][256] */
    unsigned int* ring = rowbuf + 2 * kPsRows * 256;                         /* [63][256] */

    const PyrJob job = jobs[blockIdx.x / segs];
    const int seg = blockIdx.x % segs;
    const int R = job.rows, C = job.cols;
    const size_t cells = (size_t)R * C;
    const int j = threadIdx.x;
    const bool in_map = 2 * j < C;

    for (int i = j; i < 63 * 256; i += kPsThreads) ring[i] = 0u;
    for (int i = j; i < kPsStages * kPsRows * kPsInStride; i += kPsThreads) in_ring[i] = 0u;
    __syncthreads();

    const int nblocks = R / kPsRows;
    const int nbs = (nblocks + segs - 1) / segs;
    const int b_lo = seg * nbs;                                  /* lowest block of this segment */
    const int b_top = min(nblocks, b_lo + nbs) - 1;              /* highest block it writes */
    const int b_start = min(nblocks - 1, b_top + 16);            /* warm-up starts here */
    if (b_lo > b_top)
        return;
    /* producer: this thread's 16-byte chunk of a 4-row block */
    const int ld_row = j >> 6, ld_chunk = j & 63;
    auto prefetch = [&](int b) {
        if (b >= b_lo) {
            const int stage = b % kPsStages;
            unsigned int* dst = in_ring + (stage * kPsRows + ld_row) * kPsInStride + ld_chunk * 4;
            const bool ok = ld_chunk * 8 < C;
            const uint16_t* src = job.base + (size_t)(b * kPsRows + ld_row) * C + (ok ? ld_chunk * 8 : 0);
            cp_async_16(dst, src, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    for (int k = 0; k < kPsStages - 1; ++k)
        prefetch(b_start - k);

    unsigned int* buf0 = rowbuf;
    unsigned int* buf1 = rowbuf + kPsRows * 256;
    for (int b = b_start; b >= b_lo; --b) {
        const bool wr = in_map && b <= b_top;
        asm volatile("cp.async.wait_group %0;\n" :: "n"(kPsStages - 2));
        __syncthreads();                 /* block b landed; everyone is done with block b+1 */
        prefetch(b - (kPsStages - 1));   /* refills the stage block b+1 used */

        const unsigned int* in = in_ring + (b % kPsStages) * kPsRows * kPsInStride;
        unsigned int a[kPsRows];
#pragma unroll
        for (int rr = 0; rr < kPsRows; ++rr) a[rr] = in[rr * kPsInStride + j];
        ps_level<1>(a, in, kPsInStride, buf1, ring, job, b, j, wr, cells);
        if (hmax >= 2) ps_level<2>(a, buf1, 256, buf0, ring, job, b, j, wr, cells);
        if (hmax >= 3) ps_level<3>(a, buf0, 256, buf1, ring, job, b, j, wr, cells);
        if (hmax >= 4) ps_level<4>(a, buf1, 256, buf0, ring, job, b, j, wr, cells);
        if (hmax >= 5) ps_level<5>(a, buf0, 256, buf1, ring, job, b, j, wr, cells);
        if (hmax >= 6) ps_level<6>(a, buf1, 256, buf0, ring, job, b, j, wr, cells);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
}

/* ---- streaming pyramid builder, register rings -----------------------------------
 * Same walk as k_pyramid_stream, but the vertical rings (the last `half` rows of T_h of
 * every level, 63 words per thread for six levels) live in REGISTERS instead of shared
 * memory: a ring slot is (row mod half), so the main loop is unrolled over 8 consecutive
 * 4-row blocks (32 rows = the longest ring) and every slot index becomes a compile-time
 * constant. That removes two of the four shared-memory accesses per word and level (the
 * kernel is bound by instruction issue and LSU wavefronts, not by HBM) and all ring
 * address arithmetic. Requires rows % 32 == 0 on top of k_pyramid_stream's conditions;
 * segment boundaries are multiples of 8 blocks. */
constexpr int kPs2Group = 8;          /* blocks per unrolled group */

/* The rows of out_H at the far edge of the map (rows >= R - w all repeat row R - w): only the
 * topmost blocks of a map get here, so this stays out of line and the unrolled main loop small. */
template <int H>
__device__ __noinline__ void ps2_store_edge(uint4 pv, unsigned int* __restrict__ dst, int cw, int r0, int R)
{
    constexpr int w = 1 << H;
    const int rsrc = max(R - w, 0);
    const unsigned int p[kPsRows] = { pv.x, pv.y, pv.z, pv.w };
#pragma unroll
    for (int rr = 0; rr < kPsRows; ++rr) {
        const int r = r0 + rr;
        if (r < rsrc) dst[rr * cw] = p[rr];
        else if (r == rsrc)
            for (int r2 = r; r2 < R; ++r2) dst[(r2 - r0) * cw] = p[rr];
    }
}

/* Bound mode of the streaming builder: level H goes out as the u8 bound level of csm_bounds.cuh
 * (ceil(v / 257) per cell, tiles of kBlTileR x kBlTileC cells, zero padding never touched) instead of the
 * reference's u16 level. `out` = the map's bound allocation; a thread stores its two cells of a row as
 * one 16-bit word. */
template <int H, int FIX>
__device__ __forceinline__ unsigned char* ps2b_row(unsigned char* __restrict__ out, int R, int C, int r, int j)
{
    /* (a map's bound allocation is a few MiB: 32-bit offsets) */
    const unsigned int base = FIX ? (unsigned int)bl_level_offset(H, FIX, FIX) : (unsigned int)bl_level_offset(H, R, C);
    const unsigned int tpr = FIX ? (unsigned int)bl_tiles_per_row(H, FIX) : (unsigned int)bl_tiles_per_row(H, C);
    return out + (base + bl_cell((unsigned int)(r + bl_pad_r(H)), (unsigned int)(2 * j + bl_pad_c(H)), tpr));
}

/* (every level above the first horizontal maximum is computed on encoded values, 16 bits per cell) */
__device__ __forceinline__ unsigned short ps2b_pack(unsigned int word)
{
    return (unsigned short)__byte_perm(word, 0u, 0x4420);
}

template <int H, int FIX>
__device__ __noinline__ void ps2b_store_edge(uint4 pv, unsigned char* __restrict__ out, int R, int C, int r0, int j)
{
    constexpr int w = 1 << H;
    const int rsrc = max(R - w, 0);
    const unsigned int p[kPsRows] = { pv.x, pv.y, pv.z, pv.w };
#pragma unroll
    for (int rr = 0; rr < kPsRows; ++rr) {
        const int r = r0 + rr;
        if (r < rsrc) *reinterpret_cast<unsigned short*>(ps2b_row<H, FIX>(out, R, C, r, j)) = ps2b_pack(p[rr]);
        else if (r == rsrc)
            for (int r2 = r; r2 < R; ++r2)
                *reinterpret_cast<unsigned short*>(ps2b_row<H, FIX>(out, R, C, r2, j)) = ps2b_pack(p[rr]);
    }
}

/* One level of one 4-row block; K = position of the block inside its group of 8 (compile time:
 * ring slots are registers). PS = word stride of the rows of `prev`. a[] holds this thread's
 * words of out_{H-1} and leaves as out_H. dst0 = word j of row r0 of level 1 (BND: the map's bound
 * allocation). */
template <int H, int K, int PS, int FIX, bool BND>
__device__ __forceinline__ void ps2_level(unsigned int (&a)[kPsRows], const unsigned int* __restrict__ prev,
                                          unsigned int* __restrict__ cur, unsigned int (&ring)[63],
                                          unsigned int* __restrict__ dst0, size_t cells_w_, int cw_,
                                          int R_, int C_, int r0, int j, bool wr)
{
    /* FIX > 0: square maps of FIX x FIX cells, every stride and clamp is an immediate.
     * The row buffers (`cur`, and `prev` above level 1) hold the four rows of a word index side by side
     * ([256 words][4 rows]): a thread's tap and its own output are one 128-bit access each. Level 1 taps the
     * input ring, which cp.async fills row by row (PS = word stride of its rows). */
    static_assert(kPsRows == 4, "the four rows of a block travel as one uint4");
    const int R = FIX ? FIX : R_, C = FIX ? FIX : C_, cw = FIX ? FIX / 2 : cw_;
    const size_t cells_w = FIX ? (size_t)FIX * FIX / 2 : cells_w_;
    constexpr int half = 1 << (H - 1);
    constexpr int w = 2 * half;
    constexpr int ring_off = half - 1;
    /* horizontal tap: word j + half/2 of out_{H-1}, which is clamped at column C - half */
    const int csw_prev = (H >= 2) ? (max(C - half, 0) >> 1) : 0x3fffffff;
    const bool edge = (H >= 2) && (j + (half >> 1) >= csw_prev);
    unsigned int tp[kPsRows];
    if (H == 1) {
#pragma unroll
        for (int rr = 0; rr < kPsRows; ++rr) tp[rr] = prev[rr * PS + j + 1];
    } else {
        const uint4 t4 = reinterpret_cast<const uint4*>(prev)[edge ? csw_prev : j + (half >> 1)];
        tp[0] = t4.x; tp[1] = t4.y; tp[2] = t4.z; tp[3] = t4.w;
    }
#pragma unroll
    for (int rr = kPsRows - 1; rr >= 0; --rr) {
        unsigned int t = tp[rr];
        if (H == 1) t = __byte_perm(a[rr], t, 0x5432);
        else if (edge) t = splat_lo(t);
        t = __vmaxu2(a[rr], t);
        /* bound mode: level 0 itself is never written, and the maximum commutes with the monotone encoding,
         * so the cells are encoded once, after the first horizontal maximum was taken on the raw values */
        if (H == 1 && BND) t = bl_encode2p(t);
        const int slot = ring_off + ((K * kPsRows + rr) & (half - 1));     /* constant after unrolling */
        const unsigned int old = ring[slot];
        ring[slot] = t;
        a[rr] = __vmaxu2(t, old);
    }
    reinterpret_cast<uint4*>(cur)[j] = make_uint4(a[0], a[1], a[2], a[3]);
    __syncthreads();
    /* out_H = P_H with the far edge clamped along the row; only the clamped threads re-read */
    const int csw = max(C - w, 0) >> 1;
    if (j >= csw) {
        const uint4 e = reinterpret_cast<const uint4*>(cur)[csw];
        a[0] = splat_lo(e.x); a[1] = splat_lo(e.y); a[2] = splat_lo(e.z); a[3] = splat_lo(e.w);
    }
    if (wr && BND) {
        const int rsrc = max(R - w, 0);
        unsigned char* __restrict__ out = reinterpret_cast<unsigned char*>(dst0);
        if (r0 + kPsRows - 1 < rsrc) {
            /* r0 is a multiple of the tile height: this thread's two columns of the four rows are eight
             * consecutive bytes of the tile, the warp's 32 x 8 bytes two whole lines */
            static_assert(kPsRows == kBlTileR, "a 4-row block is one row of tiles");
            *reinterpret_cast<uint2*>(ps2b_row<H, FIX>(out, R, C, r0, j)) =
                make_uint2(__byte_perm(a[0], a[1], 0x6420), __byte_perm(a[2], a[3], 0x6420));
        } else if (r0 <= rsrc) {
            ps2b_store_edge<H, FIX>(make_uint4(a[0], a[1], a[2], a[3]), out, R, C, r0, j);
        }
    } else if (wr) {
        const int rsrc = max(R - w, 0);
        unsigned int* __restrict__ dst = dst0 + (size_t)(H - 1) * cells_w;
        if (r0 + kPsRows - 1 < rsrc) {
#pragma unroll
            for (int rr = 0; rr < kPsRows; ++rr) dst[rr * cw] = a[rr];
        } else if (r0 <= rsrc) {
            ps2_store_edge<H>(make_uint4(a[0], a[1], a[2], a[3]), dst, cw, r0, R);
        }
    }
}

template <int HMAX, int K, int FIX, bool BND>
__device__ __forceinline__ void ps2_block(unsigned int (&ring)[63], const unsigned int* __restrict__ in_ring,
                                          unsigned int* __restrict__ buf0, unsigned int* __restrict__ buf1,
                                          const PyrJob& job, int b, int j, bool wr, size_t cells_w)
{
    static_assert(kPsRows == 4, "ps2_store_edge passes the four rows of a block as one uint4");
    const unsigned int* in = in_ring + (b % kPsStages) * kPsRows * kPsInStride;
    const int R = FIX ? FIX : job.rows, C = FIX ? FIX : job.cols, cw = C >> 1, r0 = b * kPsRows;
    unsigned int* dst0 = BND ? reinterpret_cast<unsigned int*>(job.levels)
                             : reinterpret_cast<unsigned int*>(job.levels) + (size_t)r0 * cw + j;
    unsigned int a[kPsRows];
#pragma unroll
    for (int rr = 0; rr < kPsRows; ++rr) a[rr] = in[rr * kPsInStride + j];
    ps2_level<1, K, kPsInStride, FIX, BND>(a, in, buf1, ring, dst0, cells_w, cw, R, C, r0, j, wr);
    if (HMAX >= 2) ps2_level<2, K, 256, FIX, BND>(a, buf1, buf0, ring, dst0, cells_w, cw, R, C, r0, j, wr);
    if (HMAX >= 3) ps2_level<3, K, 256, FIX, BND>(a, buf0, buf1, ring, dst0, cells_w, cw, R, C, r0, j, wr);
    if (HMAX >= 4) ps2_level<4, K, 256, FIX, BND>(a, buf1, buf0, ring, dst0, cells_w, cw, R, C, r0, j, wr);
    if (HMAX >= 5) ps2_level<5, K, 256, FIX, BND>(a, buf0, buf1, ring, dst0, cells_w, cw, R, C, r0, j, wr);
    if (HMAX >= 6) ps2_level<6, K, 256, FIX, BND>(a, buf1, buf0, ring, dst0, cells_w, cw, R, C, r0, j, wr);
}

#ifndef CSM_PS2_MINB
#define CSM_PS2_MINB 2      /* 108 registers. Capped at 88 (or 80: three resident CTAs) the four-row words spill and the
                               kernel runs at 150 us per 256 maps instead of 130; a batch of 256 maps is 256 CTAs anyway */
#endif
template <int HMAX, int FIX, bool BND>
__global__ void __launch_bounds__(kPsThreads, (BND && HMAX <= 5) ? CSM_PS2_MINB : 2)
k_pyramid_stream2(const PyrJob* __restrict__ jobs, int segs)
{
    extern __shared__ __align__(16) unsigned int ps_smem[];
    unsigned int* in_ring = ps_smem;                                         /* [stages][4][272] */
    unsigned int* rowbuf = in_ring + kPsStages * kPsRows * kPsInStride;      /* [2][4][256] */

    const PyrJob job = jobs[blockIdx.x / segs];
    const int seg = blockIdx.x % segs;
    const int R = FIX ? FIX : job.rows, C = FIX ? FIX : job.cols;
    const size_t cells = (size_t)R * C;
    const int j = threadIdx.x;
    const bool in_map = 2 * j < C;

    unsigned int ring[63];
#pragma unroll
    for (int i = 0; i < 63; ++i) ring[i] = 0u;
    for (int i = j; i < kPsStages * kPsRows * kPsInStride; i += kPsThreads) in_ring[i] = 0u;
    __syncthreads();

    const int nblocks = R / kPsRows;                              /* a multiple of kPs2Group */
    const int ngroups = nblocks / kPs2Group;
    const int gps = (ngroups + segs - 1) / segs;                  /* groups per segment */
    const int b_lo = seg * gps * kPs2Group;
    const int b_top = min(nblocks, b_lo + gps * kPs2Group) - 1;
    const int b_start = min(nblocks - 1, b_top + 16);             /* 64 warm-up rows >= 2^hmax - 1 */
    if (b_lo > b_top)
        return;
    const int ld_row = j >> 6, ld_chunk = j & 63;
    auto prefetch = [&](int b) {
        if (b >= b_lo) {
            const int stage = b % kPsStages;
            unsigned int* dst = in_ring + (stage * kPsRows + ld_row) * kPsInStride + ld_chunk * 4;
            const bool ok = ld_chunk * 8 < C;
            const uint16_t* src = job.base + (size_t)(b * kPsRows + ld_row) * C + (ok ? ld_chunk * 8 : 0);
            cp_async_16(dst, src, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    for (int k = 0; k < kPsStages - 1; ++k)
        prefetch(b_start - k);

    unsigned int* buf0 = rowbuf;
    unsigned int* buf1 = rowbuf + kPsRows * 256;
#define CSM_PS2_STEP(K)                                                                        \
    {                                                                                          \
        const int b = bg + (K);                                                                \
        asm volatile("cp.async.wait_group %0;\n" :: "n"(kPsStages - 2));                       \
        __syncthreads();                                                                       \
        prefetch(b - (kPsStages - 1));                                                         \
        ps2_block<HMAX, (K), FIX, BND>(ring, in_ring, buf0, buf1, job, b, j, in_map && b <= b_top, cells >> 1); \
    }
    for (int bg = b_start + 1 - kPs2Group; bg >= b_lo; bg -= kPs2Group) {
        CSM_PS2_STEP(7) CSM_PS2_STEP(6) CSM_PS2_STEP(5) CSM_PS2_STEP(4)
        CSM_PS2_STEP(3) CSM_PS2_STEP(2) CSM_PS2_STEP(1) CSM_PS2_STEP(0)
    }
#undef CSM_PS2_STEP
    asm volatile("cp.async.wait_group 0;\n" ::);
}

/* Generic sliding win x win maximum with the clamped far edge. */
__global__ void __launch_bounds__(256)
k_sliding_max(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst,
              int rows, int cols, int win)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= cols || r >= rows)
        return;
    const int rc = max(min(r, rows - win), 0);
    const int cc = max(min(c, cols - win), 0);
    unsigned int v = 0;
    for (int dr = 0; dr < win; ++dr)
        for (int dc = 0; dc < win; ++dc)
            v = max(v, ld_cell(src, rows, cols, rc + dr, cc + dc));
    dst[(size_t)r * cols + c] = (uint16_t)v;
}

/* The same sliding maximum, separable and staged in shared memory: a CTA produces a tile of
 * kSmRows x kSmCols outputs from the input rows / columns their (clamped) window origins span;
 * row-wise maxima first, then the column-wise maximum of those: 2 win operations per cell instead
 * of win^2 and every input cell read from memory once per tile. win <= kSmMaxWin. */
constexpr int kSmRows = 16, kSmCols = 128, kSmMaxWin = 32;
constexpr int kSmInRows = kSmRows + kSmMaxWin - 1, kSmInCols = kSmCols + kSmMaxWin - 1;

__global__ void __launch_bounds__(256)
k_sliding_max_tile(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int rows, int cols, int win)
{
    __shared__ uint16_t s_in[kSmInRows][kSmInCols + 1];
    __shared__ uint16_t s_h[kSmInRows][kSmCols];
    const int r0 = blockIdx.y * kSmRows, c0 = blockIdx.x * kSmCols;
    const int r1 = min(r0 + kSmRows, rows) - 1, c1 = min(c0 + kSmCols, cols) - 1;     /* last output row / col */
    /* window origin of an output cell: clamped at the far edge */
    const int rmax = max(rows - win, 0), cmax = max(cols - win, 0);
    const int or0 = min(r0, rmax), or1 = min(r1, rmax);
    const int oc0 = min(c0, cmax), oc1 = min(c1, cmax);
    const int in_rows = or1 - or0 + win, in_cols = oc1 - oc0 + win;
    for (int e = threadIdx.x; e < in_rows * in_cols; e += blockDim.x) {
        const int rr = e / in_cols, cc = e - rr * in_cols;
        s_in[rr][cc] = (uint16_t)ld_cell(src, rows, cols, or0 + rr, oc0 + cc);
    }
    __syncthreads();
    const int n_oc = oc1 - oc0 + 1;
    for (int e = threadIdx.x; e < in_rows * n_oc; e += blockDim.x) {
        const int rr = e / n_oc, cc = e - rr * n_oc;
        unsigned int v = 0;
        for (int k = 0; k < win; ++k) v = max(v, (unsigned int)s_in[rr][cc + k]);
        s_h[rr][cc] = (uint16_t)v;
    }
    __syncthreads();
    const int n_r = r1 - r0 + 1, n_c = c1 - c0 + 1;
    for (int e = threadIdx.x; e < n_r * n_c; e += blockDim.x) {
        const int rr = e / n_c, cc = e - rr * n_c;
        const int orr = min(r0 + rr, rmax) - or0, occ = min(c0 + cc, cmax) - oc0;
        unsigned int v = 0;
        for (int k = 0; k < win; ++k) v = max(v, (unsigned int)s_h[orr + k][occ]);
        dst[(size_t)(r0 + rr) * cols + c0 + cc] = (uint16_t)v;
    }
}

/* ------------------------------------------------------------------------ */
/* Projection                                                                */
/* ------------------------------------------------------------------------ */

constexpr int kProjAngles = 8;       /* candidate angles per CTA */
constexpr int kProjSat = 30000;      /* saturation of projected indices (maps are <= 16384 wide) */

/* What project_beam needs of a query, read ONCE into registers: taken through the DevQuery reference the
 * eight fields were reloaded from global memory for every beam (ncu, round 2: 8 LDG.E.64 per element, the
 * LSU data pipe at 0.73 of its peak in k_project). */
struct ProjConst
{
    double sx, sy, offx, offy, inv_res;
    int guard_hi;                           /* the guard band as a test on high words (load_proj_const) */
    const double2* beam_trig;
    const double* ranges;
};

__device__ __forceinline__ ProjConst load_proj_const(const DevQuery& Q)
{
    ProjConst P;
    P.sx = Q.sx; P.sy = Q.sy; P.offx = Q.offx; P.offy = Q.offy;
    P.inv_res = Q.inv_res;
    /* A fraction g in [0, 1) lies in the band when g < margin or 1 - g < margin (1 - g is exact from 0.5 up).
     * Non-negative doubles order like their bit patterns, so both tests run on the high words alone: integer
     * minimum and compare instead of FP64 compares and selects. That flags a superset (distances within
     * 2^-20 relative of the margin): more flags are always safe. margin <= 0 (exact reruns) flags nothing. */
    P.guard_hi = Q.margin > 0.0 ? __double2hiint(Q.margin) : -1;
    P.beam_trig = Q.beam_trig; P.ranges = Q.ranges;
    return P;
}

/* Hit cell of beam i seen from (Q.sx, Q.sy) at the candidate angle whose (cos, sin) is th:
 * cos/sin(theta + a_i) from the angle-addition formula on the per-beam table, the rest is FP64 in
 * the reference's operation order without contraction (sensor_data.hpp:190-203,
 * grid_map_geometry.cpp:113-122). Raises `flagged` when a coordinate lies within the guard band
 * of a cell boundary. */
__device__ __forceinline__ proj_t project_beam(const ProjConst& Q, const double2 th, int i, int& flagged,
                                               double& rc, double& rs)
{
    const double2 be = __ldg(Q.beam_trig + i);
    const double c = th.x * be.x - th.y * be.y;      /* cos(theta + a) */
    const double s = th.y * be.x + th.x * be.y;      /* sin(theta + a) */
    const double r = __ldg(Q.ranges + i);
    rc = __dmul_rn(r, c);
    rs = __dmul_rn(r, s);
    /* (x - off) * (1 / res) instead of the reference's (x - off) / res: the two
     * differ by < 2 ulp, i.e. floor() can only differ inside the guard band
     * that is flagged anyway (the band is ~1000x wider) */
    const double ux = __dmul_rn(__dsub_rn(__dadd_rn(Q.sx, rc), Q.offx), Q.inv_res);
    const double uy = __dmul_rn(__dsub_rn(__dadd_rn(Q.sy, rs), Q.offy), Q.inv_res);
    /* floor() without the conversion unit (FRND.F64 / F2I.F64 run on the XU pipe at a few lanes per clock: ncu
     * showed it saturated, the FP64 pipe at 0.23): ux + 1.5 * 2^52 rounded DOWN is floor(ux) + 1.5 * 2^52
     * exactly while |ux| < 2^31 (one ulp there is 1), so its low word is floor(ux) as an int and subtracting
     * the constant gives floor(ux) as a double. Anything larger (or NaN) takes the saturating conversion;
     * the clamp on integers follows either way. */
    constexpr double kFloorMagic = 6755399441055744.0;
    const double tx = __dadd_rd(ux, kFloorMagic), ty = __dadd_rd(uy, kFloorMagic);
    int ix = __double2loint(tx), iy = __double2loint(ty);
    double fx = __dsub_rn(tx, kFloorMagic), fy = __dsub_rn(ty, kFloorMagic);
    /* |u| >= 1e9 on either axis, by the high words (an OR of two of them is at least the larger; below 2^16
     * cells it stays below the limit) */
    if ((((unsigned int)__double2hiint(ux) | (unsigned int)__double2hiint(uy)) & 0x7FFFFFFFu) >= 0x41CDCD65u) {
        ix = __double2int_rd(ux); iy = __double2int_rd(uy);
        fx = floor(ux); fy = floor(uy);
    }
    const double gx = ux - fx, gy = uy - fy;
    flagged |= (int)(min(min(__double2hiint(gx), __double2hiint(1.0 - gx)),
                         min(__double2hiint(gy), __double2hiint(1.0 - gy))) <= Q.guard_hi);
    proj_t p;
    p.x = (short)max(min(ix, kProjSat), -kProjSat);
    p.y = (short)max(min(iy, kProjSat), -kProjSat);
    return p;
}

/* Candidate angle t of a query: sensorPose.theta + t * stepTheta (scan_matcher_branch_bound.cpp:
 * 159-165, scan_matcher_correlative.cpp:163-166): one rounded product, one rounded sum */
__device__ __forceinline__ double candidate_theta(const DevQuery& Q, int t)
{
    return (Q.thetas != nullptr) ? Q.thetas[t]
                                 : __dadd_rn(Q.theta0, __dmul_rn((double)(t - Q.tcenter), Q.step_t));
}

/* proj[q][t][i] = (col, row) of beam i seen from (sx, sy, thetas[t]).
 * cos/sin(theta_t + a_i) come from the angle-addition formula on the per-beam
 * table and one sincos per candidate angle (a few ulp from the reference's
 * glibc cos(theta + a), like a direct device sincos would be); the rest is
 * FP64 in the reference's operation order without contraction. Any
 * coordinate closer than `margin` to a cell boundary raises the flag: only
 * there could floor() differ from the reference (DESIGN.md "FP index parity").
 * rcs (optional): r*cos, r*sin per (t, i) for the per-candidate FP path. */
__global__ void __launch_bounds__(256)
k_project(const DevQuery* __restrict__ queries, proj_t* __restrict__ proj,
          double2* __restrict__ rcs, int* __restrict__ qflags)
{
    __shared__ double2 s_theta[kProjAngles];
    const int q = blockIdx.y;
    const DevQuery& Q = queries[q];
    const int t0 = blockIdx.x * kProjAngles;
    if (t0 >= Q.T)
        return;
    const int nt = min(kProjAngles, Q.T - t0);
    if (threadIdx.x < nt) {
        double s, c;
        sincos(candidate_theta(Q, t0 + threadIdx.x), &s, &c);
        s_theta[threadIdx.x] = make_double2(c, s);
    }
    __syncthreads();
    int flagged = 0;
    /* beam-major output (pst_t == 1): 8 consecutive threads take the angles of one beam and write
     * one contiguous 32-byte run; angle-major output: consecutive threads take consecutive beams of
     * one angle. Either way a thread walks its elements with additions only. */
    const bool angle_fastest = Q.pst_t == 1 || Q.pquad != 0;
    const int n = Q.n;
    const int tl_first = angle_fastest ? (int)(threadIdx.x & (kProjAngles - 1)) : 0;
    const int tl_step = angle_fastest ? kProjAngles : 1;
    const int i_first = angle_fastest ? (int)(threadIdx.x / kProjAngles) : (int)threadIdx.x;
    const int i_step = angle_fastest ? (int)(blockDim.x / kProjAngles) : (int)blockDim.x;
    const bool bb = Q.lx > 0;
    const int winx = Q.winx, winy = Q.winy, lx2 = 2 * Q.lx, ly2 = 2 * Q.ly;
    proj_t* __restrict__ out = proj + (size_t)Q.proj_off;
    const bool quad = Q.pquad != 0;
    const ProjConst P = load_proj_const(Q);
    const size_t pst_t = (size_t)Q.pst_t, pst_i = (size_t)Q.pst_i, tp = (size_t)Q.tp;
    double2* __restrict__ rcs_q = rcs != nullptr ? rcs + (size_t)Q.proj_off : nullptr;
    /* chunked layout: the slots past the last beam of the last chunk hold a cell far outside every map,
     * so that the sweep reads its chunks whole, without a test per beam (a child there scores 0) */
    const int n_slots = quad ? ((n + 15) & ~15) : n;
    /* position of (t, i) as proj_index gives it, walked by additions: in the chunk layout i advances by a
     * multiple of 16 (whole chunks), so the place inside the chunk never changes */
    const size_t at_step = quad ? ((size_t)(i_step >> 4) * tp) << 4 : (size_t)i_step * pst_i;
    const unsigned int edge_x = lx2 > 0 ? (unsigned int)(lx2 - 1) : 0u, edge_y = ly2 > 0 ? (unsigned int)(ly2 - 1) : 0u;
    for (int tl = tl_first; tl < nt; tl += tl_step) {
        const double2 th = s_theta[tl];
        const int t = t0 + tl;
        size_t at = quad ? ((((size_t)(i_first >> 4) * tp + (size_t)t) << 4) + (size_t)(((i_first & 3) << 2) | ((i_first >> 2) & 3)))
                         : (size_t)t * pst_t + (size_t)i_first * pst_i;
        int i = i_first;
        auto emit = [&](const proj_t p, int ii, size_t where, double rc, double rs) {
            if (bb) {
                /* branch-and-bound: a node window that straddles row / column 0 makes the coarse bound
                 * inadmissible (a lookup at a negative index reads unknown):
                 * -lx2 < x0 < 0 as one unsigned compare */
                const int x0 = (int)p.x - winx, y0 = (int)p.y - winy;
                flagged |= (int)(((unsigned int)(x0 + lx2 - 1) < edge_x) | ((unsigned int)(y0 + ly2 - 1) < edge_y)) << 2;
            }
            out[where] = p;      /* chunk layout: a warp = 4 beams x 8 angles = one 128-byte line */
            if (rcs_q != nullptr)
                rcs_q[(size_t)t * n + ii] = make_double2(rc, rs);
        };
        /* two beams per trip: the kernel is bound by the latency of one beam's chain of dependent FP64
         * operations (ncu: no pipe above 0.35), two independent chains interleave */
        for (; i + i_step < n; i += 2 * i_step, at += 2 * at_step) {
            double rc0, rs0, rc1, rs1;
            const proj_t p0 = project_beam(P, th, i, flagged, rc0, rs0);
            const proj_t p1 = project_beam(P, th, i + i_step, flagged, rc1, rs1);
            emit(p0, i, at, rc0, rs0);
            emit(p1, i + i_step, at + at_step, rc1, rs1);
        }
        for (; i < n; i += i_step, at += at_step) {
            double rc, rs;
            const proj_t p = project_beam(P, th, i, flagged, rc, rs);
            emit(p, i, at, rc, rs);
        }
        for (; i < n_slots; i += i_step, at += at_step)
            out[at] = proj_t { (short)-32768, (short)-32768 };
    }
    /* bit 0: CSM_FLAG_FP_MARGIN, bit 2: CSM_FLAG_EDGE */
    const int any = (int)__reduce_or_sync(0xffffffffu, (unsigned)flagged);
    if (any != 0 && (threadIdx.x & 31) == 0)
        atomicOr(&qflags[q], any);
}

/* ------------------------------------------------------------------------ */
/* Scoring helpers                                                           */
/* ------------------------------------------------------------------------ */

/* Whole warp: integer score of one candidate (offset ox, oy) on map m */
__device__ __forceinline__ void warp_score(const uint16_t* __restrict__ m, int rows, int cols,
                                           const proj_t* __restrict__ proj, int n, int ox, int oy,
                                           int& sumv, int& nk)
{
    /* proj: the angle's row in beam-fastest layout (stride 1) */
    const int lane = threadIdx.x & 31;
    int s = 0, k = 0;
    for (int i = lane; i < n; i += 32) {
        const proj_t p = proj[i];
        const unsigned int v = ld_cell(m, rows, cols, p.y + oy, p.x + ox);
        s += (int)v;
        k += (v != 0u);
    }
    sumv = warp_sum(s);
    nk = warp_sum(k);
}

/* Threshold comparison with exact resolution inside the guard band.
 * Called by lane 0 only (the exact path is a serial double sum). */
__device__ __forceinline__ bool passes_threshold(long long key, const DevQuery& Q,
                                                 const uint16_t* m, const proj_t* proj,
                                                 int ox, int oy)
{
    const int c = key_vs_threshold(key, Q.kthr);
    if (c != 0)
        return c > 0;
    return exact_normalized_score(m, Q.rows, Q.cols, proj, Q.pst_i, Q.n, ox, oy) > Q.kthr.thr;
}

/* ------------------------------------------------------------------------ */
/* Real-time correlative matcher                                             */
/* ------------------------------------------------------------------------ */

struct RtBlock
{
    long long coarse_key;
    long long fine_key;      /* max over the low_res x low_res block */
    int coarse_nk;
    int fine_ord;            /* (fx - x) * low_res + (fy - y) of the first maximum */
    int fine_tie;            /* another fine candidate shares fine_key */
    int pad;
};

/* One CTA per coarse cell (t, bx, by). Candidate 0 is the coarse score,
 * candidates 1..L*L the fine lattice [x, x+L) x [y, y+L) in the reference's
 * iteration order (x outer, y inner, scan_matcher_correlative.cpp:351-352). */
__global__ void __launch_bounds__(256)
k_rt_blocks(const DevQuery* __restrict__ queries, proj_t* __restrict__ proj_all,
            RtBlock* __restrict__ blocks, int low_res, int nbx, int nby, int* __restrict__ qflags, int proj_given)
{
    extern __shared__ long long s_keys[];      /* L*L fine keys, then the angle's projected indices */
    __shared__ double2 s_theta;
    const DevQuery& Q = queries[0];
    const int b = blockIdx.x;
    const int t = b / (nbx * nby);
    const int rem = b - t * nbx * nby;
    const int bx = rem / nby, by = rem - bx * nby;
    const int x = -Q.winx + bx * low_res;
    const int y = -Q.winy + by * low_res;
    /* every CTA projects its angle itself (ComputeScanIndices, scan_matcher_correlative.cpp:277-297):
     * a few FP64 operations per thread instead of a separate launch; the first CTA of an angle also
     * publishes the indices for k_finalize and raises the guard-band flag */
    proj_t* proj = reinterpret_cast<proj_t*>(s_keys + low_res * low_res);
    if (threadIdx.x == 0) {
        double sn, cs;
        sincos(candidate_theta(Q, t), &sn, &cs);
        s_theta = make_double2(cs, sn);
    }
    __syncthreads();
    {
        const bool publish = rem == 0;
        proj_t* gproj = proj_all + Q.proj_off + (size_t)t * Q.n;
        int flagged = 0;
        if (proj_given) {
            /* exact rerun: the indices were computed on the host with the reference's own libm calls */
            for (int i = threadIdx.x; i < Q.n; i += blockDim.x) proj[i] = gproj[i];
        } else {
        const ProjConst P = load_proj_const(Q);
        const int n = Q.n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double rc, rs;
            const proj_t p = project_beam(P, s_theta, i, flagged, rc, rs);
            proj[i] = p;
            if (publish) gproj[i] = p;
        }
        }
        if (publish && __any_sync(0xffffffffu, flagged) && (threadIdx.x & 31) == 0)
            atomicOr(&qflags[0], 1 /* CSM_FLAG_FP_MARGIN */);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int ncand = low_res * low_res + 1;

    for (int cand = warp; cand < ncand; cand += nwarps) {
        int sumv, nk;
        if (cand == 0) {
            warp_score(Q.coarse, Q.rows, Q.cols, proj, Q.n, x, y, sumv, nk);
            if (lane == 0) {
                blocks[b].coarse_key = make_key(sumv, nk);
                blocks[b].coarse_nk = nk;
            }
        } else {
            const int o = cand - 1;
            const int fx = x + o / low_res, fy = y + o % low_res;
            warp_score(Q.lvl[0], Q.rows, Q.cols, proj, Q.n, fx, fy, sumv, nk);
            if (lane == 0)
                s_keys[o] = make_key(sumv, nk);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long best = -1;
        int ord = 0, tie = 0;
        for (int o = 0; o < ncand - 1; ++o) {
            const long long k = s_keys[o];
            if (k > best) { best = k; ord = o; tie = 0; }
            else if (k == best) tie = 1;
        }
        blocks[b].fine_key = best;
        blocks[b].fine_ord = ord;
        blocks[b].fine_tie = tie;
    }
}

struct BestState
{
    int found, bx, by, bt;     /* bt: angle index 0..T-1 */
    int flags, n_processed, n_ignored, pad;
};

/* Replays scan_matcher_correlative.cpp:161-197 over the coarse cells in the
 * reference's order (t, x, y) on integer keys: exact also when the coarse
 * bound is not admissible. Called by one warp: the lanes
 * stage the block records in shared memory, lane 0 takes the decisions. */
constexpr int kReplayChunk = 256;

__device__ void rt_replay(const DevQuery& Q, const proj_t* __restrict__ proj_all,
                          const RtBlock* __restrict__ blocks, int low_res, int nbx, int nby,
                          RtBlock* s_blocks, BestState& out)
{
    const int lane = threadIdx.x & 31;
    bool have = false;          /* scoreMax is a fine score (else: the threshold) */
    long long cur = 0;
    int bestx = -Q.winx, besty = -Q.winy, bestt = 0, flags = 0;
    int processed = 0, ignored = 0;
    const int nb = Q.T * nbx * nby;
    int t = 0, bx = 0, by = 0;       /* block b = (t, bx, by), advanced without divisions */
    int best_ord = 0, best_x0 = 0, best_y0 = 0;
    for (int base = 0; base < nb; base += kReplayChunk) {
        const int cnt = min(kReplayChunk, nb - base);
        __syncwarp();
        for (int i = lane; i < cnt * (int)(sizeof(RtBlock) / 16); i += 32)
            reinterpret_cast<uint4*>(s_blocks)[i] = reinterpret_cast<const uint4*>(blocks + base)[i];
        __syncwarp();
        if (lane != 0)
            continue;
        for (int k = 0; k < cnt; ++k) {
            const long long coarse_key = s_blocks[k].coarse_key;
            const long long fine_key = s_blocks[k].fine_key;
            const int coarse_nk = s_blocks[k].coarse_nk;
            const int x = -Q.winx + bx * low_res, y = -Q.winy + by * low_res;
            const int tt = t;
            if (++by == nby) { by = 0; if (++bx == nbx) { bx = 0; ++t; } }
            bool coarse_ok;
            if (have) {
                coarse_ok = coarse_key > cur;
                if (coarse_key == cur) flags |= 2;   /* equal keys: doubles could round either way */
            } else {
                const int c = key_vs_threshold(coarse_key, Q.kthr);
                coarse_ok = c > 0;
                if (c == 0)
                    coarse_ok = exact_normalized_score(Q.coarse, Q.rows, Q.cols,
                                                       proj_all + Q.proj_off + (size_t)tt * Q.n, Q.pst_i, Q.n,
                                                       x, y) > Q.kthr.thr;
            }
            if (!coarse_ok || coarse_nk <= Q.nk_cut) {
                ++ignored;
                continue;
            }
            ++processed;
            bool fine_ok;
            if (have) {
                fine_ok = fine_key > cur;
            } else {
                const int c = key_vs_threshold(fine_key, Q.kthr);
                fine_ok = c > 0;
                if (c == 0) {
                    const int ord = s_blocks[k].fine_ord;
                    fine_ok = exact_normalized_score(Q.lvl[0], Q.rows, Q.cols,
                                                     proj_all + Q.proj_off + (size_t)tt * Q.n, Q.pst_i, Q.n,
                                                     x + ord / low_res, y + ord % low_res) > Q.kthr.thr;
                }
            }
            if (fine_ok) {
                have = true;
                cur = fine_key;
                best_ord = s_blocks[k].fine_ord; best_x0 = x; best_y0 = y; bestt = tt;
                if (s_blocks[k].fine_tie) flags |= 2;
            }
        }
    }
    if (have) {
        bestx = best_x0 + best_ord / low_res;
        besty = best_y0 + best_ord % low_res;
    }
    out.found = have ? 1 : 0;
    out.bx = bestx; out.by = besty;
    out.bt = have ? bestt : 0;    /* reference initial bestWinTheta = -winTheta = index 0 */
    out.flags = flags; out.n_processed = processed; out.n_ignored = ignored; out.pad = 0;
}

/* The same replay as a warp scan, 32 blocks per step, valid whenever the
 * sequential decisions reduce to "strict prefix maxima of the fine keys":
 * every block that could be accepted has coarse key >= fine key (the bound is
 * admissible there) and no key falls inside the threshold guard band. Returns
 * false (nothing decided) otherwise; the caller then replays sequentially. */
__device__ __forceinline__ long long shfl_ll(long long v, int src)
{
    return __shfl_sync(0xffffffffu, v, src);
}

__device__ bool rt_replay_scan(const DevQuery& Q, const RtBlock* __restrict__ blocks, int low_res,
                               int nbx, int nby, BestState& out)
{
    const int lane = threadIdx.x & 31;
    const int nb = Q.T * nbx * nby;
    long long cur = -1;                  /* best accepted fine key so far (-1: none yet) */
    int best_b = -1, best_ord = 0, flags = 0, processed = 0;
    for (int base = 0; base < nb; base += 32) {
        const int b = base + lane;
        const bool valid = b < nb;
        RtBlock B;
        if (valid) {
            const uint4 lo = reinterpret_cast<const uint4*>(blocks + b)[0];
            const uint4 hi = reinterpret_cast<const uint4*>(blocks + b)[1];
            B.coarse_key = (long long)(((unsigned long long)lo.y << 32) | lo.x);
            B.fine_key = (long long)(((unsigned long long)lo.w << 32) | lo.z);
            B.coarse_nk = (int)hi.x; B.fine_ord = (int)hi.y; B.fine_tie = (int)hi.z; B.pad = 0;
        } else {
            B.coarse_key = -1; B.fine_key = -1; B.coarse_nk = 0; B.fine_ord = 0; B.fine_tie = 0; B.pad = 0;
        }
        const bool nkok = valid && B.coarse_nk > Q.nk_cut;
        const int cc = key_vs_threshold(B.coarse_key, Q.kthr), cf = key_vs_threshold(B.fine_key, Q.kthr);
        const bool odd = valid && (cc == 0 || cf == 0 || (nkok && B.coarse_key < B.fine_key));
        if (__any_sync(0xffffffffu, odd))
            return false;
        const bool elig = nkok && cf > 0;
        const long long x = elig ? B.fine_key : -1;
        /* inclusive prefix maximum over the lanes */
        long long incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o && up > incl) incl = up;
        }
        long long excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = -1;
        const long long cur_b = excl > cur ? excl : cur;        /* scoreMax when block b is reached */
        const bool have_b = cur_b >= 0;
        const bool coarse_ok = have_b ? (B.coarse_key > cur_b) : (cc > 0);
        const bool proc = nkok && coarse_ok;
        const bool accepted = elig && (!have_b || B.fine_key > cur_b);
        const bool tie = (valid && have_b && B.coarse_key == cur_b) || (accepted && B.fine_tie != 0);
        processed += __popc(__ballot_sync(0xffffffffu, proc));
        if (__any_sync(0xffffffffu, tie)) flags |= 2;
        const long long top = shfl_ll(incl, 31);
        if (top > cur) {
            /* the last accepted block of this step is the first one that reaches the step's maximum */
            const unsigned int at = __ballot_sync(0xffffffffu, x == top);
            const int src = __ffs(at) - 1;
            best_b = base + src;
            best_ord = __shfl_sync(0xffffffffu, B.fine_ord, src);
            cur = top;
        }
    }
    const bool have = best_b >= 0;
    int bestx = -Q.winx, besty = -Q.winy, bestt = 0;
    if (have) {
        const int t = best_b / (nbx * nby);
        const int rem = best_b - t * nbx * nby;
        const int bx = rem / nby, by = rem - bx * nby;
        bestx = -Q.winx + bx * low_res + best_ord / low_res;
        besty = -Q.winy + by * low_res + best_ord % low_res;
        bestt = t;
    }
    out.found = have ? 1 : 0;
    out.bx = bestx; out.by = besty; out.bt = bestt;
    out.flags = flags; out.n_processed = processed; out.n_ignored = nb - processed; out.pad = 0;
    return true;
}

/* ------------------------------------------------------------------------ */
/* Branch and bound                                                          */
/* ------------------------------------------------------------------------ */

/* Level-synchronous frontier expansion.
 *
 * list(h) holds the nodes (q, t, x, y) of height h that PASSED when they were
 * scored, i.e. that the reference would expand when it pops them
 * (scan_matcher_branch_bound.cpp:191-198: drop iff score <= scoreMax or
 * knownRate <= threshold). k_bb_roots scores the root candidates (:179-182)
 * into list(hmax); k_bb_expand(h) takes every node of list(h), scores its four
 * children (:226-229) on the level h-1 map and appends the passing ones to
 * list(h-1); passing children of height 0 are leaves and raise the query's
 * incumbent with atomicMax.
 *
 * Lanes: a warp takes 8 consecutive nodes; lane = part * 8 + slot, `slot` the
 * node, `part` the quarter of the beams (i = part, part + 4, ...) the lane
 * sums. Per beam a lane loads the projected index once and gathers the four
 * children's cells. Lists keep runs of adjacent angles t of the same (x, y)
 * (roots are generated t-fastest, children are appended per child type in lane
 * order), adjacent angles and adjacent beams hit neighbouring cells, and the
 * projection is stored beam-major (proj[i][t]): a warp's index loads are 4
 * short contiguous runs and its 32 gathers of one child fall on a compact 2-D
 * patch of the map. */
struct BbWork
{
    unsigned long long* list[2];    /* node lists, ping-pong by height parity */
    unsigned int*       counts;     /* [kMaxLevels]: nodes of list(h) */
    unsigned long long* incumbent;  /* per query: packed (key, ordfield) */
    int*                stats;      /* per query: processed, ignored */
    int*                overflow;   /* set when a list is full */
    long long*          rootkey;    /* per root candidate: its key if it passed, else -1 (feeds the dive) */
    unsigned long long* tiekey;     /* per query: largest key two different leaves were seen to share (0: none) */
    unsigned long long* probe;      /* per query, kDiveStarts words: the children with the largest bound of the launch
                                       (start nodes of k_bbg_dive); null: off */
    unsigned int        probe_heights; /* bit hc: the launch that creates the children of height hc records them */
    unsigned int        capacity;
    int                 top;        /* height of the root candidates (the reference's node_height_max) */
    int                 split_shift; /* lanes per node: largest split with count * split * 2 <= lanes << split_shift */
};

/* list(h) lives in list[(h & 1) ^ 1]; the root candidates in list[top & 1] */
__device__ __forceinline__ unsigned long long* bb_list(const BbWork& W, int h) { return W.list[(h & 1) ^ 1]; }

__device__ __forceinline__ unsigned long long leaf_ordfield(const DevQuery& Q, int t, int xi, int yi)
{
    /* smaller ordinal (t, x, y order) wins ties -> larger field */
    const unsigned long long ord =
        ((unsigned long long)t * (unsigned)Q.lx + (unsigned)xi) * (unsigned)Q.ly + (unsigned)yi;
    return (kOrdMask - 1ull) - ord;     /* all-ones is reserved for "no leaf yet" */
}

/* A passing leaf raises its query's incumbent. When the word it replaces (or fails to replace) carries
 * the same key from another leaf, that key is remembered: k_finalize raises CSM_FLAG_KEY_TIE when the
 * winning key is a shared one (the reference's winner among equal double scores depends on its heap
 * order; here the smallest (t, x, y) ordinal wins). */
__device__ __forceinline__ void bb_raise_incumbent(const BbWork& W, int q, long long key, unsigned long long ordfield)
{
    const unsigned long long word = pack_best(key, ordfield);
    const unsigned long long old = atomicMax(&W.incumbent[q], word);
    if (old != word && (old >> kOrdBits) == (unsigned long long)key && (old & kOrdMask) != kOrdMask)
        atomicMax(&W.tiekey[q], (unsigned long long)key);
}

/* Root candidates of every query: (x, y) stepping by 2^top from -win, all angles
 * (scan_matcher_branch_bound.cpp:179-182), angle fastest. With `unscored` they go straight into
 * list(top) as if all of them had passed (which internal nodes get expanded never changes the
 * result): the first expand launch then scores their children, four per index load, instead of
 * k_bb_roots scoring the roots one gather per index load first. */
__global__ void __launch_bounds__(256)
k_bb_init(const DevQuery* __restrict__ queries, const unsigned int* __restrict__ root_off,
          int nq, BbWork W, int unscored)
{
    const int q = blockIdx.y;
    const DevQuery& Q = queries[q];
    const int nroots = Q.T * Q.nrx * Q.nry;
    unsigned long long* out = (unscored ? bb_list(W, W.top) : W.list[W.top & 1]) + root_off[q];
    if (unscored && q == 0 && blockIdx.x == 0 && threadIdx.x == 0)
        W.counts[W.top] = root_off[nq];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nroots; e += gridDim.x * blockDim.x) {
        const int t = e % Q.T;
        const int cell = e / Q.T;
        const int rx = cell / Q.nry, ry = cell - rx * Q.nry;
        out[e] = pack_node(q, t, rx << W.top, ry << W.top);
    }
}

constexpr int kBbSplit = 4;                 /* lanes cooperating on one node */
#ifndef CSM_BB_UNROLL
#define CSM_BB_UNROLL 2
#endif
constexpr int kBbUnroll = CSM_BB_UNROLL;    /* beams (x 4 children) a lane keeps in flight */
constexpr int kBbNodesPerWarp = 32 / kBbSplit;

/* The decision the reference takes when it pops a node, on integer keys */
__device__ __forceinline__ bool bb_passes(const DevQuery& Q, const proj_t* __restrict__ proj_all,
                                          const BbWork& W, int q, int t, int xi, int yi, int h,
                                          int s, int k, long long& key)
{
    key = make_key(s, k);
    const unsigned long long inc = *(volatile unsigned long long*)&W.incumbent[q];
    bool pass = k > Q.nk_cut && pack_best(key, kOrdMask) > inc;
    if (pass) {
        const int c = key_vs_threshold(key, Q.kthr);
        if (c < 0) pass = false;
        else if (c == 0) {
            pass = exact_normalized_score_q(Q, Q.lvl[h], proj_all + Q.proj_off, t,
                                            xi - Q.winx, yi - Q.winy) > Q.kthr.thr;
        }
    }
    return pass;
}

/* processed / ignored counters, one atomic per (warp, query, outcome) */
__device__ __forceinline__ void bb_count(const BbWork& W, bool valid, int q, bool pass)
{
    const int lane = threadIdx.x & 31;
    const int tag = valid ? (2 * q + (pass ? 0 : 1)) : -1;
    const unsigned int peers = __match_any_sync(0xffffffffu, tag);
    if (tag >= 0 && lane == __ffs(peers) - 1)
        atomicAdd(&W.stats[tag], __popc(peers));
}

/* Score the root candidates on the level `top` map; the passing ones form
 * list(top). Lanes per node as in k_bb_expand. */
__global__ void __launch_bounds__(256)
k_bb_roots(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all,
           BbWork W, unsigned int count)
{
    const int h = W.top;
    const int lane = threadIdx.x & 31;
    const unsigned int total_lanes = gridDim.x * blockDim.x;
    int split = kBbSplit;
    while (split < 32 && (((unsigned long long)count * (unsigned)(split * 2)) << max(-W.split_shift, 0)) <= ((unsigned long long)total_lanes << max(W.split_shift, 0))) split *= 2;
    const int npw = 32 / split;
    const int slot = lane & (npw - 1);
    const int part = lane / npw;
    const unsigned long long* __restrict__ in = W.list[h & 1];
    unsigned long long* __restrict__ out = bb_list(W, h);
    const unsigned int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned int nwarps = total_lanes >> 5;
    for (unsigned int base = warp_global * npw; base < count; base += nwarps * npw) {
        const unsigned int idx = base + slot;
        const bool valid = idx < count;
        int q = 0, t = 0, xi = 0, yi = 0;
        int s = 0, k = 0;
        if (valid) {
            unpack_node(in[idx], q, t, xi, yi);
            const DevQuery& Q = queries[q];
            const uint16_t* __restrict__ m = Q.lvl[h];
            const int rows = Q.rows, cols = Q.cols, n = Q.n;
            const int ox = xi - Q.winx, oy = yi - Q.winy;
            const size_t ps = (size_t)Q.pst_i;
            const proj_t* __restrict__ pp = proj_all + Q.proj_off + (size_t)t * Q.pst_t + (size_t)part * ps;
            const size_t step = ps * (size_t)split;
            const int mine = (n - part + split - 1) / split;     /* beams of this lane */
            int i = 0;
            for (; i + 8 <= mine; i += 8) {
                proj_t p[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) p[u] = pp[(size_t)(i + u) * step];
                unsigned int v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_cell_nb(m, rows, cols, p[u].y + oy, p[u].x + ox);
#pragma unroll
                for (int u = 0; u < 8; ++u) { s += (int)v[u]; k += (v[u] != 0u); }
            }
            for (; i < mine; ++i) {
                const proj_t p = pp[(size_t)i * step];
                const unsigned int v = ld_cell_nb(m, rows, cols, p.y + oy, p.x + ox);
                s += (int)v; k += (v != 0u);
            }
        }
        for (int o = npw; o < 32; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            k += __shfl_xor_sync(0xffffffffu, k, o);
        }
        bool pass = false;
        long long key = 0;
        if (valid && part == 0) {
            const DevQuery& Q = queries[q];
            pass = bb_passes(Q, proj_all, W, q, t, xi, yi, h, s, k, key);
            if (pass && h == 0)
                bb_raise_incumbent(W, q, key, leaf_ordfield(Q, t, xi, yi));
            if (W.rootkey != nullptr)
                W.rootkey[idx] = pass ? key : -1ll;
        }
        bb_count(W, valid && part == 0, q, pass);
        if (h > 0) {
            const unsigned int ballot = __ballot_sync(0xffffffffu, pass);
            if (ballot != 0u) {
                unsigned int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&W.counts[h], (unsigned int)__popc(ballot));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (pass) {
                    const unsigned int dst = slot0 + __popc(ballot & ((1u << lane) - 1u));
                    if (dst < W.capacity) out[dst] = pack_node(q, t, xi, yi);
                    else *W.overflow = 1;
                }
            }
        }
    }
}

/* Four cells (r, c), (r, c + w), (r + w, c), (r + w, c + w) of the children's level, w = 2^HC,
 * branch-free so that a lane keeps all its loads in flight: out-of-map cells read cell 0 and
 * are masked (the reference's ValueOr -> 0, grid_map.cpp:389-392). */
template <int HC>
__device__ __forceinline__ void ld_children(const uint16_t* __restrict__ m, int rows, int cols,
                                            int r, int c, unsigned int (&v)[4])
{
    constexpr int w = 1 << HC;
    const int r1 = r + w, c1 = c + w;
    const bool rk0 = (unsigned)r < (unsigned)rows, rk1 = (unsigned)r1 < (unsigned)rows;
    const bool ck0 = (unsigned)c < (unsigned)cols, ck1 = (unsigned)c1 < (unsigned)cols;
    const unsigned int R0 = (unsigned)r * (unsigned)cols, R1 = R0 + (unsigned)(w * cols);
    const bool k00 = rk0 && ck0, k01 = rk0 && ck1, k10 = rk1 && ck0, k11 = rk1 && ck1;
    const unsigned int x00 = __ldg(m + (k00 ? R0 + (unsigned)c : 0u)), x01 = __ldg(m + (k01 ? R0 + (unsigned)c1 : 0u));
    const unsigned int x10 = __ldg(m + (k10 ? R1 + (unsigned)c : 0u)), x11 = __ldg(m + (k11 ? R1 + (unsigned)c1 : 0u));
    v[0] = k00 ? x00 : 0u; v[1] = k01 ? x01 : 0u; v[2] = k10 ? x10 : 0u; v[3] = k11 ? x11 : 0u;
}

/* Expand list(HC + 1): score the four children of every node on the level HC map.
 *
 * `split` lanes share a node (a power of two, 4..32): lane = part * npw + slot
 * with npw = 32 / split nodes per warp. The kernel picks the largest split
 * that still gives every node of the list its lanes in one wave of the grid,
 * so short lists (the fine levels, small calls) are latency-bound on a few
 * beams per lane instead of 90.
 *
 * Children of height >= 1 are internal nodes: whether they pass only decides
 * how much work follows, never the result (a leaf that passes implies that all
 * its ancestors pass), so they are tested on an upper bound of
 * their key that needs no known-cell count: key <= 998 sum + 64536 n. Leaves
 * (HC == 0) are scored exactly. */
#ifndef CSM_BB_MINB
#define CSM_BB_MINB 4
#endif
template <int HC>
__global__ void __launch_bounds__(256, CSM_BB_MINB)
k_bb_expand(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all, BbWork W)
{
    constexpr int h = HC + 1;
    constexpr int w = 1 << HC;               /* spacing of the children */
    constexpr bool kLeaf = HC == 0;
#ifndef CSM_BB_COUNT_ALL
#define CSM_BB_COUNT_ALL 0
#endif
    constexpr bool kCount = kLeaf || CSM_BB_COUNT_ALL;     /* known-cell counts above the leaves too */
    const int lane = threadIdx.x & 31;
    const unsigned int count = min(W.counts[h], W.capacity);
    const unsigned int total_lanes = gridDim.x * blockDim.x;
    int split = kBbSplit;
    while (split < 32 && (((unsigned long long)count * (unsigned)(split * 2)) << max(-W.split_shift, 0)) <= ((unsigned long long)total_lanes << max(W.split_shift, 0))) split *= 2;
    const int npw = 32 / split;              /* nodes per warp */
    const int slot = lane & (npw - 1);
    const int part = lane / npw;
    const unsigned long long* __restrict__ in = bb_list(W, h);
    unsigned long long* __restrict__ out = bb_list(W, HC);
    const unsigned int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned int nwarps = total_lanes >> 5;
    for (unsigned int base = warp_global * npw; base < count; base += nwarps * npw) {
        const unsigned int idx = base + slot;
        const bool valid = idx < count;
        int q = 0, t = 0, xi = 0, yi = 0;
        unsigned int s0 = 0, s1 = 0, s2 = 0, s3 = 0, k0 = 0, k1 = 0, k2 = 0, k3 = 0;
        if (valid) {
            unpack_node(in[idx], q, t, xi, yi);
            const DevQuery& Q = queries[q];
            const uint16_t* __restrict__ m = Q.lvl[HC];
            const int rows = Q.rows, cols = Q.cols, n = Q.n;
            const int ox = xi - Q.winx, oy = yi - Q.winy;
            const unsigned int ps = (unsigned int)Q.pst_i;
            const proj_t* __restrict__ pp = proj_all + Q.proj_off + (size_t)t * Q.pst_t + (size_t)part * ps;
            const unsigned int step = ps * (unsigned int)split;
            const int mine = (n - part + split - 1) / split;     /* beams of this lane */
            int i = 0;
            /* software pipeline: the projected indices of the next group are requested before the
             * gathers of the current one are consumed, so a group costs one memory round trip, not
             * two dependent ones (the sweep is bound by load latency: long-scoreboard stalls) */
            proj_t pn[kBbUnroll];
            if (kBbUnroll <= mine) {
#pragma unroll
                for (int u = 0; u < kBbUnroll; ++u) pn[u] = pp[(unsigned int)u * step];
            }
            for (; i + kBbUnroll <= mine; i += kBbUnroll) {
                proj_t p[kBbUnroll];
#pragma unroll
                for (int u = 0; u < kBbUnroll; ++u) p[u] = pn[u];
                unsigned int v[kBbUnroll][4];
#pragma unroll
                for (int u = 0; u < kBbUnroll; ++u)
                    ld_children<HC>(m, rows, cols, p[u].y + oy, p[u].x + ox, v[u]);
                if (i + 2 * kBbUnroll <= mine) {
#pragma unroll
                    for (int u = 0; u < kBbUnroll; ++u) pn[u] = pp[(unsigned int)(i + kBbUnroll + u) * step];
                }
#pragma unroll
                for (int u = 0; u < kBbUnroll; ++u) {
                    s0 += v[u][0]; s1 += v[u][1]; s2 += v[u][2]; s3 += v[u][3];
                    if (kCount) {
                        k0 += (v[u][0] != 0u); k1 += (v[u][1] != 0u);
                        k2 += (v[u][2] != 0u); k3 += (v[u][3] != 0u);
                    }
                }
            }
            for (; i < mine; ++i) {
                const proj_t p = pp[(unsigned int)i * step];
                unsigned int v[4];
                ld_children<HC>(m, rows, cols, p.y + oy, p.x + ox, v);
                s0 += v[0]; s1 += v[1]; s2 += v[2]; s3 += v[3];
                if (kCount) { k0 += (v[0] != 0u); k1 += (v[1] != 0u); k2 += (v[2] != 0u); k3 += (v[3] != 0u); }
            }
        }
        /* Sum over the parts (sums < 2^32; the four known counts < 2^16 share one word) */
        unsigned long long a = ((unsigned long long)s0 << 32) | s1;
        unsigned long long b = ((unsigned long long)s2 << 32) | s3;
        unsigned long long kk = ((unsigned long long)k0 << 48) | ((unsigned long long)k1 << 32) |
                                ((unsigned long long)k2 << 16) | (unsigned long long)k3;
        for (int o = npw; o < 32; o <<= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            if (kCount) kk += __shfl_xor_sync(0xffffffffu, kk, o);
        }
        /* lane (slot, part < 4) now decides child `part` of node `slot`: (x + (part & 1) w, y + (part >> 1) w) */
        const bool decides = valid && part < 4;
        const unsigned long long sp = (part & 2) ? b : a;
        const int s = (int)((part & 1) ? (unsigned)sp : (unsigned)(sp >> 32));
        const int cx = xi + (part & 1) * w, cy = yi + ((part >> 1) & 1) * w;
        bool pass = false;
        if (decides) {
            const DevQuery& Q = queries[q];
            if (kLeaf) {
                const int k = (int)((kk >> (16 * (3 - (part & 3)))) & 0xffffull);
                long long key;
                pass = bb_passes(Q, proj_all, W, q, t, cx, cy, 0, s, k, key);
                if (pass)
                    bb_raise_incumbent(W, q, key, leaf_ordfield(Q, t, cx, cy));
            } else {
                /* upper bound of the key: every beam counted as known (or, with counts, the cells of
                 * the coarse level that are known: at least as many as at any leaf below) */
                const int kub = kCount ? (int)((kk >> (16 * (3 - (part & 3)))) & 0xffffull) : Q.n;
                const long long key_ub = make_key(s, kub);
                const unsigned long long inc = *(volatile unsigned long long*)&W.incumbent[q];
                pass = pack_best(key_ub, kOrdMask) > inc && key_ub > Q.kthr.fail_max && kub > Q.nk_cut;
            }
        }
        bb_count(W, decides, q, pass);
        if (!kLeaf) {
            const unsigned int ballot = __ballot_sync(0xffffffffu, pass);
            if (ballot != 0u) {
                unsigned int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&W.counts[HC], (unsigned int)__popc(ballot));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (pass) {
                    /* lane order = child type major, node minor: runs of adjacent angles stay together */
                    const unsigned int dst = slot0 + __popc(ballot & ((1u << lane) - 1u));
                    if (dst < W.capacity) out[dst] = pack_node(q, t, cx, cy);
                    else *W.overflow = 1;
                }
            }
        }
    }
}

/* ---- the sweep over bound levels ---------------------------------------------------
 * Same frontier expansion as k_bb_expand, reading the representation built for it (csm_bounds.cuh) and
 * working on GROUPS: a list entry is (query, x, y, group of 8 adjacent angles, 8-bit mask of the angles
 * still alive). Along the angle axis the frontier is made of runs (a node's bound changes slowly with
 * the angle: measured 93 % / 87 % / 78 % / 67 % / 54 % of the angles of a listed group are alive at
 * heights 5..1 of the loop-detection batch), and a warp that takes one group has its 32 lanes = 8
 * adjacent angles x 4 adjacent beams on ONE (x, y): their hit cells lie on an arc of a few cells, so
 * a gather touches one to three 128-byte tiles of the bound level, and the index load is exactly one
 * line of the quad layout [n / 4][tp][4]. (With one list entry per node a warp's 8 entries came from 2
 * to 5 different (query, x, y) cells.) Short lists give every group several warps, each on a slice of the
 * beams, summed through shared memory.
 *
 * Children of height HC >= 1 are tested on 257 * sum of the u8 bound level HC (zero padded: one clamp
 * per axis instead of a bounds test per child); leaves (HC == 0) are scored exactly on the u16 level-0
 * grid. */
#ifndef CSM_BBG_UNROLL
#define CSM_BBG_UNROLL 2
#endif
constexpr int kBbgUnroll = CSM_BBG_UNROLL;
constexpr int kBbgWarps = 8;                 /* warps per CTA */

__device__ __forceinline__ unsigned long long pack_group(int q, int tg, unsigned int mask, int xi, int yi)
{
    return ((unsigned long long)(unsigned)q << 48) | ((unsigned long long)(unsigned)tg << 40) |
           ((unsigned long long)(mask & 0xffu) << 32) | ((unsigned long long)(unsigned)xi << 16) |
           (unsigned long long)(unsigned)yi;
}

/* Root groups of every query: (x, y) stepping by 2^top from -win, every group of 8 angles, all alive:
 * the first expand launch scores their children (unscored roots, see k_bb_init). */
__global__ void __launch_bounds__(256)
k_bbg_init(const DevQuery* __restrict__ queries, const unsigned int* __restrict__ root_off, int nq, BbWork W)
{
    const int q = blockIdx.y;
    const DevQuery& Q = queries[q];
    const int G = (Q.T + 7) >> 3;
    const int nitems = G * Q.nrx * Q.nry;
    unsigned long long* out = bb_list(W, W.top) + root_off[q];
    if (q == 0 && blockIdx.x == 0 && threadIdx.x == 0)
        W.counts[W.top] = root_off[nq];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nitems; e += gridDim.x * blockDim.x) {
        const int tg = e % G;
        const int cell = e / G;
        const int rx = cell / Q.nry, ry = cell - rx * Q.nry;
        const int left = Q.T - 8 * tg;
        const unsigned int mask = left >= 8 ? 0xffu : ((1u << left) - 1u);
        out[e] = pack_group(q, tg, mask, rx << W.top, ry << W.top);
    }
}

template <int HC>
__device__ __forceinline__ void ld_children_b(const unsigned char* __restrict__ bm, unsigned int tpr, unsigned int row_w,
                                              int rows, int cols, int r, int c, unsigned int (&v)[4])
{
    constexpr int w = 1 << HC;
    constexpr int PR = bl_pad_r(HC), PC = bl_pad_c(HC);
    /* children outside the map read the zero padding: clamp the base cell into [-(w + 1), extent] */
    const unsigned int rp = (unsigned int)(min(max(r, -(w + 1)), rows) + PR);
    const unsigned int cp = (unsigned int)(min(max(c, -(w + 1)), cols) + PC);
    /* bl_cell split into its row and column parts */
    const unsigned int ro0 = (((rp >> 2) * tpr) << 7) + ((rp & 3u) << 1);
    unsigned int ro1;
    if (w >= kBlTileR) ro1 = ro0 + row_w;          /* row_w = (w / tile rows) * tpr * 128 */
    else { const unsigned int rq = rp + w; ro1 = (((rq >> 2) * tpr) << 7) + ((rq & 3u) << 1); }
    /* column part of bl_cell: c = 32 a + 2 b + p sits at 128 a + 8 b + p = 4 c - 3 p, so a step of an even w
     * columns is 4 w bytes wherever the tile boundaries fall: the second child of a row is an immediate offset */
    const unsigned int co0 = (cp << 2) - 3u * (cp & 1u);
    unsigned int dco = 4 * w;
    if ((w & 1) != 0) {            /* height 0 is never a bound level (the leaves read the u16 map); kept well-formed */
        const unsigned int cq = cp + w;
        dco = ((cq << 2) - 3u * (cq & 1u)) - co0;
    }
    const unsigned char* __restrict__ p0 = bm + (ro0 + co0);
    const unsigned char* __restrict__ p1 = bm + (ro1 + co0);
    v[0] = __ldg(p0); v[1] = __ldg(p0 + dco);
    v[2] = __ldg(p1); v[3] = __ldg(p1 + dco);
}

/* warps per group: the split that minimises (rounds of the grid) x (steps per warp), with a few steps
 * charged for the exchange through shared memory when a group is shared */
__device__ __forceinline__ int bbg_warps_per_group(unsigned int count, unsigned int ctas, int nchunks)
{
    int best = 1;
    unsigned int best_cost = 0xffffffffu;
#pragma unroll
    for (int wpi = 1; wpi <= kBbgWarps; wpi *= 2) {
        const unsigned int per_round = ctas * (unsigned int)(kBbgWarps / wpi);
        const unsigned int rounds = (count + per_round - 1) / per_round;
        const unsigned int cost = rounds * (unsigned int)((nchunks + wpi - 1) / wpi + (wpi > 1 ? 1 : 0));
        if (cost < best_cost) { best_cost = cost; best = wpi; }
    }
    return best;
}

/* ---- incumbent dives of the group sweep -----------------------------------------------------------
 * The level-synchronous sweep meets leaves only at its last launch, so on its own it prunes against the
 * score threshold alone. Any real leaf is a valid incumbent (the result, maximum key and smallest ordinal,
 * does not depend on it: a node is kept while its bound is >= the incumbent's key, so the best leaf, its
 * ancestors and every leaf that ties with it survive). After the launches that create the children of the
 * upper heights, one CTA per query descends greedily from the children with the largest bounds -- the best of
 * every fourth group of angles, so that the starts differ in angle -- keeping the kDiveBeam best children at
 * every height, down to leaves that are scored exactly; the best one becomes the query's incumbent before
 * the next launch decides anything. On loop-detection batches the launches below then score a quarter of the
 * groups they would otherwise (true positives, where the threshold prunes least, dominate them). */
constexpr int kDiveStarts = 4;               /* start nodes per query (a power of two) */
constexpr int kDiveWidth = 4;                /* nodes kept per height */
constexpr int kProbePosBits = 14;            /* leaf lattice extents up to 16383 cells, angles up to 4095 */

__device__ __forceinline__ unsigned long long pack_probe(unsigned int sv, int t, int cx, int cy)
{
    return ((unsigned long long)min(sv, 0xffffffu) << 40) | ((unsigned long long)(unsigned)t << (2 * kProbePosBits)) |
           ((unsigned long long)(unsigned)cx << kProbePosBits) | (unsigned long long)(unsigned)cy;
}

/* The four children of up to kDiveWidth nodes on level HH: 64 lanes (two warps) per node, a lane takes the
 * beams lane, lane + 64, ...; the index loads of kDiveUnroll beams are issued together, then their 4 x
 * kDiveUnroll cell loads (the dive is a chain of dependent launches-within-a-launch: its time is the latency
 * of these rounds, not their work). s_sum[node][child], s_known likewise (leaves). */
constexpr int kDiveUnroll = 6;               /* 6 x 64 lanes: one round for scans of up to 384 beams */

template <int HH>
__device__ __forceinline__ void dive_score(const DevQuery& Q, const proj_t* __restrict__ pq,
                                           const int (*s_node)[3], int n_nodes,
                                           unsigned int (*s_sum)[4], unsigned int (*s_known)[4])
{
    constexpr int w = 1 << HH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int node = warp >> 1, half = warp & 1;
    unsigned int s[4] = { 0u, 0u, 0u, 0u }, k[4] = { 0u, 0u, 0u, 0u };
    if (node < n_nodes) {
        const int t = s_node[node][0];
        const int ox = s_node[node][1] - Q.winx, oy = s_node[node][2] - Q.winy;
        const unsigned int tpr = (unsigned int)Q.bl_tpr[HH];
        const unsigned int row_w = (unsigned int)(w / kBlTileR) * (tpr << 7);
        const uint16_t* __restrict__ m = Q.lvl[0];
        const unsigned char* __restrict__ bm = Q.bl[HH];
        const int rows = Q.rows, cols = Q.cols, n = Q.n;
        for (int i0 = half * 32 + lane; i0 < n; i0 += 64 * kDiveUnroll) {
            proj_t p[kDiveUnroll];
#pragma unroll
            for (int u = 0; u < kDiveUnroll; ++u) {
                const int i = i0 + 64 * u;
                /* beams past the end read beam i0 again and are masked below */
                p[u] = pq[proj_index(Q, t, i < n ? i : i0)];
            }
            unsigned int v[kDiveUnroll][4];
#pragma unroll
            for (int u = 0; u < kDiveUnroll; ++u) {
                if (HH == 0) ld_children<0>(m, rows, cols, p[u].y + oy, p[u].x + ox, v[u]);
                else ld_children_b<HH>(bm, tpr, row_w, rows, cols, p[u].y + oy, p[u].x + ox, v[u]);
            }
#pragma unroll
            for (int u = 0; u < kDiveUnroll; ++u) {
                const bool ok = i0 + 64 * u < n;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const unsigned int x = ok ? v[u][c] : 0u;
                    s[c] += x;
                    if (HH == 0) k[c] += (x != 0u);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            s[c] += __shfl_xor_sync(0xffffffffu, s[c], o);
            if (HH == 0) k[c] += __shfl_xor_sync(0xffffffffu, k[c], o);
        }
    }
    if (lane == 0 && half == 1 && node < n_nodes) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_sum[node][c] = s[c]; s_known[node][c] = k[c]; }
    }
    __syncthreads();
    if (lane == 0 && half == 0 && node < n_nodes) {
#pragma unroll
        for (int c = 0; c < 4; ++c) { s_sum[node][c] += s[c]; s_known[node][c] += k[c]; }
    }
    __syncthreads();
}

/* One CTA of 256 threads per query; hc = height of the children the last launch created (2 <= hc) */
__global__ void __launch_bounds__(256)
k_bbg_dive(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all, BbWork W, int hc)
{
    static_assert(kDiveWidth * 2 == 8 && kDiveWidth * 4 <= 32, "two warps per node, one lane per child");
    __shared__ int s_node[2][kDiveWidth][3];               /* (t, xi, yi), ping-pong by height */
    __shared__ int s_n;
    __shared__ unsigned int s_sum[kDiveWidth][4], s_known[kDiveWidth][4];
    const int q = blockIdx.x;
    const DevQuery& Q = queries[q];
    if (threadIdx.x < 32) {
        /* lanes 0..3 fetch (and clear: the next recording launch starts over) the start words */
        unsigned long long pw = 0ull;
        if (threadIdx.x < kDiveStarts) {
            pw = W.probe[q * kDiveStarts + threadIdx.x];
            if (pw != 0ull) W.probe[q * kDiveStarts + threadIdx.x] = 0ull;
        }
        const unsigned int have = __ballot_sync(0xffffffffu, pw != 0ull);
        if (pw != 0ull) {
            const int at = __popc(have & ((1u << threadIdx.x) - 1u));
            s_node[hc & 1][at][0] = (int)((pw >> (2 * kProbePosBits)) & 0xfffull);
            s_node[hc & 1][at][1] = (int)((pw >> kProbePosBits) & ((1ull << kProbePosBits) - 1ull));
            s_node[hc & 1][at][2] = (int)(pw & ((1ull << kProbePosBits) - 1ull));
        }
        if (threadIdx.x == 0) s_n = __popc(have);
    }
    __syncthreads();
    const proj_t* __restrict__ pq = proj_all + Q.proj_off;
    /* nothing but this CTA raises this query's incumbent while it runs: one reading serves every height */
    const unsigned long long inc0 = *(volatile unsigned long long*)&W.incumbent[q];
    const long long fail_max = Q.kthr.fail_max;
    const int n_beams = Q.n;
    for (int hh = hc - 1; hh >= 0; --hh) {
        const int n_nodes = s_n;
        if (n_nodes == 0)
            return;
        const int (*cur)[3] = s_node[(hh + 1) & 1];
        switch (hh) {
        case 0: dive_score<0>(Q, pq, cur, n_nodes, s_sum, s_known); break;
        case 1: dive_score<1>(Q, pq, cur, n_nodes, s_sum, s_known); break;
        case 2: dive_score<2>(Q, pq, cur, n_nodes, s_sum, s_known); break;
        case 3: dive_score<3>(Q, pq, cur, n_nodes, s_sum, s_known); break;
        case 4: dive_score<4>(Q, pq, cur, n_nodes, s_sum, s_known); break;
        default: dive_score<5>(Q, pq, cur, n_nodes, s_sum, s_known); break;
        }
        if (threadIdx.x < 32) {
            /* lane = (node, child) */
            const int lane = threadIdx.x, nd = lane >> 2, c = lane & 3;
            const bool valid = lane < 4 * n_nodes;
            const int w = 1 << hh;
            const int t = valid ? cur[nd][0] : 0;
            const int xi = valid ? cur[nd][1] + (c & 1) * w : 0, yi = valid ? cur[nd][2] + (c >> 1) * w : 0;
            const unsigned int sv = valid ? s_sum[nd][c] : 0u;
            if (hh == 0) {
                /* every passing leaf competes for the incumbent */
                long long key;
                if (valid && bb_passes(Q, proj_all, W, q, t, xi, yi, 0, (int)sv, (int)s_known[nd][c], key))
                    bb_raise_incumbent(W, q, key, leaf_ordfield(Q, t, xi, yi));
                if (lane == 0) s_n = 0;
            } else {
                /* the kDiveWidth children with the largest bounds among those the sweep would keep */
                const long long key_ub = make_key(257ll * (long long)sv, n_beams);
                const bool pass = valid && pack_best(key_ub, kOrdMask) > inc0 && key_ub > fail_max;
                const unsigned int mine = pass ? sv + 1u : 0u;
                int rank = 0;
#pragma unroll
                for (int o = 0; o < 4 * kDiveWidth; ++o) {
                    const unsigned int other = __shfl_sync(0xffffffffu, mine, o);
                    rank += (other > mine) || (other == mine && o < lane);
                }
                const unsigned int kept = __ballot_sync(0xffffffffu, pass && rank < kDiveWidth);
                if (pass && rank < kDiveWidth) {
                    s_node[hh & 1][rank][0] = t; s_node[hh & 1][rank][1] = xi; s_node[hh & 1][rank][2] = yi;
                }
                if (lane == 0) s_n = __popc(kept);
            }
        }
        __syncthreads();
    }
}

template <int HC>
__global__ void __launch_bounds__(32 * kBbgWarps, CSM_BB_MINB)
k_bbg_expand(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all, BbWork W, int nchunks_hint)
{
    constexpr int h = HC + 1;
    constexpr int w = 1 << HC;
    constexpr bool kLeaf = HC == 0;
    __shared__ unsigned int s_part[kBbgWarps][8][8];      /* per warp: per angle, 4 sums (+ 4 known counts) */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = lane & 7, part = lane >> 3;          /* angle within the group, beam within a quad */
    const unsigned int count = min(W.counts[h], W.capacity);
    const int wpi = bbg_warps_per_group(count, gridDim.x, nchunks_hint);
    const int ipc = kBbgWarps / wpi;                      /* groups a CTA takes at a time */
    const int sub = warp & (wpi - 1), local = warp / wpi;
    const unsigned long long* __restrict__ in = bb_list(W, h);
    unsigned long long* __restrict__ out = bb_list(W, HC);
    for (unsigned int base = blockIdx.x * ipc; base < count; base += gridDim.x * ipc) {
        const unsigned int idx = base + local;
        const bool have = idx < count;
        const unsigned long long item = have ? in[idx] : 0ull;
        const int q = (int)(item >> 48), tg = (int)((item >> 40) & 0xffull);
        const unsigned int mask = (unsigned int)((item >> 32) & 0xffull);
        const int xi = (int)((item >> 16) & 0xffffull), yi = (int)(item & 0xffffull);
        const bool alive = have && ((mask >> slot) & 1u);
        const int t = tg * 8 + slot;
        unsigned int s0 = 0, s1 = 0, s2 = 0, s3 = 0, k0 = 0, k1 = 0, k2 = 0, k3 = 0;
        if (alive) {
            const DevQuery& Q = queries[q];
            const uint16_t* __restrict__ m = Q.lvl[0];
            const unsigned char* __restrict__ bm = Q.bl[HC];
            const unsigned int tpr = (unsigned int)Q.bl_tpr[HC];
            const unsigned int row_w = (unsigned int)(w / kBlTileR) * (tpr << 7);
            const int rows = Q.rows, cols = Q.cols, n = Q.n;
            const int ox = xi - Q.winx, oy = yi - Q.winy;
            /* chunk k (16 beams) of angle t starts at ((k * tp + t) << 4); this lane's beams 16 k + 4 s + part,
             * s = 0..3, are the four entries at + (part << 2): one 16-byte load. This warp takes the chunks
             * k = sub, sub + wpi, ... */
            const unsigned int tp = (unsigned int)Q.tp;
            const uint4* __restrict__ pp = reinterpret_cast<const uint4*>(
                proj_all + Q.proj_off + ((((unsigned int)sub * tp + (unsigned int)t) << 4) + ((unsigned int)part << 2)));
            const unsigned int step = (tp * (unsigned int)wpi) << 2;        /* in 16-byte words */
            const int nchunks = (n + 15) >> 4;
            const int mine = (nchunks - sub + wpi - 1) / wpi;
            uint4 nx = make_uint4(0u, 0u, 0u, 0u);
            if (mine > 0) nx = __ldg(pp);
            for (int j = 0; j < mine; ++j) {
                const uint4 cur = nx;
                if (j + 1 < mine) nx = __ldg(pp + (unsigned int)(j + 1) * step);
                const unsigned int pw[4] = { cur.x, cur.y, cur.z, cur.w };
                unsigned int v[4][4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    /* slots past the last beam hold a cell outside every map (k_project): no test per beam */
                    const int r = (int)(short)(pw[s] >> 16) + oy, c = (int)(short)(pw[s] & 0xffffu) + ox;
                    if (kLeaf) ld_children<0>(m, rows, cols, r, c, v[s]);
                    else ld_children_b<HC>(bm, tpr, row_w, rows, cols, r, c, v[s]);
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    s0 += v[s][0]; s1 += v[s][1]; s2 += v[s][2]; s3 += v[s][3];
                    if (kLeaf) {
                        k0 += (v[s][0] != 0u); k1 += (v[s][1] != 0u);
                        k2 += (v[s][2] != 0u); k3 += (v[s][3] != 0u);
                    }
                }
            }
        }
        /* sum over the four beams of a quad (lanes 8 and 16 apart) */
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o); s3 += __shfl_xor_sync(0xffffffffu, s3, o);
            if (kLeaf) {
                k0 += __shfl_xor_sync(0xffffffffu, k0, o); k1 += __shfl_xor_sync(0xffffffffu, k1, o);
                k2 += __shfl_xor_sync(0xffffffffu, k2, o); k3 += __shfl_xor_sync(0xffffffffu, k3, o);
            }
        }
        if (wpi > 1) {
            /* ... and over the warps that share the group */
            if (part == 0) {
                unsigned int* d = s_part[warp][slot];
                d[0] = s0; d[1] = s1; d[2] = s2; d[3] = s3;
                if (kLeaf) { d[4] = k0; d[5] = k1; d[6] = k2; d[7] = k3; }
            }
            __syncthreads();
            if (sub == 0) {
                s0 = s1 = s2 = s3 = 0; k0 = k1 = k2 = k3 = 0;
                for (int o = 0; o < wpi; ++o) {
                    const unsigned int* d = s_part[warp + o][slot];
                    s0 += d[0]; s1 += d[1]; s2 += d[2]; s3 += d[3];
                    if (kLeaf) { k0 += d[4]; k1 += d[5]; k2 += d[6]; k3 += d[7]; }
                }
            }
        }
        if (sub == 0) {
            /* lane (slot, part) decides child `part` of angle `slot`: (x + (part & 1) w, y + (part >> 1) w) */
            const unsigned int sv = part == 0 ? s0 : part == 1 ? s1 : part == 2 ? s2 : s3;
            const int cx = xi + (part & 1) * w, cy = yi + (part >> 1) * w;
            bool pass = false;
            if (alive) {
                const DevQuery& Q = queries[q];
                if (kLeaf) {
                    const int k = (int)(part == 0 ? k0 : part == 1 ? k1 : part == 2 ? k2 : k3);
                    long long key;
                    pass = bb_passes(Q, proj_all, W, q, t, cx, cy, 0, (int)sv, k, key);
                    if (pass)
                        bb_raise_incumbent(W, q, key, leaf_ordfield(Q, t, cx, cy));
                } else {
                    /* upper bound of the key: 257 * B >= v per cell, every beam counted as known */
                    const long long key_ub = make_key(257ll * (long long)sv, Q.n);
                    const unsigned long long inc = *(volatile unsigned long long*)&W.incumbent[q];
                    pass = pack_best(key_ub, kOrdMask) > inc && key_ub > Q.kthr.fail_max && Q.n > Q.nk_cut;
                }
            }
            const unsigned int ballot = __ballot_sync(0xffffffffu, pass);
            const unsigned int scored = __ballot_sync(0xffffffffu, alive);
            if (have && lane == 0) {
                atomicAdd(&W.stats[2 * q], __popc(ballot));
                atomicAdd(&W.stats[2 * q + 1], __popc(scored) - __popc(ballot));
            }
            if (!kLeaf && W.probe != nullptr && ((W.probe_heights >> HC) & 1u) && ballot != 0u) {
                /* the passing child with the largest bound of this group competes for a start of the dive */
                const unsigned int cand = pass ? sv : 0u;
                const unsigned int top_sv = __reduce_max_sync(0xffffffffu, cand);
                const unsigned int who = __ballot_sync(0xffffffffu, pass && cand == top_sv);
                if (lane == __ffs(who) - 1)
                    atomicMax(&W.probe[q * kDiveStarts + (tg & (kDiveStarts - 1))], pack_probe(sv, t, cx, cy));
            }
            if (!kLeaf && ballot != 0u) {
                /* one new group per child that keeps an angle alive: lanes 0..3 write them */
                const unsigned int cm = (ballot >> (8 * (lane & 3))) & 0xffu;
                const bool emit = lane < 4 && cm != 0u;
                const unsigned int eb = __ballot_sync(0xffffffffu, emit);
                unsigned int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&W.counts[HC], (unsigned int)__popc(eb));
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (emit) {
                    const unsigned int dst = slot0 + __popc(eb & ((1u << lane) - 1u));
                    if (dst < W.capacity) out[dst] = pack_group(q, tg, cm, xi + (lane & 1) * w, yi + ((lane >> 1) & 1) * w);
                    else *W.overflow = 1;
                }
            }
        }
        if (wpi > 1)
            __syncthreads();          /* s_part is rewritten by the next round */
    }
}

/* ---- incumbent dive ----------------------------------------------------------
 * The level-synchronous search only meets leaves at the last launch, so
 * without help nothing is ever pruned against a found pose (with a zero score
 * threshold it would degenerate into the exhaustive search). One CTA per query
 * therefore descends first: the kDiveBeam best root nodes, then at every height
 * the kDiveBeam best of their children, down to a leaf, whose packed (key,
 * ordinal) becomes the query's incumbent. Any real leaf is a valid incumbent:
 * the search result (maximum key, smallest ordinal) does not depend on it,
 * only the amount of work does. */
constexpr int kDiveBeam = 8;
constexpr int kDiveRootsSmem = 1024;     /* roots ranked in shared memory when there are at most this many */

__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned long long r = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = s_red[w] > r ? s_red[w] : r;
    return r;
}

__global__ void __launch_bounds__(256)
k_bb_dive(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all,
          const unsigned int* __restrict__ root_off, BbWork W)
{
    __shared__ unsigned long long s_red[8];
    __shared__ int s_beam[2][kDiveBeam][3];          /* (t, xi, yi) */
    __shared__ int s_nbeam;
    __shared__ long long s_key[4 * kDiveBeam];
    __shared__ int s_ok[4 * kDiveBeam];
    __shared__ long long s_root[kDiveRootsSmem];
    const int q = blockIdx.x;
    const DevQuery& Q = queries[q];
    const unsigned int r0 = root_off[q];
    const unsigned int R = root_off[q + 1] - r0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (R >= (1u << 20))
        return;
    /* the kDiveBeam best roots, in (key descending, index ascending) order */
    unsigned long long last = ~0ull;
    int nbeam = 0;
    if (R <= (unsigned int)kDiveRootsSmem) {
        /* few roots (single-scan matches): one pass, every passing root ranks itself against the
         * others in shared memory */
        if (threadIdx.x == 0) s_nbeam = 0;
        for (unsigned int e = threadIdx.x; e < R; e += blockDim.x) s_root[e] = W.rootkey[r0 + e];
        __syncthreads();
        for (unsigned int e = threadIdx.x; e < R; e += blockDim.x) {
            const long long k = s_root[e];
            if (k < 0) continue;
            int rank = 0;
            for (unsigned int o = 0; o < R; ++o) {
                const long long ko = s_root[o];
                rank += (ko > k || (ko == k && o < e)) ? 1 : 0;
            }
            if (rank < kDiveBeam) {
                const int t = (int)(e % (unsigned)Q.T), cell = (int)(e / (unsigned)Q.T);
                const int rx = cell / Q.nry, ry = cell - rx * Q.nry;
                s_beam[0][rank][0] = t; s_beam[0][rank][1] = rx << W.top; s_beam[0][rank][2] = ry << W.top;
                atomicAdd(&s_nbeam, 1);
            }
        }
        __syncthreads();
        nbeam = min(s_nbeam, kDiveBeam);
    } else
    for (int j = 0; j < kDiveBeam; ++j) {
        unsigned long long best = 0ull;
        for (unsigned int e = threadIdx.x; e < R; e += blockDim.x) {
            const long long k = W.rootkey[r0 + e];
            if (k < 0) continue;
            const unsigned long long v = ((unsigned long long)(k + 1) << 20) | (unsigned long long)(0xFFFFFu - e);
            if (v < last && v > best) best = v;
        }
        best = block_max_u64(best, s_red);
        if (best == 0ull)
            break;
        last = best;
        if (threadIdx.x == 0) {
            const int e = (int)(0xFFFFFu - (unsigned int)(best & 0xFFFFFull));
            const int t = e % Q.T, cell = e / Q.T;
            const int rx = cell / Q.nry, ry = cell - rx * Q.nry;
            s_beam[0][j][0] = t; s_beam[0][j][1] = rx << W.top; s_beam[0][j][2] = ry << W.top;
        }
        ++nbeam;
    }
    __syncthreads();
    int cur = 0;
    for (int h = W.top; h >= 1 && nbeam > 0; --h) {
        const int w = 1 << (h - 1);
        const uint16_t* __restrict__ m = Q.lvl[h - 1];
        const int ncand = 4 * nbeam;
        {
            /* warp b scores the four children of beam node b: one projected-index load per beam
             * serves four cell loads, six beams (24 cells) per lane in flight */
            static_assert(kDiveBeam == 8, "one warp per beam node");
            const int rows = Q.rows, cols = Q.cols, n = Q.n;
            const size_t ps = (size_t)Q.pst_i;
            const bool on = warp < nbeam;
            const int b = on ? warp : 0;
            const proj_t* __restrict__ pp = proj_all + Q.proj_off + (size_t)s_beam[cur][b][0] * Q.pst_t;
            const int ox = s_beam[cur][b][1] - Q.winx, oy = s_beam[cur][b][2] - Q.winy;
            int sv[4] = { 0, 0, 0, 0 }, kn[4] = { 0, 0, 0, 0 };
            if (on) {
                int i = lane;
                for (; i + 5 * 32 < n; i += 6 * 32) {
                    proj_t p[6];
                    unsigned int v[6][4];
#pragma unroll
                    for (int u = 0; u < 6; ++u) p[u] = pp[(size_t)(i + 32 * u) * ps];
#pragma unroll
                    for (int u = 0; u < 6; ++u) {
                        const int r = p[u].y + oy, c = p[u].x + ox;
                        v[u][0] = ld_cell_nb(m, rows, cols, r, c);
                        v[u][1] = ld_cell_nb(m, rows, cols, r, c + w);
                        v[u][2] = ld_cell_nb(m, rows, cols, r + w, c);
                        v[u][3] = ld_cell_nb(m, rows, cols, r + w, c + w);
                    }
#pragma unroll
                    for (int u = 0; u < 6; ++u)
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) { sv[ch] += (int)v[u][ch]; kn[ch] += (v[u][ch] != 0u); }
                }
                for (; i < n; i += 32) {
                    const proj_t p = pp[(size_t)i * ps];
                    const int r = p.y + oy, c = p.x + ox;
                    const unsigned int v0 = ld_cell_nb(m, rows, cols, r, c), v1 = ld_cell_nb(m, rows, cols, r, c + w);
                    const unsigned int v2 = ld_cell_nb(m, rows, cols, r + w, c), v3 = ld_cell_nb(m, rows, cols, r + w, c + w);
                    sv[0] += (int)v0; sv[1] += (int)v1; sv[2] += (int)v2; sv[3] += (int)v3;
                    kn[0] += (v0 != 0u); kn[1] += (v1 != 0u); kn[2] += (v2 != 0u); kn[3] += (v3 != 0u);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
                const int s = warp_sum(sv[ch]), k = warp_sum(kn[ch]);
                if (on && lane == 0) {
                    const long long key = make_key(s, k);
                    s_key[4 * b + ch] = key;
                    /* strictly above the threshold (outside its guard band) and known enough */
                    s_ok[4 * b + ch] = (k > Q.nk_cut && key_vs_threshold(key, Q.kthr) > 0) ? 1 : 0;
                }
            }
        }
        __syncthreads();
        if (h - 1 == 0) {
            if (threadIdx.x == 0) {
                unsigned long long best = 0ull;
                for (int c = 0; c < ncand; ++c) {
                    if (!s_ok[c]) continue;
                    const int b = c >> 2, ch = c & 3;
                    const int xi = s_beam[cur][b][1] + (ch & 1), yi = s_beam[cur][b][2] + (ch >> 1);
                    const unsigned long long v = pack_best(s_key[c], leaf_ordfield(Q, s_beam[cur][b][0], xi, yi));
                    best = v > best ? v : best;
                }
                if (best != 0ull)
                    atomicMax(&W.incumbent[q], best);
            }
            break;
        }
        /* next beam: the kDiveBeam best passing children (rank by key, then by position) */
        if (threadIdx.x == 0) s_nbeam = 0;
        __syncthreads();
        if ((int)threadIdx.x < ncand && s_ok[threadIdx.x]) {
            const int c = threadIdx.x;
            int rank = 0;
            for (int o = 0; o < ncand; ++o)
                if (s_ok[o] && (s_key[o] > s_key[c] || (s_key[o] == s_key[c] && o < c)))
                    ++rank;
            if (rank < kDiveBeam) {
                const int b = c >> 2, ch = c & 3;
                s_beam[cur ^ 1][rank][0] = s_beam[cur][b][0];
                s_beam[cur ^ 1][rank][1] = s_beam[cur][b][1] + (ch & 1) * w;
                s_beam[cur ^ 1][rank][2] = s_beam[cur][b][2] + (ch >> 1) * w;
                atomicAdd(&s_nbeam, 1);
            }
        }
        __syncthreads();
        nbeam = min(s_nbeam, kDiveBeam);
        cur ^= 1;
        __syncthreads();
    }
}

/* ------------------------------------------------------------------------ */
/* Exhaustive grid search                                                    */
/* ------------------------------------------------------------------------ */

struct GridArgs
{
    const int* mx;          /* integer cell offset of dx[k] relative to dx[0] */
    const int* my;
    const double* px;       /* sx + dx[k] (general path) */
    const double* py;
    int ndx, ndy, ndt;
    unsigned long long* best;   /* packed (key, ordfield) */
    unsigned long long* tiekey; /* largest key two different candidates were seen to share (0: none) */
    int ord_mode;               /* tie order of the candidates: 0 = (iy, ix, it), the grid search's loops;
                                   1 = (it, ix, iy), the branch-and-bound leaf order */
};

__device__ __forceinline__ unsigned long long grid_ordinal(const GridArgs& G, int iy, int ix, int it)
{
    return G.ord_mode ? ((unsigned long long)it * G.ndx + ix) * G.ndy + iy
                      : ((unsigned long long)iy * G.ndx + ix) * G.ndt + it;
}

/* Running best of a thread / warp together with the largest key seen twice */
struct BestTie
{
    unsigned long long best, tiekey;
};

__device__ __forceinline__ void best_merge(BestTie& b, unsigned long long v, unsigned long long vtie)
{
    if (v != 0ull && v != b.best && (v >> kOrdBits) == (b.best >> kOrdBits))
        vtie = max(vtie, v >> kOrdBits);
    b.tiekey = max(b.tiekey, vtie);
    b.best = max(b.best, v);
}

__device__ __forceinline__ void block_best_commit(BestTie b, unsigned long long* best, unsigned long long* tiekey)
{
    /* warp max, then one atomic per warp */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, b.best, o);
        const unsigned long long otie = __shfl_xor_sync(0xffffffffu, b.tiekey, o);
        best_merge(b, other, otie);
    }
    if ((threadIdx.x & 31) == 0 && b.best != 0ull) {
        const unsigned long long old = atomicMax(best, b.best);
        if (old != b.best && (old >> kOrdBits) == (b.best >> kOrdBits))
            b.tiekey = max(b.tiekey, b.best >> kOrdBits);
        if (b.tiekey != 0ull)
            atomicMax(tiekey, b.tiekey);
    }
}

/* Integer-shift path: all dx[k] / dy[k] are integer multiples of the
 * resolution apart, so candidate (iy, ix, it) reads cell
 * (row_i + my[iy], col_i + mx[ix]) of the angle's projected indices.
 * CTA = (angle it, 8 rows iy); warp = one row; lane = one ix. */
__global__ void __launch_bounds__(256)
k_grid_window(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all, GridArgs G)
{
    extern __shared__ proj_t s_proj[];
    const DevQuery& Q = queries[0];
    const int it = blockIdx.x;
    const proj_t* proj = proj_all + Q.proj_off + (size_t)it * Q.n;
    for (int i = threadIdx.x; i < Q.n; i += blockDim.x)
        s_proj[i] = proj[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iy = blockIdx.y * (blockDim.x >> 5) + warp;
    BestTie best = { 0ull, 0ull };
    if (iy < G.ndy) {
        const int oy = G.my[iy];
        const uint16_t* __restrict__ m = Q.lvl[0];
        for (int ixb = 0; ixb < G.ndx; ixb += 32) {
            const int ix = ixb + lane;
            if (ix >= G.ndx)
                break;
            const int ox = G.mx[ix];
            int s = 0, k = 0;
#pragma unroll 4
            for (int i = 0; i < Q.n; ++i) {
                const proj_t p = s_proj[i];
                const unsigned int v = ld_cell(m, Q.rows, Q.cols, p.y + oy, p.x + ox);
                s += (int)v;
                k += (v != 0u);
            }
            const long long key = make_key(s, k);
            if (k > Q.nk_cut) {
                const int c = key_vs_threshold(key, Q.kthr);
                bool ok = c > 0;
                if (c == 0)
                    ok = exact_normalized_score(m, Q.rows, Q.cols, proj, Q.pst_i, Q.n, ox, oy) > Q.kthr.thr;
                if (ok) {
                    const unsigned long long ord = grid_ordinal(G, iy, ix, it);
                    best_merge(best, pack_best(key, (kOrdMask - 1ull) - ord), 0ull);
                }
            }
        }
    }
    block_best_commit(best, G.best, G.tiekey);
}

/* Per-candidate FP64 path (steps that are not multiples of the resolution):
 * the reference's arithmetic per candidate and beam,
 * col = floor(((sx + dx) + r*cos - offx) / res), with device sin/cos. */
__global__ void __launch_bounds__(256)
k_grid_general(const DevQuery* __restrict__ queries, const double2* __restrict__ rcs_all,
               GridArgs G, int* __restrict__ qflags)
{
    extern __shared__ double2 s_rcs[];
    const DevQuery& Q = queries[0];
    const int it = blockIdx.x;
    const double2* rcs = rcs_all + Q.proj_off + (size_t)it * Q.n;
    for (int i = threadIdx.x; i < Q.n; i += blockDim.x)
        s_rcs[i] = rcs[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iy = blockIdx.y * (blockDim.x >> 5) + warp;
    BestTie best = { 0ull, 0ull };
    int flagged = 0;
    if (iy < G.ndy) {
        const double py = G.py[iy];
        const uint16_t* __restrict__ m = Q.lvl[0];
        for (int ixb = 0; ixb < G.ndx; ixb += 32) {
            const int ix = ixb + lane;
            if (ix >= G.ndx)
                break;
            const double px = G.px[ix];
            int s = 0, k = 0;
            double sum = 0.0;
            for (int i = 0; i < Q.n; ++i) {
                const double2 rc = s_rcs[i];
                const double ux = __ddiv_rn(__dsub_rn(__dadd_rn(px, rc.x), Q.offx), Q.res);
                const double uy = __ddiv_rn(__dsub_rn(__dadd_rn(py, rc.y), Q.offy), Q.res);
                const double fx = floor(ux), fy = floor(uy);
                const double gx = ux - fx, gy = uy - fy;
                if (gx < Q.margin || gx > 1.0 - Q.margin || gy < Q.margin || gy > 1.0 - Q.margin)
                    flagged = 1;
                const double lim = 1073741824.0;
                const int col = (int)fmin(fmax(fx, -lim), lim);
                const int row = (int)fmin(fmax(fy, -lim), lim);
                const unsigned int v = ld_cell(m, Q.rows, Q.cols, row, col);
                if (v != 0u) {
                    s += (int)v; ++k;
                    sum = __dadd_rn(sum, value_to_probability(v));
                }
            }
            const long long key = make_key(s, k);
            if (k > Q.nk_cut) {
                const int c = key_vs_threshold(key, Q.kthr);
                bool ok = c > 0;
                if (c == 0)
                    ok = __ddiv_rn(sum, (double)Q.n) > Q.kthr.thr;
                if (ok) {
                    const unsigned long long ord = grid_ordinal(G, iy, ix, it);
                    best_merge(best, pack_best(key, (kOrdMask - 1ull) - ord), 0ull);
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, flagged) && lane == 0)
        atomicOr(&qflags[0], 1);
    block_best_commit(best, G.best, G.tiekey);
}

/* Smallest min(row, col) over the non-zero cells of a map whose row or column is below `limit`
 * (`limit` if there is none): when it is at least 2^hmax, every coarse lookup at a negative index
 * covers unknown cells only, i.e. the 0 the reference reads there (grid_map.cpp:389-392) IS the
 * maximum of the window and its branch-and-bound bound stays admissible: a
 * CSM_FLAG_EDGE raised by the projection is then withdrawn (k_finalize). kMarginParts CTAs per map. */
struct MarginJob
{
    const uint16_t* base;
    int rows, cols, limit, slot;     /* the result goes to out[slot] */
};

constexpr int kMarginParts = 8;          /* CTAs per map */

/* out[slot] must hold `limit` when the kernel starts (k_low_margin_init) */
__global__ void __launch_bounds__(256)
k_low_margin_init(const MarginJob* __restrict__ jobs, int* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[jobs[i].slot] = jobs[i].limit;
}

__device__ __forceinline__ int low_margin_of(uint4 v, int r, int c, int best)
{
    const unsigned int w[4] = { v.x, v.y, v.z, v.w };
    #pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (w[k] & 0xffffu) best = min(best, min(r, c + 2 * k));
        if (w[k] >> 16) best = min(best, min(r, c + 2 * k + 1));
    }
    return best;
}

__global__ void __launch_bounds__(256)
k_low_margin(const MarginJob* __restrict__ jobs, int* __restrict__ out)
{
    const MarginJob J = jobs[blockIdx.x];
    const int lr = min(J.limit, J.rows), lc = min(J.limit, J.cols);
    const int tid = blockIdx.y * blockDim.x + threadIdx.x, nth = gridDim.y * blockDim.x;
    int best = J.limit;
    if ((J.cols & 7) == 0 && (lc & 7) == 0 && (reinterpret_cast<uintptr_t>(J.base) & 15u) == 0) {
        /* 8 cells per load: the low rows whole, then the low columns of the other rows */
        const int vpr = J.cols >> 3, vlc = lc >> 3;
        for (int e = tid; e < lr * vpr; e += nth) {
            const int r = e / vpr, c = (e - r * vpr) << 3;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(J.base + (size_t)r * J.cols + c));
            if (v.x | v.y | v.z | v.w) best = low_margin_of(v, r, c, best);
        }
        for (int e = tid; e < (J.rows - lr) * vlc; e += nth) {
            const int r = lr + e / vlc, c = (e % vlc) << 3;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(J.base + (size_t)r * J.cols + c));
            if (v.x | v.y | v.z | v.w) best = low_margin_of(v, r, c, best);
        }
    } else {
        for (int e = tid; e < lr * J.cols; e += nth) {
            const int r = e / J.cols, c = e - r * J.cols;
            if (J.base[(size_t)r * J.cols + c] != 0) best = min(best, min(r, c));
        }
        for (int e = tid; e < J.rows * lc; e += nth) {
            const int r = e / lc, c = e - r * lc;
            if (J.base[(size_t)r * J.cols + c] != 0) best = min(best, min(r, c));
        }
    }
    best = (int)__reduce_min_sync(0xffffffffu, (unsigned)best);
    if ((threadIdx.x & 31) == 0 && best < J.limit)
        atomicMin(&out[J.slot], best);
}

/* ------------------------------------------------------------------------ */
/* Finalisation                                                              */
/* ------------------------------------------------------------------------ */

struct FinalArgs
{
    int decode;               /* where the winner comes from: 1 = B&B incumbents, 2 = grid-search
                                 best word, 3 = replay of the real-time correlative blocks */
    int mode;                 /* 0 = window indices are cell offsets, 1 = grid fast, 2 = grid general */
    /* decode 1 */
    const unsigned long long* incumbent;
    const unsigned long long* tiekey;
    const int* stats;
    /* decode 2 */
    GridArgs G;
    /* decode 3 */
    const RtBlock* blocks;
    int low_res, nbx, nby;
    /* grid search: cell offsets per index / candidate positions / r*cos, r*sin */
    const int* mx;
    const int* my;
    const double* px;
    const double* py;
    const double2* rcs;
    unsigned long long* best_key;   /* device word for the cross-rank argmax */
    const int* qflags;              /* per-query flags raised by the projection */
    const int* overflow;            /* frontier overflow flag, copied behind the results */
    int query_index_base;
    int nq;
};

/* One warp per query: decode the winner, integer score and reference-order
 * double score of the winning pose on the level-0 map, packed best word. */
__global__ void __launch_bounds__(32)
k_finalize(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all,
           FinalArgs F, csm_result* __restrict__ results)
{
    __shared__ double s_prob[kMaxBeams];
    __shared__ __align__(16) RtBlock s_blocks[kReplayChunk];
    __shared__ BestState s_state;
    const int q = blockIdx.x;
    const DevQuery& Q = queries[q];
    const int lane = threadIdx.x & 31;

    if (F.decode == 3) {
        BestState st;
        if (!rt_replay_scan(Q, F.blocks, F.low_res, F.nbx, F.nby, st))
            rt_replay(Q, proj_all, F.blocks, F.low_res, F.nbx, F.nby, s_blocks, st);
        if (lane == 0) s_state = st;
    } else if (lane == 0) {
        BestState st;
        if (F.decode == 1) {
            const unsigned long long inc = F.incumbent[q];
            const unsigned long long ordf = inc & kOrdMask;
            st.found = (ordf != kOrdMask) ? 1 : 0;   /* initial incumbent carries the all-ones field */
            if (st.found) {
                unsigned long long ord = (kOrdMask - 1ull) - ordf;
                const int yi = (int)(ord % (unsigned)Q.ly); ord /= (unsigned)Q.ly;
                const int xi = (int)(ord % (unsigned)Q.lx); ord /= (unsigned)Q.lx;
                st.bx = xi - Q.winx; st.by = yi - Q.winy; st.bt = (int)ord;
            } else {
                /* reference: bestX = bestY = bestTheta = 0 (scan_matcher_branch_bound.cpp:145-147) */
                st.bx = 0; st.by = 0; st.bt = (Q.T - 1) / 2;
            }
            st.flags = (st.found && F.tiekey != nullptr && F.tiekey[q] == (inc >> kOrdBits)) ? 2 : 0;   /* CSM_FLAG_KEY_TIE */
            st.n_processed = F.stats[2 * q];
            st.n_ignored = F.stats[2 * q + 1];
        } else {
            const unsigned long long b = *F.G.best;
            st.found = b != 0ull ? 1 : 0;
            st.bx = st.by = st.bt = -1;
            if (st.found) {
                unsigned long long ord = (kOrdMask - 1ull) - (b & kOrdMask);
                if (F.G.ord_mode) {
                    st.by = (int)(ord % (unsigned)F.G.ndy); ord /= (unsigned)F.G.ndy;
                    st.bx = (int)(ord % (unsigned)F.G.ndx); ord /= (unsigned)F.G.ndx;
                    st.bt = (int)ord;
                } else {
                    st.bt = (int)(ord % (unsigned)F.G.ndt); ord /= (unsigned)F.G.ndt;
                    st.bx = (int)(ord % (unsigned)F.G.ndx); ord /= (unsigned)F.G.ndx;
                    st.by = (int)ord;
                }
            }
            st.flags = (st.found && F.G.tiekey != nullptr && *F.G.tiekey == (b >> kOrdBits)) ? 2 : 0;   /* CSM_FLAG_KEY_TIE */
            st.n_processed = F.G.ndx * F.G.ndy * F.G.ndt;
            st.n_ignored = 0;
        }
        st.pad = 0;
        s_state = st;
    }
    __syncwarp();
    const BestState s = s_state;

    const uint16_t* __restrict__ m = Q.lvl[0];
    csm_result r;
    r.found = s.found;
    r.flags = s.flags | (F.qflags != nullptr ? F.qflags[q] : 0);
    if ((r.flags & 4) && Q.low_margin != nullptr && *Q.low_margin >= Q.edge_need)
        r.flags &= ~4;        /* CSM_FLAG_EDGE: nothing known below row / column 2^hmax, the reference's bound holds */
    r.n_processed = s.n_processed;
    r.n_ignored = s.n_ignored;
    r.sum_value = 0; r.n_known = 0; r.normalized_score = 0.0;
    const bool evaluate = (F.mode == 0) || s.found;
    const int it = s.bt;
    r.best_x = s.bx; r.best_y = s.by;
    r.best_t = (F.mode == 0) ? s.bt - (Q.T - 1) / 2 : s.bt;
    int sumv = 0, nk = 0;
    if (evaluate) {
        const size_t row_off = (size_t)Q.proj_off + (size_t)it * Q.n;
        int ps = 0, pk = 0;
        int i0 = lane;
        if (F.mode != 2) {
            /* four beams per lane in flight: index loads, then cell loads, then the stores (a
             * dependent pair of memory round trips per group instead of per beam) */
            const int ox = (F.mode == 1) ? F.mx[s.bx] : s.bx;
            const int oy = (F.mode == 1) ? F.my[s.by] : s.by;
            const proj_t* __restrict__ pp = proj_all + (size_t)Q.proj_off;
            for (; i0 + 96 < Q.n; i0 += 128) {
                proj_t p[4];
                unsigned int v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) p[u] = pp[proj_index(Q, it, i0 + 32 * u)];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = ld_cell_nb(m, Q.rows, Q.cols, p[u].y + oy, p[u].x + ox);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    s_prob[i0 + 32 * u] = (v[u] != 0u) ? value_to_probability(v[u]) : 0.0;
                    ps += (int)v[u];
                    pk += (v[u] != 0u);
                }
            }
        }
        for (int i = i0; i < Q.n; i += 32) {
            int row, col;
            if (F.mode == 2) {
                const double2 rc = F.rcs[row_off + i];
                const double ux = __ddiv_rn(__dsub_rn(__dadd_rn(F.px[s.bx], rc.x), Q.offx), Q.res);
                const double uy = __ddiv_rn(__dsub_rn(__dadd_rn(F.py[s.by], rc.y), Q.offy), Q.res);
                const double lim = 1073741824.0;
                col = (int)fmin(fmax(floor(ux), -lim), lim);
                row = (int)fmin(fmax(floor(uy), -lim), lim);
            } else {
                const proj_t p = proj_all[(size_t)Q.proj_off + proj_index(Q, it, i)];
                const int ox = (F.mode == 1) ? F.mx[s.bx] : s.bx;
                const int oy = (F.mode == 1) ? F.my[s.by] : s.by;
                col = p.x + ox; row = p.y + oy;
            }
            const unsigned int v = ld_cell(m, Q.rows, Q.cols, row, col);
            /* unknown cells are skipped by the reference; adding +0.0 leaves the sum unchanged */
            s_prob[i] = (v != 0u) ? value_to_probability(v) : 0.0;
            ps += (int)v;
            pk += (v != 0u);
        }
        sumv = warp_sum(ps);
        nk = warp_sum(pk);
        __syncwarp();
        if (lane == 0) {
            /* the reference's sequential sum in scan order (score_function_pixel_accurate.cpp:21-57) */
            double sum = 0.0;
            int i = 0;
            for (; i + 8 <= Q.n; i += 8) {
                double pv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) pv[u] = s_prob[i + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) sum = __dadd_rn(sum, pv[u]);
            }
            for (; i < Q.n; ++i)
                sum = __dadd_rn(sum, s_prob[i]);
            r.normalized_score = __ddiv_rn(sum, (double)Q.n);
        }
    }
    if (lane == 0) {
        r.sum_value = sumv;
        r.n_known = nk;
        results[q] = r;
        if (s.found && F.best_key != nullptr) {
            const unsigned long long word =
                ((unsigned long long)make_key(sumv, nk) << 20) |
                (unsigned long long)(0xFFFFF - (F.query_index_base + q));
            atomicMax(F.best_key, word);
        }
        if (q == 0)     /* the overflow flag travels behind the results: one read-back */
            *reinterpret_cast<int*>(results + F.nq) = (F.overflow != nullptr) ? *F.overflow : 0;
    }
}

} /* namespace csm */
