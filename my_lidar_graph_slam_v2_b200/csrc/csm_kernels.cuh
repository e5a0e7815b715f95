/* csm_kernels.cuh -- CUDA kernels of the correlative scan matching hot path.
 * sm_100a only; compiled into libcsm_b200.so by csm_b200.cu.
 *
 * Kernel inventory (reference unit each one replaces, paths relative to the
 * reference repository):
 *   k_pyramid_level   PrecomputeGridMaps level h from level h-1
 *                     (grid_map_builder.cpp:987-1012, util.hpp:369-424)
 *   k_sliding_max     PrecomputeGridMap(map, win), any window
 *                     (grid_map_builder.cpp:1044-1065)
 *   k_project         ScanData::HitPoint + PositionToIndex for every
 *                     (query, angle, beam) (sensor_data.hpp:190-203,
 *                     grid_map_geometry.cpp:113-122)
 *   k_rt_blocks       ComputeScore on the coarse map + EvaluateHighResolutionMap
 *                     for every coarse cell (scan_matcher_correlative.cpp:301-368)
 *   k_rt_replay       the sequential accept/skip decisions of
 *                     scan_matcher_correlative.cpp:161-197 replayed on keys
 *   k_bb_roots / k_bb_dive / k_bb_expand
 *                     branch-and-bound as level-synchronous frontier expansion
 *                     (scan_matcher_branch_bound.cpp:156-231)
 *   k_grid_window     exhaustive (dy, dx, dtheta) search, integer-shift path
 *   k_grid_general    same, per-candidate FP64 projection (arbitrary steps)
 *                     (scan_matcher_grid_search.cpp:118-142)
 *   k_finalize        integer score + reference-order double score at the
 *                     winning pose, packed best word for the NCCL argmax
 */
#pragma once

#include "csm_device.cuh"
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
    int log2bs, block_cols, cols;
    size_t map_cells;
};

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
        const uint4 v = __ldg(A.data + (size_t)(first + b) * chunks_per_block + c);
        *reinterpret_cast<uint4*>(dense + (size_t)((brow << k) + r_in) * A.cols + (bcol << k) + c_in) = v;
    }
}

/* ------------------------------------------------------------------------ */
/* Precomputation                                                            */
/* ------------------------------------------------------------------------ */

struct PyrJob
{
    const uint16_t* base;   /* level 0 */
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
    if (c0 + 1 < cols) {
        *reinterpret_cast<unsigned int*>(dst + (size_t)r * cols + c0) = out[0] | (out[1] << 16);
    } else {
        dst[(size_t)r * cols + c0] = (uint16_t)out[0];
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
        unsigned int* dst = reinterpret_cast<unsigned int*>(job.levels + (size_t)(H - 1) * cells +
                                                            (size_t)r0 * C) + j;
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

__global__ void __launch_bounds__(kPsThreads, 2)
k_pyramid_stream(const PyrJob* __restrict__ jobs, int hmax)
{
    extern __shared__ __align__(16) unsigned int ps_smem[];
    unsigned int* in_ring = ps_smem;                                         /* [stages][4][272] */
    unsigned int* rowbuf = in_ring + kPsStages * kPsRows * kPsInStride;      /* [2][4][256] */
    unsigned int* ring = rowbuf + 2 * kPsRows * 256;                         /* [63][256] */

    const PyrJob job = jobs[blockIdx.x];
    const int R = job.rows, C = job.cols;
    const size_t cells = (size_t)R * C;
    const int j = threadIdx.x;
    const bool in_map = 2 * j < C;

    for (int i = j; i < 63 * 256; i += kPsThreads) ring[i] = 0u;
    for (int i = j; i < kPsStages * kPsRows * kPsInStride; i += kPsThreads) in_ring[i] = 0u;
    __syncthreads();

    const int nblocks = R / kPsRows;
    /* producer: this thread's 16-byte chunk of a 4-row block */
    const int ld_row = j >> 6, ld_chunk = j & 63;
    auto prefetch = [&](int b) {
        if (b >= 0) {
            const int stage = b % kPsStages;
            unsigned int* dst = in_ring + (stage * kPsRows + ld_row) * kPsInStride + ld_chunk * 4;
            const bool ok = ld_chunk * 8 < C;
            const uint16_t* src = job.base + (size_t)(b * kPsRows + ld_row) * C + (ok ? ld_chunk * 8 : 0);
            cp_async_16(dst, src, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
    for (int k = 0; k < kPsStages - 1; ++k)
        prefetch(nblocks - 1 - k);

    unsigned int* buf0 = rowbuf;
    unsigned int* buf1 = rowbuf + kPsRows * 256;
    for (int b = nblocks - 1; b >= 0; --b) {
        asm volatile("cp.async.wait_group %0;\n" :: "n"(kPsStages - 2));
        __syncthreads();                 /* block b landed; everyone is done with block b+1 */
        prefetch(b - (kPsStages - 1));   /* refills the stage block b+1 used */

        const unsigned int* in = in_ring + (b % kPsStages) * kPsRows * kPsInStride;
        unsigned int a[kPsRows];
#pragma unroll
        for (int rr = 0; rr < kPsRows; ++rr) a[rr] = in[rr * kPsInStride + j];
        ps_level<1>(a, in, kPsInStride, buf1, ring, job, b, j, in_map, cells);
        if (hmax >= 2) ps_level<2>(a, buf1, 256, buf0, ring, job, b, j, in_map, cells);
        if (hmax >= 3) ps_level<3>(a, buf0, 256, buf1, ring, job, b, j, in_map, cells);
        if (hmax >= 4) ps_level<4>(a, buf1, 256, buf0, ring, job, b, j, in_map, cells);
        if (hmax >= 5) ps_level<5>(a, buf0, 256, buf1, ring, job, b, j, in_map, cells);
        if (hmax >= 6) ps_level<6>(a, buf1, 256, buf0, ring, job, b, j, in_map, cells);
    }
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

/* ------------------------------------------------------------------------ */
/* Projection                                                                */
/* ------------------------------------------------------------------------ */

/* Per-scan table: (cos a_i, sin a_i) of every beam angle, computed once when
 * the scan is uploaded. */
__global__ void __launch_bounds__(256)
k_beam_trig(const double* __restrict__ angles, double2* __restrict__ trig, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    sincos(angles[i], &s, &c);
    trig[i] = make_double2(c, s);
}

constexpr int kProjAngles = 8;       /* candidate angles per CTA */
constexpr int kProjSat = 30000;      /* saturation of projected indices (maps are <= 16384 wide) */

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
        sincos(Q.thetas[t0 + threadIdx.x], &s, &c);
        s_theta[threadIdx.x] = make_double2(c, s);
    }
    __syncthreads();
    int flagged = 0;
    const int total = nt * Q.n;
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int tl = e / Q.n;
        const int i = e - tl * Q.n;
        const double2 th = s_theta[tl];
        const double2 be = Q.beam_trig[i];
        const double c = th.x * be.x - th.y * be.y;      /* cos(theta + a) */
        const double s = th.y * be.x + th.x * be.y;      /* sin(theta + a) */
        const double r = Q.ranges[i];
        const double rc = __dmul_rn(r, c);
        const double rs = __dmul_rn(r, s);
        const double ux = __ddiv_rn(__dsub_rn(__dadd_rn(Q.sx, rc), Q.offx), Q.res);
        const double uy = __ddiv_rn(__dsub_rn(__dadd_rn(Q.sy, rs), Q.offy), Q.res);
        const double fx = floor(ux), fy = floor(uy);
        const double gx = ux - fx, gy = uy - fy;
        if (gx < Q.margin || gx > 1.0 - Q.margin || gy < Q.margin || gy > 1.0 - Q.margin)
            flagged = 1;
        const double lim = (double)kProjSat;
        proj_t p;
        p.x = (short)(int)fmin(fmax(fx, -lim), lim);
        p.y = (short)(int)fmin(fmax(fy, -lim), lim);
        proj[(size_t)Q.proj_off + (size_t)(t0 + tl) * Q.pst_t + (size_t)i * Q.pst_i] = p;
        if (rcs != nullptr)
            rcs[(size_t)Q.proj_off + (size_t)(t0 + tl) * Q.n + i] = make_double2(rc, rs);
    }
    if (__any_sync(0xffffffffu, flagged) && (threadIdx.x & 31) == 0)
        atomicOr(&qflags[q], 1 /* CSM_FLAG_FP_MARGIN */);
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
k_rt_blocks(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all,
            RtBlock* __restrict__ blocks, int low_res, int nbx, int nby)
{
    extern __shared__ long long s_keys[];      /* L*L fine keys */
    const DevQuery& Q = queries[0];
    const int b = blockIdx.x;
    const int t = b / (nbx * nby);
    const int rem = b - t * nbx * nby;
    const int bx = rem / nby, by = rem - bx * nby;
    const int x = -Q.winx + bx * low_res;
    const int y = -Q.winy + by * low_res;
    const proj_t* proj = proj_all + Q.proj_off + (size_t)t * Q.n;
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
 * reference's order (t, x, y) on integer keys. One thread. Exact also when
 * the coarse bound is not admissible. */
__global__ void k_rt_replay(const DevQuery* __restrict__ queries,
                            const proj_t* __restrict__ proj_all,
                            const RtBlock* __restrict__ blocks, int low_res,
                            int nbx, int nby, BestState* __restrict__ state)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    const DevQuery& Q = queries[0];
    bool have = false;          /* scoreMax is a fine score (else: the threshold) */
    long long cur = 0;
    int bestx = -Q.winx, besty = -Q.winy, bestt = 0, flags = 0;
    int processed = 0, ignored = 0;
    const int nb = Q.T * nbx * nby;
    for (int b = 0; b < nb; ++b) {
        const RtBlock B = blocks[b];
        const int t = b / (nbx * nby);
        const int rem = b - t * nbx * nby;
        const int bx = rem / nby, by = rem - bx * nby;
        const int x = -Q.winx + bx * low_res, y = -Q.winy + by * low_res;
        const proj_t* proj = proj_all + Q.proj_off + (size_t)t * Q.n;
        bool coarse_ok;
        if (have) {
            coarse_ok = B.coarse_key > cur;
            if (B.coarse_key == cur) flags |= 2;   /* equal keys: doubles could round either way */
        } else {
            coarse_ok = passes_threshold(B.coarse_key, Q, Q.coarse, proj, x, y);
        }
        if (!coarse_ok || B.coarse_nk <= Q.nk_cut) {
            ++ignored;
            continue;
        }
        ++processed;
        bool fine_ok;
        const int fx = x + B.fine_ord / low_res, fy = y + B.fine_ord % low_res;
        if (have) fine_ok = B.fine_key > cur;
        else fine_ok = passes_threshold(B.fine_key, Q, Q.lvl[0], proj, fx, fy);
        if (fine_ok) {
            have = true;
            cur = B.fine_key;
            bestx = fx; besty = fy; bestt = t;
            if (B.fine_tie) flags |= 2;
        }
    }
    BestState s;
    s.found = have ? 1 : 0;
    s.bx = bestx; s.by = besty;
    s.bt = have ? bestt : 0;    /* reference initial bestWinTheta = -winTheta = index 0 */
    s.flags = flags; s.n_processed = processed; s.n_ignored = ignored; s.pad = 0;
    state[0] = s;
}

/* ------------------------------------------------------------------------ */
/* Branch and bound                                                          */
/* ------------------------------------------------------------------------ */

/* Level-synchronous frontier expansion, one LANE per node.
 *
 * The candidate list of height h holds unscored nodes (q, t, x, y). A warp
 * takes 32 consecutive candidates; every lane scores its node over all N
 * beams on the level-h map (its own integer sums, no warp reduction) and
 * decides like the reference does when it pops a node
 * (scan_matcher_branch_bound.cpp:191-198): drop iff score <= scoreMax or
 * knownRate <= threshold. Survivors of height h > 0 append their four
 * children (:226-229) to the list of height h-1; survivors of height 0 are
 * leaves and raise the query's incumbent with atomicMax.
 *
 * Why lanes and not warps per node: candidates are kept in runs of adjacent
 * angles t (roots are generated t-fastest, children are appended per child
 * type in lane order), and adjacent angles of the same (x, y) hit almost the
 * same cells. The 32 two-byte gathers of a warp then fall into a few 32-byte
 * sectors instead of 32, and with the projection stored beam-major
 * (proj[i][t]) the index loads of a warp are one contiguous segment. */
struct BbWork
{
    unsigned long long* list[2];    /* candidate lists, ping-pong by height parity */
    unsigned int*       counts;     /* [kMaxLevels]: number of candidates per height */
    unsigned long long* incumbent;  /* per query: packed (key, ordfield) */
    int*                stats;      /* per query: processed, ignored */
    int*                overflow;   /* set when a list is full */
    unsigned int        capacity;
    int                 hmax;
};

__device__ __forceinline__ unsigned long long leaf_ordfield(const DevQuery& Q, int t, int xi, int yi)
{
    /* smaller ordinal (t, x, y order) wins ties -> larger field */
    const unsigned long long ord =
        ((unsigned long long)t * (unsigned)Q.lx + (unsigned)xi) * (unsigned)Q.ly + (unsigned)yi;
    return (kOrdMask - 1ull) - ord;     /* all-ones is reserved for "no leaf yet" */
}

/* Root candidates of every query: (x, y) stepping by 2^hmax from -win, all
 * angles (scan_matcher_branch_bound.cpp:179-182), angle fastest. */
__global__ void __launch_bounds__(256)
k_bb_init(const DevQuery* __restrict__ queries, const unsigned int* __restrict__ root_off,
          int nq, BbWork W)
{
    const int q = blockIdx.y;
    const DevQuery& Q = queries[q];
    const int nroots = Q.T * Q.nrx * Q.nry;
    unsigned long long* out = W.list[W.hmax & 1] + root_off[q];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nroots; e += gridDim.x * blockDim.x) {
        const int t = e % Q.T;
        const int cell = e / Q.T;
        const int rx = cell / Q.nry, ry = cell - rx * Q.nry;
        out[e] = pack_node(q, t, rx << W.hmax, ry << W.hmax);
    }
    if (q == 0 && blockIdx.x == 0 && threadIdx.x == 0)
        W.counts[W.hmax] = root_off[nq];
}

constexpr int kBbSplit = 4;                 /* lanes cooperating on one node */
constexpr int kBbNodesPerWarp = 32 / kBbSplit;

__global__ void __launch_bounds__(256)
k_bb_score(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all,
           BbWork W, int h)
{
    /* lane = part * 8 + slot: `slot` selects one of 8 consecutive candidates,
     * `part` the quarter of the beams (i = part, part + 4, ...) this lane sums.
     * Lanes with equal `part` sit next to each other, so a warp's index loads
     * are 4 runs of 8 consecutive proj entries when the angles are consecutive. */
    const int lane = threadIdx.x & 31;
    const int slot = lane & (kBbNodesPerWarp - 1);
    const int part = lane / kBbNodesPerWarp;
    const unsigned int count = min(W.counts[h], W.capacity);
    const unsigned long long* __restrict__ in = W.list[h & 1];
    unsigned long long* __restrict__ out = W.list[(h & 1) ^ 1];
    const unsigned int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int w = (h > 0) ? (1 << (h - 1)) : 0;
    for (unsigned int base = warp_global * kBbNodesPerWarp; base < count; base += nwarps * kBbNodesPerWarp) {
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
            const size_t step = ps * kBbSplit;
            const int mine = (n - part + kBbSplit - 1) / kBbSplit;     /* beams of this lane */
            int i = 0;
            for (; i + 8 <= mine; i += 8) {
                proj_t p[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) p[u] = pp[(size_t)(i + u) * step];
                unsigned int v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_cell(m, rows, cols, p[u].y + oy, p[u].x + ox);
#pragma unroll
                for (int u = 0; u < 8; ++u) { s += (int)v[u]; k += (v[u] != 0u); }
            }
            for (; i < mine; ++i) {
                const proj_t p = pp[(size_t)i * step];
                const unsigned int v = ld_cell(m, rows, cols, p.y + oy, p.x + ox);
                s += (int)v; k += (v != 0u);
            }
        }
        /* sum the four parts of every node (lanes slot, slot+8, slot+16, slot+24) */
        s += __shfl_xor_sync(0xffffffffu, s, 8);  k += __shfl_xor_sync(0xffffffffu, k, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16); k += __shfl_xor_sync(0xffffffffu, k, 16);
        bool pass = false;
        long long key = 0;
        if (valid) {
            const DevQuery& Q = queries[q];
            key = make_key(s, k);
            const unsigned long long inc = *(volatile unsigned long long*)&W.incumbent[q];
            pass = k > Q.nk_cut && pack_best(key, kOrdMask) > inc;
            if (pass) {
                const int c = key_vs_threshold(key, Q.kthr);
                if (c < 0) pass = false;
                else if (c == 0) {
                    const proj_t* pp0 = proj_all + Q.proj_off + (size_t)t * Q.pst_t;
                    pass = exact_normalized_score(Q.lvl[h], Q.rows, Q.cols, pp0, Q.pst_i, Q.n,
                                                  xi - Q.winx, yi - Q.winy) > Q.kthr.thr;
                }
            }
            if (pass && h == 0 && part == 0)
                atomicMax(&W.incumbent[q], pack_best(key, leaf_ordfield(Q, t, xi, yi)));
        }
        {
            /* processed / ignored counters, one atomic per (warp, query, outcome) */
            const int tag = (valid && part == 0) ? (2 * q + (pass ? 0 : 1)) : -1;
            const unsigned int peers = __match_any_sync(0xffffffffu, tag);
            if (tag >= 0 && lane == __ffs(peers) - 1)
                atomicAdd(&W.stats[tag], __popc(peers));
        }
        if (h > 0) {
            /* every lane of a surviving node writes one child: child type = part */
            const unsigned int ballot = __ballot_sync(0xffffffffu, pass) & 0xffu;
            if (ballot != 0u) {
                const int npass = __popc(ballot);
                unsigned int slot0 = 0;
                if (lane == 0) slot0 = atomicAdd(&W.counts[h - 1], 4u * npass);
                slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                if (pass) {
                    const unsigned int rank = __popc(ballot & ((1u << slot) - 1u));
                    /* one contiguous run per child type, candidate order kept inside it */
                    const unsigned int dst = slot0 + part * npass + rank;
                    if (dst < W.capacity)
                        out[dst] = pack_node(q, t, xi + (part & 1) * w, yi + (part >> 1) * w);
                    else
                        *W.overflow = 1;
                }
            }
        }
    }
}

/* Decode the incumbents into BestState records */
__global__ void k_bb_collect(const DevQuery* __restrict__ queries, BbWork W, int nq,
                             const int* __restrict__ qflags, BestState* __restrict__ state)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq)
        return;
    const DevQuery& Q = queries[q];
    const unsigned long long inc = W.incumbent[q];
    BestState s;
    const unsigned long long ordf = inc & kOrdMask;
    s.found = (ordf != kOrdMask) ? 1 : 0;   /* initial incumbent carries the all-ones field */
    if (s.found) {
        unsigned long long ord = (kOrdMask - 1ull) - ordf;
        const int yi = (int)(ord % (unsigned)Q.ly); ord /= (unsigned)Q.ly;
        const int xi = (int)(ord % (unsigned)Q.lx); ord /= (unsigned)Q.lx;
        s.bx = xi - Q.winx; s.by = yi - Q.winy; s.bt = (int)ord;
    } else {
        /* reference: bestX = bestY = bestTheta = 0 (scan_matcher_branch_bound.cpp:145-147) */
        s.bx = 0; s.by = 0; s.bt = (Q.T - 1) / 2;
    }
    s.flags = qflags[q];
    s.n_processed = W.stats[2 * q];
    s.n_ignored = W.stats[2 * q + 1];
    s.pad = 0;
    state[q] = s;
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
    int* tie;
};

__device__ __forceinline__ void block_best_commit(unsigned long long v, unsigned long long* best)
{
    /* warp max, then one atomic per warp */
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    if ((threadIdx.x & 31) == 0 && v != 0ull)
        atomicMax(best, v);
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
    unsigned long long best = 0ull;
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
                    const unsigned long long ord =
                        ((unsigned long long)iy * G.ndx + ix) * G.ndt + it;
                    const unsigned long long v = pack_best(key, (kOrdMask - 1ull) - ord);
                    best = v > best ? v : best;
                }
            }
        }
    }
    block_best_commit(best, G.best);
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
    unsigned long long best = 0ull;
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
                    const unsigned long long ord =
                        ((unsigned long long)iy * G.ndx + ix) * G.ndt + it;
                    const unsigned long long v = pack_best(key, (kOrdMask - 1ull) - ord);
                    best = v > best ? v : best;
                }
            }
        }
    }
    if (__any_sync(0xffffffffu, flagged) && lane == 0)
        atomicOr(&qflags[0], 1);
    block_best_commit(best, G.best);
}

__global__ void k_grid_collect(GridArgs G, const int* __restrict__ qflags,
                               BestState* __restrict__ state)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    const unsigned long long b = *G.best;
    BestState s;
    s.found = b != 0ull ? 1 : 0;
    s.bx = s.by = s.bt = -1;
    if (s.found) {
        unsigned long long ord = (kOrdMask - 1ull) - (b & kOrdMask);
        s.bt = (int)(ord % (unsigned)G.ndt); ord /= (unsigned)G.ndt;
        s.bx = (int)(ord % (unsigned)G.ndx); ord /= (unsigned)G.ndx;
        s.by = (int)ord;
    }
    s.flags = qflags[0];
    s.n_processed = G.ndx * G.ndy * G.ndt;
    s.n_ignored = 0;
    s.pad = 0;
    state[0] = s;
}

/* ------------------------------------------------------------------------ */
/* Finalisation                                                              */
/* ------------------------------------------------------------------------ */

struct FinalArgs
{
    const int* mx;            /* grid search: cell offsets per index (else nullptr) */
    const int* my;
    const double* px;         /* grid search general path: sx + dx[k] (else nullptr) */
    const double* py;
    const double2* rcs;
    unsigned long long* best_key;   /* device word for the cross-rank argmax */
    const int* qflags;              /* per-query flags raised by the projection */
    int query_index_base;
    int mode;                 /* 0 = window indices are cell offsets, 1 = grid fast, 2 = grid general */
};

/* One warp per query: integer score and reference-order double score of the
 * winning pose on the level-0 map; packs the per-batch best word. */
__global__ void __launch_bounds__(32)
k_finalize(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all,
           const BestState* __restrict__ state, FinalArgs F, csm_result* __restrict__ results)
{
    __shared__ double s_prob[kMaxBeams];
    const int q = blockIdx.x;
    const DevQuery& Q = queries[q];
    const BestState s = state[q];
    const int lane = threadIdx.x & 31;
    const uint16_t* __restrict__ m = Q.lvl[0];
    csm_result r;
    r.found = s.found;
    r.flags = s.flags | (F.qflags != nullptr ? F.qflags[q] : 0);
    r.n_processed = s.n_processed;
    r.n_ignored = s.n_ignored;
    r.sum_value = 0; r.n_known = 0; r.normalized_score = 0.0;
    const bool evaluate = (F.mode == 0) || s.found;
    int it = s.bt;
    if (F.mode == 0) {
        r.best_x = s.bx; r.best_y = s.by; r.best_t = s.bt - (Q.T - 1) / 2;
    } else {
        r.best_x = s.bx; r.best_y = s.by; r.best_t = s.bt;
    }
    int sumv = 0, nk = 0;
    if (evaluate) {
        const size_t row_off = (size_t)Q.proj_off + (size_t)it * Q.n;
        int ps = 0, pk = 0;
        for (int i = lane; i < Q.n; i += 32) {
            int row, col;
            if (F.mode == 2) {
                const double2 rc = F.rcs[row_off + i];
                const double ux = __ddiv_rn(__dsub_rn(__dadd_rn(F.px[s.bx], rc.x), Q.offx), Q.res);
                const double uy = __ddiv_rn(__dsub_rn(__dadd_rn(F.py[s.by], rc.y), Q.offy), Q.res);
                const double lim = 1073741824.0;
                col = (int)fmin(fmax(floor(ux), -lim), lim);
                row = (int)fmin(fmax(floor(uy), -lim), lim);
            } else {
                const proj_t p = proj_all[(size_t)Q.proj_off + (size_t)it * Q.pst_t + (size_t)i * Q.pst_i];
                const int ox = (F.mode == 1) ? F.mx[s.bx] : s.bx;
                const int oy = (F.mode == 1) ? F.my[s.by] : s.by;
                col = p.x + ox; row = p.y + oy;
            }
            const unsigned int v = ld_cell(m, Q.rows, Q.cols, row, col);
            s_prob[i] = (v != 0u) ? value_to_probability(v) : -1.0;
            ps += (int)v;
            pk += (v != 0u);
        }
        sumv = warp_sum(ps);
        nk = warp_sum(pk);
        __syncwarp();
        if (lane == 0) {
            double sum = 0.0;
            for (int i = 0; i < Q.n; ++i) {
                const double pv = s_prob[i];
                if (pv >= 0.0)
                    sum = __dadd_rn(sum, pv);
            }
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
    }
}

} /* namespace csm */
