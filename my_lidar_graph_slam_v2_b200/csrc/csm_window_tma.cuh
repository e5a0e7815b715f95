/* csm_window_tma.cuh -- correlative score evaluation over a full (x, y, theta)
 * window with the submap tile staged in shared memory by TMA. sm_100a only.
 *
 * Replaces the inner loops of ScanMatcherGridSearch::OptimizePose
 * (scan_matcher_grid_search.cpp:118-142) on the integer-shift path: for a fixed
 * candidate angle the score volume is a correlation,
 *     S[iy][ix] = sum_i M[row_i + my[iy]][col_i + mx[ix]],
 * so the map cells a run of consecutive beams needs for ALL (iy, ix) form one
 * compact window: the bounding box of their hit cells widened by the candidate
 * window. k_window_groups cuts every angle's beams into maximal runs whose
 * window fits the shared-memory tile; k_window_tma then, per CTA = (angle,
 * block of candidate rows iy, block of candidate columns ix):
 *   - loads the tile of every run with cp.async.bulk.tensor.2d (TMA, 8-row
 *     boxes of 256 cells; cells outside the map arrive as zeros, which is the
 *     reference's ValueOr for unknown, grid_map.cpp:389-392) into one of two
 *     landing buffers behind an mbarrier, so the next runs' tiles stream in
 *     while this run is scored;
 *   - widens the landed u16 cells to 32-bit words value | known << 20, so that
 *     ONE integer add per candidate and beam accumulates both the value sum
 *     and the known-cell count (folded into 32-bit totals every 16 beams);
 *   - scores from shared memory: warp = kWtRows candidate rows, lane = 32
 *     consecutive candidate columns per 32-column chunk, i.e. every warp-level
 *     load reads 32 consecutive words (128 contiguous bytes, conflict-free), the
 *     per-candidate sums live in registers across all runs;
 *   - windows of 32 k + r columns, r <= 4 (odd windows 2 w + 1 with w a multiple of 16: the
 *     BASELINE 161 x 161 window) would leave 31 of 32 lanes of a last chunk idle: the kRem variant
 *     scores k chunks per lane and gives the r remainder columns of a warp's own rows to its first
 *     3 r lanes (one more, conflict-free load per warp and beam at the 260-word tile pitch instead
 *     of three loads with one useful lane);
 *   - reduces (key, ordinal) with warp shuffles and one atomicMax per warp.
 * No tensor cores: this is a gather-and-sum, not a contraction.
 *
 * Roofline: one shared-memory wavefront per 32 candidate-beam pairs, i.e.
 * 32 x 148 x f_SM gathers/s. */
#pragma once

#include <cuda.h>

#include "csm_kernels.cuh"

namespace csm {

constexpr int kWtMaxWarps = 20;                  /* warps per CTA: 8..20, chosen per window (host) */
constexpr int kWtRows = 3;                       /* candidate rows per warp */
constexpr int kWtChunks = 6;                     /* 32-column chunks per lane (at most) */
constexpr int kWtColsPerCta = 32 * kWtChunks;    /* 192 candidate columns */
constexpr int kWtRemMax = 4;                     /* remainder columns the transposed path takes */
constexpr int kWtPitch = 256;                    /* landing buffer row pitch in cells (TMA box width) */
constexpr int kWtTilePitch = 260;                /* scored tile row pitch in words: a multiple of 4 (16-byte
                                                    stores of the widening) that is not a multiple of 32, so
                                                    that a column of the tile spreads over 8 banks */
constexpr int kWtBoxRows = 8;                    /* rows per TMA box */
constexpr int kWtTileRows = 96;                  /* rows per tile (multiple of kWtBoxRows) */
constexpr size_t kWtStageBytes = (size_t)kWtTileRows * kWtPitch * sizeof(uint16_t);   /* TMA landing buffer */
constexpr size_t kWtTileBytes = (size_t)kWtTileRows * kWtTilePitch * sizeof(uint32_t);   /* scored tile */
constexpr int kWwPitch = 248;                    /* "wide" variant: TMA box width = landing buffer row pitch in
                                                    words; 24 mod 32, so that the rows of a column spread over banks */
constexpr size_t kWwStageBytes = (size_t)kWtTileRows * kWwPitch * sizeof(uint32_t);
constexpr unsigned int kWtKnownBit = 1u << 20;   /* tile word = value | (value != 0) << 20 */
constexpr int kWtFlush = 16;                     /* beams per packed accumulation: 16 * 65535 < 2^20 */
/* dynamic shared memory of k_window_tma for n beams: two landing buffers, the tile, two
 * mbarriers, the projected indices, and slack for the 1024-byte alignment */
__host__ __device__ constexpr size_t wt_smem_bytes(int n)
{
    return 2 * kWtStageBytes + kWtTileBytes + 16 + sizeof(proj_t) * (size_t)n + 1024;
}

/* the same for the wide variant: the landing buffers hold 32-bit words and are scored in place */
__host__ __device__ constexpr size_t wt_smem_bytes_wide(int n)
{
    return 2 * kWwStageBytes + 16 + sizeof(proj_t) * (size_t)n + 1024;
}

/* The map as the wide variant loads it: one 32-bit word value | known << 20 per cell, written once per
 * map (and again whenever level 0 changes) */
__global__ void __launch_bounds__(256)
k_widen_map(const uint4* __restrict__ cells, uint4* __restrict__ wide, size_t n8)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = cells[i];
        const unsigned int w[4] = { v.x, v.y, v.z, v.w };
        unsigned int o[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned int lo = w[k] & 0xffffu, hi = w[k] >> 16;
            o[2 * k] = lo | (lo ? kWtKnownBit : 0u);
            o[2 * k + 1] = hi | (hi ? kWtKnownBit : 0u);
        }
        wide[2 * i] = make_uint4(o[0], o[1], o[2], o[3]);
        wide[2 * i + 1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

/* One run of consecutive beams of one angle and the low corner of its hit cells */
struct WtGroup
{
    int begin, end;      /* beams [begin, end) */
    int minr, minc;
    int hspan, wspan;    /* maxr - minr, maxc - minc */
    int pad0, pad1;
};

/* Greedy runs: extend while bbox_h <= max_h and bbox_w <= max_w. One thread per angle. */
__global__ void __launch_bounds__(128)
k_window_groups(const DevQuery* __restrict__ queries, const proj_t* __restrict__ proj_all,
                WtGroup* __restrict__ groups, int* __restrict__ gcount, int max_h, int max_w)
{
    const DevQuery& Q = queries[0];
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= Q.T)
        return;
    const proj_t* __restrict__ proj = proj_all + Q.proj_off + (size_t)it * Q.n;
    WtGroup* __restrict__ out = groups + (size_t)it * Q.n;
    int ng = 0, begin = 0;
    int minr = 0, maxr = 0, minc = 0, maxc = 0;
    for (int i = 0; i < Q.n; ++i) {
        const proj_t p = proj[i];
        if (i == begin) {
            minr = maxr = p.y; minc = maxc = p.x;
            continue;
        }
        const int nminr = min(minr, (int)p.y), nmaxr = max(maxr, (int)p.y);
        const int nminc = min(minc, (int)p.x), nmaxc = max(maxc, (int)p.x);
        if (nmaxr - nminr > max_h || nmaxc - nminc > max_w) {
            out[ng++] = WtGroup { begin, i, minr, minc, maxr - minr, maxc - minc, 0, 0 };
            begin = i;
            minr = maxr = p.y; minc = maxc = p.x;
        } else {
            minr = nminr; maxr = nmaxr; minc = nminc; maxc = nmaxc;
        }
    }
    if (Q.n > 0)
        out[ng++] = WtGroup { begin, Q.n, minr, minc, maxr - minr, maxc - minc, 0, 0 };
    gcount[it] = ng;
}

__device__ __forceinline__ unsigned int smem_u32(const void* p)
{
    return (unsigned int)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int phase)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}

/* One TMA box: kWtPitch x kWtBoxRows cells of the level-0 map at (col x, row y) -> dst */
__device__ __forceinline__ void tma_load_box(const CUtensorMap* tmap, void* dst, unsigned long long* bar,
                                             int x, int y)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
        :: "r"(smem_u32(dst)), "l"((unsigned long long)tmap), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}

/* shared-memory word at a register address plus an immediate (the scoring loop keeps one address per beam
 * and row; the chunk offsets are immediates) */
template <int kOff>
__device__ __forceinline__ unsigned int lds_u32(unsigned int addr)
{
    unsigned int v;
    asm volatile("ld.shared.u32 %0, [%1+%2];\n" : "=r"(v) : "r"(addr), "n"(kOff));
    return v;
}

struct WtArgs
{
    const WtGroup* groups;
    const int* gcount;
    int dy_span, dx_span;      /* largest my / mx extent of one CTA's block */
    int unit_dx;               /* mx[ix] = mx[0] + ix: consecutive candidate columns are consecutive cells */
    int rows_per_cta;          /* kWtRows * warps of the launch */
};

/* kWide: the tensor map is over the map's 32-bit form (k_widen_map); the tiles land as words and are scored
 * where they land -- no widening pass, two landing buffers of 96 x 248 words instead of two u16 buffers and a
 * tile. */
template <bool kUnitDx, int kChunks, bool kRem, bool kWide>
__global__ void __launch_bounds__(kWtMaxWarps * 32, 1)
k_window_tma(const __grid_constant__ CUtensorMap tmap, const DevQuery* __restrict__ queries,
             const proj_t* __restrict__ proj_all, GridArgs G, WtArgs A)
{
    /* everything lives in dynamic shared memory, aligned by hand: TMA destinations need 128 bytes */
    extern __shared__ unsigned char wt_smem_raw[];
    unsigned char* wt_smem = wt_smem_raw + ((1024u - (smem_u32(wt_smem_raw) & 1023u)) & 1023u);
    constexpr size_t kStage = kWide ? kWwStageBytes : kWtStageBytes;
    constexpr size_t kTile = kWide ? 0 : kWtTileBytes;
    constexpr int kPitchW = kWide ? kWwPitch : kWtTilePitch;        /* row pitch of the scored words */
    constexpr int kAlign = kWide ? 4 : 8;                           /* cells per 16 bytes of a map row */
    uint16_t* stage0 = reinterpret_cast<uint16_t*>(wt_smem);
    uint16_t* stage1 = reinterpret_cast<uint16_t*>(wt_smem + kStage);
    uint32_t* tile = reinterpret_cast<uint32_t*>(wt_smem + 2 * kStage);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(wt_smem + 2 * kStage + kTile);
    proj_t* s_proj = reinterpret_cast<proj_t*>(wt_smem + 2 * kStage + kTile + 16);

    const DevQuery& Q = queries[0];
    const int it = blockIdx.x;
    const int iy0 = blockIdx.y * A.rows_per_cta;
    static_assert(!kRem || kUnitDx, "the remainder path assumes unit column steps");
    const int ix0 = kRem ? 0 : blockIdx.z * (32 * kChunks);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = Q.n;
    const proj_t* __restrict__ proj = proj_all + Q.proj_off + (size_t)it * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        s_proj[i] = proj[i];
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    /* this thread's candidates: rows iy0 + warp * kWtRows + rw, columns ix0 + lane + 32 m */
    const int my0 = G.my[min(iy0, G.ndy - 1)];
    const int mx0 = G.mx[min(ix0, G.ndx - 1)];
    int roff[kWtRows];                  /* tile-local row offset of each of my rows, in cells */
#pragma unroll
    for (int rw = 0; rw < kWtRows; ++rw) {
        const int iy = iy0 + warp * kWtRows + rw;
        roff[rw] = (G.my[min(iy, G.ndy - 1)] - my0) * kPitchW;
    }
    /* Unit column steps: every beam's hit cell becomes the BYTE offset of its word inside its run's tile, once
     * (each beam belongs to one run; the offsets take the place of the projected indices in shared memory), so
     * that the scoring loop is one broadcast load and one add per beam, one add per row, then loads at immediate
     * offsets -- the loop is bound by issue slots, not by shared-memory bandwidth, until it is this lean. */
    int* s_offb = reinterpret_cast<int*>(s_proj);
    if (kUnitDx) {
        const WtGroup* __restrict__ grs = A.groups + (size_t)it * n;
        const int ngr = A.gcount[it];
        for (int g = threadIdx.x; g < ngr; g += blockDim.x) {
            const WtGroup grp = grs[g];
            const int cs = (grp.minc + mx0) & (kAlign - 1);
            for (int i = grp.begin; i < grp.end; ++i) {
                const proj_t p = s_proj[i];
                s_offb[i] = (((int)p.y - grp.minr) * kPitchW + ((int)p.x - grp.minc) + cs) * 4;
            }
        }
        __syncthreads();
    }
    /* remainder columns 32 kChunks .. ndx - 1: the first kWtRows * r lanes of every warp take the
     * warp's own candidate rows, lane = column * kWtRows + row (tile words 260 apart per row and 1
     * apart per column: distinct banks, and every warp carries the same one extra load per beam) */
    bool rem_on = false;
    int rem_off = 0, rem_iy = 0, rem_ix = 0;
    unsigned int accx = 0u, sumx = 0u, cntx = 0u;
    if (kRem) {
        const int nrem = G.ndx - 32 * kChunks;
        const int rw = lane % kWtRows, rj = lane / kWtRows;
        rem_iy = iy0 + warp * kWtRows + rw; rem_ix = 32 * kChunks + rj;
        rem_on = rj < nrem && rem_iy < G.ndy;
        rem_off = rem_on ? (G.my[rem_iy] - my0) * kPitchW + rem_ix : 0;
    }
    int coff[kChunks];
#pragma unroll
    for (int m = 0; m < kChunks; ++m) {
        const int ix = ix0 + lane + 32 * m;
        coff[m] = kUnitDx ? lane + 32 * m : G.mx[min(ix, G.ndx - 1)] - mx0;
    }
    /* per candidate: packed running word (value sum in bits 0..19, known count above) that is
     * folded into the 32-bit sum / count every kWtFlush beams */
    unsigned int acc[kWtRows][kChunks], sum[kWtRows][kChunks], cnt[kWtRows][kChunks];
#pragma unroll
    for (int rw = 0; rw < kWtRows; ++rw)
#pragma unroll
        for (int m = 0; m < kChunks; ++m) { acc[rw][m] = 0u; sum[rw][m] = 0u; cnt[rw][m] = 0u; }
    auto flush = [&]() {
#pragma unroll
        for (int rw = 0; rw < kWtRows; ++rw)
#pragma unroll
            for (int m = 0; m < kChunks; ++m) {
                sum[rw][m] += acc[rw][m] & (kWtKnownBit - 1u);
                cnt[rw][m] += acc[rw][m] >> 20;
                acc[rw][m] = 0u;
            }
        if (kRem) {
            sumx += accx & (kWtKnownBit - 1u);
            cntx += accx >> 20;
            accx = 0u;
        }
    };

    const WtGroup* __restrict__ groups = A.groups + (size_t)it * n;
    const int ng = A.gcount[it];
    auto issue = [&](int g) {
        /* one thread: arm the barrier with the tile's byte count and issue its boxes */
        const WtGroup grp = groups[g];
        const int rows = grp.hspan + A.dy_span + 1;
        const int boxes = (rows + kWtBoxRows - 1) / kWtBoxRows;
        unsigned long long* bar = &s_bar[g & 1];
        uint16_t* dst = (g & 1) ? stage1 : stage0;
        constexpr size_t kBoxBytes = kWide ? (size_t)kWtBoxRows * kWwPitch * sizeof(uint32_t)
                                           : (size_t)kWtBoxRows * kWtPitch * sizeof(uint16_t);
        mbar_expect_tx(bar, (unsigned int)(boxes * kBoxBytes));
        /* the box's first column must sit on a 16-byte boundary of the map row (measured:
         * other values fault); the tile is up to 7 (wide: 3) cells wider for it */
        for (int b = 0; b < boxes; ++b)
            tma_load_box(&tmap, reinterpret_cast<unsigned char*>(dst) + (size_t)b * kBoxBytes, bar,
                         (grp.minc + mx0) & ~(kAlign - 1), grp.minr + my0 + b * kWtBoxRows);
    };
    if (threadIdx.x == 0) {
        if (ng > 0) issue(0);
        if (ng > 1) issue(1);
    }
    int pending = 0;                      /* beams folded into acc since the last flush */
    for (int g = 0; g < ng; ++g) {
        const WtGroup grp = groups[g];
        mbar_wait(&s_bar[g & 1], (unsigned int)((g >> 1) & 1));
        if (!kWide) {
        __syncthreads();                  /* the previous run has been scored: the tile is free */
        {
            /* landing buffer (u16) -> tile (u32 words carrying the known bit), 8 cells per step */
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>((g & 1) ? stage1 : stage0);
            uint4* __restrict__ tile4 = reinterpret_cast<uint4*>(tile);
            const int rows = grp.hspan + A.dy_span + 1;
            const int chunks = rows * (kWtPitch / 8);
            for (int c = threadIdx.x; c < chunks; c += blockDim.x) {
                /* landing row r, 8-cell chunk k -> tile row r (pitch kWtTilePitch words), words 8k .. 8k+7 */
                const int r = c / (kWtPitch / 8), k8 = c - r * (kWtPitch / 8);
                uint4* __restrict__ dst = tile4 + r * (kWtTilePitch / 4) + 2 * k8;
                const uint4 v = src[c];
                const unsigned int w[4] = { v.x, v.y, v.z, v.w };
                unsigned int o[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned int lo = w[k] & 0xffffu, hi = w[k] >> 16;
                    o[2 * k] = lo | (lo ? kWtKnownBit : 0u);
                    o[2 * k + 1] = hi | (hi ? kWtKnownBit : 0u);
                }
                dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
                dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        __syncthreads();                  /* tile ready, landing buffer g & 1 free */
        if (threadIdx.x == 0 && g + 2 < ng)
            issue(g + 2);
        }
        const uint32_t* __restrict__ words = kWide ? reinterpret_cast<const uint32_t*>((g & 1) ? stage1 : stage0) : tile;
        if (kUnitDx) {
            unsigned int wsh = smem_u32(words) + (unsigned int)lane * 4u;
            asm volatile("" : "+r"(wsh));             /* one register for the whole run */
            /* idle lanes read the word lane 0 reads (a broadcast, no extra bank) and mask it away */
            const int rem_off0 = __shfl_sync(0xffffffffu, rem_off, 0);
            const unsigned int rem_delta = (unsigned int)((rem_on ? rem_off : rem_off0) - lane) * 4u;
            const unsigned int rem_mask = rem_on ? 0xffffffffu : 0u;
            for (int i = grp.begin; i < grp.end; ++i) {
                const unsigned int a0 = wsh + (unsigned int)s_offb[i];
#pragma unroll
                for (int rw = 0; rw < kWtRows; ++rw) {
                    const unsigned int a = a0 + (unsigned int)(roff[rw] * 4);
                    if (kChunks > 0) acc[rw][0] += lds_u32<0>(a);
                    if (kChunks > 1) acc[rw][1 % kChunks] += lds_u32<128>(a);
                    if (kChunks > 2) acc[rw][2 % kChunks] += lds_u32<256>(a);
                    if (kChunks > 3) acc[rw][3 % kChunks] += lds_u32<384>(a);
                    if (kChunks > 4) acc[rw][4 % kChunks] += lds_u32<512>(a);
                    if (kChunks > 5) acc[rw][5 % kChunks] += lds_u32<640>(a);
                }
                if (kRem)
                    accx += lds_u32<0>(a0 + rem_delta) & rem_mask;      /* idle lanes read their own word, masked */
                if (++pending == kWtFlush) { flush(); pending = 0; }
            }
        } else {
        const int cshift = (grp.minc + mx0) & (kAlign - 1);      /* column 0 is the 16-byte aligned cell below the window */
        for (int i = grp.begin; i < grp.end; ++i) {
            const proj_t p = s_proj[i];
            const uint32_t* __restrict__ base = words + ((int)p.y - grp.minr) * kPitchW + ((int)p.x - grp.minc) + cshift;
#pragma unroll
            for (int rw = 0; rw < kWtRows; ++rw) {
                const uint32_t* __restrict__ row = base + roff[rw];
#pragma unroll
                for (int m = 0; m < kChunks; ++m)
                    acc[rw][m] += row[coff[m]];
            }
            if (++pending == kWtFlush) { flush(); pending = 0; }
        }
        }
        if (kWide) {
            __syncthreads();              /* every warp has scored this run: its landing buffer is free */
            if (threadIdx.x == 0 && g + 2 < ng)
                issue(g + 2);
        }
    }
    flush();

    /* candidates -> packed best (same decisions as k_grid_window) */
    BestTie best = { 0ull, 0ull };
    const uint16_t* __restrict__ m0 = Q.lvl[0];
    auto consider = [&](int iy, int ix, unsigned int sumv, unsigned int known) {
        const int k = (int)known;
        const long long key = make_key((long long)sumv, k);
        if (k > Q.nk_cut) {
            const int c = key_vs_threshold(key, Q.kthr);
            bool ok = c > 0;
            if (c == 0)
                ok = exact_normalized_score(m0, Q.rows, Q.cols, proj, Q.pst_i, n, G.mx[ix], G.my[iy]) > Q.kthr.thr;
            if (ok) {
                const unsigned long long ord = grid_ordinal(G, iy, ix, it);
                best_merge(best, pack_best(key, (kOrdMask - 1ull) - ord), 0ull);
            }
        }
    };
#pragma unroll
    for (int rw = 0; rw < kWtRows; ++rw) {
        const int iy = iy0 + warp * kWtRows + rw;
#pragma unroll
        for (int m = 0; m < kChunks; ++m) {
            const int ix = ix0 + lane + 32 * m;
            if (iy >= G.ndy || ix >= G.ndx)
                continue;
            consider(iy, ix, sum[rw][m], cnt[rw][m]);
        }
    }
    if (kRem && rem_on)
        consider(rem_iy, rem_ix, sumx, cntx);
    block_best_commit(best, G.best, G.tiekey);
}

} /* namespace csm */
