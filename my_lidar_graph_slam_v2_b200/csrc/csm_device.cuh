/* csm_device.cuh -- device-side data structures and helpers shared by the
 * kernels of libcsm_b200.so. sm_100a only.
 *
 * Score arithmetic (DESIGN.md "Exact integer key"):
 *   the reference accumulates sum_i p(v_i) over known cells in double
 *   (score_function_pixel_accurate.cpp:21-57) with
 *   p(v) = 0.001 + 0.998 * (v - 1) / 65534 (grid_values.hpp:26-36), so
 *   65534000 * sum = 998 * sumV + 64536 * nKnown =: key  (exact integers).
 *   Kernels carry (sumV, nKnown) per candidate, order candidates by `key`
 *   and only evaluate doubles for the final result and for threshold
 *   comparisons that fall inside the rounding guard band.
 */
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace csm {

constexpr int kMaxLevels = 8;            /* hmax <= 7 */
constexpr int kOrdBits = 26;             /* candidate ordinal bits in a packed best word */
constexpr unsigned long long kOrdMask = (1ull << kOrdBits) - 1ull;
constexpr int kMaxBeams = 4096;

/* projected cell index of one beam: (col, row), saturated to +-30000 */
typedef short2 proj_t;

/* Threshold on the normalized score expressed on integer keys:
 *   key >= pass_min  -> reference comparison `score > thr` is true
 *   key <= fail_max  -> false
 *   otherwise        -> inside the guard band: decide with the exact
 *                       sequential double sum (exact_normalized_score) */
struct KeyThreshold
{
    long long fail_max;
    long long pass_min;
    double    thr;
};

/* One query as the kernels see it (device pointers only) */
struct DevQuery
{
    const uint16_t* lvl[kMaxLevels];   /* pyramid levels, lvl[0] = uploaded grid */
    const uint16_t* coarse;            /* RT: sliding max with win = low_res */
    int rows, cols;
    double res, offx, offy;
    double inv_res;                    /* 1 / res (projection) */
    double sx, sy;                     /* sensor position (map-local) */
    const double* thetas;              /* grid search: T candidate sensor angles (host doubles); else null
                                          and angle t is theta0 + (t - tcenter) * step_t, evaluated on the
                                          device with the same two IEEE operations the reference uses */
    double theta0, step_t;
    int tcenter;
    int pad2;
    const double* angles;              /* N beam angles */
    const double* ranges;              /* N beam ranges */
    const double2* beam_trig;          /* N x (cos a_i, sin a_i) */
    int n;                             /* beams */
    int T;                             /* candidate angles */
    int winx, winy;                    /* half windows (cells) */
    int nrx, nry;                      /* B&B: number of root nodes along x / y */
    int lx, ly;                        /* leaf lattice extent along x / y */
    long long proj_off;                /* offset of this query in the projection buffer */
    int pst_t, pst_i;                  /* element strides of proj along angle / beam:
                                          (n, 1) angle-major for RT / grid search,
                                          (1, T) beam-major for branch-and-bound */
    double margin;                     /* FP guard band in cells (projection) */
    KeyThreshold kthr;
    int nk_cut;                        /* known test passes iff nKnown > nk_cut */
    int pad;
    /* refinement (csm_refine.cuh) */
    const unsigned char* alloc;        /* one byte per 2^k x 2^k block: allocated in the reference's sense */
    int alloc_log2bs, alloc_bcols;     /* k, blocks per row */
    double stepx, stepy;               /* search steps along x / y (the coarse result is an index) */
    /* bound levels of the branch-and-bound sweep (csm_bounds.cuh): bl[hc] = level hc (u8, tiled, padded),
     * bl_tpr[hc] its tiles per row; null when the sweep reads the u16 levels */
    const unsigned char* bl[kMaxLevels];
    int bl_tpr[kMaxLevels];
    const int* low_margin;             /* smallest min(row, col) of a known cell near the low edges of the map
                                          (k_low_margin); at least edge_need: a CSM_FLAG_EDGE is withdrawn */
    int edge_need, pad4;               /* 2^hmax */
    int pquad;                         /* 1: proj is stored in chunks of 16 beams, [n / 16][tp][4][4] (see proj_index) */
    int tp;                            /* quad layout: angles per beam group, T rounded up to a multiple of 8 */
};

/* Position of (angle t, beam i) in a query's block of the projection buffer. Strided layouts:
 * angle-major (pst_t = n, pst_i = 1) for the real-time and grid-search matchers, beam-major
 * (pst_t = 1, pst_i = T). Chunk layout (branch-and-bound over bound levels): [n / 16][tp][4][4], tp = T
 * rounded up to a multiple of 8: beam i of angle t sits in chunk i / 16 at [t][i % 4][(i / 4) % 4]. A lane of
 * the group sweep (one angle, beams i % 4 == part) reads its four beams of a chunk as ONE 16-byte word, the
 * 8 angles x 4 parts of a warp read 512 contiguous bytes, and at each of the four steps the warp's lanes
 * hold 8 adjacent angles x 4 ADJACENT beams. */
__device__ __forceinline__ size_t proj_index(const DevQuery& Q, int t, int i)
{
    return Q.pquad ? ((((size_t)(i >> 4) * (size_t)Q.tp + (size_t)t) << 4) + (size_t)(((i & 3) << 2) | ((i >> 2) & 3)))
                   : (size_t)t * (size_t)Q.pst_t + (size_t)i * (size_t)Q.pst_i;
}

__host__ __device__ __forceinline__ long long make_key(long long sumv, int nk)
{
    return 998ll * sumv + 64536ll * (long long)nk;
}

__device__ __forceinline__ unsigned int ld_cell(const uint16_t* __restrict__ m,
                                                int rows, int cols, int r, int c)
{
    /* Outside the map -> unknown (0), grid_map.cpp:389-392 */
    if ((unsigned)r < (unsigned)rows && (unsigned)c < (unsigned)cols)
        return (unsigned int)__ldg(m + (size_t)r * (size_t)cols + (size_t)c);
    return 0u;
}

/* Layout of the precomputed levels: dense row-major u16, like level 0. An 8x8-cell tiled
 * layout (one 128-byte line per tile, zero border) was built and measured in round 1
 * (profiles/r1_layout_ab.txt): it cut the branch-and-bound sweep by 16 % (534 -> 451 us, fewer
 * lines per divergent gather) but slowed the streaming pyramid builder from 236 to 400 us
 * (16-byte pieces of 8 lines per store instead of one full line), so it was dropped. */

/* Same as ld_cell, branch-free: an out-of-map read loads cell 0 and is masked, so that a lane
 * can keep many of these loads in flight */
__device__ __forceinline__ unsigned int ld_cell_nb(const uint16_t* __restrict__ m, int rows, int cols,
                                                   int r, int c)
{
    const bool ok = (unsigned)r < (unsigned)rows && (unsigned)c < (unsigned)cols;
    const unsigned int v = (unsigned int)__ldg(m + (ok ? (unsigned)r * (unsigned)cols + (unsigned)c : 0u));
    return ok ? v : 0u;
}

/* p(v), grid_values.hpp:26-36, same operation order, no FMA */
__device__ __forceinline__ double value_to_probability(unsigned int v)
{
    const double pmin = 1e-3;
    const double pmax = 1.0 - 1e-3;
    const double span = pmax - pmin;
    return __dadd_rn(pmin, __ddiv_rn(__dmul_rn(span, (double)((int)v - 1)), 65534.0));
}

/* Sequential double sum in scan order: the reference's own arithmetic
 * (scan_matcher_correlative.cpp:308-335). One thread. */
__device__ double exact_normalized_score(const uint16_t* __restrict__ m, int rows, int cols,
                                         const proj_t* __restrict__ proj, int stride, int n,
                                         int ox, int oy)
{
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const proj_t p = proj[(size_t)i * stride];
        const unsigned int v = ld_cell(m, rows, cols, p.y + oy, p.x + ox);
        if (v != 0u)
            sum = __dadd_rn(sum, value_to_probability(v));
    }
    return __ddiv_rn(sum, (double)n);
}

/* The same for angle t of a query in whatever layout its projection has */
__device__ double exact_normalized_score_q(const DevQuery& Q, const uint16_t* __restrict__ m,
                                           const proj_t* __restrict__ proj_q, int t, int ox, int oy)
{
    double sum = 0.0;
    for (int i = 0; i < Q.n; ++i) {
        const proj_t p = proj_q[proj_index(Q, t, i)];
        const unsigned int v = ld_cell(m, Q.rows, Q.cols, p.y + oy, p.x + ox);
        if (v != 0u)
            sum = __dadd_rn(sum, value_to_probability(v));
    }
    return __ddiv_rn(sum, (double)Q.n);
}

/* -1 fail, +1 pass, 0 guard band */
__device__ __forceinline__ int key_vs_threshold(long long key, const KeyThreshold& t)
{
    if (key >= t.pass_min) return 1;
    if (key <= t.fail_max) return -1;
    return 0;
}

__device__ __forceinline__ unsigned long long pack_best(long long key, unsigned long long ordfield)
{
    return ((unsigned long long)key << kOrdBits) | ordfield;
}

__device__ __forceinline__ int warp_sum(int v)
{
    return (int)__reduce_add_sync(0xffffffffu, (unsigned)v);
}

/* B&B frontier node: q(16) | t(16) | xi(16) | yi(16), xi = x + winx >= 0 */
__device__ __forceinline__ unsigned long long pack_node(int q, int t, int xi, int yi)
{
    return ((unsigned long long)(unsigned)q << 48) | ((unsigned long long)(unsigned)t << 32) |
           ((unsigned long long)(unsigned)xi << 16) | (unsigned long long)(unsigned)yi;
}

__device__ __forceinline__ void unpack_node(unsigned long long n, int& q, int& t, int& xi, int& yi)
{
    q = (int)(n >> 48);
    t = (int)((n >> 32) & 0xffffull);
    xi = (int)((n >> 16) & 0xffffull);
    yi = (int)(n & 0xffffull);
}

} /* namespace csm */
