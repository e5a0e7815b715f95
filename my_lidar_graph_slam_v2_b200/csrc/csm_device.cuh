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
};

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

/* Tiled layout of the precomputed levels (h >= 1): 8x8-cell tiles of 128
 * bytes (one L1/L2 line), row-major inside a tile, with a border of one
 * all-zero tile on every side; rows and cols are padded up to multiples of 8
 * and the padding cells are zero too.
 *  - Branch-and-bound gathers of one warp fall on a compact 2-D patch of cells
 *    (a few adjacent beams x a few adjacent angles): in tiles the patch touches
 *    ~2.7 lines instead of ~7 in row-major order (measured on the cfg3
 *    workload).
 *  - Out-of-map reads (the reference's ValueOr -> 0, grid_map.cpp:389-392) need
 *    no predicate: coordinates are clamped into the zero border.
 *  - The index is separable and cheap: with rp = r + 8, cp = c + 8,
 *    index = [rp * 8 + (rp >> 3) * (64 * tiles_per_row - 64)] + [cp + (cp >> 3) * 56]. */
/* Measured on B200 (cfg3, 256 queries, profiles/r1_layout_ab.txt): tiles cut the B&B
 * sweep from 534 to 451 us but the streaming pyramid builder, whose warps then write
 * 16-byte pieces of 8 different lines instead of one full line, goes from 236 to 400 us.
 * Row-major wins overall, so it is the default; -DCSM_TILED=1 builds the tiled variant. */
#ifndef CSM_TILED
#define CSM_TILED 0
#endif
__host__ __device__ __forceinline__ int padded_tiles_per_row(int cols) { return ((cols + 7) >> 3) + 2; }
__host__ __device__ __forceinline__ int padded_tile_rows(int rows) { return ((rows + 7) >> 3) + 2; }
__host__ __device__ __forceinline__ size_t tiled_cells(int rows, int cols)
{
#if CSM_TILED
    return (size_t)padded_tile_rows(rows) * (size_t)padded_tiles_per_row(cols) * 64u;
#else
    return (size_t)rows * (size_t)cols;
#endif
}
/* row stride term: 64 * tiles_per_row - 64 */
__host__ __device__ __forceinline__ int tiled_rstride(int cols) { return (padded_tiles_per_row(cols) << 6) - 64; }
__host__ __device__ __forceinline__ unsigned int tiled_row_p(int rp, int rstride)
{
    return (unsigned)rp * 8u + (unsigned)(rp >> 3) * (unsigned)rstride;
}
__host__ __device__ __forceinline__ unsigned int tiled_col_p(int cp) { return (unsigned)cp + (unsigned)(cp >> 3) * 56u; }
/* in-map cell (0 <= r < rows, 0 <= c < cols) */
__host__ __device__ __forceinline__ size_t tiled_index(int r, int c, int cols)
{
#if CSM_TILED
    return (size_t)tiled_row_p(r + 8, tiled_rstride(cols)) + (size_t)tiled_col_p(c + 8);
#else
    return (size_t)r * (size_t)cols + (size_t)c;
#endif
}
/* largest padded coordinate: clamping r + 8 into [0, tiled_rmax] lands out-of-map reads in the zero border */
__host__ __device__ __forceinline__ int tiled_rmax(int rows) { return ((rows + 7) & ~7) + 15; }

/* Cell of a precomputed level (tiled, h >= 1) or of the level-0 grid (row-major) */
__device__ __forceinline__ unsigned int ld_level(const uint16_t* __restrict__ m, int rows, int cols,
                                                 bool tiled, int r, int c)
{
    if (tiled && CSM_TILED) {
        const int rp = min(max(r + 8, 0), tiled_rmax(rows));
        const int cp = min(max(c + 8, 0), tiled_rmax(cols));
        return (unsigned int)__ldg(m + (tiled_row_p(rp, tiled_rstride(cols)) + tiled_col_p(cp)));
    }
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
                                         int ox, int oy, bool tiled = false)
{
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        const proj_t p = proj[(size_t)i * stride];
        const unsigned int v = ld_level(m, rows, cols, tiled, p.y + oy, p.x + ox);
        if (v != 0u)
            sum = __dadd_rn(sum, value_to_probability(v));
    }
    return __ddiv_rn(sum, (double)n);
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
