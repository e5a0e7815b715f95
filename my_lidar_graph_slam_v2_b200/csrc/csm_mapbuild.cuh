/* csm_mapbuild.cuh -- occupancy-map construction on the device.
 * sm_100a only; compiled into libcsm_b200.so by csm_b200.cu.
 *
 * Replaces the ray casting of GridMapBuilder::UpdateGridMap / ConstructMapFromScans
 * (mapping/grid_map_builder.cpp:390-494, 561-695): for every beam the cells between sensor and hit point
 * (BresenhamScaled at 1/100 cell, bresenham.cpp:59-236, minus the cell of the end point,
 * grid_map_builder.cpp:891-911) take a "miss" update, the hit cell a "hit" update
 * (GridBinaryBayes::UpdateOddsUnchecked, grid_binary_bayes.cpp:302-321). That update is a function of
 * the cell's u16 value alone, value' = T_miss[value] or T_hit[value]: two 65536-entry tables the host
 * evaluates once with the reference's own double arithmetic. It does not commute (the value is
 * truncated and clamped after every update), so a cell must see its updates in the reference's order:
 * scan by scan, beam by beam. The device does it in three steps:
 *   k_map_rays    one CTA per beam, one thread per grid column of its ray (the reference's integer
 *                 algorithm in closed form per column) writes one 64-bit event per touched cell: cell << shift | beam order << 1 | hit, shift = 1 + the
 *                 bits the beam count needs, so that the sort only runs over bits that vary;
 *   radix sort    of the events (CUB, part of the CUDA toolkit): per cell, in beam order;
 *   k_map_apply   one thread per cell run (its warp for the long runs next to a sensor) applies the table
 *                 chain to the cell and marks its 16 x 16 block allocated (GridMap::UpdateOddsUnchecked allocates on first write,
 *                 grid_map.cpp:649-661).
 * The map stays where the matchers read it (level 0 of a map slot): no host round trip per scan.
 */
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "csm_device.cuh"
#include "csm_b200.h"

namespace csm {

constexpr unsigned long long kMapEventNone = ~0ull;
constexpr int kMapRayThreads = 64;       /* threads of one ray's CTA, one grid column each (strided) */
constexpr int kMapPowTables = 10;        /* T^(2^j), j < 10, per kind of update */
constexpr int kMapShortRun = 16;         /* updates one thread applies by itself before its warp helps */

struct MapRaysArgs
{
    const csm_ray* rays;
    const unsigned int* offset;      /* first event slot of every ray, n + 1 entries */
    unsigned long long* events;      /* preset to kMapEventNone */
    int n, scale, rows, cols;
    int shift;                       /* event = cell << shift | order << 1 | hit */
    int* error;                      /* set when a ray leaves the map or overruns its slots */
};

__device__ __forceinline__ long long floor_div(long long a, long long b)      /* b > 0 */
{
    long long q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}
__device__ __forceinline__ long long ceil_div(long long a, long long b)       /* b > 0 */
{
    long long q = a / b;
    return (a % b != 0 && a > 0) ? q + 1 : q;
}

/* BresenhamScaled (bresenham.cpp:59-236) column by column. The reference walks the ray from the end with
 * the smaller x, carrying sub_y (the ray's height at the right border of the current column, in units of
 * 1 / denominator of a cell, wrapped into (0, denominator]) and stepping cy whenever it wraps. Unwrapped,
 * the height at the right border of column k is U_k = U_0 + k * 2 * dy * scale, an integer known without
 * walking, and the walk's cells in column k are
 *   dy > 0 :  start_y + floor(U_{k-1} / den)    ..  start_y + ceil(U_k / den) - 1       (upwards)
 *   dy <= 0:  start_y + ceil(U_{k-1} / den) - 1 ..  start_y + floor(U_k / den)          (downwards)
 * (column 0 starts at start_y; the last column ends at the height of the end point, U_{K-1} + dy * lastPixel).
 * The floor/ceil pair reproduces the reference's two cases at a border: it pushes a cell while sub_y is
 * strictly beyond the border (:152-157, :196-201) and steps WITHOUT a push when the ray meets the corner
 * exactly (:159-162, :203-206). Every column is independent, so one thread takes one column; a ray's
 * cells are distinct, and all its events carry the same order, so their order among themselves is free.
 * Column k writes from slot k + |first cell's y - start_y| of the ray's slots: at most one slot per cell is
 * skipped (at a corner), none is shared. */
__global__ void __launch_bounds__(kMapRayThreads)
k_map_rays(MapRaysArgs A)
{
    const int i = blockIdx.x;
    const csm_ray R = A.rays[i];
    int sx = R.start_x, sy = R.start_y, ex = R.end_x, ey = R.end_y;
    const int scale = A.scale;
    const int hit_x = ex / scale, hit_y = ey / scale;            /* the end index, not a miss (grid_map_builder.cpp:904-910) */
    if (sx > ex) { int t = sx; sx = ex; ex = t; t = sy; sy = ey; ey = t; }     /* ordered by x (:69-72) */
    const int start_x = sx / scale, start_y = sy / scale, end_x = ex / scale, end_y = ey / scale;
    unsigned long long* out = A.events + A.offset[i];
    const unsigned int cap = A.offset[i + 1] - A.offset[i];
    const unsigned long long tag = (unsigned long long)(unsigned)R.order << 1;
    bool bad = false;
    auto emit = [&](unsigned int slot, int x, int y) {
        if (x == hit_x && y == hit_y)
            return;
        if ((unsigned)x >= (unsigned)A.cols || (unsigned)y >= (unsigned)A.rows || slot + 1 >= cap) { bad = true; return; }
        out[slot] = ((unsigned long long)((unsigned)y * (unsigned)A.cols + (unsigned)x) << A.shift) | tag;
    };
    if (start_x == end_x) {
        /* one column of full cells (:88-102) */
        const int y0 = min(start_y, end_y), y1 = max(start_y, end_y);
        for (int y = y0 + (int)threadIdx.x; y <= y1; y += kMapRayThreads)
            emit((unsigned)(y - y0), start_x, y);
    } else {
        const long long dx = (long long)ex - sx, dy = (long long)ey - sy;
        const long long den = 2ll * scale * dx;
        const long long u0 = (2ll * (sy % scale) + 1) * dx + dy * (2 * scale - (2 * (sx % scale) + 1));
        const long long per_col = 2ll * dy * scale;
        const int K = end_x - start_x;
        for (int k = threadIdx.x; k <= K; k += kMapRayThreads) {
            const long long u_prev = u0 + (long long)(k - 1) * per_col;
            const long long u_here = (k < K) ? u_prev + per_col : u_prev + dy * (2 * (ex % scale) + 1);
            int y_first, y_last, step;
            if (dy > 0) {
                y_first = (k == 0) ? start_y : start_y + (int)floor_div(u_prev, den);
                y_last = start_y + (int)ceil_div(u_here, den) - 1;
                step = 1;
            } else {
                y_first = (k == 0) ? start_y : start_y + (int)ceil_div(u_prev, den) - 1;
                y_last = start_y + (int)floor_div(u_here, den);
                step = -1;
            }
            unsigned int slot = (unsigned)k + (unsigned)abs(y_first - start_y);
            for (int y = y_first; step > 0 ? y <= y_last : y >= y_last; y += step)
                emit(slot++, start_x + k, y);
        }
    }
    if (threadIdx.x == 0) {
        /* the hit cell (PositionToIndex of the hit point in the unscaled geometry), last slot of the ray */
        if ((unsigned)R.hit_col < (unsigned)A.cols && (unsigned)R.hit_row < (unsigned)A.rows && cap > 0)
            out[cap - 1] = ((unsigned long long)((unsigned)R.hit_row * (unsigned)A.cols + (unsigned)R.hit_col) << A.shift) |
                           tag | 1ull;
        else
            bad = true;
    }
    if (bad)
        *A.error = 1;
}

struct MapApplyArgs
{
    const unsigned long long* events;     /* sorted */
    unsigned int n;
    uint16_t* map;                        /* the true values (what the next update continues from) */
    uint16_t* view;                       /* what the matchers read: 65535 as unknown when saturated_unknown */
    int saturated_unknown;
    unsigned char* alloc;                 /* one byte per 16 x 16 block */
    const uint16_t* lut;                  /* [kind: miss, hit][j < kMapPowTables][65536]: j-th table = 2^j updates */
    int cols, log2bs, block_cols, shift;
};

/* `count` updates of one kind in a row: binary decomposition over the power tables */
__device__ __forceinline__ unsigned int map_apply_run(const uint16_t* lut, unsigned int v, unsigned int kind, unsigned int count)
{
    const uint16_t* t = lut + (size_t)kind * kMapPowTables * 65536;
    while (count >> kMapPowTables) {
        v = __ldg(t + (size_t)(kMapPowTables - 1) * 65536 + v);
        count -= 1u << (kMapPowTables - 1);
    }
    #pragma unroll 1
    for (int j = 0; count; ++j, count >>= 1)
        if (count & 1u)
            v = __ldg(t + (size_t)j * 65536 + v);
    return v;
}

/* One thread per event; the thread of a cell's FIRST event applies the cell's chain. Most chains are a
 * handful of updates. The cells around a sensor take one miss from every beam of its scan: hundreds in a
 * row, each a dependent table look-up. Those chains go to the whole warp: 32 events per coalesced load,
 * runs of equal kind found with ballots and bit scans, a run of k updates applied in popcount(k) look-ups
 * of the power tables. */
__global__ void __launch_bounds__(256)
k_map_apply(MapApplyArgs A)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int lane = threadIdx.x & 31u;
    unsigned long long e = kMapEventNone;
    bool first = false;
    if (i < A.n) {
        e = A.events[i];
        first = e != kMapEventNone && (i == 0 || (A.events[i - 1] >> A.shift) != (e >> A.shift));
    }
    const unsigned int cell = (unsigned int)(e >> A.shift);
    unsigned int v = 0, pos = i;
    bool more = false;
    if (first) {
        v = A.map[cell];
        #pragma unroll 1
        for (int k = 0; k < kMapShortRun; ++k) {
            const unsigned long long ek = (pos < A.n) ? A.events[pos] : kMapEventNone;
            if (ek == kMapEventNone || (unsigned int)(ek >> A.shift) != cell)
                break;
            v = __ldg(A.lut + (size_t)(ek & 1ull) * kMapPowTables * 65536 + v);
            ++pos;
        }
        const unsigned long long en = (pos < A.n) ? A.events[pos] : kMapEventNone;
        more = en != kMapEventNone && (unsigned int)(en >> A.shift) == cell;
    }
    /* long chains, one at a time, the warp on each */
    unsigned int todo = __ballot_sync(0xffffffffu, more);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const unsigned int c = __shfl_sync(0xffffffffu, cell, src);
        unsigned int p = __shfl_sync(0xffffffffu, pos, src);
        unsigned int w = __shfl_sync(0xffffffffu, v, src);
        unsigned int run_kind = 0, run_len = 0;
        unsigned long long en = (p + lane < A.n) ? A.events[p + lane] : kMapEventNone;
        while (true) {
            const unsigned long long ec = en;
            p += 32;
            en = (p + lane < A.n) ? A.events[p + lane] : kMapEventNone;       /* next chunk in flight */
            const unsigned int same = __ballot_sync(0xffffffffu, ec != kMapEventNone && (unsigned int)(ec >> A.shift) == c);
            const unsigned int hits = __ballot_sync(0xffffffffu, (unsigned int)(ec & 1ull));
            const int valid = (same == 0xffffffffu) ? 32 : __ffs(~same) - 1;      /* sorted: a prefix of the lanes */
            int b = 0;
            while (b < valid) {
                const unsigned int kind = (hits >> b) & 1u;
                const unsigned int differ = (kind ? ~hits : hits) >> b;
                int len = differ ? __ffs(differ) - 1 : 32 - b;
                len = min(len, valid - b);
                if (kind == run_kind)
                    run_len += len;
                else {
                    w = map_apply_run(A.lut, w, run_kind, run_len);
                    run_kind = kind; run_len = len;
                }
                b += len;
            }
            if (valid < 32)
                break;
        }
        w = map_apply_run(A.lut, w, run_kind, run_len);
        if ((int)lane == src)
            v = w;
    }
    if (first) {
        A.map[cell] = (uint16_t)v;
        A.view[cell] = (A.saturated_unknown && v == 65535u) ? (uint16_t)0 : (uint16_t)v;
        const unsigned int row = cell / (unsigned)A.cols, col = cell - row * (unsigned)A.cols;
        A.alloc[(row >> A.log2bs) * (unsigned)A.block_cols + (col >> A.log2bs)] = 1;
    }
}

/* GridMap::Resize (grid_map.cpp:842-888): the blocks that overlap old and new extent move, every other
 * cell is unknown. dst (new_rows x new_cols) from src (old extent), src cell (r, c) = dst cell
 * (r - row_min, c - col_min); the block allocation bytes move the same way. */
struct MapMoveArgs
{
    const uint16_t* src; uint16_t* dst;       /* true values */
    uint16_t* dst_view;                        /* the matchers' view of dst */
    int saturated_unknown;
    const unsigned char* src_alloc; unsigned char* dst_alloc;
    int src_rows, src_cols, dst_rows, dst_cols, row_min, col_min, log2bs;
};

__global__ void __launch_bounds__(256)
k_map_move(MapMoveArgs A)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < A.dst_cols && r < A.dst_rows) {
        const int sr = r + A.row_min, sc = c + A.col_min;
        uint16_t v = 0;
        if ((unsigned)sr < (unsigned)A.src_rows && (unsigned)sc < (unsigned)A.src_cols)
            v = A.src[(size_t)sr * A.src_cols + sc];
        A.dst[(size_t)r * A.dst_cols + c] = v;
        A.dst_view[(size_t)r * A.dst_cols + c] = (A.saturated_unknown && v == 65535) ? (uint16_t)0 : v;
    }
    const int bs = 1 << A.log2bs;
    const int dbr = A.dst_rows >> A.log2bs, dbc = A.dst_cols >> A.log2bs;
    if (r < dbr && c < dbc) {
        const int sbr = r + (A.row_min >> A.log2bs), sbc = c + (A.col_min >> A.log2bs);
        unsigned char a = 0;
        if ((unsigned)sbr < (unsigned)(A.src_rows / bs) && (unsigned)sbc < (unsigned)(A.src_cols / bs))
            a = A.src_alloc[sbr * (A.src_cols / bs) + sbc];
        A.dst_alloc[r * dbc + c] = a;
    }
}

} /* namespace csm */
