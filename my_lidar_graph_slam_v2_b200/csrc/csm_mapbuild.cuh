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
 *   k_map_rays    one thread per beam walks its ray (the reference's integer algorithm) and writes one
 *                 64-bit event per touched cell: cell << 32 | beam order << 1 | hit;
 *   radix sort    of the events (CUB, part of the CUDA toolkit): per cell, in beam order;
 *   k_map_apply   one thread per cell run applies the table chain to the cell and marks its 16 x 16
 *                 block allocated (GridMap::UpdateOddsUnchecked allocates on first write,
 *                 grid_map.cpp:649-661).
 * The map stays where the matchers read it (level 0 of a map slot): no host round trip per scan.
 */
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "csm_device.cuh"
#include "csm_b200.h"

namespace csm {

constexpr unsigned long long kMapEventNone = ~0ull;

struct MapRaysArgs
{
    const csm_ray* rays;
    const unsigned int* offset;      /* first event slot of every ray, n + 1 entries */
    unsigned long long* events;
    int n, scale, rows, cols;
    int* error;                      /* set when a ray leaves the map or overruns its slots */
};

struct RayEmitter
{
    unsigned long long* out;
    unsigned int cap, used;
    int rows, cols, order;
    int end_x, end_y;
    int bad;
    __device__ __forceinline__ void cell(int x, int y)
    {
        if (x == end_x && y == end_y)
            return;                                  /* the end cell is not a miss (grid_map_builder.cpp:904-910) */
        if ((unsigned)x >= (unsigned)cols || (unsigned)y >= (unsigned)rows || used >= cap) { bad = 1; return; }
        out[used++] = ((unsigned long long)((unsigned)y * (unsigned)cols + (unsigned)x) << 32) |
                      ((unsigned long long)(unsigned)order << 1);
    }
};

/* BresenhamScaled (bresenham.cpp:59-236), cell by cell, without the list: consecutive duplicates are
 * dropped like there (the emitter sees every full-pixel cell of the ray once). */
__device__ void walk_ray_scaled(int sx, int sy, int ex, int ey, int scale, RayEmitter& E)
{
    if (sx > ex) { int t = sx; sx = ex; ex = t; t = sy; sy = ey; ey = t; }     /* ordered by x */
    const int start_x = sx / scale, start_y = sy / scale, end_x = ex / scale, end_y = ey / scale;
    int last_x = start_x, last_y = start_y;
    bool have = false;
    auto push = [&](int x, int y) {
        if (have && x == last_x && y == last_y) return;
        have = true; last_x = x; last_y = y;
        E.cell(x, y);
    };
    if (start_x == end_x) {
        const int y0 = min(start_y, end_y), y1 = max(start_y, end_y);
        for (int y = y0; y <= y1; ++y) push(start_x, y);
        return;
    }
    const long long dx = (long long)ex - sx, dy = (long long)ey - sy;
    const long long denominator = 2ll * scale * dx;
    int cx = start_x, cy = start_y;
    push(cx, cy);
    long long sub_y = (2ll * (sy % scale) + 1) * dx;
    const int first_pixel = 2 * scale - (2 * (sx % scale) + 1);
    const int last_pixel = 2 * (ex % scale) + 1;
    const int end_full_x = max(start_x, end_x);
    sub_y += dy * first_pixel;
    if (dy > 0) {
        while (true) {
            push(cx, cy);
            while (sub_y > denominator) { sub_y -= denominator; ++cy; push(cx, cy); }
            if (sub_y == denominator) { sub_y -= denominator; ++cy; }
            ++cx;
            if (cx == end_full_x) break;
            sub_y += 2 * dy * scale;
        }
        sub_y += dy * last_pixel;
        push(cx, cy);
        while (sub_y > denominator) { sub_y -= denominator; ++cy; push(cx, cy); }
    } else {
        while (true) {
            push(cx, cy);
            while (sub_y < 0) { sub_y += denominator; --cy; push(cx, cy); }
            if (sub_y == 0) { sub_y += denominator; --cy; }
            ++cx;
            if (cx == end_full_x) break;
            sub_y += 2 * dy * scale;
        }
        sub_y += dy * last_pixel;
        push(cx, cy);
        while (sub_y < 0) { sub_y += denominator; --cy; push(cx, cy); }
    }
}

__global__ void __launch_bounds__(128)
k_map_rays(MapRaysArgs A)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    const csm_ray R = A.rays[i];
    RayEmitter E;
    E.out = A.events + A.offset[i];
    E.cap = A.offset[i + 1] - A.offset[i];
    E.used = 0;
    E.rows = A.rows; E.cols = A.cols; E.order = R.order;
    E.end_x = R.end_x / A.scale; E.end_y = R.end_y / A.scale;
    E.bad = 0;
    walk_ray_scaled(R.start_x, R.start_y, R.end_x, R.end_y, A.scale, E);
    /* the hit cell (PositionToIndex of the hit point in the unscaled geometry) */
    if ((unsigned)R.hit_col < (unsigned)A.cols && (unsigned)R.hit_row < (unsigned)A.rows && E.used < E.cap)
        E.out[E.used++] = ((unsigned long long)((unsigned)R.hit_row * (unsigned)A.cols + (unsigned)R.hit_col) << 32) |
                          ((unsigned long long)(unsigned)R.order << 1) | 1ull;
    else
        E.bad = 1;
    for (unsigned int k = E.used; k < E.cap; ++k)
        E.out[k] = kMapEventNone;                     /* sorts behind every real event */
    if (E.bad)
        *A.error = 1;
}

struct MapApplyArgs
{
    const unsigned long long* events;     /* sorted */
    unsigned int n;
    uint16_t* map;
    unsigned char* alloc;                 /* one byte per 16 x 16 block */
    const uint16_t* lut_miss;
    const uint16_t* lut_hit;
    int cols, log2bs, block_cols;
};

__global__ void __launch_bounds__(256)
k_map_apply(MapApplyArgs A)
{
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n)
        return;
    const unsigned long long e = A.events[i];
    if (e == kMapEventNone)
        return;
    const unsigned int cell = (unsigned int)(e >> 32);
    if (i > 0 && (unsigned int)(A.events[i - 1] >> 32) == cell)
        return;                                       /* not the first event of its cell */
    unsigned int v = A.map[cell];
    for (unsigned int k = i; k < A.n; ++k) {
        const unsigned long long ek = A.events[k];
        if ((unsigned int)(ek >> 32) != cell)
            break;
        v = (ek & 1ull) ? __ldg(A.lut_hit + v) : __ldg(A.lut_miss + v);
    }
    A.map[cell] = (uint16_t)v;
    const unsigned int row = cell / (unsigned)A.cols, col = cell - row * (unsigned)A.cols;
    A.alloc[(row >> A.log2bs) * (unsigned)A.block_cols + (col >> A.log2bs)] = 1;
}

/* GridMap::Resize (grid_map.cpp:842-888): the blocks that overlap old and new extent move, every other
 * cell is unknown. dst (new_rows x new_cols) from src (old extent), src cell (r, c) = dst cell
 * (r - row_min, c - col_min); the block allocation bytes move the same way. */
struct MapMoveArgs
{
    const uint16_t* src; uint16_t* dst;
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
