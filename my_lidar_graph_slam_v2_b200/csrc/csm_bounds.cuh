/* csm_bounds.cuh -- bound levels of the branch-and-bound sweep and the kernel that builds them.
 * sm_100a only; compiled into libcsm_b200.so by csm_b200.cu.
 *
 * The reference scores a node of height h on its precomputed map of window 2^h
 * (scan_matcher_branch_bound.cpp:156-170, grid_map_builder.cpp:987-1012). Above the leaves that score
 * is only ever used as an upper bound: which internal nodes get expanded changes the amount of work,
 * never the result. The sweep therefore reads its own representation of the
 * coarse levels, built for the way it gathers:
 *
 *   B_0[r][c] = ceil(v / 257)              (u8; 257 * B_0 >= v for every u16 cell value v)
 *   B_h[r][c] = max B_0[r .. r + 2^h - 1][c .. c + 2^h - 1], cells outside the map read 0
 *
 * so 257 * sum_i B_h[hit_i + (x, y)] is an upper bound of the value sum of every leaf below node
 * (x, y) of height h, at most 256 per beam above the reference's own bound (0.4 % of the score range).
 * Cells at negative indices read 0 like the reference's lookups (grid_map.cpp:389-392).
 *
 * Layout of one level: u8 cells in tiles of 4 rows x 32 columns = one 128-byte line, tiles row-major
 * over a domain padded with zeros by at least 2^h + 1 cells on every side; inside a tile the cells go
 * [16 column pairs][4 rows][2 columns], so that a 32-byte sector is a patch of 4 rows x 8 columns.
 * The 32 gathers of a warp (8 nodes at adjacent angles x 4 adjacent beams: an arc of a few cells along
 * a wall) fall into one to three lines whatever the orientation of the wall, where the row-major u16
 * levels cost one line per map row touched; a clamped index needs no per-child bounds test; the
 * builder writes half the bytes of the u16 levels, and the streaming builder, whose threads own two
 * columns of four consecutive rows, stores those eight cells as one 64-bit word: a warp writes two
 * whole lines per instruction. The L1TEX wavefront rate (one line per cycle) bounds the sweep.
 * (Tiles of 8 x 16 cells made the sweep 5 % faster but break that store pattern.)
 *
 * Two builders write the levels: the streaming pyramid kernel (k_pyramid_stream2 in its bound mode:
 * one pass over the map, the reference's window with the far edge clamped, which covers more and is
 * therefore admissible as well) for batches of maps that fit it, k_bounds_build below for any shape.
 */
#pragma once

#include "csm_device.cuh"

namespace csm {

constexpr int kBlLog2R = 2, kBlLog2C = 5;     /* a tile = 4 rows x 32 columns = 128 cells */
constexpr int kBlTileR = 1 << kBlLog2R, kBlTileC = 1 << kBlLog2C;

__host__ __device__ constexpr int bl_pad_r(int hc) { return ((1 << hc) + 1 + kBlTileR - 1) & ~(kBlTileR - 1); }
__host__ __device__ constexpr int bl_pad_c(int hc) { return ((1 << hc) + 1 + kBlTileC - 1) & ~(kBlTileC - 1); }
/* tiles per row / rows of tiles of level hc of a rows x cols map */
__host__ __device__ constexpr int bl_tiles_per_row(int hc, int cols)
{
    return (2 * bl_pad_c(hc) + ((cols + kBlTileC - 1) & ~(kBlTileC - 1))) / kBlTileC;
}
__host__ __device__ constexpr int bl_tile_rows(int hc, int rows)
{
    return (2 * bl_pad_r(hc) + ((rows + kBlTileR - 1) & ~(kBlTileR - 1))) / kBlTileR;
}
__host__ __device__ constexpr size_t bl_level_bytes(int hc, int rows, int cols)
{
    return (size_t)bl_tiles_per_row(hc, cols) * (size_t)bl_tile_rows(hc, rows) * 128u;
}
/* byte offset of level hc (1 <= hc) inside a map's bound allocation: levels 1, 2, ... back to back */
__host__ __device__ constexpr size_t bl_level_offset(int hc, int rows, int cols)
{
    size_t off = 0;
    for (int j = 1; j < hc; ++j) off += bl_level_bytes(j, rows, cols);
    return off;
}

/* byte offset of padded cell (rp, cp) = (r + pad_r, c + pad_c) in a level with `tpr` tiles per row */
__host__ __device__ __forceinline__ unsigned int bl_cell(unsigned int rp, unsigned int cp, unsigned int tpr)
{
    /* inside a tile: [16 column pairs][4 rows][2 columns] */
    return (((rp >> 2) * tpr + (cp >> 5)) << 7) + ((cp & 30u) << 2) + ((rp & 3u) << 1) + (cp & 1u);
}

/* ceil(v / 257) of both u16 halves of a word, each result in its own 16-bit lane: v = 257 hi +
 * (lo - hi) with hi = v >> 8, lo = v & 255, so the quotient rounds up to hi + 1 exactly when lo > hi;
 * 255 + lo - hi has bit 8 set exactly then and never borrows from the neighbouring lane. (The u8x4
 * SIMD compare / maximum intrinsics are emulated on sm_100a, the u16x2 ones are single instructions:
 * the levels are therefore computed on 16-bit lanes and packed to bytes on the way out.) */
__device__ __forceinline__ unsigned int bl_encode2(unsigned int v)
{
    const unsigned int hi = (v >> 8) & 0x00ff00ffu, lo = v & 0x00ff00ffu;
    const unsigned int t = (lo | 0x01000100u) - hi - 0x00010001u;
    return hi + ((t >> 8) & 0x00010001u);
}

/* The same in five instructions instead of seven: byte permutes pick the high and low bytes of both lanes
 * and the carry byte of 255 + lo - hi */
__device__ __forceinline__ unsigned int bl_encode2p(unsigned int v)
{
    const unsigned int hi = __byte_perm(v, 0u, 0x4341), lo = __byte_perm(v, 0u, 0x4240);
    const unsigned int d = lo + 0x00ff00ffu - hi;
    return hi + __byte_perm(d, 0u, 0x4341);
}

struct BlJob
{
    const uint16_t* base;     /* level 0, row-major u16 */
    unsigned char*  out;      /* the map's bound allocation (levels 1.. back to back) */
    int rows, cols;
};

/* One CTA builds levels 1..L of a 128 x 128-cell region of one map: the u16 cells of the region plus a
 * halo of 2^L - 1 rows / columns on the high side are read once (L2 serves the halo overlap), encoded
 * in shared memory (two cells per word), and every level is the four-tap maximum of the previous one
 *     B_h[r][c] = max(B_{h-1}[r][c], B_{h-1}[r][c + e], B_{h-1}[r + e][c], B_{h-1}[r + e][c + e]),  e = 2^(h-1)
 * on u16x2 words (VIMNMX.U16x2), ping-pong between two buffers. A warp computes 8 rows x 32 columns
 * per step, lane = row * 4 + group of 8 columns (one 128-bit shared-memory access per tap), which is
 * the order of two adjacent 8 x 16 tiles in memory: the region is stored as whole 128-byte lines
 * straight from the registers that computed them. The shared-memory row pitch (== 16 mod 32 words)
 * sends the 8 lanes of each quarter warp to disjoint banks. */
constexpr int kBlOutR = 128, kBlOutC = 128;
constexpr int kBlThreads = 256;

__host__ __device__ constexpr int bl_in_rows(int L) { return (kBlOutR + (1 << L) - 1 + 7) & ~7; }
__host__ __device__ constexpr int bl_in_cols(int L) { return (kBlOutC + (1 << L) - 1 + 31) & ~31; }
__host__ __device__ constexpr int bl_pitch_words(int L)
{
    /* >= in_cols / 2 words, == 16 (mod 32) */
    return 16 + 32 * ((bl_in_cols(L) / 2 - 16 + 31) / 32);
}
__host__ __device__ constexpr size_t bl_smem_bytes(int L)
{
    return (size_t)2 * bl_in_rows(L) * bl_pitch_words(L) * sizeof(unsigned int);
}

struct BlCtx
{
    unsigned char* out;       /* the map's bound allocation */
    int R, C, r0, c0;         /* map extent, origin of the region */
    int warp, lane;
};

__device__ __forceinline__ uint4 bl_max4(const uint4 a, const uint4 b)
{
    return make_uint4(__vmaxu2(a.x, b.x), __vmaxu2(a.y, b.y), __vmaxu2(a.z, b.z), __vmaxu2(a.w, b.w));
}

/* the 8 cells that start E cells to the right of `a` (an = the next 8 cells of the row) */
template <int E>
__device__ __forceinline__ uint4 bl_shift(const uint4 a, const uint4 an)
{
    if (E == 1)
        return make_uint4(__byte_perm(a.x, a.y, 0x5432), __byte_perm(a.y, a.z, 0x5432),
                          __byte_perm(a.z, a.w, 0x5432), __byte_perm(a.w, an.x, 0x5432));
    if (E == 2)
        return make_uint4(a.y, a.z, a.w, an.x);
    return make_uint4(a.z, a.w, an.x, an.y);       /* E == 4 */
}

/* Level H from level H - 1 (src) into dst and, for the region proper, into the map's level H */
template <int H, int L>
__device__ __forceinline__ void bl_level(const unsigned int* __restrict__ src, unsigned int* __restrict__ dst, const BlCtx& X)
{
    constexpr int P = bl_pitch_words(L);
    constexpr int IN_R = bl_in_rows(L), IN_Q = bl_in_cols(L) / 8;      /* uint4 groups (8 cells) per row */
    constexpr int e = 1 << (H - 1);
    /* what the higher levels still read of this one */
    constexpr int need_r = kBlOutR + (1 << L) - (1 << H), need_c = kBlOutC + (1 << L) - (1 << H);
    constexpr int tiles_r = (need_r + 7) >> 3, pairs_c = (need_c + 31) >> 5;
    const int lr = X.lane >> 2, lq = X.lane & 3;
    unsigned char* out_h = X.out + bl_level_offset(H, X.R, X.C);
    const unsigned int tpr = (unsigned int)bl_tiles_per_row(H, X.C);
    const int pad_r = bl_pad_r(H), pad_c = bl_pad_c(H);
    for (int tp = X.warp; tp < tiles_r * pairs_c; tp += kBlThreads / 32) {
        const int tr = tp / pairs_c, tc = tp - tr * pairs_c;
        const int r = tr * 8 + lr, q = tc * 4 + lq;
        const int r2 = min(r + e, IN_R - 1);
        const uint4* __restrict__ row0 = reinterpret_cast<const uint4*>(src + r * P);
        const uint4* __restrict__ row1 = reinterpret_cast<const uint4*>(src + r2 * P);
        const uint4 a = row0[q], c = row1[q];
        uint4 b, d;
        if (e < 8) {
            const int qn = min(q + 1, IN_Q - 1);
            b = bl_shift<e < 8 ? e : 1>(a, row0[qn]);
            d = bl_shift<e < 8 ? e : 1>(c, row1[qn]);
        } else {
            const int qn = min(q + e / 8, IN_Q - 1);
            b = row0[qn]; d = row1[qn];
        }
        const uint4 v = bl_max4(bl_max4(a, b), bl_max4(c, d));
        if (H < L)
            *reinterpret_cast<uint4*>(dst + r * P + 4 * q) = v;
        const int gr0 = X.r0 + tr * 8, gc0 = X.c0 + tc * 32 + lq * 8;
        if (tr * 8 < kBlOutR && tc * 32 < kBlOutC && gr0 < X.R && gc0 < X.C) {
            /* this lane's 8 cells: row gr0 + lr, columns gc0 .. gc0 + 7 = four column pairs, 8 bytes apart */
            unsigned char* __restrict__ o = out_h + bl_cell((unsigned int)(gr0 + lr + pad_r), (unsigned int)(gc0 + pad_c), tpr);
            *reinterpret_cast<unsigned short*>(o)      = (unsigned short)__byte_perm(v.x, 0u, 0x4420);
            *reinterpret_cast<unsigned short*>(o + 8)  = (unsigned short)__byte_perm(v.y, 0u, 0x4420);
            *reinterpret_cast<unsigned short*>(o + 16) = (unsigned short)__byte_perm(v.z, 0u, 0x4420);
            *reinterpret_cast<unsigned short*>(o + 24) = (unsigned short)__byte_perm(v.w, 0u, 0x4420);
        }
    }
}

template <int L>
__global__ void __launch_bounds__(kBlThreads)
k_bounds_build(const BlJob* __restrict__ jobs, int regions_x)
{
    extern __shared__ __align__(16) unsigned int bl_smem[];
    constexpr int IN_R = bl_in_rows(L), IN_C = bl_in_cols(L);
    constexpr int P = bl_pitch_words(L);
    unsigned int* bufA = bl_smem;
    unsigned int* bufB = bl_smem + IN_R * P;

    const BlJob job = jobs[blockIdx.y];
    const int R = job.rows, C = job.cols;
    const int ry = blockIdx.x / regions_x, rx = blockIdx.x - ry * regions_x;
    const int r0 = ry * kBlOutR, c0 = rx * kBlOutC;
    if (r0 >= R || c0 >= C)
        return;

    /* stage 0: u16 cells of the region + halo -> B_0 in bufA, 16 bits per cell (cells outside the map read 0) */
    {
        const bool vec = (C & 7) == 0;
        constexpr int chunks = IN_C / 8;              /* 8 cells = 16 bytes in, 16 bytes out */
        for (int e = threadIdx.x; e < IN_R * chunks; e += kBlThreads) {
            const int rr = e / chunks, k = e - rr * chunks;
            const int gr = r0 + rr, gc = c0 + 8 * k;
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            if (gr < R && gc < C) {
                const uint16_t* src = job.base + (size_t)gr * C + gc;
                if (vec) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
                    w = make_uint4(bl_encode2(v.x), bl_encode2(v.y), bl_encode2(v.z), bl_encode2(v.w));
                } else {
                    unsigned int b[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        b[u] = (gc + u < C) ? (unsigned int)__ldg(src + u) : 0u;
                    w = make_uint4(bl_encode2(b[0] | (b[1] << 16)), bl_encode2(b[2] | (b[3] << 16)),
                                   bl_encode2(b[4] | (b[5] << 16)), bl_encode2(b[6] | (b[7] << 16)));
                }
            }
            *reinterpret_cast<uint4*>(bufA + rr * P + 4 * k) = w;
        }
    }
    __syncthreads();

    BlCtx X;
    X.out = job.out; X.R = R; X.C = C; X.r0 = r0; X.c0 = c0;
    X.warp = threadIdx.x >> 5; X.lane = threadIdx.x & 31;
    bl_level<1, L>(bufA, bufB, X);
    if (L >= 2) { __syncthreads(); bl_level<(L >= 2 ? 2 : 1), L>(bufB, bufA, X); }
    if (L >= 3) { __syncthreads(); bl_level<(L >= 3 ? 3 : 1), L>(bufA, bufB, X); }
    if (L >= 4) { __syncthreads(); bl_level<(L >= 4 ? 4 : 1), L>(bufB, bufA, X); }
    if (L >= 5) { __syncthreads(); bl_level<(L >= 5 ? 5 : 1), L>(bufA, bufB, X); }
    if (L >= 6) { __syncthreads(); bl_level<(L >= 6 ? 6 : 1), L>(bufB, bufA, X); }
}

} /* namespace csm */
