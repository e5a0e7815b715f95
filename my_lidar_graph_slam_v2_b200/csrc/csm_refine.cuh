/* csm_refine.cuh -- sub-cell refinement of found poses on the device.
 *
 * Replaces, for a whole batch of detected loops at once, the CPU stage that
 * follows every successful coarse match in the reference:
 *   ScanMatcherLinearSolver::OptimizePose / OptimizeStep
 *                        (scan_matcher_linear_solver.cpp:66-170)
 *   CostSquareError::Cost, ComputeHessianAndResidual, ComputeCovariance
 *                        (cost_function_square_error.cpp:48-75,131-195)
 *   GridMap::ProbabilityOr(row, col, 0.5)          (grid_map.cpp:424-436)
 * as it is called from LoopDetectorBranchBound::Detect
 * (loop_detector_branch_bound.cpp:110-135).
 *
 * One CTA of kRefThreads threads per query. A pass over the scan evaluates
 * the squared-error cost, the Gauss-Newton Hessian and the residual vector at
 * one pose (the reference makes two passes per iteration, Cost at the new pose
 * and Hessian/residual at the same pose in the next iteration; here both come
 * from the same map samples). The 3x3 damped system is solved by thread 0 with
 * a column-pivoting Householder QR (Eigen's colPivHouseholderQr, :161).
 * Arithmetic is FP64 throughout. Sums are reduced in a fixed tree order, so
 * results are deterministic; they differ from the reference's sequential sums
 * by rounding only (north_star tolerance for refined poses: 1e-5 relative).
 *
 * Map samples need the reference's notion of allocated blocks: a cell of an
 * unallocated block reads as probability 0.5, an unknown cell (value 0) of an
 * allocated block as 0.0. DevQuery.alloc holds one byte per block.
 */
#pragma once

#include "csm_device.cuh"
#include "csm_b200.h"

namespace csm {

constexpr int kRefThreads = 384;     /* one beam per thread for scans of up to 384 beams */
constexpr int kRefWarps = kRefThreads / 32;
constexpr int kRefTerms = 10;      /* cost, H00 H01 H02 H11 H12 H22, r0 r1 r2 */

/* One byte per 2^k x 2^k block of a dense map: 1 iff the block holds a non-zero
 * cell. This is how a dense upload (no block list) defines "allocated"; the host
 * mirror uses the same rule (host/src/cost_square_error.cpp). One warp per block,
 * blockIdx.y = map of the batch. */
struct AllocJob
{
    const uint16_t* base;
    unsigned char*  alloc;
    int rows, cols;
};

__global__ void __launch_bounds__(256)
k_block_alloc(const AllocJob* __restrict__ jobs, int k)
{
    const AllocJob job = jobs[blockIdx.y];
    const int rows = job.rows, cols = job.cols;
    const int bs = 1 << k;
    const int block_rows = (rows + bs - 1) >> k, block_cols = (cols + bs - 1) >> k;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int b = warp; b < block_rows * block_cols; b += nwarps) {
        const int brow = b / block_cols, bcol = b - brow * block_cols;
        bool any = false;
        for (int e = lane; e < bs * bs; e += 32) {
            const int r = (brow << k) + (e >> k), c = (bcol << k) + (e & (bs - 1));
            if (r < rows && c < cols)
                any = any || job.base[(size_t)r * cols + c] != 0;
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0)
            job.alloc[b] = any ? 1 : 0;
    }
}

struct RefineArgs
{
    const csm_result* results;     /* start = sensor pose + step * best index of a found query; or null */
    const double* start;           /* [3 * nq] start poses when results == null */
    csm_refined* out;
    int max_iterations;            /* 0: no iteration, cost and covariance at the start pose only */
    int always;                    /* 1: evaluate every query, found or not (single-scan epilogue) */
    double convergence_threshold;
    double lambda0;
    double covariance_scale;
};

/* GridMap::ProbabilityOr(row, col, 0.5), grid_map.cpp:424-436, for the four cells around a hit
 * point. Branch-free: the allocation byte and the cell of all four are loaded at once (an index
 * outside the map reads cell 0 and is masked), so a pass costs one memory round trip per beam
 * instead of eight dependent ones. */
__device__ __forceinline__ void refine_sample4(const DevQuery& Q, const int (&row)[4], const int (&col)[4],
                                               double (&p)[4])
{
    bool inside[4];
    unsigned int a[4], v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        inside[j] = (unsigned)row[j] < (unsigned)Q.rows && (unsigned)col[j] < (unsigned)Q.cols;
        const int r = inside[j] ? row[j] : 0, c = inside[j] ? col[j] : 0;
        a[j] = __ldg(Q.alloc + (size_t)(r >> Q.alloc_log2bs) * Q.alloc_bcols + (c >> Q.alloc_log2bs));
        v[j] = __ldg(Q.lvl[0] + (size_t)r * Q.cols + c);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        p[j] = (!inside[j] || a[j] == 0u) ? 0.5 : (v[j] == 0u ? 0.0 : value_to_probability(v[j]));
}

/* Column-pivoting Householder QR solve of a 3x3 system, a x = b (row-major a).
 * Same steps as Eigen's ColPivHouseholderQR::solve: pivot on the largest remaining
 * column norm, reflect, rank from pivots above eps * size * largest pivot.
 * Fully unrolled with compile-time indices (column swaps as conditional moves), so
 * that the matrix stays in registers: the solve sits on the serial path of every
 * iteration. */
template <int K>
__device__ __forceinline__ void refine_qr_step(double (&a)[3][3], double (&b)[3], int (&perm)[3],
                                               double& max_pivot)
{
    /* column with the largest remaining norm comes first */
    double norm[3];
#pragma unroll
    for (int c = K; c < 3; ++c) {
        double s = 0.0;
#pragma unroll
        for (int r = K; r < 3; ++r) s += a[r][c] * a[r][c];
        norm[c] = s;
    }
    int best = K;
    double best_norm = norm[K];
#pragma unroll
    for (int c = K + 1; c < 3; ++c)
        if (norm[c] > best_norm) { best_norm = norm[c]; best = c; }
#pragma unroll
    for (int c = K + 1; c < 3; ++c)
        if (best == c) {
#pragma unroll
            for (int r = 0; r < 3; ++r) { const double t = a[r][K]; a[r][K] = a[r][c]; a[r][c] = t; }
            const int t = perm[K]; perm[K] = perm[c]; perm[c] = t;
        }
    /* Householder reflector H = I - tau v v^T that maps a[K..2][K] onto (beta, 0, 0) */
    double tail = 0.0;
#pragma unroll
    for (int r = K + 1; r < 3; ++r) tail += a[r][K] * a[r][K];
    const double c0 = a[K][K];
    double beta = c0, tau = 0.0, v[3] = { 0.0, 0.0, 0.0 };
    if (tail > 0.0) {
        beta = sqrt(c0 * c0 + tail);
        if (c0 >= 0.0) beta = -beta;
#pragma unroll
        for (int r = K + 1; r < 3; ++r) v[r] = a[r][K] / (c0 - beta);
        v[K] = 1.0;
        tau = (beta - c0) / beta;
    }
    a[K][K] = beta;
#pragma unroll
    for (int r = K + 1; r < 3; ++r) a[r][K] = 0.0;
    if (tau != 0.0) {
#pragma unroll
        for (int c = K + 1; c < 3; ++c) {
            double dot = 0.0;
#pragma unroll
            for (int r = K; r < 3; ++r) dot += v[r] * a[r][c];
#pragma unroll
            for (int r = K; r < 3; ++r) a[r][c] -= tau * v[r] * dot;
        }
        double dot = 0.0;
#pragma unroll
        for (int r = K; r < 3; ++r) dot += v[r] * b[r];
#pragma unroll
        for (int r = K; r < 3; ++r) b[r] -= tau * v[r] * dot;
    }
    max_pivot = fmax(max_pivot, fabs(beta));
}

__device__ __forceinline__ void refine_solve3(const double (&a_in)[9], const double (&b_in)[3], double (&x)[3])
{
    double a[3][3], b[3] = { b_in[0], b_in[1], b_in[2] };
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            a[r][c] = a_in[r * 3 + c];
    int perm[3] = { 0, 1, 2 };
    double max_pivot = 0.0;
    refine_qr_step<0>(a, b, perm, max_pivot);
    refine_qr_step<1>(a, b, perm, max_pivot);
    refine_qr_step<2>(a, b, perm, max_pivot);
    /* numerical rank as Eigen decides it; pivots come out in non-increasing magnitude */
    const double threshold = 2.220446049250313e-16 * 3.0 * max_pivot;
    int rank = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
        if (fabs(a[k][k]) > threshold) ++rank;
    double y[3] = { 0.0, 0.0, 0.0 };
#pragma unroll
    for (int k = 2; k >= 0; --k)
        if (k < rank) {
            double s = b[k];
#pragma unroll
            for (int c = k + 1; c < 3; ++c)
                if (c < rank) s -= a[k][c] * y[c];
            y[k] = s / a[k][k];
        }
#pragma unroll
    for (int j = 0; j < 3; ++j)
        x[j] = perm[0] == j ? y[0] : (perm[1] == j ? y[1] : y[2]);
}

/* One pass over the scan at sensor pose (px, py, pt): partial sums of this thread.
 * cost_function_square_error.cpp:48-75 (cost), :151-195 (Hessian, residual). */
__device__ __forceinline__ void refine_pass(const DevQuery& Q, double px, double py, double pt,
                                            double (&acc)[kRefTerms])
{
#pragma unroll
    for (int j = 0; j < kRefTerms; ++j) acc[j] = 0.0;
    const double inv_res = __ddiv_rn(1.0, Q.res);
    for (int i = threadIdx.x; i < Q.n; i += kRefThreads) {
        /* ScanData::HitPoint, sensor_data.hpp:190-203 */
        double sn, cs;
        sincos(__dadd_rn(pt, Q.angles[i]), &sn, &cs);
        const double range = Q.ranges[i];
        const double hx = __dadd_rn(px, __dmul_rn(range, cs));
        const double hy = __dadd_rn(py, __dmul_rn(range, sn));
        /* floating-point cell coordinate and the four closest cells (:21-46) */
        const double fx = __ddiv_rn(__dsub_rn(hx, Q.offx), Q.res);
        const double fy = __ddiv_rn(__dsub_rn(hy, Q.offy), Q.res);
        const double x0 = floor(fx), y0 = floor(fy);
        const double lim = 1073741824.0;
        const int xc0 = max((int)fmin(fmax(x0, -lim), lim), 0);
        const int yc0 = max((int)fmin(fmax(y0, -lim), lim), 0);
        const int xc1 = min(xc0 + 1, Q.cols - 1);
        const int yc1 = min(yc0 + 1, Q.rows - 1);
        const double dx = __dsub_rn(fx, x0), dy = __dsub_rn(fy, y0);
        const int rws[4] = { yc0, yc1, yc0, yc1 }, cls[4] = { xc0, xc0, xc1, xc1 };
        double pm[4];
        refine_sample4(Q, rws, cls, pm);
        const double m00 = pm[0], m01 = pm[1], m10 = pm[2], m11 = pm[3];
        const double smoothed = dy * (dx * m11 + (1.0 - dx) * m01) + (1.0 - dy) * (dx * m10 + (1.0 - dx) * m00);
        const double err = 1.0 - smoothed;
        acc[0] += err * err;
        const double gx = dy * (m11 - m01) + (1.0 - dy) * (m10 - m00);
        const double gy = dx * (m11 - m10) + (1.0 - dx) * (m01 - m00);
        const double gt = -(hy - py) * gx + (hx - px) * gy;
        const double g0 = gx * inv_res, g1 = gy * inv_res, g2 = gt * inv_res;
        acc[1] += g0 * g0; acc[2] += g0 * g1; acc[3] += g0 * g2;
        acc[4] += g1 * g1; acc[5] += g1 * g2; acc[6] += g2 * g2;
        acc[7] += g0 * err; acc[8] += g1 * err; acc[9] += g2 * err;
    }
}

/* Sum the partial sums over the CTA in a fixed order; the totals land in s_tot. */
__device__ __forceinline__ void refine_reduce(double (&acc)[kRefTerms], double (*s_part)[kRefTerms],
                                              double* s_tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kRefTerms; ++j) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < kRefTerms) {
        double v = s_part[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kRefWarps; ++w) v += s_part[w][threadIdx.x];
        s_tot[threadIdx.x] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kRefThreads)
k_refine(const DevQuery* __restrict__ queries, RefineArgs A)
{
    __shared__ double s_part[kRefWarps][kRefTerms];
    __shared__ double s_tot[kRefTerms];
    __shared__ double s_pose[3];
    __shared__ int s_go;
    const int q = blockIdx.x;
    const DevQuery& Q = queries[q];
    csm_refined* out = A.out + q;

    if (threadIdx.x == 0) {
        if (A.results != nullptr) {
            const csm_result r = A.results[q];
            s_go = r.found | A.always;
            /* best sensor pose of the coarse search, scan_matcher_branch_bound.cpp:237-240 */
            s_pose[0] = __dadd_rn(Q.sx, __dmul_rn(Q.stepx, (double)r.best_x));
            s_pose[1] = __dadd_rn(Q.sy, __dmul_rn(Q.stepy, (double)r.best_y));
            s_pose[2] = __dadd_rn(Q.theta0, __dmul_rn(Q.step_t, (double)r.best_t));
        } else {
            s_go = 1;
            s_pose[0] = A.start[3 * q]; s_pose[1] = A.start[3 * q + 1]; s_pose[2] = A.start[3 * q + 2];
        }
    }
    __syncthreads();
    if (!s_go) {
        if (threadIdx.x == 0) {
            csm_refined z;
            z.pose[0] = z.pose[1] = z.pose[2] = 0.0;
            for (int j = 0; j < 9; ++j) z.covariance[j] = 0.0;
            z.initial_cost = z.final_cost = 0.0;
            z.lambda = A.lambda0;
            z.iterations = 0; z.valid = 0;
            *out = z;
        }
        return;
    }

    double acc[kRefTerms];
    refine_pass(Q, s_pose[0], s_pose[1], s_pose[2], acc);
    refine_reduce(acc, s_part, s_tot);
    /* thread 0 carries the solver state (scan_matcher_linear_solver.cpp:82-110) */
    double lambda = A.lambda0, prev_cost = s_tot[0], initial_cost = s_tot[0], cost = s_tot[0];
    int iterations = 0;
    while (A.max_iterations > 0) {
        if (threadIdx.x == 0) {
            /* OptimizeStep, :143-170: (H + lambda I) d = residual */
            const double h[9] = { s_tot[1] + lambda, s_tot[2], s_tot[3],
                                  s_tot[2], s_tot[4] + lambda, s_tot[5],
                                  s_tot[3], s_tot[5], s_tot[6] + lambda };
            const double r[3] = { s_tot[7], s_tot[8], s_tot[9] };
            double d[3];
            refine_solve3(h, r, d);
            s_pose[0] += d[0]; s_pose[1] += d[1]; s_pose[2] += d[2];
        }
        __syncthreads();
        refine_pass(Q, s_pose[0], s_pose[1], s_pose[2], acc);
        refine_reduce(acc, s_part, s_tot);
        cost = s_tot[0];
        if (++iterations >= A.max_iterations || fabs(prev_cost - cost) < A.convergence_threshold)
            break;
        lambda = cost < prev_cost ? fmax(1e-8, lambda * 0.5) : fmin(1e-4, lambda * 2.0);
        prev_cost = cost;
    }
    if (threadIdx.x == 0) {
        csm_refined o;
        o.pose[0] = s_pose[0]; o.pose[1] = s_pose[1]; o.pose[2] = s_pose[2];
        /* ComputeCovariance at the final pose, cost_function_square_error.cpp:131-146:
         * scale * inverse of the (undamped) Hessian, by cofactors like Eigen's 3x3 inverse */
        const double h[9] = { s_tot[1], s_tot[2], s_tot[3], s_tot[2], s_tot[4], s_tot[5],
                              s_tot[3], s_tot[5], s_tot[6] };
        const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) +
                           h[2] * (h[3] * h[7] - h[4] * h[6]);
        const double id = 1.0 / det;
        o.covariance[0] = (h[4] * h[8] - h[5] * h[7]) * id;
        o.covariance[1] = (h[2] * h[7] - h[1] * h[8]) * id;
        o.covariance[2] = (h[1] * h[5] - h[2] * h[4]) * id;
        o.covariance[3] = (h[5] * h[6] - h[3] * h[8]) * id;
        o.covariance[4] = (h[0] * h[8] - h[2] * h[6]) * id;
        o.covariance[5] = (h[2] * h[3] - h[0] * h[5]) * id;
        o.covariance[6] = (h[3] * h[7] - h[4] * h[6]) * id;
        o.covariance[7] = (h[1] * h[6] - h[0] * h[7]) * id;
        o.covariance[8] = (h[0] * h[4] - h[1] * h[3]) * id;
        for (int j = 0; j < 9; ++j) o.covariance[j] *= A.covariance_scale;
        o.initial_cost = initial_cost;
        o.final_cost = cost;
        o.lambda = lambda;
        o.iterations = iterations;
        o.valid = 1;
        *out = o;
    }
}

} /* namespace csm */
