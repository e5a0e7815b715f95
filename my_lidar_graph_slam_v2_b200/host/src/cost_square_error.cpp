#include "csm_host/cost_square_error.hpp"

#include <algorithm>
#include <cmath>

namespace csm_host {

namespace {

constexpr int kLog2Block = 4;     /* reference block size 16 (launcher_settings_default.json:178) */

/* p(v): grid_values.hpp:26-36. A cell at 65535 reads 0.0 like an unknown one: the reference's table has
 * 65535 entries (grid_values.cpp:32-35), index 65535 lies in the zero tail of its allocation (see
 * k_saturated_unknown in csrc/csm_kernels.cuh; option "saturated_unknown" of the library). */
bool gSaturatedUnknown = true;
inline double Probability(std::uint16_t v)
{
    if (v == 0 || (v == 65535 && gSaturatedUnknown))
        return 0.0;
    const double pmin = 1e-3, pmax = 1.0 - 1e-3;
    return pmin + (pmax - pmin) * static_cast<double>(static_cast<int>(v) - 1) / 65534.0;
}

/* GridMap::ProbabilityOr(row, col, 0.5) on either form of the view. Block
 * allocation of a dense view without a bitmap is derived on first touch of a
 * block (a block counts as allocated iff it holds a non-zero cell). */
struct Sampler
{
    const GridMapView& map;
    int k, block_cols;
    mutable std::vector<std::int32_t> slot;   /* block-sparse: block -> position in `blocks`, -1 = unallocated;
                                                 dense: 1 allocated, 0 not, -2 not looked at yet */

    bool sparse;      /* block list: contiguous (`blocks`) or one heap allocation per block (`block_ptrs`) */

    explicit Sampler(const GridMapView& m) : map(m), k((m.blocks || m.block_ptrs) ? m.log2_block_size : kLog2Block),
        block_cols((m.cols + (1 << k) - 1) >> k), sparse(m.blocks != nullptr || m.block_ptrs != nullptr)
    {
        const int block_rows = (m.rows + (1 << k) - 1) >> k;
        const std::size_t nb = static_cast<std::size_t>(block_rows) * block_cols;
        if (sparse) {
            slot.assign(nb, -1);
            for (int b = 0; b < m.n_blocks; ++b)
                slot[m.block_index[b]] = b;
        } else if (m.block_allocated != nullptr) {
            slot.resize(nb);
            for (std::size_t b = 0; b < nb; ++b)
                slot[b] = m.block_allocated[b] ? 1 : 0;
        } else {
            slot.assign(nb, -2);
        }
    }

    bool DenseBlockAllocated(int brow, int bcol) const
    {
        std::int32_t& st = slot[static_cast<std::size_t>(brow) * block_cols + bcol];
        if (st == -2) {
            st = 0;
            const int r1 = std::min((brow + 1) << k, map.rows), c1 = std::min((bcol + 1) << k, map.cols);
            for (int r = brow << k; r < r1 && !st; ++r)
                for (int c = bcol << k; c < c1; ++c)
                    if (map.values[static_cast<std::size_t>(r) * map.cols + c] != 0) { st = 1; break; }
        }
        return st != 0;
    }

    /* block of (row, col) allocated? (row, col inside the map) */
    bool Allocated(int row, int col) const
    {
        const int brow = row >> k, bcol = col >> k;
        if (sparse)
            return slot[static_cast<std::size_t>(brow) * block_cols + bcol] >= 0;
        return DenseBlockAllocated(brow, bcol);
    }

    /* grid_map.cpp:424-436 */
    double At(int row, int col) const
    {
        if (row < 0 || row >= map.rows || col < 0 || col >= map.cols)
            return 0.5;
        const int brow = row >> k, bcol = col >> k;
        if (sparse) {
            const std::int32_t b = slot[static_cast<std::size_t>(brow) * block_cols + bcol];
            if (b < 0)
                return 0.5;
            const int mask = (1 << k) - 1;
            const std::size_t in_block = (static_cast<std::size_t>(row & mask) << k) + (col & mask);
            if (map.block_ptrs != nullptr)
                return Probability(map.block_ptrs[b][in_block]);
            return Probability(map.blocks[(static_cast<std::size_t>(b) << (2 * k)) + in_block]);
        }
        if (!DenseBlockAllocated(brow, bcol))
            return 0.5;
        return Probability(map.values[static_cast<std::size_t>(row) * map.cols + col]);
    }
};

struct Neighbours
{
    double dx, dy, m00, m01, m10, m11;

    double Smoothed() const
    {
        return dy * (dx * m11 + (1.0 - dx) * m01) + (1.0 - dy) * (dx * m10 + (1.0 - dx) * m00);
    }
};

Neighbours Closest(const Sampler& s, double fx, double fy)
{
    const double x0 = std::floor(fx), y0 = std::floor(fy);
    const int xc0 = std::max(static_cast<int>(x0), 0);
    const int yc0 = std::max(static_cast<int>(y0), 0);
    const int xc1 = std::min(xc0 + 1, s.map.cols - 1);
    const int yc1 = std::min(yc0 + 1, s.map.rows - 1);
    return Neighbours { fx - x0, fy - y0, s.At(yc0, xc0), s.At(yc1, xc0), s.At(yc0, xc1), s.At(yc1, xc1) };
}

inline void HitPoint(const ScanData& scan, const Pose2D& pose, std::size_t i, double& hx, double& hy)
{
    /* sensor_data.hpp:190-203 */
    /* one call; glibc's sincos returns exactly the values of sin() and cos() */
    double s, c;
    ::sincos(pose.theta + scan.angles[i], &s, &c);
    hx = pose.x + scan.ranges[i] * c;
    hy = pose.y + scan.ranges[i] * s;
}

} /* namespace */

namespace {

/* One pass over the scan: squared-error cost (:48-75) and, when h != nullptr,
 * the Gauss-Newton Hessian (:151-195) from the same map samples. The two
 * accumulations are independent, each in the reference's own order. */
double Accumulate(const GridMapView& map, const ScanData& scan, const Pose2D& pose, double* h,
                  double* residual = nullptr)
{
    const Sampler sampler(map);
    double cost = 0.0;
    const double inv_res = 1.0 / map.resolution;
    for (std::size_t i = 0; i < scan.NumOfScans(); ++i) {
        double hx, hy;
        HitPoint(scan, pose, i, hx, hy);
        const double fx = (hx - map.offset_x) / map.resolution;
        const double fy = (hy - map.offset_y) / map.resolution;
        const Neighbours n = Closest(sampler, fx, fy);
        cost += std::pow(1.0 - n.Smoothed(), 2.0);
        if (h == nullptr)
            continue;
        const double gx = n.dy * (n.m11 - n.m01) + (1.0 - n.dy) * (n.m10 - n.m00);
        const double gy = n.dx * (n.m11 - n.m10) + (1.0 - n.dx) * (n.m01 - n.m00);
        const double gt = -(hy - pose.y) * gx + (hx - pose.x) * gy;
        const double g[3] = { gx * inv_res, gy * inv_res, gt * inv_res };
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                h[r * 3 + c] += g[r] * g[c];
        if (residual != nullptr) {
            /* residualVec += mapGrad * (1 - smoothed), :187-192 */
            const double res = 1.0 - n.Smoothed();
            for (int r = 0; r < 3; ++r)
                residual[r] += g[r] * res;
        }
    }
    return cost;
}

std::array<double, 9> InverseScaled(const double h[9], double scale)
{
    const double det = h[0] * (h[4] * h[8] - h[5] * h[7]) - h[1] * (h[3] * h[8] - h[5] * h[6]) +
                       h[2] * (h[3] * h[7] - h[4] * h[6]);
    const double id = 1.0 / det;
    std::array<double, 9> cov {
        (h[4] * h[8] - h[5] * h[7]) * id, (h[2] * h[7] - h[1] * h[8]) * id, (h[1] * h[5] - h[2] * h[4]) * id,
        (h[5] * h[6] - h[3] * h[8]) * id, (h[0] * h[8] - h[2] * h[6]) * id, (h[2] * h[3] - h[0] * h[5]) * id,
        (h[3] * h[7] - h[4] * h[6]) * id, (h[1] * h[6] - h[0] * h[7]) * id, (h[0] * h[4] - h[1] * h[3]) * id };
    for (double& v : cov)
        v *= scale;
    return cov;
}

} /* namespace */

double CostSquareError::Cost(const GridMapView& map, const ScanData& scan, const Pose2D& pose) const
{
    return Accumulate(map, scan, pose, nullptr);
}

std::array<double, 9> CostSquareError::ComputeCovariance(const GridMapView& map, const ScanData& scan,
                                                         const Pose2D& pose) const
{
    double h[9] = { 0.0 };
    Accumulate(map, scan, pose, h);
    return InverseScaled(h, mCovarianceScale);
}

void CostSquareError::ComputeHessianAndResidual(const GridMapView& map, const ScanData& scan, const Pose2D& pose,
                                                double hessian[9], double residual[3]) const
{
    for (int i = 0; i < 9; ++i) hessian[i] = 0.0;
    for (int i = 0; i < 3; ++i) residual[i] = 0.0;
    Accumulate(map, scan, pose, hessian, residual);
}

std::array<double, 9> CostSquareError::CostAndCovariance(const GridMapView& map, const ScanData& scan,
                                                         const Pose2D& pose, double& cost) const
{
    double h[9] = { 0.0 };
    cost = Accumulate(map, scan, pose, h);
    return InverseScaled(h, mCovarianceScale);
}

/* ---- greedy endpoint cost --------------------------------------------------------------- */
CostGreedyEndpoint::CostGreedyEndpoint(double map_resolution, double hit_and_missed_dist,
                                       double occupancy_threshold, int kernel_size, double scaling_factor,
                                       double standard_deviation) :
    mMapResolution(map_resolution), mHitAndMissedDist(hit_and_missed_dist),
    mOccupancyThreshold(occupancy_threshold), mKernelSize(kernel_size),
    mVariance(standard_deviation * standard_deviation), mScalingFactor(scaling_factor), mDefaultCostValue(0.0)
{
    /* SetupLookupTable, cost_function_greedy_endpoint.cpp:168-199 */
    const int kernel = 2 * mKernelSize + 1;
    mCostLookupTable.assign(static_cast<std::size_t>(kernel) * kernel, 0.0);
    for (int ky = -mKernelSize; ky <= mKernelSize; ++ky)
        for (int kx = -mKernelSize; kx <= mKernelSize; ++kx) {
            const double dx = mMapResolution * kx, dy = mMapResolution * ky;
            const double squared = dx * dx + dy * dy;
            mCostLookupTable[(mKernelSize + ky) * kernel + (mKernelSize + kx)] = -std::exp(-0.5 * squared / mVariance);
        }
    const double mx = mMapResolution * (mKernelSize + 1), my = mMapResolution * (mKernelSize + 1);
    mDefaultCostValue = -std::exp(-0.5 * (mx * mx + my * my) / mVariance);
}

double CostGreedyEndpoint::Cost(const GridMapView& map, const ScanData& scan, const Pose2D& pose) const
{
    /* cost_function_greedy_endpoint.cpp:33-101. ProbabilityOr(row, col, unknown = 0.0): cells outside the
     * map, of unallocated blocks and unknown cells all read 0.0 and are skipped, so block allocation does
     * not matter here */
    const Sampler sampler(map);
    auto prob = [&](int row, int col) {
        if (row < 0 || row >= map.rows || col < 0 || col >= map.cols)
            return 0.0;
        const double p = sampler.At(row, col);
        return p == 0.5 && !sampler.Allocated(row, col) ? 0.0 : p;
    };
    const int kernel = 2 * mKernelSize + 1;
    double sum = 0.0;
    for (std::size_t i = 0; i < scan.NumOfScans(); ++i) {
        /* ScanData::HitAndMissedPoint, sensor_data.hpp:252-273 */
        const double range = scan.ranges[i];
        const double c = std::cos(pose.theta + scan.angles[i]), s = std::sin(pose.theta + scan.angles[i]);
        const double hx = pose.x + range * c, hy = pose.y + range * s;
        const double mx = pose.x + (range - mHitAndMissedDist) * c, my = pose.y + (range - mHitAndMissedDist) * s;
        const int hit_col = static_cast<int>(std::floor((hx - map.offset_x) / map.resolution));
        const int hit_row = static_cast<int>(std::floor((hy - map.offset_y) / map.resolution));
        const int mis_col = static_cast<int>(std::floor((mx - map.offset_x) / map.resolution));
        const int mis_row = static_cast<int>(std::floor((my - map.offset_y) / map.resolution));
        double best = mDefaultCostValue;
        for (int ky = -mKernelSize; ky <= mKernelSize; ++ky)
            for (int kx = -mKernelSize; kx <= mKernelSize; ++kx) {
                const double hit = prob(hit_row + ky, hit_col + kx);
                const double missed = prob(mis_row + ky, mis_col + kx);
                if (hit == 0.0 || missed == 0.0)
                    continue;
                if (hit < mOccupancyThreshold || missed > mOccupancyThreshold)
                    continue;
                best = std::min(best, mCostLookupTable[(mKernelSize + ky) * kernel + (mKernelSize + kx)]);
            }
        sum += best;
    }
    sum *= mScalingFactor;
    return sum;
}

std::array<double, 9> CostGreedyEndpoint::ComputeCovariance(const GridMapView& map, const ScanData& scan,
                                                            const Pose2D& pose) const
{
    /* ComputeGradient + ComputeCovariance, cost_function_greedy_endpoint.cpp:104-165 */
    const double dl = map.resolution, da = 1e-2;
    auto cost = [&](double x, double y, double t) { return Cost(map, scan, Pose2D { x, y, t }); };
    const double cx = cost(pose.x + dl, pose.y + 0.0, pose.theta + 0.0) - cost(pose.x - dl, pose.y - 0.0, pose.theta - 0.0);
    const double cy = cost(pose.x + 0.0, pose.y + dl, pose.theta + 0.0) - cost(pose.x - 0.0, pose.y - dl, pose.theta - 0.0);
    const double ct = cost(pose.x + 0.0, pose.y + 0.0, pose.theta + da) - cost(pose.x - 0.0, pose.y - 0.0, pose.theta - da);
    const double g[3] = { 0.5 * cx / dl, 0.5 * cy / dl, 0.5 * ct / da };
    std::array<double, 9> cov {};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            cov[r * 3 + c] = g[r] * g[c];
    cov[0] += 0.1; cov[4] += 0.1; cov[8] += 0.1;
    return cov;
}

} /* namespace csm_host */
