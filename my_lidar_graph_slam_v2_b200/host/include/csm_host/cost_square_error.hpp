/* cost_square_error.hpp -- the CPU epilogue every matcher runs once at the
 * winning pose: squared-error cost on the bilinearly smoothed map and the
 * covariance from the Gauss-Newton Hessian. Restates CostSquareError of the
 * reference (mapping/cost_function_square_error.cpp:27-36 bilinear, :48-75
 * cost, :131-195 covariance / Hessian, :233-274 gradients, :323-347 map
 * samples) in the same operation order, so the cost is bit-identical; the
 * covariance differs only by the 3x3 inverse (closed form here, Eigen there).
 * This part of the path stays on the CPU. */
#pragma once

#include "csm_host/types.hpp"

namespace csm_host {

/* CostFunction (mapping/cost_function.hpp:20-47): what a matcher needs of a cost */
class CostFunction
{
public:
    virtual ~CostFunction() = default;
    virtual double Cost(const GridMapView& map, const ScanData& scan, const Pose2D& sensor_pose) const = 0;
    virtual std::array<double, 9> ComputeCovariance(const GridMapView& map, const ScanData& scan,
                                                    const Pose2D& sensor_pose) const = 0;
};

class CostSquareError final : public CostFunction
{
public:
    explicit CostSquareError(double covariance_scale) : mCovarianceScale(covariance_scale) { }
    double CovarianceScale() const { return mCovarianceScale; }

    double Cost(const GridMapView& map, const ScanData& scan, const Pose2D& sensor_pose) const override;
    std::array<double, 9> ComputeCovariance(const GridMapView& map, const ScanData& scan,
                                            const Pose2D& sensor_pose) const override;
    /* Gauss-Newton Hessian (row-major 3x3) and residual vector at a sensor pose
     * (cost_function_square_error.cpp:151-195), the inputs of the linear-solver refiner */
    void ComputeHessianAndResidual(const GridMapView& map, const ScanData& scan, const Pose2D& sensor_pose,
                                   double hessian[9], double residual[3]) const;
    /* Both in one pass over the scan (what every matcher's epilogue needs) */
    std::array<double, 9> CostAndCovariance(const GridMapView& map, const ScanData& scan,
                                            const Pose2D& sensor_pose, double& cost) const;

private:
    double mCovarianceScale;
};

/* CostGreedyEndpoint (mapping/cost_function_greedy_endpoint.cpp:9-202), the hill-climbing matcher's
 * default cost: per beam, the best entry of a (2k+1)^2 Gaussian kernel around the hit cell among the
 * offsets where the hit cell is occupied and the cell HitAndMissedDist before it is free; covariance
 * from a central-difference gradient. Restated in the reference's operation order (bit-identical). */
class CostGreedyEndpoint final : public CostFunction
{
public:
    CostGreedyEndpoint(double map_resolution, double hit_and_missed_dist, double occupancy_threshold,
                       int kernel_size, double scaling_factor, double standard_deviation);
    double Cost(const GridMapView& map, const ScanData& scan, const Pose2D& sensor_pose) const override;
    std::array<double, 9> ComputeCovariance(const GridMapView& map, const ScanData& scan,
                                            const Pose2D& sensor_pose) const override;

private:
    double mMapResolution, mHitAndMissedDist, mOccupancyThreshold;
    int mKernelSize;
    double mVariance, mScalingFactor;
    std::vector<double> mCostLookupTable;
    double mDefaultCostValue;
};

} /* namespace csm_host */
