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

class CostSquareError
{
public:
    explicit CostSquareError(double covariance_scale) : mCovarianceScale(covariance_scale) { }
    double CovarianceScale() const { return mCovarianceScale; }

    double Cost(const GridMapView& map, const ScanData& scan, const Pose2D& sensor_pose) const;
    std::array<double, 9> ComputeCovariance(const GridMapView& map, const ScanData& scan,
                                            const Pose2D& sensor_pose) const;
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

} /* namespace csm_host */
