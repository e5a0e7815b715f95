/* scan_matchers.hpp -- GPU scan matchers behind the reference's ScanMatcher
 * plugin interface (mapping/scan_matcher.hpp:89-117: Name() and
 * OptimizePose(query) -> summary), selected by the launcher JSON type strings
 * "RealTimeCorrelative" | "BranchBound" | "GridSearch"
 * (scan_matcher_factory.cpp:200-214). Constructor parameters are the reference's:
 *   ScanMatcherCorrelative(name, costFunc, lowResolution, rangeX, rangeY, rangeTheta)
 *       scan_matcher_correlative.hpp:59-65
 *   ScanMatcherBranchBound(name, costFunc, nodeHeightMax, rangeX, rangeY, rangeTheta)
 *       scan_matcher_branch_bound.hpp:108-116 (the score function argument is gone:
 *       the device computes ScorePixelAccurate itself)
 *   ScanMatcherGridSearch(name, costFunc, rangeX, rangeY, rangeTheta, stepX, stepY, stepTheta)
 *       scan_matcher_grid_search.hpp:49-58
 * Each instance owns (or shares) one csm_handle = one CUDA stream with its own
 * device buffers, so the front-end matcher and the back-end detector can run
 * concurrently from their two threads (lidar_graph_slam.cpp:777-779). Errors
 * follow the reference convention: print and abort (util.hpp:39-72). */
#pragma once

#include "csm_b200.h"
#include "csm_host/cost_square_error.hpp"
#include "csm_host/metrics.hpp"
#include "csm_host/types.hpp"

namespace csm_host {

using CostFuncPtr = std::shared_ptr<CostSquareError>;

/* Shared ownership of a csm_handle */
class DeviceContext
{
public:
    explicit DeviceContext(int device);
    ~DeviceContext();
    DeviceContext(const DeviceContext&) = delete;
    DeviceContext& operator=(const DeviceContext&) = delete;
    csm_handle Handle() const { return mHandle; }
    int Device() const { return mDevice; }
    /* Page-locked staging area of at least `bytes` (grown on demand, reused by every call): map data
     * copied here first crosses PCIe by DMA while the call goes on */
    void* Staging(std::size_t bytes);
    /* Abort with the library's message unless rc == CSM_OK */
    void Check(int rc, const char* what) const;
    /* Matchers on this context take cost and covariance of the pose they decide on from the device
     * (csm_set_epilogue: computed behind the match in the same submission) instead of running
     * CostSquareError on the CPU afterwards. Same quantities, summed in a different order. */
    void SetDeviceEpilogue(bool on) { mDeviceEpilogue = on; }
    bool DeviceEpilogue() const { return mDeviceEpilogue; }
    /* Real-time correlative / branch-and-bound matchers on this context hand the pose they find to
     * the reference's final matcher (ScanMatcherLinearSolver over CostSquareError, what the front end
     * runs next: lidar_graph_slam_frontend.cpp:216-230) on the device, in the same submission: the
     * summary then carries the refined pose, its cost and covariance. max_iterations <= 0 = off.
     * The damping factor is carried from match to match like the solver's member. */
    void SetDeviceFinalMatcher(int max_iterations, double convergence_threshold, double initial_lambda,
                               double covariance_scale)
    {
        mFinal.max_iterations = max_iterations; mFinal.reserved = 0;
        mFinal.convergence_threshold = convergence_threshold;
        mFinal.lambda = initial_lambda; mFinal.covariance_scale = covariance_scale;
    }
    bool HasDeviceFinalMatcher() const { return mFinal.max_iterations > 0; }
    csm_refine_params& FinalMatcherParams() { return mFinal; }

private:
    csm_handle mHandle;
    int mDevice = 0;
    bool mDeviceEpilogue = false;
    csm_refine_params mFinal {};
    void* mStaging = nullptr;
    std::size_t mStagingBytes = 0;
};
using DeviceContextPtr = std::shared_ptr<DeviceContext>;

struct ScanMatchingQuery
{
    const GridMapView& grid_map;
    ScanDataPtr scan_data;
    Pose2D map_local_initial_pose;
};

class ScanMatcher
{
public:
    ScanMatcher(const std::string& name, const DeviceContextPtr& context) :
        mName(name), mContext(context) { }
    virtual ~ScanMatcher() = default;
    const std::string& Name() const { return mName; }
    virtual ScanMatchingSummary OptimizePose(const ScanMatchingQuery& query) = 0;
    const DeviceContextPtr& Context() const { return mContext; }
    /* Report the reference's "<Name>.<Metric>" value sequences here (metrics.hpp); null = off */
    void SetMetricSink(const MetricSinkPtr& sink) { mMetricSink = sink; }

protected:
    void Observe(const char* metric, double value) const
    {
        if (mMetricSink) mMetricSink->Observe(mName + "." + metric, value);
    }
    /* The ids every matcher reports after a match (scan_matcher_correlative.cpp:221-237) */
    void ObserveSummary(const ScanMatchingSummary& summary, const ScanData& scan, double micro) const;

    /* Upload the map unless a map with this id is already resident */
    std::int64_t EnsureMap(const GridMapView& map);
    /* Cost, covariance and MoveBackward at the winning sensor pose
     * (scan_matcher_correlative.cpp:203-219) */
    void Epilogue(const GridMapView& map, const ScanData& scan, const Pose2D& best_sensor_pose,
                  const CostFuncPtr& cost, ScanMatchingSummary& summary) const;

    std::string mName;
    DeviceContextPtr mContext;
    bool mEpilogueOnDevice = false;     /* the last match computed cost / covariance on the device */
    bool mFinalOnDevice = false;        /* the last match ran the final matcher on the device */
    /* switch the handle's epilogue / refiner on for the match about to be made (and off after it) */
    void BeginDeviceStages(double covariance_scale);
    void EndDeviceStages();
    MetricSinkPtr mMetricSink;
    std::vector<std::int64_t> mResidentMaps;
};

/* scan_matcher_correlative.cpp:255-274: stepX = stepY = resolution,
 * stepTheta = acos(1 - 0.5 (resolution / maxRange)^2) */
void ComputeSearchStep(double resolution, const ScanData& scan,
                       double& step_x, double& step_y, double& step_theta);

class ScanMatcherCorrelative final : public ScanMatcher
{
public:
    ScanMatcherCorrelative(const std::string& name, const CostFuncPtr& cost, int low_resolution,
                           double range_x, double range_y, double range_theta,
                           const DeviceContextPtr& context);
    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& query) override;
    /* scan_matcher_correlative.hpp:75-84 */
    ScanMatchingSummary OptimizePose(const GridMapView& map, const ScanDataPtr& scan,
                                     const Pose2D& initial_pose, double score_threshold,
                                     double known_rate_threshold);
    int LowResolution() const { return mLowResolution; }
    const CostFuncPtr& Cost() const { return mCost; }

private:
    CostFuncPtr mCost;
    int mLowResolution;
    double mRangeX, mRangeY, mRangeTheta;
};

class ScanMatcherBranchBound final : public ScanMatcher
{
public:
    ScanMatcherBranchBound(const std::string& name, const CostFuncPtr& cost, int node_height_max,
                           double range_x, double range_y, double range_theta,
                           const DeviceContextPtr& context);
    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& query) override;
    ScanMatchingSummary OptimizePose(const GridMapView& map, const ScanDataPtr& scan,
                                     const Pose2D& initial_pose, double score_threshold,
                                     double known_rate_threshold);
    int NodeHeightMax() const { return mNodeHeightMax; }
    double RangeX() const { return mRangeX; }
    double RangeY() const { return mRangeY; }
    double RangeTheta() const { return mRangeTheta; }
    const CostFuncPtr& Cost() const { return mCost; }

private:
    CostFuncPtr mCost;
    int mNodeHeightMax;
    double mRangeX, mRangeY, mRangeTheta;
};

class ScanMatcherGridSearch final : public ScanMatcher
{
public:
    ScanMatcherGridSearch(const std::string& name, const CostFuncPtr& cost,
                          double range_x, double range_y, double range_theta,
                          double step_x, double step_y, double step_theta,
                          const DeviceContextPtr& context);
    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& query) override;
    ScanMatchingSummary OptimizePose(const GridMapView& map, const ScanDataPtr& scan,
                                     const Pose2D& initial_pose, double score_threshold,
                                     double known_rate_threshold);
    const CostFuncPtr& Cost() const { return mCost; }

private:
    CostFuncPtr mCost;
    double mRangeX, mRangeY, mRangeTheta, mStepX, mStepY, mStepTheta;
};

/* ScanMatcherLinearSolver (mapping/scan_matcher_linear_solver.cpp:46-170): the sub-pixel refiner
 * the reference runs after every coarse match ("FinalScanMatcherType": "LinearSolver",
 * launcher_settings_default.json:73-96). Damped Gauss-Newton on the square-error cost: per step
 * (H + lambda I) dx = r solved by a column-pivoting Householder QR, lambda halved / doubled by the
 * cost trend and -- like in the reference -- kept across calls. Runs on the CPU (SURVEY.md 8f
 * rank 1); O(iterations x beams). */
class ScanMatcherLinearSolver final : public ScanMatcher
{
public:
    ScanMatcherLinearSolver(const std::string& name, int num_of_iterations_max, double convergence_threshold,
                            double initial_lambda, const CostFuncPtr& cost);
    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& query) override;
    double Lambda() const { return mLambda; }

private:
    Pose2D OptimizeStep(const GridMapView& map, const ScanData& scan, const Pose2D& sensor_pose) const;

    int mNumOfIterationsMax;
    double mConvergenceThreshold;
    double mLambda;
    CostFuncPtr mCost;
};

/* ScanMatcherHillClimbing (scan_matcher_hill_climbing.cpp:45-170), the reference's other final
 * matcher: coordinate descent over (+-x, +-y, +-theta) steps on the cost function, steps halved
 * whenever no move improves the cost, until MaxNumOfRefinements halvings or MaxIterations. CPU. */
class ScanMatcherHillClimbing final : public ScanMatcher
{
public:
    /* any cost function: the square-error cost or the greedy-endpoint cost (the launcher default) */
    ScanMatcherHillClimbing(const std::string& name, double linear_step, double angular_step,
                            int max_iterations, int max_num_of_refinements,
                            const std::shared_ptr<CostFunction>& cost);
    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& query) override;
    int LastNumOfRefinements() const { return mLastNumOfRefinements; }

private:
    double mLinearStep, mAngularStep;
    int mMaxIterations, mMaxNumOfRefinements;
    std::shared_ptr<CostFunction> mCost;
    int mLastNumOfRefinements = 0;
};

/* x = A^-1 b for a 3x3 system with Eigen's ColPivHouseholderQR scheme (what
 * scan_matcher_linear_solver.cpp:161 calls): Householder reflections with the largest remaining
 * column brought to the front at every step. A is row-major. */
void SolveColPivHouseholderQr3(const double a[9], const double b[3], double x[3]);

} /* namespace csm_host */
