/* slam_pipeline.hpp -- the full SLAM loop around the GPU hot path (BASELINE.json configs[4], SURVEY.md 8f
 * ranks 3 and 4): the reference's LidarGraphSlam with its front end and back end
 * (mapping/lidar_graph_slam.cpp, lidar_graph_slam_frontend.cpp:109-330, lidar_graph_slam_backend.cpp:92-198)
 * in one class, every matcher and every map on the device:
 *
 *   per scan   ProcessScan: odometry thresholds -> UpdateLatestMap (device) -> real-time correlative match
 *              + final matcher on the latest map (device) -> degeneration check -> AppendScan: pose graph
 *              node and edges, scan cast into the current local map (device)
 *   every LoopDetectionThreshold metres   RunBackendStep: GetLoopSearchHint -> LoopSearcherNearest::Search
 *              -> GetLoopDetectionQueries -> LoopDetectorBranchBound::Detect (one device batch over the
 *              resident local maps) -> AppendLoopClosingEdges -> PoseGraphOptimizer::Optimize (the seam,
 *              CPU) -> AfterLoopClosure
 *
 * What differs from the reference on purpose: the back end runs inside ProcessScan instead of on a worker
 * thread (the reference's front end waits for a running optimisation anyway, :192-193; results then do not
 * depend on thread timing), and the scan filters of the front end (outlier filter, interpolator, accumulator,
 * :165-175) are not part of this path: scans arrive as they are matched.
 * The optimiser is behind PoseGraphOptimizer (pose_graph.hpp); g2o is not in this package.
 */
#pragma once

#include <memory>
#include <vector>

#include "csm_host/carmen_log.hpp"
#include "csm_host/loop_detector.hpp"
#include "csm_host/loop_searcher.hpp"
#include "csm_host/map_builder.hpp"
#include "csm_host/pose_graph.hpp"

namespace csm_host {

/* launcher_settings_default.json, the groups this path reads; defaults are the reference's */
struct SlamSettings
{
    /* GridMapBuilder (:175-187) */
    double resolution = 0.05;
    int patch_size = 16;
    int scans_for_latest_map = 10;
    double local_map_travel_dist = 2.5;
    int overlapped_scans = 10;
    double usable_range_min = 0.01, usable_range_max = 20.0;
    double prob_hit = 0.62, prob_miss = 0.46;
    /* Frontend (:341-385) */
    double update_travel_dist = 0.5, update_angle = 0.5, update_time = 5.0;
    double loop_detection_threshold = 2.5;
    double degeneration_threshold = 10.0;
    double odometry_covariance_scale = 1e2;
    bool fuse_odometry_covariance = false;
    Pose2D initial_pose;
    /* ScanMatcherRealTimeCorrelative (:37-45) and the final matcher (:364-371), CostSquareError (:11-13) */
    int rt_low_resolution = 5;
    double rt_range_x = 0.25, rt_range_y = 0.25, rt_range_theta = 0.5;
    int final_iterations = 10;
    double final_convergence = 1e-4, final_lambda = 1e-4;
    double covariance_scale = 1e4;
    /* LoopSearcherNearest (:61-65) */
    double searcher_travel_dist = 10.0, searcher_node_dist = 5.0;
    int searcher_candidates = 2;
    /* LoopDetectorBranchBound (:128-156) */
    int bb_node_height_max = 6;
    double bb_range_x = 2.5, bb_range_y = 2.5, bb_range_theta = 0.5;
    double score_threshold = 0.55, known_rate_threshold = 0.6;
    /* true: the final matchers (front end and loop detector) are the CPU twins of ScanMatcherLinearSolver
     * on a host copy of the map (downloaded per scan / per finished local map): bit-identical to the
     * reference, for parity runs. false: they run on the device behind the search (1e-9 relative). */
    bool host_final_matchers = false;
};

struct SlamCounters
{
    int scans_in = 0, scans_processed = 0;
    int backend_steps = 0, backend_steps_with_candidates = 0;
    int loop_queries = 0, loops_detected = 0, optimizations = 0;
    int degenerations = 0;
    /* seconds */
    double t_latest_map = 0.0, t_match = 0.0, t_append = 0.0, t_backend = 0.0, t_detect = 0.0;
};

class SlamPipeline
{
public:
    SlamPipeline(const DeviceContextPtr& context, const SlamSettings& settings,
                 const std::shared_ptr<PoseGraphOptimizer>& optimizer = nullptr);

    /* LidarGraphSlamFrontend::ProcessScan: returns whether the scan was used (thresholds passed) */
    bool ProcessScan(const ScanDataPtr& scan, const Pose2D& odom_pose, double time_stamp);
    /* the last back-end iteration after the front end has finished (lidar_graph_slam_backend.cpp:83-89) */
    void Finish();
    /* slam_launcher.cpp:262-283: every scan record of a Carmen log goes to ProcessScan with the odometry pose and
     * time stamp it carries (odometry records are not consumed there either); returns the scans used */
    int RunLog(const std::vector<CarmenRecord>& records, bool finish = true);
    /* the reference's metric ids: "Frontend.*" (lidar_graph_slam_frontend.cpp:34-63, 151-320), "Backend.*"
     * (lidar_graph_slam_backend.cpp:29-56, 106-197) and those of the matchers and the detector; with
     * SaveMetrics (carmen_log.hpp) this gives the launcher's <output>.metric.json */
    void SetMetricSink(const MetricSinkPtr& sink);

    const PoseGraph& Graph() const { return mPoseGraph; }
    const GridMapBuilderGPU& Builder() const { return *mBuilder; }
    const SlamCounters& Counters() const { return mCounters; }
    const std::vector<LoopDetectionResult>& Loops() const { return mLoops; }
    /* the scan-matching summaries of the processed scans after the first (final matcher's output) */
    const std::vector<ScanMatchingSummary>& Matches() const { return mMatches; }

    /* the steps of the back end, public so that tests can chain them by hand
     * (lidar_graph_slam.cpp:273-415, 448-504, 106-194, 506-672) */
    LoopSearchHint GetLoopSearchHint() const;
    std::vector<LoopDetectionQuery> GetLoopDetectionQueries(const std::vector<LoopCandidate>& candidates);
    void AppendLoopClosingEdges(const std::vector<LoopDetectionResult>& results);
    void GetPoseGraphForOptimization(std::vector<int>& local_map_ids, std::vector<std::array<double, 3>>& local_map_poses,
                                     std::vector<int>& scan_node_ids, std::vector<std::array<double, 3>>& scan_poses,
                                     std::vector<EdgePose>& edges) const;
    void AfterLoopClosure(const std::vector<int>& local_map_ids, const std::vector<std::array<double, 3>>& local_map_poses,
                          const std::vector<int>& scan_node_ids, const std::vector<std::array<double, 3>>& scan_poses);
    void RunBackendStep();

private:
    bool CheckDegeneration(const Mat3& covariance) const;
    Mat3 ComputeOdometryCovariance(const Pose2D& relative_pose, double elapsed) const;
    void FuseOdometry(const Pose2D& odom_rel, const Mat3& odom_cov, const Pose2D& scan_rel, const Mat3& scan_cov,
                      Pose2D& fused_rel, Mat3& fused_cov) const;
    /* host copy of a device map (parity runs): cells and block allocation */
    struct HostCopy { std::vector<std::uint16_t> cells; std::vector<std::uint8_t> allocation; };
    GridMapView WithHostCopy(const DeviceGridMap& map, HostCopy& copy) const;

    DeviceContextPtr mContext;
    SlamSettings mSettings;
    std::shared_ptr<PoseGraphOptimizer> mOptimizer;
    PoseGraph mPoseGraph;
    std::unique_ptr<GridMapBuilderGPU> mBuilder;
    std::shared_ptr<ScanMatcherCorrelative> mScanMatcher;
    std::shared_ptr<ScanMatcherLinearSolver> mHostFinalMatcher;         /* parity runs */
    std::shared_ptr<ScanMatcherLinearSolver> mHostLoopFinalMatcher;
    std::unique_ptr<LoopSearcherNearest> mLoopSearcher;
    std::shared_ptr<LoopDetectorBranchBound> mLoopDetector;
    std::vector<std::unique_ptr<HostCopy>> mLocalMapCopies;             /* by LocalMapId, parity runs */
    SlamCounters mCounters;
    MetricSinkPtr mMetricSink;
    void Observe(const char* id, double value) const { if (mMetricSink) mMetricSink->Observe(id, value); }
    std::vector<LoopDetectionResult> mLoops;
    std::vector<ScanMatchingSummary> mMatches;
    /* front-end state (lidar_graph_slam_frontend.hpp) */
    int mProcessCount = 0;
    Pose2D mLastOdomPose, mLastMapUpdateOdomPose;
    double mAccumulatedTravelDist = 0.0, mAccumulatedAngle = 0.0;
    double mLastMapUpdateTime = 0.0, mLastLoopDetectionDist = 0.0;
};

} /* namespace csm_host */
