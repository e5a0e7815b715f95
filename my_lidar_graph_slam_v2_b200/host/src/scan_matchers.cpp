#include "csm_host/scan_matchers.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace csm_host {

DeviceContext::DeviceContext(int device) : mHandle(nullptr), mDevice(device)
{
    const int rc = csm_create(device, 0, &mHandle);
    if (rc != CSM_OK) {
        /* no CPU fallback: like a failed Assert in the reference (util.hpp:39-72) */
        std::fprintf(stderr, "csm_host: csm_create(device %d) failed with %d\n", device, rc);
        std::abort();
    }
}

DeviceContext::~DeviceContext()
{
    if (mHandle != nullptr) {
        csm_synchronize(mHandle);
        csm_destroy(mHandle);
    }
    if (mStaging != nullptr)
        csm_free_pinned(mStaging);
}

void* DeviceContext::Staging(std::size_t bytes)
{
    if (mStagingBytes < bytes) {
        if (mStaging != nullptr) {
            csm_synchronize(mHandle);
            csm_free_pinned(mStaging);
        }
        mStagingBytes = std::max<std::size_t>(bytes * 2, std::size_t(1) << 20);
        mStaging = csm_alloc_pinned(mStagingBytes);
        if (mStaging == nullptr) {
            std::fprintf(stderr, "csm_host: cannot allocate %zu bytes of page-locked memory\n", mStagingBytes);
            std::abort();
        }
    }
    return mStaging;
}

void DeviceContext::Check(int rc, const char* what) const
{
    if (rc == CSM_OK)
        return;
    std::fprintf(stderr, "csm_host: %s failed (%d): %s\n", what, rc, csm_last_error(mHandle));
    std::abort();
}

std::int64_t ScanMatcher::EnsureMap(const GridMapView& map)
{
    /* Anonymous maps (front-end latest map, rebuilt per scan,
     * lidar_graph_slam.cpp:233-240) are uploaded on every call under a private id */
    const std::int64_t id = map.map_id >= 0 ? map.map_id : (std::int64_t(1) << 40);
    if (map.device_resident) {
        /* built on the device (GridMapBuilderGPU): nothing to upload, but the CPU epilogue has no cells to read */
        if (!mContext->DeviceEpilogue() && !mContext->HasDeviceFinalMatcher() && map.values == nullptr) {
            std::fprintf(stderr, "csm_host: a device-resident map needs the device epilogue or final matcher\n");
            std::abort();
        }
        return map.map_id;
    }
    const bool resident = map.map_id >= 0 &&
        std::find(mResidentMaps.begin(), mResidentMaps.end(), id) != mResidentMaps.end();
    if (!resident) {
        /* through the context's page-locked staging area (the previous match on this context has
         * returned, so the area is free): the copy then runs as a DMA behind this call */
        if (map.blocks != nullptr || map.block_ptrs != nullptr) {
            const std::size_t data = (sizeof(std::uint16_t) * static_cast<std::size_t>(map.n_blocks))
                                     << (2 * map.log2_block_size);
            const std::size_t idx_off = (data + 15) & ~std::size_t(15);
            char* st = static_cast<char*>(mContext->Staging(idx_off + sizeof(std::int32_t) * map.n_blocks));
            if (map.block_ptrs != nullptr) {
                /* the reference's storage: one heap allocation per block (grid_map.cpp:522-535) */
                const std::size_t one = sizeof(std::uint16_t) << (2 * map.log2_block_size);
                for (int b = 0; b < map.n_blocks; ++b)
                    std::memcpy(st + one * static_cast<std::size_t>(b), map.block_ptrs[b], one);
            } else
            std::memcpy(st, map.blocks, data);
            std::memcpy(st + idx_off, map.block_index, sizeof(std::int32_t) * map.n_blocks);
            mContext->Check(csm_upload_grid_blocks(mContext->Handle(), id,
                                                   reinterpret_cast<const std::uint16_t*>(st),
                                                   reinterpret_cast<const std::int32_t*>(st + idx_off),
                                                   map.n_blocks, map.log2_block_size,
                                                   map.rows >> map.log2_block_size,
                                                   map.cols >> map.log2_block_size, map.resolution,
                                                   map.offset_x, map.offset_y), "csm_upload_grid_blocks");
        } else {
            const std::size_t bytes = sizeof(std::uint16_t) * static_cast<std::size_t>(map.rows) * map.cols;
            void* st = mContext->Staging(bytes);
            std::memcpy(st, map.values, bytes);
            mContext->Check(csm_upload_grid(mContext->Handle(), id, static_cast<const std::uint16_t*>(st),
                                            map.rows, map.cols, map.resolution, map.offset_x, map.offset_y),
                            "csm_upload_grid");
        }
        if (map.map_id >= 0)
            mResidentMaps.push_back(id);
    }
    return id;
}

void ScanMatcher::BeginDeviceStages(double covariance_scale)
{
    csm_handle h = mContext->Handle();
    mFinalOnDevice = mContext->HasDeviceFinalMatcher();
    mEpilogueOnDevice = !mFinalOnDevice && mContext->DeviceEpilogue();
    mContext->Check(csm_set_epilogue(h, mEpilogueOnDevice ? covariance_scale : 0.0), "csm_set_epilogue");
    if (mFinalOnDevice)
        mContext->Check(csm_set_refiner(h, &mContext->FinalMatcherParams()), "csm_set_refiner");
}

void ScanMatcher::EndDeviceStages()
{
    if (mFinalOnDevice)
        mContext->Check(csm_set_refiner(mContext->Handle(), nullptr), "csm_set_refiner");
}

void ScanMatcher::Epilogue(const GridMapView& map, const ScanData& scan, const Pose2D& best,
                           const CostFuncPtr& cost, ScanMatchingSummary& summary) const
{
    if (mContext && mFinalOnDevice) {
        /* the match ran the final matcher on the pose it found (csm_set_refiner) */
        csm_refined e;
        mContext->Check(csm_last_epilogue(mContext->Handle(), &e), "csm_last_epilogue");
        if (e.valid) {
            std::copy(e.covariance, e.covariance + 9, summary.estimated_covariance.begin());
            summary.normalized_cost = e.final_cost / static_cast<double>(scan.NumOfScans());
            summary.estimated_pose = MoveBackward(Pose2D { e.pose[0], e.pose[1], e.pose[2] },
                                                  scan.relative_sensor_pose);
            mContext->FinalMatcherParams().lambda = e.lambda;
            return;
        }
        /* no pose found: nothing was refined, fall through to the plain epilogue */
    }
    if (mContext && mContext->DeviceEpilogue() && mEpilogueOnDevice) {
        /* the match just made computed both on the device (csm_set_epilogue) */
        csm_refined e;
        mContext->Check(csm_last_epilogue(mContext->Handle(), &e), "csm_last_epilogue");
        std::copy(e.covariance, e.covariance + 9, summary.estimated_covariance.begin());
        summary.normalized_cost = e.final_cost / static_cast<double>(scan.NumOfScans());
        summary.estimated_pose = MoveBackward(best, scan.relative_sensor_pose);
        return;
    }
    double c = 0.0;
    summary.estimated_covariance = cost->CostAndCovariance(map, scan, best, c);
    summary.normalized_cost = c / static_cast<double>(scan.NumOfScans());
    summary.estimated_pose = MoveBackward(best, scan.relative_sensor_pose);
}

void ScanMatcher::ObserveSummary(const ScanMatchingSummary& s, const ScanData& scan, double micro) const
{
    if (!mMetricSink)
        return;
    Observe("OptimizationTime", micro);
    Observe("DiffTranslation", std::hypot(s.map_local_initial_pose.x - s.estimated_pose.x,
                                          s.map_local_initial_pose.y - s.estimated_pose.y));
    Observe("DiffRotation", std::fabs(s.map_local_initial_pose.theta - s.estimated_pose.theta));
    Observe("ScoreValue", s.normalized_score);
    Observe("CostValue", s.normalized_cost);
    Observe("NumOfScans", static_cast<double>(scan.NumOfScans()));
}

void ComputeSearchStep(double resolution, const ScanData& scan,
                       double& step_x, double& step_y, double& step_theta)
{
    const double max_range = *std::max_element(scan.ranges.begin(), scan.ranges.end());
    const double theta = resolution / max_range;
    step_x = resolution;
    step_y = resolution;
    step_theta = std::acos(1.0 - 0.5 * theta * theta);
}

namespace {

void FillFromDevice(const csm_result& r, ScanMatchingSummary& s)
{
    s.pose_found = r.found != 0;
    s.best_x = r.best_x; s.best_y = r.best_y; s.best_theta = r.best_t;
    s.sum_value = r.sum_value; s.n_known = r.n_known;
    s.normalized_score = r.normalized_score;
    s.flags = r.flags;
    s.n_processed = r.n_processed; s.n_ignored = r.n_ignored;
}

/* `for (d = -r; d <= r; d += s)`, scan_matcher_grid_search.cpp:118-120 */
std::vector<double> Offsets(double radius, double step)
{
    std::vector<double> out;
    for (double d = -radius; d <= radius; d += step)
        out.push_back(d);
    return out;
}

} /* namespace */

/* ---- real-time correlative ------------------------------------------------ */
ScanMatcherCorrelative::ScanMatcherCorrelative(
    const std::string& name, const CostFuncPtr& cost, int low_resolution,
    double range_x, double range_y, double range_theta, const DeviceContextPtr& context) :
    ScanMatcher(name, context), mCost(cost), mLowResolution(low_resolution),
    mRangeX(range_x), mRangeY(range_y), mRangeTheta(range_theta) { }

ScanMatchingSummary ScanMatcherCorrelative::OptimizePose(const ScanMatchingQuery& query)
{
    /* scan_matcher_correlative.cpp:92-115: thresholds 0 search the whole window */
    return OptimizePose(query.grid_map, query.scan_data, query.map_local_initial_pose, 0.0, 0.0);
}

ScanMatchingSummary ScanMatcherCorrelative::OptimizePose(
    const GridMapView& map, const ScanDataPtr& scan, const Pose2D& initial_pose,
    double score_threshold, double known_rate_threshold)
{
    csm_handle h = mContext->Handle();
    MicroTimer timer;
    const std::int64_t id = EnsureMap(map);
    mContext->Check(csm_build_coarse(h, id, mLowResolution), "csm_build_coarse");   /* ComputeCoarserMap */
    Observe("InputSetupTime", timer.ElapsedMicro());      /* enqueue time: the device work is asynchronous */
    timer.Start();
    const Pose2D sensor = Compound(initial_pose, scan->relative_sensor_pose);
    double sx, sy, st;
    ComputeSearchStep(map.resolution, *scan, sx, sy, st);
    const int win_x = static_cast<int>(std::ceil(0.5 * mRangeX / sx));
    const int win_y = static_cast<int>(std::ceil(0.5 * mRangeY / sy));
    const int win_t = static_cast<int>(std::ceil(0.5 * mRangeTheta / st));
    const double pose[3] = { sensor.x, sensor.y, sensor.theta };
    csm_result r;
    BeginDeviceStages(mCost->CovarianceScale());
    mContext->Check(csm_match_rt(h, id, scan->angles.data(), scan->ranges.data(),
                                 static_cast<int>(scan->NumOfScans()), pose, mLowResolution,
                                 win_x, win_y, win_t, sx, sy, st,
                                 score_threshold, known_rate_threshold, &r), "csm_match_rt");
    EndDeviceStages();
    ScanMatchingSummary s;
    FillFromDevice(r, s);
    s.map_local_initial_pose = initial_pose;
    const Pose2D best { sensor.x + r.best_x * sx, sensor.y + r.best_y * sy, sensor.theta + r.best_t * st };
    Epilogue(map, *scan, best, mCost, s);
    if (mMetricSink) {
        ObserveSummary(s, *scan, timer.ElapsedMicro());
        Observe("WinSizeX", win_x); Observe("WinSizeY", win_y); Observe("WinSizeTheta", win_t);
        Observe("StepSizeX", sx); Observe("StepSizeY", sy); Observe("StepSizeTheta", st);
        Observe("NumOfIgnoredNodes", r.n_ignored); Observe("NumOfProcessedNodes", r.n_processed);
    }
    return s;
}

/* ---- branch and bound --------------------------------------------------------- */
ScanMatcherBranchBound::ScanMatcherBranchBound(
    const std::string& name, const CostFuncPtr& cost, int node_height_max,
    double range_x, double range_y, double range_theta, const DeviceContextPtr& context) :
    ScanMatcher(name, context), mCost(cost), mNodeHeightMax(node_height_max),
    mRangeX(range_x), mRangeY(range_y), mRangeTheta(range_theta) { }

ScanMatchingSummary ScanMatcherBranchBound::OptimizePose(const ScanMatchingQuery& query)
{
    return OptimizePose(query.grid_map, query.scan_data, query.map_local_initial_pose, 0.0, 0.0);
}

ScanMatchingSummary ScanMatcherBranchBound::OptimizePose(
    const GridMapView& map, const ScanDataPtr& scan, const Pose2D& initial_pose,
    double score_threshold, double known_rate_threshold)
{
    csm_handle h = mContext->Handle();
    MicroTimer timer;
    const std::int64_t id = EnsureMap(map);
    mContext->Check(csm_build_pyramid(h, id, mNodeHeightMax), "csm_build_pyramid");  /* ComputeCoarserMaps */
    Observe("InputSetupTime", timer.ElapsedMicro());
    timer.Start();
    const Pose2D sensor = Compound(initial_pose, scan->relative_sensor_pose);
    double sx, sy, st;
    ComputeSearchStep(map.resolution, *scan, sx, sy, st);
    const int win_x = static_cast<int>(std::ceil(0.5 * mRangeX / sx));
    const int win_y = static_cast<int>(std::ceil(0.5 * mRangeY / sy));
    const int win_t = static_cast<int>(std::ceil(0.5 * mRangeTheta / st));
    const double pose[3] = { sensor.x, sensor.y, sensor.theta };
    csm_result r;
    BeginDeviceStages(mCost->CovarianceScale());
    mContext->Check(csm_match_bb(h, id, scan->angles.data(), scan->ranges.data(),
                                 static_cast<int>(scan->NumOfScans()), pose, mNodeHeightMax,
                                 win_x, win_y, win_t, sx, sy, st,
                                 score_threshold, known_rate_threshold, &r), "csm_match_bb");
    EndDeviceStages();
    ScanMatchingSummary s;
    FillFromDevice(r, s);
    s.map_local_initial_pose = initial_pose;
    const Pose2D best { sensor.x + sx * r.best_x, sensor.y + sy * r.best_y, sensor.theta + st * r.best_t };
    Epilogue(map, *scan, best, mCost, s);
    if (mMetricSink) {
        /* the node counts are those of the level-synchronous sweep, not of the reference's
         * best-first order (they measure work, the result does not depend on them) */
        ObserveSummary(s, *scan, timer.ElapsedMicro());
        Observe("WinSizeX", win_x); Observe("WinSizeY", win_y); Observe("WinSizeTheta", win_t);
        Observe("StepSizeX", sx); Observe("StepSizeY", sy); Observe("StepSizeTheta", st);
        Observe("NumOfIgnoredNodes", r.n_ignored); Observe("NumOfProcessedNodes", r.n_processed);
    }
    return s;
}

/* ---- grid search ----------------------------------------------------------------- */
ScanMatcherGridSearch::ScanMatcherGridSearch(
    const std::string& name, const CostFuncPtr& cost, double range_x, double range_y,
    double range_theta, double step_x, double step_y, double step_theta,
    const DeviceContextPtr& context) :
    ScanMatcher(name, context), mCost(cost), mRangeX(range_x), mRangeY(range_y),
    mRangeTheta(range_theta), mStepX(step_x), mStepY(step_y), mStepTheta(step_theta) { }

ScanMatchingSummary ScanMatcherGridSearch::OptimizePose(const ScanMatchingQuery& query)
{
    return OptimizePose(query.grid_map, query.scan_data, query.map_local_initial_pose, 0.0, 0.0);
}

ScanMatchingSummary ScanMatcherGridSearch::OptimizePose(
    const GridMapView& map, const ScanDataPtr& scan, const Pose2D& initial_pose,
    double score_threshold, double known_rate_threshold)
{
    csm_handle h = mContext->Handle();
    MicroTimer timer;
    const std::int64_t id = EnsureMap(map);
    const Pose2D sensor = Compound(initial_pose, scan->relative_sensor_pose);
    const std::vector<double> dx = Offsets(mRangeX / 2.0, mStepX);
    const std::vector<double> dy = Offsets(mRangeY / 2.0, mStepY);
    const std::vector<double> dt = Offsets(mRangeTheta / 2.0, mStepTheta);
    const double pose[3] = { sensor.x, sensor.y, sensor.theta };
    csm_result r;
    mEpilogueOnDevice = false;          /* the grid search keeps the CPU epilogue (4 ms of search per match) */
    mFinalOnDevice = false;
    mContext->Check(csm_match_grid(h, id, scan->angles.data(), scan->ranges.data(),
                                   static_cast<int>(scan->NumOfScans()), pose,
                                   dx.data(), static_cast<int>(dx.size()),
                                   dy.data(), static_cast<int>(dy.size()),
                                   dt.data(), static_cast<int>(dt.size()),
                                   score_threshold, known_rate_threshold, &r), "csm_match_grid");
    ScanMatchingSummary s;
    FillFromDevice(r, s);
    s.map_local_initial_pose = initial_pose;
    Pose2D best = sensor;     /* reference: bestSensorPose starts at the sensor pose (:113) */
    if (r.found)
        best = Pose2D { sensor.x + dx[r.best_x], sensor.y + dy[r.best_y], sensor.theta + dt[r.best_t] };
    Epilogue(map, *scan, best, mCost, s);
    if (mMetricSink) {
        /* NumOfScoreUpdates (how often the sequential loop raised its maximum,
         * scan_matcher_grid_search.cpp:136) has no parallel counterpart and is not reported */
        ObserveSummary(s, *scan, timer.ElapsedMicro());
        Observe("NumOfScoreEvaluations", static_cast<double>(dx.size()) * dy.size() * dt.size());
    }
    return s;
}

/* ---- linear-solver refiner (CPU) ------------------------------------------------------ */
void SolveColPivHouseholderQr3(const double a_in[9], const double b_in[3], double x[3])
{
    double a[3][3], b[3] = { b_in[0], b_in[1], b_in[2] };
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            a[r][c] = a_in[r * 3 + c];
    int perm[3] = { 0, 1, 2 };
    int rank = 3;
    double max_pivot = 0.0;
    for (int k = 0; k < 3; ++k) {
        /* column with the largest remaining norm comes first */
        int best = k;
        double best_norm = -1.0;
        for (int c = k; c < 3; ++c) {
            double s = 0.0;
            for (int r = k; r < 3; ++r) s += a[r][c] * a[r][c];
            if (s > best_norm) { best_norm = s; best = c; }
        }
        if (best != k) {
            for (int r = 0; r < 3; ++r) std::swap(a[r][k], a[r][best]);
            std::swap(perm[k], perm[best]);
        }
        /* Householder reflector H = I - tau v v^T that maps a[k..2][k] onto (beta, 0, 0) */
        double tail = 0.0;
        for (int r = k + 1; r < 3; ++r) tail += a[r][k] * a[r][k];
        const double c0 = a[k][k];
        double beta = c0, tau = 0.0, v[3] = { 0.0, 0.0, 0.0 };
        if (tail > 0.0) {
            beta = std::sqrt(c0 * c0 + tail);
            if (c0 >= 0.0) beta = -beta;
            for (int r = k + 1; r < 3; ++r) v[r] = a[r][k] / (c0 - beta);
            v[k] = 1.0;
            tau = (beta - c0) / beta;
        }
        a[k][k] = beta;
        for (int r = k + 1; r < 3; ++r) a[r][k] = 0.0;
        if (tau != 0.0) {
            for (int c = k + 1; c < 3; ++c) {
                double dot = 0.0;
                for (int r = k; r < 3; ++r) dot += v[r] * a[r][c];
                for (int r = k; r < 3; ++r) a[r][c] -= tau * v[r] * dot;
            }
            double dot = 0.0;
            for (int r = k; r < 3; ++r) dot += v[r] * b[r];
            for (int r = k; r < 3; ++r) b[r] -= tau * v[r] * dot;
        }
        max_pivot = std::max(max_pivot, std::fabs(beta));
    }
    /* numerical rank as Eigen decides it: pivots above epsilon * size * largest pivot */
    const double threshold = 2.220446049250313e-16 * 3.0 * max_pivot;
    rank = 0;
    for (int k = 0; k < 3; ++k)
        if (std::fabs(a[k][k]) > threshold) ++rank;
    double y[3] = { 0.0, 0.0, 0.0 };
    for (int k = rank - 1; k >= 0; --k) {
        double s = b[k];
        for (int c = k + 1; c < rank; ++c) s -= a[k][c] * y[c];
        y[k] = s / a[k][k];
    }
    for (int k = 0; k < 3; ++k) x[perm[k]] = y[k];
}

ScanMatcherLinearSolver::ScanMatcherLinearSolver(
    const std::string& name, int num_of_iterations_max, double convergence_threshold,
    double initial_lambda, const CostFuncPtr& cost) :
    ScanMatcher(name, nullptr), mNumOfIterationsMax(num_of_iterations_max),
    mConvergenceThreshold(convergence_threshold), mLambda(initial_lambda), mCost(cost) { }

Pose2D ScanMatcherLinearSolver::OptimizeStep(const GridMapView& map, const ScanData& scan,
                                             const Pose2D& sensor_pose) const
{
    /* scan_matcher_linear_solver.cpp:143-170 */
    double h[9], r[3], d[3];
    mCost->ComputeHessianAndResidual(map, scan, sensor_pose, h, r);
    h[0] += mLambda; h[4] += mLambda; h[8] += mLambda;
    SolveColPivHouseholderQr3(h, r, d);
    return Pose2D { sensor_pose.x + d[0], sensor_pose.y + d[1], sensor_pose.theta + d[2] };
}

ScanMatchingSummary ScanMatcherLinearSolver::OptimizePose(const ScanMatchingQuery& query)
{
    /* scan_matcher_linear_solver.cpp:66-140 */
    const GridMapView& map = query.grid_map;
    const ScanData& scan = *query.scan_data;
    MicroTimer timer;
    const Pose2D sensor = Compound(query.map_local_initial_pose, scan.relative_sensor_pose);
    const double initial_cost = mCost->Cost(map, scan, sensor);
    double prev_cost = initial_cost, cost = 0.0;
    Pose2D best = sensor;
    int iterations = 0;
    while (true) {
        best = OptimizeStep(map, scan, best);
        cost = mCost->Cost(map, scan, best);
        if (++iterations >= mNumOfIterationsMax || std::fabs(prev_cost - cost) < mConvergenceThreshold)
            break;
        if (cost < prev_cost)
            mLambda = std::max(1e-8, mLambda * 0.5);
        else
            mLambda = std::min(1e-4, mLambda * 2.0);
        prev_cost = cost;
    }
    ScanMatchingSummary s;
    s.pose_found = true;
    s.normalized_cost = cost / static_cast<double>(scan.NumOfScans());
    s.map_local_initial_pose = query.map_local_initial_pose;
    s.estimated_pose = MoveBackward(best, scan.relative_sensor_pose);
    s.estimated_covariance = mCost->ComputeCovariance(map, scan, best);
    s.n_processed = iterations;
    if (mMetricSink) {
        /* scan_matcher_linear_solver.cpp:125-133 */
        Observe("OptimizationTime", timer.ElapsedMicro());
        Observe("DiffTranslation", std::hypot(s.map_local_initial_pose.x - s.estimated_pose.x,
                                              s.map_local_initial_pose.y - s.estimated_pose.y));
        Observe("DiffRotation", std::fabs(s.map_local_initial_pose.theta - s.estimated_pose.theta));
        Observe("NumOfIterations", iterations);
        Observe("InitialCost", initial_cost / static_cast<double>(scan.NumOfScans()));
        Observe("FinalCost", s.normalized_cost);
        Observe("NumOfScans", static_cast<double>(scan.NumOfScans()));
    }
    return s;
}

/* ---- hill-climbing refiner (CPU) ---------------------------------------------------------- */
ScanMatcherHillClimbing::ScanMatcherHillClimbing(
    const std::string& name, double linear_step, double angular_step, int max_iterations,
    int max_num_of_refinements, const std::shared_ptr<CostFunction>& cost) :
    ScanMatcher(name, nullptr), mLinearStep(linear_step), mAngularStep(angular_step),
    mMaxIterations(max_iterations), mMaxNumOfRefinements(max_num_of_refinements), mCost(cost) { }

ScanMatchingSummary ScanMatcherHillClimbing::OptimizePose(const ScanMatchingQuery& query)
{
    /* scan_matcher_hill_climbing.cpp:63-170 */
    static const double move_x[] = { 1.0, -1.0, 0.0, 0.0, 0.0, 0.0 };
    static const double move_y[] = { 0.0, 0.0, 1.0, -1.0, 0.0, 0.0 };
    static const double move_t[] = { 0.0, 0.0, 0.0, 0.0, 1.0, -1.0 };
    const GridMapView& map = query.grid_map;
    const ScanData& scan = *query.scan_data;
    MicroTimer timer;
    const Pose2D sensor = Compound(query.map_local_initial_pose, scan.relative_sensor_pose);
    const double initial_cost = mCost->Cost(map, scan, sensor);
    double min_cost = initial_cost;
    Pose2D best = sensor;
    int iterations = 0, refinements = 0;
    double linear = mLinearStep, angular = mAngularStep;
    bool updated = false;
    do {
        double min_local = min_cost;
        Pose2D best_local = best;
        updated = false;
        for (int i = 0; i < 6; ++i) {
            Pose2D p = best;
            p.x += move_x[i] * linear;
            p.y += move_y[i] * linear;
            p.theta += move_t[i] * angular;
            const double c = mCost->Cost(map, scan, p);
            if (c < min_local) {
                min_local = c;
                best_local = p;
                updated = true;
            }
        }
        if (updated) {
            min_cost = min_local;
            best = best_local;
        } else {
            ++refinements;
            linear *= 0.5;
            angular *= 0.5;
        }
    } while ((updated || refinements < mMaxNumOfRefinements) && (++iterations < mMaxIterations));
    mLastNumOfRefinements = refinements;
    ScanMatchingSummary s;
    s.pose_found = true;
    s.normalized_cost = min_cost / static_cast<double>(scan.NumOfScans());
    s.map_local_initial_pose = query.map_local_initial_pose;
    s.estimated_pose = MoveBackward(best, scan.relative_sensor_pose);
    s.estimated_covariance = mCost->ComputeCovariance(map, scan, best);
    s.n_processed = iterations;
    if (mMetricSink) {
        /* scan_matcher_hill_climbing.cpp:153-163 */
        Observe("OptimizationTime", timer.ElapsedMicro());
        Observe("DiffTranslation", std::hypot(s.map_local_initial_pose.x - s.estimated_pose.x,
                                              s.map_local_initial_pose.y - s.estimated_pose.y));
        Observe("DiffRotation", std::fabs(s.map_local_initial_pose.theta - s.estimated_pose.theta));
        Observe("NumOfIterations", iterations);
        Observe("NumOfRefinements", refinements);
        Observe("InitialCost", initial_cost / static_cast<double>(scan.NumOfScans()));
        Observe("FinalCost", s.normalized_cost);
        Observe("NumOfScans", static_cast<double>(scan.NumOfScans()));
    }
    return s;
}

} /* namespace csm_host */
