#include "csm_host/slam_pipeline.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace csm_host {

namespace {

double Seconds()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

constexpr int kInvalidId = -1;          /* NodeId::Invalid / LocalMapId::Invalid */

} /* namespace */

SlamPipeline::SlamPipeline(const DeviceContextPtr& context, const SlamSettings& settings,
                           const std::shared_ptr<PoseGraphOptimizer>& optimizer) :
    mContext(context), mSettings(settings),
    mOptimizer(optimizer ? optimizer : std::make_shared<PoseGraphOptimizerIdentity>())
{
    const SlamSettings& s = mSettings;
    mBuilder.reset(new GridMapBuilderGPU(mContext, s.resolution, s.patch_size, s.scans_for_latest_map,
                                         s.usable_range_min, s.usable_range_max, s.prob_hit, s.prob_miss));
    mBuilder->SetLocalMapPolicy(s.local_map_travel_dist, s.overlapped_scans);
    const auto cost = std::make_shared<CostSquareError>(s.covariance_scale);
    if (s.host_final_matchers) {
        mContext->SetDeviceEpilogue(false);
        mContext->SetDeviceFinalMatcher(0, 0.0, 0.0, 0.0);
        mHostFinalMatcher = std::make_shared<ScanMatcherLinearSolver>(
            "LocalSlam.FinalScanMatcherLinearSolver", s.final_iterations, s.final_convergence, s.final_lambda, cost);
        mHostLoopFinalMatcher = std::make_shared<ScanMatcherLinearSolver>(
            "LoopDetector.FinalScanMatcherLinearSolver", s.final_iterations, s.final_convergence, s.final_lambda, cost);
    } else {
        mContext->SetDeviceFinalMatcher(s.final_iterations, s.final_convergence, s.final_lambda, s.covariance_scale);
    }
    mScanMatcher = std::make_shared<ScanMatcherCorrelative>("LocalSlam.ScanMatcherCorrelative", cost, s.rt_low_resolution,
                                                            s.rt_range_x, s.rt_range_y, s.rt_range_theta, mContext);
    mLoopSearcher.reset(new LoopSearcherNearest(s.searcher_travel_dist, s.searcher_node_dist, s.searcher_candidates));
    const auto bb = std::make_shared<ScanMatcherBranchBound>("LoopDetector.ScanMatcherBranchBound", cost, s.bb_node_height_max,
                                                             s.bb_range_x, s.bb_range_y, s.bb_range_theta, mContext);
    FinalMatcher loop_final;
    if (s.host_final_matchers)
        loop_final = MakeLinearSolverFinalMatcher(mHostLoopFinalMatcher);
    mLoopDetector = std::make_shared<LoopDetectorBranchBound>("LoopDetector.BranchBound", bb, loop_final,
                                                              s.score_threshold, s.known_rate_threshold);
    if (!s.host_final_matchers)
        mLoopDetector->UseDeviceRefiner(s.final_iterations, s.final_convergence, s.final_lambda, s.covariance_scale);
}

GridMapView SlamPipeline::WithHostCopy(const DeviceGridMap& map, HostCopy& copy) const
{
    const int bs = map.BlockSize();
    copy.cells.resize(static_cast<std::size_t>(map.Rows()) * map.Cols());
    copy.allocation.resize(static_cast<std::size_t>(map.Rows() / bs) * (map.Cols() / bs));
    csm_handle h = mContext->Handle();
    mContext->Check(csm_download_level(h, map.MapId(), 0, copy.cells.data()), "csm_download_level");
    mContext->Check(csm_map_download_allocation(h, map.MapId(), copy.allocation.data()), "csm_map_download_allocation");
    GridMapView v = map.View();
    v.values = copy.cells.data();
    v.block_allocated = copy.allocation.data();
    return v;
}

bool SlamPipeline::ProcessScan(const ScanDataPtr& scan, const Pose2D& odom_pose, double time_stamp)
{
    const SlamSettings& s = mSettings;
    /* lidar_graph_slam_frontend.cpp:117-159 */
    const Pose2D rel_odom = (mProcessCount == 0) ? Pose2D {} : InverseCompound(mLastOdomPose, odom_pose);
    mLastOdomPose = odom_pose;
    mAccumulatedTravelDist += Distance(rel_odom);
    mAccumulatedAngle += std::fabs(rel_odom.theta);
    const double elapsed = (mProcessCount == 0) ? 0.0 : time_stamp - mLastMapUpdateTime;
    const bool first = mProcessCount == 0;
    const bool needed = (mAccumulatedTravelDist >= s.update_travel_dist || mAccumulatedAngle >= s.update_angle ||
                         elapsed >= s.update_time || first) && (elapsed >= 0.0);
    ++mCounters.scans_in;
    const double t_process = Seconds();
    if (!needed) {
        Observe("Frontend.ProcessTime", (Seconds() - t_process) * 1e6);
        return false;
    }

    if (first) {
        /* AppendFirstNodeAndEdge (lidar_graph_slam.cpp:419-437): a tight covariance pins the first node */
        const Mat3 covariance { 1e-9, 0.0, 0.0, 0.0, 1e-9, 0.0, 0.0, 0.0, 1e-9 };
        const double t0 = Seconds();
        mBuilder->AppendScan(mPoseGraph, s.initial_pose, covariance, scan);
        mCounters.t_append += Seconds() - t0;
    } else {
        /* GetLatestData (lidar_graph_slam.cpp:224-270): the latest map is rebuilt from the last scans */
        double t0 = Seconds();
        mBuilder->UpdateLatestMap(mPoseGraph.scan_nodes);
        const Pose2D latest_scan_pose = mPoseGraph.scan_nodes.back().global_pose;
        const Pose2D latest_map_pose = mBuilder->LatestMapPose();
        double t1 = Seconds();
        mCounters.t_latest_map += t1 - t0;
        Observe("Frontend.ScanDataSetupTime", (t1 - t0) * 1e6);

        /* :205-230 */
        const Pose2D rel_from_last_update = InverseCompound(mLastMapUpdateOdomPose, odom_pose);
        const Pose2D initial_pose = Compound(latest_scan_pose, rel_from_last_update);
        const Pose2D map_local_initial_pose = InverseCompound(latest_map_pose, initial_pose);
        ScanMatchingSummary fin;
        if (s.host_final_matchers) {
            HostCopy copy;
            const GridMapView view = WithHostCopy(mBuilder->LatestGrid(), copy);
            const ScanMatchingSummary coarse = mScanMatcher->OptimizePose(ScanMatchingQuery { view, scan, map_local_initial_pose });
            fin = mHostFinalMatcher->OptimizePose(ScanMatchingQuery { view, scan, coarse.estimated_pose });
        } else {
            /* the device runs the final matcher behind the search: one summary for both */
            const GridMapView view = mBuilder->LatestMap();
            fin = mScanMatcher->OptimizePose(ScanMatchingQuery { view, scan, map_local_initial_pose });
        }
        if (!fin.pose_found) {
            std::fprintf(stderr, "csm_host: the front-end scan matcher found no pose\n");
            std::abort();
        }
        mMatches.push_back(fin);
        double t2 = Seconds();
        mCounters.t_match += t2 - t1;
        /* with the final matcher on the device both stages are one submission: its time is reported as the
         * scan matching time and the final stage as 0 */
        Observe("Frontend.ScanMatchingTime", (t2 - t1) * 1e6);
        Observe("Frontend.FinalScanMatchingTime", 0.0);

        /* :232-283 */
        const Pose2D global_estimated_pose = Compound(latest_map_pose, fin.estimated_pose);
        const Pose2D scan_relative_pose = InverseCompound(latest_scan_pose, global_estimated_pose);
        const Mat3 scan_covariance = ConvertCovarianceFromLocalToWorld(latest_map_pose, fin.estimated_covariance);
        Pose2D relative_pose = scan_relative_pose;
        Mat3 covariance = scan_covariance;
        if (CheckDegeneration(scan_covariance)) {
            ++mCounters.degenerations;
            const Mat3 odom_covariance = ComputeOdometryCovariance(rel_from_last_update, elapsed);
            if (s.fuse_odometry_covariance)
                FuseOdometry(rel_from_last_update, odom_covariance, scan_relative_pose, scan_covariance, relative_pose, covariance);
            else {
                relative_pose = rel_from_last_update;
                covariance = odom_covariance;
            }
        }
        mBuilder->AppendScan(mPoseGraph, relative_pose, covariance, scan);
        double t3 = Seconds();
        mCounters.t_append += t3 - t2;
        Observe("Frontend.DataUpdateTime", (t3 - t2) * 1e6);

        /* :289-303: the back end is notified every LoopDetectionThreshold metres */
        const double accum = mBuilder->AccumTravelDist();
        if (accum - mLastLoopDetectionDist >= s.loop_detection_threshold) {
            mLastLoopDetectionDist = accum;
            RunBackendStep();
            mCounters.t_backend += Seconds() - t3;
        }
    }
    ++mCounters.scans_processed;
    mProcessCount += 1;
    if (mMetricSink) {
        const double micro = (Seconds() - t_process) * 1e6;
        Observe("Frontend.ProcessTime", micro);
        Observe("Frontend.ProcessScanTime", micro);
        Observe("Frontend.IntervalTravelDist", mAccumulatedTravelDist);
        Observe("Frontend.IntervalAngle", mAccumulatedAngle);
        Observe("Frontend.IntervalTime", elapsed);
        Observe("Frontend.NumOfScans", static_cast<double>(scan->NumOfScans()));
        Observe("Frontend.ProcessFrame", static_cast<double>(mProcessCount - 1));
    }
    mAccumulatedTravelDist = 0.0;
    mAccumulatedAngle = 0.0;
    mLastMapUpdateOdomPose = odom_pose;
    mLastMapUpdateTime = time_stamp;
    return true;
}

void SlamPipeline::Finish()
{
    const double t0 = Seconds();
    RunBackendStep();
    mCounters.t_backend += Seconds() - t0;
}

bool SlamPipeline::CheckDegeneration(const Mat3& c) const
{
    /* :334-348: ratio of the eigenvalues of the translational 2 x 2 block */
    const double a = c[0], b = c[1], cc = c[3], d = c[4];
    const double mean = 0.5 * (a + d), disc = 0.25 * (a - d) * (a - d) + b * cc;
    const double root = disc > 0.0 ? std::sqrt(disc) : 0.0;       /* complex pair: equal real parts */
    const double lo = mean - root, hi = mean + root;
    return hi / lo > mSettings.degeneration_threshold;
}

Mat3 SlamPipeline::ComputeOdometryCovariance(const Pose2D& relative_pose, double elapsed) const
{
    /* :351-368 */
    const double trans = std::max(1e-1, Distance(relative_pose) / elapsed);
    const double rot = std::max(1e-1, relative_pose.theta / elapsed);
    const double k = mSettings.odometry_covariance_scale;
    return Mat3 { trans * trans * k, 0.0, 0.0, 0.0, trans * trans * k, 0.0, 0.0, 0.0, rot * rot * k };
}

void SlamPipeline::FuseOdometry(const Pose2D& odom_rel, const Mat3& odom_cov, const Pose2D& scan_rel, const Mat3& scan_cov,
                                Pose2D& fused_rel, Mat3& fused_cov) const
{
    /* :371-411 */
    const Mat3 inv_odom = Inverse(odom_cov), inv_scan = Inverse(scan_cov);
    Mat3 inv_fused;
    for (int i = 0; i < 9; ++i) inv_fused[i] = inv_odom[i] + inv_scan[i];
    fused_cov = Inverse(inv_fused);
    const double odom_theta = NormalizeAngle(odom_rel.theta), scan_theta = NormalizeAngle(scan_rel.theta);
    const double diff = scan_theta - odom_theta;
    const double fixed_odom_theta = diff > kPi ? odom_theta + 2.0 * kPi : diff < -kPi ? odom_theta - 2.0 * kPi : odom_theta;
    const double o[3] = { odom_rel.x, odom_rel.y, fixed_odom_theta }, c[3] = { scan_rel.x, scan_rel.y, scan_theta };
    double w[3];
    for (int i = 0; i < 3; ++i)
        w[i] = (inv_odom[i * 3] * o[0] + inv_odom[i * 3 + 1] * o[1] + inv_odom[i * 3 + 2] * o[2]) +
               (inv_scan[i * 3] * c[0] + inv_scan[i * 3 + 1] * c[1] + inv_scan[i * 3 + 2] * c[2]);
    fused_rel.x = fused_cov[0] * w[0] + fused_cov[1] * w[1] + fused_cov[2] * w[2];
    fused_rel.y = fused_cov[3] * w[0] + fused_cov[4] * w[1] + fused_cov[5] * w[2];
    fused_rel.theta = NormalizeAngle(fused_cov[6] * w[0] + fused_cov[7] * w[1] + fused_cov[8] * w[2]);
}

LoopSearchHint SlamPipeline::GetLoopSearchHint() const
{
    /* lidar_graph_slam.cpp:273-381 */
    const std::vector<LocalMapGPU>& maps = mBuilder->LocalMaps();
    std::size_t unfinished = 0;
    while (unfinished < maps.size() && maps[unfinished].finished)
        ++unfinished;
    LoopSearchHint hint;
    hint.accum_travel_dist = mBuilder->AccumTravelDist();
    hint.last_finished_scan_id = kInvalidId;
    hint.last_finished_map_id = kInvalidId;
    if (unfinished == 0)
        return hint;
    const int map_id_max = unfinished < maps.size() ? maps[unfinished].id : kInvalidId;
    const int node_id_max = unfinished < maps.size() ? maps[unfinished].scan_node_id_min : kInvalidId;
    for (const ScanNode& n : mPoseGraph.scan_nodes) {
        if (map_id_max != kInvalidId && (n.local_map_id >= map_id_max || n.node_id >= node_id_max))
            break;
        hint.scan_nodes.push_back(ScanNodeData { n.node_id, n.global_pose });
    }
    for (const LocalMapNode& n : mPoseGraph.local_map_nodes) {
        if (map_id_max != kInvalidId && n.local_map_id >= map_id_max)
            break;
        const LocalMapGPU& m = maps[n.local_map_id];
        hint.local_map_nodes.push_back(LocalMapData { m.id, m.scan_node_id_min, m.scan_node_id_max, m.finished });
    }
    const LocalMapData& last = hint.local_map_nodes.back();
    hint.last_finished_map_id = last.local_map_id;
    hint.last_finished_scan_id = (last.scan_node_id_min + last.scan_node_id_max) / 2;
    return hint;
}

std::vector<LoopDetectionQuery> SlamPipeline::GetLoopDetectionQueries(const std::vector<LoopCandidate>& candidates)
{
    /* lidar_graph_slam.cpp:384-415: references into the pose graph and the local maps, here resolved */
    std::vector<LoopDetectionQuery> queries;
    queries.reserve(candidates.size());
    for (const LoopCandidate& c : candidates) {
        const ScanNode& query_node = mPoseGraph.scan_nodes.at(c.query_scan_node_id);
        const ScanNode& ref_node = mPoseGraph.scan_nodes.at(c.reference_scan_node_id);
        const LocalMapGPU& ref_map = mBuilder->LocalMaps().at(c.reference_local_map_id);
        const LocalMapNode& ref_map_node = mPoseGraph.local_map_nodes.at(c.reference_local_map_id);
        LoopDetectionQuery q;
        q.scan = query_node.scan;
        q.scan_id = query_node.node_id;
        q.scan_node_id = query_node.node_id;
        q.scan_global_pose = query_node.global_pose;
        if (mSettings.host_final_matchers) {
            /* a finished local map never changes: one host copy for the CPU final matcher */
            if (mLocalMapCopies.size() <= static_cast<std::size_t>(ref_map.id))
                mLocalMapCopies.resize(ref_map.id + 1);
            if (!mLocalMapCopies[ref_map.id]) {
                mLocalMapCopies[ref_map.id].reset(new HostCopy);
                WithHostCopy(*ref_map.map, *mLocalMapCopies[ref_map.id]);
            }
            q.local_map = ref_map.map->View();
            q.local_map.values = mLocalMapCopies[ref_map.id]->cells.data();
            q.local_map.block_allocated = mLocalMapCopies[ref_map.id]->allocation.data();
        } else {
            q.local_map = ref_map.map->View();
        }
        q.local_map_global_pose = ref_map_node.global_pose;
        q.reference_scan_local_pose = ref_node.local_pose;
        queries.push_back(q);
    }
    return queries;
}

void SlamPipeline::AppendLoopClosingEdges(const std::vector<LoopDetectionResult>& results)
{
    /* lidar_graph_slam.cpp:448-504 */
    for (const LoopDetectionResult& r : results) {
        PoseGraphEdge e;
        e.local_map_id = static_cast<int>(r.local_map_id);
        e.scan_node_id = r.scan_node_id;
        e.edge_type = EdgeType::InterLocalMap;
        e.constraint_type = ConstraintType::Loop;
        e.relative_pose = NormalizeAngle(r.relative_pose);
        e.information = Inverse(r.estimated_covariance);
        mPoseGraph.edges.push_back(e);
    }
}

void SlamPipeline::GetPoseGraphForOptimization(std::vector<int>& local_map_ids, std::vector<std::array<double, 3>>& local_map_poses,
                                               std::vector<int>& scan_node_ids, std::vector<std::array<double, 3>>& scan_poses,
                                               std::vector<EdgePose>& edges) const
{
    /* lidar_graph_slam.cpp:106-194: the finished part of the graph */
    const std::vector<LocalMapGPU>& maps = mBuilder->LocalMaps();
    std::size_t unfinished = 0;
    while (unfinished < maps.size() && maps[unfinished].finished)
        ++unfinished;
    const int map_id_max = unfinished < maps.size() ? maps[unfinished].id : kInvalidId;
    const int node_id_max = unfinished < maps.size() ? maps[unfinished].scan_node_id_min : kInvalidId;
    for (const LocalMapNode& n : mPoseGraph.local_map_nodes)
        if (map_id_max == kInvalidId || n.local_map_id < map_id_max) {
            local_map_ids.push_back(n.local_map_id);
            local_map_poses.push_back({ n.global_pose.x, n.global_pose.y, n.global_pose.theta });
        }
    for (const ScanNode& n : mPoseGraph.scan_nodes)
        if (map_id_max == kInvalidId || (n.local_map_id < map_id_max && n.node_id < node_id_max)) {
            scan_node_ids.push_back(n.node_id);
            scan_poses.push_back({ n.global_pose.x, n.global_pose.y, n.global_pose.theta });
        }
    for (const PoseGraphEdge& e : mPoseGraph.edges)
        if (map_id_max == kInvalidId || (e.local_map_id < map_id_max && e.scan_node_id < node_id_max)) {
            EdgePose p;
            p.is_loop_closing = e.IsLoopClosingConstraint();
            p.local_map_index = e.local_map_id;          /* ids are positions */
            p.scan_node_index = e.scan_node_id;
            p.relative_pose = { e.relative_pose.x, e.relative_pose.y, e.relative_pose.theta };
            p.information = e.information;
            edges.push_back(p);
        }
}

void SlamPipeline::AfterLoopClosure(const std::vector<int>& local_map_ids, const std::vector<std::array<double, 3>>& local_map_poses,
                                    const std::vector<int>& scan_node_ids, const std::vector<std::array<double, 3>>& scan_poses)
{
    /* lidar_graph_slam.cpp:506-672 */
    for (std::size_t i = 0; i < local_map_ids.size(); ++i)
        mPoseGraph.local_map_nodes.at(local_map_ids[i]).global_pose =
            Pose2D { local_map_poses[i][0], local_map_poses[i][1], local_map_poses[i][2] };
    for (std::size_t i = 0; i < scan_node_ids.size(); ++i)
        mPoseGraph.scan_nodes.at(scan_node_ids[i]).global_pose = Pose2D { scan_poses[i][0], scan_poses[i][1], scan_poses[i][2] };
    const int last_map_id = *std::max_element(local_map_ids.begin(), local_map_ids.end());
    const LocalMapGPU& last_map = mBuilder->LocalMaps().at(last_map_id);
    /* the first odometry edge that took no part in this optimisation (:566-571) */
    auto it = std::find_if(mPoseGraph.edges.begin(), mPoseGraph.edges.end(), [&](const PoseGraphEdge& e) {
        return e.local_map_id == last_map.id && e.scan_node_id > last_map.scan_node_id_max; });
    if (it != mPoseGraph.edges.end()) {
        /* the nodes added since follow along their odometry edges (:590-640) */
        int processed_map = last_map.id, processed_node = last_map.scan_node_id_max;
        for (; it != mPoseGraph.edges.end(); ++it) {
            const PoseGraphEdge& e = *it;
            if (!e.IsOdometryConstraint())
                continue;
            const bool update_scan = e.local_map_id == processed_map && e.scan_node_id > processed_node;
            const bool update_map = e.local_map_id > processed_map && e.scan_node_id == processed_node;
            if (update_scan)
                mPoseGraph.scan_nodes.at(e.scan_node_id).global_pose =
                    Compound(mPoseGraph.local_map_nodes.at(e.local_map_id).global_pose, e.relative_pose);
            else if (update_map)
                mPoseGraph.local_map_nodes.at(e.local_map_id).global_pose =
                    MoveBackward(mPoseGraph.scan_nodes.at(e.scan_node_id).global_pose, e.relative_pose);
            else {
                std::fprintf(stderr, "csm_host: an odometry edge after the loop closure updates neither node\n");
                std::abort();
            }
            processed_map = e.local_map_id;
            processed_node = e.scan_node_id;
        }
    }
    mBuilder->AfterLoopClosure(mPoseGraph);
}

void SlamPipeline::RunBackendStep()
{
    /* lidar_graph_slam_backend.cpp:92-198 */
    ++mCounters.backend_steps;
    const double t_step = Seconds();
    double t_mark = t_step;
    auto lap = [&](const char* id) {
        const double now = Seconds();
        Observe(id, (now - t_mark) * 1e6);
        t_mark = now;
    };
    auto end_at = [&](const char* id) {
        Observe("Backend.ProcessTime", (Seconds() - t_step) * 1e6);
        Observe(id, static_cast<double>(mCounters.backend_steps - 1));
    };
    const LoopSearchHint hint = GetLoopSearchHint();
    lap("Backend.LoopSearchSetupTime");
    if (hint.local_map_nodes.empty() || hint.scan_nodes.empty())
        return end_at("Backend.EndAtLoopSearchSetup");
    const std::vector<LoopCandidate> candidates = mLoopSearcher->Search(hint);
    lap("Backend.LoopSearchTime");
    if (candidates.empty())
        return end_at("Backend.EndAtLoopSearch");
    ++mCounters.backend_steps_with_candidates;
    const std::vector<LoopDetectionQuery> queries = GetLoopDetectionQueries(candidates);
    lap("Backend.LoopDetectionSetupTime");
    const double t0 = Seconds();
    const std::vector<LoopDetectionResult> results = mLoopDetector->Detect(queries);
    mCounters.t_detect += Seconds() - t0;
    mCounters.loop_queries += static_cast<int>(queries.size());
    lap("Backend.LoopDetectionTime");
    if (results.empty())
        return end_at("Backend.EndAtLoopDetection");
    mCounters.loops_detected += static_cast<int>(results.size());
    mLoops.insert(mLoops.end(), results.begin(), results.end());
    AppendLoopClosingEdges(results);
    lap("Backend.PoseGraphAppendTime");
    std::vector<int> map_ids, node_ids;
    std::vector<std::array<double, 3>> map_poses, node_poses;
    std::vector<EdgePose> edges;
    GetPoseGraphForOptimization(map_ids, map_poses, node_ids, node_poses, edges);
    lap("Backend.OptimizationSetupTime");
    mOptimizer->Optimize(map_poses, node_poses, edges);
    ++mCounters.optimizations;
    lap("Backend.OptimizationTime");
    AfterLoopClosure(map_ids, map_poses, node_ids, node_poses);
    lap("Backend.PoseGraphUpdateTime");
    Observe("Backend.ProcessStepTime", (Seconds() - t_step) * 1e6);
    end_at("Backend.EndAtLoopClosure");
}

int SlamPipeline::RunLog(const std::vector<CarmenRecord>& records, bool finish)
{
    int used = 0;
    for (const CarmenRecord& rec : records)
        if (rec.kind == CarmenRecord::Kind::Scan && rec.scan)
            used += ProcessScan(rec.scan, rec.odom_pose, rec.time_stamp) ? 1 : 0;
    if (finish)
        Finish();
    return used;
}

void SlamPipeline::SetMetricSink(const MetricSinkPtr& sink)
{
    mMetricSink = sink;
    mScanMatcher->SetMetricSink(sink);
    mLoopDetector->SetMetricSink(sink);
}

} /* namespace csm_host */
