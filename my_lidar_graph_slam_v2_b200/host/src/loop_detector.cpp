#include "csm_host/loop_detector.hpp"

#include <cstdio>
#include <cstdlib>
#include <map>

namespace csm_host {

LoopDetectorBranchBound::LoopDetectorBranchBound(
    const std::string& name, const std::shared_ptr<ScanMatcherBranchBound>& scan_matcher,
    const FinalMatcher& final_matcher, double score_threshold, double known_rate_threshold) :
    LoopDetector(name), mScanMatcher(scan_matcher), mFinalMatcher(final_matcher),
    mScoreThreshold(score_threshold), mKnownRateThreshold(known_rate_threshold)
{
    /* loop_detector_branch_bound.cpp:54-55 */
    if (!(score_threshold > 0.0 && score_threshold <= 1.0) ||
        !(known_rate_threshold > 0.0 && known_rate_threshold <= 1.0)) {
        std::fprintf(stderr, "csm_host: loop detector thresholds must be in (0, 1]\n");
        std::abort();
    }
}

std::vector<LoopDetectionResult> LoopDetectorBranchBound::Detect(
    const std::vector<LoopDetectionQuery>& queries)
{
    std::vector<LoopDetectionResult> results;
    mLastResults.clear();
    if (queries.empty())
        return results;
    const DeviceContextPtr& ctx = mScanMatcher->Context();
    csm_handle h = ctx->Handle();
    const int hmax = mScanMatcher->NodeHeightMax();
    const int nq = static_cast<int>(queries.size());

    /* first touch of a local map: upload + pyramid, cached by LocalMapId
     * (loop_detector_branch_bound.cpp:83-89) */
    std::vector<std::int64_t> new_maps;
    for (const LoopDetectionQuery& q : queries) {
        const GridMapView& m = q.local_map;
        if (m.map_id < 0) {
            std::fprintf(stderr, "csm_host: loop detection maps need a LocalMapId\n");
            std::abort();
        }
        if (mCachedMaps.insert(m.map_id).second) {
            ctx->Check(csm_upload_grid(h, m.map_id, m.values, m.rows, m.cols, m.resolution,
                                       m.offset_x, m.offset_y), "csm_upload_grid");
            new_maps.push_back(m.map_id);
        }
        if (mCachedScans.insert(q.scan_id).second)
            ctx->Check(csm_upload_scan(h, q.scan_id, q.scan->angles.data(), q.scan->ranges.data(),
                                       static_cast<int>(q.scan->NumOfScans())), "csm_upload_scan");
    }
    if (!new_maps.empty())
        ctx->Check(csm_build_pyramids(h, static_cast<int>(new_maps.size()), new_maps.data(), hmax),
                   "csm_build_pyramids");

    /* per query: initial pose InverseCompound(map, scan) (:97-98), sensor pose,
     * steps and windows with the reference's expressions */
    std::vector<csm_loop_query> dq(nq);
    std::map<std::pair<std::int64_t, double>, std::array<double, 3>> steps;
    for (int i = 0; i < nq; ++i) {
        const LoopDetectionQuery& q = queries[i];
        const Pose2D init = InverseCompound(q.local_map_global_pose, q.scan_global_pose);
        const Pose2D sensor = Compound(init, q.scan->relative_sensor_pose);
        auto key = std::make_pair(q.scan_id, q.local_map.resolution);
        auto it = steps.find(key);
        if (it == steps.end()) {
            std::array<double, 3> st;
            ComputeSearchStep(q.local_map.resolution, *q.scan, st[0], st[1], st[2]);
            it = steps.emplace(key, st).first;
        }
        const std::array<double, 3>& st = it->second;
        csm_loop_query& d = dq[i];
        d.map_id = q.local_map.map_id;
        d.scan_id = q.scan_id;
        d.sensor_pose[0] = sensor.x; d.sensor_pose[1] = sensor.y; d.sensor_pose[2] = sensor.theta;
        d.win_x = static_cast<int>(std::ceil(0.5 * mScanMatcher->RangeX() / st[0]));
        d.win_y = static_cast<int>(std::ceil(0.5 * mScanMatcher->RangeY() / st[1]));
        d.win_t = static_cast<int>(std::ceil(0.5 * mScanMatcher->RangeTheta() / st[2]));
        d.reserved = 0;
        d.step_x = st[0]; d.step_y = st[1]; d.step_t = st[2];
        d.score_thr = mScoreThreshold;
        d.known_thr = mKnownRateThreshold;
    }

    mLastResults.resize(nq);
    ctx->Check(csm_loop_batch(h, dq.data(), nq, hmax, mQueryIndexBase, mLastResults.data()),
               "csm_loop_batch");

    for (int i = 0; i < nq; ++i) {
        const csm_result& r = mLastResults[i];
        if (!r.found)
            continue;                      /* :106-108 */
        const LoopDetectionQuery& q = queries[i];
        const csm_loop_query& d = dq[i];
        const Pose2D best { d.sensor_pose[0] + d.step_x * r.best_x, d.sensor_pose[1] + d.step_y * r.best_y,
                            d.sensor_pose[2] + d.step_t * r.best_t };
        LoopDetectionResult out;
        out.relative_pose = MoveBackward(best, q.scan->relative_sensor_pose);
        out.estimated_covariance = mScanMatcher->Cost()->ComputeCovariance(q.local_map, *q.scan, best);
        if (mFinalMatcher) {
            /* sub-pixel refinement around the reference scan's local pose (:110-127) */
            const ScanMatchingSummary fin = mFinalMatcher(q.local_map, q.scan, q.reference_scan_local_pose,
                                                          out.relative_pose);
            out.relative_pose = fin.estimated_pose;
            out.estimated_covariance = fin.estimated_covariance;
        }
        out.local_map_pose = q.local_map_global_pose;
        out.local_map_id = q.local_map.map_id;
        out.scan_node_id = q.scan_node_id;
        out.normalized_score = r.normalized_score;
        out.query_index = i;
        results.push_back(out);
    }
    return results;
}

} /* namespace csm_host */
