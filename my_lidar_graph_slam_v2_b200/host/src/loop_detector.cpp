#include "csm_host/loop_detector.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>

namespace csm_host {

FinalMatcher MakeLinearSolverFinalMatcher(const std::shared_ptr<ScanMatcherLinearSolver>& solver)
{
    return [solver](const GridMapView& map, const ScanDataPtr& scan, const Pose2D&, const Pose2D& initial_pose) {
        return solver->OptimizePose(ScanMatchingQuery { map, scan, initial_pose });
    };
}

namespace {

/* loop_detector_branch_bound.cpp:54-55 (the same two Asserts in all three detectors) */
void CheckThresholds(double score_threshold, double known_rate_threshold)
{
    if (!(score_threshold > 0.0 && score_threshold <= 1.0) ||
        !(known_rate_threshold > 0.0 && known_rate_threshold <= 1.0)) {
        std::fprintf(stderr, "csm_host: loop detector thresholds must be in (0, 1]\n");
        std::abort();
    }
}

/* The reference's per-query loop (loop_detector_correlative.cpp:68-146,
 * loop_detector_grid_search.cpp:64-129) around a coarse matcher with thresholds */
template <typename Matcher>
std::vector<LoopDetectionResult> DetectPerQuery(
    const std::vector<LoopDetectionQuery>& queries, Matcher& matcher, const FinalMatcher& final_matcher,
    double score_threshold, double known_rate_threshold,
    const std::function<void(const char*, double)>& observe)
{
    std::vector<LoopDetectionResult> results;
    MicroTimer timer;
    for (std::size_t i = 0; i < queries.size(); ++i) {
        const LoopDetectionQuery& q = queries[i];
        if (q.local_map.map_id < 0) {
            std::fprintf(stderr, "csm_host: loop detection maps need a LocalMapId\n");
            std::abort();
        }
        timer.Start();
        /* pose of the scan node in the map-local frame */
        const Pose2D init = InverseCompound(q.local_map_global_pose, q.scan_global_pose);
        const ScanMatchingSummary coarse = matcher.OptimizePose(q.local_map, q.scan, init,
                                                                score_threshold, known_rate_threshold);
        if (!coarse.pose_found)
            continue;
        LoopDetectionResult out;
        out.relative_pose = coarse.estimated_pose;
        out.estimated_covariance = coarse.estimated_covariance;
        if (final_matcher) {
            const ScanMatchingSummary fin = final_matcher(q.local_map, q.scan, q.reference_scan_local_pose,
                                                          coarse.estimated_pose);
            out.relative_pose = fin.estimated_pose;
            out.estimated_covariance = fin.estimated_covariance;
        }
        out.local_map_pose = q.local_map_global_pose;
        out.local_map_id = q.local_map.map_id;
        out.scan_node_id = q.scan_node_id;
        out.normalized_score = coarse.normalized_score;
        out.query_index = static_cast<int>(i);
        results.push_back(out);
        observe("LoopDetectionTime", timer.ElapsedMicro());
    }
    observe("NumOfQueries", static_cast<double>(queries.size()));
    observe("NumOfDetections", static_cast<double>(results.size()));
    return results;
}

} /* namespace */

LoopDetectorBranchBound::LoopDetectorBranchBound(
    const std::string& name, const std::shared_ptr<ScanMatcherBranchBound>& scan_matcher,
    const FinalMatcher& final_matcher, double score_threshold, double known_rate_threshold) :
    LoopDetector(name), mScanMatcher(scan_matcher), mFinalMatcher(final_matcher),
    mScoreThreshold(score_threshold), mKnownRateThreshold(known_rate_threshold)
{
    CheckThresholds(score_threshold, known_rate_threshold);
}

LoopDetectorCorrelative::LoopDetectorCorrelative(
    const std::string& name, const std::shared_ptr<ScanMatcherCorrelative>& scan_matcher,
    const FinalMatcher& final_matcher, double score_threshold, double known_rate_threshold) :
    LoopDetector(name), mScanMatcher(scan_matcher), mFinalMatcher(final_matcher),
    mScoreThreshold(score_threshold), mKnownRateThreshold(known_rate_threshold)
{
    CheckThresholds(score_threshold, known_rate_threshold);
}

std::vector<LoopDetectionResult> LoopDetectorCorrelative::Detect(const std::vector<LoopDetectionQuery>& queries)
{
    /* the matcher keeps map and coarse map resident by LocalMapId (the reference's mPrecompMaps,
     * loop_detector_correlative.cpp:83-90) */
    return DetectPerQuery(queries, *mScanMatcher, mFinalMatcher, mScoreThreshold, mKnownRateThreshold,
                          [this](const char* m, double v) { Observe(m, v); });
}

LoopDetectorGridSearch::LoopDetectorGridSearch(
    const std::string& name, const std::shared_ptr<ScanMatcherGridSearch>& scan_matcher,
    const FinalMatcher& final_matcher, double score_threshold, double known_rate_threshold) :
    LoopDetector(name), mScanMatcher(scan_matcher), mFinalMatcher(final_matcher),
    mScoreThreshold(score_threshold), mKnownRateThreshold(known_rate_threshold)
{
    CheckThresholds(score_threshold, known_rate_threshold);
}

std::vector<LoopDetectionResult> LoopDetectorGridSearch::Detect(const std::vector<LoopDetectionQuery>& queries)
{
    return DetectPerQuery(queries, *mScanMatcher, mFinalMatcher, mScoreThreshold, mKnownRateThreshold,
                          [this](const char* m, double v) { Observe(m, v); });
}

namespace {

/* Upload the first-touch maps of one chunk. Block-sparse views whose block
 * buffers follow one another in memory (a pinned staging area filled map by
 * map) go in one batched call = one PCIe copy; anything else map by map. */
void UploadNewMaps(const DeviceContextPtr& ctx, const std::vector<const GridMapView*>& maps)
{
    if (maps.empty())
        return;
    csm_handle h = ctx->Handle();
    const GridMapView& m0 = *maps[0];
    bool batch = m0.blocks != nullptr && maps.size() > 1;
    std::size_t nblk = 0;
    for (std::size_t i = 0; i < maps.size() && batch; ++i) {
        const GridMapView& m = *maps[i];
        batch = m.blocks != nullptr && m.rows == m0.rows && m.cols == m0.cols &&
                m.log2_block_size == m0.log2_block_size && m.resolution == m0.resolution &&
                m.blocks == m0.blocks + (nblk << (2 * m0.log2_block_size)) &&
                m.block_index == m0.block_index + nblk;
        nblk += static_cast<std::size_t>(m.n_blocks);
    }
    if (batch) {
        std::vector<std::int64_t> ids(maps.size());
        std::vector<std::int32_t> counts(maps.size());
        std::vector<double> ox(maps.size()), oy(maps.size());
        for (std::size_t i = 0; i < maps.size(); ++i) {
            ids[i] = maps[i]->map_id; counts[i] = maps[i]->n_blocks;
            ox[i] = maps[i]->offset_x; oy[i] = maps[i]->offset_y;
        }
        ctx->Check(csm_upload_grids_blocks(h, static_cast<int>(maps.size()), ids.data(), m0.blocks,
                                           m0.block_index, counts.data(), m0.log2_block_size,
                                           m0.rows >> m0.log2_block_size, m0.cols >> m0.log2_block_size,
                                           m0.resolution, ox.data(), oy.data()), "csm_upload_grids_blocks");
        return;
    }
    bool dense_batch = m0.blocks == nullptr && maps.size() > 1;
    for (std::size_t i = 0; i < maps.size() && dense_batch; ++i)
        dense_batch = maps[i]->blocks == nullptr && maps[i]->rows == m0.rows && maps[i]->cols == m0.cols &&
                      maps[i]->resolution == m0.resolution;
    if (dense_batch) {
        /* one call: a single copy when the grids follow one another in host memory */
        std::vector<std::int64_t> ids(maps.size());
        std::vector<const std::uint16_t*> ptrs(maps.size());
        std::vector<double> ox(maps.size()), oy(maps.size());
        for (std::size_t i = 0; i < maps.size(); ++i) {
            ids[i] = maps[i]->map_id; ptrs[i] = maps[i]->values;
            ox[i] = maps[i]->offset_x; oy[i] = maps[i]->offset_y;
        }
        ctx->Check(csm_upload_grids(h, static_cast<int>(maps.size()), ids.data(), ptrs.data(), m0.rows, m0.cols,
                                    m0.resolution, ox.data(), oy.data()), "csm_upload_grids");
        return;
    }
    for (const GridMapView* m : maps) {
        if (m->blocks != nullptr)
            ctx->Check(csm_upload_grid_blocks(h, m->map_id, m->blocks, m->block_index, m->n_blocks,
                                              m->log2_block_size, m->rows >> m->log2_block_size,
                                              m->cols >> m->log2_block_size, m->resolution,
                                              m->offset_x, m->offset_y), "csm_upload_grid_blocks");
        else
            ctx->Check(csm_upload_grid(h, m->map_id, m->values, m->rows, m->cols, m->resolution,
                                       m->offset_x, m->offset_y), "csm_upload_grid");
    }
}

} /* namespace */

void LoopDetectorBranchBound::UseDeviceRefiner(int num_of_iterations_max, double convergence_threshold,
                                               double initial_lambda, double covariance_scale)
{
    mDeviceRefiner = true;
    mFinalMatcher = FinalMatcher();
    mRefineParams.max_iterations = num_of_iterations_max;
    mRefineParams.reserved = 0;
    mRefineParams.convergence_threshold = convergence_threshold;
    mRefineParams.lambda = initial_lambda;
    mRefineParams.covariance_scale = covariance_scale;
}

void LoopDetectorBranchBound::SetPipelineLanes(const std::vector<DeviceContextPtr>& extra_lanes)
{
    mExtraLanes = extra_lanes;
    /* all lanes upload on the first lane's copy stream: strictly in call order, so that a lane's
     * maps land (and its search starts) while the later groups are still crossing PCIe */
    for (const DeviceContextPtr& c : mExtraLanes)
        c->Check(csm_share_copy_stream(c->Handle(), mScanMatcher->Context()->Handle()), "csm_share_copy_stream");
    mLaneMaps.assign(1 + extra_lanes.size(), std::set<std::int64_t>());
    mLaneScans.assign(1 + extra_lanes.size(), std::set<std::int64_t>());
    mCachedMaps.clear();
    mCachedScans.clear();
}

std::vector<LoopDetectionResult> LoopDetectorBranchBound::Detect(
    const std::vector<LoopDetectionQuery>& queries)
{
    std::vector<LoopDetectionResult> results;
    mLastResults.clear();
    mLastRefined.clear();
    mBestWord = 0;
    if (queries.empty())
        return results;
    MicroTimer timer;
    const DeviceContextPtr& ctx = mScanMatcher->Context();
    const int hmax = mScanMatcher->NodeHeightMax();
    const int nq = static_cast<int>(queries.size());

    /* per query: initial pose InverseCompound(map, scan) (:97-98), sensor pose,
     * steps and windows with the reference's expressions */
    std::vector<csm_loop_query> dq(nq);
    for (int i = 0; i < nq; ++i)
        if (queries[i].local_map.map_id < 0) {
            std::fprintf(stderr, "csm_host: loop detection maps need a LocalMapId\n");
            std::abort();
        }
    /* (evaluated after the uploads have been enqueued: the copies start before the host does this) */
    auto fill_queries = [&]() {
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
    };
    mLastResults.resize(nq);
    if (mDeviceRefiner)
        mLastRefined.resize(nq);
    /* read back the oldest batch in flight on a context */
    auto finish = [&](const DeviceContextPtr& c, int f0, int fc) {
        if (mDeviceRefiner)
            c->Check(csm_loop_batch_finish_refined(c->Handle(), mLastResults.data() + f0,
                                                   mLastRefined.data() + f0, fc), "csm_loop_batch_finish_refined");
        else
            c->Check(csm_loop_batch_finish(c->Handle(), mLastResults.data() + f0, fc), "csm_loop_batch_finish");
    };
    auto set_refiner = [&](const DeviceContextPtr& c) {
        c->Check(csm_set_refiner(c->Handle(), mDeviceRefiner ? &mRefineParams : nullptr), "csm_set_refiner");
    };

    if (!mExtraLanes.empty()) {
        /* ---- pipelined over lanes: one search batch per upload group, on the lane of its maps ---- */
        const int lanes = NumOfLanes();
        const int chunk = std::max(1, mChunkSize);
        const int ugroup = std::max(1, std::min(mUploadChunk, chunk));
        auto lane_ctx = [&](int l) -> const DeviceContextPtr& { return l == 0 ? ctx : mExtraLanes[l - 1]; };
        /* a segment = one search batch of up to mChunkSize consecutive queries whose maps belong to
         * one lane; its first-touch maps go up in groups of mUploadChunk */
        struct Segment { int first, count, lane; };
        std::vector<Segment> segments;
        for (int i = 0; i < nq; ++i) {
            const std::int64_t id = queries[i].local_map.map_id;
            const int lane = static_cast<int>((id / chunk) % lanes);
            if (segments.empty() || segments.back().lane != lane || segments.back().count >= chunk)
                segments.push_back(Segment { i, 0, lane });
            ++segments.back().count;
        }
        std::vector<std::vector<int>> in_flight(lanes);       /* segment indices, oldest first */
        auto finish_oldest = [&](int lane) {
            const Segment& sg = segments[in_flight[lane].front()];
            finish(lane_ctx(lane), sg.first, sg.count);
            in_flight[lane].erase(in_flight[lane].begin());
        };
        /* pass 1: every upload is enqueued first, so that PCIe never waits for the host to prepare a
         * search batch; pass 2: pyramids and the search batch of every segment behind its own uploads */
        std::vector<std::vector<std::vector<std::int64_t>>> fresh_ids(segments.size());   /* per upload group */
        for (std::size_t si = 0; si < segments.size(); ++si) {
            const Segment& sg = segments[si];
            const DeviceContextPtr& c = lane_ctx(sg.lane);
            for (int first = sg.first; first < sg.first + sg.count; first += ugroup) {
                const int last = std::min(sg.first + sg.count, first + ugroup);
                std::vector<const GridMapView*> fresh;
                fresh_ids[si].emplace_back();
                for (int i = first; i < last; ++i) {
                    const GridMapView& m = queries[i].local_map;
                    if (mLaneMaps[sg.lane].insert(m.map_id).second) {
                        fresh.push_back(&m);
                        fresh_ids[si].back().push_back(m.map_id);
                    }
                }
                UploadNewMaps(c, fresh);
            }
        }
        fill_queries();
        const bool trace = std::getenv("CSM_HOST_TRACE") != nullptr;
        if (trace) std::fprintf(stderr, "lanes: uploads enqueued at %.0f us\n", timer.ElapsedMicro());
        for (std::size_t si = 0; si < segments.size(); ++si) {
            const Segment& sg = segments[si];
            const DeviceContextPtr& c = lane_ctx(sg.lane);
            for (int i = sg.first; i < sg.first + sg.count; ++i) {
                const LoopDetectionQuery& q = queries[i];
                if (mLaneScans[sg.lane].insert(q.scan_id).second)
                    c->Check(csm_upload_scan(c->Handle(), q.scan_id, q.scan->angles.data(), q.scan->ranges.data(),
                                             static_cast<int>(q.scan->NumOfScans())), "csm_upload_scan");
            }
            for (const std::vector<std::int64_t>& group : fresh_ids[si])
                if (!group.empty())
                    c->Check(csm_build_pyramids(c->Handle(), static_cast<int>(group.size()), group.data(), hmax),
                             "csm_build_pyramids");
            set_refiner(c);
            if (in_flight[sg.lane].size() >= 4)         /* the library keeps at most 4 batches in flight */
                finish_oldest(sg.lane);
            c->Check(csm_loop_batch_enqueue(c->Handle(), dq.data() + sg.first, sg.count, hmax,
                                            mQueryIndexBase + sg.first), "csm_loop_batch_enqueue");
            in_flight[sg.lane].push_back(static_cast<int>(si));
            if (trace) std::fprintf(stderr, "lanes: segment %zu enqueued at %.0f us\n", si, timer.ElapsedMicro());
        }
        for (int lane = 0; lane < lanes; ++lane)
            while (!in_flight[lane].empty()) {
                finish_oldest(lane);
                if (trace) std::fprintf(stderr, "lanes: lane %d finished a batch at %.0f us\n", lane, timer.ElapsedMicro());
            }
    } else {
    csm_handle h = ctx->Handle();
    const int chunk = std::max(1, mChunkSize);
    const int nchunks = (nq + chunk - 1) / chunk;

    /* first touch of a local map: upload + pyramid, cached by LocalMapId
     * (loop_detector_branch_bound.cpp:83-89). All uploads are enqueued first, in
     * groups of mUploadChunk maps: they stream over PCIe on the copy stream while
     * the groups that have landed are expanded and precomputed, and the search
     * batches (mChunkSize queries) whose maps are complete run behind them. */
    const int ugroup = std::max(1, std::min(mUploadChunk, chunk));
    std::vector<std::vector<std::int64_t>> new_maps;      /* per upload group */
    std::vector<int> group_end;                           /* query index one past each group */
    for (int first = 0; first < nq; first += ugroup) {
        const int last = std::min(nq, first + ugroup);
        std::vector<const GridMapView*> fresh;
        new_maps.emplace_back();
        for (int i = first; i < last; ++i) {
            const GridMapView& m = queries[i].local_map;
            if (mCachedMaps.insert(m.map_id).second) {
                fresh.push_back(&m);
                new_maps.back().push_back(m.map_id);
            }
        }
        group_end.push_back(last);
        UploadNewMaps(ctx, fresh);
    }
    for (const LoopDetectionQuery& q : queries)
        if (mCachedScans.insert(q.scan_id).second)
            ctx->Check(csm_upload_scan(h, q.scan_id, q.scan->angles.data(), q.scan->ranges.data(),
                                       static_cast<int>(q.scan->NumOfScans())), "csm_upload_scan");
    std::size_t next_group = 0;
    fill_queries();
    set_refiner(ctx);
    int finished = 0;      /* chunks whose results have been read back */
    for (int c = 0; c < nchunks; ++c) {
        const int first = c * chunk, count = std::min(nq, first + chunk) - first;
        /* pyramids of every upload group this batch touches (in upload order) */
        for (; next_group < new_maps.size() &&
               (next_group == 0 || group_end[next_group - 1] < first + count); ++next_group)
            if (!new_maps[next_group].empty())
                ctx->Check(csm_build_pyramids(h, static_cast<int>(new_maps[next_group].size()),
                                              new_maps[next_group].data(), hmax), "csm_build_pyramids");
        if (c - finished >= 4) {        /* the library keeps at most 4 batches in flight */
            const int f0 = finished * chunk, fc = std::min(nq, f0 + chunk) - f0;
            finish(ctx, f0, fc);
            ++finished;
        }
        ctx->Check(csm_loop_batch_enqueue(h, dq.data() + first, count, hmax, mQueryIndexBase + first),
                   "csm_loop_batch_enqueue");
    }
    for (; finished < nchunks; ++finished) {
        const int f0 = finished * chunk, fc = std::min(nq, f0 + chunk) - f0;
        finish(ctx, f0, fc);
    }
    }

    /* packed best word over the whole call (what the device keeps per handle, here over all lanes) */
    for (int i = 0; i < nq; ++i) {
        const csm_result& r = mLastResults[i];
        if (!r.found)
            continue;
        const std::uint64_t key = static_cast<std::uint64_t>(998ll * r.sum_value + 64536ll * r.n_known);
        const std::uint64_t word = (key << 20) | static_cast<std::uint64_t>(0xFFFFF - (mQueryIndexBase + i));
        mBestWord = std::max(mBestWord, word);
    }

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
        if (mDeviceRefiner && mLastRefined[i].valid) {
            /* the device ran the final matcher on the coarse sensor pose; what the reference
             * returns is its pose moved back to the robot frame and its covariance (:132-135) */
            const csm_refined& f = mLastRefined[i];
            out.relative_pose = MoveBackward(Pose2D { f.pose[0], f.pose[1], f.pose[2] },
                                             q.scan->relative_sensor_pose);
            std::copy(f.covariance, f.covariance + 9, out.estimated_covariance.begin());
            mRefineParams.lambda = f.lambda;      /* damping state for the next Detect */
        } else if (mFinalMatcher) {
            /* sub-pixel refinement around the reference scan's local pose (:110-127); its
             * pose and covariance are what the reference returns (:132-135) */
            const ScanMatchingSummary fin = mFinalMatcher(q.local_map, q.scan, q.reference_scan_local_pose,
                                                          out.relative_pose);
            out.relative_pose = fin.estimated_pose;
            out.estimated_covariance = fin.estimated_covariance;
        } else if (mCoarseCovariance) {
            out.estimated_covariance = mScanMatcher->Cost()->ComputeCovariance(q.local_map, *q.scan, best);
        }
        out.local_map_pose = q.local_map_global_pose;
        out.local_map_id = q.local_map.map_id;
        out.scan_node_id = q.scan_node_id;
        out.normalized_score = r.normalized_score;
        out.query_index = i;
        results.push_back(out);
    }
    if (mMetricSink) {
        /* one batch instead of one observation per query: the whole call, divided evenly */
        const double micro = timer.ElapsedMicro();
        for (std::size_t i = 0; i < results.size(); ++i)
            Observe("LoopDetectionTime", micro / static_cast<double>(nq));
        Observe("NumOfQueries", nq);
        Observe("NumOfDetections", static_cast<double>(results.size()));
        Observe("PrecompMapMemoryUsage", static_cast<double>(mCachedMaps.size()) *
                static_cast<double>(hmax + 1) * queries[0].local_map.rows * queries[0].local_map.cols * 2.0);
    }
    return results;
}

} /* namespace csm_host */
