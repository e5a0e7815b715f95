#include "csm_host/loop_detector.hpp"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>

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
 * loop_detector_grid_search.cpp:64-129) around a coarse matcher with thresholds. A single match is a
 * handful of dependent launches that leave most of the device idle, so the coarse stage may run on several
 * matchers at once (each with its own device context, i.e. stream): query i goes to matcher
 * LocalMapId mod L, which keeps that map and its coarse map resident, one host thread per matcher. The
 * final matcher then runs over the found poses in query order on the calling thread, like the reference
 * (its damping state carries from one query to the next). */
template <typename Matcher>
std::vector<LoopDetectionResult> DetectPerQuery(
    const std::vector<LoopDetectionQuery>& queries, const std::vector<Matcher*>& matchers,
    const FinalMatcher& final_matcher, double score_threshold, double known_rate_threshold,
    const std::function<void(const char*, double)>& observe)
{
    MicroTimer timer;
    const std::size_t nq = queries.size();
    const std::size_t lanes = matchers.size();
    for (const LoopDetectionQuery& q : queries)
        if (q.local_map.map_id < 0) {
            std::fprintf(stderr, "csm_host: loop detection maps need a LocalMapId\n");
            std::abort();
        }
    std::vector<ScanMatchingSummary> coarse(nq);
    auto run_lane = [&](std::size_t lane) {
        for (std::size_t i = 0; i < nq; ++i) {
            const LoopDetectionQuery& q = queries[i];
            if (static_cast<std::size_t>(q.local_map.map_id) % lanes != lane)
                continue;
            /* pose of the scan node in the map-local frame */
            const Pose2D init = InverseCompound(q.local_map_global_pose, q.scan_global_pose);
            coarse[i] = matchers[lane]->OptimizePose(q.local_map, q.scan, init, score_threshold, known_rate_threshold);
        }
    };
    if (lanes == 1)
        run_lane(0);
    else {
        std::vector<std::thread> threads;
        for (std::size_t lane = 1; lane < lanes; ++lane)
            threads.emplace_back(run_lane, lane);
        run_lane(0);
        for (std::thread& t : threads) t.join();
    }
    const double coarse_micro = timer.ElapsedMicro();
    std::vector<LoopDetectionResult> results;
    for (std::size_t i = 0; i < nq; ++i) {
        if (!coarse[i].pose_found)
            continue;
        const LoopDetectionQuery& q = queries[i];
        timer.Start();
        LoopDetectionResult out;
        out.relative_pose = coarse[i].estimated_pose;
        out.estimated_covariance = coarse[i].estimated_covariance;
        if (final_matcher) {
            const ScanMatchingSummary fin = final_matcher(q.local_map, q.scan, q.reference_scan_local_pose,
                                                          coarse[i].estimated_pose);
            out.relative_pose = fin.estimated_pose;
            out.estimated_covariance = fin.estimated_covariance;
        }
        out.local_map_pose = q.local_map_global_pose;
        out.local_map_id = q.local_map.map_id;
        out.scan_node_id = q.scan_node_id;
        out.normalized_score = coarse[i].normalized_score;
        out.query_index = static_cast<int>(i);
        results.push_back(out);
        /* the coarse stage ran concurrently: its time is divided evenly over the queries */
        observe("LoopDetectionTime", timer.ElapsedMicro() + coarse_micro / static_cast<double>(nq));
    }
    observe("NumOfQueries", static_cast<double>(nq));
    observe("NumOfDetections", static_cast<double>(results.size()));
    return results;
}

template <typename Matcher>
std::vector<Matcher*> MatcherList(const std::shared_ptr<Matcher>& first, const std::vector<std::shared_ptr<Matcher>>& extra)
{
    std::vector<Matcher*> all { first.get() };
    for (const auto& m : extra) all.push_back(m.get());
    return all;
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
    return DetectPerQuery(queries, MatcherList(mScanMatcher, mExtraMatchers), mFinalMatcher, mScoreThreshold,
                          mKnownRateThreshold, [this](const char* m, double v) { Observe(m, v); });
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
    return DetectPerQuery(queries, MatcherList(mScanMatcher, mExtraMatchers), mFinalMatcher, mScoreThreshold,
                          mKnownRateThreshold, [this](const char* m, double v) { Observe(m, v); });
}

/* Gathers blocks that live in separate heap allocations (the reference's GridMap storage,
 * grid_map.cpp:522-535) into page-locked staging, with a small pool of threads: one staging area per
 * upload group of a Detect (a group's H2D copy reads its area while the next group is gathered),
 * reused by the following Detect calls (every call ends with all its copies complete). */
class BlockGatherer
{
public:
    explicit BlockGatherer(int n_threads)
    {
        for (int i = 1; i < n_threads; ++i)
            mWorkers.emplace_back([this] { Work(); });
    }
    ~BlockGatherer()
    {
        {
            std::lock_guard<std::mutex> lock(mMutex);
            mStop = true;
            ++mGeneration;
        }
        mWake.notify_all();
        for (std::thread& t : mWorkers) t.join();
        for (void* p : mAreas) csm_free_pinned(p);
    }
    void* Area(std::size_t k, std::size_t bytes)
    {
        if (mAreas.size() <= k) { mAreas.resize(k + 1, nullptr); mAreaBytes.resize(k + 1, 0); }
        if (mAreaBytes[k] < bytes) {
            if (mAreas[k]) csm_free_pinned(mAreas[k]);
            mAreas[k] = csm_alloc_pinned(bytes + bytes / 4);
            mAreaBytes[k] = mAreas[k] ? bytes + bytes / 4 : 0;
            if (!mAreas[k]) { std::fprintf(stderr, "csm_host: out of page-locked memory\n"); std::abort(); }
        }
        return mAreas[k];
    }
    /* dst[b] <- src[b] (bytes each) for b in [0, src.size()): Start wakes the workers and returns, Finish lets
     * the caller take chunks too and waits for the rest. Between the two the caller is free: the detector
     * enqueues the previous group's upload and a segment's search while the next group is being gathered. */
    void Start(std::vector<const std::uint16_t*>&& src, std::uint16_t* dst, std::size_t bytes)
    {
        mOwned = std::move(src);
        mSrc = mOwned.data(); mDst = reinterpret_cast<char*>(dst); mBytes = bytes; mCount = mOwned.size();
        mNext.store(0);
        mBusy.store(static_cast<int>(mWorkers.size()));
        {
            std::lock_guard<std::mutex> lock(mMutex);
            ++mGeneration;
        }
        mWake.notify_all();
        mStarted = true;
    }
    void Finish()
    {
        if (!mStarted)
            return;
        Chunks();
        while (mBusy.load(std::memory_order_acquire) != 0)
            std::this_thread::yield();
        mStarted = false;
    }
    bool Started() const { return mStarted; }

private:
    static void CopyBlock(char* dst, const char* src, std::size_t bytes)
    {
#if defined(__SSE2__)
        if ((reinterpret_cast<std::uintptr_t>(dst) & 15u) == 0 && (bytes & 63u) == 0) {
            for (std::size_t o = 0; o < bytes; o += 64) {
                const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o));
                const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o + 16));
                const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o + 32));
                const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + o + 48));
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + o), a);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + o + 16), b);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + o + 32), c);
                _mm_stream_si128(reinterpret_cast<__m128i*>(dst + o + 48), d);
            }
            return;
        }
#endif
        std::memcpy(dst, src, bytes);
    }
    void Chunks()
    {
        constexpr std::size_t kChunk = 64, kAhead = 6;
        for (;;) {
            const std::size_t b0 = mNext.fetch_add(kChunk);
            if (b0 >= mCount) break;
            const std::size_t b1 = std::min(mCount, b0 + kChunk);
            for (std::size_t b = b0; b < b1; ++b) {
                /* the blocks are scattered heap allocations: ask for the lines of a later block while this
                 * one is copied, and write the staging area past the cache (it is read once, by the DMA) */
                if (b + kAhead < mCount) {
                    const char* next = reinterpret_cast<const char*>(mSrc[b + kAhead]);
                    for (std::size_t o = 0; o < mBytes; o += 64)
                        __builtin_prefetch(next + o, 0, 0);
                }
                CopyBlock(mDst + b * mBytes, reinterpret_cast<const char*>(mSrc[b]), mBytes);
            }
        }
#if defined(__SSE2__)
        _mm_sfence();           /* the streamed stores are visible before the group is handed to the copy engine */
#endif
    }
    void Work()
    {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mMutex);
                mWake.wait(lock, [&] { return mGeneration != seen; });
                seen = mGeneration;
                if (mStop) return;
            }
            Chunks();
            mBusy.fetch_sub(1, std::memory_order_release);
        }
    }
    std::vector<std::thread> mWorkers;
    std::mutex mMutex;
    std::condition_variable mWake;
    unsigned long long mGeneration = 0;
    bool mStop = false;
    std::atomic<std::size_t> mNext { 0 };
    std::atomic<int> mBusy { 0 };
    std::vector<const std::uint16_t*> mOwned;
    bool mStarted = false;
    const std::uint16_t* const* mSrc = nullptr;
    char* mDst = nullptr;
    std::size_t mBytes = 0, mCount = 0;
    std::vector<void*> mAreas;
    std::vector<std::size_t> mAreaBytes;
};

namespace {

/* Upload the first-touch maps of one upload group. Block-sparse views whose block
 * buffers follow one another in memory (a pinned staging area filled map by
 * map) go in one batched call = one PCIe copy; views whose blocks are separate heap allocations
 * are gathered into staging area `area` first; anything else map by map. */
/* An upload group whose maps keep their blocks in separate heap allocations: where the gathered blocks
 * go and what the upload call needs */
struct HeapGroup
{
    bool valid = false;
    char* stage = nullptr;
    std::int32_t* index = nullptr;
    std::vector<std::int64_t> ids;
    std::vector<std::int32_t> counts;
    std::vector<double> ox, oy;
    int log2bs = 0, block_rows = 0, block_cols = 0;
    double resolution = 0.0;
};

/* Lays the group out in page-locked staging area `area` and starts the gather; false = not such a group */
bool HeapGroupStart(const std::vector<const GridMapView*>& maps, BlockGatherer* gatherer, std::size_t area, HeapGroup& g)
{
    g.valid = false;
    if (maps.empty() || gatherer == nullptr)
        return false;
    const GridMapView& m0 = *maps[0];
    bool heap = m0.block_ptrs != nullptr;
    for (std::size_t i = 0; i < maps.size() && heap; ++i) {
        const GridMapView& m = *maps[i];
        heap = m.block_ptrs != nullptr && m.rows == m0.rows && m.cols == m0.cols &&
               m.log2_block_size == m0.log2_block_size && m.resolution == m0.resolution;
    }
    if (!heap)
        return false;
    std::size_t total = 0;
    for (const GridMapView* m : maps) total += static_cast<std::size_t>(m->n_blocks);
    const std::size_t block_bytes = sizeof(std::uint16_t) << (2 * m0.log2_block_size);
    const std::size_t index_off = (total * block_bytes + 255) & ~static_cast<std::size_t>(255);
    g.stage = static_cast<char*>(gatherer->Area(area, index_off + total * sizeof(std::int32_t) + 256));
    std::vector<const std::uint16_t*> src(total);
    g.index = reinterpret_cast<std::int32_t*>(g.stage + index_off);
    g.ids.resize(maps.size()); g.counts.resize(maps.size()); g.ox.resize(maps.size()); g.oy.resize(maps.size());
    std::size_t b = 0;
    for (std::size_t i = 0; i < maps.size(); ++i) {
        const GridMapView& m = *maps[i];
        std::copy(m.block_ptrs, m.block_ptrs + m.n_blocks, src.begin() + b);
        std::copy(m.block_index, m.block_index + m.n_blocks, g.index + b);
        b += static_cast<std::size_t>(m.n_blocks);
        g.ids[i] = m.map_id; g.counts[i] = m.n_blocks; g.ox[i] = m.offset_x; g.oy[i] = m.offset_y;
    }
    g.log2bs = m0.log2_block_size; g.block_rows = m0.rows >> m0.log2_block_size; g.block_cols = m0.cols >> m0.log2_block_size;
    g.resolution = m0.resolution;
    gatherer->Start(std::move(src), reinterpret_cast<std::uint16_t*>(g.stage), block_bytes);
    g.valid = true;
    return true;
}

/* ... enqueues the copy of a gathered group (after BlockGatherer::Finish) */
void HeapGroupUpload(const DeviceContextPtr& ctx, const HeapGroup& g)
{
    ctx->Check(csm_upload_grids_blocks(ctx->Handle(), static_cast<int>(g.ids.size()), g.ids.data(),
                                       reinterpret_cast<const std::uint16_t*>(g.stage), g.index, g.counts.data(),
                                       g.log2bs, g.block_rows, g.block_cols, g.resolution, g.ox.data(), g.oy.data()),
               "csm_upload_grids_blocks");
}

void UploadNewMaps(const DeviceContextPtr& ctx, const std::vector<const GridMapView*>& maps,
                   BlockGatherer* gatherer, std::size_t area)
{
    if (maps.empty())
        return;
    csm_handle h = ctx->Handle();
    const GridMapView& m0 = *maps[0];
    {
        HeapGroup g;
        if (HeapGroupStart(maps, gatherer, area, g)) {
            gatherer->Finish();
            HeapGroupUpload(ctx, g);
            return;
        }
    }
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
    bool dense_batch = m0.values != nullptr && m0.blocks == nullptr && maps.size() > 1;
    for (std::size_t i = 0; i < maps.size() && dense_batch; ++i)
        dense_batch = maps[i]->values != nullptr && maps[i]->blocks == nullptr && maps[i]->rows == m0.rows &&
                      maps[i]->cols == m0.cols && maps[i]->resolution == m0.resolution;
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
        if (m->block_ptrs != nullptr) {
            /* a lone heap-allocated map: flatten its blocks into a temporary */
            const std::size_t cells = static_cast<std::size_t>(1) << (2 * m->log2_block_size);
            std::vector<std::uint16_t> tmp(cells * static_cast<std::size_t>(m->n_blocks));
            for (int b = 0; b < m->n_blocks; ++b)
                std::memcpy(tmp.data() + cells * b, m->block_ptrs[b], cells * sizeof(std::uint16_t));
            ctx->Check(csm_upload_grid_blocks(h, m->map_id, tmp.data(), m->block_index, m->n_blocks,
                                              m->log2_block_size, m->rows >> m->log2_block_size,
                                              m->cols >> m->log2_block_size, m->resolution,
                                              m->offset_x, m->offset_y), "csm_upload_grid_blocks");
            ctx->Check(csm_synchronize(h), "csm_synchronize");      /* tmp goes away */
        } else if (m->blocks != nullptr)
            ctx->Check(csm_upload_grid_blocks(h, m->map_id, m->blocks, m->block_index, m->n_blocks,
                                              m->log2_block_size, m->rows >> m->log2_block_size,
                                              m->cols >> m->log2_block_size, m->resolution,
                                              m->offset_x, m->offset_y), "csm_upload_grid_blocks");
        else
            ctx->Check(csm_upload_grid(h, m->map_id, m->values, m->rows, m->cols, m->resolution,
                                       m->offset_x, m->offset_y), "csm_upload_grid");
    }
}

/* scans travel with the Detect call that uses them: call-local ids, released when the call ends */
constexpr std::int64_t kCallScanIdBase = static_cast<std::int64_t>(1) << 60;

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
    mMapLane.clear();
}

void LoopDetectorBranchBound::SetGatherThreads(int n)
{
    mGatherThreads = std::max(1, n);
    mGatherer.reset();
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
    const int lanes = NumOfLanes();
    auto lane_ctx = [&](int l) -> const DeviceContextPtr& { return l == 0 ? ctx : mExtraLanes[l - 1]; };

    std::vector<csm_loop_query> dq(nq);
    bool any_heap = false;
    for (int i = 0; i < nq; ++i) {
        if (queries[i].local_map.map_id < 0) {
            std::fprintf(stderr, "csm_host: loop detection maps need a LocalMapId\n");
            std::abort();
        }
        any_heap = any_heap || queries[i].local_map.block_ptrs != nullptr;
    }
    if (any_heap && !mGatherer) {
        const int hw = static_cast<int>(std::thread::hardware_concurrency());
        mGatherer = std::make_shared<BlockGatherer>(mGatherThreads > 0 ? mGatherThreads : std::max(1, std::min(16, hw)));
    }
    /* the scans of this call: every distinct ScanData object gets a call-local id (the reference keeps
     * no per-scan state; LoopDetectionQuery::scan_id is not trusted to be unique across calls) */
    std::vector<const ScanData*> scans;
    std::vector<int> scan_of(nq);
    for (int i = 0; i < nq; ++i) {
        const ScanData* s = queries[i].scan.get();
        auto it = std::find(scans.begin(), scans.end(), s);
        scan_of[i] = static_cast<int>(it - scans.begin());
        if (it == scans.end()) scans.push_back(s);
    }
    /* per query: initial pose InverseCompound(map, scan) (:97-98), sensor pose, steps and windows with
     * the reference's expressions (evaluated after the uploads have been enqueued: the copies start
     * before the host does this) */
    auto fill_queries = [&]() {
        std::map<std::pair<int, double>, std::array<double, 3>> steps;
        for (int i = 0; i < nq; ++i) {
            const LoopDetectionQuery& q = queries[i];
            const Pose2D init = InverseCompound(q.local_map_global_pose, q.scan_global_pose);
            const Pose2D sensor = Compound(init, q.scan->relative_sensor_pose);
            auto key = std::make_pair(scan_of[i], q.local_map.resolution);
            auto it = steps.find(key);
            if (it == steps.end()) {
                std::array<double, 3> st;
                ComputeSearchStep(q.local_map.resolution, *q.scan, st[0], st[1], st[2]);
                it = steps.emplace(key, st).first;
            }
            const std::array<double, 3>& st = it->second;
            csm_loop_query& d = dq[i];
            d.map_id = q.local_map.map_id;
            d.scan_id = kCallScanIdBase + scan_of[i];
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
    /* read back the oldest batch in flight on a context; CSM_E_CAPACITY comes back to the caller */
    auto finish = [&](const DeviceContextPtr& c, int f0, int fc) -> int {
        const int rc = mDeviceRefiner
            ? csm_loop_batch_finish_refined(c->Handle(), mLastResults.data() + f0, mLastRefined.data() + f0, fc)
            : csm_loop_batch_finish(c->Handle(), mLastResults.data() + f0, fc);
        if (rc != CSM_OK && rc != CSM_E_CAPACITY)
            c->Check(rc, "csm_loop_batch_finish");
        return rc;
    };

    /* A segment = one search batch: up to mChunkSize consecutive queries on one lane. A map that is
     * resident stays on its lane; first-touch maps take the lanes in turn, batch by batch, in the
     * order they arrive (whatever their ids). Its first-touch maps go up in groups of mUploadChunk. */
    const int chunk = std::max(1, mChunkSize);
    const int ugroup = std::max(1, std::min(mUploadChunk, chunk));
    for (int i = 0; i < nq; ++i)
        if (queries[i].local_map.device_resident && lanes > 1) {
            std::fprintf(stderr, "csm_host: device-resident local maps live on the detector's own context; "
                                 "pipeline lanes cannot see them\n");
            std::abort();
        }
    struct Segment { int first, count, lane; bool fresh; };
    std::vector<Segment> segments;
    for (int i = 0; i < nq; ++i) {
        const auto res = mMapLane.find(queries[i].local_map.map_id);
        const bool fresh = res == mMapLane.end();
        /* the tail batch starts where mTailChunk queries are left (first-touch maps only: that is where it pays) */
        const bool tail_start = mTailChunk > 0 && fresh && nq - i == mTailChunk && nq > mTailChunk;
        if (!tail_start && !segments.empty() && segments.back().count < chunk &&
            ((fresh && segments.back().fresh) || (!fresh && segments.back().lane == res->second))) {
            ++segments.back().count;
        } else {
            const int lane = fresh ? (mArrivals++ % lanes) : res->second;
            segments.push_back(Segment { i, 1, lane, fresh });
        }
        if (fresh)
            mMapLane[queries[i].local_map.map_id] = segments.back().lane;
    }
    /* Segment by segment: the first-touch maps of a segment are gathered and sent in upload groups, then its
     * levels and its search batch are enqueued behind them -- the device works on a segment while the host
     * gathers the blocks of the next one (the gather, not PCIe, is the longest leg of a cold Detect). */
    std::vector<std::vector<std::vector<std::int64_t>>> fresh_ids(segments.size());   /* per upload group */
    /* the upload groups of the whole call, in order; the gather of group k + 1 is started as soon as group k
     * has been gathered, so that it runs while this thread enqueues copies, levels and search batches */
    struct UploadJob { std::size_t si; std::vector<const GridMapView*> fresh; HeapGroup heap; bool started = false; };
    std::vector<UploadJob> ujobs;
    for (std::size_t si = 0; si < segments.size(); ++si) {
        const Segment& sg = segments[si];
        if (!sg.fresh)
            continue;
        std::set<std::int64_t> seen;
        /* the very first group of a call may be smaller than the others (SetFirstGroupDivisor): nothing crosses
         * PCIe before it is gathered */
        int step = ujobs.empty() ? std::max(std::min(8, ugroup), ugroup / mFirstGroupDivisor) : ugroup;
        for (int first = sg.first; first < sg.first + sg.count; first += step, step = ugroup) {
            const int last = std::min(sg.first + sg.count, first + step);
            UploadJob job;
            job.si = si;
            fresh_ids[si].emplace_back();
            for (int i = first; i < last; ++i) {
                const GridMapView& m = queries[i].local_map;
                if (seen.insert(m.map_id).second) {
                    /* a map built on the device (GridMapBuilderGPU) has nothing to upload: only its levels are due */
                    if (!m.device_resident)
                        job.fresh.push_back(&m);
                    fresh_ids[si].back().push_back(m.map_id);
                }
            }
            ujobs.push_back(std::move(job));
        }
    }
    std::size_t next_job = 0;
    const bool trace = std::getenv("CSM_HOST_TRACE") != nullptr;
    auto start_job = [&](std::size_t k) {
        if (mGatherer && k < ujobs.size() && !ujobs[k].started && !mGatherer->Started())
            ujobs[k].started = HeapGroupStart(ujobs[k].fresh, mGatherer.get(), k, ujobs[k].heap);
    };
    auto upload_segment = [&](std::size_t si) {
        while (next_job < ujobs.size() && ujobs[next_job].si == si) {
            UploadJob& job = ujobs[next_job];
            const DeviceContextPtr& c = lane_ctx(segments[si].lane);
            start_job(next_job);
            if (job.started) {
                mGatherer->Finish();
                if (trace) std::fprintf(stderr, "lanes: group %zu gathered at %.0f us\n", next_job, timer.ElapsedMicro());
                start_job(next_job + 1);              /* the workers go on with the next group */
                HeapGroupUpload(c, job.heap);
                if (trace) std::fprintf(stderr, "lanes: group %zu copy enqueued at %.0f us\n", next_job, timer.ElapsedMicro());
            } else
                UploadNewMaps(c, job.fresh, mGatherer.get(), next_job);
            ++next_job;
        }
    };
    start_job(0);                 /* the workers gather the first group while this thread fills the descriptors */
    fill_queries();
    if (trace) std::fprintf(stderr, "lanes: descriptors filled at %.0f us\n", timer.ElapsedMicro());
    std::vector<std::vector<int>> in_flight(lanes);       /* segment indices, oldest first */
    std::vector<int> overflowed;                          /* segments to search again in smaller batches */
    auto finish_oldest = [&](int lane) {
        const int si = in_flight[lane].front();
        if (finish(lane_ctx(lane), segments[si].first, segments[si].count) == CSM_E_CAPACITY)
            overflowed.push_back(si);
        in_flight[lane].erase(in_flight[lane].begin());
    };
    std::vector<std::set<int>> lane_scans(lanes);
    for (std::size_t si = 0; si < segments.size(); ++si) {
        const Segment& sg = segments[si];
        const DeviceContextPtr& c = lane_ctx(sg.lane);
        upload_segment(si);
        if (trace) std::fprintf(stderr, "lanes: segment %zu uploads enqueued at %.0f us\n", si, timer.ElapsedMicro());
        for (int i = sg.first; i < sg.first + sg.count; ++i)
            if (lane_scans[sg.lane].insert(scan_of[i]).second) {
                const ScanData& s = *scans[scan_of[i]];
                c->Check(csm_upload_scan(c->Handle(), kCallScanIdBase + scan_of[i], s.angles.data(), s.ranges.data(),
                                         static_cast<int>(s.NumOfScans())), "csm_upload_scan");
            }
        if (trace) std::fprintf(stderr, "lanes: segment %zu scans up at %.0f us\n", si, timer.ElapsedMicro());
        for (const std::vector<std::int64_t>& group : fresh_ids[si])
            if (!group.empty())
                c->Check(csm_build_pyramids(c->Handle(), static_cast<int>(group.size()), group.data(), hmax),
                         "csm_build_pyramids");
        if (trace) std::fprintf(stderr, "lanes: segment %zu levels enqueued at %.0f us\n", si, timer.ElapsedMicro());
        c->Check(csm_set_refiner(c->Handle(), mDeviceRefiner ? &mRefineParams : nullptr), "csm_set_refiner");
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
    /* a batch whose frontier lists overflowed (a scan that matches many maps well at the coarse
     * levels): halve it until it fits; a single query that still overflows dives for an incumbent
     * first (fewer nodes), and only then is it an error */
    std::function<void(const DeviceContextPtr&, int, int)> rerun = [&](const DeviceContextPtr& c, int first, int count) {
        ++mCapacityRetries;
        if (count == 1)
            c->Check(csm_set_option(c->Handle(), "bb_dive", 1), "csm_set_option");
        const int half = count == 1 ? 1 : count / 2;
        for (int f = first; f < first + count; f += half) {
            const int n = std::min(half, first + count - f);
            c->Check(csm_loop_batch_enqueue(c->Handle(), dq.data() + f, n, hmax, mQueryIndexBase + f),
                     "csm_loop_batch_enqueue");
            const int rc = finish(c, f, n);
            if (rc == CSM_E_CAPACITY) {
                if (count == 1) c->Check(rc, "csm_loop_batch_finish (a single query overflows the frontier lists)");
                rerun(c, f, n);
            }
        }
        if (count == 1)
            c->Check(csm_set_option(c->Handle(), "bb_dive", 2), "csm_set_option");
    };
    for (int si : overflowed)
        rerun(lane_ctx(segments[si].lane), segments[si].first, segments[si].count);
    for (int lane = 0; lane < lanes; ++lane)
        for (int s : lane_scans[lane])
            lane_ctx(lane)->Check(csm_release_scan(lane_ctx(lane)->Handle(), kCallScanIdBase + s), "csm_release_scan");

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
        Observe("PrecompMapMemoryUsage", static_cast<double>(mMapLane.size()) *
                static_cast<double>(hmax + 1) * queries[0].local_map.rows * queries[0].local_map.cols * 2.0);
    }
    return results;
}

/* ---- several GPUs, one process ------------------------------------------------------------- */
LoopDetectorBranchBoundMultiGPU::LoopDetectorBranchBoundMultiGPU(
    const std::string& name, const std::vector<std::shared_ptr<LoopDetectorBranchBound>>& shards,
    const std::vector<DeviceContextPtr>& contexts) :
    LoopDetector(name), mShards(shards), mContexts(contexts)
{
    if (shards.empty() || shards.size() != contexts.size()) {
        std::fprintf(stderr, "csm_host: the multi-GPU detector needs one shard and one context per GPU\n");
        std::abort();
    }
}

void LoopDetectorBranchBoundMultiGPU::UseNcclExchange()
{
    std::vector<csm_handle> hs;
    for (const DeviceContextPtr& c : mContexts) hs.push_back(c->Handle());
    mContexts[0]->Check(csm_comm_init_all(hs.data(), static_cast<int>(hs.size())), "csm_comm_init_all");
    mNccl = true;
}

std::vector<LoopDetectionResult> LoopDetectorBranchBoundMultiGPU::Detect(const std::vector<LoopDetectionQuery>& queries)
{
    const int G = NumOfGpus();
    const int nq = static_cast<int>(queries.size());
    /* shard g takes the queries whose local map it owns: LocalMapId mod G */
    std::vector<std::vector<LoopDetectionQuery>> part(G);
    std::vector<std::vector<int>> origin(G);
    for (int i = 0; i < nq; ++i) {
        const std::int64_t id = queries[i].local_map.map_id;
        const int g = static_cast<int>(((id % G) + G) % G);
        part[g].push_back(queries[i]);
        origin[g].push_back(i);
    }
    mLastShardSizes.assign(G, 0);
    std::vector<std::vector<LoopDetectionResult>> found(G);
    std::vector<std::thread> threads;
    for (int g = 0; g < G; ++g) {
        mLastShardSizes[g] = static_cast<int>(part[g].size());
        threads.emplace_back([&, g] { found[g] = mShards[g]->Detect(part[g]); });
    }
    for (std::thread& t : threads) t.join();
    /* results in query order, like the reference's concatenation of its cores' vectors */
    mLastResults.assign(nq, csm_result {});
    std::vector<LoopDetectionResult> results;
    std::vector<std::uint64_t> words(G, 0);
    for (int g = 0; g < G; ++g) {
        const std::vector<csm_result>& lr = mShards[g]->LastResults();
        for (std::size_t k = 0; k < lr.size(); ++k) {
            const int i = origin[g][k];
            mLastResults[i] = lr[k];
            if (lr[k].found) {
                const std::uint64_t key = static_cast<std::uint64_t>(998ll * lr[k].sum_value + 64536ll * lr[k].n_known);
                words[g] = std::max(words[g], (key << 20) | static_cast<std::uint64_t>(0xFFFFF - i));
            }
        }
        for (LoopDetectionResult r : found[g]) {
            r.query_index = origin[g][r.query_index];
            results.push_back(r);
        }
    }
    std::sort(results.begin(), results.end(),
              [](const LoopDetectionResult& a, const LoopDetectionResult& b) { return a.query_index < b.query_index; });
    mBestWord = *std::max_element(words.begin(), words.end());
    if (mNccl) {
        /* the same maximum through one 8-byte all-reduce over NVLink: every GPU ends with the best word */
        std::vector<csm_handle> hs;
        for (const DeviceContextPtr& c : mContexts) hs.push_back(c->Handle());
        std::vector<int> tickets(G, -1);
        mContexts[0]->Check(csm_comm_allreduce_words_all(hs.data(), G, words.data(), tickets.data()),
                            "csm_comm_allreduce_words_all");
        for (int g = 0; g < G; ++g) {
            std::uint64_t w = 0;
            mContexts[g]->Check(csm_comm_best_result(hs[g], tickets[g], &w), "csm_comm_best_result");
            if (w != mBestWord) {
                std::fprintf(stderr, "csm_host: the all-reduced best word differs from the host maximum\n");
                std::abort();
            }
        }
    }
    Observe("NumOfQueries", nq);
    Observe("NumOfDetections", static_cast<double>(results.size()));
    return results;
}

} /* namespace csm_host */
