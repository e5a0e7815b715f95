/* loop_detector.hpp -- GPU loop detectors behind the reference's LoopDetector
 * plugin interface (mapping/loop_detector.hpp:97-116: Detect(queries) ->
 * results), JSON type strings "BranchBound" | "RealTimeCorrelative" |
 * "GridSearch" (loop_detector_factory.cpp:202-209). Constructor parameters
 * follow the reference classes (loop_detector_branch_bound.cpp:41-56,
 * loop_detector_correlative.cpp:40-56, loop_detector_grid_search.cpp:33-49);
 * the final sub-pixel matcher is an optional callback (CPU) or the device refiner.
 * The branch-and-bound detector matches all queries in device batches; the other
 * two run the reference's per-query loop on the GPU matchers, on several device
 * contexts at once when given (SetConcurrentMatchers). */
#pragma once

#include <cstdint>
#include <functional>
#include <set>
#include <unordered_map>

#include "csm_host/scan_matchers.hpp"

namespace csm_host {

using FinalMatcher = std::function<ScanMatchingSummary(const GridMapView& map, const ScanDataPtr& scan,
                                                       const Pose2D& center, const Pose2D& initial_pose)>;

/* The reference's default refinement stage: a ScanMatcherLinearSolver run on the coarse pose
 * (loop_detector_branch_bound.cpp:110-127; the map centre argument is not used by this matcher) */
FinalMatcher MakeLinearSolverFinalMatcher(const std::shared_ptr<ScanMatcherLinearSolver>& solver);

class LoopDetector
{
public:
    explicit LoopDetector(const std::string& name) : mName(name) { }
    virtual ~LoopDetector() = default;
    const std::string& Name() const { return mName; }
    virtual std::vector<LoopDetectionResult> Detect(const std::vector<LoopDetectionQuery>& queries) = 0;
    /* "<Name>.InputSetupTime | LoopDetectionTime | NumOfQueries | NumOfDetections"
     * (loop_detector_branch_bound.cpp:24-38) go here; null = off */
    void SetMetricSink(const MetricSinkPtr& sink) { mMetricSink = sink; }

protected:
    void Observe(const char* metric, double value) const
    {
        if (mMetricSink) mMetricSink->Observe(mName + "." + metric, value);
    }

    std::string mName;
    MetricSinkPtr mMetricSink;
};

class LoopDetectorBranchBound final : public LoopDetector
{
public:
    LoopDetectorBranchBound(const std::string& name,
                            const std::shared_ptr<ScanMatcherBranchBound>& scan_matcher,
                            const FinalMatcher& final_matcher,
                            double score_threshold, double known_rate_threshold);

    /* loop_detector_branch_bound.cpp:59-156: per query, build the pyramid of a
     * local map the first time it is seen (cached by LocalMapId, never
     * evicted), match with thresholds, refine and emit a result per success.
     * All queries are matched in one device batch. */
    std::vector<LoopDetectionResult> Detect(const std::vector<LoopDetectionQuery>& queries) override;

    /* Queries are matched in chunks of this many (default 128): the uploads of
     * later chunks overlap the search of earlier ones */
    void SetChunkSize(int n) { mChunkSize = n; }
    /* First-touch maps are uploaded in groups of this many (default 64, at most the chunk size) */
    void SetUploadChunk(int n) { mUploadChunk = n; }
    /* The last search batch of a call holds at most this many queries (0 = like the others): what is left to do
     * once the last map has been gathered -- its copy, its levels, its search -- is what a cold Detect waits for
     * at the end, so a short last batch ends sooner; the batches before it overlap the gather anyway. */
    void SetTailChunk(int n) { mTailChunk = n; }
    /* the first upload group of a Detect holds 1 / n of the maps of the others (default 1 = all groups alike: measured, no gain from a smaller first group: the copy engine is the longest leg either way) */
    void SetFirstGroupDivisor(int n) { mFirstGroupDivisor = n < 1 ? 1 : n; }
    /* Without a final matcher the result carries the covariance of the cost
     * function at the coarse pose (computed on the CPU); switch it off when
     * the caller refines the poses itself */
    void SetCoarseCovariance(bool on) { mCoarseCovariance = on; }
    /* Run the reference's default final matcher (ScanMatcherLinearSolver with CostSquareError,
     * "FinalScanMatcherType": "LinearSolver") on the device, batched behind the search, instead of a
     * CPU final matcher: every detected loop comes back with its refined pose and covariance
     * (csm_set_refiner). `initial_lambda` is the solver's damping state; it is carried from one
     * Detect to the next like the reference's member (scan_matcher_linear_solver.cpp:100-104),
     * per batch instead of per query. Replaces any CPU final matcher. */
    void UseDeviceRefiner(int num_of_iterations_max, double convergence_threshold, double initial_lambda,
                          double covariance_scale);
    /* Per-query refinement outcomes of the last Detect (valid == 0 where none) */
    const std::vector<csm_refined>& LastRefined() const { return mLastRefined; }
    /* Pipeline lanes: additional device contexts (handles, i.e. streams and workspaces) on the SAME
     * device. With lanes, every search batch (mChunkSize consecutive queries) runs on the lane its
     * maps belong to (lane = (LocalMapId / chunk size) mod lanes, stable across calls like the map ->
     * GPU rule of the sharded detector), behind the uploads and pyramids of its own maps only: the
     * search of a batch that has landed runs while the following maps are still crossing PCIe, and
     * the latency-bound launches of neighbouring batches overlap on the device. Results are the same as without lanes. Clears the caches. */
    void SetPipelineLanes(const std::vector<DeviceContextPtr>& extra_lanes);
    int NumOfLanes() const { return 1 + static_cast<int>(mExtraLanes.size()); }
    /* Packed best word of the last Detect, max over all lanes: (key << 20 | (0xFFFFF - global query
     * index)) of the best found query, 0 if none -- what csm_best_key_device holds for one handle */
    std::uint64_t BestWord() const { return mBestWord; }
    /* Forget which maps are resident (the next Detect uploads them again) */
    void ClearCache() { mMapLane.clear(); }
    /* Threads that gather heap-allocated blocks (GridMapView::block_ptrs) into page-locked staging
     * (default: min(16, hardware threads)) */
    void SetGatherThreads(int n);
    /* Batches that overflowed the device's frontier lists and were searched again in smaller
     * batches (CSM_E_CAPACITY is recoverable), over the detector's life */
    int NumOfCapacityRetries() const { return mCapacityRetries; }
    /* Sharded use: global index of queries[0] (packed best word) */
    void SetQueryIndexBase(int base) { mQueryIndexBase = base; }
    /* Per-query device results of the last Detect, in query order */
    const std::vector<csm_result>& LastResults() const { return mLastResults; }

private:
    std::shared_ptr<ScanMatcherBranchBound> mScanMatcher;
    FinalMatcher mFinalMatcher;
    double mScoreThreshold, mKnownRateThreshold;
    std::unordered_map<std::int64_t, int> mMapLane;     /* resident maps -> the lane that holds them */
    std::vector<csm_result> mLastResults;
    std::vector<csm_refined> mLastRefined;
    std::vector<DeviceContextPtr> mExtraLanes;
    std::shared_ptr<class BlockGatherer> mGatherer;
    int mGatherThreads = 0;
    int mArrivals = 0;          /* first-touch batches seen so far: they take the lanes in turn */
    int mCapacityRetries = 0;
    std::uint64_t mBestWord = 0;
    bool mDeviceRefiner = false;
    csm_refine_params mRefineParams {};
    int mQueryIndexBase = 0;
    int mChunkSize = 128;
    int mUploadChunk = 64;
    int mTailChunk = 0;
    int mFirstGroupDivisor = 1;
    bool mCoarseCovariance = true;
};

/* The same detector over several GPUs of one box, one process: the model is the reference's
 * LoopDetectorFPGAParallel (loop_detector_fpga_parallel.cpp:32-68), which splits the queries over its
 * two accelerator cores and concatenates their results. Queries go to GPU LocalMapId mod G (a map is
 * uploaded and precomputed once, on the GPU that owns it, whatever batch it comes back in); every shard
 * is a LoopDetectorBranchBound on its own device context(s), driven by its own host thread; results
 * come back in query order. The packed best word of the call is exchanged with one 8-byte NCCL
 * all-reduce(max) over the shards' handles when UseNcclExchange() was called (otherwise it is the
 * maximum the host takes over the shards' words; the two agree). */
class LoopDetectorBranchBoundMultiGPU final : public LoopDetector
{
public:
    LoopDetectorBranchBoundMultiGPU(const std::string& name,
                                    const std::vector<std::shared_ptr<LoopDetectorBranchBound>>& shards,
                                    const std::vector<DeviceContextPtr>& contexts);
    std::vector<LoopDetectionResult> Detect(const std::vector<LoopDetectionQuery>& queries) override;
    int NumOfGpus() const { return static_cast<int>(mShards.size()); }
    /* csm_comm_init_all over the shards' first contexts */
    void UseNcclExchange();
    std::uint64_t BestWord() const { return mBestWord; }
    const std::vector<csm_result>& LastResults() const { return mLastResults; }
    const std::vector<int>& LastShardSizes() const { return mLastShardSizes; }

private:
    std::vector<std::shared_ptr<LoopDetectorBranchBound>> mShards;
    std::vector<DeviceContextPtr> mContexts;
    std::vector<csm_result> mLastResults;
    std::vector<int> mLastShardSizes;
    std::uint64_t mBestWord = 0;
    bool mNccl = false;
};

/* loop_detector_correlative.cpp:59-159: per query the coarse map (window = the matcher's low
 * resolution) of a local map is built at its first use and stays cached on the device by
 * LocalMapId; real-time correlative match with thresholds, then the final matcher. */
class LoopDetectorCorrelative final : public LoopDetector
{
public:
    LoopDetectorCorrelative(const std::string& name,
                            const std::shared_ptr<ScanMatcherCorrelative>& scan_matcher,
                            const FinalMatcher& final_matcher,
                            double score_threshold, double known_rate_threshold);
    std::vector<LoopDetectionResult> Detect(const std::vector<LoopDetectionQuery>& queries) override;
    /* More matchers of the same parameters, each on its own device context (the same GPU or another one):
     * the coarse stage of a Detect then runs on all of them at once, query i on matcher LocalMapId mod L.
     * Same results as with one matcher. */
    void SetConcurrentMatchers(const std::vector<std::shared_ptr<ScanMatcherCorrelative>>& extra) { mExtraMatchers = extra; }

private:
    std::shared_ptr<ScanMatcherCorrelative> mScanMatcher;
    std::vector<std::shared_ptr<ScanMatcherCorrelative>> mExtraMatchers;
    FinalMatcher mFinalMatcher;
    double mScoreThreshold, mKnownRateThreshold;
};

/* loop_detector_grid_search.cpp:52-161: exhaustive grid search with thresholds per query */
class LoopDetectorGridSearch final : public LoopDetector
{
public:
    LoopDetectorGridSearch(const std::string& name,
                           const std::shared_ptr<ScanMatcherGridSearch>& scan_matcher,
                           const FinalMatcher& final_matcher,
                           double score_threshold, double known_rate_threshold);
    std::vector<LoopDetectionResult> Detect(const std::vector<LoopDetectionQuery>& queries) override;
    void SetConcurrentMatchers(const std::vector<std::shared_ptr<ScanMatcherGridSearch>>& extra) { mExtraMatchers = extra; }

private:
    std::shared_ptr<ScanMatcherGridSearch> mScanMatcher;
    std::vector<std::shared_ptr<ScanMatcherGridSearch>> mExtraMatchers;
    FinalMatcher mFinalMatcher;
    double mScoreThreshold, mKnownRateThreshold;
};

} /* namespace csm_host */
