/* loop_detector.hpp -- batched GPU loop detector behind the reference's
 * LoopDetector plugin interface (mapping/loop_detector.hpp:97-116:
 * Detect(queries) -> results), JSON type string "BranchBound"
 * (loop_detector_factory.cpp:202-209). Constructor parameters follow
 * LoopDetectorBranchBound (loop_detector_branch_bound.cpp:41-56); the final
 * sub-pixel matcher is an optional callback because it stays on the CPU
 * (SURVEY.md 8f rank 1). */
#pragma once

#include <functional>
#include <set>

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

protected:
    std::string mName;
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
    /* Without a final matcher the result carries the covariance of the cost
     * function at the coarse pose (computed on the CPU); switch it off when
     * the caller refines the poses itself */
    void SetCoarseCovariance(bool on) { mCoarseCovariance = on; }
    /* Forget which maps are resident (the next Detect uploads them again) */
    void ClearCache() { mCachedMaps.clear(); mCachedScans.clear(); }
    /* Sharded use: global index of queries[0] (packed best word) */
    void SetQueryIndexBase(int base) { mQueryIndexBase = base; }
    /* Per-query device results of the last Detect, in query order */
    const std::vector<csm_result>& LastResults() const { return mLastResults; }

private:
    std::shared_ptr<ScanMatcherBranchBound> mScanMatcher;
    FinalMatcher mFinalMatcher;
    double mScoreThreshold, mKnownRateThreshold;
    std::set<std::int64_t> mCachedMaps;
    std::set<std::int64_t> mCachedScans;
    std::vector<csm_result> mLastResults;
    int mQueryIndexBase = 0;
    int mChunkSize = 128;
    int mUploadChunk = 64;
    bool mCoarseCovariance = true;
};

} /* namespace csm_host */
