/* loop_searcher.hpp -- the caller that produces the loop-detection query batch
 * (SURVEY.md section 8f, rank 3): LoopSearcherNearest of the reference
 * (mapping/loop_searcher_nearest.hpp:37-68, loop_searcher_nearest.cpp:59-170;
 * interface mapping/loop_searcher.hpp:25-101). Pure host logic on pose-graph
 * summaries, no device work: it decides WHICH (query scan node, reference scan
 * node, reference local map) triples the detector is asked about. With the
 * batched detector the useful NumOfCandidateNodes grows from the reference's
 * default 2 (launcher_settings_default.json:64) to hundreds. */
#ifndef CSM_HOST_LOOP_SEARCHER_HPP
#define CSM_HOST_LOOP_SEARCHER_HPP

#include <cstdint>
#include <string>
#include <vector>

#include "csm_host/types.hpp"

namespace csm_host {

/* ScanNodeData (pose_graph_node.hpp:76-95) */
struct ScanNodeData
{
    int node_id;
    Pose2D global_pose;
};

/* LocalMapData (grid_map_builder.hpp:53-85), without the bounding box the searcher never reads */
struct LocalMapData
{
    int local_map_id;
    int scan_node_id_min, scan_node_id_max;
    bool finished;
};

/* LoopSearchHint (loop_searcher.hpp:25-53): nodes in ascending id order, like the reference's IdMap */
struct LoopSearchHint
{
    std::vector<ScanNodeData> scan_nodes;
    std::vector<LocalMapData> local_map_nodes;
    double accum_travel_dist;
    int last_finished_scan_id;
    int last_finished_map_id;
};

/* LoopCandidate (loop_searcher.hpp:62-82) */
struct LoopCandidate
{
    int query_scan_node_id;
    int reference_scan_node_id;
    int reference_local_map_id;
};

class LoopSearcher
{
public:
    virtual ~LoopSearcher() = default;
    virtual std::vector<LoopCandidate> Search(const LoopSearchHint& hint) = 0;
};

class LoopSearcherNearest final : public LoopSearcher
{
public:
    /* same parameters as the reference's constructor (loop_searcher_nearest.hpp:41-48) */
    LoopSearcherNearest(double travel_dist_threshold, double node_dist_threshold, int num_of_candidate_nodes) :
        mTravelDistThreshold(travel_dist_threshold), mNodeDistThreshold(node_dist_threshold),
        mNumOfCandidateNodes(num_of_candidate_nodes) { }

    std::vector<LoopCandidate> Search(const LoopSearchHint& hint) override;
    /* squared node distances of the candidates returned by the last Search, in the same order
     * (the reference observes them as the LoopSearcherNearest.NodeDist metric, :163-164) */
    const std::vector<double>& LastNodeDistances() const { return mLastDistances; }

private:
    double mTravelDistThreshold, mNodeDistThreshold;
    int mNumOfCandidateNodes;
    std::vector<double> mLastDistances;
};

} /* namespace csm_host */

#endif
