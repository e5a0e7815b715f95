#include "csm_host/loop_searcher.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace csm_host {

namespace {

struct CandidateDistance
{
    int ref_map_id, ref_scan_id, query_scan_id;
    double dist_sq;
};

[[noreturn]] void Fail(const char* what)
{
    /* the reference Asserts (util.hpp:39-72) */
    std::fprintf(stderr, "csm_host: LoopSearcherNearest: %s\n", what);
    std::abort();
}

/* position of the node with this id (the nodes are sorted by id) */
std::size_t IndexOfNode(const std::vector<ScanNodeData>& nodes, int id)
{
    const auto it = std::lower_bound(nodes.begin(), nodes.end(), id,
                                     [](const ScanNodeData& n, int v) { return n.node_id < v; });
    if (it == nodes.end() || it->node_id != id)
        Fail("scan node id not in the hint");
    return static_cast<std::size_t>(it - nodes.begin());
}

} /* namespace */

/* loop_searcher_nearest.cpp:59-170 */
std::vector<LoopCandidate> LoopSearcherNearest::Search(const LoopSearchHint& hint)
{
    mLastDistances.clear();
    const std::vector<ScanNodeData>& scan_nodes = hint.scan_nodes;
    const std::vector<LocalMapData>& maps = hint.local_map_nodes;
    if (maps.empty() || maps.back().local_map_id != hint.last_finished_map_id)
        Fail("the last local map of the hint must be the last finished one");       /* :75 */
    const double accum_travel_dist = hint.accum_travel_dist;
    const double node_dist_threshold_sq = std::pow(mNodeDistThreshold, 2.0);          /* :68 */

    /* query scan nodes: those of the latest (last finished) local map, :77-79 */
    const LocalMapData& query_map = maps.back();
    const std::size_t q0 = IndexOfNode(scan_nodes, query_map.scan_node_id_min);
    const std::size_t q1 = IndexOfNode(scan_nodes, query_map.scan_node_id_max);

    std::vector<CandidateDistance> distances;
    double node_travel_dist = 0.0;
    bool first = true, done = false;
    Pose2D prev { 0.0, 0.0, 0.0 };
    /* reference local maps: all but the last, :82-83 */
    for (std::size_t m = 0; m + 1 < maps.size() && !done; ++m) {
        const LocalMapData& ref_map = maps[m];
        if (!ref_map.finished)
            Fail("a reference local map is not finished");                            /* :91 */
        const std::size_t r0 = IndexOfNode(scan_nodes, ref_map.scan_node_id_min);
        const std::size_t r1 = IndexOfNode(scan_nodes, ref_map.scan_node_id_max);
        for (std::size_t r = r0; r <= r1; ++r) {
            const Pose2D& ref_pose = scan_nodes[r].global_pose;
            /* travel distance along the reference nodes, :108-110 */
            node_travel_dist += first ? 0.0 : std::hypot(prev.x - ref_pose.x, prev.y - ref_pose.y);
            prev = ref_pose;
            first = false;
            /* nodes too close in travel distance to the current one end the search, :114-115 */
            if (accum_travel_dist - node_travel_dist < mTravelDistThreshold) {
                done = true;
                break;
            }
            for (std::size_t q = q0; q <= q1; ++q) {
                const Pose2D& query_pose = scan_nodes[q].global_pose;
                const double d = (ref_pose.x - query_pose.x) * (ref_pose.x - query_pose.x) +
                                 (ref_pose.y - query_pose.y) * (ref_pose.y - query_pose.y);     /* :123 */
                if (d < node_dist_threshold_sq)
                    distances.push_back(CandidateDistance { ref_map.local_map_id, scan_nodes[r].node_id,
                                                            scan_nodes[q].node_id, d });
            }
        }
    }
    if (distances.empty())
        return { };
    /* the closest NumOfCandidateNodes pairs, in the order std::nth_element leaves them, :142-157 */
    const std::size_t n = std::min(static_cast<std::size_t>(mNumOfCandidateNodes), distances.size());
    std::nth_element(distances.begin(), distances.begin() + n, distances.end(),
                     [](const CandidateDistance& a, const CandidateDistance& b) { return a.dist_sq < b.dist_sq; });
    std::vector<LoopCandidate> out;
    out.reserve(n);
    for (std::size_t i = 0; i < n; ++i) {
        out.push_back(LoopCandidate { distances[i].query_scan_id, distances[i].ref_scan_id, distances[i].ref_map_id });
        mLastDistances.push_back(distances[i].dist_sq);
    }
    return out;
}

} /* namespace csm_host */
