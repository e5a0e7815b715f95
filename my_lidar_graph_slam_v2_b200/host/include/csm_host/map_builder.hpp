/* map_builder.hpp -- the map-construction half of the reference's GridMapBuilder
 * (mapping/grid_map_builder.cpp) with the map kept ON THE DEVICE.
 *
 * The reference rebuilds its "latest map" from the last few scans for every new scan
 * (UpdateLatestMap -> ConstructMapFromScans, grid_map_builder.cpp:497-532, 561-695), hands a deep copy to
 * the front end (lidar_graph_slam.cpp:224-270) and the scan matcher reads it: with a GPU matcher that is
 * a host-side ray casting plus an upload per scan. Here the host keeps what is cheap and must follow the
 * reference to the bit -- the scan nodes, the bounding box of their hit points, the geometry bookkeeping
 * of GridMap::Resize (grid_map.cpp:842-911), the sub-pixel indices of every beam -- and the device does
 * the ray casting into the map the matchers read (csm_map_*): no map crosses PCIe.
 */
#pragma once

#include <cstdint>
#include <vector>

#include "csm_host/scan_matchers.hpp"

namespace csm_host {

/* pose_graph_node.hpp:41-73, what map construction reads of a scan node */
struct ScanNodeView
{
    Pose2D global_pose;
    ScanDataPtr scan;
};

class GridMapBuilderGPU
{
public:
    /* grid_map_builder.hpp:146-155 (the parameters that concern the latest map) */
    GridMapBuilderGPU(const DeviceContextPtr& context, double map_resolution, int patch_size,
                      int num_of_scans_for_latest_map, double usable_range_min, double usable_range_max,
                      double prob_hit, double prob_miss, std::int64_t device_map_id = (std::int64_t(1) << 41),
                      bool reference_table_end = true);

    /* GridMapBuilder::UpdateLatestMap: the latest map from the last NumOfScansForLatestMap nodes */
    void UpdateLatestMap(const std::vector<ScanNodeView>& scan_nodes);

    /* The latest map as the matchers take it: resident on the device under map_id (no host cells) */
    GridMapView LatestMap() const;
    const Pose2D& LatestMapPose() const { return mLatestMapPose; }
    int Rows() const { return mRows; }
    int Cols() const { return mCols; }
    double OffsetX() const { return mOffX; }
    double OffsetY() const { return mOffY; }
    int BlockSize() const { return 1 << mLog2BlockSize; }
    const DeviceContextPtr& Context() const { return mContext; }
    /* rays handed to the device by the last update */
    int LastNumOfRays() const { return mLastRays; }

    /* The tables of GridBinaryBayes::UpdateOddsUnchecked (grid_binary_bayes.cpp:302-321) for one odds
     * value: table[v] = the cell value after the update of a cell that holds v */
    /* reference_table_end: a cell at 65535 behaves as in the compiled reference (it reads past the
     * end of the reference's 65535-entry odds table, see map_builder.cpp) */
    static std::vector<std::uint16_t> UpdateTable(double odds, bool reference_table_end = true);
    static constexpr int SubpixelScale = 100;       /* grid_map_builder.hpp:294 */

private:
    struct Index { int x, y; };
    Index PositionToIndex(double x, double y) const;
    void Resize(double min_x, double min_y, double max_x, double max_y);

    DeviceContextPtr mContext;
    std::int64_t mMapId;
    double mResolution;
    int mLog2BlockSize;
    int mNumOfScansForLatestMap;
    double mUsableRangeMin, mUsableRangeMax;
    double mOddsHit, mOddsMiss;
    /* GridMap geometry (grid_map.hpp, grid_map_geometry.hpp) */
    int mBlockRows, mBlockCols, mRows, mCols;
    double mOffX, mOffY;
    Pose2D mLatestMapPose;
    int mLastRays = 0;
};

} /* namespace csm_host */
