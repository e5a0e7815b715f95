/* map_builder.hpp -- the reference's GridMapBuilder (mapping/grid_map_builder.hpp:88-300,
 * grid_map_builder.cpp) with every map kept ON THE DEVICE.
 *
 * The reference rebuilds its "latest map" from the last few scans for every new scan
 * (UpdateLatestMap -> ConstructMapFromScans, grid_map_builder.cpp:497-532, 561-695), hands a deep copy to
 * the front end (lidar_graph_slam.cpp:224-270) and the scan matcher reads it; it also casts every scan into
 * the current local map (UpdateGridMap, :390-494), which the loop detector reads once it is finished. With
 * GPU matchers that is host-side ray casting plus an upload per scan and per local map. Here the host keeps
 * what is cheap and must follow the reference to the bit -- the pose graph, the hit points of the scans and
 * their bounding box (libm sine and cosine), the geometry bookkeeping of GridMap::Resize / Expand
 * (grid_map.cpp:842-946) -- and the device casts the rays into the maps the matchers read (csm_map_*):
 * no map crosses PCIe, in either direction.
 */
#pragma once

#include <cstdint>
#include <cmath>
#include <memory>
#include <unordered_map>
#include <vector>

#include "csm_host/pose_graph.hpp"
#include "csm_host/scan_matchers.hpp"

namespace csm_host {

/* pose_graph_node.hpp:41-73, what map construction reads of a scan node */
struct ScanNodeView
{
    Pose2D global_pose;
    ScanDataPtr scan;
};

/* GridMap<GridBinaryBayes> (grid_map_new/grid_map.hpp) whose cells live on the device under `map_id`:
 * the geometry (grid_map_geometry.hpp) and its Resize / Expand rules on the host, the cells behind csm_map_*. */
class DeviceGridMap
{
public:
    struct Index { int x, y; };
    /* the hit points of one scan in the map's frame, and its sensor position */
    struct ScanHits
    {
        Pose2D sensor;
        std::vector<double> x, y;
        /* `fast`: x, y come from rotating the scan's own polar points (r cos a, r sin a) by the sensor
         * heading instead of libm's cos / sin of (heading + a): a few 1e-14 m away from the reference's values.
         * They are only ever floored to cell indices; whoever floors them checks the distance to the cell
         * boundary and re-evaluates the beam the reference's way when it is closer than the guard band
         * (`scan`, `beam`: what that takes). */
        bool fast = false;
        const ScanData* scan = nullptr;
        std::vector<int> beam;
        /* the reference's own arithmetic for kept beam k (sensor_data.hpp:190-203) */
        void Exact(std::size_t k, double& hx, double& hy) const
        {
            const double c = std::cos(sensor.theta + scan->angles[beam[k]]);
            const double s = std::sin(sensor.theta + scan->angles[beam[k]]);
            hx = sensor.x + scan->ranges[beam[k]] * c;
            hy = sensor.y + scan->ranges[beam[k]] * s;
        }
    };

    /* GridMap(resolution, blockSize, 1.0, 1.0) (grid_map.cpp:74-99, 226-246) */
    DeviceGridMap(const DeviceContextPtr& context, std::int64_t map_id, double resolution, int log2_block_size);
    ~DeviceGridMap();
    DeviceGridMap(const DeviceGridMap&) = delete;
    DeviceGridMap& operator=(const DeviceGridMap&) = delete;

    Index PositionToIndex(double x, double y) const;                       /* grid_map_geometry.cpp:113-122 */
    void Resize(double min_x, double min_y, double max_x, double max_y);   /* grid_map.cpp:891-911 */
    void Expand(double min_x, double min_y, double max_x, double max_y);   /* grid_map.cpp:914-946 */
    void ResetValues();                                                    /* grid_map.cpp: every block dropped */
    /* :642-692 / :445-480: per beam the sub-pixel indices of sensor and hit point and the hit cell, in the
     * order of `hits`; the device casts the rays. Returns the number of beams. */
    int InsertScans(const std::vector<ScanHits>& hits, int subpixel_scale);
    /* beams of the last InsertScans that sat inside the guard band and were re-evaluated exactly */
    int LastExactBeams() const { return mLastExactBeams; }
    /* |fraction to the nearest cell boundary| below which a floored fast coordinate is not trusted, in cells
     * (of the geometry it is floored in): 1e-9 against a possible difference of ~1e-10 sub-pixel cells */
    static constexpr double kGuardBand = 1e-9;
    static double& GuardBand() { static double band = kGuardBand; return band; }      /* tests widen it */
    static bool NearBoundary(double cells) { const double f = cells - std::floor(cells); return f < GuardBand() || 1.0 - f < GuardBand(); }

    GridMapView View() const;            /* device_resident: the matchers read the map where it is */
    std::int64_t MapId() const { return mMapId; }
    int Rows() const { return mRows; }
    int Cols() const { return mCols; }
    double Resolution() const { return mResolution; }
    double OffsetX() const { return mOffX; }
    double OffsetY() const { return mOffY; }
    int BlockSize() const { return 1 << mLog2BlockSize; }
    /* grid_map_geometry.cpp:125-135 (IndexToPosition: the cell's minimum corner) */
    void IndexToPosition(int row, int col, double& x, double& y) const
    { x = mOffX + mResolution * col; y = mOffY + mResolution * row; }

private:
    void ResizeIndex(int box_min_x, int box_min_y, int box_max_x, int box_max_y);   /* grid_map.cpp:842-888 */

    DeviceContextPtr mContext;
    std::int64_t mMapId;
    double mResolution;
    int mLog2BlockSize;
    int mBlockRows, mBlockCols, mRows, mCols;
    double mOffX, mOffY;
    int mLastExactBeams = 0;
};

/* grid_map_builder.hpp:29-85 */
struct LocalMapGPU
{
    int id = 0;
    std::unique_ptr<DeviceGridMap> map;
    int scan_node_id_min = 0, scan_node_id_max = 0;
    bool finished = false;
};

class GridMapBuilderGPU
{
public:
    /* grid_map_builder.hpp:146-155. The latest map lives under `latest_map_device_id`, local map k under
     * device map id k (its LocalMapId: what the loop detector's cache is keyed by). */
    GridMapBuilderGPU(const DeviceContextPtr& context, double map_resolution, int patch_size,
                      int num_of_scans_for_latest_map, double usable_range_min, double usable_range_max,
                      double prob_hit, double prob_miss, std::int64_t latest_map_device_id = (std::int64_t(1) << 41),
                      bool reference_table_end = true);
    /* TravelDistThresholdForLocalMap, NumOfOverlappedScans (launcher_settings_default.json:180-181) */
    void SetLocalMapPolicy(double travel_dist_threshold, int num_of_overlapped_scans)
    { mTravelDistThreshold = travel_dist_threshold; mNumOfOverlappedScans = num_of_overlapped_scans; }

    /* GridMapBuilder::AppendScan (:120-135): a new scan node and its odometry edge(s), a new local map when
     * the robot has travelled far enough (UpdatePoseGraph, :290-385; AppendLocalMap, :189-286), the scan cast
     * into the current local map (UpdateGridMap, :390-494). Returns whether a local map was inserted. */
    bool AppendScan(PoseGraph& pose_graph, const Pose2D& relative_scan_pose, const Mat3& scan_pose_covariance,
                    const ScanDataPtr& scan);
    /* GridMapBuilder::UpdateLatestMap (:497-532): the latest map from the last NumOfScansForLatestMap nodes */
    void UpdateLatestMap(const std::vector<ScanNode>& scan_nodes);
    void UpdateLatestMap(const std::vector<ScanNodeView>& scan_nodes);
    /* GridMapBuilder::AfterLoopClosure (:138-144) */
    void AfterLoopClosure(const PoseGraph& pose_graph);
    /* GridMapBuilder::FinishLocalMap (:147-165) */
    void FinishLocalMap();

    const std::vector<LocalMapGPU>& LocalMaps() const { return mLocalMaps; }
    double AccumTravelDist() const { return mAccumTravelDist; }
    int LatestScanIdMin() const { return mLatestScanIdMin; }
    int LatestScanIdMax() const { return mLatestScanIdMax; }

    /* The latest map as the matchers take it: resident on the device (no host cells) */
    GridMapView LatestMap() const { return mLatest.View(); }
    const DeviceGridMap& LatestGrid() const { return mLatest; }
    const Pose2D& LatestMapPose() const { return mLatestMapPose; }
    int Rows() const { return mLatest.Rows(); }
    int Cols() const { return mLatest.Cols(); }
    double OffsetX() const { return mLatest.OffsetX(); }
    double OffsetY() const { return mLatest.OffsetY(); }
    int BlockSize() const { return mLatest.BlockSize(); }
    const DeviceContextPtr& Context() const { return mContext; }
    /* rays handed to the device by the last map update */
    int LastNumOfRays() const { return mLastRays; }

    /* The tables of GridBinaryBayes::UpdateOddsUnchecked (grid_binary_bayes.cpp:302-321) for one odds
     * value: table[v] = the cell value after the update of a cell that holds v.
     * reference_table_end: a cell at 65535 behaves as in the compiled reference (it reads past the
     * end of the reference's 65535-entry odds table, see map_builder.cpp) */
    static std::vector<std::uint16_t> UpdateTable(double odds, bool reference_table_end = true);
    static constexpr int SubpixelScale = 100;       /* grid_map_builder.hpp:294 */

private:
    /* ConstructMapFromScans (:561-695): `map` rebuilt from `count` scan nodes in the frame `map_pose` */
    void ConstructMapFromScans(const Pose2D& map_pose, DeviceGridMap& map, const ScanNodeView* nodes, int count);
    /* ComputeBoundingBoxAndScanPointsMapLocal (:820-872) and the first loop of ConstructMapFromScans */
    DeviceGridMap::ScanHits HitsOf(const Pose2D& map_pose, const Pose2D& global_scan_pose, const ScanDataPtr& scan);
    /* bounding box of sensor and hit points; when a fast coordinate decides a floor of Resize / Expand inside
     * the guard band, every hit point is re-evaluated exactly and the box formed again */
    void BoundingBox(std::vector<DeviceGridMap::ScanHits>& hits, const DeviceGridMap& map, bool construct,
                     double& min_x, double& min_y, double& max_x, double& max_y) const;
    /* a scan's polar points (r cos a, r sin a), libm once per scan: the scans of the latest map come back
     * ten times, the map frame changes every time */
    struct Polar { ScanDataPtr keep; std::vector<double> px, py; };      /* `keep`: the address stays this scan's */
    const Polar& PolarOf(const ScanDataPtr& scan);
    void AppendLocalMap(PoseGraph& pose_graph, const Pose2D& scan_pose, const Mat3& covariance, int scan_node_id);
    void UpdateGridMap(const PoseGraph& pose_graph);

    DeviceContextPtr mContext;
    double mResolution;
    int mLog2BlockSize;
    int mNumOfScansForLatestMap;
    double mUsableRangeMin, mUsableRangeMax;
    double mOddsHit, mOddsMiss;
    DeviceGridMap mLatest;
    Pose2D mLatestMapPose;
    int mLatestScanIdMin = 0, mLatestScanIdMax = 0;
    std::vector<LocalMapGPU> mLocalMaps;
    double mAccumTravelDist = 0.0, mTravelDistLastLocalMap = 0.0;
    double mTravelDistThreshold = 2.5;
    int mNumOfOverlappedScans = 10;
    int mLastRays = 0;
    bool mFastHitPoints = true;
    std::unordered_map<const ScanData*, Polar> mPolar;

public:
    /* false: every hit point with libm, like the reference (the fast path gives the same cells; tests compare) */
    void SetFastHitPoints(bool on) { mFastHitPoints = on; }
};

} /* namespace csm_host */
