/* types.hpp -- plain data types of the host-side plugin mirror.
 *
 * They carry exactly what the reference's plugin interface passes around
 * (paths relative to the reference repository):
 *   RobotPose2D<double>            pose.hpp:20-60           -> Pose2D
 *   Sensor::ScanData<double>       sensor/sensor_data.hpp:65-178 -> ScanData
 *   Mapping::GridMap               grid_map_new/grid_map.hpp:27  -> GridMapView (dense flattening)
 *   ScanMatchingQuery / Summary    mapping/scan_matcher.hpp:28-83
 *   LoopDetectionQuery / Result    mapping/loop_detector.hpp:27-92
 * A maintainer integrating into the reference converts at this boundary (see
 * INTEGRATION.md); nothing here depends on Boost, Eigen or the reference headers.
 */
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace csm_host {

struct Pose2D
{
    double x = 0.0, y = 0.0, theta = 0.0;
};

/* pose.hpp:154-166 */
inline Pose2D Compound(const Pose2D& start, const Pose2D& diff)
{
    const double s = std::sin(start.theta), c = std::cos(start.theta);
    return Pose2D { c * diff.x - s * diff.y + start.x,
                    s * diff.x + c * diff.y + start.y,
                    start.theta + diff.theta };
}

/* pose.hpp:183-198 */
inline Pose2D InverseCompound(const Pose2D& start, const Pose2D& end)
{
    const double s = std::sin(start.theta), c = std::cos(start.theta);
    const double dx = end.x - start.x, dy = end.y - start.y;
    return Pose2D { c * dx + s * dy, -s * dx + c * dy, end.theta - start.theta };
}

/* pose.hpp:211-224 */
inline Pose2D MoveBackward(const Pose2D& end, const Pose2D& diff)
{
    const double theta = end.theta - diff.theta;
    const double s = std::sin(theta), c = std::cos(theta);
    return Pose2D { end.x - c * diff.x + s * diff.y, end.y - s * diff.x - c * diff.y, theta };
}

/* One laser scan: beam angles / ranges and the sensor pose on the robot */
struct ScanData
{
    std::vector<double> angles;
    std::vector<double> ranges;
    Pose2D relative_sensor_pose;
    double min_range = 0.0;           /* ScanData::MinRange / MaxRange (sensor_data.hpp:106-108): map */
    double max_range = 1e300;         /* construction skips beams outside (min, max) */
    std::size_t NumOfScans() const { return ranges.size(); }
};
using ScanDataPtr = std::shared_ptr<const ScanData>;

/* View of an occupancy grid: value 0 = unknown, 1..65535 <-> probability
 * 0.001..0.999 (grid_binary_bayes.hpp:163-176); cell (row, col) covers
 * [offset + res * col, offset + res * (col + 1)). `map_id` >= 0 names a
 * finished local map (LocalMapId) whose device copy and pyramid are cached;
 * -1 = anonymous map, uploaded on every call.
 *
 * Two forms:
 *  - block-sparse (`blocks` non-null): the reference's own storage, the
 *    n_blocks allocated 2^k x 2^k blocks back to back (row-major inside a
 *    block) and their positions block_row * (cols >> k) + block_col
 *    (grid_map.cpp:262-266, 522-535). Nothing is flattened and only these
 *    bytes are uploaded. Block allocation is exact.
 *  - dense (`values` non-null): row-major u16. `block_allocated` (optional,
 *    (rows/16) x (cols/16)) tells which 16x16 blocks the reference map has
 *    allocated; it only matters for the CPU cost function (cells of
 *    unallocated blocks read 0.5 there, grid_map.cpp:424-436 with
 *    cost_function_square_error.cpp:340-344). When null, a block counts as
 *    allocated iff it holds a non-zero cell. */
struct GridMapView
{
    const std::uint16_t* values = nullptr;
    int rows = 0, cols = 0;
    double resolution = 0.0;
    double offset_x = 0.0, offset_y = 0.0;
    std::int64_t map_id = -1;
    const std::uint8_t* block_allocated = nullptr;
    const std::uint16_t* blocks = nullptr;
    const std::int32_t* block_index = nullptr;
    int n_blocks = 0;
    int log2_block_size = 4;
    /*  - block-sparse, blocks where the reference keeps them (`block_ptrs` non-null): every allocated
     *    block is its own heap allocation there (grid_map.cpp:522-535), block_ptrs[b] points at block
     *    block_index[b]. The loop detector gathers them into page-locked staging with several threads,
     *    group by group, while the previous group crosses PCIe. */
    const std::uint16_t* const* block_ptrs = nullptr;
    /*  - device-resident (`device_resident`): the map was built on the device under map_id
     *    (GridMapBuilderGPU) and has no host cells; matchers read it where it is. Their cost /
     *    covariance epilogue must then run on the device too (DeviceContext::SetDeviceEpilogue /
     *    SetDeviceFinalMatcher). */
    bool device_resident = false;
};

/* scan_matcher.hpp:56-83 */
struct ScanMatchingSummary
{
    bool pose_found = false;
    double normalized_cost = 0.0;
    Pose2D map_local_initial_pose;
    Pose2D estimated_pose;
    std::array<double, 9> estimated_covariance {};   /* row-major 3x3 */
    /* what the device decided (window indices, integer score, double score) */
    int best_x = 0, best_y = 0, best_theta = 0;
    std::int64_t sum_value = 0;
    int n_known = 0;
    double normalized_score = 0.0;
    int flags = 0;
    int n_processed = 0, n_ignored = 0;
};

/* loop_detector.hpp:27-55 with the references resolved */
struct LoopDetectionQuery
{
    ScanDataPtr scan;                 /* mQueryScanNode.mScanData */
    std::int64_t scan_id = 0;         /* identifies the scan data across queries of a batch */
    int scan_node_id = 0;             /* mQueryScanNode.mNodeId */
    Pose2D scan_global_pose;          /* mQueryScanNode.mGlobalPose */
    GridMapView local_map;            /* mReferenceLocalMap.mMap, map_id = LocalMapId */
    Pose2D local_map_global_pose;     /* mReferenceLocalMapNode.mGlobalPose */
    Pose2D reference_scan_local_pose; /* mReferenceScanNode.mLocalPose (map centre for the refiner) */
};

/* loop_detector.hpp:65-92 */
struct LoopDetectionResult
{
    Pose2D relative_pose;
    Pose2D local_map_pose;
    std::int64_t local_map_id = -1;
    int scan_node_id = 0;
    std::array<double, 9> estimated_covariance {};
    double normalized_score = 0.0;
    int query_index = 0;
};

} /* namespace csm_host */
