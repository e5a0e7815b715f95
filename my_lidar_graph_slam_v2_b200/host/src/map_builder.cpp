#include "csm_host/map_builder.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

namespace csm_host {

namespace {

int ToNearestPowerOf2(int x)        /* grid_map.cpp:22-35 */
{
    --x;
    x |= (x >> 1); x |= (x >> 2); x |= (x >> 4); x |= (x >> 8); x |= (x >> 16);
    return x + 1;
}

/* GridBinaryBayes constants and conversions (grid_binary_bayes.hpp:163-176, .cpp:345-388;
 * grid_values.hpp:13-61), same expressions, same order */
constexpr double kProbMin = 1e-3;
constexpr double kProbMax = 1.0 - kProbMin;
constexpr std::uint16_t kValueMin = 1, kValueMax = 65535;

double ValueToProbability(std::uint16_t value)
{
    return kProbMin + (kProbMax - kProbMin) * static_cast<double>(value - kValueMin) /
           static_cast<double>(kValueMax - kValueMin);
}

double ProbabilityToOdds(double prob)
{
    if (prob == 0.0) return 1.0;
    if (prob < kProbMin) return kProbMin / (1.0 - kProbMin);
    if (prob > kProbMax) return kProbMax / (1.0 - kProbMax);
    return prob / (1.0 - prob);
}

double OddsToProbability(double odds)
{
    if (odds < 0.0) return 0.0;
    const double prob = odds / (1.0 + odds);
    return std::clamp(prob, kProbMin, kProbMax);
}

std::uint16_t ProbabilityToValue(double prob)
{
    if (prob == 0.0) return 0;
    if (prob < kProbMin) return kValueMin;
    if (prob > kProbMax) return kValueMax;
    return static_cast<std::uint16_t>(kValueMin + (prob - kProbMin) * static_cast<double>(kValueMax - kValueMin) /
                                      (kProbMax - kProbMin));
}

int IndexToBlock(int idx, int log2bs)        /* grid_map.cpp:803-815, the quirk for negatives included */
{
    return idx >= 0 ? (idx >> log2bs) : ((idx >> log2bs) - 1);
}

} /* namespace */

std::vector<std::uint16_t> GridMapBuilderGPU::UpdateTable(double odds, bool reference_table_end)
{
    std::vector<std::uint16_t> table(65536);
    /* an unknown cell takes the observation itself (grid_binary_bayes.cpp:309-313) */
    table[0] = ProbabilityToValue(OddsToProbability(odds));
    for (int v = 1; v < 65536; ++v) {
        /* ValueToOddsLookup[v] = ValueToOdds(v) (grid_values.cpp:75-84) */
        const double prob = ValueToProbability(static_cast<std::uint16_t>(v));
        const double old_odds = prob / (1.0 - prob);
        table[v] = ProbabilityToValue(OddsToProbability(old_odds * odds));
    }
    /* The reference's lookup holds ValueMax - ValueMin + 1 = 65535 entries (grid_values.cpp:72-74), so a
     * cell that sits at ValueMax = 65535 reads one element past its end on the next update
     * (grid_binary_bayes.cpp:316). The vector is 524280 bytes, which glibc serves from its own mmap'd
     * chunk; the element past the end lies in that chunk's zero tail, so the reference computes with old
     * odds 0.0 and the saturated cell drops to ValueMin -- after a hit as well as after a miss. The
     * reference compiled with g++ does exactly that, and it is what "identical maps" means here;
     * reference_table_end = false continues the table instead (65535 stays saturated under hits). */
    if (reference_table_end)
        table[65535] = ProbabilityToValue(OddsToProbability(0.0 * odds));
    return table;
}

GridMapBuilderGPU::GridMapBuilderGPU(const DeviceContextPtr& context, double map_resolution, int patch_size,
                                     int num_of_scans_for_latest_map, double usable_range_min,
                                     double usable_range_max, double prob_hit, double prob_miss,
                                     std::int64_t device_map_id, bool reference_table_end) :
    mContext(context), mMapId(device_map_id), mResolution(map_resolution),
    mLog2BlockSize(__builtin_ctz(ToNearestPowerOf2(patch_size))),
    mNumOfScansForLatestMap(num_of_scans_for_latest_map),
    mUsableRangeMin(usable_range_min), mUsableRangeMax(usable_range_max),
    mOddsHit(ProbabilityToOdds(prob_hit)), mOddsMiss(ProbabilityToOdds(prob_miss))
{
    /* mLatestMap(mapResolution, patchSize, 1.0, 1.0) (grid_map_builder.cpp:80; grid_map.cpp:74-99, 226-246) */
    const int bs = 1 << mLog2BlockSize;
    const int desired_rows = static_cast<int>(std::ceil(1.0 / mResolution));
    const int desired_cols = static_cast<int>(std::ceil(1.0 / mResolution));
    mBlockRows = (desired_rows + bs - 1) >> mLog2BlockSize;
    mBlockCols = (desired_cols + bs - 1) >> mLog2BlockSize;
    mRows = mBlockRows << mLog2BlockSize;
    mCols = mBlockCols << mLog2BlockSize;
    mOffX = 0.0; mOffY = 0.0;
    csm_handle h = mContext->Handle();
    const std::vector<std::uint16_t> miss = UpdateTable(mOddsMiss, reference_table_end),
                                     hit = UpdateTable(mOddsHit, reference_table_end);
    mContext->Check(csm_map_set_update_tables(h, miss.data(), hit.data()), "csm_map_set_update_tables");
    mContext->Check(csm_map_create(h, mMapId, mRows, mCols, mLog2BlockSize, mResolution, mOffX, mOffY), "csm_map_create");
}

GridMapBuilderGPU::Index GridMapBuilderGPU::PositionToIndex(double x, double y) const
{
    /* grid_map_geometry.cpp:113-122 */
    return Index { static_cast<int>(std::floor((x - mOffX) / mResolution)),
                   static_cast<int>(std::floor((y - mOffY) / mResolution)) };
}

void GridMapBuilderGPU::Resize(double min_x, double min_y, double max_x, double max_y)
{
    /* GridMap::Resize(BoundingBox<double>) (grid_map.cpp:891-911) ... */
    const Index idx_min = PositionToIndex(min_x - mResolution, min_y - mResolution);
    const Index idx_max = PositionToIndex(max_x + mResolution, max_y + mResolution);
    const int box_min_x = idx_min.x, box_min_y = idx_min.y, box_max_x = idx_max.x + 1, box_max_y = idx_max.y + 1;
    /* ... -> Resize(BoundingBox<int>) (:842-888) */
    const int bs = 1 << mLog2BlockSize;
    const int block_min_x = IndexToBlock(box_min_x, mLog2BlockSize), block_min_y = IndexToBlock(box_min_y, mLog2BlockSize);
    const int block_max_x = IndexToBlock(box_max_x + bs - 1, mLog2BlockSize);
    const int block_max_y = IndexToBlock(box_max_y + bs - 1, mLog2BlockSize);
    const int block_rows = block_max_y - block_min_y, block_cols = block_max_x - block_min_x;
    const int row_min = block_min_y << mLog2BlockSize, col_min = block_min_x << mLog2BlockSize;
    const int rows = block_rows << mLog2BlockSize, cols = block_cols << mLog2BlockSize;
    /* GridMapGeometry::Resize (grid_map_geometry.cpp:63-75) */
    mBlockRows = block_rows; mBlockCols = block_cols;
    mRows = rows; mCols = cols;
    mOffX += mResolution * col_min;
    mOffY += mResolution * row_min;
    mContext->Check(csm_map_resize(mContext->Handle(), mMapId, rows, cols, row_min, col_min, mOffX, mOffY),
                    "csm_map_resize");
}

void GridMapBuilderGPU::UpdateLatestMap(const std::vector<ScanNodeView>& scan_nodes)
{
    if (scan_nodes.empty()) {
        std::fprintf(stderr, "csm_host: UpdateLatestMap needs at least one scan node\n");
        std::abort();
    }
    /* grid_map_builder.cpp:506-518: the last NumOfScansForLatestMap nodes; the map's frame is the pose of
     * the first of them */
    const int count = std::min(static_cast<int>(scan_nodes.size()), mNumOfScansForLatestMap);
    const std::size_t first = scan_nodes.size() - static_cast<std::size_t>(count);
    mLatestMapPose = scan_nodes[first].global_pose;

    /* ConstructMapFromScans, first pass (:578-633): hit points and bounding box in the map's frame */
    double min_x = std::numeric_limits<double>::max(), min_y = std::numeric_limits<double>::max();
    double max_x = std::numeric_limits<double>::min(), max_y = std::numeric_limits<double>::min();
    struct NodeHits { Pose2D sensor; std::vector<double> x, y; };
    std::vector<NodeHits> hits(count);
    for (int k = 0; k < count; ++k) {
        const ScanNodeView& node = scan_nodes[first + k];
        const ScanData& scan = *node.scan;
        const Pose2D global_sensor = Compound(node.global_pose, scan.relative_sensor_pose);
        const Pose2D local_sensor = InverseCompound(mLatestMapPose, global_sensor);
        hits[k].sensor = local_sensor;
        min_x = std::min(min_x, local_sensor.x); min_y = std::min(min_y, local_sensor.y);
        max_x = std::max(max_x, local_sensor.x); max_y = std::max(max_y, local_sensor.y);
        const double min_range = std::max(mUsableRangeMin, scan.min_range);
        const double max_range = std::min(mUsableRangeMax, scan.max_range);
        for (std::size_t i = 0; i < scan.NumOfScans(); ++i) {
            const double range = scan.ranges[i];
            if (range >= max_range || range <= min_range)
                continue;
            /* ScanData::HitPoint (sensor_data.hpp:190-203) */
            const double c = std::cos(local_sensor.theta + scan.angles[i]);
            const double s = std::sin(local_sensor.theta + scan.angles[i]);
            const double hx = local_sensor.x + range * c, hy = local_sensor.y + range * s;
            hits[k].x.push_back(hx); hits[k].y.push_back(hy);
            min_x = std::min(min_x, hx); min_y = std::min(min_y, hy);
            max_x = std::max(max_x, hx); max_y = std::max(max_y, hy);
        }
    }
    /* :636-638 */
    Resize(min_x, min_y, max_x, max_y);
    csm_handle h = mContext->Handle();
    mContext->Check(csm_map_reset_values(h, mMapId), "csm_map_reset_values");

    /* second pass (:642-692): per beam the sub-pixel indices of sensor and hit point and the hit cell */
    const double scaled_res = mResolution / SubpixelScale;           /* grid_map_geometry.cpp:48-60 */
    std::vector<csm_ray> rays;
    int order = 0;
    for (int k = 0; k < count; ++k) {
        const int sx = static_cast<int>(std::floor((hits[k].sensor.x - mOffX) / scaled_res));
        const int sy = static_cast<int>(std::floor((hits[k].sensor.y - mOffY) / scaled_res));
        for (std::size_t i = 0; i < hits[k].x.size(); ++i) {
            const Index hit = PositionToIndex(hits[k].x[i], hits[k].y[i]);
            csm_ray r;
            r.start_x = sx; r.start_y = sy;
            r.end_x = static_cast<int>(std::floor((hits[k].x[i] - mOffX) / scaled_res));
            r.end_y = static_cast<int>(std::floor((hits[k].y[i] - mOffY) / scaled_res));
            r.hit_col = hit.x; r.hit_row = hit.y;
            r.order = order++;
            r.reserved = 0;
            rays.push_back(r);
        }
    }
    mLastRays = static_cast<int>(rays.size());
    mContext->Check(csm_map_insert_rays(h, mMapId, rays.data(), mLastRays, SubpixelScale), "csm_map_insert_rays");
}

GridMapView GridMapBuilderGPU::LatestMap() const
{
    GridMapView v;
    v.rows = mRows; v.cols = mCols;
    v.resolution = mResolution;
    v.offset_x = mOffX; v.offset_y = mOffY;
    v.map_id = mMapId;
    v.log2_block_size = mLog2BlockSize;
    v.device_resident = true;
    return v;
}

} /* namespace csm_host */
