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

/* ---- 3 x 3 helpers of the pose graph (pose_graph.hpp) ---- */
Mat3 Multiply(const Mat3& a, const Mat3& b)
{
    Mat3 c {};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c[i * 3 + j] = a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j] + a[i * 3 + 2] * b[2 * 3 + j];
    return c;
}

Mat3 Transpose(const Mat3& a)
{
    return Mat3 { a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8] };
}

Mat3 Inverse(const Mat3& a)
{
    /* cofactors over the determinant, what Eigen does for a fixed 3 x 3 */
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double inv_det = 1.0 / (a[0] * c00 + a[1] * c01 + a[2] * c02);
    return Mat3 { c00 * inv_det, (a[2] * a[7] - a[1] * a[8]) * inv_det, (a[1] * a[5] - a[2] * a[4]) * inv_det,
                  c01 * inv_det, (a[0] * a[8] - a[2] * a[6]) * inv_det, (a[2] * a[3] - a[0] * a[5]) * inv_det,
                  c02 * inv_det, (a[1] * a[6] - a[0] * a[7]) * inv_det, (a[0] * a[4] - a[1] * a[3]) * inv_det };
}

Mat3 RotateCovariance(double angle, const Mat3& cov)
{
    const double c = std::cos(angle), s = std::sin(angle);
    const Mat3 r { c, -s, 0.0, s, c, 0.0, 0.0, 0.0, 1.0 };
    return Multiply(Multiply(r, cov), Transpose(r));
}

/* ---- DeviceGridMap ---- */
DeviceGridMap::DeviceGridMap(const DeviceContextPtr& context, std::int64_t map_id, double resolution, int log2_block_size) :
    mContext(context), mMapId(map_id), mResolution(resolution), mLog2BlockSize(log2_block_size)
{
    /* GridMap(resolution, blockSize, 1.0, 1.0) (grid_map.cpp:74-99, 226-246) */
    const int bs = 1 << mLog2BlockSize;
    const int desired_rows = static_cast<int>(std::ceil(1.0 / mResolution));
    const int desired_cols = static_cast<int>(std::ceil(1.0 / mResolution));
    mBlockRows = (desired_rows + bs - 1) >> mLog2BlockSize;
    mBlockCols = (desired_cols + bs - 1) >> mLog2BlockSize;
    mRows = mBlockRows << mLog2BlockSize;
    mCols = mBlockCols << mLog2BlockSize;
    mOffX = 0.0; mOffY = 0.0;
    mContext->Check(csm_map_create(mContext->Handle(), mMapId, mRows, mCols, mLog2BlockSize, mResolution, mOffX, mOffY),
                    "csm_map_create");
}

DeviceGridMap::~DeviceGridMap()
{
    csm_release_grid(mContext->Handle(), mMapId);
}

DeviceGridMap::Index DeviceGridMap::PositionToIndex(double x, double y) const
{
    return Index { static_cast<int>(std::floor((x - mOffX) / mResolution)),
                   static_cast<int>(std::floor((y - mOffY) / mResolution)) };
}

void DeviceGridMap::ResizeIndex(int box_min_x, int box_min_y, int box_max_x, int box_max_y)
{
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

void DeviceGridMap::Resize(double min_x, double min_y, double max_x, double max_y)
{
    /* "avoid the rounding errors": one cell of slack on every side */
    const Index idx_min = PositionToIndex(min_x - mResolution, min_y - mResolution);
    const Index idx_max = PositionToIndex(max_x + mResolution, max_y + mResolution);
    ResizeIndex(idx_min.x, idx_min.y, idx_max.x + 1, idx_max.y + 1);
}

void DeviceGridMap::Expand(double min_x, double min_y, double max_x, double max_y)
{
    const Index idx_min = PositionToIndex(min_x - mResolution, min_y - mResolution);
    const Index idx_max = PositionToIndex(max_x + mResolution, max_y + mResolution);
    const int box_min_x = idx_min.x, box_min_y = idx_min.y, box_max_x = idx_max.x + 1, box_max_y = idx_max.y + 1;
    /* Expand(BoundingBox<int>) (:914-930): nothing to do when both corners are inside */
    auto inside = [this](int row, int col) { return row >= 0 && row < mRows && col >= 0 && col < mCols; };
    if (inside(box_min_y, box_min_x) && inside(box_max_y - 1, box_max_x - 1))
        return;
    ResizeIndex(std::min(0, box_min_x), std::min(0, box_min_y), std::max(mCols, box_max_x), std::max(mRows, box_max_y));
}

void DeviceGridMap::ResetValues()
{
    mContext->Check(csm_map_reset_values(mContext->Handle(), mMapId), "csm_map_reset_values");
}

int DeviceGridMap::InsertScans(const std::vector<ScanHits>& hits, int subpixel_scale)
{
    const double scaled_res = mResolution / subpixel_scale;           /* grid_map_geometry.cpp:48-60 */
    std::size_t total = 0;
    for (const ScanHits& h : hits) total += h.x.size();
    std::vector<csm_ray> rays;
    rays.reserve(total);
    int order = 0;
    mLastExactBeams = 0;
    for (const ScanHits& h : hits) {
        const int sx = static_cast<int>(std::floor((h.sensor.x - mOffX) / scaled_res));
        const int sy = static_cast<int>(std::floor((h.sensor.y - mOffY) / scaled_res));
        for (std::size_t i = 0; i < h.x.size(); ++i) {
            double hx = h.x[i], hy = h.y[i];
            double ex = (hx - mOffX) / scaled_res, ey = (hy - mOffY) / scaled_res;
            double cx = (hx - mOffX) / mResolution, cy = (hy - mOffY) / mResolution;
            if (h.fast && (NearBoundary(ex) || NearBoundary(ey) || NearBoundary(cx) || NearBoundary(cy))) {
                h.Exact(i, hx, hy);
                ex = (hx - mOffX) / scaled_res; ey = (hy - mOffY) / scaled_res;
                cx = (hx - mOffX) / mResolution; cy = (hy - mOffY) / mResolution;
                ++mLastExactBeams;
            }
            csm_ray r;
            r.start_x = sx; r.start_y = sy;
            r.end_x = static_cast<int>(std::floor(ex));
            r.end_y = static_cast<int>(std::floor(ey));
            r.hit_col = static_cast<int>(std::floor(cx)); r.hit_row = static_cast<int>(std::floor(cy));
            r.order = order++;
            r.reserved = 0;
            rays.push_back(r);
        }
    }
    mContext->Check(csm_map_insert_rays(mContext->Handle(), mMapId, rays.data(), static_cast<int>(rays.size()), subpixel_scale),
                    "csm_map_insert_rays");
    return static_cast<int>(rays.size());
}

GridMapView DeviceGridMap::View() const
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

/* ---- GridMapBuilderGPU ---- */
GridMapBuilderGPU::GridMapBuilderGPU(const DeviceContextPtr& context, double map_resolution, int patch_size,
                                     int num_of_scans_for_latest_map, double usable_range_min,
                                     double usable_range_max, double prob_hit, double prob_miss,
                                     std::int64_t latest_map_device_id, bool reference_table_end) :
    mContext(context), mResolution(map_resolution),
    mLog2BlockSize(__builtin_ctz(ToNearestPowerOf2(patch_size))),
    mNumOfScansForLatestMap(num_of_scans_for_latest_map),
    mUsableRangeMin(usable_range_min), mUsableRangeMax(usable_range_max),
    mOddsHit(ProbabilityToOdds(prob_hit)), mOddsMiss(ProbabilityToOdds(prob_miss)),
    /* mLatestMap(mapResolution, patchSize, 1.0, 1.0) (grid_map_builder.cpp:80) */
    mLatest(context, latest_map_device_id, map_resolution, __builtin_ctz(ToNearestPowerOf2(patch_size)))
{
    const std::vector<std::uint16_t> miss = UpdateTable(mOddsMiss, reference_table_end),
                                     hit = UpdateTable(mOddsHit, reference_table_end);
    mContext->Check(csm_map_set_update_tables(mContext->Handle(), miss.data(), hit.data()), "csm_map_set_update_tables");
}

const GridMapBuilderGPU::Polar& GridMapBuilderGPU::PolarOf(const ScanDataPtr& scan_ptr)
{
    const ScanData& scan = *scan_ptr;
    auto it = mPolar.find(&scan);
    if (it != mPolar.end())
        return it->second;
    if (mPolar.size() > 256)
        mPolar.clear();                    /* the scans in use come back within a few calls */
    Polar& p = mPolar[&scan];
    p.keep = scan_ptr;
    const std::size_t n = scan.NumOfScans();
    p.px.resize(n); p.py.resize(n);
    for (std::size_t i = 0; i < n; ++i) {
        p.px[i] = scan.ranges[i] * std::cos(scan.angles[i]);
        p.py[i] = scan.ranges[i] * std::sin(scan.angles[i]);
    }
    return p;
}

DeviceGridMap::ScanHits GridMapBuilderGPU::HitsOf(const Pose2D& map_pose, const Pose2D& global_scan_pose,
                                                  const ScanDataPtr& scan_ptr)
{
    const ScanData& scan = *scan_ptr;
    DeviceGridMap::ScanHits h;
    const Pose2D global_sensor = Compound(global_scan_pose, scan.relative_sensor_pose);
    h.sensor = InverseCompound(map_pose, global_sensor);
    const double min_range = std::max(mUsableRangeMin, scan.min_range);
    const double max_range = std::min(mUsableRangeMax, scan.max_range);
    const std::size_t n = scan.NumOfScans();
    h.x.reserve(n); h.y.reserve(n);
    h.scan = &scan;
    h.fast = mFastHitPoints;
    if (h.fast) {
        /* sensor + R(heading) * (r cos a, r sin a): within a few 1e-14 m of the reference's hit point, which
         * is all a floor needs outside its guard band (DeviceGridMap::ScanHits) */
        const Polar& p = PolarOf(scan_ptr);
        const double c = std::cos(h.sensor.theta), s = std::sin(h.sensor.theta);
        h.beam.reserve(n);
        for (std::size_t i = 0; i < n; ++i) {
            const double range = scan.ranges[i];
            if (range >= max_range || range <= min_range)
                continue;
            h.x.push_back(h.sensor.x + (c * p.px[i] - s * p.py[i]));
            h.y.push_back(h.sensor.y + (s * p.px[i] + c * p.py[i]));
            h.beam.push_back(static_cast<int>(i));
        }
        return h;
    }
    for (std::size_t i = 0; i < n; ++i) {
        const double range = scan.ranges[i];
        if (range >= max_range || range <= min_range)
            continue;
        /* ScanData::HitPoint (sensor_data.hpp:190-203) */
        const double c = std::cos(h.sensor.theta + scan.angles[i]);
        const double s = std::sin(h.sensor.theta + scan.angles[i]);
        h.x.push_back(h.sensor.x + range * c);
        h.y.push_back(h.sensor.y + range * s);
    }
    return h;
}

void GridMapBuilderGPU::BoundingBox(std::vector<DeviceGridMap::ScanHits>& hits, const DeviceGridMap& map, bool construct,
                                    double& min_x, double& min_y, double& max_x, double& max_y) const
{
    for (int pass = 0; pass < 2; ++pass) {
        if (construct) {
            /* ConstructMapFromScans: the box starts at (max double, min POSITIVE double), as the reference
             * has it (:582-585) */
            min_x = min_y = std::numeric_limits<double>::max();
            max_x = max_y = std::numeric_limits<double>::min();
        } else {
            /* UpdateGridMap: the box starts at the sensor (:832-838) */
            min_x = max_x = hits[0].sensor.x;
            min_y = max_y = hits[0].sensor.y;
        }
        bool fast = false;
        for (const DeviceGridMap::ScanHits& h : hits) {
            fast = fast || h.fast;
            min_x = std::min(min_x, h.sensor.x); min_y = std::min(min_y, h.sensor.y);
            max_x = std::max(max_x, h.sensor.x); max_y = std::max(max_y, h.sensor.y);
            for (std::size_t i = 0; i < h.x.size(); ++i) {
                min_x = std::min(min_x, h.x[i]); min_y = std::min(min_y, h.y[i]);
                max_x = std::max(max_x, h.x[i]); max_y = std::max(max_y, h.y[i]);
            }
        }
        if (!fast)
            return;
        /* the floors Resize / Expand take of the box (grid_map.cpp:896-903, 936-943) */
        const double res = map.Resolution();
        const bool near = DeviceGridMap::NearBoundary((min_x - res - map.OffsetX()) / res) ||
                          DeviceGridMap::NearBoundary((min_y - res - map.OffsetY()) / res) ||
                          DeviceGridMap::NearBoundary((max_x + res - map.OffsetX()) / res) ||
                          DeviceGridMap::NearBoundary((max_y + res - map.OffsetY()) / res);
        if (!near)
            return;
        for (DeviceGridMap::ScanHits& h : hits) {
            if (!h.fast) continue;
            for (std::size_t i = 0; i < h.x.size(); ++i)
                h.Exact(i, h.x[i], h.y[i]);
            h.fast = false;
        }
    }
}

void GridMapBuilderGPU::ConstructMapFromScans(const Pose2D& map_pose, DeviceGridMap& map, const ScanNodeView* nodes, int count)
{
    /* first pass (:578-633): hit points and bounding box in the map's frame */
    std::vector<DeviceGridMap::ScanHits> hits;
    hits.reserve(count);
    for (int k = 0; k < count; ++k)
        hits.push_back(HitsOf(map_pose, nodes[k].global_pose, nodes[k].scan));
    double min_x, min_y, max_x, max_y;
    BoundingBox(hits, map, true, min_x, min_y, max_x, max_y);
    /* :636-638 */
    map.Resize(min_x, min_y, max_x, max_y);
    map.ResetValues();
    /* second pass (:642-692) */
    mLastRays = map.InsertScans(hits, SubpixelScale);
}

void GridMapBuilderGPU::UpdateLatestMap(const std::vector<ScanNodeView>& scan_nodes)
{
    if (scan_nodes.empty()) {
        std::fprintf(stderr, "csm_host: UpdateLatestMap needs at least one scan node\n");
        std::abort();
    }
    /* :506-518: the last NumOfScansForLatestMap nodes; the map's frame is the pose of the first of them */
    const int count = std::min(static_cast<int>(scan_nodes.size()), mNumOfScansForLatestMap);
    const std::size_t first = scan_nodes.size() - static_cast<std::size_t>(count);
    mLatestScanIdMin = static_cast<int>(first);
    mLatestScanIdMax = static_cast<int>(scan_nodes.size()) - 1;
    mLatestMapPose = scan_nodes[first].global_pose;
    ConstructMapFromScans(mLatestMapPose, mLatest, scan_nodes.data() + first, count);
}

void GridMapBuilderGPU::UpdateLatestMap(const std::vector<ScanNode>& scan_nodes)
{
    if (scan_nodes.empty()) {
        std::fprintf(stderr, "csm_host: UpdateLatestMap needs at least one scan node\n");
        std::abort();
    }
    const int count = std::min(static_cast<int>(scan_nodes.size()), mNumOfScansForLatestMap);
    const std::size_t first = scan_nodes.size() - static_cast<std::size_t>(count);
    std::vector<ScanNodeView> views;
    views.reserve(count);
    for (std::size_t k = first; k < scan_nodes.size(); ++k)
        views.push_back(ScanNodeView { scan_nodes[k].global_pose, scan_nodes[k].scan });
    mLatestScanIdMin = scan_nodes[first].node_id;
    mLatestScanIdMax = scan_nodes.back().node_id;
    mLatestMapPose = scan_nodes[first].global_pose;
    ConstructMapFromScans(mLatestMapPose, mLatest, views.data(), count);
}

void GridMapBuilderGPU::FinishLocalMap()
{
    if (!mLocalMaps.empty())
        mLocalMaps.back().finished = true;
}

void GridMapBuilderGPU::AppendLocalMap(PoseGraph& pose_graph, const Pose2D& scan_pose, const Mat3& covariance, int scan_node_id)
{
    FinishLocalMap();
    const int local_map_id = pose_graph.local_map_nodes.empty() ? 0 : pose_graph.local_map_nodes.back().local_map_id + 1;
    const Pose2D& local_map_pose = scan_pose;
    if (!mLocalMaps.empty()) {
        /* the inter-map odometry edge from the old local map to the new scan node (:208-240) */
        const LocalMapNode& old_node = pose_graph.local_map_nodes.back();
        PoseGraphEdge e;
        e.local_map_id = old_node.local_map_id; e.scan_node_id = scan_node_id;
        e.edge_type = EdgeType::InterLocalMap; e.constraint_type = ConstraintType::Odometry;
        e.relative_pose = NormalizeAngle(InverseCompound(old_node.global_pose, scan_pose));
        e.information = Inverse(ConvertCovarianceFromWorldToLocal(old_node.global_pose, covariance));
        pose_graph.edges.push_back(e);
    }
    pose_graph.local_map_nodes.push_back(LocalMapNode { local_map_id, local_map_pose });
    LocalMapGPU lm;
    lm.id = local_map_id;
    lm.map.reset(new DeviceGridMap(mContext, static_cast<std::int64_t>(local_map_id), mResolution, mLog2BlockSize));
    lm.scan_node_id_min = lm.scan_node_id_max = scan_node_id;
    if (!mLocalMaps.empty()) {
        /* initialised with the last scans of the previous local map (:259-278) */
        const LocalMapGPU& last = mLocalMaps.back();
        const int n = static_cast<int>(std::min<std::size_t>(pose_graph.scan_nodes.size(), mNumOfOverlappedScans));
        const int id_max = last.scan_node_id_max, id_min = id_max - (n - 1);
        std::vector<ScanNodeView> views;
        for (int id = id_min; id <= id_max; ++id)
            views.push_back(ScanNodeView { pose_graph.scan_nodes[id].global_pose, pose_graph.scan_nodes[id].scan });
        ConstructMapFromScans(local_map_pose, *lm.map, views.data(), static_cast<int>(views.size()));
    }
    mLocalMaps.push_back(std::move(lm));
    mTravelDistLastLocalMap = 0.0;
}

bool GridMapBuilderGPU::AppendScan(PoseGraph& pose_graph, const Pose2D& relative_scan_pose,
                                   const Mat3& scan_pose_covariance, const ScanDataPtr& scan)
{
    /* UpdatePoseGraph (:290-385) */
    const int scan_node_id = pose_graph.scan_nodes.empty() ? 0 : pose_graph.scan_nodes.back().node_id + 1;
    const Pose2D prev_scan_pose = pose_graph.scan_nodes.empty() ? Pose2D {} : pose_graph.scan_nodes.back().global_pose;
    const Pose2D scan_pose = Compound(prev_scan_pose, relative_scan_pose);
    mAccumTravelDist += Distance(relative_scan_pose);
    mTravelDistLastLocalMap += Distance(relative_scan_pose);
    const bool travel = mTravelDistLastLocalMap >= mTravelDistThreshold;
    const bool last_finished = !mLocalMaps.empty() && mLocalMaps.back().finished;
    const bool inserted = travel || last_finished || mLocalMaps.empty();
    if (inserted)
        AppendLocalMap(pose_graph, scan_pose, scan_pose_covariance, scan_node_id);
    const LocalMapNode& map_node = pose_graph.local_map_nodes.back();
    const Pose2D map_local_scan_pose = NormalizeAngle(InverseCompound(map_node.global_pose, scan_pose));
    ScanNode node;
    node.node_id = scan_node_id; node.local_map_id = mLocalMaps.back().id;
    node.local_pose = map_local_scan_pose; node.scan = scan; node.global_pose = scan_pose;
    pose_graph.scan_nodes.push_back(node);
    PoseGraphEdge e;
    e.local_map_id = map_node.local_map_id; e.scan_node_id = scan_node_id;
    e.edge_type = EdgeType::IntraLocalMap; e.constraint_type = ConstraintType::Odometry;
    e.relative_pose = map_local_scan_pose;
    e.information = Inverse(ConvertCovarianceFromWorldToLocal(map_node.global_pose, scan_pose_covariance));
    pose_graph.edges.push_back(e);
    UpdateGridMap(pose_graph);
    return inserted;
}

void GridMapBuilderGPU::UpdateGridMap(const PoseGraph& pose_graph)
{
    /* :390-494: the latest scan into the latest (unfinished) local map */
    LocalMapGPU& lm = mLocalMaps.back();
    const LocalMapNode& map_node = pose_graph.local_map_nodes.back();
    const ScanNode& scan_node = pose_graph.scan_nodes.back();
    std::vector<DeviceGridMap::ScanHits> hits(1);
    hits[0] = HitsOf(map_node.global_pose, scan_node.global_pose, scan_node.scan);
    double min_x, min_y, max_x, max_y;
    BoundingBox(hits, *lm.map, false, min_x, min_y, max_x, max_y);
    lm.map->Expand(min_x, min_y, max_x, max_y);
    mLastRays = lm.map->InsertScans(hits, SubpixelScale);
    lm.scan_node_id_max = scan_node.node_id;
}

void GridMapBuilderGPU::AfterLoopClosure(const PoseGraph& pose_graph)
{
    /* UpdateAccumTravelDist (:535-558) */
    mAccumTravelDist = 0.0;
    for (std::size_t i = 0; i + 1 < pose_graph.scan_nodes.size(); ++i)
        mAccumTravelDist += Distance(pose_graph.scan_nodes[i].global_pose, pose_graph.scan_nodes[i + 1].global_pose);
}

} /* namespace csm_host */
