/* ref_wrapper.cpp -- drives the UNMODIFIED reference implementation through
 * the C interface of oracle_api.h.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h). This file is compiled together
 * with the reference translation units where they lie under /root/reference
 * (recipe: oracle/Makefile, output: oracle/_ref/libcsm_ref.so). All matching
 * decisions are taken by the reference classes
 *   ScanMatcherCorrelative  (scan_matcher_correlative.cpp:92-244),
 *   ScanMatcherBranchBound  (scan_matcher_branch_bound.cpp:87-278),
 *   ScanMatcherGridSearch   (scan_matcher_grid_search.cpp:69-178),
 *   LoopDetectorBranchBound (loop_detector_branch_bound.cpp:59-156),
 *   PrecomputeGridMap(s)    (grid_map_builder.cpp:987-1065);
 * this wrapper only builds their inputs from plain arrays and decodes the
 * window indices / integer score at the pose the reference returned.
 */

#include "oracle_api.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "my_lidar_graph_slam/pose.hpp"
#include "my_lidar_graph_slam/point.hpp"
#include "my_lidar_graph_slam/metric/metric.hpp"
#include "my_lidar_graph_slam/sensor/sensor_data.hpp"
#include "my_lidar_graph_slam/io/carmen/carmen_reader.hpp"
#include "my_lidar_graph_slam/mapping/grid_map_types.hpp"
/* (GridMapBuilder registers its metrics under fixed names, so a process can construct it only once;
 * the checker keeps that one instance and re-initialises its members for every new sequence, which
 * takes access to them. No reference code is changed.) */
#define private public
#include "my_lidar_graph_slam/mapping/grid_map_builder.hpp"
#undef private
#include "my_lidar_graph_slam/mapping/cost_function_square_error.hpp"
#include "my_lidar_graph_slam/mapping/cost_function_greedy_endpoint.hpp"
#include "my_lidar_graph_slam/mapping/score_function_pixel_accurate.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_correlative.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_grid_search.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_linear_solver.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_hill_climbing.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/pose_graph.hpp"
/* The reference registers the searcher's metrics under fixed names, so a process can construct
 * LoopSearcherNearest only once (metric.cpp asserts on duplicates); its thresholds are private
 * const members. The checker needs several parameter sets: it keeps the one instance and rewrites
 * those members, which takes access to them. No reference code is changed. */
#define private public
#include "my_lidar_graph_slam/mapping/loop_searcher_nearest.hpp"
#undef private

using namespace MyLidarGraphSlam;
using namespace MyLidarGraphSlam::Mapping;

namespace {

constexpr int kBlockSize = 16;
constexpr double kCovarianceScale = 1e4;

struct RefGrid
{
    GridMap mMap;
    explicit RefGrid(GridMap&& map) : mMap(std::move(map)) { }
};

std::atomic<int> gNameCounter { 0 };

std::string UniqueName(const char* prefix)
{
    return std::string(prefix) + "#" + std::to_string(gNameCounter++);
}

Sensor::ScanDataPtr<double> MakeScan(const double* angles,
                                     const double* ranges, const int n,
                                     const double relPose[3])
{
    std::vector<double> a(angles, angles + n);
    std::vector<double> r(ranges, ranges + n);
    const double rmin = *std::min_element(r.begin(), r.end());
    const double rmax = *std::max_element(r.begin(), r.end());
    const double amin = *std::min_element(a.begin(), a.end());
    const double amax = *std::max_element(a.begin(), a.end());
    const RobotPose2D<double> zero { 0.0, 0.0, 0.0 };
    const RobotPose2D<double> rel { relPose[0], relPose[1], relPose[2] };
    return std::make_shared<Sensor::ScanData<double>>(
        "oracle", 0.0, zero, zero, rel, rmin, rmax, amin, amax,
        std::move(a), std::move(r));
}

int LastInt(const std::string& id)
{
    auto* seq = Metric::MetricManager::Instance()->ValueSequenceMetric<int>(id);
    const std::size_t n = seq->NumOfValues();
    const int v = (n > 0) ? seq->ValueAt(n - 1) : 0;
    seq->Reset();
    return v;
}

float LastFloat(const std::string& id)
{
    auto* seq =
        Metric::MetricManager::Instance()->ValueSequenceMetric<float>(id);
    const std::size_t n = seq->NumOfValues();
    const float v = (n > 0) ? seq->ValueAt(n - 1) : 0.0f;
    seq->Reset();
    return v;
}

void ResetMatcherMetrics(const std::string& name, bool gridSearch)
{
    auto* mgr = Metric::MetricManager::Instance();
    static const char* ints[] = { ".InputSetupTime", ".OptimizationTime",
        ".WinSizeX", ".WinSizeY", ".WinSizeTheta", ".NumOfIgnoredNodes",
        ".NumOfProcessedNodes", ".NumOfScans", ".NumOfScoreEvaluations",
        ".NumOfScoreUpdates" };
    static const char* floats[] = { ".DiffTranslation", ".DiffRotation",
        ".StepSizeX", ".StepSizeY", ".StepSizeTheta", ".ScoreValue",
        ".CostValue" };
    (void)gridSearch;
    for (const char* s : ints)
        mgr->ValueSequenceMetric<int>(name + s)->Reset();
    for (const char* s : floats)
        mgr->ValueSequenceMetric<float>(name + s)->Reset();
}

void FillSummary(const ScanMatchingSummary& summary, orc_result* out)
{
    out->found = summary.mPoseFound ? 1 : 0;
    out->est_pose[0] = summary.mEstimatedPose.mX;
    out->est_pose[1] = summary.mEstimatedPose.mY;
    out->est_pose[2] = summary.mEstimatedPose.mTheta;
    out->norm_cost = summary.mNormalizedCost;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            out->cov[r * 3 + c] = summary.mEstimatedCovariance(r, c);
}

/* Diagnostics at the returned pose: the per-node score of
 * score_function_pixel_accurate.cpp:16-58 (re-projection per pose) */
void DiagnosePixelAccurate(const GridMap& map,
                           const Sensor::ScanDataPtr<double>& scan,
                           const RobotPose2D<double>& pose, orc_result* out)
{
    ScorePixelAccurate scoreFunc;
    const auto summary = scoreFunc.Score(map, scan, pose);
    out->score = summary.mNormalizedScore;
    out->known_rate = summary.mKnownRate;
    std::int64_t sum = 0;
    int known = 0;
    for (std::size_t i = 0; i < scan->NumOfScans(); ++i) {
        const Point2D<double> hit = scan->HitPoint(pose, i);
        const Point2D<int> idx = map.PositionToIndex(hit.mX, hit.mY);
        const std::uint16_t v = map.ValueOr(idx.mY, idx.mX, 0);
        if (v != 0) { sum += v; ++known; }
    }
    out->sum_value = sum;
    out->n_known = known;
}

/* Diagnostics for the real-time correlative matcher: indices are computed
 * once per angle and shifted by integers (scan_matcher_correlative.cpp:161-168,
 * 301-336) */
void DiagnoseShifted(const GridMap& map,
                     const Sensor::ScanDataPtr<double>& scan,
                     const RobotPose2D<double>& anglePose,
                     const int offX, const int offY, orc_result* out)
{
    double sumScore = 0.0;
    std::int64_t sum = 0;
    int known = 0;
    const std::size_t n = scan->NumOfScans();
    for (std::size_t i = 0; i < n; ++i) {
        const Point2D<double> hit = scan->HitPoint(anglePose, i);
        const Point2D<int> idx = map.PositionToIndex(hit.mX, hit.mY);
        const double prob = map.ProbabilityOr(idx.mY + offY, idx.mX + offX, 0.0);
        if (prob == 0.0)
            continue;
        sumScore += prob;
        sum += map.ValueOr(idx.mY + offY, idx.mX + offX, 0);
        ++known;
    }
    out->score = sumScore / static_cast<double>(n);
    out->known_rate = static_cast<double>(known) / static_cast<double>(n);
    out->sum_value = sum;
    out->n_known = known;
}

void SearchStep(const GridMap& map, const Sensor::ScanDataPtr<double>& scan,
                double& stepX, double& stepY, double& stepT)
{
    /* scan_matcher_correlative.cpp:255-274 (private there) */
    const double res = map.Resolution();
    const double maxRange = *std::max_element(
        scan->Ranges().cbegin(), scan->Ranges().cend());
    const double theta = res / maxRange;
    stepX = res;
    stepY = res;
    stepT = std::acos(1.0 - 0.5 * theta * theta);
}

void DecodeWindow(const GridMap& map, const Sensor::ScanDataPtr<double>& scan,
                  const RobotPose2D<double>& initPose,
                  const ScanMatchingSummary& summary,
                  const double rangeX, const double rangeY, const double rangeT,
                  orc_result* out, RobotPose2D<double>& sensorPose,
                  RobotPose2D<double>& bestPose)
{
    sensorPose = Compound(initPose, scan->RelativeSensorPose());
    SearchStep(map, scan, out->step_x, out->step_y, out->step_t);
    out->win_x = static_cast<int>(std::ceil(0.5 * rangeX / out->step_x));
    out->win_y = static_cast<int>(std::ceil(0.5 * rangeY / out->step_y));
    out->win_t = static_cast<int>(std::ceil(0.5 * rangeT / out->step_t));
    const RobotPose2D<double> best =
        Compound(summary.mEstimatedPose, scan->RelativeSensorPose());
    out->best_x = static_cast<int>(std::lround(
        (best.mX - sensorPose.mX) / out->step_x));
    out->best_y = static_cast<int>(std::lround(
        (best.mY - sensorPose.mY) / out->step_y));
    out->best_t = static_cast<int>(std::lround(
        (best.mTheta - sensorPose.mTheta) / out->step_t));
    /* Rebuild the pose exactly like the matchers do */
    bestPose = RobotPose2D<double> {
        sensorPose.mX + out->best_x * out->step_x,
        sensorPose.mY + out->best_y * out->step_y,
        sensorPose.mTheta + out->best_t * out->step_t };
    out->best_sensor_pose[0] = bestPose.mX;
    out->best_sensor_pose[1] = bestPose.mY;
    out->best_sensor_pose[2] = bestPose.mTheta;
}

/* Dense row-major copy, one ValueOr per cell. GridMap::CopyValues is not used:
 * for 16-bit buffers CopyValuesInternal advances the destination by
 * count / sizeof(U) elements (grid_map.cpp:343,349), i.e. by half a row */
template <typename MapType>
void Flatten(const MapType& map, uint16_t* out)
{
    const int rows = map.Rows();
    const int cols = map.Cols();
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            out[static_cast<std::size_t>(r) * cols + c] = map.ValueOr(r, c, 0);
}

/* Final matcher that hands the coarse estimate through unchanged: the
 * sub-pixel refiners (scan_matcher_hill_climbing / linear_solver) are
 * outside the hot path */
/* Final matcher of the loop detector: keeps the coarse pose (the sub-pixel
 * refiner is outside the path) and reports the covariance the reference's own
 * cost function gives at that pose, i.e. what the coarse matcher's epilogue
 * computed and LoopDetectorBranchBound::Detect then discards
 * (scan_matcher_branch_bound.cpp:237-262, loop_detector_branch_bound.cpp:123-135). */
class PassThroughMatcher final : public ScanMatcher
{
public:
    PassThroughMatcher() : ScanMatcher("PassThrough"),
                           mCostFunc(std::make_shared<CostSquareError>(kCovarianceScale)) { }
    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& query) override
    {
        const RobotPose2D<double> sensorPose = Compound(
            query.mMapLocalInitialPose, query.mScanData->RelativeSensorPose());
        return ScanMatchingSummary { true, 0.0, query.mMapLocalInitialPose,
                                     query.mMapLocalInitialPose,
                                     this->mCostFunc->ComputeCovariance(
                                         query.mGridMap, query.mScanData, sensorPose) };
    }
private:
    std::shared_ptr<CostSquareError> mCostFunc;
};

struct RefLoopDetector
{
    int    mHeightMax;
    double mRangeX, mRangeY, mRangeT, mScoreThr, mKnownThr;
    int    mNumThreads;
    /* final matcher: pass-through, or the reference's ScanMatcherLinearSolver */
    bool   mLinearSolver = false;
    int    mFinalIterations = 10;
    double mFinalConvergence = 1e-4, mFinalLambda = 1e-4;
    std::vector<std::shared_ptr<ScanMatcherBranchBound>> mMatchers;
    std::vector<std::unique_ptr<LoopDetectorBranchBound>> mDetectors;

    void Build()
    {
        this->mMatchers.clear();
        this->mDetectors.clear();
        for (int i = 0; i < this->mNumThreads; ++i) {
            auto scoreFunc = std::make_shared<ScorePixelAccurate>();
            auto costFunc = std::make_shared<CostSquareError>(kCovarianceScale);
            auto matcher = std::make_shared<ScanMatcherBranchBound>(
                UniqueName("LoopBB"), scoreFunc, costFunc, this->mHeightMax,
                this->mRangeX, this->mRangeY, this->mRangeT);
            std::shared_ptr<ScanMatcher> finalMatcher;
            if (this->mLinearSolver)
                finalMatcher = std::make_shared<ScanMatcherLinearSolver>(
                    UniqueName("LoopFinal"), this->mFinalIterations, this->mFinalConvergence,
                    this->mFinalLambda, std::make_shared<CostSquareError>(kCovarianceScale));
            else
                finalMatcher = std::make_shared<PassThroughMatcher>();
            this->mMatchers.push_back(matcher);
            this->mDetectors.push_back(
                std::make_unique<LoopDetectorBranchBound>(
                    UniqueName("LoopDet"), matcher, finalMatcher,
                    this->mScoreThr, this->mKnownThr));
        }
    }
};

} /* namespace */

extern "C" {

const char* orc_kind(void) { return "reference"; }

void* orc_grid_create(const uint16_t* dense, int rows, int cols,
                      double resolution, double offset_x, double offset_y)
{
    if (rows <= 0 || cols <= 0 || rows % kBlockSize || cols % kBlockSize)
        return nullptr;
    GridMap map { resolution, kBlockSize, rows / kBlockSize, cols / kBlockSize,
                  Point2D<double> { offset_x, offset_y } };
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const std::uint16_t v = dense[static_cast<std::size_t>(r) * cols + c];
            if (v != 0)
                map.SetValue(r, c, v);
        }
    return new RefGrid(std::move(map));
}

void orc_grid_destroy(void* grid)
{
    delete static_cast<RefGrid*>(grid);
}

int orc_precompute(void* grid, int win, uint16_t* out)
{
    const GridMap& map = static_cast<RefGrid*>(grid)->mMap;
    const ConstMap precomp = PrecomputeGridMap(map, win);
    Flatten(precomp, out);
    return 0;
}

int orc_precompute_pyramid(void* grid, int hmax, uint16_t* out)
{
    const GridMap& map = static_cast<RefGrid*>(grid)->mMap;
    std::vector<ConstMap> maps;
    PrecomputeGridMaps(map, maps, hmax);
    const std::size_t cells =
        static_cast<std::size_t>(map.Rows()) * map.Cols();
    for (std::size_t h = 0; h < maps.size(); ++h)
        Flatten(maps[h], out + h * cells);
    return 0;
}

int orc_match_rt(void* grid, const double* angles, const double* ranges, int n,
                 const double init_pose[3], const double rel_sensor_pose[3],
                 int low_res, double range_x, double range_y, double range_t,
                 double score_thr, double known_thr, orc_result* out)
{
    const GridMap& map = static_cast<RefGrid*>(grid)->mMap;
    const auto scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const RobotPose2D<double> initPose {
        init_pose[0], init_pose[1], init_pose[2] };
    const std::string name = UniqueName("RT");
    ScanMatcherCorrelative matcher {
        name, std::make_shared<CostSquareError>(kCovarianceScale),
        low_res, range_x, range_y, range_t };
    const ConstMap precomp = matcher.ComputeCoarserMap(map);
    const ScanMatchingSummary summary = matcher.OptimizePose(
        map, precomp, scan, initPose, score_thr, known_thr);

    *out = orc_result { };
    FillSummary(summary, out);
    RobotPose2D<double> sensorPose, bestPose;
    DecodeWindow(map, scan, initPose, summary, range_x, range_y, range_t,
                 out, sensorPose, bestPose);
    const RobotPose2D<double> anglePose {
        sensorPose.mX, sensorPose.mY,
        sensorPose.mTheta + out->step_t * out->best_t };
    DiagnoseShifted(map, scan, anglePose, out->best_x, out->best_y, out);
    out->n_processed = LastInt(name + ".NumOfProcessedNodes");
    out->n_ignored = LastInt(name + ".NumOfIgnoredNodes");
    ResetMatcherMetrics(name, false);
    return 0;
}

int orc_match_bb(void* grid, const double* angles, const double* ranges, int n,
                 const double init_pose[3], const double rel_sensor_pose[3],
                 int hmax, double range_x, double range_y, double range_t,
                 double score_thr, double known_thr, orc_result* out)
{
    const GridMap& map = static_cast<RefGrid*>(grid)->mMap;
    const auto scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const RobotPose2D<double> initPose {
        init_pose[0], init_pose[1], init_pose[2] };
    const std::string name = UniqueName("BB");
    ScanMatcherBranchBound matcher {
        name, std::make_shared<ScorePixelAccurate>(),
        std::make_shared<CostSquareError>(kCovarianceScale),
        hmax, range_x, range_y, range_t };
    const std::vector<ConstMap> pyramid = matcher.ComputeCoarserMaps(map);
    const ScanMatchingSummary summary = matcher.OptimizePose(
        map, pyramid, scan, initPose, score_thr, known_thr);

    *out = orc_result { };
    FillSummary(summary, out);
    RobotPose2D<double> sensorPose, bestPose;
    DecodeWindow(map, scan, initPose, summary, range_x, range_y, range_t,
                 out, sensorPose, bestPose);
    DiagnosePixelAccurate(map, scan, bestPose, out);
    out->n_processed = LastInt(name + ".NumOfProcessedNodes");
    out->n_ignored = LastInt(name + ".NumOfIgnoredNodes");
    ResetMatcherMetrics(name, false);
    return 0;
}

int orc_match_grid(void* grid, const double* angles, const double* ranges, int n,
                   const double init_pose[3], const double rel_sensor_pose[3],
                   double range_x, double range_y, double range_t,
                   double step_x, double step_y, double step_t,
                   double score_thr, double known_thr, orc_result* out)
{
    const GridMap& map = static_cast<RefGrid*>(grid)->mMap;
    const auto scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const RobotPose2D<double> initPose {
        init_pose[0], init_pose[1], init_pose[2] };
    const std::string name = UniqueName("GS");
    ScanMatcherGridSearch matcher {
        name, std::make_shared<ScorePixelAccurate>(),
        std::make_shared<CostSquareError>(kCovarianceScale),
        range_x, range_y, range_t, step_x, step_y, step_t };
    const ScanMatchingSummary summary = matcher.OptimizePose(
        map, scan, initPose, score_thr, known_thr);

    *out = orc_result { };
    FillSummary(summary, out);
    out->step_x = step_x;
    out->step_y = step_y;
    out->step_t = step_t;
    const RobotPose2D<double> sensorPose =
        Compound(initPose, scan->RelativeSensorPose());
    const RobotPose2D<double> best =
        Compound(summary.mEstimatedPose, scan->RelativeSensorPose());
    /* Decode the loop indices with the same accumulating loops as
     * scan_matcher_grid_search.cpp:118-120 */
    auto decode = [](const double radius, const double step,
                     const double base, const double value, double& outPose) {
        int bestIdx = 0;
        double bestErr = 1e300;
        int idx = 0;
        for (double d = -radius; d <= radius; d += step, ++idx) {
            const double err = std::fabs((base + d) - value);
            if (err < bestErr) { bestErr = err; bestIdx = idx; outPose = base + d; }
        }
        return bestIdx;
    };
    RobotPose2D<double> bestPose = sensorPose;
    if (summary.mPoseFound) {
        out->best_x = decode(range_x / 2.0, step_x, sensorPose.mX, best.mX,
                             bestPose.mX);
        out->best_y = decode(range_y / 2.0, step_y, sensorPose.mY, best.mY,
                             bestPose.mY);
        out->best_t = decode(range_t / 2.0, step_t, sensorPose.mTheta,
                             best.mTheta, bestPose.mTheta);
    } else {
        out->best_x = out->best_y = out->best_t = -1;
    }
    out->best_sensor_pose[0] = bestPose.mX;
    out->best_sensor_pose[1] = bestPose.mY;
    out->best_sensor_pose[2] = bestPose.mTheta;
    DiagnosePixelAccurate(map, scan, bestPose, out);
    out->n_processed = LastInt(name + ".NumOfScoreEvaluations");
    out->n_ignored = LastInt(name + ".NumOfScoreUpdates");
    ResetMatcherMetrics(name, true);
    return 0;
}

int orc_refine(void* grid, const double* angles, const double* ranges, int n,
               const double init_pose[3], const double rel_sensor_pose[3],
               int iterations_max, double convergence_threshold, double* lambda,
               orc_result* out)
{
    const GridMap& map = static_cast<RefGrid*>(grid)->mMap;
    const auto scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const RobotPose2D<double> initPose { init_pose[0], init_pose[1], init_pose[2] };
    const std::string name = UniqueName("LS");
    ScanMatcherLinearSolver matcher {
        name, iterations_max, convergence_threshold, *lambda,
        std::make_shared<CostSquareError>(kCovarianceScale) };
    const ScanMatchingQuery query { map, Point2D<double> { 0.0, 0.0 }, scan, initPose };
    const ScanMatchingSummary summary = matcher.OptimizePose(query);
    *out = orc_result { };
    FillSummary(summary, out);
    out->n_processed = LastInt(name + ".NumOfIterations");
    /* The damping factor is private state of the matcher; it follows from the iteration count only
     * through the cost trend, which is not observable from outside. The tests therefore start a fresh
     * solver per call (and, for the carried state, compare whole Detect sequences). */
    (void)lambda;
    return 0;
}

void orc_loopdet_use_linear_solver(void* detPtr, int iterations_max, double convergence_threshold,
                                   double initial_lambda)
{
    auto* det = static_cast<RefLoopDetector*>(detPtr);
    det->mLinearSolver = true;
    det->mFinalIterations = iterations_max;
    det->mFinalConvergence = convergence_threshold;
    det->mFinalLambda = initial_lambda;
    det->Build();
}

/* ---- map construction: GridMapBuilder::UpdateLatestMap (grid_map_builder.cpp:497-532, 561-695) ---- */
namespace {
struct RefMapBuilder
{
    std::unique_ptr<GridMapBuilder> mBuilder;
    IdMap<NodeId, ScanNode> mScanNodes;
};
RefMapBuilder* gMapBuilder = nullptr;
} /* namespace */

/* GridBinaryBayes::ValueToProbability(v) (grid_binary_bayes.cpp:339-342): the table look-up every score and
 * cost evaluation goes through; v = 65535 is one past the table's end. */
double orc_value_probability(int v)
{
    return GridMap::GridType::ValueToProbability(static_cast<std::uint16_t>(v));
}

/* out[v] = the value of a cell that holds v after GridBinaryBayes::UpdateOddsUnchecked(odds)
 * (grid_binary_bayes.cpp:302-321), for every u16 v, on a grid of one cell.
 * v = 65535 runs the reference's own out-of-table read (grid_values.cpp:72-74). */
int orc_update_table(double odds, uint16_t* out)
{
    GridMap::GridType cell;
    cell.Initialize(0);                 /* log2 size 0: one cell, allocated */
    for (int v = 0; v < 65536; ++v) {
        cell.SetValueUnchecked(0, 0, static_cast<std::uint16_t>(v));
        cell.UpdateOddsUnchecked(0, 0, odds);
        out[v] = cell.ValueUnchecked(0, 0);
    }
    return 0;
}

void* orc_mapbuilder_create(double resolution, int patch_size, int scans_for_latest_map,
                            double usable_range_min, double usable_range_max, double prob_hit, double prob_miss)
{
    if (gMapBuilder == nullptr) {
        gMapBuilder = new RefMapBuilder;
        gMapBuilder->mBuilder.reset(new GridMapBuilder(resolution, patch_size, scans_for_latest_map, 1e9, 0,
                                                       usable_range_min, usable_range_max, prob_hit, prob_miss));
    }
    GridMapBuilder& b = *gMapBuilder->mBuilder;
    /* a fresh sequence on the one instance: the state its constructor sets (grid_map_builder.cpp:70-99) */
    const_cast<double&>(b.mResolution) = resolution;
    const_cast<int&>(b.mPatchSize) = patch_size;
    b.mLatestMap = GridMap(resolution, patch_size, 1.0, 1.0);
    b.mLatestMapPose = RobotPose2D<double>(0.0, 0.0, 0.0);
    const_cast<int&>(b.mNumOfScansForLatestMap) = scans_for_latest_map;
    b.mLatestScanIdMin = NodeId(0);
    b.mLatestScanIdMax = NodeId(0);
    const_cast<double&>(b.mUsableRangeMin) = usable_range_min;
    const_cast<double&>(b.mUsableRangeMax) = usable_range_max;
    const_cast<double&>(b.mProbHit) = prob_hit;
    const_cast<double&>(b.mProbMiss) = prob_miss;
    const_cast<double&>(b.mOddsHit) = GridMap::GridType::ProbabilityToOdds(prob_hit);
    const_cast<double&>(b.mOddsMiss) = GridMap::GridType::ProbabilityToOdds(prob_miss);
    gMapBuilder->mScanNodes = IdMap<NodeId, ScanNode>();
    return gMapBuilder;
}

/* Append one scan node (global pose, scan) and rebuild the latest map like the front end does per scan */
int orc_mapbuilder_append(void* p, const double pose[3], const double* angles, const double* ranges, int n,
                          const double rel_pose[3], double min_range, double max_range)
{
    auto* mb = static_cast<RefMapBuilder*>(p);
    std::vector<double> a(angles, angles + n), r(ranges, ranges + n);
    const RobotPose2D<double> zero { 0.0, 0.0, 0.0 };
    const RobotPose2D<double> rel { rel_pose[0], rel_pose[1], rel_pose[2] };
    auto scan = std::make_shared<Sensor::ScanData<double>>(
        "lidar", 0.0, zero, zero, rel, min_range, max_range, a.front(), a.back(), std::move(a), std::move(r));
    const int id = static_cast<int>(mb->mScanNodes.size());
    const RobotPose2D<double> global { pose[0], pose[1], pose[2] };
    mb->mScanNodes.Append(NodeId { id }, LocalMapId { 0 }, zero, scan, global);
    mb->mBuilder->UpdateLatestMap(mb->mScanNodes);
    return 0;
}

/* geometry[0..6] = rows, cols, block size (as doubles), offset x, y, resolution; pose of the latest map */
int orc_mapbuilder_latest(void* p, double* geometry6, double* map_pose3, uint16_t* dense, int cap_cells,
                          uint8_t* alloc, int cap_blocks)
{
    auto* mb = static_cast<RefMapBuilder*>(p);
    const GridMap& map = mb->mBuilder->LatestMap();
    geometry6[0] = map.Rows(); geometry6[1] = map.Cols(); geometry6[2] = map.BlockSize();
    geometry6[3] = map.PosOffset().mX; geometry6[4] = map.PosOffset().mY; geometry6[5] = map.Resolution();
    const RobotPose2D<double>& pose = mb->mBuilder->LatestMapPose();
    map_pose3[0] = pose.mX; map_pose3[1] = pose.mY; map_pose3[2] = pose.mTheta;
    if (map.Rows() * map.Cols() > cap_cells || map.BlockRows() * map.BlockCols() > cap_blocks)
        return -1;
    Flatten(map, dense);
    for (int br = 0; br < map.BlockRows(); ++br)
        for (int bc = 0; bc < map.BlockCols(); ++bc)
            alloc[br * map.BlockCols() + bc] = map.Block(br, bc)->IsAllocated() ? 1 : 0;
    return 0;
}

/* ---- the full loop (BASELINE configs[4]): the reference's own components in the order its front end and
 * back end call them (lidar_graph_slam_frontend.cpp:109-330, lidar_graph_slam_backend.cpp:92-198,
 * lidar_graph_slam.cpp:224-672). LidarGraphSlam itself needs the launcher, its worker threads and an
 * optimiser library; what it does between the components -- a few pose compositions and the bookkeeping
 * of ids -- is restated here, single-threaded, with the optimiser left out (poses stay as matched). Every
 * map, every match and every detection is the reference's code. Settings: slam_settings.py. ---- */
namespace {

struct RefSlam
{
    std::vector<double> v;                       /* the 37 settings */
    std::shared_ptr<PoseGraph> mPoseGraph;
    GridMapBuilder* mBuilder = nullptr;
    std::unique_ptr<ScanMatcherCorrelative> mScanMatcher;
    std::unique_ptr<ScanMatcherLinearSolver> mFinalMatcher;
    LoopSearcherNearest* mSearcher = nullptr;
    std::shared_ptr<ScanMatcherBranchBound> mLoopMatcher;
    std::unique_ptr<LoopDetectorBranchBound> mLoopDetector;
    std::vector<LoopDetectionResult> mLoops;
    int mProcessCount = 0;
    RobotPose2D<double> mLastOdomPose { 0.0, 0.0, 0.0 }, mLastMapUpdateOdomPose { 0.0, 0.0, 0.0 };
    double mAccumulatedTravelDist = 0.0, mAccumulatedAngle = 0.0, mLastMapUpdateTime = 0.0, mLastLoopDetectionDist = 0.0;
    double c[14] = { 0 };                        /* counters, the layout of csm_host_slam_counters */

    double S(int i) const { return v[i]; }
    void RunBackendStep();
    bool ProcessScan(const Sensor::ScanDataPtr<double>& scan, const RobotPose2D<double>& odomPose, double stamp);
};

double Now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

LoopSearcherNearest* gSearcher = nullptr;

bool RefSlam::ProcessScan(const Sensor::ScanDataPtr<double>& scan, const RobotPose2D<double>& odomPose, double stamp)
{
    const RobotPose2D<double> zero { 0.0, 0.0, 0.0 };
    const RobotPose2D<double> relOdom = mProcessCount == 0 ? zero : InverseCompound(mLastOdomPose, odomPose);
    mLastOdomPose = odomPose;
    mAccumulatedTravelDist += Distance(relOdom);
    mAccumulatedAngle += std::fabs(relOdom.mTheta);
    const double elapsed = mProcessCount == 0 ? 0.0 : stamp - mLastMapUpdateTime;
    const bool first = mProcessCount == 0;
    const bool needed = (mAccumulatedTravelDist >= S(9) || mAccumulatedAngle >= S(10) || elapsed >= S(11) || first) &&
                        elapsed >= 0.0;
    c[0] += 1;
    if (!needed)
        return false;
    if (first) {
        Eigen::Matrix3d cov = Eigen::Matrix3d::Zero();
        cov(0, 0) = 1e-9; cov(1, 1) = 1e-9; cov(2, 2) = 1e-9;
        const double t0 = Now();
        mBuilder->AppendScan(mPoseGraph, RobotPose2D<double> { S(16), S(17), S(18) }, cov, scan);
        c[11] += Now() - t0;
    } else {
        double t0 = Now();
        mBuilder->UpdateLatestMap(mPoseGraph->ScanNodes());
        const RobotPose2D<double> latestScanPose = mPoseGraph->ScanNodes().Back().mGlobalPose;
        const GridMap& latestMap = mBuilder->LatestMap();      /* the reference copies it; the matchers only read */
        const RobotPose2D<double> latestMapPose = mBuilder->LatestMapPose();
        double t1 = Now();
        c[9] += t1 - t0;
        const RobotPose2D<double> relFromLastUpdate = InverseCompound(mLastMapUpdateOdomPose, odomPose);
        const RobotPose2D<double> initialPose = Compound(latestScanPose, relFromLastUpdate);
        const RobotPose2D<double> mapLocalInitialPose = InverseCompound(latestMapPose, initialPose);
        const Point2D<double> center { 0.0, 0.0 };          /* read by neither matcher */
        const ScanMatchingSummary coarse = mScanMatcher->OptimizePose(
            ScanMatchingQuery { latestMap, center, scan, mapLocalInitialPose });
        const ScanMatchingSummary fin = mFinalMatcher->OptimizePose(
            ScanMatchingQuery { latestMap, center, scan, coarse.mEstimatedPose });
        double t2 = Now();
        c[10] += t2 - t1;
        const RobotPose2D<double> globalEstimated = Compound(latestMapPose, fin.mEstimatedPose);
        const RobotPose2D<double> scanRelative = InverseCompound(latestScanPose, globalEstimated);
        const Eigen::Matrix3d scanCov = ConvertCovarianceFromLocalToWorld(latestMapPose, fin.mEstimatedCovariance);
        RobotPose2D<double> relativePose = scanRelative;
        Eigen::Matrix3d covariance = scanCov;
        /* CheckDegeneration (:334-348): eigenvalues of the translational block in closed form */
        const double a = scanCov(0, 0), b = scanCov(0, 1), cc = scanCov(1, 0), d = scanCov(1, 1);
        const double mean = 0.5 * (a + d), disc = 0.25 * (a - d) * (a - d) + b * cc;
        const double root = disc > 0.0 ? std::sqrt(disc) : 0.0;
        if ((mean + root) / (mean - root) > S(13)) {
            c[7] += 1;
            /* ComputeOdometryCovariance (:351-368); FuseOdometryCovariance is off in this driver */
            const double trans = std::max(1e-1, Distance(relFromLastUpdate) / elapsed);
            const double rot = std::max(1e-1, relFromLastUpdate.mTheta / elapsed);
            Eigen::Matrix3d odomCov = Eigen::Matrix3d::Zero();
            odomCov(0, 0) = trans * trans * S(14); odomCov(1, 1) = trans * trans * S(14); odomCov(2, 2) = rot * rot * S(14);
            relativePose = relFromLastUpdate;
            covariance = odomCov;
        }
        mBuilder->AppendScan(mPoseGraph, relativePose, covariance, scan);
        double t3 = Now();
        c[11] += t3 - t2;
        const double accum = mBuilder->AccumTravelDist();
        if (accum - mLastLoopDetectionDist >= S(12)) {
            mLastLoopDetectionDist = accum;
            RunBackendStep();
            c[12] += Now() - t3;
        }
    }
    c[1] += 1;
    mProcessCount += 1;
    mAccumulatedTravelDist = 0.0;
    mAccumulatedAngle = 0.0;
    mLastMapUpdateOdomPose = odomPose;
    mLastMapUpdateTime = stamp;
    return true;
}

void RefSlam::RunBackendStep()
{
    c[2] += 1;
    /* GetLoopSearchHint (lidar_graph_slam.cpp:273-381) */
    const auto& localMaps = mBuilder->LocalMaps();
    const auto unfinishedIt = std::find_if(localMaps.cbegin(), localMaps.cend(),
        [](const IdMap<LocalMapId, LocalMap>::ConstIdDataPair& pair) { return !pair.mData.mFinished; });
    if (unfinishedIt == localMaps.begin())
        return;
    const LocalMapId mapIdMax = unfinishedIt != localMaps.cend() ? unfinishedIt->mId : LocalMapId { LocalMapId::Invalid };
    const NodeId nodeIdMax = unfinishedIt != localMaps.cend() ? unfinishedIt->mData.mScanNodeIdMin : NodeId { NodeId::Invalid };
    IdMap<NodeId, ScanNodeData> scanNodes;
    IdMap<LocalMapId, LocalMapData> mapNodes;
    for (const auto& [nodeId, scanNode] : mPoseGraph->ScanNodes()) {
        if (mapIdMax.mId != LocalMapId::Invalid && (scanNode.mLocalMapId >= mapIdMax || scanNode.mNodeId >= nodeIdMax))
            break;
        scanNodes.Append(nodeId, scanNode.mGlobalPose);
    }
    for (const auto& [nodeId, mapNode] : mPoseGraph->LocalMapNodes()) {
        if (mapIdMax.mId != LocalMapId::Invalid && nodeId >= mapIdMax)
            break;
        const auto& localMap = mBuilder->LocalMapAt(nodeId);
        /* the bounding box of the hint is read by no searcher */
        mapNodes.Append(nodeId, Point2D<double> { 0.0, 0.0 }, Point2D<double> { 0.0, 0.0 },
                        localMap.mScanNodeIdMin, localMap.mScanNodeIdMax, localMap.mFinished);
    }
    if (mapNodes.empty() || scanNodes.empty())
        return;
    const LocalMapId lastMapId = mapNodes.IdMax();
    const auto& lastMap = mBuilder->LocalMapAt(lastMapId);
    const NodeId lastScanId { (lastMap.mScanNodeIdMin.mId + lastMap.mScanNodeIdMax.mId) / 2 };
    const LoopSearchHint hint { std::move(scanNodes), std::move(mapNodes), mBuilder->AccumTravelDist(), lastScanId, lastMapId };
    const LoopCandidateVector candidates = mSearcher->Search(hint);
    if (candidates.empty())
        return;
    c[3] += 1;
    /* GetLoopDetectionQueries (:384-415) */
    LoopDetectionQueryVector queries;
    queries.reserve(candidates.size());
    for (const auto& cand : candidates)
        queries.emplace_back(mPoseGraph->ScanNodes().at(cand.mQueryScanNodeId),
                             mPoseGraph->ScanNodes().at(cand.mReferenceScanNodeId),
                             mBuilder->LocalMapAt(cand.mReferenceLocalMapId),
                             mPoseGraph->LocalMapNodes().at(cand.mReferenceLocalMapId));
    const double t0 = Now();
    const LoopDetectionResultVector results = mLoopDetector->Detect(queries);
    c[13] += Now() - t0;
    c[4] += static_cast<double>(queries.size());
    if (results.empty())
        return;
    c[5] += static_cast<double>(results.size());
    /* AppendLoopClosingEdges (:448-504) */
    for (const auto& r : results) {
        mLoops.push_back(r);
        mPoseGraph->Edges().emplace_back(r.mLocalMapNodeId, r.mScanNodeId, EdgeType::InterLocalMap, ConstraintType::Loop,
                                         NormalizeAngle(r.mRelativePose), r.mEstimatedCovMat.inverse());
    }
    /* the optimiser would run here (lidar_graph_slam_backend.cpp:170-172) on the finished part of the graph
     * (lidar_graph_slam.cpp:106-194); without one those poses stay. AfterLoopClosure (:506-672) then lets the
     * nodes added since follow along their odometry edges, which re-derives their poses even when nothing
     * moved: restated here so that both arms round the same way. */
    c[6] += 1;
    const auto firstOpen = std::find_if(localMaps.cbegin(), localMaps.cend(),
        [](const IdMap<LocalMapId, LocalMap>::ConstIdDataPair& pair) { return !pair.mData.mFinished; });
    const auto& lastDone = std::prev(firstOpen)->mData;
    auto& edges = mPoseGraph->Edges();
    auto it = std::find_if(edges.cbegin(), edges.cend(), [&lastDone](const PoseGraphEdge& e) {
        return e.mLocalMapNodeId == lastDone.mId && e.mScanNodeId > lastDone.mScanNodeIdMax; });
    if (it != edges.cend()) {
        LocalMapId doneMap = lastDone.mId;
        NodeId doneNode = lastDone.mScanNodeIdMax;
        for (; it != edges.cend(); ++it) {
            if (!it->IsOdometryConstraint())
                continue;
            if (it->mLocalMapNodeId == doneMap && it->mScanNodeId > doneNode)
                mPoseGraph->ScanNodes().at(it->mScanNodeId).mGlobalPose =
                    Compound(mPoseGraph->LocalMapNodes().at(it->mLocalMapNodeId).mGlobalPose, it->mRelativePose);
            else if (it->mLocalMapNodeId > doneMap && it->mScanNodeId == doneNode)
                mPoseGraph->LocalMapNodes().at(it->mLocalMapNodeId).mGlobalPose =
                    MoveBackward(mPoseGraph->ScanNodes().at(it->mScanNodeId).mGlobalPose, it->mRelativePose);
            doneMap = it->mLocalMapNodeId;
            doneNode = it->mScanNodeId;
        }
    }
    mBuilder->AfterLoopClosure(mPoseGraph);
}

} /* namespace */

void* orc_slam_create(const double* v, int n)
{
    if (n != 37)
        return nullptr;
    auto* s = new RefSlam;
    s->v.assign(v, v + n);
    /* the builder registers its metrics under fixed names: one instance per process, reset here */
    void* mb = orc_mapbuilder_create(v[0], static_cast<int>(v[1]), static_cast<int>(v[2]), v[5], v[6], v[7], v[8]);
    GridMapBuilder& b = *static_cast<RefMapBuilder*>(mb)->mBuilder;
    const_cast<double&>(b.mTravelDistThreshold) = v[3];
    const_cast<std::size_t&>(b.mNumOfOverlappedScans) = static_cast<std::size_t>(v[4]);
    b.mLocalMaps = IdMap<LocalMapId, LocalMap>();
    b.mAccumTravelDist = 0.0;
    b.mTravelDistLastLocalMap = 0.0;
    b.mLastRobotPose = RobotPose2D<double>(0.0, 0.0, 0.0);
    b.mRobotPoseLastLocalMap = RobotPose2D<double>(0.0, 0.0, 0.0);
    s->mBuilder = &b;
    s->mPoseGraph = std::make_shared<PoseGraph>();
    s->mScanMatcher.reset(new ScanMatcherCorrelative(UniqueName("SlamRT"), std::make_shared<CostSquareError>(v[26]),
                                                     static_cast<int>(v[19]), v[20], v[21], v[22]));
    s->mFinalMatcher.reset(new ScanMatcherLinearSolver(UniqueName("SlamFinal"), static_cast<int>(v[23]), v[24], v[25],
                                                       std::make_shared<CostSquareError>(v[26])));
    if (gSearcher == nullptr)
        gSearcher = new LoopSearcherNearest(v[27], v[28], static_cast<int>(v[29]));
    const_cast<double&>(gSearcher->mTravelDistThreshold) = v[27];
    const_cast<double&>(gSearcher->mNodeDistThreshold) = v[28];
    const_cast<int&>(gSearcher->mNumOfCandidateNodes) = static_cast<int>(v[29]);
    s->mSearcher = gSearcher;
    s->mLoopMatcher = std::make_shared<ScanMatcherBranchBound>(
        UniqueName("SlamBB"), std::make_shared<ScorePixelAccurate>(), std::make_shared<CostSquareError>(v[26]),
        static_cast<int>(v[30]), v[31], v[32], v[33]);
    std::shared_ptr<ScanMatcher> loopFinal = std::make_shared<ScanMatcherLinearSolver>(
        UniqueName("SlamLoopFinal"), static_cast<int>(v[23]), v[24], v[25], std::make_shared<CostSquareError>(v[26]));
    s->mLoopDetector.reset(new LoopDetectorBranchBound(UniqueName("SlamLoopDet"), s->mLoopMatcher, loopFinal, v[34], v[35]));
    return s;
}

void orc_slam_destroy(void* p) { delete static_cast<RefSlam*>(p); }

int orc_slam_run(void* p, int n_scans, int n_beams, const double* angles, const double* ranges,
                 const double* odom_poses, const double* time_stamps, double min_range, double max_range, int finish)
{
    auto* s = static_cast<RefSlam*>(p);
    const RobotPose2D<double> zero { 0.0, 0.0, 0.0 };
    int used = 0;
    for (int k = 0; k < n_scans; ++k) {
        std::vector<double> a(angles, angles + n_beams);
        std::vector<double> r(ranges + static_cast<std::size_t>(k) * n_beams, ranges + static_cast<std::size_t>(k + 1) * n_beams);
        auto scan = std::make_shared<Sensor::ScanData<double>>(
            "lidar", time_stamps[k], zero, zero, zero, min_range, max_range, a.front(), a.back(), std::move(a), std::move(r));
        used += s->ProcessScan(scan, RobotPose2D<double> { odom_poses[3 * k], odom_poses[3 * k + 1], odom_poses[3 * k + 2] },
                               time_stamps[k]) ? 1 : 0;
    }
    if (finish) {
        const double t0 = Now();
        s->RunBackendStep();
        s->c[12] += Now() - t0;
    }
    return used;
}

void orc_slam_counters(void* p, double* out)
{
    auto* s = static_cast<RefSlam*>(p);
    std::copy(s->c, s->c + 14, out);
    out[8] = 0.0;         /* no optimiser behind the seam in this arm */
}

int orc_slam_num_scan_nodes(void* p) { return static_cast<int>(static_cast<RefSlam*>(p)->mPoseGraph->ScanNodes().size()); }
int orc_slam_num_local_maps(void* p) { return static_cast<int>(static_cast<RefSlam*>(p)->mBuilder->LocalMaps().size()); }
int orc_slam_num_edges(void* p) { return static_cast<int>(static_cast<RefSlam*>(p)->mPoseGraph->Edges().size()); }
int orc_slam_num_loops(void* p) { return static_cast<int>(static_cast<RefSlam*>(p)->mLoops.size()); }

void orc_slam_scan_nodes(void* p, double* out7)
{
    int i = 0;
    for (const auto& [id, n] : static_cast<RefSlam*>(p)->mPoseGraph->ScanNodes()) {
        double* o = out7 + 7 * i++;
        o[0] = n.mGlobalPose.mX; o[1] = n.mGlobalPose.mY; o[2] = n.mGlobalPose.mTheta;
        o[3] = n.mLocalPose.mX; o[4] = n.mLocalPose.mY; o[5] = n.mLocalPose.mTheta; o[6] = n.mLocalMapId.mId;
    }
}

void orc_slam_local_maps(void* p, double* out10)
{
    auto* s = static_cast<RefSlam*>(p);
    int i = 0;
    for (const auto& [id, m] : s->mBuilder->LocalMaps()) {
        double* o = out10 + 10 * i++;
        const RobotPose2D<double>& pose = s->mPoseGraph->LocalMapNodes().at(id).mGlobalPose;
        o[0] = pose.mX; o[1] = pose.mY; o[2] = pose.mTheta;
        o[3] = m.mScanNodeIdMin.mId; o[4] = m.mScanNodeIdMax.mId; o[5] = m.mFinished ? 1.0 : 0.0;
        o[6] = m.mMap.Rows(); o[7] = m.mMap.Cols(); o[8] = m.mMap.PosOffset().mX; o[9] = m.mMap.PosOffset().mY;
    }
}

int orc_slam_local_map_cells(void* p, int id, uint16_t* dense, int cap_cells, uint8_t* alloc, int cap_blocks)
{
    auto* s = static_cast<RefSlam*>(p);
    if (id < 0 || id >= static_cast<int>(s->mBuilder->LocalMaps().size()))
        return -1;
    const GridMap& map = s->mBuilder->LocalMapAt(LocalMapId { id }).mMap;
    if (map.Rows() * map.Cols() > cap_cells || map.BlockRows() * map.BlockCols() > cap_blocks)
        return -2;
    Flatten(map, dense);
    for (int br = 0; br < map.BlockRows(); ++br)
        for (int bc = 0; bc < map.BlockCols(); ++bc)
            alloc[br * map.BlockCols() + bc] = map.Block(br, bc)->IsAllocated() ? 1 : 0;
    return 0;
}

void orc_slam_edges(void* p, double* out7)
{
    int i = 0;
    for (const auto& e : static_cast<RefSlam*>(p)->mPoseGraph->Edges()) {
        double* o = out7 + 7 * i++;
        o[0] = e.mLocalMapNodeId.mId; o[1] = e.mScanNodeId.mId; o[2] = e.mEdgeType == EdgeType::InterLocalMap ? 1.0 : 0.0;
        o[3] = e.IsLoopClosingConstraint() ? 1.0 : 0.0;
        o[4] = e.mRelativePose.mX; o[5] = e.mRelativePose.mY; o[6] = e.mRelativePose.mTheta;
    }
}

void orc_slam_loops(void* p, double* out6)
{
    int i = 0;
    for (const auto& r : static_cast<RefSlam*>(p)->mLoops) {
        double* o = out6 + 6 * i++;
        o[0] = r.mLocalMapNodeId.mId; o[1] = r.mScanNodeId.mId;
        o[2] = r.mRelativePose.mX; o[3] = r.mRelativePose.mY; o[4] = r.mRelativePose.mTheta;
        o[5] = 0.0;        /* the reference's result carries no score */
    }
}

void* orc_loopdet_create(int hmax, double range_x, double range_y, double range_t,
                         double score_thr, double known_thr, int n_threads)
{
    auto* det = new RefLoopDetector;
    det->mHeightMax = hmax;
    det->mRangeX = range_x;
    det->mRangeY = range_y;
    det->mRangeT = range_t;
    det->mScoreThr = score_thr;
    det->mKnownThr = known_thr;
    det->mNumThreads = std::max(1, n_threads);
    det->Build();
    return det;
}

void orc_loopdet_destroy(void* det)
{
    delete static_cast<RefLoopDetector*>(det);
}

void orc_loopdet_clear_cache(void* det)
{
    /* The reference never evicts its pyramid cache
     * (loop_detector_branch_bound.cpp:83-89): rebuild the detectors */
    static_cast<RefLoopDetector*>(det)->Build();
}

int orc_loopdet_detect(void* detPtr, int n_queries,
                       void* const* grids, const int32_t* map_ids,
                       const double* map_global_poses,
                       const int32_t* scan_idx, const double* scan_global_poses,
                       int n_scans, int n_beams,
                       const double* angles, const double* ranges,
                       orc_result* out, double* elapsed_s)
{
    auto* det = static_cast<RefLoopDetector*>(detPtr);
    const double rel[3] = { 0.0, 0.0, 0.0 };

    /* Build the pose-graph objects the queries refer to */
    std::vector<Sensor::ScanDataPtr<double>> scans;
    for (int s = 0; s < n_scans; ++s)
        scans.push_back(MakeScan(angles + static_cast<std::size_t>(s) * n_beams,
                                 ranges + static_cast<std::size_t>(s) * n_beams,
                                 n_beams, rel));

    std::vector<std::unique_ptr<LocalMap>> localMaps;
    std::vector<std::unique_ptr<LocalMapNode>> localMapNodes;
    std::vector<std::unique_ptr<ScanNode>> scanNodes;
    std::vector<std::unique_ptr<ScanNode>> refScanNodes;
    localMaps.reserve(n_queries);
    for (int q = 0; q < n_queries; ++q) {
        const LocalMapId mapId { map_ids[q] };
        GridMap copy = static_cast<RefGrid*>(grids[q])->mMap;
        auto localMap = std::make_unique<LocalMap>(
            mapId, std::move(copy), NodeId { 0 });
        localMap->mFinished = true;
        localMaps.push_back(std::move(localMap));
        const RobotPose2D<double> mapPose { map_global_poses[3 * q],
            map_global_poses[3 * q + 1], map_global_poses[3 * q + 2] };
        localMapNodes.push_back(std::make_unique<LocalMapNode>(mapId, mapPose));
        const RobotPose2D<double> scanPose { scan_global_poses[3 * q],
            scan_global_poses[3 * q + 1], scan_global_poses[3 * q + 2] };
        const RobotPose2D<double> zero { 0.0, 0.0, 0.0 };
        /* Query scan node: Id = query index so results can be mapped back */
        scanNodes.push_back(std::make_unique<ScanNode>(
            NodeId { q }, LocalMapId { -1 }, zero, scans[scan_idx[q]], scanPose));
        /* Reference scan node inside the local map (only its local pose and
         * map Id are read, loop_detector_branch_bound.cpp:110-118) */
        refScanNodes.push_back(std::make_unique<ScanNode>(
            NodeId { n_queries + q }, mapId, zero, scans[scan_idx[q]], mapPose));
    }

    const int nThreads = det->mNumThreads;
    std::vector<LoopDetectionResultVector> results(nThreads);
    std::vector<double> times(nThreads, 0.0);

    auto worker = [&](const int t) {
        const int begin = static_cast<int>(
            static_cast<long long>(n_queries) * t / nThreads);
        const int end = static_cast<int>(
            static_cast<long long>(n_queries) * (t + 1) / nThreads);
        LoopDetectionQueryVector queries;
        queries.reserve(end - begin);
        for (int q = begin; q < end; ++q)
            queries.emplace_back(*scanNodes[q], *refScanNodes[q],
                                 *localMaps[q], *localMapNodes[q]);
        const auto t0 = std::chrono::steady_clock::now();
        results[t] = det->mDetectors[t]->Detect(queries);
        const auto t1 = std::chrono::steady_clock::now();
        times[t] = std::chrono::duration<double>(t1 - t0).count();
    };

    if (nThreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> threads;
        for (int t = 0; t < nThreads; ++t)
            threads.emplace_back(worker, t);
        for (auto& th : threads)
            th.join();
    }

    if (elapsed_s != nullptr)
        *elapsed_s = *std::max_element(times.begin(), times.end());

    for (int q = 0; q < n_queries; ++q) {
        out[q] = orc_result { };
        out[q].best_x = out[q].best_y = out[q].best_t = 0;
    }

    for (int t = 0; t < nThreads; ++t) {
        for (const auto& result : results[t]) {
            const int q = result.mScanNodeId.mId;
            orc_result* o = &out[q];
            const GridMap& map = localMaps[q]->mMap;
            const auto& scan = scans[scan_idx[q]];
            const RobotPose2D<double> initPose = InverseCompound(
                localMapNodes[q]->mGlobalPose, scanNodes[q]->mGlobalPose);
            const ScanMatchingSummary summary {
                true, 0.0, initPose, result.mRelativePose,
                result.mEstimatedCovMat };
            FillSummary(summary, o);
            RobotPose2D<double> sensorPose, bestPose;
            DecodeWindow(map, scan, initPose, summary, det->mRangeX,
                         det->mRangeY, det->mRangeT, o, sensorPose, bestPose);
            DiagnosePixelAccurate(map, scan, bestPose, o);
        }
    }

    /* Keep the metric registry from growing without bound */
    for (int t = 0; t < nThreads; ++t)
        ResetMatcherMetrics(det->mMatchers[t]->Name(), false);

    return 0;
}

int orc_hill_climb(void* grid, const double* angles, const double* ranges, int n,
                   const double init_pose[3], const double rel_sensor_pose[3],
                   double linear_step, double angular_step, int max_iterations,
                   int max_num_of_refinements, const double* greedy, orc_result* out)
{
    const GridMap& map = static_cast<RefGrid*>(grid)->mMap;
    const auto scan = MakeScan(angles, ranges, n, rel_sensor_pose);
    const RobotPose2D<double> initPose { init_pose[0], init_pose[1], init_pose[2] };
    const std::string name = UniqueName("HC");
    CostFuncPtr cost;
    if (greedy != nullptr)
        cost = std::make_shared<CostGreedyEndpoint>(greedy[0], greedy[1], greedy[2], static_cast<int>(greedy[3]),
                                                    greedy[4], greedy[5]);
    else
        cost = std::make_shared<CostSquareError>(kCovarianceScale);
    ScanMatcherHillClimbing matcher {
        name, linear_step, angular_step, max_iterations, max_num_of_refinements, cost };
    const ScanMatchingQuery query { map, Point2D<double> { 0.0, 0.0 }, scan, initPose };
    const ScanMatchingSummary summary = matcher.OptimizePose(query);
    *out = orc_result { };
    FillSummary(summary, out);
    out->n_processed = LastInt(name + ".NumOfIterations");
    out->n_ignored = LastInt(name + ".NumOfRefinements");
    return 0;
}

/* The reference's LoopSearcherNearest on a hint rebuilt from plain arrays */
int orc_loop_search(int n_scans, const int32_t* scan_ids, const double* scan_poses,
                    int n_maps, const int32_t* map_ids, const int32_t* map_scan_min,
                    const int32_t* map_scan_max, const int32_t* map_finished,
                    double accum_travel_dist, int last_finished_scan_id, int last_finished_map_id,
                    double travel_dist_threshold, double node_dist_threshold,
                    int num_of_candidate_nodes, int32_t* out_ids, int cap)
{
    IdMap<NodeId, ScanNodeData> scanNodes;
    for (int i = 0; i < n_scans; ++i)
        scanNodes.Append(NodeId { scan_ids[i] },
                         RobotPose2D<double> { scan_poses[3 * i], scan_poses[3 * i + 1], scan_poses[3 * i + 2] });
    IdMap<LocalMapId, LocalMapData> localMaps;
    for (int i = 0; i < n_maps; ++i)
        localMaps.Append(LocalMapId { map_ids[i] },
                         Point2D<double> { 0.0, 0.0 }, Point2D<double> { 0.0, 0.0 },
                         NodeId { map_scan_min[i] }, NodeId { map_scan_max[i] }, map_finished[i] != 0);
    const LoopSearchHint hint { std::move(scanNodes), std::move(localMaps), accum_travel_dist,
                                NodeId { last_finished_scan_id }, LocalMapId { last_finished_map_id } };
    static LoopSearcherNearest* searcher = nullptr;
    if (searcher == nullptr)
        searcher = new LoopSearcherNearest(travel_dist_threshold, node_dist_threshold, num_of_candidate_nodes);
    const_cast<double&>(searcher->mTravelDistThreshold) = travel_dist_threshold;
    const_cast<double&>(searcher->mNodeDistThreshold) = node_dist_threshold;
    const_cast<int&>(searcher->mNumOfCandidateNodes) = num_of_candidate_nodes;
    const LoopCandidateVector candidates = searcher->Search(hint);
    const int n = std::min(static_cast<int>(candidates.size()), cap);
    for (int i = 0; i < n; ++i) {
        out_ids[3 * i] = candidates[i].mQueryScanNodeId.mId;
        out_ids[3 * i + 1] = candidates[i].mReferenceScanNodeId.mId;
        out_ids[3 * i + 2] = candidates[i].mReferenceLocalMapId.mId;
    }
    return n;
}

/* ---- io/carmen/carmen_reader.cpp and the metric strings of metric/metric.hpp ------------------------- */
struct RefCarmen { std::vector<MyLidarGraphSlam::Sensor::SensorDataPtr> records; };

void* orc_carmen_load(const char* text)
{
    auto* p = new RefCarmen;
    std::istringstream in { std::string(text) };
    MyLidarGraphSlam::IO::Carmen::CarmenLogReader reader;
    reader.Load(in, p->records);
    return p;
}
void orc_carmen_destroy(void* p) { delete static_cast<RefCarmen*>(p); }
int orc_carmen_count(void* p) { return static_cast<int>(static_cast<RefCarmen*>(p)->records.size()); }
int orc_carmen_total_beams(void* p)
{
    std::size_t n = 0;
    for (const auto& r : static_cast<RefCarmen*>(p)->records)
        if (auto s = std::dynamic_pointer_cast<MyLidarGraphSlam::Sensor::ScanData<double>>(r))
            n += s->NumOfScans();
    return static_cast<int>(n);
}
/* the layout of csm_host_carmen_export: 15 values per record, then angles and ranges back to back */
void orc_carmen_export(void* p, double* head15, double* angles, double* ranges)
{
    using namespace MyLidarGraphSlam::Sensor;
    std::size_t at = 0;
    for (const auto& r : static_cast<RefCarmen*>(p)->records) {
        std::fill(head15, head15 + 15, 0.0);
        head15[1] = r->TimeStamp();
        if (auto o = std::dynamic_pointer_cast<OdometryData<double>>(r)) {
            head15[0] = 0.0;
            head15[2] = o->Pose().mX; head15[3] = o->Pose().mY; head15[4] = o->Pose().mTheta;
            head15[5] = o->Velocity().mX; head15[6] = o->Velocity().mTheta;
        } else if (auto s = std::dynamic_pointer_cast<ScanData<double>>(r)) {
            head15[0] = 1.0;
            head15[2] = s->OdomPose().mX; head15[3] = s->OdomPose().mY; head15[4] = s->OdomPose().mTheta;
            head15[5] = s->Velocity().mX; head15[6] = s->Velocity().mTheta;
            head15[7] = s->RelativeSensorPose().mX; head15[8] = s->RelativeSensorPose().mY;
            head15[9] = s->RelativeSensorPose().mTheta;
            head15[10] = s->MinRange(); head15[11] = s->MaxRange();
            head15[12] = s->MinAngle(); head15[13] = s->MaxAngle();
            head15[14] = static_cast<double>(s->NumOfScans());
            std::copy(s->Angles().begin(), s->Angles().end(), angles + at);
            std::copy(s->Ranges().begin(), s->Ranges().end(), ranges + at);
            at += s->NumOfScans();
        }
        head15 += 15;
    }
}
int orc_carmen_sensor_id(void* p, int i, char* buf, int cap)
{
    const std::string& id = static_cast<RefCarmen*>(p)->records.at(static_cast<std::size_t>(i))->SensorId();
    std::snprintf(buf, static_cast<std::size_t>(cap), "%s", id.c_str());
    return static_cast<int>(id.size());
}

/* a ValueSequence<int> (kind 0), <float> (1) or <uint64_t> (2) observes the values; returns what its
 * ToPropertyTree puts under "Values" (metric.hpp:611-621 -> VecToString, :42-59) */
int orc_metric_values_string(int kind, const double* values, int n, char* buf, int cap)
{
    using namespace MyLidarGraphSlam::Metric;
    std::string s;
    if (kind == 0) {
        ValueSequence<int> seq { "t" };
        for (int i = 0; i < n; ++i) seq.Observe(values[i]);
        s = VecToString(*seq.Values());
    } else if (kind == 1) {
        ValueSequence<float> seq { "t" };
        for (int i = 0; i < n; ++i) seq.Observe(values[i]);
        s = VecToString(*seq.Values());
    } else {
        ValueSequence<std::uint64_t> seq { "t" };
        for (int i = 0; i < n; ++i) seq.Observe(values[i]);
        s = VecToString(*seq.Values());
    }
    std::snprintf(buf, static_cast<std::size_t>(cap), "%s", s.c_str());
    return static_cast<int>(s.size());
}

} /* extern "C" */
