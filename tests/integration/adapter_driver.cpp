/* adapter_driver.cpp -- TEST INFRASTRUCTURE: runs the drop-in classes of csm_gpu_adapter.hpp (derived
 * from the reference's ScanMatcher / LoopDetector) next to the reference's own CPU classes, both
 * through the BASE-CLASS virtuals and on the reference's own query / result types, and hands both
 * outcomes to the Python test (tests/test_integration_adapter.py).
 *
 * Built by oracle/Makefile (target `adapter`) from the reference's headers under /root/reference with the
 * shims of oracle/ref_shim, into oracle/_ref/libcsm_adapter.so (together with the reference's translation
 * units, like oracle/_ref/libcsm_ref.so). Only tests load it. */
#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "csm_gpu_adapter.hpp"

#include "my_lidar_graph_slam/mapping/cost_function_square_error.hpp"
#include "my_lidar_graph_slam/mapping/score_function_pixel_accurate.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_correlative.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_branch_bound.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_grid_search.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher_linear_solver.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector_branch_bound.hpp"

using namespace MyLidarGraphSlam;
using namespace MyLidarGraphSlam::Mapping;

namespace {

std::atomic<int> gCounter { 0 };
std::string Unique(const char* prefix) { return std::string(prefix) + "@adapter" + std::to_string(gCounter++); }

GridMap MakeMap(const uint16_t* dense, int rows, int cols, double res, double offx, double offy)
{
    GridMap map { res, 16, rows / 16, cols / 16, Point2D<double> { offx, offy } };
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const std::uint16_t v = dense[static_cast<std::size_t>(r) * cols + c];
            if (v != 0) map.SetValue(r, c, v);
        }
    return map;
}

Sensor::ScanDataPtr<double> MakeScan(const double* angles, const double* ranges, int n, const double rel[3])
{
    std::vector<double> a(angles, angles + n), r(ranges, ranges + n);
    const RobotPose2D<double> zero { 0.0, 0.0, 0.0 };
    const RobotPose2D<double> relPose { rel[0], rel[1], rel[2] };
    return std::make_shared<Sensor::ScanData<double>>(
        "lidar", 0.0, zero, zero, relPose, 0.01, 100.0, a.front(), a.back(), std::move(a), std::move(r));
}

/* 14 doubles per summary: found, normalized cost, estimated pose (3), covariance (9) */
void Export(const ScanMatchingSummary& s, double* out)
{
    out[0] = s.mPoseFound ? 1.0 : 0.0;
    out[1] = s.mNormalizedCost;
    out[2] = s.mEstimatedPose.mX; out[3] = s.mEstimatedPose.mY; out[4] = s.mEstimatedPose.mTheta;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b)
            out[5 + 3 * a + b] = s.mEstimatedCovariance(a, b);
}

} /* namespace */

extern "C" {

/* One single-scan match through ScanMatcher::OptimizePose(const ScanMatchingQuery&) of the reference's
 * CPU class and of the GPU drop-in. kind: 0 = real-time correlative (param = LowResolution), 1 = branch and
 * bound (param = NodeHeightMax), 2 = grid search (step = SearchStep*). Returns 0. */
int adp_match_case(int kind, const uint16_t* dense, int rows, int cols, double res, double offx, double offy,
                   const double* angles, const double* ranges, int n, const double init_pose[3],
                   const double rel_pose[3], int param, const double range[3], const double step[3],
                   double* out_cpu14, double* out_gpu14, int* gpu_flags)
{
    const GridMap map = MakeMap(dense, rows, cols, res, offx, offy);
    const auto scan = MakeScan(angles, ranges, n, rel_pose);
    const RobotPose2D<double> init { init_pose[0], init_pose[1], init_pose[2] };
    const ScanMatchingQuery query { map, Point2D<double> { 0.0, 0.0 }, scan, init };
    auto cost = [] { return std::make_shared<CostSquareError>(1e4); };
    std::shared_ptr<ScanMatcher> cpu, gpu;
    if (kind == 0) {
        cpu = std::make_shared<ScanMatcherCorrelative>(Unique("RT"), cost(), param, range[0], range[1], range[2]);
        gpu = CreateScanMatcherGPU("RealTimeCorrelativeGPU", Unique("RTGPU"), cost(), param, range[0], range[1], range[2]);
    } else if (kind == 1) {
        cpu = std::make_shared<ScanMatcherBranchBound>(Unique("BB"), std::make_shared<ScorePixelAccurate>(), cost(),
                                                       param, range[0], range[1], range[2]);
        gpu = CreateScanMatcherGPU("BranchBoundGPU", Unique("BBGPU"), cost(), param, range[0], range[1], range[2]);
    } else {
        cpu = std::make_shared<ScanMatcherGridSearch>(Unique("GS"), std::make_shared<ScorePixelAccurate>(), cost(),
                                                      range[0], range[1], range[2], step[0], step[1], step[2]);
        gpu = CreateScanMatcherGPU("GridSearchGPU", Unique("GSGPU"), cost(), 0, range[0], range[1], range[2],
                                   step[0], step[1], step[2]);
    }
    if (!cpu || !gpu)
        return -1;
    Export(cpu->OptimizePose(query), out_cpu14);       /* both through the base-class virtual */
    Export(gpu->OptimizePose(query), out_gpu14);
    *gpu_flags = kind == 0 ? static_cast<ScanMatcherCorrelativeGPU*>(gpu.get())->LastResult().flags
               : kind == 1 ? static_cast<ScanMatcherBranchBoundGPU*>(gpu.get())->LastResult().flags
                           : static_cast<ScanMatcherGridSearchGPU*>(gpu.get())->LastResult().flags;
    return 0;
}

/* LoopDetector::Detect(const LoopDetectionQueryVector&) of the reference's LoopDetectorBranchBound and of
 * LoopDetectorBranchBoundGPU on the same queries (one scan against n_queries finished local maps), both
 * with the reference's ScanMatcherLinearSolver as final matcher (device_refiner != 0: the GPU detector
 * runs that solver on the device instead). Per detector: the number of results, and per result 2 ids
 * (local map, scan node) and 15 doubles (relative pose 3, local map pose 3, covariance 9). */
int adp_loop_case(int n_queries, const uint16_t* dense, int rows, int cols, double res,
                  const double* offx, const double* offy, const int32_t* map_ids,
                  const double* map_poses, const double* scan_poses,
                  const double* angles, const double* ranges, int n,
                  int hmax, const double range[3], double score_thr, double known_thr, int device_refiner,
                  int repeats, int* n_cpu, int32_t* ids_cpu, double* out_cpu, int* n_gpu, int32_t* ids_gpu,
                  double* out_gpu, int* gpu_flags_or)
{
    const double rel[3] = { 0.0, 0.0, 0.0 };
    const auto scan = MakeScan(angles, ranges, n, rel);
    const std::size_t cells = static_cast<std::size_t>(rows) * cols;
    std::vector<std::unique_ptr<LocalMap>> localMaps;
    std::vector<std::unique_ptr<LocalMapNode>> localMapNodes;
    std::vector<std::unique_ptr<ScanNode>> scanNodes, refScanNodes;
    const RobotPose2D<double> zero { 0.0, 0.0, 0.0 };
    for (int q = 0; q < n_queries; ++q) {
        const LocalMapId mapId { map_ids[q] };
        auto localMap = std::make_unique<LocalMap>(mapId, MakeMap(dense + q * cells, rows, cols, res, offx[q], offy[q]), NodeId { 0 });
        localMap->mFinished = true;
        localMaps.push_back(std::move(localMap));
        const RobotPose2D<double> mapPose { map_poses[3 * q], map_poses[3 * q + 1], map_poses[3 * q + 2] };
        localMapNodes.push_back(std::make_unique<LocalMapNode>(mapId, mapPose));
        const RobotPose2D<double> scanPose { scan_poses[3 * q], scan_poses[3 * q + 1], scan_poses[3 * q + 2] };
        scanNodes.push_back(std::make_unique<ScanNode>(NodeId { q }, LocalMapId { -1 }, zero, scan, scanPose));
        refScanNodes.push_back(std::make_unique<ScanNode>(NodeId { n_queries + q }, mapId, zero, scan, mapPose));
    }
    LoopDetectionQueryVector queries;          /* the reference's own query vector */
    for (int q = 0; q < n_queries; ++q)
        queries.emplace_back(*scanNodes[q], *refScanNodes[q], *localMaps[q], *localMapNodes[q]);

    auto solver = [] {
        return std::make_shared<ScanMatcherLinearSolver>(Unique("Final"), 10, 1e-4, 1e-4, std::make_shared<CostSquareError>(1e4));
    };
    auto bbCpu = std::make_shared<ScanMatcherBranchBound>(Unique("LoopBB"), std::make_shared<ScorePixelAccurate>(),
                                                          std::make_shared<CostSquareError>(1e4), hmax, range[0], range[1], range[2]);
    auto bbGpu = std::make_shared<ScanMatcherBranchBoundGPU>(Unique("LoopBBGPU"), std::make_shared<CostSquareError>(1e4),
                                                             hmax, range[0], range[1], range[2]);
    auto gpuDetector = std::make_shared<LoopDetectorBranchBoundGPU>(Unique("LoopDetGPU"), bbGpu, solver(), score_thr, known_thr);
    if (device_refiner)
        gpuDetector->UseDeviceRefiner(10, 1e-4, 1e-4, 1e4);
    std::shared_ptr<LoopDetector> cpu = std::make_shared<LoopDetectorBranchBound>(Unique("LoopDet"), bbCpu, solver(), score_thr, known_thr);
    std::shared_ptr<LoopDetector> gpu = gpuDetector;

    auto put = [](const LoopDetectionResultVector& results, int* count, int32_t* ids, double* out) {
        *count = static_cast<int>(results.size());
        for (std::size_t i = 0; i < results.size(); ++i) {
            const LoopDetectionResult& r = results[i];
            ids[2 * i] = r.mLocalMapNodeId.mId; ids[2 * i + 1] = r.mScanNodeId.mId;
            double* o = out + 15 * i;
            o[0] = r.mRelativePose.mX; o[1] = r.mRelativePose.mY; o[2] = r.mRelativePose.mTheta;
            o[3] = r.mLocalMapPose.mX; o[4] = r.mLocalMapPose.mY; o[5] = r.mLocalMapPose.mTheta;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    o[6 + 3 * a + b] = r.mEstimatedCovMat(a, b);
        }
    };
    /* `repeats` calls each: from the second on both detectors find their maps cached */
    for (int rep = 0; rep < repeats; ++rep) {
        put(cpu->Detect(queries), n_cpu, ids_cpu, out_cpu);
        put(gpu->Detect(queries), n_gpu, ids_gpu, out_gpu);
    }
    int flags = 0;
    for (const csm_result& r : gpuDetector->LastResults()) flags |= r.flags;
    *gpu_flags_or = flags;
    return 0;
}

} /* extern "C" */
