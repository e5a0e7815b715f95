/* csm_gpu_adapter.hpp -- the drop-in, on the reference's OWN types.
 *
 * What a maintainer adds to sterngerlach/my-lidar-graph-slam-v2 to run the correlative scan matching /
 * loop-detection hot path on a B200: classes that derive from the reference's plugin interfaces
 *     Mapping::ScanMatcher   (include/my_lidar_graph_slam/mapping/scan_matcher.hpp:89-117)
 *     Mapping::LoopDetector  (include/my_lidar_graph_slam/mapping/loop_detector.hpp:97-116)
 * take the reference's ScanMatchingQuery / LoopDetectionQueryVector, and return its ScanMatchingSummary /
 * LoopDetectionResultVector, and bind to the C ABI of include/csm_b200.h (libcsm_b200.so) underneath.
 * Constructor parameters are those of the CPU classes (scan_matcher_correlative.hpp:59-65,
 * scan_matcher_branch_bound.hpp:110-117, scan_matcher_grid_search.hpp:48-57,
 * loop_detector_branch_bound.hpp:72-77) plus the CUDA device; the factories at the end take the type
 * strings that go next to scan_matcher_factory.cpp:194-217 / loop_detector_factory.cpp:197-212.
 *
 * In this repository the header is compiled against /root/reference/include (with the header shims
 * of oracle/ref_shim standing in for Eigen / Boost) by oracle/Makefile (target `adapter`) and driven
 * through the base-class virtuals by tests/integration/adapter_driver.cpp next to the reference's
 * own CPU classes: tests/test_integration_adapter.py. It is test infrastructure here, product code in
 * the reference tree.
 */
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "csm_b200.h"

#include "my_lidar_graph_slam/pose.hpp"
#include "my_lidar_graph_slam/point.hpp"
#include "my_lidar_graph_slam/util.hpp"
#include "my_lidar_graph_slam/sensor/sensor_data.hpp"
#include "my_lidar_graph_slam/mapping/grid_map_types.hpp"
#include "my_lidar_graph_slam/mapping/cost_function.hpp"
#include "my_lidar_graph_slam/mapping/scan_matcher.hpp"
#include "my_lidar_graph_slam/mapping/loop_detector.hpp"
#include "my_lidar_graph_slam/mapping/pose_graph.hpp"

namespace MyLidarGraphSlam {
namespace Mapping {

/* One csm_handle (a CUDA stream and its device-resident maps) per matcher / detector instance: the
 * front end and the back end call different instances from different threads
 * (lidar_graph_slam.cpp:777-779) and never share one. */
class CsmDevice final
{
public:
    explicit CsmDevice(const int device) : mHandle(nullptr)
    { Assert(csm_create(device, 0, &this->mHandle) == CSM_OK); }
    ~CsmDevice() { csm_destroy(this->mHandle); }
    CsmDevice(const CsmDevice&) = delete;
    CsmDevice& operator=(const CsmDevice&) = delete;
    inline csm_handle Handle() const { return this->mHandle; }
    /* The reference's convention: Assert (print + abort, util.hpp:39-72); "no pose found" is a result */
    inline void Check(const int status) const
    {
        if (status != CSM_OK)
            std::fprintf(stderr, "csm: %s\n", csm_last_error(this->mHandle));
        Assert(status == CSM_OK);
    }

    /* The map goes over in its own storage form: the blocks that are allocated (each one a separate
     * heap allocation, grid_map.cpp:522-535; GridMap::Block grid_map.hpp:198-199,
     * GridBinaryBayes::IsAllocated / Data grid_binary_bayes.hpp:69-79), gathered back to back into
     * page-locked staging, plus their positions. Unallocated blocks never cross PCIe. */
    void Upload(const GridMap& map, const std::int64_t id)
    {
        const int cells = map.BlockSize() * map.BlockSize();
        const int nBlocksMax = map.BlockRows() * map.BlockCols();
        const std::size_t need = static_cast<std::size_t>(nBlocksMax) * (cells * sizeof(std::uint16_t) + sizeof(std::int32_t));
        if (this->mStagingBytes < need) {
            if (this->mStaging != nullptr) csm_free_pinned(this->mStaging);
            this->mStaging = csm_alloc_pinned(need);
            Assert(this->mStaging != nullptr);
            this->mStagingBytes = need;
        }
        auto* blocks = static_cast<std::uint16_t*>(this->mStaging);
        auto* index = reinterpret_cast<std::int32_t*>(
            static_cast<char*>(this->mStaging) + static_cast<std::size_t>(nBlocksMax) * cells * sizeof(std::uint16_t));
        int n = 0;
        for (int br = 0; br < map.BlockRows(); ++br)
            for (int bc = 0; bc < map.BlockCols(); ++bc) {
                const auto* block = map.Block(br, bc);
                if (!block->IsAllocated())
                    continue;
                std::memcpy(blocks + static_cast<std::size_t>(n) * cells, block->Data(), cells * sizeof(std::uint16_t));
                index[n++] = br * map.BlockCols() + bc;
            }
        this->Check(csm_upload_grid_blocks(this->mHandle, id, blocks, index, n, map.Log2BlockSize(),
                                           map.BlockRows(), map.BlockCols(), map.Resolution(),
                                           map.PosOffset().mX, map.PosOffset().mY));
        /* the staging area is reused by the next upload: wait for this copy */
        this->Check(csm_synchronize(this->mHandle));
    }

private:
    csm_handle mHandle;
    void* mStaging = nullptr;
    std::size_t mStagingBytes = 0;
};

/* scan_matcher_correlative.cpp:255-274 / scan_matcher_branch_bound.cpp:293-312 */
inline void CsmSearchStep(const GridMap& gridMap, const Sensor::ScanDataPtr<double>& scanData,
                          double& stepX, double& stepY, double& stepTheta)
{
    const auto maxRangeIt = std::max_element(scanData->Ranges().begin(), scanData->Ranges().end());
    const double maxRange = *maxRangeIt;
    const double theta = gridMap.Resolution() / maxRange;
    stepX = gridMap.Resolution();
    stepY = gridMap.Resolution();
    stepTheta = std::acos(1.0 - 0.5 * theta * theta);
}

constexpr std::int64_t kCsmAnonymousMap = static_cast<std::int64_t>(1) << 40;   /* the per-scan latest map */

/* "BranchBoundGPU": ScanMatcherBranchBound (scan_matcher_branch_bound.cpp:87-278) on the device */
class ScanMatcherBranchBoundGPU final : public ScanMatcher
{
public:
    ScanMatcherBranchBoundGPU(const std::string& scanMatcherName, const CostFuncPtr& costFunc,
                              const int nodeHeightMax, const double rangeX, const double rangeY,
                              const double rangeTheta, const int device = 0) :
        ScanMatcher(scanMatcherName), mCostFunc(costFunc), mNodeHeightMax(nodeHeightMax),
        mRangeX(rangeX), mRangeY(rangeY), mRangeTheta(rangeTheta), mDevice(new CsmDevice(device)) { }

    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& queryInfo) override
    {
        this->mDevice->Upload(queryInfo.mGridMap, kCsmAnonymousMap);
        return this->OptimizePose(queryInfo.mGridMap, kCsmAnonymousMap, queryInfo.mScanData,
                                  queryInfo.mMapLocalInitialPose, 0.0, 0.0);
    }

    /* `mapId`: the map as uploaded (a finished local map is uploaded and precomputed once) */
    ScanMatchingSummary OptimizePose(const GridMap& gridMap, const std::int64_t mapId,
                                     const Sensor::ScanDataPtr<double>& scanData,
                                     const RobotPose2D<double>& mapLocalInitialPose,
                                     const double normalizedScoreThreshold, const double knownRateThreshold)
    {
        const RobotPose2D<double> sensorPose = Compound(mapLocalInitialPose, scanData->RelativeSensorPose());
        double stepX, stepY, stepTheta;
        CsmSearchStep(gridMap, scanData, stepX, stepY, stepTheta);
        const int winX = static_cast<int>(std::ceil(0.5 * this->mRangeX / stepX));
        const int winY = static_cast<int>(std::ceil(0.5 * this->mRangeY / stepY));
        const int winTheta = static_cast<int>(std::ceil(0.5 * this->mRangeTheta / stepTheta));
        csm_handle h = this->mDevice->Handle();
        this->mDevice->Check(csm_build_pyramid(h, mapId, this->mNodeHeightMax));
        const double pose[3] = { sensorPose.mX, sensorPose.mY, sensorPose.mTheta };
        csm_result r;
        this->mDevice->Check(csm_match_bb(h, mapId, scanData->Angles().data(), scanData->Ranges().data(),
                                          static_cast<int>(scanData->NumOfScans()), pose, this->mNodeHeightMax,
                                          winX, winY, winTheta, stepX, stepY, stepTheta,
                                          normalizedScoreThreshold, knownRateThreshold, &r));
        this->mLastResult = r;
        /* scan_matcher_branch_bound.cpp:237-262: best pose from the window indices, the CPU epilogue unchanged */
        const RobotPose2D<double> bestSensorPose { sensorPose.mX + r.best_x * stepX,
                                                   sensorPose.mY + r.best_y * stepY,
                                                   sensorPose.mTheta + r.best_t * stepTheta };
        const double cost = this->mCostFunc->Cost(gridMap, scanData, bestSensorPose);
        const double normalizedCost = cost / scanData->NumOfScans();
        const RobotPose2D<double> estimatedPose = MoveBackward(bestSensorPose, scanData->RelativeSensorPose());
        const Eigen::Matrix3d estimatedCovariance = this->mCostFunc->ComputeCovariance(gridMap, scanData, bestSensorPose);
        return ScanMatchingSummary { r.found != 0, normalizedCost, mapLocalInitialPose, estimatedPose, estimatedCovariance };
    }

    inline int NodeHeightMax() const { return this->mNodeHeightMax; }
    inline double RangeX() const { return this->mRangeX; }
    inline double RangeY() const { return this->mRangeY; }
    inline double RangeTheta() const { return this->mRangeTheta; }
    inline CsmDevice& Device() { return *this->mDevice; }
    inline const csm_result& LastResult() const { return this->mLastResult; }

private:
    CostFuncPtr mCostFunc;
    const int mNodeHeightMax;
    const double mRangeX, mRangeY, mRangeTheta;
    std::unique_ptr<CsmDevice> mDevice;
    csm_result mLastResult {};
};

/* "RealTimeCorrelativeGPU": ScanMatcherCorrelative (scan_matcher_correlative.cpp:92-244) on the device */
class ScanMatcherCorrelativeGPU final : public ScanMatcher
{
public:
    ScanMatcherCorrelativeGPU(const std::string& scanMatcherName, const CostFuncPtr& costFunc,
                              const int lowResolution, const double rangeX, const double rangeY,
                              const double rangeTheta, const int device = 0) :
        ScanMatcher(scanMatcherName), mCostFunc(costFunc), mLowResolution(lowResolution),
        mRangeX(rangeX), mRangeY(rangeY), mRangeTheta(rangeTheta), mDevice(new CsmDevice(device)) { }

    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& queryInfo) override
    {
        const GridMap& gridMap = queryInfo.mGridMap;
        const auto& scanData = queryInfo.mScanData;
        csm_handle h = this->mDevice->Handle();
        this->mDevice->Upload(gridMap, kCsmAnonymousMap);
        this->mDevice->Check(csm_build_coarse(h, kCsmAnonymousMap, this->mLowResolution));
        const RobotPose2D<double> sensorPose = Compound(queryInfo.mMapLocalInitialPose, scanData->RelativeSensorPose());
        double stepX, stepY, stepTheta;
        CsmSearchStep(gridMap, scanData, stepX, stepY, stepTheta);
        const int winX = static_cast<int>(std::ceil(0.5 * this->mRangeX / stepX));
        const int winY = static_cast<int>(std::ceil(0.5 * this->mRangeY / stepY));
        const int winTheta = static_cast<int>(std::ceil(0.5 * this->mRangeTheta / stepTheta));
        const double pose[3] = { sensorPose.mX, sensorPose.mY, sensorPose.mTheta };
        csm_result r;
        this->mDevice->Check(csm_match_rt(h, kCsmAnonymousMap, scanData->Angles().data(), scanData->Ranges().data(),
                                          static_cast<int>(scanData->NumOfScans()), pose, this->mLowResolution,
                                          winX, winY, winTheta, stepX, stepY, stepTheta, 0.0, 0.0, &r));
        this->mLastResult = r;
        /* scan_matcher_correlative.cpp:199-219 */
        const RobotPose2D<double> bestSensorPose { sensorPose.mX + r.best_x * stepX,
                                                   sensorPose.mY + r.best_y * stepY,
                                                   sensorPose.mTheta + r.best_t * stepTheta };
        const double cost = this->mCostFunc->Cost(gridMap, scanData, bestSensorPose);
        const RobotPose2D<double> estimatedPose = MoveBackward(bestSensorPose, scanData->RelativeSensorPose());
        const Eigen::Matrix3d estimatedCovariance = this->mCostFunc->ComputeCovariance(gridMap, scanData, bestSensorPose);
        return ScanMatchingSummary { r.found != 0, cost / scanData->NumOfScans(), queryInfo.mMapLocalInitialPose,
                                     estimatedPose, estimatedCovariance };
    }
    inline const csm_result& LastResult() const { return this->mLastResult; }

private:
    CostFuncPtr mCostFunc;
    const int mLowResolution;
    const double mRangeX, mRangeY, mRangeTheta;
    std::unique_ptr<CsmDevice> mDevice;
    csm_result mLastResult {};
};

/* "GridSearchGPU": ScanMatcherGridSearch (scan_matcher_grid_search.cpp:69-178) on the device */
class ScanMatcherGridSearchGPU final : public ScanMatcher
{
public:
    ScanMatcherGridSearchGPU(const std::string& scanMatcherName, const CostFuncPtr& costFunc,
                             const double rangeX, const double rangeY, const double rangeTheta,
                             const double stepX, const double stepY, const double stepTheta, const int device = 0) :
        ScanMatcher(scanMatcherName), mCostFunc(costFunc), mRangeX(rangeX), mRangeY(rangeY), mRangeTheta(rangeTheta),
        mStepX(stepX), mStepY(stepY), mStepTheta(stepTheta), mDevice(new CsmDevice(device)) { }

    ScanMatchingSummary OptimizePose(const ScanMatchingQuery& queryInfo) override
    {
        const GridMap& gridMap = queryInfo.mGridMap;
        const auto& scanData = queryInfo.mScanData;
        csm_handle h = this->mDevice->Handle();
        this->mDevice->Upload(gridMap, kCsmAnonymousMap);
        const RobotPose2D<double> sensorPose = Compound(queryInfo.mMapLocalInitialPose, scanData->RelativeSensorPose());
        /* the candidate offsets of the reference's accumulating loops (scan_matcher_grid_search.cpp:118-120) */
        auto offsets = [](const double r, const double s) {
            std::vector<double> v;
            for (double d = -r; d <= r; d += s) v.push_back(d);
            return v;
        };
        const std::vector<double> dx = offsets(this->mRangeX / 2.0, this->mStepX);
        const std::vector<double> dy = offsets(this->mRangeY / 2.0, this->mStepY);
        const std::vector<double> dt = offsets(this->mRangeTheta / 2.0, this->mStepTheta);
        const double pose[3] = { sensorPose.mX, sensorPose.mY, sensorPose.mTheta };
        csm_result r;
        this->mDevice->Check(csm_match_grid(h, kCsmAnonymousMap, scanData->Angles().data(), scanData->Ranges().data(),
                                            static_cast<int>(scanData->NumOfScans()), pose,
                                            dx.data(), static_cast<int>(dx.size()), dy.data(), static_cast<int>(dy.size()),
                                            dt.data(), static_cast<int>(dt.size()), 0.0, 0.0, &r));
        this->mLastResult = r;
        const RobotPose2D<double> bestSensorPose = r.found
            ? RobotPose2D<double> { sensorPose.mX + dx[r.best_x], sensorPose.mY + dy[r.best_y], sensorPose.mTheta + dt[r.best_t] }
            : sensorPose;
        const double cost = this->mCostFunc->Cost(gridMap, scanData, bestSensorPose);
        const RobotPose2D<double> estimatedPose = MoveBackward(bestSensorPose, scanData->RelativeSensorPose());
        const Eigen::Matrix3d estimatedCovariance = this->mCostFunc->ComputeCovariance(gridMap, scanData, bestSensorPose);
        return ScanMatchingSummary { r.found != 0, cost / scanData->NumOfScans(), queryInfo.mMapLocalInitialPose,
                                     estimatedPose, estimatedCovariance };
    }
    inline const csm_result& LastResult() const { return this->mLastResult; }

private:
    CostFuncPtr mCostFunc;
    const double mRangeX, mRangeY, mRangeTheta, mStepX, mStepY, mStepTheta;
    std::unique_ptr<CsmDevice> mDevice;
    csm_result mLastResult {};
};

/* "BranchBoundGPU" loop detector: LoopDetectorBranchBound::Detect (loop_detector_branch_bound.cpp:59-156)
 * with all queries of the call matched in one device batch. The final matcher is the caller's
 * (any ScanMatcher, run on the CPU through its virtual OptimizePose like the reference does, :110-127);
 * UseDeviceRefiner() instead runs the reference's default, ScanMatcherLinearSolver with CostSquareError,
 * on the device in the same batch. */
class LoopDetectorBranchBoundGPU final : public LoopDetector
{
public:
    LoopDetectorBranchBoundGPU(const std::string& loopDetectorName,
                               const std::shared_ptr<ScanMatcherBranchBoundGPU>& scanMatcher,
                               const std::shared_ptr<ScanMatcher>& finalScanMatcher,
                               const double scoreThreshold, const double knownRateThreshold) :
        LoopDetector(loopDetectorName), mScanMatcher(scanMatcher), mFinalScanMatcher(finalScanMatcher),
        mScoreThreshold(scoreThreshold), mKnownRateThreshold(knownRateThreshold)
    {
        Assert(scoreThreshold > 0.0 && scoreThreshold <= 1.0);               /* :54-55 */
        Assert(knownRateThreshold > 0.0 && knownRateThreshold <= 1.0);
    }

    void UseDeviceRefiner(const int numOfIterationsMax, const double convergenceThreshold,
                          const double initialLambda, const double covarianceScale)
    {
        this->mRefine.max_iterations = numOfIterationsMax;
        this->mRefine.reserved = 0;
        this->mRefine.convergence_threshold = convergenceThreshold;
        this->mRefine.lambda = initialLambda;
        this->mRefine.covariance_scale = covarianceScale;
        this->mDeviceRefiner = true;
    }

    LoopDetectionResultVector Detect(const LoopDetectionQueryVector& queries) override
    {
        LoopDetectionResultVector results;
        if (queries.empty())
            return results;
        CsmDevice& dev = this->mScanMatcher->Device();
        csm_handle h = dev.Handle();
        const int hmax = this->mScanMatcher->NodeHeightMax();
        const int nq = static_cast<int>(queries.size());

        std::vector<csm_loop_query> dq(nq);
        std::vector<std::int64_t> newMaps;
        /* the scans of this call get call-local ids (the reference keeps no per-scan state) */
        std::vector<const Sensor::ScanData<double>*> scans;
        for (int i = 0; i < nq; ++i) {
            const auto& query = queries[i];
            const auto& scanNode = query.mQueryScanNode;
            const auto& localMap = query.mReferenceLocalMap;
            const auto& localMapNode = query.mReferenceLocalMapNode;
            Assert(localMap.mId == localMapNode.mLocalMapId);                /* :75 */
            Assert(localMap.mFinished);                                      /* :77 */
            const std::int64_t mapId = localMap.mId.mId;
            /* first touch of a finished (immutable) local map: upload; its pyramid is built below and
             * stays cached by LocalMapId like mPrecompMaps (:83-89) */
            if (this->mCachedMaps.insert(mapId).second) {
                dev.Upload(localMap.mMap, mapId);
                newMaps.push_back(mapId);
            }
            const Sensor::ScanData<double>* scan = scanNode.mScanData.get();
            auto it = std::find(scans.begin(), scans.end(), scan);
            const std::int64_t scanId = (static_cast<std::int64_t>(1) << 60) + (it - scans.begin());
            if (it == scans.end()) {
                scans.push_back(scan);
                dev.Check(csm_upload_scan(h, scanId, scan->Angles().data(), scan->Ranges().data(),
                                          static_cast<int>(scan->NumOfScans())));
            }
            /* :97-98 and scan_matcher_branch_bound.cpp:124-146 */
            const RobotPose2D<double> mapLocalScanPose = InverseCompound(localMapNode.mGlobalPose, scanNode.mGlobalPose);
            const RobotPose2D<double> sensorPose = Compound(mapLocalScanPose, scan->RelativeSensorPose());
            double stepX, stepY, stepTheta;
            CsmSearchStep(localMap.mMap, scanNode.mScanData, stepX, stepY, stepTheta);
            csm_loop_query& d = dq[i];
            d.map_id = mapId;
            d.scan_id = scanId;
            d.sensor_pose[0] = sensorPose.mX; d.sensor_pose[1] = sensorPose.mY; d.sensor_pose[2] = sensorPose.mTheta;
            d.win_x = static_cast<int>(std::ceil(0.5 * this->mScanMatcher->RangeX() / stepX));
            d.win_y = static_cast<int>(std::ceil(0.5 * this->mScanMatcher->RangeY() / stepY));
            d.win_t = static_cast<int>(std::ceil(0.5 * this->mScanMatcher->RangeTheta() / stepTheta));
            d.reserved = 0;
            d.step_x = stepX; d.step_y = stepY; d.step_t = stepTheta;
            d.score_thr = this->mScoreThreshold;
            d.known_thr = this->mKnownRateThreshold;
        }
        if (!newMaps.empty())
            dev.Check(csm_build_pyramids(h, static_cast<int>(newMaps.size()), newMaps.data(), hmax));
        dev.Check(csm_set_refiner(h, this->mDeviceRefiner ? &this->mRefine : nullptr));
        this->mLastResults.assign(nq, csm_result {});
        std::vector<csm_refined> refined(nq);
        dev.Check(csm_loop_batch_enqueue(h, dq.data(), nq, hmax, 0));
        if (this->mDeviceRefiner)
            dev.Check(csm_loop_batch_finish_refined(h, this->mLastResults.data(), refined.data(), nq));
        else
            dev.Check(csm_loop_batch_finish(h, this->mLastResults.data(), nq));
        for (std::size_t s = 0; s < scans.size(); ++s)
            dev.Check(csm_release_scan(h, (static_cast<std::int64_t>(1) << 60) + static_cast<std::int64_t>(s)));

        for (int i = 0; i < nq; ++i) {
            const csm_result& r = this->mLastResults[i];
            if (!r.found)
                continue;                                                    /* :106-108 */
            const auto& query = queries[i];
            const auto& scanNode = query.mQueryScanNode;
            const auto& localMap = query.mReferenceLocalMap;
            const auto& localMapNode = query.mReferenceLocalMapNode;
            const csm_loop_query& d = dq[i];
            const RobotPose2D<double> bestSensorPose { d.sensor_pose[0] + r.best_x * d.step_x,
                                                       d.sensor_pose[1] + r.best_y * d.step_y,
                                                       d.sensor_pose[2] + r.best_t * d.step_t };
            if (this->mDeviceRefiner && refined[i].valid) {
                const csm_refined& f = refined[i];
                Eigen::Matrix3d cov;
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        cov(a, b) = f.covariance[3 * a + b];
                this->mRefine.lambda = f.lambda;       /* the solver's damping state (scan_matcher_linear_solver.cpp:100-104) */
                results.emplace_back(MoveBackward(RobotPose2D<double> { f.pose[0], f.pose[1], f.pose[2] },
                                                  scanNode.mScanData->RelativeSensorPose()),
                                     localMapNode.mGlobalPose, localMapNode.mLocalMapId, scanNode.mNodeId, cov);
                continue;
            }
            /* :110-135: the caller's final matcher on the coarse estimate, on the CPU */
            const auto& refScanNode = query.mReferenceScanNode;
            Assert(refScanNode.mLocalMapId == localMapNode.mLocalMapId);
            const Point2D<double> localMapCenterPos { refScanNode.mLocalPose.mX, refScanNode.mLocalPose.mY };
            const RobotPose2D<double> coarsePose = MoveBackward(bestSensorPose, scanNode.mScanData->RelativeSensorPose());
            const ScanMatchingQuery finalQuery { localMap.mMap, localMapCenterPos, scanNode.mScanData, coarsePose };
            const ScanMatchingSummary finalSummary = this->mFinalScanMatcher->OptimizePose(finalQuery);
            Assert(finalSummary.mPoseFound);
            results.emplace_back(finalSummary.mEstimatedPose, localMapNode.mGlobalPose,
                                 localMapNode.mLocalMapId, scanNode.mNodeId, finalSummary.mEstimatedCovariance);
        }
        return results;
    }

    inline const std::vector<csm_result>& LastResults() const { return this->mLastResults; }

private:
    std::shared_ptr<ScanMatcherBranchBoundGPU> mScanMatcher;
    std::shared_ptr<ScanMatcher> mFinalScanMatcher;
    const double mScoreThreshold, mKnownRateThreshold;
    std::set<std::int64_t> mCachedMaps;
    std::vector<csm_result> mLastResults;
    bool mDeviceRefiner = false;
    csm_refine_params mRefine {};
};

/* The selection a maintainer adds next to scan_matcher_factory.cpp:194-217 ("ScanMatcherType") and
 * loop_detector_factory.cpp:197-212 ("LoopDetectorType"): the CPU type strings with a "GPU" suffix,
 * the same configuration keys (LowResolutionMapWinSize / NodeHeightMax / SearchRange{X,Y,Theta} /
 * SearchStep{X,Y,Theta}), plus an optional "Device". */
inline std::shared_ptr<ScanMatcher> CreateScanMatcherGPU(
    const std::string& scanMatcherType, const std::string& name, const CostFuncPtr& costFunc,
    const int lowResolutionOrNodeHeightMax, const double rangeX, const double rangeY, const double rangeTheta,
    const double stepX = 0.0, const double stepY = 0.0, const double stepTheta = 0.0, const int device = 0)
{
    if (scanMatcherType == "RealTimeCorrelativeGPU")
        return std::make_shared<ScanMatcherCorrelativeGPU>(name, costFunc, lowResolutionOrNodeHeightMax,
                                                           rangeX, rangeY, rangeTheta, device);
    if (scanMatcherType == "BranchBoundGPU")
        return std::make_shared<ScanMatcherBranchBoundGPU>(name, costFunc, lowResolutionOrNodeHeightMax,
                                                           rangeX, rangeY, rangeTheta, device);
    if (scanMatcherType == "GridSearchGPU")
        return std::make_shared<ScanMatcherGridSearchGPU>(name, costFunc, rangeX, rangeY, rangeTheta,
                                                          stepX, stepY, stepTheta, device);
    return nullptr;
}

} /* namespace Mapping */
} /* namespace MyLidarGraphSlam */
