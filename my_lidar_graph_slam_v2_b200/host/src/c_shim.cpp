/* c_shim.cpp -- extern "C" test entry points over the C++ plugin mirror, so that
 * the parity tests (Python, ctypes) can drive the adapter classes exactly the
 * way the reference drives its matchers / detectors. Not part of the C ABI. */
#include <csignal>
#include <cstring>
#include <execinfo.h>
#include <fstream>
#include <sstream>
#include <memory>
#include <string>
#include <unistd.h>
#include <vector>

#include <cstdio>
#include "csm_host/loop_detector.hpp"
#include "csm_host/loop_searcher.hpp"
#include "csm_host/map_builder.hpp"
#include "csm_host/slam_pipeline.hpp"

using namespace csm_host;

extern "C" {

struct csm_host_summary
{
    int32_t found, best_x, best_y, best_t;
    int64_t sum_value;
    int32_t n_known, flags;
    double  score;
    double  norm_cost;
    double  est_pose[3];
    double  cov[9];
};

static void Export(const ScanMatchingSummary& s, csm_host_summary* out)
{
    out->found = s.pose_found ? 1 : 0;
    out->best_x = s.best_x; out->best_y = s.best_y; out->best_t = s.best_theta;
    out->sum_value = s.sum_value; out->n_known = s.n_known; out->flags = s.flags;
    out->score = s.normalized_score;
    out->norm_cost = s.normalized_cost;
    out->est_pose[0] = s.estimated_pose.x; out->est_pose[1] = s.estimated_pose.y;
    out->est_pose[2] = s.estimated_pose.theta;
    std::memcpy(out->cov, s.estimated_covariance.data(), sizeof(double) * 9);
}

/* debugging aid: print the native call stack when the process takes a SIGSEGV */
static void BacktraceOnSegv(int sig)
{
    void* frames[64];
    const int n = backtrace(frames, 64);
    backtrace_symbols_fd(frames, n, 2);
    signal(sig, SIG_DFL);
    raise(sig);
}
void csm_host_install_backtrace() { signal(SIGSEGV, BacktraceOnSegv); }

void* csm_host_context_create(int device) { return new DeviceContextPtr(std::make_shared<DeviceContext>(device)); }
void csm_host_context_destroy(void* ctx) { delete static_cast<DeviceContextPtr*>(ctx); }
/* the csm_handle (C ABI) behind a context: timing options, direct C-ABI calls in tests */
/* Matchers created on this context take cost / covariance from the device (csm_set_epilogue) */
void csm_host_context_set_device_epilogue(void* ctx, int on)
{
    (*static_cast<DeviceContextPtr*>(ctx))->SetDeviceEpilogue(on != 0);
}

/* ... and run the reference's final matcher (linear solver) on the found pose on the device; iterations <= 0 = off */
void csm_host_context_set_device_final_matcher(void* ctx, int iterations_max, double convergence_threshold,
                                               double initial_lambda, double covariance_scale)
{
    (*static_cast<DeviceContextPtr*>(ctx))->SetDeviceFinalMatcher(iterations_max, convergence_threshold,
                                                                  initial_lambda, covariance_scale);
}

void* csm_host_context_handle(void* ctx) { return (*static_cast<DeviceContextPtr*>(ctx))->Handle(); }

static GridMapView View(const uint16_t* values, int rows, int cols, double res, double ox, double oy, int64_t id)
{
    GridMapView v;
    v.values = values; v.rows = rows; v.cols = cols; v.resolution = res;
    v.offset_x = ox; v.offset_y = oy; v.map_id = id;
    return v;
}

static ScanDataPtr Scan(const double* angles, const double* ranges, int n, const double rel[3])
{
    auto s = std::make_shared<ScanData>();
    s->angles.assign(angles, angles + n);
    s->ranges.assign(ranges, ranges + n);
    s->relative_sensor_pose = Pose2D { rel[0], rel[1], rel[2] };
    return s;
}

/* CPU epilogue alone: cost / N and covariance at a given sensor pose (no device needed) */
int csm_host_cost(const uint16_t* values, int rows, int cols, double res, double off_x, double off_y,
                  const double* angles, const double* ranges, int n, const double sensor_pose[3],
                  double covariance_scale, double* norm_cost, double cov[9])
{
    const CostSquareError cost(covariance_scale);
    const GridMapView map = View(values, rows, cols, res, off_x, off_y, -1);
    const double rel[3] = { 0.0, 0.0, 0.0 };
    const ScanDataPtr scan = Scan(angles, ranges, n, rel);
    const Pose2D pose { sensor_pose[0], sensor_pose[1], sensor_pose[2] };
    *norm_cost = cost.Cost(map, *scan, pose) / static_cast<double>(n);
    const std::array<double, 9> c = cost.ComputeCovariance(map, *scan, pose);
    std::memcpy(cov, c.data(), sizeof(double) * 9);
    return 0;
}

/* kind: 0 = RealTimeCorrelative, 1 = BranchBound, 2 = GridSearch.
 * iparam: lowResolution / nodeHeightMax; range[3]; step[3] (grid search only). */
static int MatchView(void* ctx, int kind, const GridMapView& map, const double* angles, const double* ranges, int n,
                     const double init_pose[3], const double rel_pose[3], int iparam,
                     const double range[3], const double step[3], double score_thr, double known_thr,
                     double covariance_scale, csm_host_summary* out);

int csm_host_match(void* ctx, int kind, const uint16_t* values, int rows, int cols, double res,
                   double off_x, double off_y, const double* angles, const double* ranges, int n,
                   const double init_pose[3], const double rel_pose[3], int iparam,
                   const double range[3], const double step[3], double score_thr, double known_thr,
                   double covariance_scale, csm_host_summary* out)
{
    return MatchView(ctx, kind, View(values, rows, cols, res, off_x, off_y, -1), angles, ranges, n, init_pose,
                     rel_pose, iparam, range, step, score_thr, known_thr, covariance_scale, out);
}

/* The same with the map in block-sparse form (the reference's storage): n_blocks blocks of
 * (1 << log2bs)^2 cells and their positions */
int csm_host_match_blocks(void* ctx, int kind, const uint16_t* blocks, const int32_t* block_index, int n_blocks,
                          int log2bs, int rows, int cols, double res, double off_x, double off_y,
                          const double* angles, const double* ranges, int n,
                          const double init_pose[3], const double rel_pose[3], int iparam,
                          const double range[3], const double step[3], double score_thr, double known_thr,
                          double covariance_scale, csm_host_summary* out)
{
    GridMapView map = View(nullptr, rows, cols, res, off_x, off_y, -1);
    map.blocks = blocks; map.block_index = block_index; map.n_blocks = n_blocks; map.log2_block_size = log2bs;
    return MatchView(ctx, kind, map, angles, ranges, n, init_pose, rel_pose, iparam, range, step, score_thr,
                     known_thr, covariance_scale, out);
}

static int MatchView(void* ctx, int kind, const GridMapView& map, const double* angles, const double* ranges, int n,
                     const double init_pose[3], const double rel_pose[3], int iparam,
                     const double range[3], const double step[3], double score_thr, double known_thr,
                     double covariance_scale, csm_host_summary* out)
{
    const DeviceContextPtr& c = *static_cast<DeviceContextPtr*>(ctx);
    const auto cost = std::make_shared<CostSquareError>(covariance_scale);
    const ScanDataPtr scan = Scan(angles, ranges, n, rel_pose);
    const Pose2D init { init_pose[0], init_pose[1], init_pose[2] };
    ScanMatchingSummary s;
    if (kind == 0) {
        ScanMatcherCorrelative m("RealTimeCorrelativeGPU", cost, iparam, range[0], range[1], range[2], c);
        s = m.OptimizePose(map, scan, init, score_thr, known_thr);
    } else if (kind == 1) {
        ScanMatcherBranchBound m("BranchBoundGPU", cost, iparam, range[0], range[1], range[2], c);
        s = m.OptimizePose(map, scan, init, score_thr, known_thr);
    } else if (kind == 2) {
        ScanMatcherGridSearch m("GridSearchGPU", cost, range[0], range[1], range[2],
                                step[0], step[1], step[2], c);
        s = m.OptimizePose(map, scan, init, score_thr, known_thr);
    } else {
        return -1;
    }
    Export(s, out);
    return 0;
}

/* Metric ids and values a detector + its matcher observed, as "id=v,v,...;id=...": what the
 * reference's MetricManager would hold */
static void DumpMetrics(const MetricRecorder& rec, char* buf, int cap)
{
    if (buf == nullptr || cap <= 0)
        return;
    std::string text;
    for (const auto& kv : rec.Values()) {
        text += kv.first + "=";
        for (std::size_t i = 0; i < kv.second.size(); ++i)
            text += (i ? "," : "") + std::to_string(kv.second[i]);
        text += ";";
    }
    std::strncpy(buf, text.c_str(), static_cast<std::size_t>(cap - 1));
    buf[cap - 1] = 0;
}

/* matchers the Correlative / GridSearch detectors of csm_host_loop_detect_kind run at once (1 = the plain loop) */
static int gDetectConcurrency = 1;
void csm_host_set_detect_concurrency(int n) { gDetectConcurrency = n < 1 ? 1 : n; }

/* LoopDetector{Correlative, BranchBound, GridSearch}::Detect (kind 0 / 1 / 2 as in csm_host_match)
 * over n_queries maps and one shared scan. values: n_queries grids of rows x cols, consecutive.
 * out[q].found = 0 when no result. metrics: optional text buffer (see DumpMetrics). */
int csm_host_loop_detect_kind(void* ctx, int kind, int n_queries, const uint16_t* values, int rows, int cols,
                              double res, const double* off_x, const double* off_y, const int64_t* map_ids,
                              const double* map_poses, const double* scan_poses,
                              const double* angles, const double* ranges, int n,
                              int iparam, const double range[3], const double step[3],
                              double score_thr, double known_thr,
                              double covariance_scale, csm_host_summary* out, char* metrics, int metrics_cap)
{
    const DeviceContextPtr& c = *static_cast<DeviceContextPtr*>(ctx);
    const auto cost = std::make_shared<CostSquareError>(covariance_scale);
    const auto rec = std::make_shared<MetricRecorder>();
    std::shared_ptr<ScanMatcherBranchBound> bb;
    std::unique_ptr<LoopDetector> det;
    /* the extra contexts live as long as the process: creating one costs milliseconds */
    static std::vector<DeviceContextPtr> pool;
    std::vector<DeviceContextPtr> extra_ctx;
    for (int k = 1; k < gDetectConcurrency && kind != 1; ++k) {
        if (static_cast<int>(pool.size()) < k || pool[k - 1]->Device() != c->Device()) {
            pool.resize(std::max<std::size_t>(pool.size(), k));
            pool[k - 1] = std::make_shared<DeviceContext>(c->Device());
        }
        extra_ctx.push_back(pool[k - 1]);
    }
    if (kind == 0) {
        auto m = std::make_shared<ScanMatcherCorrelative>("LoopRTGPU", cost, iparam, range[0], range[1], range[2], c);
        m->SetMetricSink(rec);
        auto* d = new LoopDetectorCorrelative("LoopDetectorCorrelativeGPU", m, FinalMatcher(), score_thr, known_thr);
        std::vector<std::shared_ptr<ScanMatcherCorrelative>> extra;
        for (const DeviceContextPtr& e : extra_ctx)
            extra.push_back(std::make_shared<ScanMatcherCorrelative>("LoopRTGPU", cost, iparam, range[0], range[1], range[2], e));
        d->SetConcurrentMatchers(extra);
        det.reset(d);
    } else if (kind == 1) {
        bb = std::make_shared<ScanMatcherBranchBound>("LoopBBGPU", cost, iparam, range[0], range[1], range[2], c);
        bb->SetMetricSink(rec);
        det.reset(new LoopDetectorBranchBound("LoopDetectorBranchBoundGPU", bb, FinalMatcher(), score_thr, known_thr));
    } else if (kind == 2) {
        auto m = std::make_shared<ScanMatcherGridSearch>("LoopGridGPU", cost, range[0], range[1], range[2],
                                                         step[0], step[1], step[2], c);
        m->SetMetricSink(rec);
        auto* d = new LoopDetectorGridSearch("LoopDetectorGridSearchGPU", m, FinalMatcher(), score_thr, known_thr);
        std::vector<std::shared_ptr<ScanMatcherGridSearch>> extra;
        for (const DeviceContextPtr& e : extra_ctx)
            extra.push_back(std::make_shared<ScanMatcherGridSearch>("LoopGridGPU", cost, range[0], range[1], range[2],
                                                                    step[0], step[1], step[2], e));
        d->SetConcurrentMatchers(extra);
        det.reset(d);
    } else {
        return -1;
    }
    det->SetMetricSink(rec);
    const double rel[3] = { 0.0, 0.0, 0.0 };
    const ScanDataPtr scan = Scan(angles, ranges, n, rel);
    std::vector<LoopDetectionQuery> queries(n_queries);
    const size_t cells = static_cast<size_t>(rows) * cols;
    for (int q = 0; q < n_queries; ++q) {
        LoopDetectionQuery& lq = queries[q];
        lq.scan = scan;
        lq.scan_id = 0;
        lq.scan_node_id = q;
        lq.scan_global_pose = Pose2D { scan_poses[3 * q], scan_poses[3 * q + 1], scan_poses[3 * q + 2] };
        lq.local_map = View(values + q * cells, rows, cols, res, off_x[q], off_y[q], map_ids[q]);
        lq.local_map_global_pose = Pose2D { map_poses[3 * q], map_poses[3 * q + 1], map_poses[3 * q + 2] };
    }
    const std::vector<LoopDetectionResult> results = det->Detect(queries);
    for (int q = 0; q < n_queries; ++q)
        std::memset(&out[q], 0, sizeof(csm_host_summary));
    for (const LoopDetectionResult& r : results) {
        csm_host_summary& o = out[r.scan_node_id];
        o.found = 1;
        if (kind == 1) {
            const csm_result& d = static_cast<LoopDetectorBranchBound*>(det.get())->LastResults()[r.query_index];
            o.best_x = d.best_x; o.best_y = d.best_y; o.best_t = d.best_t;
            o.sum_value = d.sum_value; o.n_known = d.n_known; o.flags = d.flags;
        }
        o.score = r.normalized_score;
        o.est_pose[0] = r.relative_pose.x; o.est_pose[1] = r.relative_pose.y; o.est_pose[2] = r.relative_pose.theta;
        std::memcpy(o.cov, r.estimated_covariance.data(), sizeof(double) * 9);
    }
    for (int q = 0; q < n_queries; ++q) {
        csm_release_grid(c->Handle(), map_ids[q]);
        for (const DeviceContextPtr& e : extra_ctx)
            csm_release_grid(e->Handle(), map_ids[q]);
    }
    DumpMetrics(*rec, metrics, metrics_cap);
    det.reset();             /* before the contexts its extra matchers run on */
    return 0;
}

int csm_host_loop_detect(void* ctx, int n_queries, const uint16_t* values, int rows, int cols, double res,
                         const double* off_x, const double* off_y, const int64_t* map_ids,
                         const double* map_poses, const double* scan_poses,
                         const double* angles, const double* ranges, int n,
                         int hmax, const double range[3], double score_thr, double known_thr,
                         double covariance_scale, csm_host_summary* out)
{
    return csm_host_loop_detect_kind(ctx, 1, n_queries, values, rows, cols, res, off_x, off_y, map_ids, map_poses,
                                     scan_poses, angles, ranges, n, hmax, range, range, score_thr, known_thr,
                                     covariance_scale, out, nullptr, 0);
}

/* ScanMatcherLinearSolver::OptimizePose on the CPU (no device): `lambda` is the solver's damping
 * state, read and written back so that a sequence of calls behaves like one solver instance */
int csm_host_refine(const uint16_t* values, int rows, int cols, double res, double off_x, double off_y,
                    const double* angles, const double* ranges, int n, const double init_pose[3],
                    const double rel_pose[3], int iterations_max, double convergence_threshold,
                    double* lambda, double covariance_scale, csm_host_summary* out)
{
    const auto cost = std::make_shared<CostSquareError>(covariance_scale);
    const GridMapView map = View(values, rows, cols, res, off_x, off_y, -1);
    const ScanDataPtr scan = Scan(angles, ranges, n, rel_pose);
    ScanMatcherLinearSolver solver("LinearSolver", iterations_max, convergence_threshold, *lambda, cost);
    const ScanMatchingSummary s = solver.OptimizePose(
        ScanMatchingQuery { map, scan, Pose2D { init_pose[0], init_pose[1], init_pose[2] } });
    *lambda = solver.Lambda();
    Export(s, out);
    out->best_t = s.n_processed;      /* iterations */
    return 0;
}

/* LoopSearcherNearest::Search (CPU, no device) on a pose-graph summary: scan nodes (ids ascending, 3 doubles
 * of global pose each) and local maps (ids ascending, scan-node id range, finished flag). Writes up to `cap`
 * candidates as (query scan node, reference scan node, reference local map) id triples and their squared node
 * distances; returns how many. */
int csm_host_loop_search(int n_scans, const int* scan_ids, const double* scan_poses,
                         int n_maps, const int* map_ids, const int* map_scan_min, const int* map_scan_max,
                         const int* map_finished, double accum_travel_dist, int last_finished_scan_id,
                         int last_finished_map_id, double travel_dist_threshold, double node_dist_threshold,
                         int num_of_candidate_nodes, int* out_ids, double* out_dist_sq, int cap)
{
    LoopSearchHint hint;
    for (int i = 0; i < n_scans; ++i)
        hint.scan_nodes.push_back(ScanNodeData { scan_ids[i], Pose2D { scan_poses[3 * i], scan_poses[3 * i + 1],
                                                                       scan_poses[3 * i + 2] } });
    for (int i = 0; i < n_maps; ++i)
        hint.local_map_nodes.push_back(LocalMapData { map_ids[i], map_scan_min[i], map_scan_max[i],
                                                      map_finished[i] != 0 });
    hint.accum_travel_dist = accum_travel_dist;
    hint.last_finished_scan_id = last_finished_scan_id;
    hint.last_finished_map_id = last_finished_map_id;
    LoopSearcherNearest searcher(travel_dist_threshold, node_dist_threshold, num_of_candidate_nodes);
    const std::vector<LoopCandidate> c = searcher.Search(hint);
    const int n = std::min(static_cast<int>(c.size()), cap);
    for (int i = 0; i < n; ++i) {
        out_ids[3 * i] = c[i].query_scan_node_id;
        out_ids[3 * i + 1] = c[i].reference_scan_node_id;
        out_ids[3 * i + 2] = c[i].reference_local_map_id;
        out_dist_sq[i] = searcher.LastNodeDistances()[i];
    }
    return n;
}

/* ScanMatcherHillClimbing::OptimizePose on the CPU (no device). out->best_t = iterations,
 * out->best_x = refinements (step halvings). greedy == null: square-error cost (covariance_scale);
 * else the greedy-endpoint cost with greedy = { MapResolution, HitAndMissedDist, OccupancyThreshold,
 * KernelSize, ScalingFactor, StandardDeviation }. */
int csm_host_hill_climb(const uint16_t* values, int rows, int cols, double res, double off_x, double off_y,
                        const double* angles, const double* ranges, int n, const double init_pose[3],
                        const double rel_pose[3], double linear_step, double angular_step, int max_iterations,
                        int max_num_of_refinements, double covariance_scale, const double* greedy,
                        csm_host_summary* out)
{
    std::shared_ptr<CostFunction> cost;
    if (greedy != nullptr)
        cost = std::make_shared<CostGreedyEndpoint>(greedy[0], greedy[1], greedy[2], static_cast<int>(greedy[3]),
                                                    greedy[4], greedy[5]);
    else
        cost = std::make_shared<CostSquareError>(covariance_scale);
    const GridMapView map = View(values, rows, cols, res, off_x, off_y, -1);
    const ScanDataPtr scan = Scan(angles, ranges, n, rel_pose);
    ScanMatcherHillClimbing matcher("HillClimbing", linear_step, angular_step, max_iterations,
                                    max_num_of_refinements, cost);
    const ScanMatchingSummary s = matcher.OptimizePose(
        ScanMatchingQuery { map, scan, Pose2D { init_pose[0], init_pose[1], init_pose[2] } });
    Export(s, out);
    out->best_t = s.n_processed;
    out->best_x = matcher.LastNumOfRefinements();
    return 0;
}

/* ---- persistent loop detector (bench.py e2e path) ------------------------------ */
struct HostLoopDet
{
    DeviceContextPtr ctx;
    std::shared_ptr<ScanMatcherBranchBound> matcher;
    std::shared_ptr<ScanMatcherLinearSolver> refiner;
    std::unique_ptr<LoopDetectorBranchBound> det;
    double score_thr = 0.0, known_thr = 0.0;
};

void* csm_host_loopdet_create(void* ctx, int hmax, const double range[3], double score_thr,
                              double known_thr, double covariance_scale)
{
    auto* d = new HostLoopDet;
    d->ctx = *static_cast<DeviceContextPtr*>(ctx);
    const auto cost = std::make_shared<CostSquareError>(covariance_scale);
    d->matcher = std::make_shared<ScanMatcherBranchBound>("LoopBBGPU", cost, hmax, range[0], range[1],
                                                          range[2], d->ctx);
    d->det.reset(new LoopDetectorBranchBound("LoopDetectorBranchBoundGPU", d->matcher, FinalMatcher(),
                                             score_thr, known_thr));
    d->score_thr = score_thr; d->known_thr = known_thr;
    return d;
}

/* Refine every detected loop with a ScanMatcherLinearSolver like the reference's default
 * configuration ("FinalScanMatcherType": "LinearSolver") */
void csm_host_loopdet_use_linear_solver(void* det, int iterations_max, double convergence_threshold,
                                        double initial_lambda, double covariance_scale)
{
    auto* d = static_cast<HostLoopDet*>(det);
    d->refiner = std::make_shared<ScanMatcherLinearSolver>(
        "LoopDetector.FinalScanMatcherLinearSolver", iterations_max, convergence_threshold, initial_lambda,
        std::make_shared<CostSquareError>(covariance_scale));
    d->det.reset(new LoopDetectorBranchBound("LoopDetectorBranchBoundGPU", d->matcher,
                                             MakeLinearSolverFinalMatcher(d->refiner), d->score_thr, d->known_thr));
}

/* The same refinement on the device, batched behind the search (csm_set_refiner) */
void csm_host_loopdet_use_device_refiner(void* det, int iterations_max, double convergence_threshold,
                                         double initial_lambda, double covariance_scale)
{
    auto* d = static_cast<HostLoopDet*>(det);
    d->refiner.reset();
    d->det->UseDeviceRefiner(iterations_max, convergence_threshold, initial_lambda, covariance_scale);
}

/* n pipeline lanes in all (n - 1 additional device contexts on the detector's device) */
void csm_host_loopdet_set_lanes(void* det, int n)
{
    auto* d = static_cast<HostLoopDet*>(det);
    std::vector<DeviceContextPtr> extra;
    for (int i = 1; i < n; ++i)
        extra.push_back(std::make_shared<DeviceContext>(d->ctx->Device()));
    d->det->SetPipelineLanes(extra);
}

/* Packed best word of the last Detect over all lanes (LoopDetectorBranchBound::BestWord) */
unsigned long long csm_host_loopdet_best_word(void* det)
{
    return static_cast<HostLoopDet*>(det)->det->BestWord();
}

void csm_host_loopdet_destroy(void* det) { delete static_cast<HostLoopDet*>(det); }

void csm_host_loopdet_configure(void* det, int chunk_size, int coarse_covariance, int query_index_base)
{
    auto* d = static_cast<HostLoopDet*>(det);
    /* bits 0..11 search batch, 12..15 tail batch / 16, 16.. upload group */
    d->det->SetChunkSize(chunk_size & 0xfff);
    d->det->SetTailChunk(((chunk_size >> 12) & 0xf) * 16);
    if ((chunk_size >> 16) > 0)
        d->det->SetUploadChunk(chunk_size >> 16);
    d->det->SetCoarseCovariance(coarse_covariance != 0);
    d->det->SetQueryIndexBase(query_index_base);
}

void csm_host_loopdet_clear_cache(void* det) { static_cast<HostLoopDet*>(det)->det->ClearCache(); }

void* csm_host_loopdet_handle(void* det) { return static_cast<HostLoopDet*>(det)->ctx->Handle(); }

/* The reference's storage of a batch of maps, rebuilt from a contiguous block list: every allocated
 * block in its own heap allocation (grid_map.cpp:522-535), in the order the blocks were first written
 * (here: map by map). What the adapter sees when it is handed GridMap objects. */
struct HeapMaps
{
    std::vector<std::unique_ptr<uint16_t[]>> blocks;
    std::vector<const uint16_t*> ptrs;
    std::vector<int32_t> index, counts;
    std::vector<size_t> first;
    int log2bs = 4;
};

void* csm_host_heap_maps_create(const uint16_t* blocks, const int32_t* block_index, const int32_t* block_count,
                                int n_maps, int log2bs)
{
    auto* hm = new HeapMaps;
    hm->log2bs = log2bs;
    const size_t cells = static_cast<size_t>(1) << (2 * log2bs);
    size_t total = 0;
    for (int m = 0; m < n_maps; ++m) { hm->first.push_back(total); total += static_cast<size_t>(block_count[m]); }
    hm->counts.assign(block_count, block_count + n_maps);
    hm->index.assign(block_index, block_index + total);
    hm->blocks.reserve(total);
    hm->ptrs.reserve(total);
    for (size_t b = 0; b < total; ++b) {
        hm->blocks.emplace_back(new uint16_t[cells]);
        std::memcpy(hm->blocks.back().get(), blocks + b * cells, cells * sizeof(uint16_t));
        hm->ptrs.push_back(hm->blocks.back().get());
    }
    return hm;
}

void csm_host_heap_maps_destroy(void* p) { delete static_cast<HeapMaps*>(p); }

/* n_queries queries, query q on map q (dense `values`, contiguous block-sparse `blocks`, or heap blocks)
 * and one shared scan */
static std::vector<LoopDetectionQuery> BuildQueries(
    int n_queries, const uint16_t* values, const uint16_t* blocks, const int32_t* block_index,
    const int32_t* block_count, const HeapMaps* heap, int log2bs, int rows, int cols, double res,
    const double* off_x, const double* off_y, const int64_t* map_ids, const double* map_poses,
    const double* scan_poses, const double* angles, const double* ranges, int n)
{
    const double rel[3] = { 0.0, 0.0, 0.0 };
    const ScanDataPtr scan = Scan(angles, ranges, n, rel);
    std::vector<LoopDetectionQuery> queries(n_queries);
    const size_t cells = static_cast<size_t>(rows) * cols;
    size_t nblk = 0;
    for (int q = 0; q < n_queries; ++q) {
        LoopDetectionQuery& lq = queries[q];
        lq.scan = scan;
        lq.scan_id = 0;
        lq.scan_node_id = q;
        lq.scan_global_pose = Pose2D { scan_poses[3 * q], scan_poses[3 * q + 1], scan_poses[3 * q + 2] };
        lq.local_map = View(values ? values + q * cells : nullptr, rows, cols, res, off_x[q], off_y[q], map_ids[q]);
        if (heap != nullptr) {
            lq.local_map.block_ptrs = heap->ptrs.data() + heap->first[q];
            lq.local_map.block_index = heap->index.data() + heap->first[q];
            lq.local_map.n_blocks = heap->counts[q];
            lq.local_map.log2_block_size = heap->log2bs;
        } else if (blocks != nullptr) {
            lq.local_map.blocks = blocks + (nblk << (2 * log2bs));
            lq.local_map.block_index = block_index + nblk;
            lq.local_map.n_blocks = block_count[q];
            lq.local_map.log2_block_size = log2bs;
            nblk += static_cast<size_t>(block_count[q]);
        }
        lq.local_map_global_pose = Pose2D { map_poses[3 * q], map_poses[3 * q + 1], map_poses[3 * q + 2] };
    }
    return queries;
}

static int ExportResults(const std::vector<LoopDetectionResult>& results, const std::vector<csm_result>& last,
                         int n_queries, csm_host_summary* out)
{
    for (int q = 0; q < n_queries; ++q)
        std::memset(&out[q], 0, sizeof(csm_host_summary));
    for (const LoopDetectionResult& r : results) {
        csm_host_summary& o = out[r.scan_node_id];
        const csm_result& dr = last[r.query_index];
        o.found = 1;
        o.best_x = dr.best_x; o.best_y = dr.best_y; o.best_t = dr.best_t;
        o.sum_value = dr.sum_value; o.n_known = dr.n_known; o.flags = dr.flags;
        o.score = dr.normalized_score;
        o.est_pose[0] = r.relative_pose.x; o.est_pose[1] = r.relative_pose.y; o.est_pose[2] = r.relative_pose.theta;
        std::memcpy(o.cov, r.estimated_covariance.data(), sizeof(double) * 9);
    }
    return static_cast<int>(results.size());
}

/* Detect over n_queries block-sparse maps (map q owns block_count[q] consecutive
 * blocks) and one shared scan. values == dense alternative when blocks is null. */
int csm_host_loopdet_detect(void* det, int n_queries, const uint16_t* values,
                            const uint16_t* blocks, const int32_t* block_index, const int32_t* block_count,
                            int log2bs, int rows, int cols, double res,
                            const double* off_x, const double* off_y, const int64_t* map_ids,
                            const double* map_poses, const double* scan_poses,
                            const double* angles, const double* ranges, int n, csm_host_summary* out)
{
    auto* d = static_cast<HostLoopDet*>(det);
    const std::vector<LoopDetectionQuery> queries = BuildQueries(
        n_queries, values, blocks, block_index, block_count, nullptr, log2bs, rows, cols, res, off_x, off_y,
        map_ids, map_poses, scan_poses, angles, ranges, n);
    return ExportResults(d->det->Detect(queries), d->det->LastResults(), n_queries, out);
}

/* The same from maps whose blocks are separate heap allocations (csm_host_heap_maps_create) */
int csm_host_loopdet_detect_heap(void* det, int n_queries, const void* heap_maps, int rows, int cols, double res,
                                 const double* off_x, const double* off_y, const int64_t* map_ids,
                                 const double* map_poses, const double* scan_poses,
                                 const double* angles, const double* ranges, int n, csm_host_summary* out)
{
    auto* d = static_cast<HostLoopDet*>(det);
    const std::vector<LoopDetectionQuery> queries = BuildQueries(
        n_queries, nullptr, nullptr, nullptr, nullptr, static_cast<const HeapMaps*>(heap_maps), 4, rows, cols, res,
        off_x, off_y, map_ids, map_poses, scan_poses, angles, ranges, n);
    return ExportResults(d->det->Detect(queries), d->det->LastResults(), n_queries, out);
}

void csm_host_loopdet_set_first_group_divisor(void* det, int n) { static_cast<HostLoopDet*>(det)->det->SetFirstGroupDivisor(n); }
void csm_host_loopdet_set_gather_threads(void* det, int n) { static_cast<HostLoopDet*>(det)->det->SetGatherThreads(n); }
int csm_host_loopdet_capacity_retries(void* det) { return static_cast<HostLoopDet*>(det)->det->NumOfCapacityRetries(); }

/* ---- the detector over several GPUs of one process ----------------------------------------- */
struct HostMultiDet
{
    std::vector<DeviceContextPtr> ctx;
    std::vector<std::shared_ptr<ScanMatcherBranchBound>> matcher;
    std::vector<std::shared_ptr<LoopDetectorBranchBound>> shard;
    std::unique_ptr<LoopDetectorBranchBoundMultiGPU> det;
};

void* csm_host_multidet_create(int n_gpus, int hmax, const double range[3], double score_thr, double known_thr,
                               double covariance_scale, int lanes, int refine_iterations,
                               double convergence_threshold, double initial_lambda)
{
    auto* d = new HostMultiDet;
    for (int g = 0; g < n_gpus; ++g) {
        d->ctx.push_back(std::make_shared<DeviceContext>(g));
        const auto cost = std::make_shared<CostSquareError>(covariance_scale);
        d->matcher.push_back(std::make_shared<ScanMatcherBranchBound>("LoopBBGPU", cost, hmax, range[0], range[1],
                                                                      range[2], d->ctx[g]));
        auto shard = std::make_shared<LoopDetectorBranchBound>("LoopDetectorBranchBoundGPU", d->matcher[g],
                                                               FinalMatcher(), score_thr, known_thr);
        shard->SetCoarseCovariance(false);
        if (refine_iterations > 0)
            shard->UseDeviceRefiner(refine_iterations, convergence_threshold, initial_lambda, covariance_scale);
        std::vector<DeviceContextPtr> extra;
        for (int l = 1; l < lanes; ++l)
            extra.push_back(std::make_shared<DeviceContext>(g));
        if (!extra.empty())
            shard->SetPipelineLanes(extra);
        d->shard.push_back(shard);
    }
    d->det.reset(new LoopDetectorBranchBoundMultiGPU("LoopDetectorBranchBoundMultiGPU", d->shard, d->ctx));
    return d;
}

void csm_host_multidet_destroy(void* det) { delete static_cast<HostMultiDet*>(det); }
void csm_host_multidet_use_nccl(void* det) { static_cast<HostMultiDet*>(det)->det->UseNcclExchange(); }
void csm_host_multidet_clear_cache(void* det)
{
    for (auto& s : static_cast<HostMultiDet*>(det)->shard) s->ClearCache();
}
void csm_host_multidet_configure(void* det, int chunk_size, int upload_chunk, int gather_threads)
{
    for (auto& s : static_cast<HostMultiDet*>(det)->shard) {
        s->SetChunkSize(chunk_size);
        s->SetUploadChunk(upload_chunk);
        if (gather_threads > 0) s->SetGatherThreads(gather_threads);
    }
}
unsigned long long csm_host_multidet_best_word(void* det) { return static_cast<HostMultiDet*>(det)->det->BestWord(); }
void csm_host_multidet_shard_sizes(void* det, int* out)
{
    const std::vector<int>& s = static_cast<HostMultiDet*>(det)->det->LastShardSizes();
    for (size_t g = 0; g < s.size(); ++g) out[g] = s[g];
}

int csm_host_multidet_detect(void* det, int n_queries, const uint16_t* blocks, const int32_t* block_index,
                             const int32_t* block_count, const void* heap_maps, int log2bs, int rows, int cols,
                             double res, const double* off_x, const double* off_y, const int64_t* map_ids,
                             const double* map_poses, const double* scan_poses,
                             const double* angles, const double* ranges, int n, csm_host_summary* out)
{
    auto* d = static_cast<HostMultiDet*>(det);
    const std::vector<LoopDetectionQuery> queries = BuildQueries(
        n_queries, nullptr, blocks, block_index, block_count, static_cast<const HeapMaps*>(heap_maps), log2bs,
        rows, cols, res, off_x, off_y, map_ids, map_poses, scan_poses, angles, ranges, n);
    return ExportResults(d->det->Detect(queries), d->det->LastResults(), n_queries, out);
}

/* GridMapBuilderGPU::UpdateTable: the 65536-entry table the device applies per update (no device needed) */
void csm_host_map_update_table(double odds, int reference_table_end, uint16_t* out)
{
    const std::vector<std::uint16_t> t = GridMapBuilderGPU::UpdateTable(odds, reference_table_end != 0);
    std::copy(t.begin(), t.end(), out);
}

/* ---- map construction on the device: GridMapBuilderGPU ------------------------------------------ */
struct HostMapBuilder
{
    DeviceContextPtr ctx;
    std::unique_ptr<GridMapBuilderGPU> builder;
    std::vector<ScanNodeView> nodes;
};

void* csm_host_mapbuilder_create(void* ctx, double resolution, int patch_size, int scans_for_latest_map,
                                 double usable_range_min, double usable_range_max, double prob_hit, double prob_miss)
{
    auto* b = new HostMapBuilder;
    b->ctx = *static_cast<DeviceContextPtr*>(ctx);
    b->builder.reset(new GridMapBuilderGPU(b->ctx, resolution, patch_size, scans_for_latest_map, usable_range_min,
                                           usable_range_max, prob_hit, prob_miss));
    return b;
}

void csm_host_mapbuilder_destroy(void* p) { delete static_cast<HostMapBuilder*>(p); }
/* 0: every hit point with libm like the reference; 1 (default): rotated polar points behind a guard band */
/* test knob: the guard band of the fast hit points, in cells (0.5 and above: every beam is re-evaluated) */
void csm_host_mapbuilder_set_guard_band(double cells) { DeviceGridMap::GuardBand() = cells; }
int csm_host_mapbuilder_last_exact_beams(void* p) { return static_cast<HostMapBuilder*>(p)->builder->LatestGrid().LastExactBeams(); }
void csm_host_mapbuilder_set_fast_hit_points(void* p, int on) { static_cast<HostMapBuilder*>(p)->builder->SetFastHitPoints(on != 0); }

/* Append one scan node (global pose, scan) and rebuild the latest map, like the front end does per scan.
 * Returns the number of beams cast. */
int csm_host_mapbuilder_append(void* p, const double pose[3], const double* angles, const double* ranges, int n,
                               const double rel_pose[3], double min_range, double max_range)
{
    auto* b = static_cast<HostMapBuilder*>(p);
    auto scan = std::make_shared<ScanData>();
    scan->angles.assign(angles, angles + n);
    scan->ranges.assign(ranges, ranges + n);
    scan->relative_sensor_pose = Pose2D { rel_pose[0], rel_pose[1], rel_pose[2] };
    scan->min_range = min_range; scan->max_range = max_range;
    b->nodes.push_back(ScanNodeView { Pose2D { pose[0], pose[1], pose[2] }, scan });
    b->builder->UpdateLatestMap(b->nodes);
    return b->builder->LastNumOfRays();
}

/* geometry6 = rows, cols, block size, offset x, y, resolution; the map itself and its block allocation */
int csm_host_mapbuilder_latest(void* p, double* geometry6, double* map_pose3, uint16_t* dense, int cap_cells,
                               uint8_t* alloc, int cap_blocks)
{
    auto* b = static_cast<HostMapBuilder*>(p);
    const GridMapBuilderGPU& g = *b->builder;
    geometry6[0] = g.Rows(); geometry6[1] = g.Cols(); geometry6[2] = g.BlockSize();
    geometry6[3] = g.OffsetX(); geometry6[4] = g.OffsetY(); geometry6[5] = g.LatestMap().resolution;
    map_pose3[0] = g.LatestMapPose().x; map_pose3[1] = g.LatestMapPose().y; map_pose3[2] = g.LatestMapPose().theta;
    const int blocks = (g.Rows() / g.BlockSize()) * (g.Cols() / g.BlockSize());
    if (g.Rows() * g.Cols() > cap_cells || blocks > cap_blocks)
        return -1;
    csm_handle h = b->ctx->Handle();
    b->ctx->Check(csm_map_download_cells(h, g.LatestMap().map_id, dense), "csm_map_download_cells");
    b->ctx->Check(csm_map_download_allocation(h, g.LatestMap().map_id, alloc), "csm_map_download_allocation");
    return 0;
}

/* The front end's pair on the RESIDENT latest map: real-time correlative match + the final matcher on the
 * device (the context must have SetDeviceFinalMatcher / SetDeviceEpilogue) */
int csm_host_mapbuilder_match_rt(void* p, const double* angles, const double* ranges, int n, const double rel_pose[3],
                                 const double init_pose[3], int low_resolution, const double range[3],
                                 double covariance_scale, csm_host_summary* out)
{
    auto* b = static_cast<HostMapBuilder*>(p);
    try {
        const auto cost = std::make_shared<CostSquareError>(covariance_scale);
        const ScanDataPtr scan = Scan(angles, ranges, n, rel_pose);
        ScanMatcherCorrelative m("RealTimeCorrelativeGPU", cost, low_resolution, range[0], range[1], range[2], b->ctx);
        const ScanMatchingSummary s = m.OptimizePose(b->builder->LatestMap(), scan,
                                                     Pose2D { init_pose[0], init_pose[1], init_pose[2] }, 0.0, 0.0);
        Export(s, out);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "csm_host_mapbuilder_match_rt: %s\n", e.what());
        return -1;
    }
    return 0;
}

/* ---- the full loop: SlamPipeline ------------------------------------------------------------------ */
struct HostSlam
{
    DeviceContextPtr ctx;
    std::shared_ptr<PoseGraphOptimizerIdentity> optimizer;
    std::unique_ptr<SlamPipeline> slam;
    std::shared_ptr<MetricRecorder> metrics;
};

/* `v`: the SlamSettings fields in declaration order (slam_pipeline.hpp), initial_pose as three values:
 * 37 doubles (slam_settings.py packs them for both arms) */
void* csm_host_slam_create(void* ctx, const double* v, int n)
{
    if (n != 37)
        return nullptr;
    SlamSettings s;
    int i = 0;
    s.resolution = v[i++]; s.patch_size = static_cast<int>(v[i++]); s.scans_for_latest_map = static_cast<int>(v[i++]);
    s.local_map_travel_dist = v[i++]; s.overlapped_scans = static_cast<int>(v[i++]);
    s.usable_range_min = v[i++]; s.usable_range_max = v[i++]; s.prob_hit = v[i++]; s.prob_miss = v[i++];
    s.update_travel_dist = v[i++]; s.update_angle = v[i++]; s.update_time = v[i++];
    s.loop_detection_threshold = v[i++]; s.degeneration_threshold = v[i++]; s.odometry_covariance_scale = v[i++];
    s.fuse_odometry_covariance = v[i++] != 0.0;
    s.initial_pose.x = v[i++]; s.initial_pose.y = v[i++]; s.initial_pose.theta = v[i++];
    s.rt_low_resolution = static_cast<int>(v[i++]); s.rt_range_x = v[i++]; s.rt_range_y = v[i++]; s.rt_range_theta = v[i++];
    s.final_iterations = static_cast<int>(v[i++]); s.final_convergence = v[i++]; s.final_lambda = v[i++];
    s.covariance_scale = v[i++];
    s.searcher_travel_dist = v[i++]; s.searcher_node_dist = v[i++]; s.searcher_candidates = static_cast<int>(v[i++]);
    s.bb_node_height_max = static_cast<int>(v[i++]); s.bb_range_x = v[i++]; s.bb_range_y = v[i++]; s.bb_range_theta = v[i++];
    s.score_threshold = v[i++]; s.known_rate_threshold = v[i++];
    s.host_final_matchers = v[i++] != 0.0;
    auto* p = new HostSlam;
    p->ctx = *static_cast<DeviceContextPtr*>(ctx);
    p->optimizer = std::make_shared<PoseGraphOptimizerIdentity>();
    p->slam.reset(new SlamPipeline(p->ctx, s, p->optimizer));
    return p;
}

void csm_host_slam_destroy(void* p) { delete static_cast<HostSlam*>(p); }

/* n_scans scans of n_beams each (row-major ranges, shared angles), their odometry poses and time stamps:
 * the whole trajectory in one call, so that the timing holds no Python. Returns the scans used. */
int csm_host_slam_run(void* p, int n_scans, int n_beams, const double* angles, const double* ranges,
                      const double* odom_poses, const double* time_stamps, double min_range, double max_range,
                      int finish)
{
    auto* hs = static_cast<HostSlam*>(p);
    int used = 0;
    for (int k = 0; k < n_scans; ++k) {
        auto scan = std::make_shared<ScanData>();
        scan->angles.assign(angles, angles + n_beams);
        scan->ranges.assign(ranges + static_cast<std::size_t>(k) * n_beams, ranges + static_cast<std::size_t>(k + 1) * n_beams);
        scan->min_range = min_range; scan->max_range = max_range;
        used += hs->slam->ProcessScan(scan, Pose2D { odom_poses[3 * k], odom_poses[3 * k + 1], odom_poses[3 * k + 2] },
                                      time_stamps[k]) ? 1 : 0;
    }
    if (finish)
        hs->slam->Finish();
    return used;
}

/* 14 values: scans in, processed, back-end steps, steps with candidates, loop queries, loops detected,
 * optimisations, degenerations, optimiser calls seen behind the seam; seconds in latest map, match,
 * append, back end, detect */
void csm_host_slam_counters(void* p, double* out)
{
    auto* hs = static_cast<HostSlam*>(p);
    const SlamCounters& c = hs->slam->Counters();
    const double v[14] = { double(c.scans_in), double(c.scans_processed), double(c.backend_steps),
                           double(c.backend_steps_with_candidates), double(c.loop_queries), double(c.loops_detected),
                           double(c.optimizations), double(c.degenerations), double(hs->optimizer->Calls()),
                           c.t_latest_map, c.t_match, c.t_append, c.t_backend, c.t_detect };
    std::copy(v, v + 14, out);
}

int csm_host_slam_num_scan_nodes(void* p) { return static_cast<int>(static_cast<HostSlam*>(p)->slam->Graph().scan_nodes.size()); }
int csm_host_slam_num_local_maps(void* p) { return static_cast<int>(static_cast<HostSlam*>(p)->slam->Builder().LocalMaps().size()); }
int csm_host_slam_num_edges(void* p) { return static_cast<int>(static_cast<HostSlam*>(p)->slam->Graph().edges.size()); }
int csm_host_slam_num_loops(void* p) { return static_cast<int>(static_cast<HostSlam*>(p)->slam->Loops().size()); }

/* per scan node: global pose (3), local pose (3), local map id */
void csm_host_slam_scan_nodes(void* p, double* out7)
{
    const PoseGraph& g = static_cast<HostSlam*>(p)->slam->Graph();
    for (std::size_t i = 0; i < g.scan_nodes.size(); ++i) {
        const ScanNode& n = g.scan_nodes[i];
        double* o = out7 + 7 * i;
        o[0] = n.global_pose.x; o[1] = n.global_pose.y; o[2] = n.global_pose.theta;
        o[3] = n.local_pose.x; o[4] = n.local_pose.y; o[5] = n.local_pose.theta; o[6] = n.local_map_id;
    }
}

/* per local map: global pose (3), first and last scan node, finished, rows, cols, offset x, y */
void csm_host_slam_local_maps(void* p, double* out10)
{
    auto* hs = static_cast<HostSlam*>(p);
    const PoseGraph& g = hs->slam->Graph();
    const std::vector<LocalMapGPU>& maps = hs->slam->Builder().LocalMaps();
    for (std::size_t i = 0; i < maps.size(); ++i) {
        double* o = out10 + 10 * i;
        const Pose2D& pose = g.local_map_nodes[i].global_pose;
        o[0] = pose.x; o[1] = pose.y; o[2] = pose.theta;
        o[3] = maps[i].scan_node_id_min; o[4] = maps[i].scan_node_id_max; o[5] = maps[i].finished ? 1.0 : 0.0;
        o[6] = maps[i].map->Rows(); o[7] = maps[i].map->Cols(); o[8] = maps[i].map->OffsetX(); o[9] = maps[i].map->OffsetY();
    }
}

/* cells (rows x cols) and block allocation of local map `id` */
int csm_host_slam_local_map_cells(void* p, int id, uint16_t* dense, int cap_cells, uint8_t* alloc, int cap_blocks)
{
    auto* hs = static_cast<HostSlam*>(p);
    const std::vector<LocalMapGPU>& maps = hs->slam->Builder().LocalMaps();
    if (id < 0 || id >= static_cast<int>(maps.size()))
        return -1;
    const DeviceGridMap& m = *maps[id].map;
    const int bs = m.BlockSize();
    if (m.Rows() * m.Cols() > cap_cells || (m.Rows() / bs) * (m.Cols() / bs) > cap_blocks)
        return -2;
    hs->ctx->Check(csm_map_download_cells(hs->ctx->Handle(), m.MapId(), dense), "csm_map_download_cells");
    hs->ctx->Check(csm_map_download_allocation(hs->ctx->Handle(), m.MapId(), alloc), "csm_map_download_allocation");
    return 0;
}

/* per edge: local map id, scan node id, inter-local-map?, loop?, relative pose (3) */
void csm_host_slam_edges(void* p, double* out7)
{
    const PoseGraph& g = static_cast<HostSlam*>(p)->slam->Graph();
    for (std::size_t i = 0; i < g.edges.size(); ++i) {
        const PoseGraphEdge& e = g.edges[i];
        double* o = out7 + 7 * i;
        o[0] = e.local_map_id; o[1] = e.scan_node_id; o[2] = e.edge_type == EdgeType::InterLocalMap ? 1.0 : 0.0;
        o[3] = e.IsLoopClosingConstraint() ? 1.0 : 0.0;
        o[4] = e.relative_pose.x; o[5] = e.relative_pose.y; o[6] = e.relative_pose.theta;
    }
}

/* per detected loop: local map id, scan node id, relative pose (3), normalized score */
void csm_host_slam_loops(void* p, double* out6)
{
    const std::vector<LoopDetectionResult>& loops = static_cast<HostSlam*>(p)->slam->Loops();
    for (std::size_t i = 0; i < loops.size(); ++i) {
        double* o = out6 + 6 * i;
        o[0] = static_cast<double>(loops[i].local_map_id); o[1] = loops[i].scan_node_id;
        o[2] = loops[i].relative_pose.x; o[3] = loops[i].relative_pose.y; o[4] = loops[i].relative_pose.theta;
        o[5] = loops[i].normalized_score;
    }
}

/* ---- Carmen logs and the metrics file (carmen_log.hpp) ------------------------------------------------ */
struct HostCarmen { std::vector<CarmenRecord> records; };

void* csm_host_carmen_load(const char* text_or_path, int is_path)
{
    auto* p = new HostCarmen;
    CarmenLogReader reader;
    bool ok;
    if (is_path) {
        ok = reader.LoadFile(text_or_path, p->records);
    } else {
        std::istringstream in { std::string(text_or_path) };
        ok = reader.Load(in, p->records);
    }
    if (!ok) {
        delete p;
        return nullptr;
    }
    return p;
}
void csm_host_carmen_destroy(void* p) { delete static_cast<HostCarmen*>(p); }
int csm_host_carmen_count(void* p) { return static_cast<int>(static_cast<HostCarmen*>(p)->records.size()); }
int csm_host_carmen_total_beams(void* p)
{
    std::size_t n = 0;
    for (const CarmenRecord& r : static_cast<HostCarmen*>(p)->records)
        if (r.scan) n += r.scan->ranges.size();
    return static_cast<int>(n);
}
/* per record 15 values: kind, time stamp, odometry pose, forward and angular velocity, sensor pose on the robot,
 * min / max range, min / max angle, beams; then every scan's angles and ranges back to back */
void csm_host_carmen_export(void* p, double* head15, double* angles, double* ranges)
{
    std::size_t at = 0;
    for (const CarmenRecord& r : static_cast<HostCarmen*>(p)->records) {
        const ScanData* s = r.scan.get();
        const double v[15] = { double(int(r.kind)), r.time_stamp, r.odom_pose.x, r.odom_pose.y, r.odom_pose.theta,
                               r.velocity.x, r.velocity.theta,
                               s ? s->relative_sensor_pose.x : 0.0, s ? s->relative_sensor_pose.y : 0.0,
                               s ? s->relative_sensor_pose.theta : 0.0, s ? s->min_range : 0.0, s ? s->max_range : 0.0,
                               r.min_angle, r.max_angle, s ? double(s->ranges.size()) : 0.0 };
        std::copy(v, v + 15, head15);
        head15 += 15;
        if (s) {
            std::copy(s->angles.begin(), s->angles.end(), angles + at);
            std::copy(s->ranges.begin(), s->ranges.end(), ranges + at);
            at += s->ranges.size();
        }
    }
}
int csm_host_carmen_sensor_id(void* p, int i, char* buf, int cap)
{
    const std::string& id = static_cast<HostCarmen*>(p)->records.at(static_cast<std::size_t>(i)).sensor_id;
    std::snprintf(buf, static_cast<std::size_t>(cap), "%s", id.c_str());
    return static_cast<int>(id.size());
}

/* a synthetic run as a Carmen log: per scan an ODOM record (optional) and a ROBOTLASER1 (format 0) or FLASER
 * (format 1, with the PARAM records that carry the beam geometry) record */
int csm_host_carmen_write(const char* path, int format, int n_scans, int n_beams, double start_angle,
                          double angular_resolution, double max_range, const double* ranges,
                          const double* odom_poses, const double laser_on_robot[3], const double* time_stamps,
                          int with_odom)
{
    std::ofstream out(path);
    if (!out)
        return -1;
    CarmenLogWriter w(out);
    auto g17 = [](double v) { char b[40]; std::snprintf(b, sizeof b, "%.17g", v); return std::string(b); };
    if (format == 1) {
        w.Param("Laser.MinRange", "0");
        w.Param("Laser.MaxRange", g17(max_range));
        w.Param("Laser.AngleIncrement", g17(angular_resolution));
        w.Param("Laser.MinAngle", g17(start_angle));
    }
    const Pose2D rel { laser_on_robot[0], laser_on_robot[1], laser_on_robot[2] };
    std::vector<double> r(static_cast<std::size_t>(n_beams));
    for (int k = 0; k < n_scans; ++k) {
        const Pose2D robot { odom_poses[3 * k], odom_poses[3 * k + 1], odom_poses[3 * k + 2] };
        const Pose2D laser = Compound(robot, rel);
        r.assign(ranges + static_cast<std::size_t>(k) * n_beams, ranges + static_cast<std::size_t>(k + 1) * n_beams);
        if (with_odom)
            w.Odom(robot, 0.0, 0.0, time_stamps[k]);
        if (format == 1)
            w.OldLaser("FLASER", r, laser, robot, time_stamps[k]);
        else
            w.RobotLaser("ROBOTLASER1", start_angle, angular_resolution, max_range, r, laser, robot, time_stamps[k]);
    }
    return out ? 0 : -2;
}

int csm_host_metric_values_string(const char* id, const double* values, int n, char* buf, int cap)
{
    const std::string s = MetricValuesToString(id, std::vector<double>(values, values + n));
    std::snprintf(buf, static_cast<std::size_t>(cap), "%s", s.c_str());
    return static_cast<int>(s.size());
}

/* WriteMetricsJson over a recorder that observed values[k] under ids[k] (newline separated, in order) */
int csm_host_metrics_json(const char* ids, const int* counts, int n_ids, const double* values, char* buf, int cap)
{
    MetricRecorder rec;
    std::istringstream in { std::string(ids) };
    std::string id;
    for (int k = 0; k < n_ids && std::getline(in, id); ++k)
        for (int i = 0; i < counts[k]; ++i)
            rec.Observe(id, *values++);
    std::ostringstream out;
    WriteMetricsJson(out, rec);
    std::snprintf(buf, static_cast<std::size_t>(cap), "%s", out.str().c_str());
    return static_cast<int>(out.str().size());
}

int csm_host_slam_run_carmen(void* slam, void* records, int finish)
{
    return static_cast<HostSlam*>(slam)->slam->RunLog(static_cast<HostCarmen*>(records)->records, finish != 0);
}
void csm_host_slam_record_metrics(void* slam)
{
    auto* hs = static_cast<HostSlam*>(slam);
    hs->metrics = std::make_shared<MetricRecorder>();
    hs->slam->SetMetricSink(hs->metrics);
}
/* writes path + ".metric.json" (slam_launcher.cpp:171-181); -1 when no recorder is set */
int csm_host_slam_save_metrics(void* slam, const char* path)
{
    auto* hs = static_cast<HostSlam*>(slam);
    if (!hs->metrics)
        return -1;
    return SaveMetrics(path, *hs->metrics) ? 0 : -2;
}

} /* extern "C" */
