/* oracle_api.h -- C interface shared by the two CPU checkers.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product library
 * (my_lidar_graph_slam_v2_b200/csrc, include/csm_b200.h) includes, links or calls
 * anything declared here. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py load these libraries.
 *
 * Two shared libraries export exactly this interface:
 *   oracle/_ref/libcsm_ref.so   the UNMODIFIED reference translation units
 *                               from /root/reference compiled with the header
 *                               shims in oracle/ref_shim (see oracle/Makefile),
 *                               driven by oracle/ref_wrapper.cpp;
 *   oracle/libcsm_port.so       oracle/port.cpp, a plain C++ restatement of
 *                               the same algorithms (each function cites the
 *                               reference file:line it follows).
 */
#ifndef CSM_ORACLE_API_H
#define CSM_ORACLE_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_result
{
    int32_t found;          /* ScanMatchingSummary::mPoseFound */
    int32_t best_x;         /* best window index along x (cells; dx[] index for grid search) */
    int32_t best_y;         /* best window index along y */
    int32_t best_t;         /* best window index along theta */
    int32_t win_x;          /* half window, cells (0 for grid search) */
    int32_t win_y;
    int32_t win_t;
    int32_t n_known;        /* number of known (non-zero) cells hit at the best pose */
    int64_t sum_value;      /* sum of the u16 cell values over those cells */
    double  step_x;
    double  step_y;
    double  step_t;
    double  score;          /* normalized score at the best pose, reference double arithmetic */
    double  known_rate;
    int32_t n_processed;    /* reference metric NumOfProcessedNodes (evaluations for grid search) */
    int32_t n_ignored;      /* reference metric NumOfIgnoredNodes (updates for grid search) */
    double  best_sensor_pose[3];
    double  est_pose[3];    /* ScanMatchingSummary::mEstimatedPose */
    double  norm_cost;      /* ScanMatchingSummary::mNormalizedCost */
    double  cov[9];         /* ScanMatchingSummary::mEstimatedCovariance, row-major */
} orc_result;

/* Short tag: "reference" or "port" */
const char* orc_kind(void);

/* Dense row-major u16 grid (0 = unknown). rows and cols must be multiples of 16
 * (reference block size, launcher_settings_default.json:177-178). */
void* orc_grid_create(const uint16_t* dense, int rows, int cols,
                      double resolution, double offset_x, double offset_y);
void  orc_grid_destroy(void* grid);

/* PrecomputeGridMap(map, win): grid_map_builder.cpp:1044-1065. out: rows*cols */
int orc_precompute(void* grid, int win, uint16_t* out);
/* PrecomputeGridMaps(map, out, hmax): grid_map_builder.cpp:987-1012.
 * out: (hmax+1)*rows*cols, level h uses win = 2^h */
int orc_precompute_pyramid(void* grid, int hmax, uint16_t* out);

/* ScanMatcherCorrelative::OptimizePose: scan_matcher_correlative.cpp:92-244 */
int orc_match_rt(void* grid, const double* angles, const double* ranges, int n,
                 const double init_pose[3], const double rel_sensor_pose[3],
                 int low_res, double range_x, double range_y, double range_t,
                 double score_thr, double known_thr, orc_result* out);

/* ScanMatcherBranchBound::OptimizePose: scan_matcher_branch_bound.cpp:87-278 */
int orc_match_bb(void* grid, const double* angles, const double* ranges, int n,
                 const double init_pose[3], const double rel_sensor_pose[3],
                 int hmax, double range_x, double range_y, double range_t,
                 double score_thr, double known_thr, orc_result* out);

/* ScanMatcherGridSearch::OptimizePose: scan_matcher_grid_search.cpp:69-178 */
int orc_match_grid(void* grid, const double* angles, const double* ranges, int n,
                   const double init_pose[3], const double rel_sensor_pose[3],
                   double range_x, double range_y, double range_t,
                   double step_x, double step_y, double step_t,
                   double score_thr, double known_thr, orc_result* out);

/* ScanMatcherLinearSolver::OptimizePose: scan_matcher_linear_solver.cpp:66-140 (the reference's
 * default final matcher). `lambda` is the solver's damping state: read at entry, written back at
 * exit, so that a sequence of calls behaves like one solver instance. out: est_pose, norm_cost,
 * cov; n_processed = number of iterations. */
int orc_refine(void* grid, const double* angles, const double* ranges, int n,
               const double init_pose[3], const double rel_sensor_pose[3],
               int iterations_max, double convergence_threshold, double* lambda,
               orc_result* out);

/* Use a ScanMatcherLinearSolver (iterations_max, convergence_threshold, initial_lambda) as the
 * final matcher of the loop detector instead of the pass-through one. */
void orc_loopdet_use_linear_solver(void* det, int iterations_max, double convergence_threshold,
                                   double initial_lambda);

/* LoopDetectorBranchBound: loop_detector_branch_bound.cpp:59-156, with a
 * pass-through final matcher (the sub-pixel refiner is outside the path).
 * Queries are split into n_threads contiguous ranges, one detector (and one
 * pyramid cache) per thread, like loop_detector_fpga_parallel.cpp:41-56. */
/* ---- map construction: GridMapBuilder::UpdateLatestMap (grid_map_builder.cpp:497-532, 561-695) on the
 * reference's GridMapBuilder (one instance per process, re-initialised by orc_mapbuilder_create) ---- */
void* orc_mapbuilder_create(double resolution, int patch_size, int scans_for_latest_map,
                            double usable_range_min, double usable_range_max, double prob_hit, double prob_miss);
int orc_mapbuilder_append(void* builder, const double pose[3], const double* angles, const double* ranges, int n,
                          const double rel_pose[3], double min_range, double max_range);
int orc_mapbuilder_latest(void* builder, double* geometry6, double* map_pose3, uint16_t* dense, int cap_cells,
                          uint8_t* alloc, int cap_blocks);

void* orc_loopdet_create(int hmax, double range_x, double range_y, double range_t,
                         double score_thr, double known_thr, int n_threads);
void  orc_loopdet_destroy(void* det);
/* Forget all cached pyramids (next Detect is a first touch for every map) */
void  orc_loopdet_clear_cache(void* det);
/* scan q uses angles/ranges[scan_idx[q]*n_beams ...]; poses are 3 doubles per query.
 * out[q].found == 0 when the reference emits no result for query q.
 * elapsed_s: wall time of the Detect calls (max over threads). */
int orc_loopdet_detect(void* det, int n_queries,
                       void* const* grids, const int32_t* map_ids,
                       const double* map_global_poses,
                       const int32_t* scan_idx, const double* scan_global_poses,
                       int n_scans, int n_beams,
                       const double* angles, const double* ranges,
                       orc_result* out, double* elapsed_s);

/* ScanMatcherHillClimbing::OptimizePose (scan_matcher_hill_climbing.cpp:63-170) over CostSquareError.
 * out: est_pose, norm_cost, cov; n_processed = iterations, n_ignored = step halvings. Returns -1
 * when this checker does not provide it. */
int orc_hill_climb(void* grid, const double* angles, const double* ranges, int n,
                   const double init_pose[3], const double rel_sensor_pose[3],
                   double linear_step, double angular_step, int max_iterations,
                   int max_num_of_refinements, const double* greedy, orc_result* out);
/* greedy == NULL: CostSquareError; else CostGreedyEndpoint(MapResolution, HitAndMissedDist,
 * OccupancyThreshold, KernelSize, ScalingFactor, StandardDeviation) (cost_function_greedy_endpoint.cpp:9-30) */

/* LoopSearcherNearest::Search (loop_searcher_nearest.cpp:59-170) on a pose-graph summary: scan nodes
 * (ids ascending, 3 doubles of global pose each), local maps (ids ascending, scan-node id range,
 * finished flag). Writes up to cap candidates as (query scan node, reference scan node, reference
 * local map) triples in the order the reference returns them; returns their number, -1 when this
 * checker does not provide it. */
int orc_loop_search(int n_scans, const int32_t* scan_ids, const double* scan_poses,
                    int n_maps, const int32_t* map_ids, const int32_t* map_scan_min,
                    const int32_t* map_scan_max, const int32_t* map_finished,
                    double accum_travel_dist, int last_finished_scan_id, int last_finished_map_id,
                    double travel_dist_threshold, double node_dist_threshold,
                    int num_of_candidate_nodes, int32_t* out_ids, int cap);

/* out[v], v = 0 .. 65535: the cell value after one GridBinaryBayes::UpdateOddsUnchecked(odds) of a cell at v (ref only) */
int orc_update_table(double odds, uint16_t* out65536);
/* GridBinaryBayes::ValueToProbability(v) (ref only) */
double orc_value_probability(int v);

/* The full loop with the reference's own components (ref only); settings: slam_settings.py; the tables
 * have the layouts of csm_host_slam_* (host/src/c_shim.cpp) */
void* orc_slam_create(const double* settings, int n);
void orc_slam_destroy(void* slam);
int orc_slam_run(void* slam, int n_scans, int n_beams, const double* angles, const double* ranges,
                 const double* odom_poses, const double* time_stamps, double min_range, double max_range, int finish);
void orc_slam_counters(void* slam, double* out14);
int orc_slam_num_scan_nodes(void* slam);
int orc_slam_num_local_maps(void* slam);
int orc_slam_num_edges(void* slam);
int orc_slam_num_loops(void* slam);
void orc_slam_scan_nodes(void* slam, double* out7);
void orc_slam_local_maps(void* slam, double* out10);
int orc_slam_local_map_cells(void* slam, int id, uint16_t* dense, int cap_cells, uint8_t* alloc, int cap_blocks);
void orc_slam_edges(void* slam, double* out7);
void orc_slam_loops(void* slam, double* out6);

#ifdef __cplusplus
}
#endif

#endif /* CSM_ORACLE_API_H */
