/* csm_b200.h -- C ABI of the B200-native correlative scan matching /
 * loop-detection hot path (libcsm_b200.so).
 *
 * This is the drop-in boundary: plain C types, no exceptions, no torch types.
 * It is what a host-side adapter class deriving from the reference's
 * ScanMatcher / LoopDetector plugin interfaces binds to (see INTEGRATION.md and
 * my_lidar_graph_slam_v2_b200/host/). The boundary sits where the reference's
 * own accelerator offload puts it: the register/DMA calls of
 * scan_matcher_correlative_fpga.cpp:277-311 (SetParameterRegisters /
 * SendScanData / SendGridMap / ReceiveResult), with a device-side map cache
 * keyed by LocalMapId like scan_matcher_correlative_fpga.cpp:261-270,301-304.
 *
 * Citations are paths relative to the reference repository root.
 *
 * Conventions
 *  - Every function returns CSM_OK (0) or a negative CSM_E_* status;
 *    csm_last_error(h) gives a message. "No pose found" is a normal result
 *    (csm_result.found == 0), not an error (scan_matcher_correlative.cpp:201).
 *  - One handle per matcher / detector instance. A handle owns one CUDA
 *    stream and all its device buffers; the library has no process-global
 *    mutable state, so the front-end matcher and the back-end loop detector
 *    (two threads, lidar_graph_slam.cpp:777-779) use separate handles
 *    concurrently. A single handle must not be used from two threads at once.
 *  - All host pointers are borrowed for the duration of the call only.
 *  - There is no CPU fallback: every entry point fails with CSM_E_CUDA when
 *    no CUDA device is usable.
 */
#ifndef CSM_B200_H
#define CSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSM_OK              0
#define CSM_E_INVALID      -1   /* bad argument */
#define CSM_E_CUDA         -2   /* CUDA runtime error (see csm_last_error) */
#define CSM_E_NOT_FOUND    -3   /* unknown map / scan id */
#define CSM_E_CAPACITY     -4   /* a device-side queue overflowed */
#define CSM_E_UNSUPPORTED  -5   /* parameter outside the supported range */

/* csm_result.flags */
#define CSM_FLAG_FP_MARGIN   1  /* a projected hit point lies within the FP
                                 * guard band of a cell boundary (DESIGN.md
                                 * "FP index parity") and the exact rerun is
                                 * switched off (option "exact_rerun" = 0):
                                 * the result may differ from glibc's */
#define CSM_FLAG_KEY_TIE     2  /* the winning integer key is shared by
                                 * another candidate; first in reference
                                 * iteration order was taken */
#define CSM_FLAG_EDGE        4  /* B&B: a node window reaches below row /
                                 * column 0 of a map that has known cells in
                                 * its first 2^hmax rows or columns (the
                                 * reference's bound is not admissible there,
                                 * SURVEY.md A.11) */
#define CSM_FLAG_EXACT       8  /* a projected hit point lay within the FP
                                 * guard band of a cell boundary, where device
                                 * and glibc trigonometry may round to
                                 * different cells: the result was recomputed
                                 * from indices evaluated on the host with the
                                 * reference's own libm calls and operation
                                 * order (replaces CSM_FLAG_FP_MARGIN) */

typedef struct csm_context* csm_handle;

/* Result of one match. Mirrors what the reference matchers decide before the
 * CPU epilogue (Cost / ComputeCovariance / MoveBackward,
 * scan_matcher_correlative.cpp:199-219). */
typedef struct csm_result
{
    int32_t found;            /* poseFound: best score > score threshold */
    int32_t best_x;           /* RT / B&B: window index in cells (bestWinX, Node::mX);
                                 grid search: index into dx[] (-1 if not found) */
    int32_t best_y;
    int32_t best_t;
    int64_t sum_value;        /* sum of u16 cell values over known cells at the best pose */
    int32_t n_known;          /* number of known (non-zero) cells at the best pose */
    int32_t flags;            /* CSM_FLAG_* */
    double  normalized_score; /* sum of probabilities / N, accumulated in scan
                                 order in double like score_function_pixel_accurate.cpp:21-57 */
    int32_t n_processed;      /* nodes expanded / coarse cells refined on the device */
    int32_t n_ignored;        /* nodes pruned / coarse cells skipped on the device */
} csm_result;

/* One loop-detection query (loop_detector.hpp:27-55 after the adapter has
 * resolved references to ids and computed the map-local sensor pose with
 * InverseCompound + Compound, loop_detector_branch_bound.cpp:97-98,
 * scan_matcher_branch_bound.cpp:124-125). */
typedef struct csm_loop_query
{
    int64_t map_id;           /* LocalMapId::mId of the reference local map (uploaded, pyramid built) */
    int64_t scan_id;          /* id of a scan uploaded with csm_upload_scan */
    double  sensor_pose[3];   /* map-local sensor pose (x, y, theta) */
    int32_t win_x;            /* half windows: ceil(0.5 * range / step), :141-146 */
    int32_t win_y;
    int32_t win_t;
    int32_t reserved;
    double  step_x;           /* search steps, scan_matcher_branch_bound.cpp:293-312 */
    double  step_y;
    double  step_t;
    double  score_thr;        /* normalized score threshold */
    double  known_thr;        /* known-rate threshold */
} csm_loop_query;

/* Parameters of the refinement stage that follows a successful coarse match in
 * the reference's loop detectors: ScanMatcherLinearSolver with CostSquareError
 * ("FinalScanMatcherType": "LinearSolver", launcher_settings_default.json:28-35,
 * scan_matcher_linear_solver.cpp:46-64, cost_function_square_error.cpp:16-24). */
typedef struct csm_refine_params
{
    int32_t max_iterations;         /* NumOfIterationsMax */
    int32_t reserved;
    double  convergence_threshold;  /* ConvergenceThreshold, on the summed squared error */
    double  lambda;                 /* damping factor every query of the next batch starts from */
    double  covariance_scale;       /* CostSquareError CovarianceScale */
} csm_refine_params;

/* Outcome of the refinement of one query (valid == 0: nothing to refine). */
typedef struct csm_refined
{
    double  pose[3];          /* refined map-local SENSOR pose (the adapter applies MoveBackward,
                                 scan_matcher_linear_solver.cpp:113-114) */
    double  covariance[9];    /* ComputeCovariance at the refined pose, row-major (:117-118) */
    double  initial_cost;     /* sum of squared errors at the coarse pose (not normalized) */
    double  final_cost;       /* ... at the refined pose */
    double  lambda;           /* damping factor after the last iteration */
    int32_t iterations;
    int32_t valid;
} csm_refined;

/* One pose to refine on its own (csm_refine_batch) */
typedef struct csm_refine_query
{
    int64_t map_id;
    int64_t scan_id;
    double  sensor_pose[3];   /* map-local sensor pose to start from */
} csm_refine_query;

/* ---- lifetime ---------------------------------------------------------- */
int  csm_version(void);
int  csm_device_count(void);
/* device: CUDA ordinal. flags: reserved, pass 0. */
int  csm_create(int device, unsigned flags, csm_handle* out);
int  csm_destroy(csm_handle h);
const char* csm_last_error(csm_handle h);
/* The handle's CUDA stream (cudaStream_t) for event timing by the caller */
void* csm_stream(csm_handle h);
/* Block until everything enqueued on the handle's stream has finished */
int  csm_synchronize(csm_handle h);
/* Number of kernels this handle has launched so far */
int64_t csm_launch_count(csm_handle h);
/* Number of results this handle has recomputed exactly so far (CSM_FLAG_EXACT) */
int64_t csm_exact_rerun_count(csm_handle h);
/* Make `h` enqueue its host-to-device grid uploads on `owner`'s copy stream. Handles that serve
 * as pipeline lanes of one detector (same host thread, same device) then upload strictly in call
 * order: copies issued on different streams share the link and would all land together at the
 * end, whereas a lane's search should start when ITS maps have landed. `owner` must outlive `h`. */
int csm_share_copy_stream(csm_handle h, csm_handle owner);
/* Tuning / test knobs. (The environment variable CSM_OPTIONS="name=value,name=value" applies csm_set_option to
 * every handle a process creates: A/B runs through host code that does not expose the knobs.)
 *  "pyramid_mode": 0 = automatic, 1 = level-by-level kernels, 2 = streaming
 *      single-pass kernel (when the maps fit its layout), 3 = the streaming
 *      kernel variant that keeps its row rings in shared memory;
 *  "bb_dive": every branch-and-bound query first descends greedily (beam of 8)
 *      to a leaf that seeds its incumbent: 0 = never (plain level sweep), 1 =
 *      always, 2 (default) = only for calls of at most 4 queries. Results are
 *      identical either way; only the number of nodes scored changes;
 *  "bb_skip_top": 1 (default) = the branch-and-bound sweep starts one height
 *      below hmax on the same leaf lattice (identical results, one launch less);
 *  "bb_probe": the sweep over bound levels seeds its incumbents early: after the launch that creates the
 *      children of height 4 (1, default), after those of heights 5 and 4 (2), or never (0), every query descends
 *      greedily from the children with the largest bounds to leaves that are scored exactly. Results are
 *      identical either way; only the number of nodes scored below changes;
 *  "bb_bounds": 1 (default) = sweeps that neither score their roots nor dive read the
 *      u8 bound levels (tiled upper bounds of the reference's coarse levels, built by
 *      csm_build_pyramids for batches of maps or on first use), 0 = the reference's u16
 *      levels. Results are identical either way; a few percent more nodes are expanded;
 *  "window_mode": grid search, 0 (default) = TMA shared-memory tile kernel when
 *      the window fits (tiles of 32-bit value | known words, scored where they land),
 *      1 = global-memory kernel, 2 = require the TMA kernel, 3 = require the TMA
 *      kernel on u16 tiles that a pass per tile widens;
 *  "exact_rerun": 1 (default) = a result whose projection raised the FP guard-band
 *      flag is recomputed from host-evaluated indices (CSM_FLAG_EXACT), 0 = it is
 *      returned as is with CSM_FLAG_FP_MARGIN; "fp_margin_scale": test knob, multiplies
 *      the guard band (a huge value flags every result);
 *  "bb_sweep_ctas_per_sm": n > 0 = the branch-and-bound sweep launches at most n CTAs per SM (default 0:
 *      as many as fit, 4, which fill the register file). With batches of several handles in flight on
 *      one GPU, 2 leave room for the level builder and the projection of the other batches (+3 % steps/s,
 *      +7 % with cached levels on cfg3); alone on the GPU the full grid is faster. Results do not depend on it;
 *  "timing": 1 / 2 = record CUDA events after every kernel (csm_debug_timings);
 *  "accumulate_best_key": 1 = loop batches keep (do not reset) the packed
 *      best word, so that a Detect call split into several batches ends with
 *      the maximum over all of them; "reset_best_key": clears it now. */
int csm_set_option(csm_handle h, const char* name, int value);
/* Page-locked host memory for the caller's upload buffers (fast H2D) */
void* csm_alloc_pinned(size_t bytes);
void  csm_free_pinned(void* p);

/* ---- grid maps ---------------------------------------------------------
 * Replaces GridMap storage lookups (grid_map.cpp:385-397,424-436) and
 * SendGridMap of the FPGA path (scan_matcher_correlative_fpga.cpp:301-304).
 * `dense` is the row-major u16 flattening of the map (0 = unknown, 1..65535
 * <-> p in [0.001, 0.999], grid_binary_bayes.hpp:163-176); cell (row, col)
 * covers [off + res*col, off + res*(col+1)) (grid_map_geometry.cpp:113-122).
 * Re-uploading an id replaces the map and drops its precomputed levels. */
int csm_upload_grid(csm_handle h, int64_t map_id, const uint16_t* dense,
                    int rows, int cols, double resolution,
                    double offset_x, double offset_y);
/* Same, but `dense_dev` is a device pointer on the handle's device (the
 * copy is device-to-device on the handle's stream). */
int csm_upload_grid_device(csm_handle h, int64_t map_id, const uint16_t* dense_dev,
                           int rows, int cols, double resolution,
                           double offset_x, double offset_y);
/* n maps of identical shape and resolution in one call. The copies run
 * asynchronously on the handle's copy stream (use pinned buffers) and overlap
 * kernels working on maps uploaded by earlier calls; whatever later touches
 * one of these maps waits for this call's copies only. Page-locked buffers
 * must stay unchanged until csm_synchronize or a synchronous call that uses
 * one of these maps (csm_match_*, csm_loop_batch_finish, csm_download_level)
 * has returned. */
int csm_upload_grids(csm_handle h, int n, const int64_t* map_ids, const uint16_t* const* dense,
                     int rows, int cols, double resolution,
                     const double* offset_x, const double* offset_y);
/* Block-sparse upload: the reference stores a map as block_rows x block_cols
 * blocks of 2^k x 2^k cells and allocates only the blocks that were ever
 * written (grid_map.hpp:27, grid_map.cpp:262-266, 522-535); unallocated
 * blocks read as unknown (0). The adapter hands over exactly that: the
 * allocated blocks back to back (`blocks`, 4^k u16 each, row-major inside a
 * block) and their positions (`block_index[b]` = block_row * block_cols +
 * block_col). Only these bytes cross PCIe; a kernel expands them into the
 * dense level-0 grid when the map is first used. Equivalent to
 * csm_upload_grid with the flattened map. 3 <= log2_block_size <= 6. */
int csm_upload_grid_blocks(csm_handle h, int64_t map_id, const uint16_t* blocks,
                           const int32_t* block_index, int n_blocks, int log2_block_size,
                           int block_rows, int block_cols, double resolution,
                           double offset_x, double offset_y);
/* n maps of identical geometry in one call: map i owns block_count[i]
 * consecutive entries of `blocks` / `block_index`. One upload group, like
 * csm_upload_grids. */
int csm_upload_grids_blocks(csm_handle h, int n, const int64_t* map_ids,
                            const uint16_t* blocks, const int32_t* block_index,
                            const int32_t* block_count, int log2_block_size,
                            int block_rows, int block_cols, double resolution,
                            const double* offset_x, const double* offset_y);
int csm_release_grid(csm_handle h, int64_t map_id);

/* PrecomputeGridMap(map, win) (grid_map_builder.cpp:1044-1065): sliding
 * win x win maximum with the far edge clamped (util.hpp:369-424). */
int csm_build_coarse(csm_handle h, int64_t map_id, int win);
/* PrecomputeGridMaps(map, out, hmax) (grid_map_builder.cpp:987-1012):
 * levels 1..hmax with win = 2^h; level 0 is the uploaded grid itself. */
int csm_build_pyramid(csm_handle h, int64_t map_id, int hmax);
/* The same for n maps in one batch of launches (loop detection, first touch
 * of each map: loop_detector_branch_bound.cpp:83-89). Asynchronous: returns
 * after enqueueing; any later call on the handle is ordered after it. */
int csm_build_pyramids(csm_handle h, int n, const int64_t* map_ids, int hmax);
/* Mark the precomputed levels of these maps stale (device memory is kept), so
 * that the next build recomputes them: benchmarking the first-touch path. */
int csm_drop_pyramids(csm_handle h, int n, const int64_t* map_ids);
/* Copy a precomputed level back to the host (tests / debugging).
 * level >= 0: pyramid level; level < 0: the coarse map built with win = -level. */
int csm_download_level(csm_handle h, int64_t map_id, int level, uint16_t* out);

/* ---- map construction on the device ----------------------------------------
 * Replaces the ray casting of GridMapBuilder::UpdateGridMap / ConstructMapFromScans
 * (mapping/grid_map_builder.cpp:390-494, 561-695): the latest map / the growing local map stays on
 * the device, where the matchers read it, instead of being rebuilt on the CPU and uploaded per scan.
 * The adapter keeps the reference's geometry bookkeeping (GridMap::Resize / Expand,
 * grid_map.cpp:842-946, on the host: a few integers) and computes, per beam, what only needs the scan
 * and the poses: the sub-pixel indices of sensor and hit point and the hit cell
 * (grid_map_builder.cpp:445-463). Cell updates are bit-identical to the reference's: the update
 * v -> v' of GridBinaryBayes::UpdateOddsUnchecked (grid_binary_bayes.cpp:302-321) depends on v and the
 * odds only, so the adapter hands over the two tables T_miss[v], T_hit[v] (65536 u16 each, evaluated
 * with the reference's own arithmetic), and every cell sees its updates in the reference's order
 * (`order` ascending: scan by scan, beam by beam; within a beam the missed cells, then the hit cell). */
typedef struct csm_ray
{
    int32_t start_x, start_y;   /* sensor, index in the geometry scaled by `subpixel_scale` (col, row) */
    int32_t end_x, end_y;       /* hit point, same */
    int32_t hit_col, hit_row;   /* PositionToIndex(hit point) in the map's own geometry */
    int32_t order;              /* position of the beam in the reference's update sequence */
    int32_t reserved;
} csm_ray;
int csm_map_set_update_tables(csm_handle h, const uint16_t* t_miss, const uint16_t* t_hit);
/* A map of rows x cols unknown cells, no block allocated (GridMap construction / ResetValues on a map
 * whose blocks are dropped). rows, cols multiples of 2^log2_block_size. Replaces map_id. */
int csm_map_create(csm_handle h, int64_t map_id, int rows, int cols, int log2_block_size,
                   double resolution, double offset_x, double offset_y);
/* GridMap::Resize: new extent rows x cols whose cell (0, 0) is the old cell (row_min, col_min); cells
 * and block allocation of the overlap move, the rest is unknown / unallocated. */
int csm_map_resize(csm_handle h, int64_t map_id, int rows, int cols, int row_min, int col_min,
                   double offset_x, double offset_y);
/* GridMap::ResetValues: every cell unknown, block allocation kept. */
int csm_map_reset_values(csm_handle h, int64_t map_id);
/* Insert n beams. Asynchronous; precomputed levels of the map are dropped. A beam that leaves the map
 * is an error reported by the next csm_synchronize / csm_map_download_allocation. */
int csm_map_insert_rays(csm_handle h, int64_t map_id, const csm_ray* rays, int n, int subpixel_scale);
/* Block allocation bytes of the map (block_rows * block_cols, 1 = allocated); synchronous. */
int csm_map_download_allocation(csm_handle h, int64_t map_id, uint8_t* out);
/* The map's true cell values, rows x cols (GridMap's own u16). csm_download_level(level 0) returns what the
 * matchers read instead: the same cells with 65535 as 0 while option "saturated_unknown" is on. */
int csm_map_download_cells(csm_handle h, int64_t map_id, uint16_t* out);

/* ---- scans ---------------------------------------------------------------
 * Replaces SendScanData (scan_matcher_correlative_fpga.cpp:296-299): beam
 * angles and ranges of one ScanData (sensor/sensor_data.hpp:69-89). */
int csm_upload_scan(csm_handle h, int64_t scan_id,
                    const double* angles, const double* ranges, int n);
int csm_release_scan(csm_handle h, int64_t scan_id);

/* ---- single-scan matchers ---------------------------------------------------
 * sensor_pose = Compound(initial pose, relative sensor pose); steps and half
 * windows are computed by the caller with the reference's double expressions
 * (scan_matcher_correlative.cpp:141-146,255-274). */

/* ScanMatcherCorrelative::OptimizePose, scan_matcher_correlative.cpp:118-201.
 * The coarse map for win = low_res must have been built. */
int csm_match_rt(csm_handle h, int64_t map_id,
                 const double* angles, const double* ranges, int n,
                 const double sensor_pose[3], int low_res,
                 int win_x, int win_y, int win_t,
                 double step_x, double step_y, double step_t,
                 double score_thr, double known_thr, csm_result* out);

/* ScanMatcherBranchBound::OptimizePose, scan_matcher_branch_bound.cpp:111-235.
 * The pyramid up to hmax must have been built. */
int csm_match_bb(csm_handle h, int64_t map_id,
                 const double* angles, const double* ranges, int n,
                 const double sensor_pose[3], int hmax,
                 int win_x, int win_y, int win_t,
                 double step_x, double step_y, double step_t,
                 double score_thr, double known_thr, csm_result* out);

/* ScanMatcherGridSearch::OptimizePose, scan_matcher_grid_search.cpp:84-142.
 * dx/dy/dt are the loop values generated by the caller with the reference's
 * accumulating `for (d = -r; d <= r; d += s)` loops (:118-120). */
int csm_match_grid(csm_handle h, int64_t map_id,
                   const double* angles, const double* ranges, int n,
                   const double sensor_pose[3],
                   const double* dx, int ndx, const double* dy, int ndy,
                   const double* dt, int ndt,
                   double score_thr, double known_thr, csm_result* out);

/* ---- loop detection ----------------------------------------------------------
 * LoopDetectorBranchBound::Detect, loop_detector_branch_bound.cpp:59-156,
 * coarse stage: all queries of this rank's shard are matched in one batch of
 * level-synchronous branch-and-bound launches. results[q] is per query, in
 * query order (the reference emits one result per successful query).
 *
 * csm_loop_batch_enqueue returns after enqueueing the work (including the
 * read-back of the results into page-locked memory) on the handle's stream;
 * csm_loop_batch_finish waits for the OLDEST batch in flight and hands out
 * its nq results. Up to 4 batches may be in flight, so a Detect call split
 * into chunks keeps the GPU busy while the host prepares the next chunk.
 * csm_loop_batch = enqueue + finish. The single-scan matchers refuse to run
 * while a batch is in flight on the same handle.
 *
 * After a batch, csm_best_key_device(h) points to one uint64 on the device:
 * max over the batch of (key << 20 | (0xFFFFF - (query_index_base + q))), 0 if no query
 * found a pose, where key = 998 * sum_value + 64536 * n_known is the exact
 * integer image of the reference's double score. Ranks reduce it
 * with one 8-byte all-reduce(max) over NCCL; csm_decode_best_key splits it. */
int csm_loop_batch_enqueue(csm_handle h, const csm_loop_query* queries, int nq, int hmax,
                           int query_index_base);
int csm_loop_batch_finish(csm_handle h, csm_result* results, int nq);
int csm_loop_batch(csm_handle h, const csm_loop_query* queries, int nq, int hmax,
                   int query_index_base, csm_result* results);
/* Refinement on the device (ScanMatcherLinearSolver::OptimizePose,
 * scan_matcher_linear_solver.cpp:66-140, as LoopDetectorBranchBound::Detect runs
 * it on every detected loop, loop_detector_branch_bound.cpp:110-135).
 * csm_set_refiner(h, p) makes every following loop batch refine the poses it
 * finds (one more kernel behind the search, results read back with the batch);
 * p == NULL switches it off. csm_loop_batch_finish_refined is
 * csm_loop_batch_finish plus the per-query refinement outcomes.
 * The reference's solver carries its damping factor from one query to the next
 * (a member, :100-104); here every query of a batch starts from p->lambda and
 * reports where it ended, and the adapter carries the last one into the next
 * batch. The factor stays within [1e-8, 1e-4] against Hessian entries of 1e3 and
 * more, so this changes refined poses by far less than the 1e-5 tolerance.
 * Cells of unallocated blocks read as 0.5 (grid_map.cpp:424-436): maps uploaded
 * block-sparse keep their block list for this; for dense uploads a block counts
 * as allocated iff it (16 x 16 cells) holds a non-zero cell. */
int csm_set_refiner(csm_handle h, const csm_refine_params* p);
int csm_loop_batch_finish_refined(csm_handle h, csm_result* results, csm_refined* refined, int nq);
/* Refine n given poses (synchronous). Uses p, not the handle's refiner setting. */
int csm_refine_batch(csm_handle h, const csm_refine_query* queries, int n,
                     const csm_refine_params* p, csm_refined* out);
/* The epilogue of the single-scan matchers on the device: CostSquareError::Cost and
 * ComputeCovariance at the pose the matcher decided on, found or not
 * (scan_matcher_correlative.cpp:203-219, scan_matcher_branch_bound.cpp:241-252;
 * cost_function_square_error.cpp:48-75,131-146), computed behind the match in the same
 * submission instead of a second pass on the CPU. csm_set_epilogue(h, scale > 0) switches it
 * on for the following csm_match_rt / csm_match_bb calls (0 = off); csm_last_epilogue returns
 * the outcome of the last such call: pose = the decided sensor pose, final_cost = the summed
 * squared error there, covariance = scale * inverse Hessian, iterations = 0.
 * With the epilogue off and a refiner set (csm_set_refiner), csm_match_rt / csm_match_bb run that
 * refiner on the pose they find -- the final matcher the front end calls after its scan matcher
 * (lidar_graph_slam_frontend.cpp:216-230, "FinalScanMatcherType": "LinearSolver") -- and
 * csm_last_epilogue returns its outcome (valid == 0 when no pose was found). */
int csm_set_epilogue(csm_handle h, double covariance_scale);
int csm_last_epilogue(csm_handle h, csm_refined* out);
/* ---- the whole first-touch step in one call -----------------------------------
 * What LoopDetectorBranchBound::Detect does with its device context for one batch whose level-0
 * grids are resident: (re)build what the search reads above level 0 for the n_maps maps (drop != 0:
 * as on first touch, even if a previous call built it), enqueue the search batch with its
 * refinement and read-back, and, when the handle has a communicator (csm_comm_init_*), start the
 * exchange of the packed best word behind it. Returns after enqueueing; *ticket (may be NULL)
 * identifies the exchange (csm_comm_best_result). Finish with csm_loop_batch_finish(_refined). */
int csm_detect_step_enqueue(csm_handle h, const int64_t* map_ids, int n_maps, int drop,
                            const csm_loop_query* queries, int nq, int hmax, int query_index_base,
                            int* ticket);

/* ---- multi-GPU: exchange of the packed best word over NCCL --------------------
 * Queries of a Detect are independent (loop_detector_branch_bound.cpp:68) and shard over GPUs like
 * the reference's two accelerator cores (loop_detector_fpga_parallel.cpp:41-56); the only exchange
 * is the 8-byte all-reduce(max) of the packed best word (see csm_best_key_device). The library
 * calls NCCL itself (libnccl.so.2, loaded at run time), on a side stream of the handle behind an
 * event, so that the exchange of one batch overlaps the kernels of the next.
 *  - one process per GPU: rank 0 calls csm_comm_unique_id, hands the 128 bytes to the other
 *    ranks (any side channel), every rank calls csm_comm_init_rank;
 *  - one process, several GPUs: csm_comm_init_all(handles, n), one handle per device, then
 *    csm_comm_allreduce_best_all on all of them at once.
 * csm_comm_allreduce_best starts the exchange of the word the last batch left (returns a ticket,
 * up to 8 may be in flight); csm_comm_best_result waits for it and returns the reduced word. */
int csm_comm_unique_id(void* id128);
int csm_comm_init_rank(csm_handle h, const void* id128, int rank, int world);
int csm_comm_init_all(csm_handle* handles, int n);
int csm_comm_allreduce_best(csm_handle h, int* ticket);
int csm_comm_allreduce_best_all(csm_handle* handles, int n, int* tickets);
/* The same exchange for a word the caller formed on the host (a Detect that ran on several lanes
 * or handles keeps its best word there): no dependence on the compute stream. */
int csm_comm_allreduce_word(csm_handle h, uint64_t word, int* ticket);
int csm_comm_allreduce_words_all(csm_handle* handles, int n, const uint64_t* words, int* tickets);
int csm_comm_best_result(csm_handle h, int ticket, uint64_t* word);
int csm_comm_destroy(csm_handle h);

/* Phase timing: after csm_set_option(h, "timing", 1) the library records a CUDA
 * event on the handle's stream after every kernel of a loop batch (or of a
 * streaming pyramid build). csm_debug_timings waits for the last one and
 * returns the durations in ms of the phases of the LAST such call, with their
 * names separated by ';' in `names`. Returns the number of phases. With
 * "timing" = 2 the marks accumulate across calls (uploads included) until the
 * option is set again, and the values are times since the first mark. */
int csm_debug_timings(csm_handle h, char* names, size_t names_cap, float* ms, int max_n);
/* Debug: size of the node list of every height after the last batch (8 entries):
 * the nodes of that height that passed and were expanded */
int csm_debug_frontier_counts(csm_handle h, unsigned int* out8);
/* Debug: one bound level of the branch-and-bound sweep (csm_bounds.cuh: u8 upper bounds
 * ceil(v / 257) of the sliding 2^level x 2^level maximum, cells outside the map 0), untiled into
 * out[rows * cols]. Builds levels 1..level for this map when they are missing. 1 <= level <= 6. */
int csm_debug_bound_level(csm_handle h, int64_t map_id, int level, uint8_t* out);
/* Debug: the node list of height `level` (q:16 | t:16 | xi:16 | yi:16 per entry) as the last batch left it;
 * with option "bb_stop_level" = level the sweep stops there. Returns the number of entries copied. */
int csm_debug_node_list(csm_handle h, int level, uint64_t* out, int cap);
void* csm_best_key_device(csm_handle h);
void csm_decode_best_key(uint64_t best_key, int64_t* key, int32_t* query_index);

#ifdef __cplusplus
}
#endif

#endif /* CSM_B200_H */
