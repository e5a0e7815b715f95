/* csm_b200.cu -- host side of libcsm_b200.so: handles, device-resident map /
 * scan caches, launch sequences, and the extern "C" entry points declared in
 * include/csm_b200.h. Compile: nvcc -gencode arch=compute_100a,code=sm_100a.
 *
 * There is no CPU path in this file: every entry point enqueues CUDA kernels
 * on the handle's stream or fails with CSM_E_CUDA.
 */

#include "csm_b200.h"
#include "csm_kernels.cuh"
#include "csm_window_tma.cuh"
#include "csm_refine.cuh"
#include "csm_mapbuild.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

using namespace csm;

namespace {

/* NCCL, loaded at run time: inside a process that already carries a libnccl.so.2 (PyTorch brings its
 * own) dlopen returns that one, a plain C++ host gets the system library. Only the handful of entry
 * points the best-word exchange needs; the declarations follow nccl.h (2.x ABI). */
namespace nccl {
typedef struct ncclComm* comm_t;
struct unique_id { char internal[128]; };
constexpr int kUint64 = 5, kMax = 2;       /* ncclUint64, ncclMax */
struct Api
{
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommInitAll)(comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
const Api& api()
{
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (lib == nullptr) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (lib == nullptr) return;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
        a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(lib, "ncclAllReduce"));
        a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(dlsym(lib, "ncclGroupStart"));
        a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommInitAll && a.CommDestroy && a.AllReduce &&
               a.GroupStart && a.GroupEnd && a.GetErrorString;
    });
    return a;
}
} /* namespace nccl */

constexpr int64_t kTempScanId = INT64_MIN;

/* Device allocation shared by the level-0 grids of one batched upload, so
 * that a batch whose host buffers are contiguous moves with a single copy */
struct ArenaBlock
{
    void* p = nullptr;
    cudaStream_t stream = nullptr;
    ~ArenaBlock() { if (p) cudaFreeAsync(p, stream); }
};

/* Block-sparse upload in flight: the allocated blocks of a batch of maps sit in
 * a device staging buffer (filled by the copy stream); the first consumer of
 * any of the maps expands the whole batch into the dense level-0 grids. */
struct BlockScatter
{
    void* stage = nullptr;            /* [block data][block index][prefix], device */
    size_t data_bytes = 0, index_off = 0, prefix_off = 0;
    uint16_t* dense = nullptr;        /* level 0 of the first map; maps are contiguous */
    unsigned char* alloc = nullptr;   /* block-allocation bytes of the first map; maps are contiguous */
    int n_maps = 0, max_count = 0, nb_map = 0;
    int log2bs = 0, block_cols = 0, rows = 0, cols = 0;
    std::vector<int> prefix;          /* host copy, alive until the H2D copy has run */
    bool done = false;
    cudaStream_t stream = nullptr;
    ~BlockScatter() { if (stage) cudaFreeAsync(stage, stream); }
};

struct MapSlot
{
    int rows = 0, cols = 0;
    double res = 0.0, offx = 0.0, offy = 0.0;
    uint16_t* base = nullptr;      /* level 0: what the matchers read (65535 as unknown, see k_saturated_unknown) */
    uint16_t* raw = nullptr;       /* maps made by csm_map_create: the true values the next update continues from */
    std::shared_ptr<ArenaBlock> base_block;   /* owner of `base` when it lives in a batch arena */
    uint16_t* levels = nullptr;    /* levels 1..hmax */
    int levels_alloc = 0;          /* number of levels the allocation holds */
    int hmax = 0;                  /* number of levels currently valid */
    uint16_t* coarse = nullptr;
    int coarse_win = 0;
    uint32_t* wide = nullptr;      /* level 0 as value | known << 20 words: what the wide TMA window kernel loads */
    bool wide_valid = false;
    /* bound levels of the branch-and-bound sweep (csm_bounds.cuh): levels 1..bounds_alloc allocated
     * (padding zeroed once), 1..bounds_levels valid for the current contents */
    unsigned char* bounds = nullptr;
    int bounds_alloc = 0;
    int bounds_levels = -1;        /* -1: nothing valid; L >= 0: serves searches with hmax <= L + 1 */
    unsigned long long mark = 0;   /* scratch: "already taken" in the list being built (compared with csm_context::mark) */
    int margin_slot = -1;          /* this map's word in the handle's low-margin array (k_low_margin) */
    bool margin_valid = false;     /* ... holds the value of the current contents */
    cudaEvent_t pending_upload = nullptr;   /* copy-stream event the next consumer must wait for */
    std::shared_ptr<BlockScatter> pending_scatter;   /* block-sparse upload not yet expanded */
    /* which blocks are allocated in the reference's sense (refinement reads unallocated cells as 0.5):
     * one byte per block, from the block list of a block-sparse upload or derived from the cells */
    unsigned char* alloc = nullptr;
    std::shared_ptr<ArenaBlock> alloc_block;  /* owner of `alloc` when it lives in a batch arena */
    int alloc_log2bs = 0, alloc_bytes = 0;
    bool alloc_valid = false;
};

struct ScanSlot
{
    void* block = nullptr;         /* one device allocation: angles | ranges | trig */
    double* angles = nullptr;
    double* ranges = nullptr;
    double2* trig = nullptr;       /* (cos a_i, sin a_i) */
    int n = 0;
    double max_range = 0.0;
    /* host copy: a result flagged CSM_FLAG_FP_MARGIN is recomputed from indices evaluated on the host
     * with the reference's own libm calls (exact_rerun_*) */
    std::vector<double> h_angles, h_ranges;
};

/* What finish_results needs to know about a batch in flight to post-process flagged results */
struct BatchRecord
{
    std::vector<csm_loop_query> queries;
    int hmax = 0;
    bool inline_scan = false;
    std::vector<double> angles, ranges;      /* the scan that arrived with the call (inline_scan) */
};

/* Where the small per-batch arrays live inside the plan buffer (d_plan):
 * a region pulled from page-locked host memory followed by a zeroed region. */
struct PlanView
{
    size_t off_queries = 0, off_thetas = 0, off_inc = 0, off_rootoff = 0, off_extra = 0, off_scan = 0;
    size_t pulled_bytes = 0, zero_off = 0, zero_bytes = 0, trig_off = 0, total = 0;
    DevQuery* queries = nullptr;
    double* thetas = nullptr;
    unsigned long long* inc = nullptr;
    unsigned int* rootoff = nullptr;
    char* extra = nullptr;
    ScanSlot scan;                 /* scan arriving with the call (single-scan matchers) */
    int* qflags = nullptr;
    unsigned long long* tiekey = nullptr;   /* per query: largest key shared by two candidates */
    unsigned long long* probe = nullptr;    /* per query: start nodes of the incumbent dive (k_bbg_dive) */
    int* stats = nullptr;
    unsigned int* counts = nullptr;
    int* overflow = nullptr;
};

struct DevBuf
{
    void* p = nullptr;
    size_t bytes = 0;
};

} /* namespace */

struct csm_context
{
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;        /* compute stream */
    cudaStream_t copy_stream = nullptr;   /* host-to-device uploads of grids, overlaps compute */
    bool owns_copy_stream = true;         /* false: borrowed from another handle (csm_share_copy_stream) */
    cudaEvent_t upload_events[16] = { nullptr };
    int upload_event_next = 0;
    cudaEvent_t compute_mark = nullptr;
    bool upload_open = false;             /* uploads enqueued since the last event record */
    std::string err;
    int64_t launches = 0;
    std::unordered_map<int64_t, MapSlot> maps;
    std::unordered_map<int64_t, ScanSlot> scans;

    /* workspaces, grown on demand */
    DevBuf d_plan, d_proj, d_rcs, d_results, d_bestkey;
    DevBuf d_list[2];
    DevBuf d_rtblocks, d_pyrjobs, d_rootkey, d_wtgroups, d_bljobs;
    int bb_capacity = 0;           /* test knob: upper limit of the frontier lists (entries), 0 = automatic */
    int pyramid_segs = 0;          /* test knob: row segments per map of the streaming builder (0 = automatic) */
    int bounds_mode = 0;           /* builder of the bound levels: 0 auto, 1 k_bounds_build, 2 the streaming kernel */
    int bb_stop_level = 0;         /* debug: the sweep stops once list(bb_stop_level) is complete (csm_debug_node_list) */
    int bb_bounds = 1;             /* 1: batched searches sweep the u8 bound levels (csm_bounds.cuh), 0: the u16 levels */
    int bbx_ctas_per_sm = 0;
    int bb_sweep_ctas_cap = 0;     /* option (A/B runs): at most this many sweep CTAs per SM (0 = what fits) */
    int bb_ctas_per_sm = 0;        /* resident CTAs per SM of the B&B sweep kernels (occupancy query, lazily) */
    int bb_split_shift = 0;
    int bb_skip_top = 1;           /* 1: the B&B sweep starts one height below hmax (same results) */
    int bb_probe = 1;              /* the group sweep descends greedily to a leaf per query (k_bbg_dive: same results,
                                      fewer nodes): 1 = after the launch of height 4, 2 = after heights 5 and 4, 0 = never.
                                      Measured on cfg3 (256 queries): 625 k / 719 k / 707 k queries/s for 0 / 1 / 2 */
    int window_mode = 0;           /* grid search, integer-shift path: 0 auto (TMA tiles when possible),
                                      1 plain global-memory kernel, 2 require the TMA kernel */
    PlanView plan_view;                   /* layout of the last staged batch */
    unsigned int frontier_capacity = 0;
    /* pinned staging: eight upload areas used in turn (an area is reused
     * only after the copies that read it have completed) + one result area */
    static constexpr int kUploadAreas = 8;
    void* h_up[kUploadAreas] = { nullptr };
    size_t h_up_bytes[kUploadAreas] = { 0 };
    cudaEvent_t h_up_done[kUploadAreas] = { nullptr };
    int h_up_next = 0;
    /* result areas (pinned): one per batch in flight, used in turn */
    static constexpr int kResultSlots = 4;
    void* h_res[kResultSlots] = { nullptr, nullptr, nullptr, nullptr };
    size_t h_res_bytes[kResultSlots] = { 0, 0, 0, 0 };
    cudaEvent_t h_res_done[kResultSlots] = { nullptr, nullptr, nullptr, nullptr };
    int res_head = 0;              /* slot of the oldest batch in flight */
    int res_count = 0;             /* batches in flight */
    int res_nq[kResultSlots] = { 0, 0, 0, 0 };
    BatchRecord res_rec[kResultSlots];
    void* h_exact = nullptr;       /* pinned: result of an exact rerun / low-margin scan */
    /* exchange of the packed best word over NCCL (csm_comm_*): a ring of words, each copied out of
     * d_bestkey on the compute stream, reduced in place on a high-priority side stream, read back */
    static constexpr int kCommRing = 8;
    nccl::comm_t comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    cudaStream_t comm_stream = nullptr;
    unsigned long long* d_words = nullptr;        /* kCommRing device words */
    unsigned long long* h_words = nullptr;        /* kCommRing pinned words */
    cudaEvent_t comm_ready[kCommRing] = { nullptr };
    cudaEvent_t comm_done[kCommRing] = { nullptr };
    int comm_next = 0;
    int64_t exact_reruns = 0;      /* flagged results recomputed exactly so far */
    int exact_rerun = 1;           /* option: recompute results whose projection raised the FP guard-band flag */
    double fp_margin_scale = 1.0;  /* option (tests): multiplies the guard band */
    int saturated_unknown = 1;     /* option: a cell at 65535 reads as unknown, like in the compiled reference */
    /* map construction (csm_map_*): update tables, ray / event workspaces, error word */
    DevBuf d_maptables, d_mapwork, d_maperror;
    bool map_tables_set = false;
    int* h_maperror = nullptr;     /* pinned copy of the error word, read by the next synchronising call */
    DevBuf d_margin, d_marginjobs; /* low-margin words of the maps (one int each), job table of the scan */
    int margin_slots = 0;
    unsigned long long mark = 0;   /* bumped for every list that must hold each map once */
    bool res_refined[kResultSlots] = { false, false, false, false };
    /* last pyramid job table on the device (skips the re-upload when unchanged) */
    std::vector<PyrJob> jobs_on_device;
    /* options (csm_set_option) */
    int pyramid_mode = 0;          /* 0 auto, 1 level-by-level, 2 streaming, 3 streaming with shared-memory rings */
    int bb_dive = 2;               /* a beam dive per query seeds the incumbents before the level sweep:
                                      0 never, 1 always, 2 only for calls of at most 4 queries (there the
                                      latency of the dive is small against the nodes it saves) */
    int accumulate_best_key = 0;   /* 1: batches do not reset the packed best word */
    double epilogue_scale = 0.0;   /* > 0: single-scan matches also return cost and covariance at the
                                      pose they decide on (csm_set_epilogue), computed behind k_finalize */
    csm_refined last_epilogue {};
    bool last_epilogue_set = false;
    bool refine_on = false;        /* loop batches refine the poses they find (csm_set_refiner) */
    csm_refine_params refine {};
    DevBuf d_refine_in;            /* csm_refine_batch: queries and start poses */
    DevBuf d_allocjobs;            /* ensure_alloc: job table */
    /* last (thresholds, beams) -> integer cut-offs (fill_common) */
    int memo_n = -1, memo_nk_cut = 0;
    double memo_score_thr = -1.0, memo_known_thr = -1.0;
    KeyThreshold memo_kthr {};
    /* "timing" option: CUDA events between the phases of the last loop batch / pyramid build */
    int timing = 0;
    std::vector<cudaEvent_t> tev;
    std::vector<std::string> tnames;
    size_t tcount = 0;
};

namespace {

#define CSM_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);       \
            return CSM_E_CUDA;                                                 \
        }                                                                      \
    } while (0)

#define CSM_LAUNCH_CHECK()                                                     \
    do {                                                                       \
        ++h->launches;                                                         \
        cudaError_t e_ = cudaGetLastError();                                   \
        if (e_ != cudaSuccess) {                                               \
            h->err = std::string("kernel launch: ") + cudaGetErrorString(e_);  \
            return CSM_E_CUDA;                                                 \
        }                                                                      \
    } while (0)

/* timing: mark the end of a phase on the compute stream (no-op unless enabled) */
void phase_mark(csm_handle h, const char* name, cudaStream_t on = nullptr)
{
    if (!h->timing)
        return;
    if (h->timing == 1 && std::strcmp(name, "start") == 0)
        h->tcount = 0;             /* mode 1: keep the phases of the last call only */
    if (h->tcount >= 256)
        return;
    if (h->tcount >= h->tev.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess)
            return;
        h->tev.push_back(e);
        h->tnames.emplace_back();
    }
    h->tnames[h->tcount] = name;
    cudaEventRecord(h->tev[h->tcount], on ? on : h->stream);
    ++h->tcount;
}

int fail(csm_handle h, int code, const std::string& msg)
{
    h->err = msg;
    return code;
}

int ensure(csm_handle h, DevBuf& b, size_t bytes)
{
    if (b.bytes >= bytes && b.p != nullptr)
        return CSM_OK;
    if (b.p != nullptr)
        CSM_CUDA(cudaFreeAsync(b.p, h->stream));
    b.p = nullptr;
    b.bytes = 0;
    size_t want = std::max(bytes, (size_t)256);
    want = (want + 255) & ~(size_t)255;
    CSM_CUDA(cudaMallocAsync(&b.p, want, h->stream));
    b.bytes = want;
    return CSM_OK;
}

/* Pinned upload area that is safe to overwrite. Call upload_committed() after
 * enqueueing the copies that read it. */
int acquire_upload(csm_handle h, size_t bytes, char** out)
{
    const int k = h->h_up_next;
    h->h_up_next = (h->h_up_next + 1) % csm_context::kUploadAreas;
    if (h->h_up_done[k] == nullptr)
        CSM_CUDA(cudaEventCreateWithFlags(&h->h_up_done[k], cudaEventDisableTiming));
    else
        CSM_CUDA(cudaEventSynchronize(h->h_up_done[k]));
    if (h->h_up_bytes[k] < bytes) {
        if (h->h_up[k] != nullptr)
            CSM_CUDA(cudaFreeHost(h->h_up[k]));
        h->h_up[k] = nullptr;
        h->h_up_bytes[k] = 0;
        const size_t want = std::max(bytes * 2, (size_t)1 << 18);
        CSM_CUDA(cudaHostAlloc(&h->h_up[k], want, cudaHostAllocDefault));
        h->h_up_bytes[k] = want;
    }
    *out = static_cast<char*>(h->h_up[k]);
    return CSM_OK;
}

int upload_committed(csm_handle h)
{
    const int k = (h->h_up_next + csm_context::kUploadAreas - 1) % csm_context::kUploadAreas;
    CSM_CUDA(cudaEventRecord(h->h_up_done[k], h->stream));
    return CSM_OK;
}

/* dst (device, 16-byte aligned, capacity rounded up to 16) <- src (pinned host) */
int pull_to_device(csm_handle h, void* dst, const void* src_pinned, size_t bytes)
{
    const unsigned int n16 = (unsigned int)((bytes + 15) / 16);
    if (n16 == 0)
        return CSM_OK;
    const unsigned int blocks = std::min<unsigned int>((n16 + 255) / 256, 64u);
    k_pull<<<blocks, 256, 0, h->stream>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src_pinned), n16);
    CSM_LAUNCH_CHECK();
    return CSM_OK;
}

/* Enqueue the device-to-host copy of the nq results (and the overflow flag)
 * of the batch just launched into the next free pinned result area. */
size_t refined_offset(int nq) { return sizeof(csm_result) * (size_t)nq + 16; }

int enqueue_readback(csm_handle h, int nq, bool refined = false)
{
    if (h->res_count >= csm_context::kResultSlots)
        return fail(h, CSM_E_CAPACITY, "too many loop batches in flight: call csm_loop_batch_finish");
    const int k = (h->res_head + h->res_count) % csm_context::kResultSlots;
    const size_t bytes = refined_offset(nq) + sizeof(csm_refined) * (size_t)nq + 64;
    if (h->h_res_bytes[k] < bytes) {
        if (h->h_res[k] != nullptr)
            CSM_CUDA(cudaFreeHost(h->h_res[k]));
        h->h_res[k] = nullptr;
        h->h_res_bytes[k] = 0;
        const size_t want = std::max(bytes * 2, (size_t)1 << 16);
        CSM_CUDA(cudaHostAlloc(&h->h_res[k], want, cudaHostAllocDefault));
        h->h_res_bytes[k] = want;
    }
    if (h->h_res_done[k] == nullptr)
        CSM_CUDA(cudaEventCreateWithFlags(&h->h_res_done[k], cudaEventDisableTiming));
    /* results followed by the overflow flag (written by k_finalize): one copy */
    /* ... and by the refinement outcomes when the batch refined its poses */
    CSM_CUDA(cudaMemcpyAsync(h->h_res[k], h->d_results.p,
                             refined_offset(nq) + (refined ? sizeof(csm_refined) * (size_t)nq : 0),
                             cudaMemcpyDeviceToHost, h->stream));
    h->res_refined[k] = refined;
    CSM_CUDA(cudaEventRecord(h->h_res_done[k], h->stream));
    h->res_nq[k] = nq;
    ++h->res_count;
    return CSM_OK;
}

int postprocess_flags(csm_handle h, const BatchRecord& rec, csm_result* results, int nq, csm_refined* refined);

/* Wait for the oldest batch in flight and hand out its results */
int finish_results(csm_handle h, csm_result* results, int nq, csm_refined* refined = nullptr)
{
    if (h->res_count <= 0)
        return fail(h, CSM_E_INVALID, "no batch in flight");
    const int k = h->res_head;
    if (nq != h->res_nq[k])
        return fail(h, CSM_E_INVALID, "loop batch: finish does not match the enqueued batch");
    h->res_head = (h->res_head + 1) % csm_context::kResultSlots;
    --h->res_count;
    CSM_CUDA(cudaEventSynchronize(h->h_res_done[k]));
    const char* hp = static_cast<const char*>(h->h_res[k]);
    std::memcpy(results, hp, sizeof(csm_result) * (size_t)nq);
    if (refined != nullptr) {
        if (h->res_refined[k])
            std::memcpy(refined, hp + refined_offset(nq), sizeof(csm_refined) * (size_t)nq);
        else
            std::memset(refined, 0, sizeof(csm_refined) * (size_t)nq);
    }
    if (*reinterpret_cast<const int*>(hp + sizeof(csm_result) * (size_t)nq) != 0)
        return fail(h, CSM_E_CAPACITY, "branch-and-bound frontier overflow; split the batch");
    /* flagged results: exact rerun for the FP guard band, a look at the map for the low-edge flag */
    bool any = false;
    for (int q = 0; q < nq && !any; ++q)
        any = (results[q].flags & CSM_FLAG_FP_MARGIN) != 0;
    if (any && (int)h->res_rec[k].queries.size() == nq)
        return postprocess_flags(h, h->res_rec[k], results, nq, h->res_refined[k] ? refined : nullptr);
    return CSM_OK;
}

/* Dense uploads: cells at 65535 become unknown in the matchers' copy (k_saturated_unknown), behind the copy
 * on the same stream. Device allocations are 256-byte aligned; arena slots are whole maps of an even number of
 * columns, so every map starts on a 4-byte boundary at least: the 16-byte path needs 16. */
int saturated_pass(csm_handle h, uint16_t* cells, size_t bytes, cudaStream_t stream)
{
    if (!h->saturated_unknown || bytes == 0)
        return CSM_OK;
    if (bytes % 4 != 0 || (reinterpret_cast<uintptr_t>(cells) & 15u) != 0)
        return fail(h, CSM_E_UNSUPPORTED, "grid: the map must start on a 16-byte boundary and hold an even number of cells");
    const size_t n_words = bytes / 4;
    const unsigned blocks = (unsigned)std::max<size_t>(1, std::min<size_t>((n_words / 4 + 255) / 256, 148 * 8));
    k_saturated_unknown<<<blocks, 256, 0, stream>>>(reinterpret_cast<unsigned int*>(cells), n_words);
    CSM_LAUNCH_CHECK();
    ++h->launches;
    return CSM_OK;
}

void free_map(csm_handle h, MapSlot& m)
{
    if (m.base && !m.base_block) cudaFreeAsync(m.base, h->stream);
    if (m.raw) cudaFreeAsync(m.raw, h->stream);
    if (m.levels) cudaFreeAsync(m.levels, h->stream);
    if (m.coarse) cudaFreeAsync(m.coarse, h->stream);
    if (m.wide) cudaFreeAsync(m.wide, h->stream);
    if (m.bounds) cudaFreeAsync(m.bounds, h->stream);
    if (m.alloc && !m.alloc_block) cudaFreeAsync(m.alloc, h->stream);
    m = MapSlot();
}

void free_scan(csm_handle h, ScanSlot& s)
{
    if (s.block) cudaFreeAsync(s.block, h->stream);
    s = ScanSlot();
}

/* Uploads from host memory run on the copy stream so that they overlap the
 * kernels of maps uploaded earlier. close_upload_group() records one event
 * for everything uploaded since the previous group; consumers make the
 * compute stream wait for the events of the maps they touch. */
int close_upload_group(csm_handle h)
{
    if (!h->upload_open)
        return CSM_OK;
    cudaEvent_t& ev = h->upload_events[h->upload_event_next];
    h->upload_event_next = (h->upload_event_next + 1) & 15;
    if (ev == nullptr)
        CSM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CSM_CUDA(cudaEventRecord(ev, h->copy_stream));
    for (auto& kv : h->maps)
        if (kv.second.pending_upload == reinterpret_cast<cudaEvent_t>(1))
            kv.second.pending_upload = ev;
    h->upload_open = false;
    return CSM_OK;
}

int wait_uploads(csm_handle h, const std::vector<MapSlot*>& slots)
{
    int rc = close_upload_group(h);
    if (rc) return rc;
    cudaEvent_t seen[16];
    int nseen = 0;
    for (MapSlot* m : slots) {
        cudaEvent_t ev = m->pending_upload;
        if (ev == nullptr)
            continue;
        m->pending_upload = nullptr;
        bool dup = false;
        for (int i = 0; i < nseen; ++i) dup = dup || seen[i] == ev;
        if (dup) continue;
        if (nseen < 16) seen[nseen++] = ev;
        CSM_CUDA(cudaStreamWaitEvent(h->stream, ev, 0));
    }
    for (MapSlot* m : slots) {
        if (!m->pending_scatter)
            continue;
        std::shared_ptr<BlockScatter> bs = std::move(m->pending_scatter);
        m->pending_scatter.reset();
        if (bs->done)
            continue;
        bs->done = true;
        const size_t map_cells = (size_t)bs->rows * bs->cols;
        CSM_CUDA(cudaMemsetAsync(bs->dense, 0, map_cells * sizeof(uint16_t) * bs->n_maps, h->stream));
        if (bs->alloc != nullptr)
            CSM_CUDA(cudaMemsetAsync(bs->alloc, 0, (size_t)bs->nb_map * bs->n_maps, h->stream));
        if (bs->max_count > 0) {
            ScatterArgs A;
            A.data = static_cast<const uint4*>(bs->stage);
            A.index = reinterpret_cast<const int*>(static_cast<const char*>(bs->stage) + bs->index_off);
            A.prefix = reinterpret_cast<const int*>(static_cast<const char*>(bs->stage) + bs->prefix_off);
            A.dense = bs->dense;
            A.alloc = bs->alloc; A.nb_map = bs->nb_map;
            A.log2bs = bs->log2bs; A.block_cols = bs->block_cols; A.cols = bs->cols;
            A.map_cells = map_cells;
            A.saturated_unknown = h->saturated_unknown;
            const int chunks = (1 << bs->log2bs) * ((1 << bs->log2bs) >> 3);
            const long long work = (long long)bs->max_count * chunks;
            dim3 grid((unsigned)std::min<long long>((work + 255) / 256, 64), (unsigned)bs->n_maps);
            k_scatter_blocks<<<grid, 256, 0, h->stream>>>(A);
            CSM_LAUNCH_CHECK();
        }
        CSM_CUDA(cudaFreeAsync(bs->stage, h->stream));
        bs->stage = nullptr;
        if (h->timing == 2) phase_mark(h, "expand blocks");
    }
    return CSM_OK;
}

/* Before a slot is replaced or freed: transfers still in flight into it (a copy-stream upload, a
 * block-sparse batch that has not been expanded and shares its arena with sibling maps) are ordered
 * before whatever the compute stream does next, and the pending expansion runs now. */
int settle_slot(csm_handle h, MapSlot& m)
{
    if (m.pending_upload == nullptr && !m.pending_scatter)
        return CSM_OK;
    return wait_uploads(h, std::vector<MapSlot*>{ &m });
}

/* The n level-0 grids of a batch share one device arena (so that a contiguous
 * host batch moves with one copy and a block-sparse batch is cleared with one
 * memset). The arena of a previous identical batch is reused. */
int bind_batch_arena(csm_handle h, int n, const int64_t* map_ids, int rows, int cols,
                     std::vector<MapSlot*>& slots, bool& fresh_alloc)
{
    const size_t bytes = (size_t)rows * cols * sizeof(uint16_t);
    slots.resize(n);
    for (int i = 0; i < n; ++i)
        slots[i] = &h->maps[map_ids[i]];
    bool reuse = slots[0]->base_block != nullptr && slots[0]->base == slots[0]->base_block->p;
    for (int i = 0; i < n && reuse; ++i)
        reuse = slots[i]->base_block == slots[0]->base_block && slots[i]->rows == rows &&
                slots[i]->cols == cols &&
                reinterpret_cast<char*>(slots[i]->base) == static_cast<char*>(slots[0]->base_block->p) + (size_t)i * bytes;
    fresh_alloc = false;
    for (int i = 0; i < n; ++i) {
        const int src = settle_slot(h, *slots[i]);
        if (src) return src;
    }
    if (!reuse) {
        auto block = std::make_shared<ArenaBlock>();
        block->stream = h->stream;
        CSM_CUDA(cudaMallocAsync(&block->p, bytes * n, h->stream));
        for (int i = 0; i < n; ++i) {
            MapSlot& m = *slots[i];
            const bool keep_levels = m.rows == rows && m.cols == cols;
            uint16_t* levels = keep_levels ? m.levels : nullptr;
            const int levels_alloc = keep_levels ? m.levels_alloc : 0;
            uint16_t* coarse = keep_levels ? m.coarse : nullptr;
            unsigned char* bounds = keep_levels ? m.bounds : nullptr;
            const int bounds_alloc = keep_levels ? m.bounds_alloc : 0;
            if (keep_levels) { m.levels = nullptr; m.coarse = nullptr; m.bounds = nullptr; }
            free_map(h, m);
            m.levels = levels; m.levels_alloc = levels_alloc; m.coarse = coarse;
            m.bounds = bounds; m.bounds_alloc = bounds_alloc;
            m.rows = rows; m.cols = cols;
            m.base_block = block;
            m.base = reinterpret_cast<uint16_t*>(static_cast<char*>(block->p) + (size_t)i * bytes);
        }
        fresh_alloc = true;
    }
    return CSM_OK;
}

/* Block-allocation bytes for the refinement stage. Block-sparse uploads bring them (set by
 * k_scatter_blocks); for dense uploads they are derived here, once per upload: a 16 x 16 block
 * counts as allocated iff it holds a non-zero cell. All maps that need it go in ONE launch (their
 * bytes share one allocation). The maps' uploads must have been waited for. */
int ensure_alloc(csm_handle h, const std::vector<MapSlot*>& slots)
{
    const int k = 4;
    std::vector<MapSlot*> todo;
    size_t total = 0;
    int max_blocks = 0;
    for (MapSlot* m : slots) {
        if (m->alloc_valid || std::find(todo.begin(), todo.end(), m) != todo.end())
            continue;
        todo.push_back(m);
        const int bytes = ((m->rows + 15) >> k) * ((m->cols + 15) >> k);
        total += ((size_t)bytes + 15) & ~(size_t)15;
        max_blocks = std::max(max_blocks, bytes);
    }
    if (todo.empty())
        return CSM_OK;
    auto block = std::make_shared<ArenaBlock>();
    block->stream = h->stream;
    CSM_CUDA(cudaMallocAsync(&block->p, total, h->stream));
    std::vector<AllocJob> jobs(todo.size());
    size_t off = 0;
    for (size_t i = 0; i < todo.size(); ++i) {
        MapSlot* m = todo[i];
        const int bytes = ((m->rows + 15) >> k) * ((m->cols + 15) >> k);
        if (m->alloc && !m->alloc_block) CSM_CUDA(cudaFreeAsync(m->alloc, h->stream));
        m->alloc_block = block;
        m->alloc = static_cast<unsigned char*>(block->p) + off;
        m->alloc_bytes = bytes;
        m->alloc_log2bs = k;
        jobs[i] = AllocJob { m->base, m->alloc, m->rows, m->cols };
        off += ((size_t)bytes + 15) & ~(size_t)15;
    }
    const size_t jb = sizeof(AllocJob) * jobs.size();
    int rc = ensure(h, h->d_allocjobs, jb);
    if (rc) return rc;
    char* hp = nullptr;
    if ((rc = acquire_upload(h, jb, &hp))) return rc;
    std::memcpy(hp, jobs.data(), jb);
    if ((rc = pull_to_device(h, h->d_allocjobs.p, hp, jb))) return rc;
    if ((rc = upload_committed(h))) return rc;
    dim3 grid((unsigned)std::max(1, std::min((max_blocks + 7) / 8, 64)), (unsigned)jobs.size());
    k_block_alloc<<<grid, 256, 0, h->stream>>>(static_cast<const AllocJob*>(h->d_allocjobs.p), k);
    CSM_LAUNCH_CHECK();
    for (MapSlot* m : todo)
        m->alloc_valid = true;
    return CSM_OK;
}

/* `normalized > thr` on integer keys, see csm_device.cuh */
KeyThreshold make_key_threshold(double thr, int n)
{
    KeyThreshold k;
    k.thr = thr;
    if (thr < 0.0) {
        k.fail_max = -1; k.pass_min = 0;
    } else if (thr == 0.0) {
        k.fail_max = 0; k.pass_min = 1;
    } else {
        const long double t = (long double)thr * (long double)n * 65534000.0L;
        const long long cut = (long long)floorl(t);
        k.fail_max = cut - 1;
        k.pass_min = cut + 2;
    }
    return k;
}

/* `double(nKnown) / double(n) > thr` passes iff nKnown > cut */
int make_known_cut(double thr, int n)
{
    int cut = -1;
    for (int k = 0; k <= n; ++k)
        if ((double)k / (double)n <= thr)
            cut = k;
    return cut;
}

double fp_margin(const double pose[3], const MapSlot& m, double max_range, double extent)
{
    /* Guard band (cells) around cell boundaries inside which device and glibc
     * trigonometry (a few ulp apart) or per-node re-projection rounding could
     * give different floor() results. Bounds every intermediate magnitude. */
    const double mag = std::fabs(pose[0]) + std::fabs(pose[1]) + std::fabs(m.offx) +
                       std::fabs(m.offy) + max_range + extent +
                       (double)std::max(m.rows, m.cols) * m.res;
    return 1024.0 * 1.1102230246251565e-16 * mag / m.res + 1e-12;
}

/* ---- exact reruns ------------------------------------------------------------------------------------
 * The device forms cos / sin(theta + a) from an angle-addition formula and multiplies by 1 / res, a few ulp
 * away from the reference's glibc calls and its division; only a hit point within the guard band of a cell
 * boundary can land in another cell because of that, and the projection flags exactly those. A flagged
 * result is recomputed here from what the HOST evaluates with the reference's own libm calls in the
 * reference's operation order (sensor_data.hpp:190-203, grid_map_geometry.cpp:113-122): r cos(theta + a),
 * r sin(theta + a) per (angle, beam); the device then repeats the reference's per-candidate arithmetic
 * ((x + r cos) - off) / res with IEEE division (k_grid_general) over the whole candidate lattice, scoring
 * on the GPU as always. Rare (about 4e-10 per angle and beam), so the exhaustive lattice is affordable. */
double2 host_hit_terms(double theta, double angle, double range)
{
    /* HitPoint: cosTheta = cos(theta + angle), x = sensor.x + range * cosTheta */
    const double c = std::cos(theta + angle), s = std::sin(theta + angle);
    return make_double2(range * c, range * s);
}

int ensure_exact_area(csm_handle h)
{
    if (h->h_exact == nullptr)
        CSM_CUDA(cudaHostAlloc(&h->h_exact, 4096, cudaHostAllocDefault));
    return CSM_OK;
}

int launch_setup(csm_handle h, const SetupArgs& A)
{
    const unsigned int work = std::max(std::max(A.n16, A.z16), (unsigned int)std::max(A.n_beams, 1));
    const unsigned int blocks = std::min<unsigned int>((work + 255) / 256, 64u);
    k_setup<<<blocks, 256, 0, h->stream>>>(A);
    CSM_LAUNCH_CHECK();
    return CSM_OK;
}

int upload_scan_impl(csm_handle h, int64_t scan_id, const double* angles,
                     const double* ranges, int n)
{
    if (n <= 0 || n > kMaxBeams || angles == nullptr || ranges == nullptr)
        return fail(h, CSM_E_INVALID, "scan: need 1 <= n <= 4096 beams");
    ScanSlot& s = h->scans[scan_id];
    const size_t half = ((sizeof(double) * n + 15) / 16) * 16;
    if (s.n != n) {
        free_scan(h, s);
        CSM_CUDA(cudaMallocAsync(&s.block, 2 * half + sizeof(double2) * n, h->stream));
        s.angles = static_cast<double*>(s.block);
        s.ranges = reinterpret_cast<double*>(static_cast<char*>(s.block) + half);
        s.trig = reinterpret_cast<double2*>(static_cast<char*>(s.block) + 2 * half);
        s.n = n;
    }
    /* staged through pinned memory and pulled by one kernel that also fills the trig table */
    char* hp = nullptr;
    int rc = acquire_upload(h, 2 * half, &hp);
    if (rc) return rc;
    std::memcpy(hp, angles, sizeof(double) * n);
    std::memcpy(hp + half, ranges, sizeof(double) * n);
    SetupArgs A;
    std::memset(&A, 0, sizeof(A));
    A.dst = static_cast<uint4*>(s.block);
    A.src_host = reinterpret_cast<const uint4*>(hp);
    A.n16 = (unsigned int)(2 * half / 16);
    A.angles_host = reinterpret_cast<const double*>(hp);
    A.trig = s.trig;
    A.n_beams = n;
    if ((rc = launch_setup(h, A))) return rc;
    if ((rc = upload_committed(h))) return rc;
    s.max_range = *std::max_element(ranges, ranges + n);
    s.h_angles.assign(angles, angles + n);
    s.h_ranges.assign(ranges, ranges + n);
    return CSM_OK;
}

int build_levels(csm_handle h, const std::vector<MapSlot*>& slots, int hmax)
{
    if (hmax < 0 || hmax >= kMaxLevels)
        return fail(h, CSM_E_UNSUPPORTED, "pyramid: 0 <= hmax <= 7");
    std::vector<PyrJob> jobs;
    int max_rows = 0, max_cols = 0;
    {
        const int wrc = wait_uploads(h, slots);
        if (wrc) return wrc;
    }
    for (MapSlot* m : slots) {
        if (m->hmax >= hmax)
            continue;
        if (m->levels_alloc < hmax) {
            if (m->levels) CSM_CUDA(cudaFreeAsync(m->levels, h->stream));
            m->levels = nullptr;
            m->levels_alloc = 0;
            const size_t bytes = (size_t)hmax * m->rows * m->cols * sizeof(uint16_t);
            CSM_CUDA(cudaMallocAsync((void**)&m->levels, bytes, h->stream));
            m->levels_alloc = hmax;
        }
        m->hmax = hmax;
        jobs.push_back(PyrJob { m->base, m->levels, m->rows, m->cols });
        max_rows = std::max(max_rows, m->rows);
        max_cols = std::max(max_cols, m->cols);
    }
    if (jobs.empty())
        return CSM_OK;
    const size_t jb = jobs.size() * sizeof(PyrJob);
    const bool same = h->jobs_on_device.size() == jobs.size() &&
                      std::memcmp(h->jobs_on_device.data(), jobs.data(), jb) == 0;
    if (!same) {
        int rc = ensure(h, h->d_pyrjobs, jb);
        if (rc) return rc;
        char* hp = nullptr;
        if ((rc = acquire_upload(h, jb, &hp))) return rc;
        std::memcpy(hp, jobs.data(), jb);
        if ((rc = pull_to_device(h, h->d_pyrjobs.p, hp, jb))) return rc;
        if ((rc = upload_committed(h))) return rc;
        h->jobs_on_device = jobs;
    }
    /* Batches of maps that fit its layout take the streaming builder (one pass,
     * level 0 read once, every level written once). */
    bool stream_ok = hmax >= 1 && hmax <= 6 && h->pyramid_mode != 1 &&
                     (h->pyramid_mode >= 2 || jobs.size() >= 8);
    for (const PyrJob& j : jobs)
        stream_ok = stream_ok && j.cols <= 512 && (j.cols % 8) == 0 && (j.rows % kPsRows) == 0;
    if (stream_ok) {
        phase_mark(h, "start");
        /* two CTAs fit an SM: split every map into row segments until the grid fills them */
        int min_rows = jobs[0].rows;
        bool rows32 = true;
        for (const PyrJob& j : jobs) {
            min_rows = std::min(min_rows, j.rows);
            rows32 = rows32 && (j.rows % (kPsRows * kPs2Group)) == 0;
        }
        int segs = (int)std::min<size_t>(4, (size_t)(2 * h->sm_count) / jobs.size());
        segs = std::max(1, std::min(segs, min_rows / 128));         /* segments of at least 128 rows */
        const unsigned int grid = (unsigned)(jobs.size() * segs);
        const PyrJob* dj = static_cast<const PyrJob*>(h->d_pyrjobs.p);
        if (rows32 && h->pyramid_mode != 3) {
            /* rings in registers (main loop unrolled over 32 rows) */
            const size_t smem2 = sizeof(unsigned int) * ((size_t)kPsStages * kPsRows * kPsInStride + 2 * kPsRows * 256);
            bool sq512 = true;          /* the common submap size gets immediates for every stride */
            for (const PyrJob& j : jobs) sq512 = sq512 && j.rows == 512 && j.cols == 512;
#define CSM_PS2_LAUNCH(HM)                                                                             \
            if (sq512) k_pyramid_stream2<HM, 512, false><<<grid, kPsThreads, smem2, h->stream>>>(dj, segs);   \
            else k_pyramid_stream2<HM, 0, false><<<grid, kPsThreads, smem2, h->stream>>>(dj, segs);
            switch (hmax) {
            case 1: CSM_PS2_LAUNCH(1) break;
            case 2: CSM_PS2_LAUNCH(2) break;
            case 3: CSM_PS2_LAUNCH(3) break;
            case 4: CSM_PS2_LAUNCH(4) break;
            case 5: CSM_PS2_LAUNCH(5) break;
            default: CSM_PS2_LAUNCH(6) break;
            }
#undef CSM_PS2_LAUNCH
            CSM_LAUNCH_CHECK();
            phase_mark(h, "k_pyramid_stream");
            return CSM_OK;
        }
        const size_t smem = sizeof(unsigned int) *
            ((size_t)kPsStages * kPsRows * kPsInStride + 2 * kPsRows * 256 + 63 * 256);
        k_pyramid_stream<<<grid, kPsThreads, smem, h->stream>>>(dj, hmax, segs);
        CSM_LAUNCH_CHECK();
        phase_mark(h, "k_pyramid_stream");
        return CSM_OK;
    }
    /* Otherwise level by level; maps are processed in chunks so that level h-1
     * of a chunk is still in L2 when level h reads it. */
    const size_t map_bytes = (size_t)max_rows * max_cols * sizeof(uint16_t);
    const size_t chunk = std::max<size_t>(1, ((size_t)24 << 20) / std::max<size_t>(map_bytes, 1));
    for (size_t first = 0; first < jobs.size(); first += chunk) {
        const size_t count = std::min(chunk, jobs.size() - first);
        for (int lvl = 1; lvl <= hmax; ++lvl) {
            dim3 grid((max_cols / 2 + 255) / 256, max_rows, (unsigned)count);
            k_pyramid_level<<<grid, 256, 0, h->stream>>>(
                static_cast<const PyrJob*>(h->d_pyrjobs.p) + first, lvl);
            CSM_LAUNCH_CHECK();
        }
    }
    return CSM_OK;
}

/* Bound levels 1..L (csm_bounds.cuh) of every slot that lacks them: what a search with hmax = L + 1
 * reads above the leaves. One launch for all maps. */
int build_bounds(csm_handle h, const std::vector<MapSlot*>& slots, int L)
{
    if (L < 0 || L > 6)
        return fail(h, CSM_E_UNSUPPORTED, "bound levels: 1 <= hmax <= 7");
    {
        const int wrc = wait_uploads(h, slots);
        if (wrc) return wrc;
    }
    std::vector<BlJob> jobs;
    int max_rows = 0, max_cols = 0;
    const unsigned long long mark = ++h->mark;
    for (MapSlot* m : slots) {
        if (m->bounds_levels >= L || m->mark == mark)
            continue;
        m->mark = mark;
        if (L > 0 && m->bounds_alloc < L) {
            if (m->bounds) CSM_CUDA(cudaFreeAsync(m->bounds, h->stream));
            m->bounds = nullptr;
            m->bounds_alloc = 0;
            const size_t bytes = bl_level_offset(L + 1, m->rows, m->cols);
            CSM_CUDA(cudaMallocAsync((void**)&m->bounds, bytes, h->stream));
            /* the padding is never written again: the builder stores whole tiles of the map region only */
            CSM_CUDA(cudaMemsetAsync(m->bounds, 0, bytes, h->stream));
            m->bounds_alloc = L;
        }
        m->bounds_levels = L;
        if (L == 0)
            continue;
        jobs.push_back(BlJob { m->base, m->bounds, m->rows, m->cols });
        max_rows = std::max(max_rows, m->rows);
        max_cols = std::max(max_cols, m->cols);
    }
    if (jobs.empty())
        return CSM_OK;
    /* batches of maps that fit the streaming builder take it in its bound mode (one pass over the map,
     * level 0 read once, every level written once as u8 tiles) */
    bool stream_ok = L >= 1 && L <= 6 && h->bounds_mode != 1 && (h->bounds_mode == 2 || jobs.size() >= 8);
    bool sq512 = true;
    int min_rows = jobs[0].rows;
    for (const BlJob& j : jobs) {
        stream_ok = stream_ok && j.cols <= 512 && (j.cols % 8) == 0 && (j.rows % (kPsRows * kPs2Group)) == 0;
        sq512 = sq512 && j.rows == 512 && j.cols == 512;
        min_rows = std::min(min_rows, j.rows);
    }
    if (h->bounds_mode == 2 && !stream_ok)
        return fail(h, CSM_E_UNSUPPORTED, "bound levels: the streaming builder cannot take these maps");
    if (stream_ok) {
        std::vector<PyrJob> pj(jobs.size());
        for (size_t i = 0; i < jobs.size(); ++i)
            pj[i] = PyrJob { jobs[i].base, reinterpret_cast<uint16_t*>(jobs[i].out), jobs[i].rows, jobs[i].cols };
        const size_t pjb = pj.size() * sizeof(PyrJob);
        int rc = ensure(h, h->d_pyrjobs, pjb);
        if (rc) return rc;
        char* hp = nullptr;
        if ((rc = acquire_upload(h, pjb, &hp))) return rc;
        std::memcpy(hp, pj.data(), pjb);
        if ((rc = pull_to_device(h, h->d_pyrjobs.p, hp, pjb))) return rc;
        if ((rc = upload_committed(h))) return rc;
        h->jobs_on_device.clear();
        phase_mark(h, "start");
        int segs = (int)std::min<size_t>(4, (size_t)((L <= 5 ? CSM_PS2_MINB : 2) * h->sm_count) / pj.size());
        if (h->pyramid_segs > 0) segs = h->pyramid_segs;
        segs = std::max(1, std::min(segs, min_rows / 128));
        const unsigned int grid = (unsigned)(pj.size() * segs);
        const PyrJob* dj = static_cast<const PyrJob*>(h->d_pyrjobs.p);
        const size_t smem2 = sizeof(unsigned int) * ((size_t)kPsStages * kPsRows * kPsInStride + 2 * kPsRows * 256);
#define CSM_PS2B_LAUNCH(HM)                                                                            \
        if (sq512) k_pyramid_stream2<HM, 512, true><<<grid, kPsThreads, smem2, h->stream>>>(dj, segs);     \
        else k_pyramid_stream2<HM, 0, true><<<grid, kPsThreads, smem2, h->stream>>>(dj, segs);
        switch (L) {
        case 1: CSM_PS2B_LAUNCH(1) break;
        case 2: CSM_PS2B_LAUNCH(2) break;
        case 3: CSM_PS2B_LAUNCH(3) break;
        case 4: CSM_PS2B_LAUNCH(4) break;
        case 5: CSM_PS2B_LAUNCH(5) break;
        default: CSM_PS2B_LAUNCH(6) break;
        }
#undef CSM_PS2B_LAUNCH
        CSM_LAUNCH_CHECK();
        phase_mark(h, "k_pyramid_stream(bounds)");
        return CSM_OK;
    }
    const size_t jb = jobs.size() * sizeof(BlJob);
    int rc = ensure(h, h->d_bljobs, jb);
    if (rc) return rc;
    char* hp = nullptr;
    if ((rc = acquire_upload(h, jb, &hp))) return rc;
    std::memcpy(hp, jobs.data(), jb);
    if ((rc = pull_to_device(h, h->d_bljobs.p, hp, jb))) return rc;
    if ((rc = upload_committed(h))) return rc;
    phase_mark(h, "start");
    const int regions_x = (max_cols + kBlOutC - 1) / kBlOutC, regions_y = (max_rows + kBlOutR - 1) / kBlOutR;
    const BlJob* dj = static_cast<const BlJob*>(h->d_bljobs.p);
    for (size_t first = 0; first < jobs.size(); first += 65535) {
        const dim3 grid((unsigned)(regions_x * regions_y), (unsigned)std::min<size_t>(65535, jobs.size() - first));
        const size_t smem = bl_smem_bytes(L);
        switch (L) {
        case 1: k_bounds_build<1><<<grid, kBlThreads, smem, h->stream>>>(dj + first, regions_x); break;
        case 2: k_bounds_build<2><<<grid, kBlThreads, smem, h->stream>>>(dj + first, regions_x); break;
        case 3: k_bounds_build<3><<<grid, kBlThreads, smem, h->stream>>>(dj + first, regions_x); break;
        case 4: k_bounds_build<4><<<grid, kBlThreads, smem, h->stream>>>(dj + first, regions_x); break;
        case 5: k_bounds_build<5><<<grid, kBlThreads, smem, h->stream>>>(dj + first, regions_x); break;
        default: k_bounds_build<6><<<grid, kBlThreads, smem, h->stream>>>(dj + first, regions_x); break;
        }
        CSM_LAUNCH_CHECK();
    }
    phase_mark(h, "k_bounds_build");
    return CSM_OK;
}

struct QueryPlan
{
    std::vector<DevQuery> dq;
    std::vector<double> thetas;
    std::vector<size_t> theta_off;
    std::vector<unsigned long long> inc_init;
    std::vector<unsigned int> root_off;      /* B&B only */
    std::vector<char> extra;                 /* grid search: offsets and positions */
    const double* scan_angles = nullptr;     /* scan arriving with the call */
    const double* scan_ranges = nullptr;
    long long proj_total = 0;
    int max_tn = 0;
    int max_t = 0;
    int max_roots = 0;
};

size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

/* Decide where everything small lives in d_plan. Sizes must be final; the
 * contents (queries etc.) are filled afterwards with these device addresses. */
int layout_plan(csm_handle h, int nq, size_t n_thetas, size_t n_rootoff, size_t extra_bytes,
                int scan_n, PlanView& V)
{
    V = PlanView();
    size_t off = 0;
    V.off_queries = off; off += align16(sizeof(DevQuery) * (size_t)nq);
    V.off_thetas = off;  off += align16(sizeof(double) * n_thetas);
    V.off_inc = off;     off += align16(sizeof(unsigned long long) * (size_t)nq);
    V.off_rootoff = off; off += align16(sizeof(unsigned int) * n_rootoff);
    V.off_extra = off;   off += align16(extra_bytes);
    V.off_scan = off;    off += 2 * align16(sizeof(double) * (size_t)scan_n);
    V.pulled_bytes = off;
    V.zero_off = off;
    const size_t off_qflags = off; off += align16(sizeof(int) * (size_t)nq);
    const size_t off_tiekey = off; off += align16(sizeof(unsigned long long) * (size_t)nq);
    const size_t off_probe = off;  off += align16(sizeof(unsigned long long) * kDiveStarts * (size_t)nq);
    const size_t off_stats = off;  off += align16(sizeof(int) * 2 * (size_t)nq);
    const size_t off_counts = off; off += align16(sizeof(unsigned int) * kMaxLevels);
    const size_t off_overflow = off; off += 16;
    V.zero_bytes = off - V.zero_off;
    V.trig_off = off; off += sizeof(double2) * (size_t)scan_n;
    V.total = off;
    int rc = ensure(h, h->d_plan, V.total);
    if (rc) return rc;
    char* base = static_cast<char*>(h->d_plan.p);
    V.queries = reinterpret_cast<DevQuery*>(base + V.off_queries);
    V.thetas = reinterpret_cast<double*>(base + V.off_thetas);
    V.inc = reinterpret_cast<unsigned long long*>(base + V.off_inc);
    V.rootoff = reinterpret_cast<unsigned int*>(base + V.off_rootoff);
    V.extra = base + V.off_extra;
    V.qflags = reinterpret_cast<int*>(base + off_qflags);
    V.tiekey = reinterpret_cast<unsigned long long*>(base + off_tiekey);
    V.probe = reinterpret_cast<unsigned long long*>(base + off_probe);
    V.stats = reinterpret_cast<int*>(base + off_stats);
    V.counts = reinterpret_cast<unsigned int*>(base + off_counts);
    V.overflow = reinterpret_cast<int*>(base + off_overflow);
    if (scan_n > 0) {
        V.scan.angles = reinterpret_cast<double*>(base + V.off_scan);
        V.scan.ranges = reinterpret_cast<double*>(base + V.off_scan + align16(sizeof(double) * (size_t)scan_n));
        V.scan.trig = reinterpret_cast<double2*>(base + V.trig_off);
        V.scan.n = scan_n;
    }
    h->plan_view = V;
    return CSM_OK;
}

int fill_common(csm_handle h, DevQuery& Q, const MapSlot& m, const ScanSlot& s,
                double score_thr, double known_thr)
{
    if (!(score_thr >= 0.0) || !(known_thr >= 0.0))
        return fail(h, CSM_E_UNSUPPORTED, "thresholds must be >= 0");
    std::memset(&Q, 0, sizeof(Q));
    Q.lvl[0] = m.base;
    for (int l = 1; l <= m.hmax && l < kMaxLevels; ++l)
        Q.lvl[l] = m.levels + (size_t)(l - 1) * m.rows * m.cols;
    for (int l = 1; l <= m.bounds_levels && l < kMaxLevels; ++l) {
        Q.bl[l] = m.bounds + bl_level_offset(l, m.rows, m.cols);
        Q.bl_tpr[l] = bl_tiles_per_row(l, m.cols);
    }
    Q.coarse = m.coarse;
    Q.rows = m.rows; Q.cols = m.cols;
    Q.res = m.res; Q.offx = m.offx; Q.offy = m.offy;
    Q.inv_res = 1.0 / m.res;
    Q.angles = s.angles; Q.ranges = s.ranges; Q.beam_trig = s.trig; Q.n = s.n;
    if (m.alloc_valid) {
        Q.alloc = m.alloc;
        Q.alloc_log2bs = m.alloc_log2bs;
        Q.alloc_bcols = (m.cols + (1 << m.alloc_log2bs) - 1) >> m.alloc_log2bs;
    }
    /* the integer images of the two thresholds depend on (threshold, n) only: queries of a
     * batch share them */
    if (h->memo_n != s.n || h->memo_score_thr != score_thr || h->memo_known_thr != known_thr) {
        h->memo_kthr = make_key_threshold(score_thr, s.n);
        h->memo_nk_cut = make_known_cut(known_thr, s.n);
        h->memo_n = s.n; h->memo_score_thr = score_thr; h->memo_known_thr = known_thr;
    }
    Q.kthr = h->memo_kthr;
    Q.nk_cut = h->memo_nk_cut;
    return CSM_OK;
}

/* Stage the plan laid out by layout_plan with ONE kernel: pull queries, candidate
 * angles, initial incumbents, root offsets, extras and the per-call scan from
 * pinned memory; zero the per-batch counters; fill the scan's trig table. */
int commit_plan(csm_handle h, QueryPlan& plan, const PlanView& V, bool want_rcs)
{
    const int nq = (int)plan.dq.size();
    int rc;
    if ((rc = ensure(h, h->d_proj, sizeof(proj_t) * (size_t)plan.proj_total))) return rc;
    if (want_rcs && (rc = ensure(h, h->d_rcs, sizeof(double2) * (size_t)plan.proj_total))) return rc;
    if ((rc = ensure(h, h->d_results, refined_offset(nq) + sizeof(csm_refined) * (size_t)nq))) return rc;
    {
        const bool fresh = h->d_bestkey.p == nullptr;
        if ((rc = ensure(h, h->d_bestkey, 8))) return rc;
        if (fresh)      /* accumulating batches never clear the word themselves */
            CSM_CUDA(cudaMemsetAsync(h->d_bestkey.p, 0, 8, h->stream));
    }
    for (int q = 0; q < nq; ++q)
        plan.dq[q].thetas = plan.thetas.empty() ? nullptr : V.thetas + plan.theta_off[q];
    char* hp = nullptr;
    if ((rc = acquire_upload(h, V.pulled_bytes, &hp))) return rc;
    std::memcpy(hp + V.off_queries, plan.dq.data(), sizeof(DevQuery) * nq);
    std::memcpy(hp + V.off_thetas, plan.thetas.data(), sizeof(double) * plan.thetas.size());
    std::memcpy(hp + V.off_inc, plan.inc_init.data(), sizeof(unsigned long long) * nq);
    if (!plan.root_off.empty())
        std::memcpy(hp + V.off_rootoff, plan.root_off.data(), sizeof(unsigned int) * plan.root_off.size());
    if (!plan.extra.empty())
        std::memcpy(hp + V.off_extra, plan.extra.data(), plan.extra.size());
    SetupArgs A;
    std::memset(&A, 0, sizeof(A));
    if (V.scan.n > 0) {
        const size_t half = align16(sizeof(double) * (size_t)V.scan.n);
        std::memcpy(hp + V.off_scan, plan.scan_angles, sizeof(double) * V.scan.n);
        std::memcpy(hp + V.off_scan + half, plan.scan_ranges, sizeof(double) * V.scan.n);
        A.angles_host = reinterpret_cast<const double*>(hp + V.off_scan);
        A.trig = V.scan.trig;
        A.n_beams = V.scan.n;
    }
    A.dst = static_cast<uint4*>(h->d_plan.p);
    A.src_host = reinterpret_cast<const uint4*>(hp);
    A.n16 = (unsigned int)(V.pulled_bytes / 16);
    A.zero = reinterpret_cast<uint4*>(static_cast<char*>(h->d_plan.p) + V.zero_off);
    A.z16 = (unsigned int)(V.zero_bytes / 16);
    A.best_key = h->accumulate_best_key ? nullptr : static_cast<unsigned long long*>(h->d_bestkey.p);
    if ((rc = launch_setup(h, A))) return rc;
    return upload_committed(h);
}

int launch_project(csm_handle h, const QueryPlan& plan, const PlanView& V, bool want_rcs)
{
    const int nq = (int)plan.dq.size();
    dim3 grid(std::max(1, (plan.max_t + kProjAngles - 1) / kProjAngles), nq);
    k_project<<<grid, 256, 0, h->stream>>>(
        V.queries, static_cast<proj_t*>(h->d_proj.p),
        want_rcs ? static_cast<double2*>(h->d_rcs.p) : nullptr, V.qflags);
    CSM_LAUNCH_CHECK();
    return CSM_OK;
}

int ensure_frontier(csm_handle h, int nq, unsigned int total_roots)
{
    /* candidate lists: 4 children per survivor; sized for a few thousand
     * candidates per query and at least all roots */
    unsigned int cap = (unsigned int)std::min<long long>(
        std::max<long long>((long long)nq * 16384, 1ll << 20), 1ll << 24);
    if (h->bb_capacity > 0) cap = std::min(cap, (unsigned int)h->bb_capacity);     /* test knob: forces overflows */
    cap = std::max(cap, total_roots);
    int rc;
    for (int l = 0; l < 2; ++l)
        if ((rc = ensure(h, h->d_list[l], sizeof(unsigned long long) * cap))) return rc;
    h->frontier_capacity = cap;
    return CSM_OK;
}

/* cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link) */
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    /* the entry point is a property of the process, not of a device; a function-local static with an
     * initialiser is set exactly once, also with handles created from several threads */
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<EncodeTiledFn>(p);
        return static_cast<EncodeTiledFn>(nullptr);
    }();
    return fn;
}

/* TMA descriptor of a level-0 map: 2-D u16 tensor (cols fastest), boxes of kWtPitch x kWtBoxRows
 * cells, out-of-bounds cells filled with zeros */
bool make_map_tensor(const MapSlot& m, CUtensorMap* out)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr || (m.cols % 8) != 0)
        return false;
    const cuuint64_t dims[2] = { (cuuint64_t)m.cols, (cuuint64_t)m.rows };
    const cuuint64_t strides[1] = { (cuuint64_t)m.cols * sizeof(uint16_t) };
    const cuuint32_t box[2] = { (cuuint32_t)kWtPitch, (cuuint32_t)kWtBoxRows };
    const cuuint32_t estr[2] = { 1, 1 };
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, m.base, dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* The same over the map's 32-bit form (MapSlot::wide): boxes of kWwPitch x kWtBoxRows words */
bool make_map_tensor_wide(const MapSlot& m, CUtensorMap* out)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr || (m.cols % 8) != 0 || m.wide == nullptr)
        return false;
    const cuuint64_t dims[2] = { (cuuint64_t)m.cols, (cuuint64_t)m.rows };
    const cuuint64_t strides[1] = { (cuuint64_t)m.cols * sizeof(uint32_t) };
    const cuuint32_t box[2] = { (cuuint32_t)kWwPitch, (cuuint32_t)kWtBoxRows };
    const cuuint32_t estr[2] = { 1, 1 };
    return fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, m.wide, dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

/* value | known << 20 words of level 0, once per map contents (k_widen_map) */
int ensure_wide(csm_handle h, MapSlot& m)
{
    if (m.wide_valid)
        return CSM_OK;
    const size_t cells = (size_t)m.rows * m.cols;
    if (m.wide == nullptr)
        CSM_CUDA(cudaMallocAsync((void**)&m.wide, cells * sizeof(uint32_t), h->stream));
    const size_t n8 = cells / 8;           /* cols % 8 == 0 */
    const unsigned blocks = (unsigned)std::min<size_t>((n8 + 255) / 256, 148 * 8);
    k_widen_map<<<blocks, 256, 0, h->stream>>>(reinterpret_cast<const uint4*>(m.base), reinterpret_cast<uint4*>(m.wide), n8);
    CSM_LAUNCH_CHECK();
    ++h->launches;
    m.wide_valid = true;
    return CSM_OK;
}

/* scan: when non-null, the (single) query uses a scan that arrives with this
 * call (angles / ranges on the host) instead of a scan uploaded before. */
struct InlineScan
{
    const double* angles;
    const double* ranges;
    int n;
};

/* Low-margin word of every map of the batch that lacks one (first touch): one small launch. The
 * branch-and-bound projection flags node windows that reach below row / column 0 (CSM_FLAG_EDGE);
 * k_finalize withdraws the flag when the map knows nothing in its first 2^hmax rows and columns. */
int ensure_margins(csm_handle h, const std::vector<MapSlot*>& slots)
{
    constexpr int kLimit = 1 << (kMaxLevels - 1);
    std::vector<MapSlot*> todo;
    const unsigned long long mark = ++h->mark;
    for (MapSlot* m : slots)
        if (!m->margin_valid && m->mark != mark) {
            m->mark = mark;
            todo.push_back(m);
        }
    if (todo.empty())
        return CSM_OK;
    int rc;
    for (MapSlot* m : todo)
        if (m->margin_slot < 0) m->margin_slot = h->margin_slots++;
    if (h->d_margin.bytes < sizeof(int) * (size_t)h->margin_slots) {
        /* grow, keeping the words of the other maps */
        DevBuf bigger;
        const size_t want = std::max<size_t>(sizeof(int) * (size_t)h->margin_slots * 2, (size_t)1 << 16);
        if ((rc = ensure(h, bigger, want))) return rc;
        if (h->d_margin.p) {
            CSM_CUDA(cudaMemcpyAsync(bigger.p, h->d_margin.p, h->d_margin.bytes, cudaMemcpyDeviceToDevice, h->stream));
            CSM_CUDA(cudaFreeAsync(h->d_margin.p, h->stream));
        }
        h->d_margin = bigger;
    }
    if ((rc = wait_uploads(h, todo))) return rc;
    std::vector<MarginJob> jobs(todo.size());
    for (size_t i = 0; i < todo.size(); ++i)
        jobs[i] = MarginJob { todo[i]->base, todo[i]->rows, todo[i]->cols, kLimit, todo[i]->margin_slot };
    const size_t jb = sizeof(MarginJob) * jobs.size();
    if ((rc = ensure(h, h->d_marginjobs, jb))) return rc;
    char* hp = nullptr;
    if ((rc = acquire_upload(h, jb, &hp))) return rc;
    std::memcpy(hp, jobs.data(), jb);
    if ((rc = pull_to_device(h, h->d_marginjobs.p, hp, jb))) return rc;
    if ((rc = upload_committed(h))) return rc;
    k_low_margin_init<<<(unsigned)((jobs.size() + 255) / 256), 256, 0, h->stream>>>(
        static_cast<const MarginJob*>(h->d_marginjobs.p), static_cast<int*>(h->d_margin.p), (int)jobs.size());
    CSM_LAUNCH_CHECK();
    k_low_margin<<<dim3((unsigned)jobs.size(), kMarginParts), 256, 0, h->stream>>>(
        static_cast<const MarginJob*>(h->d_marginjobs.p), static_cast<int*>(h->d_margin.p));
    CSM_LAUNCH_CHECK();
    for (MapSlot* m : todo)
        m->margin_valid = true;
    return CSM_OK;
}

int bb_enqueue(csm_handle h, const csm_loop_query* queries, int nq, int hmax, int query_base,
               const InlineScan* inline_scan)
{
    if (nq <= 0 || queries == nullptr)
        return fail(h, CSM_E_INVALID, "loop batch: nq must be positive");
    if (nq > 65535)
        return fail(h, CSM_E_UNSUPPORTED, "loop batch: at most 65535 queries per call");
    if (query_base < 0 || (long long)query_base + nq > 0x100000ll)
        return fail(h, CSM_E_UNSUPPORTED, "loop batch: query_index_base + nq must stay within 2^20 (packed best word)");
    if (h->res_count >= csm_context::kResultSlots)
        return fail(h, CSM_E_CAPACITY, "too many loop batches in flight: call csm_loop_batch_finish");
    if (hmax < 0 || hmax >= kMaxLevels)
        return fail(h, CSM_E_UNSUPPORTED, "branch-and-bound: 0 <= hmax <= 7");
    if (inline_scan && (inline_scan->n <= 0 || inline_scan->n > kMaxBeams || !inline_scan->angles ||
                        !inline_scan->ranges))
        return fail(h, CSM_E_INVALID, "scan: need 1 <= n <= 4096 beams");
    /* pass 1: validate, count candidate angles */
    std::vector<MapSlot*> used_slots(nq);
    size_t n_thetas = 0;
    for (int q = 0; q < nq; ++q) {
        const csm_loop_query& in = queries[q];
        auto mi = h->maps.find(in.map_id);
        if (mi == h->maps.end())
            return fail(h, CSM_E_NOT_FOUND, "loop batch: unknown map id " + std::to_string(in.map_id));
        used_slots[q] = &mi->second;
        if (!inline_scan && h->scans.find(in.scan_id) == h->scans.end())
            return fail(h, CSM_E_NOT_FOUND, "loop batch: unknown scan id " + std::to_string(in.scan_id));
        if (in.win_x < 0 || in.win_y < 0 || in.win_t < 0)
            return fail(h, CSM_E_INVALID, "loop batch: negative window");
        n_thetas += (size_t)(2 * in.win_t + 1);
    }
    int rc;
    /* Which internal nodes are expanded never changes the result, and on
     * loop-detection windows the hmax-level bound is so loose that nearly every root passes: by
     * default the roots are not scored at all but expanded unconditionally, which trades R + 4 p R
     * scored nodes (p ~ 1) for 4 R and a launch on a latency-bound chain. The dive needs the
     * root keys and keeps the scored roots. */
    const int top = hmax;
    const bool dive = top >= 1 && (h->bb_dive == 1 || (h->bb_dive == 2 && nq <= 4));
    const bool unscored_roots = top >= 1 && h->bb_skip_top && !dive;
    /* the sweep over bound levels (csm_bounds.cuh) whenever no kernel needs a u16 level above 0 */
    bool use_bounds = h->bb_bounds != 0 && unscored_roots;
    for (int q = 0; q < nq && use_bounds; ++q)
        use_bounds = 2 * queries[q].win_t + 1 <= 2048;       /* 8-bit angle group index of the group sweep */
    /* what the search reads above level 0, built here for the maps that lack it (first touch) */
    if (use_bounds) {
        if ((rc = build_bounds(h, used_slots, hmax - 1))) return rc;
    } else {
        std::vector<MapSlot*> missing;
        for (MapSlot* m : used_slots)
            if (m->hmax < hmax && std::find(missing.begin(), missing.end(), m) == missing.end())
                missing.push_back(m);
        if (!missing.empty() && (rc = build_levels(h, missing, hmax))) return rc;
    }
    if ((rc = ensure_margins(h, used_slots))) return rc;
    const bool epilogue = inline_scan != nullptr && nq == 1 && h->epilogue_scale > 0.0;
    const bool refine = h->refine_on || epilogue;
    if (refine) {
        if ((rc = wait_uploads(h, used_slots))) return rc;
        if ((rc = ensure_alloc(h, used_slots))) return rc;
    }
    PlanView V;
    if ((rc = layout_plan(h, nq, 0, (size_t)nq + 1, 0, inline_scan ? inline_scan->n : 0, V))) return rc;
    ScanSlot inl = V.scan;
    if (inline_scan)
        inl.max_range = *std::max_element(inline_scan->ranges, inline_scan->ranges + inline_scan->n);

    QueryPlan plan;
    plan.dq.resize(nq);
    plan.theta_off.resize(nq);
    plan.inc_init.resize(nq);
    plan.root_off.assign(nq + 1, 0u);
    if (inline_scan) { plan.scan_angles = inline_scan->angles; plan.scan_ranges = inline_scan->ranges; }
    const int wsz = 1 << hmax;
    for (int q = 0; q < nq; ++q) {
        const csm_loop_query& in = queries[q];
        const MapSlot& m = *used_slots[q];
        const ScanSlot& s = inline_scan ? inl : h->scans.find(in.scan_id)->second;
        DevQuery& Q = plan.dq[q];
        if ((rc = fill_common(h, Q, m, s, in.score_thr, in.known_thr))) return rc;
        Q.sx = in.sensor_pose[0];
        Q.sy = in.sensor_pose[1];
        Q.low_margin = static_cast<const int*>(h->d_margin.p) + m.margin_slot;
        Q.edge_need = 1 << hmax;
        Q.stepx = in.step_x; Q.stepy = in.step_y;
        Q.T = 2 * in.win_t + 1;
        Q.winx = in.win_x; Q.winy = in.win_y;
        /* leaf lattice of the reference: roots every 2^hmax cells from -win (:179-182) */
        Q.lx = ((2 * in.win_x) / wsz + 1) * wsz;
        Q.ly = ((2 * in.win_y) / wsz + 1) * wsz;
        Q.nrx = Q.lx >> top;
        Q.nry = Q.ly >> top;
        if (Q.T > 65535 || Q.lx > 8192 || Q.ly > 8192 || in.win_x > 8192 || in.win_y > 8192 ||
            (unsigned long long)Q.T * Q.lx * Q.ly >= kOrdMask - 1ull)
            return fail(h, CSM_E_UNSUPPORTED, "branch-and-bound: search lattice exceeds 2^26 leaves");
        Q.proj_off = plan.proj_total;
        Q.pst_t = 1; Q.pst_i = Q.T;            /* beam-major for the lane-per-node B&B */
        Q.pquad = use_bounds ? 1 : 0;          /* groups of four beams for the sweep over bound levels */
        Q.tp = (Q.T + 7) & ~7;
        plan.proj_total += (long long)Q.tp * ((Q.n + 15) & ~15);
        plan.max_tn = std::max(plan.max_tn, Q.T * Q.n);
        plan.max_t = std::max(plan.max_t, Q.T);
        plan.max_roots = std::max(plan.max_roots, Q.T * Q.nrx * Q.nry);
        const double extent = (double)(std::max(in.win_x, in.win_y) + wsz) * m.res;
        Q.margin = h->fp_margin_scale * fp_margin(in.sensor_pose, m, s.max_range, extent);
        plan.theta_off[q] = 0;
        Q.theta0 = in.sensor_pose[2]; Q.step_t = in.step_t; Q.tcenter = in.win_t;
        plan.inc_init[q] = ((unsigned long long)Q.kthr.fail_max << kOrdBits) | kOrdMask;
        /* list entries of the roots: one per node, or one per group of 8 angles (group sweep) */
        plan.root_off[q + 1] = plan.root_off[q] + (unsigned int)((use_bounds ? (Q.T + 7) / 8 : Q.T) * Q.nrx * Q.nry);
    }
    if ((rc = wait_uploads(h, used_slots))) return rc;
    if ((rc = ensure_frontier(h, nq, plan.root_off[nq]))) return rc;
    phase_mark(h, "start");
    if ((rc = commit_plan(h, plan, V, false))) return rc;
    phase_mark(h, "k_setup");
    if ((rc = launch_project(h, plan, V, false))) return rc;
    phase_mark(h, "k_project");

    BbWork W;
    std::memset(&W, 0, sizeof(W));
    W.list[0] = static_cast<unsigned long long*>(h->d_list[0].p);
    W.list[1] = static_cast<unsigned long long*>(h->d_list[1].p);
    W.counts = V.counts;
    W.incumbent = V.inc;
    W.tiekey = V.tiekey;
    W.stats = V.stats;
    W.overflow = V.overflow;
    W.capacity = h->frontier_capacity;
    W.top = top;
    W.split_shift = h->bb_split_shift;
    if (use_bounds && h->bb_probe && top >= 3) {
        /* a start node's lattice position takes 14 + 14 bits and its angle 12; option values above 2 are a mask
         * of heights (bit hc), for experiments */
        bool fits = true;
        for (int q = 0; q < nq && fits; ++q)
            fits = plan.dq[q].lx < (1 << kProbePosBits) && plan.dq[q].ly < (1 << kProbePosBits) && plan.dq[q].T <= 4096;
        unsigned int heights = h->bb_probe == 1 ? (1u << 4) : h->bb_probe == 2 ? ((1u << 5) | (1u << 4)) : (unsigned int)h->bb_probe;
        /* the first launch creates the children of height top - 1: a shallower search dives once, from there */
        if ((heights & ((1u << top) - 1u) & ~3u) == 0u) heights = 1u << (top - 1);
        heights &= ((1u << top) - 1u) & ~3u & 0x3fu;
        if (fits && heights != 0u) { W.probe = V.probe; W.probe_heights = heights; }
    }
    if (dive) {
        if ((rc = ensure(h, h->d_rootkey, sizeof(long long) * (size_t)plan.root_off[nq]))) return rc;
        W.rootkey = static_cast<long long*>(h->d_rootkey.p);
    }
    const DevQuery* dq = V.queries;
    const proj_t* proj = static_cast<const proj_t*>(h->d_proj.p);

    {
        dim3 grid(std::max(1, std::min((plan.max_roots + 255) / 256, 64)), nq);
        if (use_bounds) k_bbg_init<<<grid, 256, 0, h->stream>>>(dq, V.rootoff, nq, W);
        else k_bb_init<<<grid, 256, 0, h->stream>>>(dq, V.rootoff, nq, W, unscored_roots ? 1 : 0);
        CSM_LAUNCH_CHECK();
    }
    {
        const unsigned int n_roots = plan.root_off[nq];
        /* one resident wave: the kernels walk their list with a grid-stride loop and choose the lanes
         * per node from the grid size, so CTAs beyond what fits at once would only queue behind the
         * first wave (a second round of load latency on short lists) */
        if (h->bb_ctas_per_sm == 0) {
            int a = 0, b = 0, c = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_bb_expand<0>, 256, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_bb_expand<3>, 256, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, k_bb_roots, 256, 0);
            int d = 0, e = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d, k_bbg_expand<0>, 32 * kBbgWarps, 0);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&e, k_bbg_expand<3>, 32 * kBbgWarps, 0);
            h->bb_ctas_per_sm = std::max(1, std::min(std::min(a > 0 ? a : 8, b > 0 ? b : 8), c > 0 ? c : 8));
            h->bbx_ctas_per_sm = std::max(1, std::min(d > 0 ? d : 8, e > 0 ? e : 8));
        }
        int per_sm = use_bounds ? h->bbx_ctas_per_sm : h->bb_ctas_per_sm;
        if (h->bb_sweep_ctas_cap > 0) per_sm = std::min(per_sm, h->bb_sweep_ctas_cap);
        const int full = h->sm_count * per_sm;
        if (!unscored_roots) {
            const int root_blocks = (int)std::min<unsigned int>((n_roots + 7) / 8, (unsigned int)full);
            k_bb_roots<<<std::max(root_blocks, 1), 256, 0, h->stream>>>(dq, proj, W, n_roots);
            CSM_LAUNCH_CHECK();
        }
        phase_mark(h, "k_bb_init+k_bb_roots");
        if (dive) {
            k_bb_dive<<<nq, 256, 0, h->stream>>>(dq, proj, V.rootoff, W);
            CSM_LAUNCH_CHECK();
            phase_mark(h, "k_bb_dive");
        }
        /* a list never holds more than 4^k times the roots: small calls get small grids */
        unsigned long long bound = n_roots;
        for (int lvl = top; lvl >= 1 && lvl > h->bb_stop_level; --lvl) {
            /* per-node lists: 8 nodes per warp; group lists: a CTA per group when the list is short */
            const int blocks = (int)std::min<unsigned long long>(use_bounds ? bound : (bound + 7) / 8, (unsigned long long)full);
            const dim3 g((unsigned)std::max(blocks, 1));
            if (use_bounds) {
                const int gt = 32 * kBbgWarps;
                const int nchunks_hint = (plan.dq[0].n + 15) / 16;
                switch (lvl - 1) {
                case 0: k_bbg_expand<0><<<g, gt, 0, h->stream>>>(dq, proj, W, nchunks_hint); break;
                case 1: k_bbg_expand<1><<<g, gt, 0, h->stream>>>(dq, proj, W, nchunks_hint); break;
                case 2: k_bbg_expand<2><<<g, gt, 0, h->stream>>>(dq, proj, W, nchunks_hint); break;
                case 3: k_bbg_expand<3><<<g, gt, 0, h->stream>>>(dq, proj, W, nchunks_hint); break;
                case 4: k_bbg_expand<4><<<g, gt, 0, h->stream>>>(dq, proj, W, nchunks_hint); break;
                case 5: k_bbg_expand<5><<<g, gt, 0, h->stream>>>(dq, proj, W, nchunks_hint); break;
                default: k_bbg_expand<6><<<g, gt, 0, h->stream>>>(dq, proj, W, nchunks_hint); break;
                }
            } else
            switch (lvl - 1) {          /* height of the children: compile-time for the index arithmetic */
            case 0: k_bb_expand<0><<<g, 256, 0, h->stream>>>(dq, proj, W); break;
            case 1: k_bb_expand<1><<<g, 256, 0, h->stream>>>(dq, proj, W); break;
            case 2: k_bb_expand<2><<<g, 256, 0, h->stream>>>(dq, proj, W); break;
            case 3: k_bb_expand<3><<<g, 256, 0, h->stream>>>(dq, proj, W); break;
            case 4: k_bb_expand<4><<<g, 256, 0, h->stream>>>(dq, proj, W); break;
            case 5: k_bb_expand<5><<<g, 256, 0, h->stream>>>(dq, proj, W); break;
            default: k_bb_expand<6><<<g, 256, 0, h->stream>>>(dq, proj, W); break;
            }
            CSM_LAUNCH_CHECK();
            if (h->timing) {
                const std::string nm = std::string(use_bounds ? "k_bbg_expand<" : "k_bb_expand<") + std::to_string(lvl - 1) + ">";
                phase_mark(h, nm.c_str());
            }
            if (use_bounds && W.probe != nullptr && ((W.probe_heights >> (lvl - 1)) & 1u) && lvl - 1 > h->bb_stop_level) {
                k_bbg_dive<<<nq, 256, 0, h->stream>>>(dq, proj, W, lvl - 1);
                CSM_LAUNCH_CHECK();
                if (h->timing) phase_mark(h, (std::string("k_bbg_dive<") + std::to_string(lvl - 1) + ">").c_str());
            }
            bound = std::min<unsigned long long>(bound * 4, (unsigned long long)h->frontier_capacity);
        }
    }
    FinalArgs F;
    std::memset(&F, 0, sizeof(F));
    F.decode = 1;
    F.incumbent = V.inc;
    F.tiekey = V.tiekey;
    F.stats = V.stats;
    F.best_key = static_cast<unsigned long long*>(h->d_bestkey.p);
    F.query_index_base = query_base;
    F.mode = 0;
    F.qflags = V.qflags;
    F.overflow = V.overflow;
    F.nq = nq;
    k_finalize<<<nq, 32, 0, h->stream>>>(dq, proj, F, static_cast<csm_result*>(h->d_results.p));
    CSM_LAUNCH_CHECK();
    phase_mark(h, "k_finalize");
    if (refine) {
        /* ScanMatcherLinearSolver on every pose found, loop_detector_branch_bound.cpp:110-135 */
        RefineArgs R;
        std::memset(&R, 0, sizeof(R));
        R.results = static_cast<const csm_result*>(h->d_results.p);
        R.out = reinterpret_cast<csm_refined*>(static_cast<char*>(h->d_results.p) + refined_offset(nq));
        if (epilogue) {
            /* single-scan match: Cost and ComputeCovariance at the decided pose, found or not
             * (scan_matcher_branch_bound.cpp:241-252) */
            R.max_iterations = 0;
            R.always = 1;
            R.covariance_scale = h->epilogue_scale;
        } else {
            R.max_iterations = h->refine.max_iterations;
            R.convergence_threshold = h->refine.convergence_threshold;
            R.lambda0 = h->refine.lambda;
            R.covariance_scale = h->refine.covariance_scale;
        }
        k_refine<<<nq, kRefThreads, 0, h->stream>>>(dq, R);
        CSM_LAUNCH_CHECK();
        phase_mark(h, "k_refine");
    }
    rc = enqueue_readback(h, nq, refine);
    phase_mark(h, "readback");
    if (rc == CSM_OK) {
        BatchRecord& rec = h->res_rec[(h->res_head + h->res_count - 1) % csm_context::kResultSlots];
        rec.queries.assign(queries, queries + nq);
        rec.hmax = hmax;
        rec.inline_scan = inline_scan != nullptr;
        if (inline_scan) {
            rec.angles.assign(inline_scan->angles, inline_scan->angles + inline_scan->n);
            rec.ranges.assign(inline_scan->ranges, inline_scan->ranges + inline_scan->n);
        }
    }
    return rc;
}

/* Exhaustive search of a candidate lattice (positions px x py, angles thetas) with the reference's
 * per-candidate arithmetic from host-evaluated hit terms; out->best_* are lattice indices (ix, iy, it). */
int exact_lattice_search(csm_handle h, MapSlot& m, const double* angles, const double* ranges, int n,
                         const std::vector<double>& thetas, const std::vector<double>& px,
                         const std::vector<double>& py, int ord_mode, double score_thr, double known_thr,
                         csm_result* out)
{
    const int ndx = (int)px.size(), ndy = (int)py.size(), ndt = (int)thetas.size();
    if (ndt > 65535 || (unsigned long long)ndx * ndy * ndt >= kOrdMask - 1ull)
        return fail(h, CSM_E_UNSUPPORTED, "exact rerun: lattice exceeds 2^26 candidates");
    int rc;
    if ((rc = wait_uploads(h, std::vector<MapSlot*>{ &m }))) return rc;
    if ((rc = ensure_exact_area(h))) return rc;
    /* the reference's hit terms, on the host */
    std::vector<double2> rcs((size_t)ndt * n);
    for (int t = 0; t < ndt; ++t)
        for (int i = 0; i < n; ++i)
            rcs[(size_t)t * n + i] = host_hit_terms(thetas[t], angles[i], ranges[i]);
    PlanView V;
    const size_t ob = align16(sizeof(int) * (size_t)(ndx + ndy));
    const size_t pb = sizeof(double) * (size_t)(ndx + ndy);
    if ((rc = layout_plan(h, 1, 0, 0, ob + pb, 0, V))) return rc;
    ScanSlot dummy;
    dummy.n = n;
    QueryPlan plan;
    plan.dq.resize(1);
    plan.theta_off.assign(1, 0);
    plan.inc_init.assign(1, 0ull);
    DevQuery& Q = plan.dq[0];
    if ((rc = fill_common(h, Q, m, dummy, score_thr, known_thr))) return rc;
    Q.T = ndt;
    Q.proj_off = 0;
    Q.pst_t = n; Q.pst_i = 1;
    Q.margin = 0.0;              /* nothing to guard: these are the reference's own operations */
    plan.proj_total = (long long)ndt * n;
    plan.extra.assign(ob + pb, 0);
    {
        double* pos = reinterpret_cast<double*>(plan.extra.data() + ob);
        for (int k = 0; k < ndx; ++k) pos[k] = px[k];
        for (int k = 0; k < ndy; ++k) pos[ndx + k] = py[k];
    }
    if ((rc = commit_plan(h, plan, V, true))) return rc;
    CSM_CUDA(cudaMemcpyAsync(h->d_rcs.p, rcs.data(), sizeof(double2) * rcs.size(), cudaMemcpyHostToDevice, h->stream));
    GridArgs G;
    G.mx = reinterpret_cast<const int*>(V.extra);
    G.my = G.mx + ndx;
    G.px = reinterpret_cast<const double*>(V.extra + ob);
    G.py = G.px + ndx;
    G.ndx = ndx; G.ndy = ndy; G.ndt = ndt;
    G.best = V.inc;
    G.tiekey = V.tiekey;
    G.ord_mode = ord_mode;
    const DevQuery* dq = V.queries;
    const proj_t* proj = static_cast<const proj_t*>(h->d_proj.p);
    dim3 grid(ndt, (ndy + 7) / 8);
    k_grid_general<<<grid, 256, sizeof(double2) * n, h->stream>>>(dq, static_cast<const double2*>(h->d_rcs.p), G, V.qflags);
    CSM_LAUNCH_CHECK();
    FinalArgs F;
    std::memset(&F, 0, sizeof(F));
    F.decode = 2;
    F.G = G;
    F.mx = G.mx; F.my = G.my; F.px = G.px; F.py = G.py;
    F.rcs = static_cast<const double2*>(h->d_rcs.p);
    F.mode = 2;
    F.qflags = V.qflags;
    F.nq = 1;
    k_finalize<<<1, 32, 0, h->stream>>>(dq, proj, F, static_cast<csm_result*>(h->d_results.p));
    CSM_LAUNCH_CHECK();
    CSM_CUDA(cudaMemcpyAsync(h->h_exact, h->d_results.p, sizeof(csm_result), cudaMemcpyDeviceToHost, h->stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));
    *out = *static_cast<const csm_result*>(h->h_exact);
    ++h->exact_reruns;
    return CSM_OK;
}

/* ScanMatcherBranchBound::OptimizePose for one query, exactly: with admissible bounds the best-first
 * search returns the best leaf among the leaves that pass both thresholds, so
 * the rerun scores the whole leaf lattice with the node poses of scan_matcher_branch_bound.cpp:159-162. */
int exact_rerun_bb(csm_handle h, const csm_loop_query& q, const double* angles, const double* ranges, int n,
                   int hmax, csm_result* r)
{
    auto mi = h->maps.find(q.map_id);
    if (mi == h->maps.end())
        return fail(h, CSM_E_NOT_FOUND, "exact rerun: unknown map id");
    const int wsz = 1 << hmax;
    const int lx = ((2 * q.win_x) / wsz + 1) * wsz, ly = ((2 * q.win_y) / wsz + 1) * wsz, T = 2 * q.win_t + 1;
    std::vector<double> thetas(T), px(lx), py(ly);
    for (int t = 0; t < T; ++t) thetas[t] = q.sensor_pose[2] + (t - q.win_t) * q.step_t;
    for (int x = 0; x < lx; ++x) px[x] = q.sensor_pose[0] + (x - q.win_x) * q.step_x;
    for (int y = 0; y < ly; ++y) py[y] = q.sensor_pose[1] + (y - q.win_y) * q.step_y;
    csm_result e;
    const int rc = exact_lattice_search(h, mi->second, angles, ranges, n, thetas, px, py, 1, q.score_thr, q.known_thr, &e);
    if (rc) return rc;
    r->flags = ((r->flags & ~CSM_FLAG_FP_MARGIN) & ~CSM_FLAG_KEY_TIE) | (e.flags & CSM_FLAG_KEY_TIE) | CSM_FLAG_EXACT;
    r->found = e.found;
    if (e.found) {
        r->best_x = e.best_x - q.win_x; r->best_y = e.best_y - q.win_y; r->best_t = e.best_t - q.win_t;
        r->sum_value = e.sum_value; r->n_known = e.n_known; r->normalized_score = e.normalized_score;
    } else {
        r->best_x = r->best_y = r->best_t = 0;      /* scan_matcher_branch_bound.cpp:145-147 */
    }
    return CSM_OK;
}

/* Refine n given poses (synchronous); max_iterations == 0 evaluates cost and covariance only */
int refine_poses(csm_handle h, const csm_refine_query* queries, int n, const csm_refine_params* p, csm_refined* out)
{
    int rc;
    std::vector<MapSlot*> slots(n);
    for (int q = 0; q < n; ++q) {
        auto mi = h->maps.find(queries[q].map_id);
        if (mi == h->maps.end())
            return fail(h, CSM_E_NOT_FOUND, "refine batch: unknown map id " + std::to_string(queries[q].map_id));
        if (h->scans.find(queries[q].scan_id) == h->scans.end())
            return fail(h, CSM_E_NOT_FOUND, "refine batch: unknown scan id " + std::to_string(queries[q].scan_id));
        slots[q] = &mi->second;
    }
    if ((rc = wait_uploads(h, slots))) return rc;
    if ((rc = ensure_alloc(h, slots))) return rc;
    const size_t q_bytes = align16(sizeof(DevQuery) * (size_t)n);
    const size_t in_bytes = q_bytes + align16(sizeof(double) * 3 * (size_t)n);
    if ((rc = ensure(h, h->d_refine_in, in_bytes + sizeof(csm_refined) * (size_t)n))) return rc;
    char* hp = nullptr;
    if ((rc = acquire_upload(h, in_bytes, &hp))) return rc;
    for (int q = 0; q < n; ++q) {
        DevQuery Q;
        if ((rc = fill_common(h, Q, *slots[q], h->scans.find(queries[q].scan_id)->second, 0.0, 0.0))) return rc;
        std::memcpy(hp + sizeof(DevQuery) * (size_t)q, &Q, sizeof(Q));
        std::memcpy(hp + q_bytes + sizeof(double) * 3 * (size_t)q, queries[q].sensor_pose, sizeof(double) * 3);
    }
    if ((rc = pull_to_device(h, h->d_refine_in.p, hp, in_bytes))) return rc;
    if ((rc = upload_committed(h))) return rc;
    char* base = static_cast<char*>(h->d_refine_in.p);
    RefineArgs R;
    std::memset(&R, 0, sizeof(R));
    R.start = reinterpret_cast<const double*>(base + q_bytes);
    R.out = reinterpret_cast<csm_refined*>(base + in_bytes);
    R.max_iterations = p->max_iterations;
    R.always = p->max_iterations == 0 ? 1 : 0;
    R.convergence_threshold = p->convergence_threshold;
    R.lambda0 = p->lambda;
    R.covariance_scale = p->covariance_scale;
    k_refine<<<n, kRefThreads, 0, h->stream>>>(reinterpret_cast<const DevQuery*>(base), R);
    CSM_LAUNCH_CHECK();
    CSM_CUDA(cudaMemcpyAsync(out, R.out, sizeof(csm_refined) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));
    return CSM_OK;
}

/* Results of a batch that came back flagged CSM_FLAG_FP_MARGIN: exact rerun (and, when the batch
 * refined its poses and the rerun moved the pose, the refinement again from the new pose). */
int postprocess_flags(csm_handle h, const BatchRecord& rec, csm_result* results, int nq, csm_refined* refined)
{
    int rc;
    if (!h->exact_rerun)
        return CSM_OK;
    for (int q = 0; q < nq; ++q) {
        if (!(results[q].flags & CSM_FLAG_FP_MARGIN)) continue;
        const csm_loop_query& lq = rec.queries[q];
        const double* angles = nullptr; const double* ranges = nullptr; int n = 0;
        if (rec.inline_scan) {
            angles = rec.angles.data(); ranges = rec.ranges.data(); n = (int)rec.angles.size();
        } else {
            auto si = h->scans.find(lq.scan_id);
            if (si == h->scans.end()) continue;
            angles = si->second.h_angles.data(); ranges = si->second.h_ranges.data(); n = si->second.n;
        }
        const csm_result before = results[q];
        if ((rc = exact_rerun_bb(h, lq, angles, ranges, n, rec.hmax, &results[q]))) return rc;
        const csm_result& after = results[q];
        const bool moved = before.found != after.found || before.best_x != after.best_x ||
                           before.best_y != after.best_y || before.best_t != after.best_t;
        if (refined != nullptr && moved) {
            /* the refinement started from the other pose: run it again from the exact one */
            std::memset(&refined[q], 0, sizeof(csm_refined));
            const bool epilogue = rec.inline_scan && h->epilogue_scale > 0.0;
            if (after.found || epilogue) {
                csm_refine_query rq;
                rq.map_id = lq.map_id;
                rq.scan_id = lq.scan_id;
                rq.sensor_pose[0] = lq.sensor_pose[0] + lq.step_x * after.best_x;
                rq.sensor_pose[1] = lq.sensor_pose[1] + lq.step_y * after.best_y;
                rq.sensor_pose[2] = lq.sensor_pose[2] + lq.step_t * after.best_t;
                if (rec.inline_scan) {
                    if ((rc = upload_scan_impl(h, kTempScanId, angles, ranges, n))) return rc;
                    rq.scan_id = kTempScanId;
                }
                csm_refine_params p = h->refine;
                if (epilogue) { p.max_iterations = 0; p.covariance_scale = h->epilogue_scale; }
                if ((rc = refine_poses(h, &rq, 1, &p, &refined[q]))) return rc;
            }
        }
    }
    return CSM_OK;
}

/* Kernels that need more dynamic shared memory than the default limit: the attribute is per device,
 * so every handle sets it for its own device when it is created (no process-wide flags). */
cudaError_t set_kernel_attributes()
{
    cudaError_t e;
    const int ps = (int)(sizeof(unsigned int) * ((size_t)kPsStages * kPsRows * kPsInStride + 2 * kPsRows * 256 + 63 * 256));
    if ((e = cudaFuncSetAttribute(k_pyramid_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, ps)) != cudaSuccess) return e;
#define CSM_BL_ATTR(L) \
    if ((e = cudaFuncSetAttribute(k_bounds_build<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bl_smem_bytes(L))) != cudaSuccess) return e;
    CSM_BL_ATTR(1) CSM_BL_ATTR(2) CSM_BL_ATTR(3) CSM_BL_ATTR(4) CSM_BL_ATTR(5) CSM_BL_ATTR(6)
#undef CSM_BL_ATTR
    const int wt = (int)wt_smem_bytes(kMaxBeams), ww = (int)wt_smem_bytes_wide(kMaxBeams);
#define CSM_WT_ATTR(UNIT, CHUNKS, REM)                                                                                        \
    if ((e = cudaFuncSetAttribute(k_window_tma<UNIT, CHUNKS, REM, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, wt)) != cudaSuccess) return e; \
    if ((e = cudaFuncSetAttribute(k_window_tma<UNIT, CHUNKS, REM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ww)) != cudaSuccess) return e;
    CSM_WT_ATTR(true, 1, true) CSM_WT_ATTR(true, 2, true) CSM_WT_ATTR(true, 3, true)
    CSM_WT_ATTR(true, 4, true) CSM_WT_ATTR(true, 5, true) CSM_WT_ATTR(true, 6, true)
    CSM_WT_ATTR(true, kWtChunks, false) CSM_WT_ATTR(false, kWtChunks, false)
#undef CSM_WT_ATTR
    return cudaSuccess;
}

int check_refine_params(csm_handle h, const csm_refine_params* p)
{
    if (p->max_iterations < 1 || p->max_iterations > 1000 || !(p->convergence_threshold >= 0.0) ||
        !(p->lambda >= 0.0) || !(p->covariance_scale > 0.0))
        return fail(h, CSM_E_INVALID, "refiner: need 1 <= max_iterations <= 1000, threshold >= 0, "
                                      "lambda >= 0, covariance_scale > 0");
    return CSM_OK;
}

} /* namespace */

/* ======================================================================== */
extern "C" {

int csm_version(void) { return 100; }

int csm_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

int csm_create(int device, unsigned flags, csm_handle* out)
{
    (void)flags;
    if (out == nullptr)
        return CSM_E_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n)
        return CSM_E_CUDA;
    if (cudaSetDevice(device) != cudaSuccess)
        return CSM_E_CUDA;
    if (set_kernel_attributes() != cudaSuccess)
        return CSM_E_CUDA;
    csm_handle h = new csm_context;
    h->device = device;
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete h;
        return CSM_E_CUDA;
    }
    if (cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->compute_mark, cudaEventDisableTiming) != cudaSuccess) {
        cudaStreamDestroy(h->stream);
        delete h;
        return CSM_E_CUDA;
    }
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    /* CSM_OPTIONS="name=value,name=value": csm_set_option calls for every handle of the process, for A/B runs
     * through host code that does not expose the knobs (unknown names and bad values are ignored) */
    if (const char* env = std::getenv("CSM_OPTIONS")) {
        std::string all(env);
        size_t at = 0;
        while (at < all.size()) {
            size_t end = all.find(',', at);
            if (end == std::string::npos) end = all.size();
            const std::string kv = all.substr(at, end - at);
            const size_t eq = kv.find('=');
            if (eq != std::string::npos && eq > 0)
                csm_set_option(h, kv.substr(0, eq).c_str(), std::atoi(kv.c_str() + eq + 1));
            at = end + 1;
        }
    }
    *out = h;
    return CSM_OK;
}

int csm_destroy(csm_handle h)
{
    if (h == nullptr)
        return CSM_E_INVALID;
    cudaSetDevice(h->device);
    csm_comm_destroy(h);
    cudaStreamSynchronize(h->copy_stream);
    cudaStreamSynchronize(h->stream);
    for (auto& kv : h->maps) free_map(h, kv.second);
    for (auto& kv : h->scans) free_scan(h, kv.second);
    DevBuf* bufs[] = { &h->d_plan, &h->d_proj, &h->d_rcs, &h->d_results, &h->d_bestkey,
                       &h->d_rtblocks, &h->d_pyrjobs, &h->d_rootkey, &h->d_wtgroups, &h->d_refine_in, &h->d_allocjobs, &h->d_bljobs, &h->d_margin, &h->d_marginjobs,
                       &h->d_maptables, &h->d_mapwork, &h->d_maperror };
    for (DevBuf* b : bufs)
        if (b->p) cudaFreeAsync(b->p, h->stream);
    for (int l = 0; l < 2; ++l)
        if (h->d_list[l].p) cudaFreeAsync(h->d_list[l].p, h->stream);
    cudaStreamSynchronize(h->stream);
    if (h->h_exact) cudaFreeHost(h->h_exact);
    if (h->h_maperror) cudaFreeHost(h->h_maperror);
    for (int k = 0; k < csm_context::kUploadAreas; ++k) {
        if (h->h_up[k]) cudaFreeHost(h->h_up[k]);
        if (h->h_up_done[k]) cudaEventDestroy(h->h_up_done[k]);
    }
    for (int k = 0; k < csm_context::kResultSlots; ++k) {
        if (h->h_res[k]) cudaFreeHost(h->h_res[k]);
        if (h->h_res_done[k]) cudaEventDestroy(h->h_res_done[k]);
    }
    for (int k = 0; k < 16; ++k)
        if (h->upload_events[k]) cudaEventDestroy(h->upload_events[k]);
    for (cudaEvent_t e : h->tev) cudaEventDestroy(e);
    cudaEventDestroy(h->compute_mark);
    if (h->owns_copy_stream) cudaStreamDestroy(h->copy_stream);
    cudaStreamDestroy(h->stream);
    delete h;
    return CSM_OK;
}

const char* csm_last_error(csm_handle h)
{
    return h ? h->err.c_str() : "null handle";
}

void* csm_stream(csm_handle h) { return h ? (void*)h->stream : nullptr; }

int csm_synchronize(csm_handle h)
{
    if (!h) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    CSM_CUDA(cudaStreamSynchronize(h->copy_stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));
    if (h->h_maperror != nullptr && *h->h_maperror != 0) {
        *h->h_maperror = 0;
        return fail(h, CSM_E_INVALID, "map construction: a beam left the map (resize the map to the scan's bounding box first)");
    }
    return CSM_OK;
}

int64_t csm_launch_count(csm_handle h) { return h ? h->launches : 0; }

int64_t csm_exact_rerun_count(csm_handle h) { return h ? h->exact_reruns : 0; }

int csm_set_option(csm_handle h, const char* name, int value)
{
    if (!h || !name) return CSM_E_INVALID;
    if (std::strcmp(name, "pyramid_mode") == 0 && value >= 0 && value <= 3) {
        h->pyramid_mode = value;
        return CSM_OK;
    }
    if (std::strcmp(name, "bb_dive") == 0 && value >= 0 && value <= 2) { h->bb_dive = value; return CSM_OK; }
    if (std::strcmp(name, "accumulate_best_key") == 0) { h->accumulate_best_key = value; return CSM_OK; }
    if (std::strcmp(name, "bb_split_shift") == 0) { h->bb_split_shift = std::max(-4, std::min(value, 4)); return CSM_OK; }
    if (std::strcmp(name, "exact_rerun") == 0) { h->exact_rerun = value != 0; return CSM_OK; }
    if (std::strcmp(name, "saturated_unknown") == 0) { h->saturated_unknown = value != 0; return CSM_OK; }
    if (std::strcmp(name, "fp_margin_scale") == 0) { h->fp_margin_scale = value > 0 ? (double)value : 1.0; return CSM_OK; }
    if (std::strcmp(name, "bb_capacity") == 0) { h->bb_capacity = std::max(0, value); return CSM_OK; }
    if (std::strcmp(name, "bb_sweep_ctas_per_sm") == 0) { h->bb_sweep_ctas_cap = std::max(0, value); return CSM_OK; }
    if (std::strcmp(name, "pyramid_segs") == 0) { h->pyramid_segs = std::max(0, std::min(value, 4)); return CSM_OK; }
    if (std::strcmp(name, "bounds_mode") == 0 && value >= 0 && value <= 2) { h->bounds_mode = value; return CSM_OK; }
    if (std::strcmp(name, "bb_stop_level") == 0) { h->bb_stop_level = std::max(0, std::min(value, kMaxLevels - 1)); return CSM_OK; }
    if (std::strcmp(name, "bb_bounds") == 0) { h->bb_bounds = value != 0; return CSM_OK; }
    if (std::strcmp(name, "bb_skip_top") == 0) { h->bb_skip_top = value != 0; return CSM_OK; }
    if (std::strcmp(name, "bb_probe") == 0 && value >= 0) { h->bb_probe = value; return CSM_OK; }
    if (std::strcmp(name, "window_mode") == 0 && value >= 0 && value <= 3) { h->window_mode = value; return CSM_OK; }
    if (std::strcmp(name, "timing") == 0) { h->timing = value; h->tcount = 0; return CSM_OK; }
    if (std::strcmp(name, "accumulate_best_key") == 0 && value != 0 && h->d_bestkey.p == nullptr) {
        CSM_CUDA(cudaSetDevice(h->device));
        int rc = ensure(h, h->d_bestkey, 8);
        if (rc) return rc;
        CSM_CUDA(cudaMemsetAsync(h->d_bestkey.p, 0, 8, h->stream));
    }
    if (std::strcmp(name, "reset_best_key") == 0) {
        CSM_CUDA(cudaSetDevice(h->device));
        int rc = ensure(h, h->d_bestkey, 8);
        if (rc) return rc;
        CSM_CUDA(cudaMemsetAsync(h->d_bestkey.p, 0, 8, h->stream));
        return CSM_OK;
    }
    return fail(h, CSM_E_INVALID, std::string("unknown option ") + name);
}

void* csm_alloc_pinned(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess)
        return nullptr;
    return p;
}

void csm_free_pinned(void* p)
{
    if (p) cudaFreeHost(p);
}

static int upload_grid_impl(csm_handle h, int64_t map_id, const uint16_t* dense, bool on_device,
                            int rows, int cols, double res, double offx, double offy)
{
    if (!h) return CSM_E_INVALID;
    if (dense == nullptr || rows <= 0 || cols <= 0 || rows > 16384 || cols > 16384 || (cols & 1) ||
        !(res > 0.0))
        return fail(h, CSM_E_INVALID, "grid: need 0 < rows, cols <= 16384, even cols, resolution > 0");
    CSM_CUDA(cudaSetDevice(h->device));
    MapSlot& m = h->maps[map_id];
    {
        const int src = settle_slot(h, m);
        if (src) return src;
    }
    bool fresh_alloc = false;
    if (m.rows != rows || m.cols != cols || m.base == nullptr) {
        free_map(h, m);
        CSM_CUDA(cudaMallocAsync((void**)&m.base, (size_t)rows * cols * sizeof(uint16_t), h->stream));
        m.rows = rows; m.cols = cols;
        fresh_alloc = true;
    }
    /* precomputed levels belong to the previous contents (allocations are kept) */
    m.hmax = 0;
    m.bounds_levels = -1;
    m.margin_valid = false;
    m.coarse_win = 0;
    m.wide_valid = false;
    m.pending_scatter.reset();
    m.alloc_valid = false;
    m.res = res; m.offx = offx; m.offy = offy;
    const size_t bytes = (size_t)rows * cols * sizeof(uint16_t);
    if (on_device) {
        CSM_CUDA(cudaMemcpyAsync(m.base, dense, bytes, cudaMemcpyDeviceToDevice, h->stream));
        return saturated_pass(h, m.base, bytes, h->stream);
    }
    /* Host uploads go to the copy stream. It first waits for the compute work
     * enqueued so far: kernels of the previous batch may still read the old
     * contents of this slot, and a fresh allocation is ordered on the compute
     * stream. */
    if (!h->upload_open || fresh_alloc) {
        CSM_CUDA(cudaEventRecord(h->compute_mark, h->stream));
        CSM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->compute_mark, 0));
        h->upload_open = true;
    }
    CSM_CUDA(cudaMemcpyAsync(m.base, dense, bytes, cudaMemcpyHostToDevice, h->copy_stream));
    m.pending_upload = reinterpret_cast<cudaEvent_t>(1);
    return saturated_pass(h, m.base, bytes, h->copy_stream);
}

int csm_upload_grid(csm_handle h, int64_t map_id, const uint16_t* dense,
                    int rows, int cols, double resolution, double offset_x, double offset_y)
{
    return upload_grid_impl(h, map_id, dense, false, rows, cols, resolution, offset_x, offset_y);
}

int csm_upload_grid_device(csm_handle h, int64_t map_id, const uint16_t* dense_dev,
                           int rows, int cols, double resolution, double offset_x, double offset_y)
{
    return upload_grid_impl(h, map_id, dense_dev, true, rows, cols, resolution, offset_x, offset_y);
}

int csm_upload_grids(csm_handle h, int n, const int64_t* map_ids, const uint16_t* const* dense,
                     int rows, int cols, double resolution,
                     const double* offset_x, const double* offset_y)
{
    if (!h) return CSM_E_INVALID;
    if (n <= 0 || !map_ids || !dense || !offset_x || !offset_y)
        return fail(h, CSM_E_INVALID, "upload_grids: empty batch");
    if (rows <= 0 || cols <= 0 || rows > 16384 || cols > 16384 || (cols & 1) || !(resolution > 0.0))
        return fail(h, CSM_E_INVALID, "grid: need 0 < rows, cols <= 16384, even cols, resolution > 0");
    CSM_CUDA(cudaSetDevice(h->device));
    const size_t bytes = (size_t)rows * cols * sizeof(uint16_t);
    bool host_contiguous = true;
    for (int i = 1; i < n; ++i)
        host_contiguous = host_contiguous && dense[i] != nullptr &&
            reinterpret_cast<const char*>(dense[i]) == reinterpret_cast<const char*>(dense[0]) + (size_t)i * bytes;
    if (!host_contiguous || n == 1 || dense[0] == nullptr) {
        for (int i = 0; i < n; ++i) {
            const int rc = upload_grid_impl(h, map_ids[i], dense[i], false, rows, cols, resolution,
                                            offset_x[i], offset_y[i]);
            if (rc) return rc;
        }
        return close_upload_group(h);
    }
    std::vector<MapSlot*> slots;
    bool fresh_alloc = false;
    {
        const int arc = bind_batch_arena(h, n, map_ids, rows, cols, slots, fresh_alloc);
        if (arc) return arc;
    }
    for (int i = 0; i < n; ++i) {
        MapSlot& m = *slots[i];
        m.hmax = 0;
        m.bounds_levels = -1;
        m.margin_valid = false;
        m.coarse_win = 0;
        m.wide_valid = false;
        m.res = resolution; m.offx = offset_x[i]; m.offy = offset_y[i];
        m.pending_upload = reinterpret_cast<cudaEvent_t>(1);
        m.pending_scatter.reset();
        m.alloc_valid = false;
    }
    if (!h->upload_open || fresh_alloc) {
        CSM_CUDA(cudaEventRecord(h->compute_mark, h->stream));
        CSM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->compute_mark, 0));
        h->upload_open = true;
    }
    CSM_CUDA(cudaMemcpyAsync(slots[0]->base, dense[0], bytes * n, cudaMemcpyHostToDevice, h->copy_stream));
    {
        const int src = saturated_pass(h, slots[0]->base, bytes * n, h->copy_stream);
        if (src) return src;
    }
    /* one upload group per batch call: consumers of these maps wait for this
     * batch only, later batches keep streaming in behind the kernels */
    return close_upload_group(h);
}

int csm_upload_grids_blocks(csm_handle h, int n, const int64_t* map_ids,
                            const uint16_t* blocks, const int32_t* block_index,
                            const int32_t* block_count, int log2_block_size,
                            int block_rows, int block_cols, double resolution,
                            const double* offset_x, const double* offset_y)
{
    if (!h) return CSM_E_INVALID;
    if (n <= 0 || !map_ids || !block_count || !offset_x || !offset_y)
        return fail(h, CSM_E_INVALID, "upload_grids_blocks: empty batch");
    if (log2_block_size < 3 || log2_block_size > 6)
        return fail(h, CSM_E_UNSUPPORTED, "upload_grids_blocks: block size must be 8, 16, 32 or 64");
    if (block_rows <= 0 || block_cols <= 0 || !(resolution > 0.0))
        return fail(h, CSM_E_INVALID, "upload_grids_blocks: bad geometry");
    const long long rows_ll = (long long)block_rows << log2_block_size;
    const long long cols_ll = (long long)block_cols << log2_block_size;
    if (rows_ll > 16384 || cols_ll > 16384)
        return fail(h, CSM_E_INVALID, "grid: need rows, cols <= 16384");
    const int rows = (int)rows_ll, cols = (int)cols_ll;
    const int nb_map = block_rows * block_cols;
    auto bs = std::make_shared<BlockScatter>();
    bs->prefix.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        if (block_count[i] < 0 || block_count[i] > nb_map)
            return fail(h, CSM_E_INVALID, "upload_grids_blocks: block count out of range");
        bs->prefix[i + 1] = bs->prefix[i] + block_count[i];
        bs->max_count = std::max(bs->max_count, (int)block_count[i]);
    }
    const int total = bs->prefix[n];
    if (total > 0 && (!blocks || !block_index))
        return fail(h, CSM_E_INVALID, "upload_grids_blocks: null block data");
    for (int i = 0; i < n; ++i) {
        /* every listed block inside the map, no block listed twice */
        std::vector<char> seen(nb_map, 0);
        for (int b = bs->prefix[i]; b < bs->prefix[i + 1]; ++b) {
            const int bi = block_index[b];
            if (bi < 0 || bi >= nb_map || seen[bi])
                return fail(h, CSM_E_INVALID, "upload_grids_blocks: bad or repeated block index");
            seen[bi] = 1;
        }
    }
    CSM_CUDA(cudaSetDevice(h->device));
    std::vector<MapSlot*> slots;
    bool fresh_alloc = false;
    {
        const int arc = bind_batch_arena(h, n, map_ids, rows, cols, slots, fresh_alloc);
        if (arc) return arc;
    }
    {
        /* one byte per block and map, set by the expansion kernel from the block list */
        bool reuse = slots[0]->alloc_block != nullptr && slots[0]->alloc == slots[0]->alloc_block->p;
        for (int i = 0; i < n && reuse; ++i)
            reuse = slots[i]->alloc_block == slots[0]->alloc_block && slots[i]->alloc_bytes == nb_map &&
                    slots[i]->alloc == static_cast<unsigned char*>(slots[0]->alloc_block->p) + (size_t)i * nb_map;
        if (!reuse) {
            auto block = std::make_shared<ArenaBlock>();
            block->stream = h->stream;
            CSM_CUDA(cudaMallocAsync(&block->p, (size_t)nb_map * n, h->stream));
            for (int i = 0; i < n; ++i) {
                MapSlot& m = *slots[i];
                if (m.alloc && !m.alloc_block) CSM_CUDA(cudaFreeAsync(m.alloc, h->stream));
                m.alloc_block = block;
                m.alloc = static_cast<unsigned char*>(block->p) + (size_t)i * nb_map;
                m.alloc_bytes = nb_map;
            }
        }
        for (int i = 0; i < n; ++i) {
            slots[i]->alloc_log2bs = log2_block_size;
            slots[i]->alloc_valid = true;      /* once the pending expansion has run */
        }
        bs->alloc = slots[0]->alloc;
        bs->nb_map = nb_map;
    }
    const size_t block_bytes = sizeof(uint16_t) << (2 * log2_block_size);
    bs->data_bytes = block_bytes * (size_t)total;
    bs->index_off = (bs->data_bytes + 255) & ~(size_t)255;
    bs->prefix_off = (bs->index_off + sizeof(int) * (size_t)total + 255) & ~(size_t)255;
    const size_t stage_bytes = bs->prefix_off + sizeof(int) * (size_t)(n + 1);
    bs->stream = h->stream;
    CSM_CUDA(cudaMallocAsync(&bs->stage, stage_bytes, h->stream));
    bs->dense = slots[0]->base;
    bs->n_maps = n;
    bs->log2bs = log2_block_size; bs->block_cols = block_cols; bs->rows = rows; bs->cols = cols;
    for (int i = 0; i < n; ++i) {
        MapSlot& m = *slots[i];
        m.hmax = 0;
        m.bounds_levels = -1;
        m.margin_valid = false;
        m.coarse_win = 0;
        m.wide_valid = false;
        m.res = resolution; m.offx = offset_x[i]; m.offy = offset_y[i];
        m.pending_upload = reinterpret_cast<cudaEvent_t>(1);
        m.pending_scatter = bs;
    }
    /* the staging buffer was allocated on the compute stream: order the copies after it */
    CSM_CUDA(cudaEventRecord(h->compute_mark, h->stream));
    CSM_CUDA(cudaStreamWaitEvent(h->copy_stream, h->compute_mark, 0));
    h->upload_open = true;
    char* stage = static_cast<char*>(bs->stage);
    if (total > 0) {
        CSM_CUDA(cudaMemcpyAsync(stage, blocks, bs->data_bytes, cudaMemcpyHostToDevice, h->copy_stream));
        CSM_CUDA(cudaMemcpyAsync(stage + bs->index_off, block_index, sizeof(int) * (size_t)total,
                                 cudaMemcpyHostToDevice, h->copy_stream));
    }
    CSM_CUDA(cudaMemcpyAsync(stage + bs->prefix_off, bs->prefix.data(), sizeof(int) * (size_t)(n + 1),
                             cudaMemcpyHostToDevice, h->copy_stream));
    if (h->timing == 2) phase_mark(h, "h2d blocks (copy stream)", h->copy_stream);
    return close_upload_group(h);
}

int csm_upload_grid_blocks(csm_handle h, int64_t map_id, const uint16_t* blocks,
                           const int32_t* block_index, int n_blocks, int log2_block_size,
                           int block_rows, int block_cols, double resolution,
                           double offset_x, double offset_y)
{
    const int32_t count = n_blocks;
    return csm_upload_grids_blocks(h, 1, &map_id, blocks, block_index, &count, log2_block_size,
                                   block_rows, block_cols, resolution, &offset_x, &offset_y);
}

int csm_release_grid(csm_handle h, int64_t map_id)
{
    if (!h) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    auto it = h->maps.find(map_id);
    if (it == h->maps.end())
        return fail(h, CSM_E_NOT_FOUND, "release: unknown map id");
    {
        const int wrc = wait_uploads(h, std::vector<MapSlot*>{ &it->second });
        if (wrc) return wrc;
    }
    free_map(h, it->second);
    h->maps.erase(it);
    return CSM_OK;
}

int csm_build_coarse(csm_handle h, int64_t map_id, int win)
{
    if (!h) return CSM_E_INVALID;
    auto it = h->maps.find(map_id);
    if (it == h->maps.end())
        return fail(h, CSM_E_NOT_FOUND, "build_coarse: unknown map id");
    if (win <= 0 || win > 4096)
        return fail(h, CSM_E_INVALID, "build_coarse: window must be in 1..4096");
    CSM_CUDA(cudaSetDevice(h->device));
    MapSlot& m = it->second;
    if (m.coarse_win == win && m.coarse != nullptr)
        return CSM_OK;
    {
        const int wrc = wait_uploads(h, std::vector<MapSlot*>{ &m });
        if (wrc) return wrc;
    }
    if (m.coarse == nullptr)
        CSM_CUDA(cudaMallocAsync((void**)&m.coarse, (size_t)m.rows * m.cols * sizeof(uint16_t), h->stream));
    if (win >= 2 && win <= kSmMaxWin) {
        dim3 grid((m.cols + kSmCols - 1) / kSmCols, (m.rows + kSmRows - 1) / kSmRows);
        k_sliding_max_tile<<<grid, 256, 0, h->stream>>>(m.base, m.coarse, m.rows, m.cols, win);
    } else {
        dim3 grid((m.cols + 255) / 256, m.rows);
        k_sliding_max<<<grid, 256, 0, h->stream>>>(m.base, m.coarse, m.rows, m.cols, win);
    }
    CSM_LAUNCH_CHECK();
    m.coarse_win = win;
    return CSM_OK;
}

int csm_build_pyramids(csm_handle h, int n, const int64_t* map_ids, int hmax)
{
    if (!h) return CSM_E_INVALID;
    if (n <= 0 || map_ids == nullptr)
        return fail(h, CSM_E_INVALID, "build_pyramids: empty id list");
    CSM_CUDA(cudaSetDevice(h->device));
    std::vector<MapSlot*> slots;
    slots.reserve(n);
    for (int i = 0; i < n; ++i) {
        auto it = h->maps.find(map_ids[i]);
        if (it == h->maps.end())
            return fail(h, CSM_E_NOT_FOUND, "build_pyramids: unknown map id " + std::to_string(map_ids[i]));
        slots.push_back(&it->second);
    }
    /* a batch of maps is precomputed for the batched search: its bound levels (csm_bounds.cuh) unless the
     * sweep is set to read the reference's u16 levels; one map alone gets the reference's levels */
    if (n > 1 && h->bb_bounds && h->bb_skip_top && h->bb_dive != 1 && hmax >= 1 && hmax < kMaxLevels)
        return build_bounds(h, slots, hmax - 1);
    return build_levels(h, slots, hmax);
}

int csm_drop_pyramids(csm_handle h, int n, const int64_t* map_ids)
{
    if (!h) return CSM_E_INVALID;
    for (int i = 0; i < n; ++i) {
        auto it = h->maps.find(map_ids[i]);
        if (it == h->maps.end())
            return fail(h, CSM_E_NOT_FOUND, "drop_pyramids: unknown map id");
        it->second.hmax = 0;
        it->second.bounds_levels = -1;
        it->second.coarse_win = 0;
        it->second.wide_valid = false;
    }
    return CSM_OK;
}

int csm_build_pyramid(csm_handle h, int64_t map_id, int hmax)
{
    return csm_build_pyramids(h, 1, &map_id, hmax);
}

int csm_download_level(csm_handle h, int64_t map_id, int level, uint16_t* out)
{
    if (!h || !out) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    auto it = h->maps.find(map_id);
    if (it == h->maps.end())
        return fail(h, CSM_E_NOT_FOUND, "download: unknown map id");
    MapSlot& m = it->second;
    {
        const int wrc = wait_uploads(h, std::vector<MapSlot*>{ &m });
        if (wrc) return wrc;
    }
    if (level > 0 && level > m.hmax && level <= m.bounds_levels + 1) {
        /* the map was precomputed for the sweep only: build the reference's levels now */
        const int brc = build_levels(h, std::vector<MapSlot*>{ &m }, level);
        if (brc) return brc;
    }
    const size_t cells = (size_t)m.rows * m.cols;
    const uint16_t* src = nullptr;
    if (level == 0) src = m.base;
    else if (level > 0 && level <= m.hmax) src = m.levels + (size_t)(level - 1) * cells;
    else if (level < 0 && m.coarse_win == -level) src = m.coarse;
    if (src == nullptr)
        return fail(h, CSM_E_INVALID, "download: level not built");
    CSM_CUDA(cudaMemcpyAsync(out, src, cells * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));
    return CSM_OK;
}

int csm_upload_scan(csm_handle h, int64_t scan_id, const double* angles, const double* ranges, int n)
{
    if (!h) return CSM_E_INVALID;
    if (scan_id == kTempScanId)
        return fail(h, CSM_E_INVALID, "scan id reserved");
    CSM_CUDA(cudaSetDevice(h->device));
    return upload_scan_impl(h, scan_id, angles, ranges, n);
}

int csm_release_scan(csm_handle h, int64_t scan_id)
{
    if (!h) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    auto it = h->scans.find(scan_id);
    if (it == h->scans.end())
        return fail(h, CSM_E_NOT_FOUND, "release: unknown scan id");
    free_scan(h, it->second);
    h->scans.erase(it);
    return CSM_OK;
}

int csm_loop_batch_enqueue(csm_handle h, const csm_loop_query* queries, int nq, int hmax,
                           int query_index_base)
{
    if (!h) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    return bb_enqueue(h, queries, nq, hmax, query_index_base, nullptr);
}

int csm_loop_batch_finish(csm_handle h, csm_result* results, int nq)
{
    if (!h || !results) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    return finish_results(h, results, nq);
}

int csm_set_refiner(csm_handle h, const csm_refine_params* p)
{
    if (!h) return CSM_E_INVALID;
    if (p == nullptr) {
        h->refine_on = false;
        return CSM_OK;
    }
    const int rc = check_refine_params(h, p);
    if (rc) return rc;
    h->refine = *p;
    h->refine_on = true;
    return CSM_OK;
}

int csm_share_copy_stream(csm_handle h, csm_handle owner)
{
    if (!h || !owner || h == owner) return CSM_E_INVALID;
    if (h->device != owner->device)
        return fail(h, CSM_E_INVALID, "share_copy_stream: handles on different devices");
    CSM_CUDA(cudaSetDevice(h->device));
    CSM_CUDA(cudaStreamSynchronize(h->copy_stream));
    if (h->owns_copy_stream) CSM_CUDA(cudaStreamDestroy(h->copy_stream));
    h->copy_stream = owner->copy_stream;
    h->owns_copy_stream = false;
    return CSM_OK;
}

int csm_set_epilogue(csm_handle h, double covariance_scale)
{
    if (!h) return CSM_E_INVALID;
    if (!(covariance_scale >= 0.0))
        return fail(h, CSM_E_INVALID, "epilogue: covariance_scale must be >= 0");
    h->epilogue_scale = covariance_scale;
    std::memset(&h->last_epilogue, 0, sizeof(h->last_epilogue));
    h->last_epilogue_set = false;
    return CSM_OK;
}

int csm_last_epilogue(csm_handle h, csm_refined* out)
{
    if (!h || !out) return CSM_E_INVALID;
    if (!h->last_epilogue_set)
        return fail(h, CSM_E_INVALID, "epilogue: no single-scan match with csm_set_epilogue / csm_set_refiner on has run");
    *out = h->last_epilogue;
    return CSM_OK;
}

int csm_loop_batch_finish_refined(csm_handle h, csm_result* results, csm_refined* refined, int nq)
{
    if (!h || !results || !refined) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    return finish_results(h, results, nq, refined);
}

int csm_refine_batch(csm_handle h, const csm_refine_query* queries, int n,
                     const csm_refine_params* p, csm_refined* out)
{
    if (!h) return CSM_E_INVALID;
    if (n <= 0 || !queries || !p || !out)
        return fail(h, CSM_E_INVALID, "refine batch: empty");
    const int rc = check_refine_params(h, p);
    if (rc) return rc;
    CSM_CUDA(cudaSetDevice(h->device));
    return refine_poses(h, queries, n, p, out);
}

int csm_loop_batch(csm_handle h, const csm_loop_query* queries, int nq, int hmax,
                   int query_index_base, csm_result* results)
{
    int rc = csm_loop_batch_enqueue(h, queries, nq, hmax, query_index_base);
    if (rc) return rc;
    return csm_loop_batch_finish(h, results, nq);
}

/* ---- map construction on the device -------------------------------------------------------------- */
int csm_map_set_update_tables(csm_handle h, const uint16_t* t_miss, const uint16_t* t_hit)
{
    if (!h || !t_miss || !t_hit) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    /* the tables of 2^j updates in a row, j < kMapPowTables: composition of the table with itself */
    const size_t one = 65536;
    std::vector<uint16_t> pw(2 * (size_t)kMapPowTables * one);
    for (int kind = 0; kind < 2; ++kind) {
        uint16_t* t = pw.data() + (size_t)kind * kMapPowTables * one;
        std::memcpy(t, kind ? t_hit : t_miss, one * sizeof(uint16_t));
        for (int j = 1; j < kMapPowTables; ++j)
            for (size_t v = 0; v < one; ++v)
                t[j * one + v] = t[(j - 1) * one + t[(j - 1) * one + v]];
    }
    int rc = ensure(h, h->d_maptables, pw.size() * sizeof(uint16_t));
    if (rc) return rc;
    CSM_CUDA(cudaMemcpyAsync(h->d_maptables.p, pw.data(), pw.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, h->stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));       /* pw is about to go */
    h->map_tables_set = true;
    return CSM_OK;
}

static void invalidate_derived(MapSlot& m)
{
    m.hmax = 0;
    m.bounds_levels = -1;
    m.margin_valid = false;
    m.coarse_win = 0;
    m.wide_valid = false;
}

int csm_map_create(csm_handle h, int64_t map_id, int rows, int cols, int log2_block_size,
                   double resolution, double offset_x, double offset_y)
{
    if (!h) return CSM_E_INVALID;
    if (rows <= 0 || cols <= 0 || rows > 16384 || cols > 16384 || log2_block_size < 3 || log2_block_size > 6 ||
        (rows & ((1 << log2_block_size) - 1)) || (cols & ((1 << log2_block_size) - 1)) || !(resolution > 0.0))
        return fail(h, CSM_E_INVALID, "map: rows / cols must be positive multiples of the block size (8..64), resolution > 0");
    CSM_CUDA(cudaSetDevice(h->device));
    MapSlot& m = h->maps[map_id];
    int rc = settle_slot(h, m);
    if (rc) return rc;
    free_map(h, m);
    const size_t bytes = (size_t)rows * cols * sizeof(uint16_t);
    CSM_CUDA(cudaMallocAsync((void**)&m.base, bytes, h->stream));
    CSM_CUDA(cudaMemsetAsync(m.base, 0, bytes, h->stream));
    CSM_CUDA(cudaMallocAsync((void**)&m.raw, bytes, h->stream));
    CSM_CUDA(cudaMemsetAsync(m.raw, 0, bytes, h->stream));
    m.rows = rows; m.cols = cols;
    m.res = resolution; m.offx = offset_x; m.offy = offset_y;
    m.alloc_log2bs = log2_block_size;
    m.alloc_bytes = (rows >> log2_block_size) * (cols >> log2_block_size);
    CSM_CUDA(cudaMallocAsync((void**)&m.alloc, (size_t)m.alloc_bytes, h->stream));
    CSM_CUDA(cudaMemsetAsync(m.alloc, 0, (size_t)m.alloc_bytes, h->stream));
    m.alloc_valid = true;
    invalidate_derived(m);
    return CSM_OK;
}

int csm_map_resize(csm_handle h, int64_t map_id, int rows, int cols, int row_min, int col_min,
                   double offset_x, double offset_y)
{
    if (!h) return CSM_E_INVALID;
    auto it = h->maps.find(map_id);
    if (it == h->maps.end() || it->second.alloc == nullptr || !it->second.alloc_valid || it->second.alloc_block)
        return fail(h, CSM_E_NOT_FOUND, "map resize: not a map made by csm_map_create");
    MapSlot& m = it->second;
    const int k = m.alloc_log2bs, mask = (1 << k) - 1;
    if (rows <= 0 || cols <= 0 || rows > 16384 || cols > 16384 || (rows & mask) || (cols & mask) ||
        (row_min & mask) || (col_min & mask))
        return fail(h, CSM_E_INVALID, "map resize: extents and origin must be multiples of the block size");
    CSM_CUDA(cudaSetDevice(h->device));
    int rc = settle_slot(h, m);
    if (rc) return rc;
    uint16_t* base = nullptr;
    uint16_t* raw = nullptr;
    unsigned char* alloc = nullptr;
    const int nb = (rows >> k) * (cols >> k);
    CSM_CUDA(cudaMallocAsync((void**)&base, (size_t)rows * cols * sizeof(uint16_t), h->stream));
    CSM_CUDA(cudaMallocAsync((void**)&raw, (size_t)rows * cols * sizeof(uint16_t), h->stream));
    CSM_CUDA(cudaMallocAsync((void**)&alloc, (size_t)nb, h->stream));
    MapMoveArgs A;
    A.src = m.raw; A.dst = raw; A.dst_view = base; A.saturated_unknown = h->saturated_unknown;
    A.src_alloc = m.alloc; A.dst_alloc = alloc;
    A.src_rows = m.rows; A.src_cols = m.cols; A.dst_rows = rows; A.dst_cols = cols;
    A.row_min = row_min; A.col_min = col_min; A.log2bs = k;
    dim3 grid((cols + 255) / 256, rows);
    k_map_move<<<grid, 256, 0, h->stream>>>(A);
    CSM_LAUNCH_CHECK();
    if (m.levels) { CSM_CUDA(cudaFreeAsync(m.levels, h->stream)); m.levels = nullptr; m.levels_alloc = 0; }
    if (m.coarse) { CSM_CUDA(cudaFreeAsync(m.coarse, h->stream)); m.coarse = nullptr; }
    if (m.wide) { CSM_CUDA(cudaFreeAsync(m.wide, h->stream)); m.wide = nullptr; m.wide_valid = false; }
    if (m.bounds) { CSM_CUDA(cudaFreeAsync(m.bounds, h->stream)); m.bounds = nullptr; m.bounds_alloc = 0; }
    CSM_CUDA(cudaFreeAsync(m.base, h->stream));
    CSM_CUDA(cudaFreeAsync(m.raw, h->stream));
    CSM_CUDA(cudaFreeAsync(m.alloc, h->stream));
    m.base = base; m.raw = raw; m.alloc = alloc; m.alloc_bytes = nb;
    m.rows = rows; m.cols = cols; m.offx = offset_x; m.offy = offset_y;
    invalidate_derived(m);
    return CSM_OK;
}

int csm_map_reset_values(csm_handle h, int64_t map_id)
{
    if (!h) return CSM_E_INVALID;
    auto it = h->maps.find(map_id);
    if (it == h->maps.end() || it->second.base == nullptr)
        return fail(h, CSM_E_NOT_FOUND, "map reset: unknown map id");
    MapSlot& m = it->second;
    CSM_CUDA(cudaSetDevice(h->device));
    int rc = settle_slot(h, m);
    if (rc) return rc;
    CSM_CUDA(cudaMemsetAsync(m.base, 0, (size_t)m.rows * m.cols * sizeof(uint16_t), h->stream));
    if (m.raw)
        CSM_CUDA(cudaMemsetAsync(m.raw, 0, (size_t)m.rows * m.cols * sizeof(uint16_t), h->stream));
    invalidate_derived(m);
    return CSM_OK;
}

int csm_map_insert_rays(csm_handle h, int64_t map_id, const csm_ray* rays, int n, int subpixel_scale)
{
    if (!h) return CSM_E_INVALID;
    if (n < 0 || (n > 0 && !rays) || subpixel_scale <= 0)
        return fail(h, CSM_E_INVALID, "map insert: bad arguments");
    if (!h->map_tables_set)
        return fail(h, CSM_E_INVALID, "map insert: csm_map_set_update_tables first");
    auto it = h->maps.find(map_id);
    if (it == h->maps.end() || it->second.alloc == nullptr || !it->second.alloc_valid || it->second.alloc_block)
        return fail(h, CSM_E_NOT_FOUND, "map insert: not a map made by csm_map_create");
    if (it->second.raw == nullptr)
        return fail(h, CSM_E_NOT_FOUND, "map insert: not a map made by csm_map_create");
    if (n == 0)
        return CSM_OK;
    MapSlot& m = it->second;
    CSM_CUDA(cudaSetDevice(h->device));
    int rc = settle_slot(h, m);
    if (rc) return rc;
    /* event slots per beam: the cells of its ray (at most |dx| + |dy| + 1 full cells) and its hit */
    const size_t rays_bytes = align16(sizeof(csm_ray) * (size_t)n);
    const size_t off_bytes = align16(sizeof(unsigned int) * (size_t)(n + 1));
    char* hp = nullptr;
    if ((rc = acquire_upload(h, rays_bytes + off_bytes, &hp))) return rc;
    std::memcpy(hp, rays, sizeof(csm_ray) * (size_t)n);
    unsigned int* off = reinterpret_cast<unsigned int*>(hp + rays_bytes);
    unsigned long long total = 0;
    /* |a / s - b / s| <= |a - b| / s + 1 for the truncating division: an upper bound without a division per
     * coordinate (the reciprocal is rounded up a little so that the product never falls short) */
    const double inv_scale = (1.0 + 1e-9) / (double)subpixel_scale;
    for (int i = 0; i < n; ++i) {
        off[i] = (unsigned int)total;
        const long long span = std::llabs((long long)rays[i].end_x - rays[i].start_x) +
                               std::llabs((long long)rays[i].end_y - rays[i].start_y);
        total += (unsigned long long)((double)span * inv_scale) + 6;
    }
    off[n] = (unsigned int)total;
    if (total >= (1ull << 31))
        return fail(h, CSM_E_UNSUPPORTED, "map insert: too many cell updates in one call");
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                   (int)total, 0, 64, h->stream);
    const size_t ev_bytes = align16(sizeof(unsigned long long) * (size_t)total);
    sort_bytes = align16(sort_bytes);
    if ((rc = ensure(h, h->d_mapwork, rays_bytes + off_bytes + 2 * ev_bytes + sort_bytes + 256))) return rc;
    if (h->d_maperror.p == nullptr) {
        if ((rc = ensure(h, h->d_maperror, 16))) return rc;
        CSM_CUDA(cudaMemsetAsync(h->d_maperror.p, 0, 16, h->stream));
        CSM_CUDA(cudaHostAlloc((void**)&h->h_maperror, 16, cudaHostAllocDefault));
        *h->h_maperror = 0;
    }
    char* w = static_cast<char*>(h->d_mapwork.p);
    if ((rc = pull_to_device(h, w, hp, rays_bytes + off_bytes))) return rc;
    if ((rc = upload_committed(h))) return rc;
    unsigned long long* ev_in = reinterpret_cast<unsigned long long*>(w + rays_bytes + off_bytes);
    unsigned long long* ev_out = reinterpret_cast<unsigned long long*>(w + rays_bytes + off_bytes + ev_bytes);
    void* sort_tmp = w + rays_bytes + off_bytes + 2 * ev_bytes;
    MapRaysArgs R;
    R.rays = reinterpret_cast<const csm_ray*>(w);
    R.offset = reinterpret_cast<const unsigned int*>(w + rays_bytes);
    R.events = ev_in;
    R.n = n; R.scale = subpixel_scale; R.rows = m.rows; R.cols = m.cols;
    R.error = static_cast<int*>(h->d_maperror.p);
    /* `order` must be below 2^(shift - 1): the adapter numbers the beams of one call 0 .. n - 1 */
    int order_max = 0;
    bool ascending = true;          /* rays handed over in update order (what the adapter does) */
    for (int i = 0; i < n; ++i) {
        if (rays[i].order < 0)
            return fail(h, CSM_E_INVALID, "map insert: negative beam order");
        ascending = ascending && rays[i].order >= order_max;
        order_max = std::max(order_max, rays[i].order);
    }
    int order_bits = 1;
    while ((1ll << order_bits) <= (long long)order_max) ++order_bits;
    int cell_bits = 1;
    while ((1ull << cell_bits) < (unsigned long long)m.rows * m.cols) ++cell_bits;
    R.shift = order_bits + 1;
    if (R.shift + cell_bits + 1 > 64)
        return fail(h, CSM_E_UNSUPPORTED, "map insert: cell index and beam order do not fit one event");
    CSM_CUDA(cudaMemsetAsync(ev_in, 0xff, sizeof(unsigned long long) * (size_t)total, h->stream));    /* kMapEventNone */
    k_map_rays<<<n, kMapRayThreads, 0, h->stream>>>(R);
    CSM_LAUNCH_CHECK();
    /* Per cell, in beam order. When the rays arrive in update order their events are WRITTEN in that order
     * (a ray's slots follow the previous ray's, its hit is its last slot) and the radix sort is stable, so
     * sorting on the cell index alone keeps every cell's events in the reference's order; otherwise the order
     * bits are sorted too. One bit above the cell index sends the unused slots (all ones) to the end. */
    if (cub::DeviceRadixSort::SortKeys(sort_tmp, sort_bytes, ev_in, ev_out, (int)total, ascending ? R.shift : 0,
                                       std::min(64, R.shift + cell_bits + 1), h->stream) != cudaSuccess)
        return fail(h, CSM_E_CUDA, "map insert: radix sort failed");
    ++h->launches;
    MapApplyArgs P;
    P.events = ev_out; P.n = (unsigned int)total;
    P.map = m.raw; P.view = m.base; P.saturated_unknown = h->saturated_unknown; P.alloc = m.alloc;
    P.lut = static_cast<const uint16_t*>(h->d_maptables.p);
    P.cols = m.cols; P.log2bs = m.alloc_log2bs; P.block_cols = m.cols >> m.alloc_log2bs; P.shift = R.shift;
    k_map_apply<<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(P);
    CSM_LAUNCH_CHECK();
    CSM_CUDA(cudaMemcpyAsync(h->h_maperror, h->d_maperror.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    invalidate_derived(m);
    return CSM_OK;
}

int csm_map_download_cells(csm_handle h, int64_t map_id, uint16_t* out)
{
    if (!h || !out) return CSM_E_INVALID;
    auto it = h->maps.find(map_id);
    if (it == h->maps.end() || it->second.raw == nullptr)
        return fail(h, CSM_E_NOT_FOUND, "map cells: not a map made by csm_map_create");
    CSM_CUDA(cudaSetDevice(h->device));
    MapSlot& m = it->second;
    CSM_CUDA(cudaMemcpyAsync(out, m.raw, (size_t)m.rows * m.cols * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
    return csm_synchronize(h);
}

int csm_map_download_allocation(csm_handle h, int64_t map_id, uint8_t* out)
{
    if (!h || !out) return CSM_E_INVALID;
    auto it = h->maps.find(map_id);
    if (it == h->maps.end() || it->second.alloc == nullptr)
        return fail(h, CSM_E_NOT_FOUND, "map allocation: unknown map id");
    CSM_CUDA(cudaSetDevice(h->device));
    int rc = wait_uploads(h, std::vector<MapSlot*>{ &it->second });
    if (rc) return rc;
    CSM_CUDA(cudaMemcpyAsync(out, it->second.alloc, (size_t)it->second.alloc_bytes, cudaMemcpyDeviceToHost, h->stream));
    return csm_synchronize(h);
}

/* ---- NCCL exchange of the packed best word ------------------------------------------------- */
static int comm_fail(csm_handle h, const char* what, int rc)
{
    return fail(h, CSM_E_CUDA, std::string(what) + ": " + (nccl::api().GetErrorString ? nccl::api().GetErrorString(rc) : "nccl"));
}

static int comm_setup(csm_handle h)
{
    int lo = 0, hi = 0;
    CSM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    /* the 8-byte all-reduce runs beside kernels that fill every SM: highest priority, so that its CTA
     * is placed as soon as a slot frees */
    CSM_CUDA(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, hi));
    CSM_CUDA(cudaMalloc((void**)&h->d_words, sizeof(unsigned long long) * csm_context::kCommRing));
    CSM_CUDA(cudaMemset(h->d_words, 0, sizeof(unsigned long long) * csm_context::kCommRing));
    CSM_CUDA(cudaHostAlloc((void**)&h->h_words, sizeof(unsigned long long) * csm_context::kCommRing, cudaHostAllocDefault));
    for (int i = 0; i < csm_context::kCommRing; ++i) {
        CSM_CUDA(cudaEventCreateWithFlags(&h->comm_ready[i], cudaEventDisableTiming));
        CSM_CUDA(cudaEventCreateWithFlags(&h->comm_done[i], cudaEventDisableTiming));
    }
    const bool fresh = h->d_bestkey.p == nullptr;
    const int rc = ensure(h, h->d_bestkey, 8);
    if (rc == CSM_OK && fresh)
        CSM_CUDA(cudaMemsetAsync(h->d_bestkey.p, 0, 8, h->stream));
    return rc;
}

int csm_comm_unique_id(void* id128)
{
    if (!id128 || !nccl::api().ok) return CSM_E_UNSUPPORTED;
    nccl::unique_id id;
    if (nccl::api().GetUniqueId(&id) != 0) return CSM_E_CUDA;
    std::memcpy(id128, &id, sizeof(id));
    return CSM_OK;
}

int csm_comm_init_rank(csm_handle h, const void* id128, int rank, int world)
{
    if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return CSM_E_INVALID;
    if (!nccl::api().ok) return fail(h, CSM_E_UNSUPPORTED, "libnccl.so.2 not found");
    if (h->comm) return fail(h, CSM_E_INVALID, "communicator already initialised");
    CSM_CUDA(cudaSetDevice(h->device));
    nccl::unique_id id;
    std::memcpy(&id, id128, sizeof(id));
    const int rc = nccl::api().CommInitRank(&h->comm, world, id, rank);
    if (rc != 0) return comm_fail(h, "ncclCommInitRank", rc);
    h->comm_rank = rank; h->comm_world = world;
    return comm_setup(h);
}

int csm_comm_init_all(csm_handle* handles, int n)
{
    if (!handles || n < 1) return CSM_E_INVALID;
    for (int i = 0; i < n; ++i)
        if (!handles[i] || handles[i]->comm) return CSM_E_INVALID;
    csm_handle h = handles[0];
    if (!nccl::api().ok) return fail(h, CSM_E_UNSUPPORTED, "libnccl.so.2 not found");
    std::vector<int> devs(n);
    std::vector<nccl::comm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = handles[i]->device;
    const int rc = nccl::api().CommInitAll(comms.data(), n, devs.data());
    if (rc != 0) return comm_fail(h, "ncclCommInitAll", rc);
    for (int i = 0; i < n; ++i) {
        handles[i]->comm = comms[i];
        handles[i]->comm_rank = i; handles[i]->comm_world = n;
        if (cudaSetDevice(handles[i]->device) != cudaSuccess) return CSM_E_CUDA;
        const int src = comm_setup(handles[i]);
        if (src) return src;
    }
    return CSM_OK;
}

/* The exchange in three steps, so that several handles of one process can go through one NCCL group
 * (the collectives of a group are only launched by ncclGroupEnd: what must follow them on the stream is
 * enqueued after it): word of the last batch (or a host word) -> ring slot; all-reduce(max) in place on
 * the side stream; read-back and the "done" event. */
static int comm_prepare(csm_handle h, int* slot, const uint64_t* host_word)
{
    if (!h->comm) return fail(h, CSM_E_INVALID, "no communicator: call csm_comm_init_rank / csm_comm_init_all");
    const int i = h->comm_next;
    h->comm_next = (h->comm_next + 1) % csm_context::kCommRing;
    CSM_CUDA(cudaEventSynchronize(h->comm_done[i]));           /* the slot's previous exchange has been read */
    if (host_word != nullptr) {
        /* a word the caller formed on the host (a Detect over several lanes): no tie to the compute stream */
        h->h_words[i] = *host_word;
        CSM_CUDA(cudaMemcpyAsync(h->d_words + i, h->h_words + i, 8, cudaMemcpyHostToDevice, h->comm_stream));
    } else {
        CSM_CUDA(cudaMemcpyAsync(h->d_words + i, h->d_bestkey.p, 8, cudaMemcpyDeviceToDevice, h->stream));
        CSM_CUDA(cudaEventRecord(h->comm_ready[i], h->stream));
        CSM_CUDA(cudaStreamWaitEvent(h->comm_stream, h->comm_ready[i], 0));
    }
    *slot = i;
    return CSM_OK;
}

static int comm_reduce(csm_handle h, int i)
{
    const int rc = nccl::api().AllReduce(h->d_words + i, h->d_words + i, 1, nccl::kUint64, nccl::kMax, h->comm, h->comm_stream);
    if (rc != 0) return comm_fail(h, "ncclAllReduce", rc);
    return CSM_OK;
}

static int comm_complete(csm_handle h, int i)
{
    CSM_CUDA(cudaMemcpyAsync(h->h_words + i, h->d_words + i, 8, cudaMemcpyDeviceToHost, h->comm_stream));
    CSM_CUDA(cudaEventRecord(h->comm_done[i], h->comm_stream));
    return CSM_OK;
}

static int comm_enqueue(csm_handle h, int* ticket, const uint64_t* host_word = nullptr)
{
    int i = 0, rc;
    if ((rc = comm_prepare(h, &i, host_word))) return rc;
    if ((rc = comm_reduce(h, i))) return rc;
    if ((rc = comm_complete(h, i))) return rc;
    if (ticket) *ticket = i;
    return CSM_OK;
}

static int comm_enqueue_all(csm_handle* handles, int n, const uint64_t* words, int* tickets)
{
    std::vector<int> slot(n, 0);
    int rc;
    for (int i = 0; i < n; ++i) {
        if (cudaSetDevice(handles[i]->device) != cudaSuccess) return CSM_E_CUDA;
        if ((rc = comm_prepare(handles[i], &slot[i], words ? words + i : nullptr))) return rc;
    }
    int nrc = nccl::api().GroupStart();
    if (nrc != 0) return comm_fail(handles[0], "ncclGroupStart", nrc);
    int first = CSM_OK;
    for (int i = 0; i < n && first == CSM_OK; ++i) {
        if (cudaSetDevice(handles[i]->device) != cudaSuccess) { first = CSM_E_CUDA; break; }
        first = comm_reduce(handles[i], slot[i]);
    }
    nrc = nccl::api().GroupEnd();
    if (nrc != 0 && first == CSM_OK) first = comm_fail(handles[0], "ncclGroupEnd", nrc);
    if (first) return first;
    for (int i = 0; i < n; ++i) {
        if (cudaSetDevice(handles[i]->device) != cudaSuccess) return CSM_E_CUDA;
        if ((rc = comm_complete(handles[i], slot[i]))) return rc;
        if (tickets) tickets[i] = slot[i];
    }
    return CSM_OK;
}

int csm_comm_allreduce_best(csm_handle h, int* ticket)
{
    if (!h) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    return comm_enqueue(h, ticket);
}

int csm_comm_allreduce_word(csm_handle h, uint64_t word, int* ticket)
{
    if (!h) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    return comm_enqueue(h, ticket, &word);
}

int csm_comm_allreduce_words_all(csm_handle* handles, int n, const uint64_t* words, int* tickets)
{
    if (!handles || n < 1 || !words || !nccl::api().ok) return CSM_E_INVALID;
    return comm_enqueue_all(handles, n, words, tickets);
}

int csm_comm_allreduce_best_all(csm_handle* handles, int n, int* tickets)
{
    if (!handles || n < 1 || !nccl::api().ok) return CSM_E_INVALID;
    return comm_enqueue_all(handles, n, nullptr, tickets);
}

int csm_comm_best_result(csm_handle h, int ticket, uint64_t* word)
{
    if (!h || !word || ticket < 0 || ticket >= csm_context::kCommRing || !h->comm) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    CSM_CUDA(cudaEventSynchronize(h->comm_done[ticket]));
    *word = h->h_words[ticket];
    return CSM_OK;
}

int csm_comm_destroy(csm_handle h)
{
    if (!h) return CSM_E_INVALID;
    if (!h->comm) return CSM_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->comm_stream);
    nccl::api().CommDestroy(h->comm);
    h->comm = nullptr;
    for (int i = 0; i < csm_context::kCommRing; ++i) {
        if (h->comm_ready[i]) cudaEventDestroy(h->comm_ready[i]);
        if (h->comm_done[i]) cudaEventDestroy(h->comm_done[i]);
        h->comm_ready[i] = h->comm_done[i] = nullptr;
    }
    if (h->d_words) cudaFree(h->d_words);
    if (h->h_words) cudaFreeHost(h->h_words);
    h->d_words = nullptr; h->h_words = nullptr;
    cudaStreamDestroy(h->comm_stream);
    h->comm_stream = nullptr;
    return CSM_OK;
}

int csm_detect_step_enqueue(csm_handle h, const int64_t* map_ids, int n_maps, int drop,
                            const csm_loop_query* queries, int nq, int hmax, int query_index_base, int* ticket)
{
    if (!h) return CSM_E_INVALID;
    int rc;
    if (n_maps > 0) {
        if (drop && (rc = csm_drop_pyramids(h, n_maps, map_ids))) return rc;
        if ((rc = csm_build_pyramids(h, n_maps, map_ids, hmax))) return rc;
    }
    if ((rc = csm_loop_batch_enqueue(h, queries, nq, hmax, query_index_base))) return rc;
    if (h->comm && h->comm_world > 1)
        return comm_enqueue(h, ticket);
    if (ticket) *ticket = -1;
    return CSM_OK;
}

int csm_debug_bound_level(csm_handle h, int64_t map_id, int level, uint8_t* out)
{
    if (!h || !out) return CSM_E_INVALID;
    auto it = h->maps.find(map_id);
    if (it == h->maps.end())
        return fail(h, CSM_E_NOT_FOUND, "bound level: unknown map id");
    if (level < 1 || level > 6)
        return fail(h, CSM_E_INVALID, "bound level: 1 <= level <= 6");
    CSM_CUDA(cudaSetDevice(h->device));
    MapSlot& m = it->second;
    if (m.bounds_levels < level) {
        const int brc = build_bounds(h, std::vector<MapSlot*>{ &m }, level);
        if (brc) return brc;
    }
    const size_t bytes = bl_level_bytes(level, m.rows, m.cols);
    std::vector<unsigned char> raw(bytes);
    CSM_CUDA(cudaMemcpyAsync(raw.data(), m.bounds + bl_level_offset(level, m.rows, m.cols), bytes,
                             cudaMemcpyDeviceToHost, h->stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));
    const int tpr = bl_tiles_per_row(level, m.cols), pr = bl_pad_r(level), pc = bl_pad_c(level);
    /* everything outside the map must have stayed zero */
    for (int rp = 0; rp < bl_tile_rows(level, m.rows) * kBlTileR; ++rp)
        for (int cp = 0; cp < tpr * kBlTileC; ++cp) {
            const unsigned char v = raw[bl_cell((unsigned int)rp, (unsigned int)cp, (unsigned int)tpr)];
            const int r = rp - pr, c = cp - pc;
            if (r >= 0 && r < m.rows && c >= 0 && c < m.cols)
                out[(size_t)r * m.cols + c] = v;
            else if (v != 0)
                return fail(h, CSM_E_INVALID, "bound level: non-zero padding");
        }
    return CSM_OK;
}

int csm_debug_node_list(csm_handle h, int level, uint64_t* out, int cap)
{
    if (!h || !out || level < 0 || level >= kMaxLevels || !h->plan_view.counts) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    unsigned int counts[kMaxLevels];
    CSM_CUDA(cudaMemcpyAsync(counts, h->plan_view.counts, sizeof(counts), cudaMemcpyDeviceToHost, h->stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));
    const int n = (int)std::min<unsigned int>(std::min(counts[level], h->frontier_capacity), (unsigned int)std::max(cap, 0));
    if (n > 0) {
        CSM_CUDA(cudaMemcpyAsync(out, h->d_list[(level & 1) ^ 1].p, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, h->stream));
        CSM_CUDA(cudaStreamSynchronize(h->stream));
    }
    return n;
}

int csm_debug_frontier_counts(csm_handle h, unsigned int* out8)
{
    if (!h || !out8 || !h->plan_view.counts) return CSM_E_INVALID;
    CSM_CUDA(cudaSetDevice(h->device));
    CSM_CUDA(cudaMemcpyAsync(out8, h->plan_view.counts, sizeof(unsigned int) * kMaxLevels, cudaMemcpyDeviceToHost, h->stream));
    CSM_CUDA(cudaStreamSynchronize(h->stream));
    return CSM_OK;
}

int csm_debug_timings(csm_handle h, char* names, size_t names_cap, float* ms, int max_n)
{
    if (!h || !ms || max_n <= 0) return 0;
    if (h->tcount < 2) return 0;
    if (cudaEventSynchronize(h->tev[h->tcount - 1]) != cudaSuccess) return 0;
    std::string all;
    int n = 0;
    for (size_t i = 0; i < h->tcount; ++i)
        cudaEventSynchronize(h->tev[i]);
    for (size_t i = 1; i < h->tcount && n < max_n; ++i, ++n) {
        float t = 0.0f;
        /* mode 2: time since the first mark (a timeline across streams); mode 1: phase durations */
        cudaEventElapsedTime(&t, h->tev[h->timing == 2 ? 0 : i - 1], h->tev[i]);
        ms[n] = t;
        all += h->tnames[i];
        all += ';';
    }
    if (names && names_cap > 0) {
        std::strncpy(names, all.c_str(), names_cap - 1);
        names[names_cap - 1] = '\0';
    }
    return n;
}

void* csm_best_key_device(csm_handle h)
{
    return h ? h->d_bestkey.p : nullptr;
}

void csm_decode_best_key(uint64_t best_key, int64_t* key, int32_t* query_index)
{
    if (key) *key = (int64_t)(best_key >> 20);
    if (query_index) *query_index = best_key ? (int32_t)(0xFFFFF - (best_key & 0xFFFFF)) : -1;
}

int csm_match_bb(csm_handle h, int64_t map_id,
                 const double* angles, const double* ranges, int n,
                 const double sensor_pose[3], int hmax,
                 int win_x, int win_y, int win_t,
                 double step_x, double step_y, double step_t,
                 double score_thr, double known_thr, csm_result* out)
{
    if (!h || !out || !sensor_pose) return CSM_E_INVALID;
    if (h->res_count != 0)
        return fail(h, CSM_E_INVALID, "match_bb: a loop batch is in flight on this handle");
    CSM_CUDA(cudaSetDevice(h->device));
    csm_loop_query q;
    std::memset(&q, 0, sizeof(q));
    q.map_id = map_id;
    q.scan_id = kTempScanId;
    q.sensor_pose[0] = sensor_pose[0]; q.sensor_pose[1] = sensor_pose[1]; q.sensor_pose[2] = sensor_pose[2];
    q.win_x = win_x; q.win_y = win_y; q.win_t = win_t;
    q.step_x = step_x; q.step_y = step_y; q.step_t = step_t;
    q.score_thr = score_thr; q.known_thr = known_thr;
    const InlineScan scan { angles, ranges, n };
    const int rc = bb_enqueue(h, &q, 1, hmax, 0, &scan);
    if (rc) return rc;
    const bool extra = h->epilogue_scale > 0.0 || h->refine_on;
    const int frc = finish_results(h, out, 1, extra ? &h->last_epilogue : nullptr);
    h->last_epilogue_set = frc == CSM_OK && extra;
    return frc;
}

static int match_rt_impl(csm_handle h, int64_t map_id,
                         const double* angles, const double* ranges, int n,
                         const double sensor_pose[3], int low_res,
                         int win_x, int win_y, int win_t,
                         double step_x, double step_y, double step_t,
                         double score_thr, double known_thr, csm_result* out, bool exact)
{
    if (!h || !out || !sensor_pose) return CSM_E_INVALID;
    if (h->res_count != 0)
        return fail(h, CSM_E_INVALID, "match_rt: a loop batch is in flight on this handle");
    CSM_CUDA(cudaSetDevice(h->device));
    auto mi = h->maps.find(map_id);
    if (mi == h->maps.end())
        return fail(h, CSM_E_NOT_FOUND, "match_rt: unknown map id");
    MapSlot& m = mi->second;
    {
        const int wrc = wait_uploads(h, std::vector<MapSlot*>{ &m });
        if (wrc) return wrc;
    }
    if (low_res <= 0 || low_res > 64)
        return fail(h, CSM_E_UNSUPPORTED, "match_rt: 1 <= low_res <= 64");
    if (m.coarse == nullptr || m.coarse_win != low_res)
        return fail(h, CSM_E_INVALID, "match_rt: coarse map for this low_res not built");
    if (win_x < 0 || win_y < 0 || win_t < 0)
        return fail(h, CSM_E_INVALID, "match_rt: negative window");
    if (win_x > 8000 || win_y > 8000 || win_t > 32000)
        return fail(h, CSM_E_UNSUPPORTED, "match_rt: window too large");
    if (n <= 0 || n > kMaxBeams || angles == nullptr || ranges == nullptr)
        return fail(h, CSM_E_INVALID, "scan: need 1 <= n <= 4096 beams");

    int rc;
    const bool epilogue = h->epilogue_scale > 0.0;
    const bool refine_final = !epilogue && h->refine_on;      /* the final matcher on the found pose */
    if ((epilogue || refine_final) && (rc = ensure_alloc(h, std::vector<MapSlot*>{ &m }))) return rc;
    PlanView V;
    const int T = 2 * win_t + 1;
    if ((rc = layout_plan(h, 1, 0, 0, 0, n, V))) return rc;
    ScanSlot s = V.scan;
    s.max_range = *std::max_element(ranges, ranges + n);

    QueryPlan plan;
    plan.dq.resize(1);
    plan.theta_off.assign(1, 0);
    plan.inc_init.assign(1, 0ull);
    plan.scan_angles = angles; plan.scan_ranges = ranges;
    DevQuery& Q = plan.dq[0];
    if ((rc = fill_common(h, Q, m, s, score_thr, known_thr))) return rc;
    Q.sx = sensor_pose[0]; Q.sy = sensor_pose[1];
    Q.T = T;
    Q.winx = win_x; Q.winy = win_y;
    Q.proj_off = 0;
    Q.pst_t = Q.n; Q.pst_i = 1;
    plan.proj_total = (long long)Q.T * Q.n;
    plan.max_tn = Q.T * Q.n;
    plan.max_t = Q.T;
    Q.margin = h->fp_margin_scale * fp_margin(sensor_pose, m, s.max_range, (double)(std::max(win_x, win_y) + low_res) * m.res);
    Q.theta0 = sensor_pose[2]; Q.step_t = step_t; Q.tcenter = win_t;
    Q.stepx = step_x; Q.stepy = step_y;
    const int nbx = (2 * win_x) / low_res + 1;
    const int nby = (2 * win_y) / low_res + 1;
    const int nblocks = Q.T * nbx * nby;
    if ((rc = ensure(h, h->d_rtblocks, sizeof(RtBlock) * (size_t)nblocks))) return rc;
    phase_mark(h, "start");
    if ((rc = commit_plan(h, plan, V, false))) return rc;
    phase_mark(h, "k_setup");
    const DevQuery* dq = V.queries;
    proj_t* proj = static_cast<proj_t*>(h->d_proj.p);
    std::vector<proj_t> host_proj;
    if (exact) {
        /* exact rerun: ComputeScanIndices (scan_matcher_correlative.cpp:277-297) on the host, with the
         * reference's libm calls and operation order; the kernels take the indices as given */
        host_proj.resize((size_t)T * n);
        for (int t = 0; t < T; ++t) {
            const double theta = sensor_pose[2] + step_t * (t - win_t);
            for (int i = 0; i < n; ++i) {
                const double2 hit = host_hit_terms(theta, angles[i], ranges[i]);
                const double col = std::floor(((sensor_pose[0] + hit.x) - m.offx) / m.res);
                const double row = std::floor(((sensor_pose[1] + hit.y) - m.offy) / m.res);
                proj_t p;
                p.x = (short)std::max(std::min(col, (double)kProjSat), -(double)kProjSat);
                p.y = (short)std::max(std::min(row, (double)kProjSat), -(double)kProjSat);
                host_proj[(size_t)t * n + i] = p;
            }
        }
        CSM_CUDA(cudaMemcpyAsync(proj, host_proj.data(), sizeof(proj_t) * host_proj.size(), cudaMemcpyHostToDevice, h->stream));
        ++h->exact_reruns;
    }
    /* projection happens inside k_rt_blocks (one launch less on this latency-bound path) */
    k_rt_blocks<<<nblocks, 256, sizeof(long long) * low_res * low_res + sizeof(proj_t) * Q.n, h->stream>>>(
        dq, proj, static_cast<RtBlock*>(h->d_rtblocks.p), low_res, nbx, nby, V.qflags, exact ? 1 : 0);
    CSM_LAUNCH_CHECK();
    phase_mark(h, "k_rt_blocks");
    FinalArgs F;
    std::memset(&F, 0, sizeof(F));
    F.decode = 3;
    F.blocks = static_cast<const RtBlock*>(h->d_rtblocks.p);
    F.low_res = low_res; F.nbx = nbx; F.nby = nby;
    F.best_key = static_cast<unsigned long long*>(h->d_bestkey.p);
    F.mode = 0;
    F.qflags = V.qflags;
    F.nq = 1;
    k_finalize<<<1, 32, 0, h->stream>>>(dq, proj, F, static_cast<csm_result*>(h->d_results.p));
    CSM_LAUNCH_CHECK();
    phase_mark(h, "k_finalize");
    if (epilogue || refine_final) {
        RefineArgs R;
        std::memset(&R, 0, sizeof(R));
        R.results = static_cast<const csm_result*>(h->d_results.p);
        R.out = reinterpret_cast<csm_refined*>(static_cast<char*>(h->d_results.p) + refined_offset(1));
        if (epilogue) {
            /* Cost and ComputeCovariance at the decided pose, found or not (scan_matcher_correlative.cpp:203-219) */
            R.max_iterations = 0;
            R.always = 1;
            R.covariance_scale = h->epilogue_scale;
        } else {
            /* the caller's final matcher (ScanMatcherLinearSolver) on the pose found,
             * lidar_graph_slam_frontend.cpp:216-230 */
            R.max_iterations = h->refine.max_iterations;
            R.convergence_threshold = h->refine.convergence_threshold;
            R.lambda0 = h->refine.lambda;
            R.covariance_scale = h->refine.covariance_scale;
        }
        k_refine<<<1, kRefThreads, 0, h->stream>>>(dq, R);
        CSM_LAUNCH_CHECK();
        phase_mark(h, "k_refine");
    }
    const bool extra = epilogue || refine_final;
    if ((rc = enqueue_readback(h, 1, extra))) return rc;
    h->res_rec[(h->res_head + h->res_count - 1) % csm_context::kResultSlots].queries.clear();   /* no batch record: handled below */
    phase_mark(h, "readback");
    rc = finish_results(h, out, 1, extra ? &h->last_epilogue : nullptr);
    h->last_epilogue_set = rc == CSM_OK && extra;
    return rc;
}

int csm_match_rt(csm_handle h, int64_t map_id,
                 const double* angles, const double* ranges, int n,
                 const double sensor_pose[3], int low_res,
                 int win_x, int win_y, int win_t,
                 double step_x, double step_y, double step_t,
                 double score_thr, double known_thr, csm_result* out)
{
    int rc = match_rt_impl(h, map_id, angles, ranges, n, sensor_pose, low_res, win_x, win_y, win_t,
                           step_x, step_y, step_t, score_thr, known_thr, out, false);
    if (rc == CSM_OK && (out->flags & CSM_FLAG_FP_MARGIN) && h->exact_rerun) {
        /* a hit point within the FP guard band of a cell boundary: again, from host-evaluated indices */
        rc = match_rt_impl(h, map_id, angles, ranges, n, sensor_pose, low_res, win_x, win_y, win_t,
                           step_x, step_y, step_t, score_thr, known_thr, out, true);
        if (rc == CSM_OK) out->flags |= CSM_FLAG_EXACT;
    }
    return rc;
}

int csm_match_grid(csm_handle h, int64_t map_id,
                   const double* angles, const double* ranges, int n,
                   const double sensor_pose[3],
                   const double* dx, int ndx, const double* dy, int ndy,
                   const double* dt, int ndt,
                   double score_thr, double known_thr, csm_result* out)
{
    if (!h || !out || !sensor_pose || !dx || !dy || !dt) return CSM_E_INVALID;
    if (h->res_count != 0)
        return fail(h, CSM_E_INVALID, "match_grid: a loop batch is in flight on this handle");
    CSM_CUDA(cudaSetDevice(h->device));
    auto mi = h->maps.find(map_id);
    if (mi == h->maps.end())
        return fail(h, CSM_E_NOT_FOUND, "match_grid: unknown map id");
    MapSlot& m = mi->second;
    {
        const int wrc = wait_uploads(h, std::vector<MapSlot*>{ &m });
        if (wrc) return wrc;
    }
    if (ndx <= 0 || ndy <= 0 || ndt <= 0 || ndt > 65535 ||
        (unsigned long long)ndx * ndy * ndt >= kOrdMask - 1ull)
        return fail(h, CSM_E_UNSUPPORTED, "match_grid: window exceeds 2^26 candidates");
    if (n <= 0 || n > kMaxBeams || angles == nullptr || ranges == nullptr)
        return fail(h, CSM_E_INVALID, "scan: need 1 <= n <= 4096 beams");

    /* integer-shift path iff every dx[k] - dx[0] (dy likewise) is an integer
     * number of cells up to rounding noise */
    std::vector<int> offs(ndx + ndy);
    double dev = 0.0;
    for (int k = 0; k < ndx; ++k) {
        const double e = (dx[k] - dx[0]) / m.res;
        const double r = std::nearbyint(e);
        dev = std::max(dev, std::fabs(e - r));
        offs[k] = (int)r;
    }
    for (int k = 0; k < ndy; ++k) {
        const double e = (dy[k] - dy[0]) / m.res;
        const double r = std::nearbyint(e);
        dev = std::max(dev, std::fabs(e - r));
        offs[ndx + k] = (int)r;
    }
    bool fast = dev < 1e-7;
    for (int v : offs)
        fast = fast && std::abs(v) <= 8192;

    int rc;
    PlanView V;
    const size_t ob = align16(sizeof(int) * offs.size());
    const size_t pb = sizeof(double) * (size_t)(ndx + ndy);
    if ((rc = layout_plan(h, 1, (size_t)ndt, 0, ob + pb, n, V))) return rc;
    ScanSlot s = V.scan;
    s.max_range = *std::max_element(ranges, ranges + n);

    QueryPlan plan;
    plan.dq.resize(1);
    plan.theta_off.assign(1, 0);
    plan.inc_init.assign(1, 0ull);
    plan.scan_angles = angles; plan.scan_ranges = ranges;
    DevQuery& Q = plan.dq[0];
    if ((rc = fill_common(h, Q, m, s, score_thr, known_thr))) return rc;
    Q.sx = sensor_pose[0] + dx[0];
    Q.sy = sensor_pose[1] + dy[0];
    Q.T = ndt;
    Q.proj_off = 0;
    Q.pst_t = Q.n; Q.pst_i = 1;
    plan.proj_total = (long long)ndt * Q.n;
    plan.max_tn = ndt * Q.n;
    plan.max_t = ndt;
    const double extent = std::fabs(dx[0]) + std::fabs(dx[ndx - 1]) + std::fabs(dy[0]) + std::fabs(dy[ndy - 1]);
    Q.margin = h->fp_margin_scale * fp_margin(sensor_pose, m, s.max_range, extent) + (fast ? 2.0 * dev : 0.0);
    for (int k = 0; k < ndt; ++k)
        plan.thetas.push_back(sensor_pose[2] + dt[k]);

    plan.extra.resize(ob + pb);
    std::memcpy(plan.extra.data(), offs.data(), sizeof(int) * offs.size());
    {
        double* pos = reinterpret_cast<double*>(plan.extra.data() + ob);
        for (int k = 0; k < ndx; ++k) pos[k] = sensor_pose[0] + dx[k];
        for (int k = 0; k < ndy; ++k) pos[ndx + k] = sensor_pose[1] + dy[k];
    }
    phase_mark(h, "start");
    if ((rc = commit_plan(h, plan, V, !fast))) return rc;
    if ((rc = launch_project(h, plan, V, !fast))) return rc;
    phase_mark(h, "k_setup+k_project");

    GridArgs G;
    G.mx = reinterpret_cast<const int*>(V.extra);
    G.my = G.mx + ndx;
    G.px = reinterpret_cast<const double*>(V.extra + ob);
    G.py = G.px + ndx;
    G.ndx = ndx; G.ndy = ndy; G.ndt = ndt;
    G.best = V.inc;                       /* zero-initialised by the plan */
    G.tiekey = V.tiekey;
    G.ord_mode = 0;
    const DevQuery* dq = V.queries;
    const proj_t* proj = static_cast<const proj_t*>(h->d_proj.p);
    dim3 grid(ndt, (ndy + 7) / 8);
    /* integer-shift path: TMA-staged shared-memory tiles when the window fits the tile */
    bool tma = false;
    CUtensorMap tmap;
    WtArgs WA;
    std::memset(&WA, 0, sizeof(WA));
    if (fast && h->window_mode != 1) {
        int dy_span = 0, dx_span = 0;
        bool monotone = true, unit = true;
        for (int k = 1; k < ndx; ++k) { monotone = monotone && offs[k] >= offs[k - 1]; unit = unit && offs[k] == offs[k - 1] + 1; }
        for (int k = 1; k < ndy; ++k) monotone = monotone && offs[ndx + k] >= offs[ndx + k - 1];
        /* warps per CTA (3 candidate rows each): as many as possible (one CTA per SM, so the
         * warps are all the latency hiding there is) with the fewest idle rows in the last CTA */
        int warps = 8;
        {
            int best_padded = 1 << 30;
            for (int w = 8; w <= kWtMaxWarps; ++w) {
                const int rows = w * kWtRows;
                const int padded = ((ndy + rows - 1) / rows) * rows;
                if (padded <= best_padded) { best_padded = padded; warps = w; }
            }
        }
        const int rows_per_cta = warps * kWtRows;
        WA.rows_per_cta = rows_per_cta;
        /* unit column steps and 32 k + r columns, r <= kWtRemMax, in one block of columns: k chunks per
         * lane, the r remainder columns on a few lanes of every warp */
        const int rem_chunks = ndx / 32, rem_cols = ndx % 32;
        const bool rem_variant = unit && rem_chunks >= 1 && rem_chunks <= kWtChunks && rem_cols >= 1 &&
                                 rem_cols <= kWtRemMax;
        const int cols_per_cta = rem_variant ? ndx : kWtColsPerCta;
        for (int b = 0; b < ndy; b += rows_per_cta)
            dy_span = std::max(dy_span, offs[ndx + std::min(ndy, b + rows_per_cta) - 1] - offs[ndx + b]);
        for (int b = 0; b < ndx; b += cols_per_cta)
            dx_span = std::max(dx_span, offs[std::min(ndx, b + cols_per_cta) - 1] - offs[b]);
        if (unit) dx_span = std::max(dx_span, std::min(ndx, cols_per_cta) - 1);
        WA.dy_span = dy_span; WA.dx_span = dx_span; WA.unit_dx = unit ? 1 : 0;
        /* lanes of a partly filled last chunk read up to 32 ceil(cols / 32) - 1 columns to the right */
        const int lane_reach = unit ? (rem_variant ? ndx : kWtColsPerCta) - 1 : dx_span;
        /* wide: the tiles land as 32-bit words and are scored in place (window_mode 3 = the u16 tiles that
         * a pass per tile widens, the round-1 kernel) */
        bool wide = h->window_mode != 3 && (m.cols % 8) == 0 && encode_tiled_fn() != nullptr;
        int max_h = kWtTileRows - 1 - dy_span;
        int max_w = (wide ? kWwPitch - 4 : kWtPitch - 8) - lane_reach;
        if (wide && (max_w < 0 || !monotone || max_h < 0))
            wide = false, max_w = kWtPitch - 8 - lane_reach;
        if (wide) {
            if ((rc = ensure_wide(h, m))) return rc;
            tma = make_map_tensor_wide(m, &tmap);
        } else
            tma = monotone && max_h >= 0 && max_w >= 0 && make_map_tensor(m, &tmap);
        if (tma) {
            if ((rc = ensure(h, h->d_wtgroups, sizeof(WtGroup) * (size_t)ndt * Q.n + sizeof(int) * (size_t)ndt))) return rc;
            WtGroup* groups = static_cast<WtGroup*>(h->d_wtgroups.p);
            int* gcount = reinterpret_cast<int*>(groups + (size_t)ndt * Q.n);
            WA.groups = groups; WA.gcount = gcount;
            k_window_groups<<<(ndt + 127) / 128, 128, 0, h->stream>>>(dq, proj, groups, gcount, max_h, max_w);
            CSM_LAUNCH_CHECK();
            const size_t smem = wide ? wt_smem_bytes_wide(Q.n) : wt_smem_bytes(Q.n);
            dim3 wgrid(ndt, (ndy + rows_per_cta - 1) / rows_per_cta, (ndx + cols_per_cta - 1) / cols_per_cta);
            phase_mark(h, "k_window_groups");
#define CSM_WT_LAUNCH(UNIT, CHUNKS, REM)                                                                        \
            if (wide) k_window_tma<UNIT, CHUNKS, REM, true><<<wgrid, warps * 32, smem, h->stream>>>(tmap, dq, proj, G, WA); \
            else k_window_tma<UNIT, CHUNKS, REM, false><<<wgrid, warps * 32, smem, h->stream>>>(tmap, dq, proj, G, WA);
            if (rem_variant) {
                switch (rem_chunks) {
                case 1: CSM_WT_LAUNCH(true, 1, true) break;
                case 2: CSM_WT_LAUNCH(true, 2, true) break;
                case 3: CSM_WT_LAUNCH(true, 3, true) break;
                case 4: CSM_WT_LAUNCH(true, 4, true) break;
                case 5: CSM_WT_LAUNCH(true, 5, true) break;
                default: CSM_WT_LAUNCH(true, 6, true) break;
                }
            } else if (unit) {
                CSM_WT_LAUNCH(true, kWtChunks, false)
            } else {
                CSM_WT_LAUNCH(false, kWtChunks, false)
            }
#undef CSM_WT_LAUNCH
            CSM_LAUNCH_CHECK();
            phase_mark(h, "k_window_tma");
        } else if (h->window_mode == 2 || h->window_mode == 3) {
            return fail(h, CSM_E_UNSUPPORTED, "match_grid: the TMA window kernel cannot take this window / map");
        }
    }
    if (tma) {
        /* launched above */
    } else if (fast) {
        k_grid_window<<<grid, 256, sizeof(proj_t) * Q.n, h->stream>>>(dq, proj, G);
    } else {
        k_grid_general<<<grid, 256, sizeof(double2) * Q.n, h->stream>>>(
            dq, static_cast<const double2*>(h->d_rcs.p), G, V.qflags);
    }
    CSM_LAUNCH_CHECK();
    FinalArgs F;
    std::memset(&F, 0, sizeof(F));
    F.decode = 2;
    F.G = G;
    F.mx = G.mx; F.my = G.my; F.px = G.px; F.py = G.py;
    F.rcs = fast ? nullptr : static_cast<const double2*>(h->d_rcs.p);
    F.best_key = static_cast<unsigned long long*>(h->d_bestkey.p);
    F.mode = fast ? 1 : 2;
    F.qflags = V.qflags;
    F.nq = 1;
    phase_mark(h, "k_grid_window/general");
    k_finalize<<<1, 32, 0, h->stream>>>(dq, proj, F, static_cast<csm_result*>(h->d_results.p));
    CSM_LAUNCH_CHECK();
    phase_mark(h, "k_finalize");
    if ((rc = enqueue_readback(h, 1))) return rc;
    h->res_rec[(h->res_head + h->res_count - 1) % csm_context::kResultSlots].queries.clear();
    if ((rc = finish_results(h, out, 1))) return rc;
    if ((out->flags & CSM_FLAG_FP_MARGIN) && h->exact_rerun) {
        /* a hit point within the FP guard band of a cell boundary: the whole window again with the
         * reference's per-candidate arithmetic from host-evaluated hit terms */
        std::vector<double> thetas(ndt), px(ndx), py(ndy);
        for (int k = 0; k < ndt; ++k) thetas[k] = sensor_pose[2] + dt[k];
        for (int k = 0; k < ndx; ++k) px[k] = sensor_pose[0] + dx[k];
        for (int k = 0; k < ndy; ++k) py[k] = sensor_pose[1] + dy[k];
        csm_result e;
        if ((rc = exact_lattice_search(h, m, angles, ranges, n, thetas, px, py, 0, score_thr, known_thr, &e))) return rc;
        e.flags = (e.flags & CSM_FLAG_KEY_TIE) | CSM_FLAG_EXACT;
        *out = e;
    }
    return CSM_OK;
}

} /* extern "C" */
