// libnvsm_b200.so — host side of the B200-native NVSM/LSE training step and its C ABI
// (include/nvsm_b200.h). One nvsm_model owns the parameter tables, the optimiser state and
// a fixed per-step workspace in HBM; every step is a short, allocation-free sequence of
// kernel launches on one stream.
#include "../../include/nvsm_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tcgen05.cuh"
#include "gemm_tcgen05_2cta.cuh"
#include "kernels.cuh"
#include "microbench.cuh"
#include "ops.cuh"
#include "nccl_dyn.h"
#include "peer_allreduce.cuh"
#include "pull_update.cuh"
#include "sampler.cuh"
#include "score_ring.cuh"
#include "similarity.cuh"

using namespace nvsm;

namespace {

thread_local std::string g_error;

int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return 1;
}

#define CU(expr)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail("%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define TRY(expr)                 \
    do {                          \
        int rc_ = (expr);         \
        if (rc_ != 0) return rc_; \
    } while (0)

enum Phase {
    PH_H2D = 0, PH_GATHER, PH_GEMM_FWD, PH_BN_STATS, PH_SCORE, PH_BN_BWD, PH_GEMM_GT, PH_GEMM_GP,
    PH_UPD_ENTITIES, PH_UPD_WORDS, PH_UPD_TRANSFORM, PH_ALLREDUCE, PH_BUCKETS, PH_COUNT
};
const char* kPhaseNames[PH_COUNT] = {
    "h2d", "gather_mean", "gemm_fwd", "bn_stats", "score_loss_bwd", "bn_backward", "gemm_grad_transform",
    "gemm_grad_phrase", "update_entities", "update_words", "update_transform", "allreduce", "bucket_build"};

struct BatchSlot {
    idx_t* features = nullptr;   // [maxB*n]
    float* fweights = nullptr;   // [maxB*n]
    idx_t* ids = nullptr;        // [maxB*R]
    idx_t* labels = nullptr;     // [maxB] positive ids (device sampler input)
    float* weights = nullptr;    // [maxB]
    long B = 0;
    cudaEvent_t ready = nullptr;     // H2D done
    cudaEvent_t consumed = nullptr;  // last step reading it has been enqueued and finished
    bool ever_consumed = false;
    bool in_use = false;  // a forward result reads it and no `consumed` event covers that yet
    long fw_ones = 0, w_ones = 0;   // leading elements of fweights / weights known to hold 1.0f (uniform weights: no H2D)
};

struct TableOpt {  // optimiser state of one embedding table
    float* m = nullptr;    // Adam first moment [N, dim]
    float* v = nullptr;    // Adam second moment: [N] (sparse / dense-update) or [N, dim] (full)
    float* acc = nullptr;  // Adagrad [N]
    float* agg = nullptr;  // full Adam: scatter-added gradient of the running step [N, dim], kept zero between steps
    unsigned long t = 1;
};

}  // namespace

struct nvsm_model {
    nvsm_config cfg;
    int device = 0, num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;      // reference buckets of the pull update are built here, under the forward pass
    cudaEvent_t buckets_ready = nullptr, buckets_consumed = nullptr;
    // fused steps: the entity-table update only needs the forward pass, so it runs on the auxiliary stream under
    // batch-norm backward and the two backward GEMMs of the main stream (start_entity_update)
    cudaEvent_t score_done = nullptr, entity_done = nullptr, word_rows_done = nullptr;
    bool entity_async = false, entity_async_running = false;
    float* rowtmp_e = nullptr;   // entity-side row scratch (the word side may run concurrently)
    bool buckets_in_flight = false, buckets_ever_consumed = false;
    bool own_stream = false;
    long V, D;
    int dw, dd, n, z, R;
    long maxB;

    // parameters
    float *W = nullptr, *E = nullptr, *T = nullptr, *b = nullptr;
    float *P_lo = nullptr, *Gp_lo = nullptr, *Tt_lo = nullptr, *Tr_lo = nullptr;  // x - rn_tf32(x) operands (3xTF32)
    float* Tr = nullptr;  // T rounded to tf32 [dw, dd]: B operand of the grad_phrase tensor-core GEMM
    float* Tt = nullptr;  // T transposed [dd, dw]: K-major B operand of the forward tensor-core GEMM
    TableOpt optW, optE;
    float *T_a = nullptr, *b_a = nullptr, *T_v = nullptr, *b_v = nullptr;  // transform acc/m, v
    unsigned long t_transform = 1;

    // batches
    std::vector<BatchSlot> slots;  // cfg.num_batch_slots staged + 2 live (host-fed) slots
    int next_live = 0;
    int live_slots = 3;   // host-fed batches in flight. With 2 the upload of batch k+2 (H2D, id validation, device sampling) had to
                          // wait for step k to release its slot and then ran under step k+1's gather / forward GEMM (e2e timeline
                          // r2B: gather 75 -> 83 us, GEMM 76 -> 81 us); with 3 it starts as soon as the host enqueues it, under
                          // step k's backward / update: C2 e2e 0.619 -> 0.611 ms, C3 1.043 -> 1.031 ms (NVSM_LIVE_SLOTS=2|3|4)
    BatchSlot* cur = nullptr;      // batch of the running forward result
    long B = 0;                    // local instances of the running step
    long Bglobal = 0;

    // per-step workspace
    float *P = nullptr, *Z = nullptr, *Gp = nullptr, *gP = nullptr;
    float* Y = nullptr;  // post-activation projections, kept only for the pull-style (full Adam) update
    float *probs = nullptr, *mult = nullptr, *rowtmp = nullptr;
    float *mean = nullptr, *invstd = nullptr, *mean_dy = nullptr, *mean_dyx = nullptr;
    float *bn_scale = nullptr, *bn_shift = nullptr;  // invstd, bias - mean * invstd (ring score kernel)
    bool score_shifted = false;                       // backward column sums carry the bias term
    float* stat_part = nullptr;  // [2 * num_sms][2*dd] per-block partial column statistics
    double* dsums = nullptr;   // [2*dd fwd sums][dd var sums][2*dd bwd col sums][1 loss]
    float *gT = nullptr, *gb = nullptr, *gT_part = nullptr;
    int gt_splits = 1;
    int gt_nparts = 0;          // split-K partials of the running step's grad_transform
    bool gt_reduced = true;     // gT holds their sum (single GPU: summed inside transform_update_kernel or on demand)
    // pull-style full Adam: per-step reference buckets (counting sort by row)
    bool pull = false;          // reference buckets are built every step (full Adam, SGD, Adagrad)
    float* wcoef = nullptr;     // word-side Adagrad coefficients fw / sqrt(mean acc + eps) [maxB * n]
    int *e_counts = nullptr, *e_offsets = nullptr, *e_refs = nullptr;
    int *w_counts = nullptr, *w_offsets = nullptr, *w_refs = nullptr;
    int* scan_tmp = nullptr;  // block totals of the two-level scan
    // rows with more than kHeavyRefs references: work list + partial sums of pull_heavy_kernel, per table
    // (the two table updates may run concurrently on different streams)
    HeavyWork heavy_e{}, heavy_w{};
    HeavyWork *heavy_e_dev = nullptr, *heavy_w_dev = nullptr;   // device copies of the descriptors (row kernels)
    bool no_heavy = false;      // NVSM_NO_HEAVY=1: one warp per row whatever its reference count (A/B measurements)
    int ldP = 0;          // row stride of P (and Tt): d_w rounded up to 32 floats on the tensor-core path so
                          // that every 128-byte TMA box row is 128-byte aligned (d_w = 300 -> 320)
    bool use_tc = false;  // projection GEMMs on tcgen05 (gemm_mode != FP32 and shapes allow)
    bool t_copies_stale = true;   // Tt / Tr (+ lo) do not match T: refreshed by transform_update_kernel or the next forward
    float* scratch = nullptr;  // inspection buffer max(B*R*dd, ...) allocated on demand
    size_t scratch_bytes = 0;
    // loss read-back ring: every forward ends with an async D2H of its loss sum
    static constexpr int kCostRing = 16;
    double* loss_host = nullptr;  // pinned [kCostRing]
    cudaEvent_t loss_ev[kCostRing] = {nullptr};
    long loss_B[kCostRing] = {0};
    long forward_count = 0;
    bool have_forward = false, have_gradients = false;
    int* id_flags = nullptr;      // mapped pinned [2]: a word id / an entity id of an uploaded batch was out of range

    // instrumentation
    long launches = 0;
    bool profiling = false;   // serial phase timing: the stream overlaps are switched off so that phases add up
    bool timeline = false;    // overlaps kept: every phase reports (start, end) against one origin event, whatever its stream
    cudaEvent_t tl_origin = nullptr;
    struct Interval { int phase; float start_ms, end_ms; };
    std::vector<Interval> intervals;
    struct Ev { int phase; cudaEvent_t a, b; };
    std::vector<Ev> pending;
    std::vector<cudaEvent_t> ev_pool;
    double phase_ms[PH_COUNT] = {0};
    int open_phase = -1;
    cudaEvent_t open_ev = nullptr;

    // device sampler (sampler.cuh): engine state double-buffered on the device
    unsigned int* rng_dev = nullptr;   // [2]
    int rng_cur = 0;
    bool rng_seeded = false;
    int *smp_counts = nullptr, *smp_offsets = nullptr, *smp_scan = nullptr, *smp_error = nullptr;
    int2* smp_rej_items = nullptr;     // rare-rejection path of the sampler (sampler.cuh: RejList)
    int* smp_rej_count = nullptr;      // [2], by call parity
    unsigned long smp_calls = 0;
    double* smp_cdf = nullptr;         // [D] cumulative distribution of the negatives (null: uniform, the reference's)
    long smp_capacity = 0;             // candidate chunks allocated

    // L2 Normalizer (cpp/cuda_utils.cu:3-141): per-n-gram norms of P; per-reference |E_d| and normalised score;
    // effective multipliers mult / |E_d| and the per-row self coefficient of the entity gradient
    bool l2_phrase = false, l2_entity = false;
    float *p_norms = nullptr, *enorm = nullptr, *escore = nullptr, *mult_eff = nullptr, *kself = nullptr;
    bool entity_prep_done = false;

    // RepresentationSimilarity objective (similarity.cuh): pairs over the entity or the word table
    bool has_text = true, has_pair = false, pair_entities = false;
    float text_scale = 1.0f, pair_scale = 1.0f;   // w_k / sum_k w_k of the mixtures
    long maxN = 0, pair_N = 0;
    idx_t* pair_ids = nullptr;
    float *pair_w = nullptr, *pair_probs = nullptr, *pair_mult = nullptr, *pair_G = nullptr, *pair_msq = nullptr;
    double* pair_loss = nullptr;        // device [1]
    double* pair_loss_host = nullptr;   // pinned [1]
    bool have_pair_forward = false;

    // multi-GPU
    NcclComm comm = nullptr;
    int nranks = 1, rank = 0;
    // small reductions over NVLink peer memory (peer_allreduce.cuh); NCCL is the fallback and carries grad_transform
    bool peer_ready = false;
    PeerXchg peer{};
    PeerXchg* peer_dev = nullptr;                // device copy (kernels that take it by pointer)
    unsigned int* xchg_counter = nullptr;        // arrival counter of the score kernels' last-block exchange
    bool no_fused_xchg = false;
    double* peer_inbox = nullptr;                // this rank's inbox (exported with CUDA IPC)
    unsigned long long* peer_flags = nullptr;
    void* peer_mapped[2 * kPeerMaxRanks] = {nullptr};   // IPC mappings to close
    unsigned long long peer_epoch[kPeerKinds] = {0, 0, 0, 0, 0};
    bool gt_push_side = false;                   // NVSM_GT_PUSH_SIDE=1: the push kernel on the communication stream (measured slower)
    bool gt_xchg_pending = false;                // grad_transform went through gt_reduce_push_kernel: update_transform sums the inbox
    int gt_xchg_chunk = 0;
    int* peer_error = nullptr;
    cudaStream_t comm_stream = nullptr;          // grad_transform all-reduce, under grad_phrase and the word update
    cudaStream_t gt_stream = nullptr;            // fused steps: grad_transform GEMM (+ all-reduce) under the word update
    cudaEvent_t dx_ready = nullptr;
    bool gt_side = false, in_fused_step = false, no_fused_reduce = false, no_fused_gt = false, skip_gt_reduce = false;
    // experiment knobs of the per-step path (environment, read ONCE in nvsm_create; -1 / 0 = not set)
    struct Knobs {
        int tc_2cta = -1, tc_stages = 0, tc_kb = 0, score_w = 0, score_s = 0, stats_bps = 0, sgd_sparse = -1;
        bool no_ring = false, no_fused_stats = false, no_overlap = false, sampler_scan = false;
    } knobs;
    int buckets_at = 0;
    int pdl = 0;               // programmatic dependent launches along the main-stream kernel chain: bit mask over the
                               // kPdl* launch sites (NVSM_PDL)
    cudaEvent_t build_gate = nullptr;
    cudaEvent_t gt_ready = nullptr, gt_reduced_ev = nullptr;
    bool gt_allreduce_pending = false;
    // NVSM_SPARSE_ALLGATHER: every rank applies the table updates of ALL rows (exact single-GPU trajectory).
    int sparse_mode = NVSM_SPARSE_LOCAL;
    BatchSlot ag_slot;                  // gathered features / weights / ids of the global batch
    float *ag_mult = nullptr, *ag_act = nullptr, *ag_gP = nullptr, *ag_rowtmp = nullptr;
    float* ag_escore = nullptr;
    int *ag_e_refs = nullptr, *ag_w_refs = nullptr;

    double* fwd_sums() { return dsums; }
    double* var_sums() { return dsums + 2 * dd; }
    double* bwd_sums() { return dsums + 3 * dd; }
    double* loss_acc() { return dsums + 5 * dd; }
};

namespace {

// ------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------
template <typename T>
int dev_alloc(T** p, size_t count, bool zero = true) {
    CU(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    if (zero) {
        // cudaMemset runs on the legacy default stream and is asynchronous; the model's streams are non-blocking, so
        // without this wait a write enqueued on them right after the allocation (nvsm_sampler_seed -> rng_dev) can be
        // overtaken by the zero fill.
        CU(cudaMemset(*p, 0, std::max<size_t>(count, 1) * sizeof(T)));
        CU(cudaStreamSynchronize(cudaStreamLegacy));
    }
    return 0;
}

cudaEvent_t get_event(nvsm_model* m) {
    if (!m->ev_pool.empty()) {
        cudaEvent_t e = m->ev_pool.back();
        m->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void phase_begin(nvsm_model* m, int phase) {
    if (!m->profiling && !m->timeline) return;
    m->open_phase = phase;
    m->open_ev = get_event(m);
    cudaEventRecord(m->open_ev, m->stream);
}

void phase_end(nvsm_model* m) {
    if ((!m->profiling && !m->timeline) || m->open_phase < 0) return;
    cudaEvent_t e = get_event(m);
    cudaEventRecord(e, m->stream);
    m->pending.push_back({m->open_phase, m->open_ev, e});
    m->open_phase = -1;
}

int collect_phases(nvsm_model* m) {
    if (m->pending.empty()) return 0;
    CU(cudaStreamSynchronize(m->stream));
    if (m->aux_stream) CU(cudaStreamSynchronize(m->aux_stream));
    if (m->gt_stream) CU(cudaStreamSynchronize(m->gt_stream));
    if (m->copy_stream) CU(cudaStreamSynchronize(m->copy_stream));
    for (auto& ev : m->pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev.a, ev.b);
        m->phase_ms[ev.phase] += ms;
        if (m->timeline && m->tl_origin) {
            float t0 = 0.f;
            if (cudaEventElapsedTime(&t0, m->tl_origin, ev.a) == cudaSuccess) m->intervals.push_back({ev.phase, t0, t0 + ms});
            else (void)cudaGetLastError();
        }
        m->ev_pool.push_back(ev.a);
        m->ev_pool.push_back(ev.b);
    }
    m->pending.clear();
    return 0;
}

constexpr long kPullMinRows = 8192;   // pull-style SGD / Adagrad below this many table rows: atomics win

int grid_for(const nvsm_model* m, long work_items, int items_per_block, int blocks_per_sm) {
    long need = (work_items + items_per_block - 1) / items_per_block;
    long cap = (long)m->num_sms * blocks_per_sm;
    return (int)std::max<long>(1, std::min(need, cap));
}

#define LAUNCH(m, kernel, grid, block, smem, ...)                                   \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), (m)->stream>>>(__VA_ARGS__);              \
        (m)->launches++;                                                            \
        cudaError_t le_ = cudaPeekAtLastError();                                    \
        if (le_ != cudaSuccess)                                                     \
            return fail("launch %s: %s (%s:%d)", #kernel, cudaGetErrorString(le_),  \
                        __FILE__, __LINE__);                                        \
    } while (0)

enum { kPdlGemmFwd = 0, kPdlStats = 1, kPdlScore = 2, kPdlBnBwd = 3, kPdlGemmGt = 4, kPdlGemmGp = 5, kPdlNone = 30 };
// Same, as a programmatic dependent of the kernel in front of it in the stream (m->pdl; see common.cuh:pdl_wait). Only for
// kernels that execute pdl_wait() before their first dependent global access.
#define LAUNCH_PDL(m, bit, kernel, grid, block, smem, ...)                                            \
    do {                                                                                           \
        if (!((m)->pdl >> (bit) & 1)) {                                                                      \
            kernel<<<(grid), (block), (smem), (m)->stream>>>(__VA_ARGS__);                         \
        } else {                                                                                   \
            cudaLaunchConfig_t cfg_ = {};                                                          \
            cfg_.gridDim = dim3(grid); cfg_.blockDim = dim3(block);                                \
            cfg_.dynamicSmemBytes = (smem); cfg_.stream = (m)->stream;                             \
            cudaLaunchAttribute at_[1];                                                            \
            at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                        \
            at_[0].val.programmaticStreamSerializationAllowed = 1;                                 \
            cfg_.attrs = at_; cfg_.numAttrs = 1;                                                   \
            cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                                        \
        }                                                                                          \
        (m)->launches++;                                                                           \
        cudaError_t le_ = cudaPeekAtLastError();                                                   \
        if (le_ != cudaSuccess)                                                                    \
            return fail("launch %s: %s (%s:%d)", #kernel, cudaGetErrorString(le_),                 \
                        __FILE__, __LINE__);                                                       \
    } while (0)

ActParams act_params(const nvsm_model* m, bool use_bn) {
    ActParams a;
    a.nonlinearity = m->cfg.nonlinearity;
    a.use_bn = use_bn ? 1 : 0;
    // func::clip(-1, 1): bounds one ulp outside (include/cuNVSM/cuda_utils.h:91-95)
    a.clip_min = std::nextafter(-1.0f, -1.0f - 1e-5f);
    a.clip_max = std::nextafter(1.0f, 1.0f + 1e-5f);
    a.mean = m->mean;
    a.invstd = m->invstd;
    a.bias = m->b;
    return a;
}

bool vec4_ok(int dim) { return dim % 4 == 0; }

// ------------------------------------------------------------------------------------
// GEMM dispatch (fp32 SIMT path)
// ------------------------------------------------------------------------------------
template <bool TA, bool TB>
int run_sgemm(nvsm_model* m, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb, float* C,
              int ldc, int splits, float alpha, const float* bias) {
    int kps = (K + splits - 1) / splits;
    kps = (kps + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
    dim3 grid((N + GEMM_BN - 1) / GEMM_BN, (M + GEMM_BM - 1) / GEMM_BM, (K + kps - 1) / kps);
    LAUNCH(m, (sgemm_kernel<TA, TB>), grid, 256, 0, M, N, K, A, lda, Bm, ldb, C, ldc, kps, alpha, bias);
    return 0;
}


// ------------------------------------------------------------------------------------
// GEMM dispatch (tcgen05 / TMA path)
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D fp32 row-major tensor [outer, inner] with a {box_inner, box_outer} box, 128B swizzle, zero OOB fill.
int make_tensor_map(CUtensorMap* map, const float* base, long inner, long outer, long row_stride_elems,
                    int box_inner, int box_outer, bool atom32 = false) {
    // K-major boxes: the swizzle span equals the box row (32 fp32 = 128 B, or 16 fp32 = 64 B at k-block 16)
    const CUtensorMapSwizzle swz = atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                          : (box_inner == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail("cuTensorMapEncodeTiled is unavailable in this driver");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride_elems * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) inner=%ld outer=%ld box=%dx%d", (int)r, inner, outer, box_inner, box_outer);
    return 0;
}

// 227 KB opt-in limit minus the kernel's static shared memory (barriers), with headroom.
constexpr uint32_t kTcMaxDynSmem = 232448u - 2048u - 8192u;   // (8 KB static: the fused column-statistics strips)

// Dynamic shared memory above 48 KB is an opt-in per function AND per device: called by nvsm_create for the device of
// every model (a process-wide "done once" flag would leave a second device without it).
int set_kernel_attributes() {
    const int mx = (int)kTcMaxDynSmem;
    CU(cudaFuncSetAttribute(tc::gemm_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    CU(cudaFuncSetAttribute(tc::gemm_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
#define NVSM_TC_ATTR(A_, B_, S_, K_) CU(cudaFuncSetAttribute(tc::gemm_tc_kernel<A_, B_, S_, K_>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))
    NVSM_TC_ATTR(false, false, false, 32); NVSM_TC_ATTR(false, false, true, 32); NVSM_TC_ATTR(true, true, false, 32); NVSM_TC_ATTR(true, true, true, 32);
    NVSM_TC_ATTR(false, false, false, 16); NVSM_TC_ATTR(false, false, true, 16); NVSM_TC_ATTR(true, true, false, 16); NVSM_TC_ATTR(true, true, true, 16);
#undef NVSM_TC_ATTR
#define NVSM_RING_ATTR(N_) \
    CU(cudaFuncSetAttribute(score_ring_kernel<N_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    CU(cudaFuncSetAttribute(score_ring_kernel<N_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    NVSM_RING_ATTR(1); NVSM_RING_ATTR(2); NVSM_RING_ATTR(3); NVSM_RING_ATTR(4); NVSM_RING_ATTR(8);
#undef NVSM_RING_ATTR
    return 0;
}

bool tc_shapes_ok(int dw, int dd) {
    // TMA needs 16-byte row strides; grad_transform's MN-major B tile needs dd % 32 == 0.
    return dw % 4 == 0 && dd % 32 == 0 && dd <= 512 && dw <= 512;
}

// C[M, N] = alpha * A . B (+ bias), reduction length K.
//   mn_major == false: A is [M, K] row-major, Bm is [N, K] row-major      (both K-major)
//   mn_major == true : A is [K, M] row-major, Bm is [K, N] row-major      (both MN-major)
// splits > 1 writes `splits` partial products to C + z * split_stride.
int run_gemm_tc(nvsm_model* m, bool mn_major, int M, int N, int K, const float* A, int lda, const float* Bm, int ldb,
                float* C, int ldc, int splits, long split_stride, float alpha, const float* bias,
                int* splits_out = nullptr, const float* A_lo = nullptr, const float* B_lo = nullptr,
                float* stat_part = nullptr, int* stat_rows_out = nullptr, int pdl_bit = kPdlNone) {
    // stat_part: fused column statistics of C (tc::Params::stat_part); *stat_rows_out = partial rows written, 0 when
    // this launch could not fuse them (the caller then runs col_stats4_kernel)
    const bool split3 = A_lo != nullptr && B_lo != nullptr;
    if (stat_rows_out) *stat_rows_out = 0;
    {
        // Two-SM (cta_group::2) kernel for the K-major, un-split GEMMs (forward, grad_phrase); see gemm_tcgen05_2cta.cuh.
        // Default: on when one 256-wide tile covers N (the forward projection: B traffic per SM halves, measured
        // 70.2 -> 62.3 us on C2); grad_phrase (N = 300 -> two 160-wide tiles) measured equal (56.2 vs 56.6 us) and
        // stays on the single-SM kernel. NVSM_TC_2CTA=0 / 1 forces it off / on for every eligible GEMM.
        const bool want = m->knobs.tc_2cta >= 0 ? m->knobs.tc_2cta != 0 : (N > 160 && N <= 256);
        if (want && !mn_major && splits <= 1 && m->num_sms >= 2) {
            tc::Params p;
            p.M = M; p.N = N; p.K = K;
            const int n_pad = (N + 15) / 16 * 16;
            p.n_tiles = (n_pad + 255) / 256;
            p.bn = ((n_pad + p.n_tiles - 1) / p.n_tiles + 15) / 16 * 16;   // <= 256, bn / 2 a multiple of 8
            p.m_tiles = (M + 255) / 256;
            p.splits = 1; p.kb = 32;
            p.kb_per_split = (K + tc::kBlockK - 1) / tc::kBlockK;
            p.stage_bytes = (tc::kATileBytes + (uint32_t)(p.bn / 2) * 128u) * (split3 ? 2u : 1u);
            p.stages = (int)std::min<uint32_t>(8u, (kTcMaxDynSmem - 1024u) / p.stage_bytes);
            if (m->knobs.tc_stages) p.stages = std::max(2, std::min(p.stages, m->knobs.tc_stages));
            if (p.stages < 2) return fail("tensor-core GEMM: tile does not fit in shared memory");
            p.tmem_cols = 512;   // whole TMEM: both CTAs of the pair must get base 0
            p.C = C; p.ldc = ldc; p.split_stride = 0; p.alpha = alpha; p.bias = bias;
            const bool fuse_stats = stat_part && p.n_tiles == 1 && !bias && N % 16 == 0 && N <= tc::kStatCols;
            p.stat_part = fuse_stats ? stat_part : nullptr;
            CUtensorMap tmA, tmB, tmAlo, tmBlo;
            TRY(make_tensor_map(&tmA, A, K, M, lda, tc::kBlockK, tc::kBlockM));
            TRY(make_tensor_map(&tmB, Bm, K, N, ldb, tc::kBlockK, p.bn / 2));
            TRY(make_tensor_map(&tmAlo, split3 ? A_lo : A, K, M, lda, tc::kBlockK, tc::kBlockM));
            TRY(make_tensor_map(&tmBlo, split3 ? B_lo : Bm, K, N, ldb, tc::kBlockK, p.bn / 2));
            const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
            const int num_tiles = p.m_tiles * p.n_tiles;
            const int grid = 2 * std::min(num_tiles, m->num_sms / 2);
            if (split3) LAUNCH_PDL(m, pdl_bit, (tc::gemm_tc2_kernel<true>), grid, tc::kThreads, smem, tmA, tmB, tmAlo, tmBlo, p);
            else LAUNCH_PDL(m, pdl_bit, (tc::gemm_tc2_kernel<false>), grid, tc::kThreads, smem, tmA, tmB, tmAlo, tmBlo, p);
            if (splits_out) *splits_out = 1;
            if (stat_rows_out && fuse_stats) *stat_rows_out = grid * 4;
            return 0;
        }
    }
    tc::Params p;
    p.M = M; p.N = N; p.K = K;
    const int unit = mn_major ? 32 : 16;
    const int n_pad = (N + unit - 1) / unit * unit;
    const int max_bn = 256;   // (3xTF32: two 96 KB stages measured faster than three 64 KB stages with bn = 128)
    p.n_tiles = (n_pad + max_bn - 1) / max_bn;
    p.bn = ((n_pad + p.n_tiles - 1) / p.n_tiles + unit - 1) / unit * unit;   // <= 256
    p.m_tiles = (M + tc::kBlockM - 1) / tc::kBlockM;
    // k-block per stage: 16 halves the stage (deeper ring, twice the TMA requests and barrier round trips); NVSM_TC_KB overrides
    int kb = 32;   // measured on C2 (3xTF32): k-block 16 = 4-deep ring 72 / 88 / 68 us (fwd, gT, gP) vs k-block 32 = 2-deep 70 / 71 / 57 us
    if (m->knobs.tc_kb == 16 || m->knobs.tc_kb == 32) kb = m->knobs.tc_kb;
    p.kb = kb;
    const int num_kb = (K + kb - 1) / kb;
    splits = std::max(1, std::min(splits, num_kb));
    p.kb_per_split = (num_kb + splits - 1) / splits;
    p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.stage_bytes = ((uint32_t)tc::kBlockM * kb * 4u + (uint32_t)p.bn * kb * 4u) * (split3 ? 2u : 1u);
    p.stages = (int)std::min<uint32_t>(8u, (kTcMaxDynSmem - 1024u) / p.stage_bytes);
    if (m->knobs.tc_stages) p.stages = std::max(2, std::min(p.stages, m->knobs.tc_stages));
    if (p.stages < 2) return fail("tensor-core GEMM: tile does not fit in shared memory");
    p.tmem_cols = 32;
    while ((int)p.tmem_cols < 2 * p.bn) p.tmem_cols <<= 1;
    p.C = C; p.ldc = ldc; p.split_stride = split_stride; p.alpha = alpha; p.bias = bias;
    const bool fuse_stats = stat_part && p.n_tiles == 1 && p.splits == 1 && !bias && N % 16 == 0 && N <= tc::kStatCols;
    p.stat_part = fuse_stats ? stat_part : nullptr;
    CUtensorMap tmA, tmB, tmAlo, tmBlo;
    if (!mn_major) {
        TRY(make_tensor_map(&tmA, A, K, M, lda, kb, tc::kBlockM));
        TRY(make_tensor_map(&tmB, Bm, K, N, ldb, kb, p.bn));
        TRY(make_tensor_map(&tmAlo, split3 ? A_lo : A, K, M, lda, kb, tc::kBlockM));
        TRY(make_tensor_map(&tmBlo, split3 ? B_lo : Bm, K, N, ldb, kb, p.bn));
    } else {
        TRY(make_tensor_map(&tmA, A, M, K, lda, 32, kb, true));
        TRY(make_tensor_map(&tmB, Bm, N, K, ldb, 32, kb, true));
        TRY(make_tensor_map(&tmAlo, split3 ? A_lo : A, M, K, lda, 32, kb, true));
        TRY(make_tensor_map(&tmBlo, split3 ? B_lo : Bm, N, K, ldb, 32, kb, true));
    }
    const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
    const int num_tiles = p.m_tiles * p.n_tiles * p.splits;
    const int grid = std::min(num_tiles, m->num_sms);
#define NVSM_TC_LAUNCH(A_, B_, S_, K_) LAUNCH_PDL(m, pdl_bit, (tc::gemm_tc_kernel<A_, B_, S_, K_>), grid, tc::kThreads, smem, tmA, tmB, tmAlo, tmBlo, p)
    if (kb == 32) {
        if (!mn_major && !split3) NVSM_TC_LAUNCH(false, false, false, 32);
        else if (!mn_major) NVSM_TC_LAUNCH(false, false, true, 32);
        else if (!split3) NVSM_TC_LAUNCH(true, true, false, 32);
        else NVSM_TC_LAUNCH(true, true, true, 32);
    } else {
        if (!mn_major && !split3) NVSM_TC_LAUNCH(false, false, false, 16);
        else if (!mn_major) NVSM_TC_LAUNCH(false, false, true, 16);
        else if (!split3) NVSM_TC_LAUNCH(true, true, false, 16);
        else NVSM_TC_LAUNCH(true, true, true, 16);
    }
#undef NVSM_TC_LAUNCH
    if (splits_out) *splits_out = p.splits;
    if (stat_rows_out && fuse_stats) *stat_rows_out = grid * 4;
    return 0;
}

// out[c][r] = rn_tf32(in[r][c]) (K-major B operand of the forward GEMM); copy[r][c] = rn_tf32(in[r][c])
// (K-major B operand of the grad_phrase GEMM).
__global__ void transpose_kernel(const float* __restrict__ in, int rows, int cols, float* __restrict__ out, int ld_out,
                                 float* __restrict__ copy, float* __restrict__ out_lo, float* __restrict__ copy_lo) {
    __shared__ float tile_lo[32][33];
    __shared__ float tile[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    for (int r = blockIdx.y * 32 + threadIdx.y; r < min(rows, (int)(blockIdx.y + 1) * 32); r += blockDim.y)
        if (c < cols) {
            const float raw = in[(long)r * cols + c];
            const float x = round_tf32(raw);
            tile[r - blockIdx.y * 32][threadIdx.x] = x;
            tile_lo[r - blockIdx.y * 32][threadIdx.x] = raw - x;
            copy[(long)r * cols + c] = x;
            if (copy_lo) copy_lo[(long)r * cols + c] = raw - x;
        }
    __syncthreads();
    const int r2 = blockIdx.y * 32 + threadIdx.x;
    for (int c2 = blockIdx.x * 32 + threadIdx.y; c2 < min(cols, (int)(blockIdx.x + 1) * 32); c2 += blockDim.y)
        if (r2 < rows) {
            out[(long)c2 * ld_out + r2] = tile[threadIdx.x][c2 - blockIdx.x * 32];
            if (out_lo) out_lo[(long)c2 * ld_out + r2] = tile_lo[threadIdx.x][c2 - blockIdx.x * 32];
        }
}

// kind >= 0 names one of the kPeerKinds per-step reduction sites that may use the NVLink peer exchange.
int allreduce(nvsm_model* m, void* buf, size_t count, bool is_double, int kind = -1) {
    if (m->nranks <= 1) return 0;
    phase_end(m);
    phase_begin(m, PH_ALLREDUCE);
    if (m->peer_ready && is_double && kind >= 0 && kind < kPeerKinds && (long)count <= m->peer.slot_doubles) {
        const unsigned long long epoch = ++m->peer_epoch[kind];
        LAUNCH(m, peer_allreduce_kernel, 1, 256, 0, m->peer, (double*)buf, (int)count, kind, epoch, m->peer_error);
        phase_end(m);
        return 0;
    }
    int rc = nccl_api().AllReduce(buf, buf, count, is_double ? kNcclFloat64 : kNcclFloat32, kNcclSum, m->comm,
                                  (void*)m->stream);
    phase_end(m);
    if (rc != 0) return fail("ncclAllReduce: %s", nccl_api().GetErrorString(rc));
    return 0;
}

int allgather_bytes(nvsm_model* m, const void* src, void* dst, size_t bytes_per_rank) {
    int rc = nccl_api().AllGather(src, dst, bytes_per_rank, kNcclInt8, m->comm, (void*)m->stream);
    if (rc != 0) return fail("ncclAllGather: %s", nccl_api().GetErrorString(rc));
    return 0;
}

// The two per-step reductions on the critical path run INSIDE compute kernels when the NVLink peer exchange is up
// (peer_allreduce.cuh: col_stats_reduce_finalize_kernel<true>, peer_sums_tail); NVSM_NO_FUSED_XCHG=1 restores the
// stand-alone one-block all-reduce launches (A/B measurements), and without peer memory they are ncclAllReduce calls.
bool fused_xchg(const nvsm_model* m) {
    return m->nranks > 1 && m->peer_ready && m->peer_dev && !m->no_fused_xchg && 2 * m->dd + 1 <= m->peer.slot_doubles &&
           (m->dd + 7) / 8 <= kPeerFlagStride;
}

bool exact_sparse(const nvsm_model* m) { return m->nranks > 1 && m->sparse_mode == NVSM_SPARSE_ALLGATHER; }

// ------------------------------------------------------------------------------------
// forward: Model::compute_cost on a device-resident batch
// ------------------------------------------------------------------------------------
template <int VEC, int NCH>
int launch_score(nvsm_model* m, const ScoreParams& sp) {
    const int grid = grid_for(m, sp.B, 8, 3);
    const size_t smem = (2 * (size_t)sp.dd + 1) * sizeof(float);
    LAUNCH(m, (score_kernel<VEC, NCH>), grid, 256, smem, sp);
    return 0;
}

template <int NCH>
int launch_score_ring(nvsm_model* m, const ScoreRingParams& q, int grid, size_t smem) {
    if (q.s.dd == NCH * 128) LAUNCH_PDL(m, kPdlScore, (score_ring_kernel<NCH, true>), grid, q.warps * 32, smem, q);
    else LAUNCH_PDL(m, kPdlScore, (score_ring_kernel<NCH, false>), grid, q.warps * 32, smem, q);
    return 0;
}

// Staged (cp.async) variant when the rows are 16-byte multiples; returns -1 when it does not apply.
// Policy: one stage per warp and as many warps per SM as shared memory allows (two 8-warp blocks for the NVSM
// shape) -- measured on C2: 16 warps/SM x 1 stage 92.9 us vs 8 warps/SM x 2-stage prefetch ring 124.6 us; the
// per-n-gram work is a dependent LDS -> FMA -> shuffle chain, so extra warps hide more than a deeper ring does.
// The >= 2-stage ring is kept for stages so large that fewer than 8 single-stage warps fit.
int try_score_ring(nvsm_model* m, const ScoreParams& sp) {
    const int dd = sp.dd, R = sp.R;
    if (!vec4_ok(dd) || R > 32 || dd > 1024 || m->knobs.no_ring) return -1;
    if (sp.enorm) return -1;   // entity normalisation lives in the register variant (score_kernel)
    const size_t stage_bytes = (size_t)(R + 1) * dd * 4;
    const size_t fixed = (size_t)(4 * dd + 2) * 4 + 8 * 4 * 8;
    const size_t budget = 227 * 1024 - 1024;
    int W = 8, S = 0;
    const bool ew = m->knobs.score_w > 0, es = m->knobs.score_s > 0;
    if (ew) W = std::max(1, std::min(8, m->knobs.score_w));
    if (!ew && !es) {
        // single stage: blocks of W warps, as many blocks per SM as fit; the block size that keeps the most warps
        // resident wins (C3, R = 17: 18 KB per stage -> one 8-warp block per SM = 12.5 % occupancy in ncu r2a, three
        // 4-warp blocks = 12 warps), ties go to the larger block
        int best_w = 0, best_warps = 0;
        for (int cand = 8; cand >= 2; cand /= 2) {
            const size_t blocks = std::min<size_t>(budget / (cand * stage_bytes + fixed + 1024), 2048 / (cand * 32));
            const int warps = (int)blocks * cand;
            if (warps >= 8 && warps > best_warps) { best_warps = warps; best_w = cand; }
        }
        if (best_w) { W = best_w; S = 1; } else W = 8;
    }
    if (S == 0) {
        for (; W >= 2; W /= 2) {
            S = (int)std::min<size_t>(4, (budget - fixed) / (W * stage_bytes));
            if (S >= 2 || ew) break;
        }
        if (es) S = std::min(S, std::max(1, m->knobs.score_s));
    }
    if (S < 1) return -1;
    ScoreRingParams q;
    q.s = sp; q.stages = S; q.warps = W; q.bn_scale = m->bn_scale; q.bn_shift = m->bn_shift;
    const size_t smem = (size_t)W * S * stage_bytes + (size_t)(4 * dd + 2) * 4 + (size_t)W * S * 8;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2048 / (W * 32), (227 * 1024) / (smem + 1024)));
    const int grid = (int)std::max<long>(1, std::min<long>((sp.B + W - 1) / W, (long)m->num_sms * per_sm));
    const int nch = (dd / 4 + 31) / 32;
    if (nch <= 1) return launch_score_ring<1>(m, q, grid, smem);
    if (nch <= 2) return launch_score_ring<2>(m, q, grid, smem);
    if (nch <= 3) return launch_score_ring<3>(m, q, grid, smem);
    if (nch <= 4) return launch_score_ring<4>(m, q, grid, smem);
    return launch_score_ring<8>(m, q, grid, smem);
}

int dispatch_score(nvsm_model* m, const ScoreParams& sp) {
    const int dd = sp.dd;
    m->score_shifted = false;
    {
        const int rc = try_score_ring(m, sp);
        if (rc == 0) { m->score_shifted = sp.act.use_bn != 0; return 0; }
        if (rc > 0) return rc;
    }
    if (vec4_ok(dd)) {
        const int nch = (dd / 4 + 31) / 32;
        if (nch <= 1) return launch_score<4, 1>(m, sp);
        if (nch <= 2) return launch_score<4, 2>(m, sp);
        if (nch <= 3) return launch_score<4, 3>(m, sp);
        if (nch <= 4) return launch_score<4, 4>(m, sp);
        if (nch <= 8) return launch_score<4, 8>(m, sp);
    } else {
        const int nch = (dd + 31) / 32;
        if (nch <= 1) return launch_score<1, 1>(m, sp);
        if (nch <= 4) return launch_score<1, 4>(m, sp);
        if (nch <= 16) return launch_score<1, 16>(m, sp);
    }
    return fail("entity_repr_size %d is not supported (max 1024 when a multiple of 4, else 512)", dd);
}

int start_bucket_build(nvsm_model* m, BatchSlot* s);

int forward(nvsm_model* m, BatchSlot* s) {
    const long B = s->B;
    if (B <= 0) return fail("empty batch");
    if (!m->has_text) return fail("this handle was created for a RepresentationSimilarity objective: use nvsm_similarity_compute_cost");
    m->cur = s;
    s->in_use = true;
    m->B = B;
    m->Bglobal = B * m->nranks;
    m->have_forward = false;
    m->have_gradients = false;
    CU(cudaStreamWaitEvent(m->stream, s->ready, 0));
    // Reference buckets of the pull update: built on the auxiliary stream under the forward pass. buckets_at (experiment
    // knob NVSM_BUCKETS_AT) = main-stream point behind which the build may start: 0 step start, 1 gather, 2 forward GEMM.
    auto build_at = [&](int stage) -> int {
        if (!(m->pull && !exact_sparse(m)) || m->buckets_at != stage) return 0;   // (gathered ids only exist at update time)
        if (stage > 0) {
            CU(cudaEventRecord(m->build_gate, m->stream));
            CU(cudaStreamWaitEvent(m->aux_stream, m->build_gate, 0));
        }
        return start_bucket_build(m, s);
    };
    TRY(build_at(0));
    const bool bn = m->cfg.batch_normalization != 0;
    const int dw = m->dw, dd = m->dd;

    // Batch-wide accumulators. On the tensor-core batch-norm path nothing needs a memset operation in the stream: the
    // forward sums are overwritten by col_stats_reduce_finalize_kernel, which also zeroes the backward sums + loss.
    const bool zero_in_finalize = m->cfg.batch_normalization != 0 && m->use_tc;
    if (!zero_in_finalize) CU(cudaMemsetAsync(m->dsums, 0, (5 * (size_t)dd + 1) * sizeof(double), m->stream));

    // (1) phrase representations: weighted mean of the word rows. (A cp.async-ring variant like
    // score_ring_kernel and a fully unrolled window were both measured slower: the gather has almost
    // no arithmetic per row, so plain loads at high occupancy win.)
    phase_begin(m, PH_GATHER);
    {
        const int grid = grid_for(m, B, 8, 8);
        const int nvec = dw / 4;
        if (vec4_ok(dw) && nvec <= 128 && (long)m->V * nvec < (1L << 32)) {
            // K float4 per lane, L active lanes per n-gram, LG = group width (power of two)
            const int K = (nvec + 31) / 32, L = (nvec + K - 1) / K;
            int LG = 1;
            while (LG < L) LG <<= 1;
            const int tf = m->use_tc ? 1 : 0;
#define NVSM_GATHER(KK, GG) LAUNCH(m, (gather_mean_lanes_kernel<KK, GG>), grid_for(m, B / (32 / GG) + 1, 8, 8), 256, 0, m->W, dw, \
                                   s->features, s->fweights, B, m->n, L, m->P, m->ldP, tf, m->P_lo)
            if (K == 1 && LG <= 4) NVSM_GATHER(1, 4);
            else if (K == 1 && LG == 8) NVSM_GATHER(1, 8);
            else if (K == 1 && LG == 16) NVSM_GATHER(1, 16);
            else if (K == 1) NVSM_GATHER(1, 32);
            else if (K == 2) NVSM_GATHER(2, 32);
            else if (K == 3) NVSM_GATHER(3, 32);
            else NVSM_GATHER(4, 32);
#undef NVSM_GATHER
        } else if (vec4_ok(dw))
            LAUNCH(m, gather_mean_kernel<4>, grid, 256, 0, m->W, dw, s->features, s->fweights, B, m->n, m->P, m->ldP, m->use_tc ? 1 : 0, m->P_lo);
        else
            LAUNCH(m, gather_mean_kernel<1>, grid, 256, 0, m->W, dw, s->features, s->fweights, B, m->n, m->P, m->ldP, m->use_tc ? 1 : 0, m->P_lo);
    }
    if (m->l2_phrase)   // Normalizer::forward on the phrase representations, in place (cpp/objective.cu:134-140)
        LAUNCH(m, row_l2_normalize_kernel, grid_for(m, B, 8, 8), 256, 0, m->P, m->P_lo, B, dw, m->ldP, m->use_tc ? 1 : 0, m->p_norms);
    phase_end(m);
    TRY(build_at(1));

    // (2) projection Z = P . T (+ b when batch-norm is off).
    int stat_rows = 0;   // > 0: the GEMM epilogue wrote that many partial rows of column statistics into stat_part
    phase_begin(m, PH_GEMM_FWD);
    if (m->use_tc) {
        if (m->t_copies_stale) {   // after initialize / set_tensor; update() keeps the copies current otherwise
            LAUNCH(m, transpose_kernel, dim3((dd + 31) / 32, (dw + 31) / 32), dim3(32, 8), 0, m->T, dw, dd, m->Tt, m->ldP, m->Tr, m->Tt_lo, m->Tr_lo);
            m->t_copies_stale = false;
        }
        // batch-norm: the column sums / sums of squares of Z come out of the GEMM's epilogue (no second pass over Z)
        float* const fuse = (bn && !m->knobs.no_fused_stats) ? m->stat_part : (float*)nullptr;
        TRY(run_gemm_tc(m, false, (int)B, dd, dw, m->P, m->ldP, m->Tt, m->ldP, m->Z, dd, 1, 0, 1.0f, bn ? nullptr : m->b, nullptr,
                        m->P_lo, m->Tt_lo, fuse, &stat_rows, kPdlGemmFwd));
    } else {
        TRY((run_sgemm<false, false>(m, (int)B, dd, dw, m->P, m->ldP, m->T, dd, m->Z, dd, 1, 1.0f, bn ? nullptr : m->b)));
    }
    phase_end(m);

    TRY(build_at(2));

    // (3) batch statistics over the (global) batch.
    if (bn && m->use_tc) {
        // One pass over the L2-resident Z for sums and sums of squares (fp32 over ~100-row
        // slabs, double across slabs); one all-reduce, then mean / invstd.
        phase_begin(m, PH_BN_STATS);
        {
            const int nvec = dd / 4, tpr = std::min(nvec, 256), rpp = 256 / tpr;
            int bps = 4;   // blocks per SM (NVSM_STATS_BPS: experiment knob)
            if (m->knobs.stats_bps) bps = std::max(1, std::min(8, m->knobs.stats_bps));
            int nblk = stat_rows;
            if (nblk == 0) {
                nblk = grid_for(m, B, 16, bps);
                LAUNCH(m, col_stats4_kernel, nblk, 256, (size_t)rpp * 2 * dd * sizeof(float), m->Z, B, dd, m->stat_part);
            }
            if (m->nranks <= 1) {
                LAUNCH_PDL(m, kPdlStats, col_stats_reduce_finalize_kernel<false>, (dd + 7) / 8, 1024, 0, m->stat_part, nblk, dd, (double)m->Bglobal,
                       1e-4 /* cpp/objective.cu:114 */, m->fwd_sums(), m->mean, m->invstd, m->b, m->bn_scale, m->bn_shift,
                       (const PeerXchg*)nullptr, 0ull, (int*)nullptr, m->bwd_sums(), 2 * dd + 1);
            } else if (fused_xchg(m)) {
                // partial rows -> local sums -> NVLink exchange -> global mean / invstd in ONE launch (per-block flags)
                const unsigned long long epoch = ++m->peer_epoch[0];
                LAUNCH_PDL(m, kPdlStats, col_stats_reduce_finalize_kernel<true>, (dd + 7) / 8, 1024, 0, m->stat_part, nblk, dd, (double)m->Bglobal,
                       1e-4 /* cpp/objective.cu:114 */, m->fwd_sums(), m->mean, m->invstd, m->b, m->bn_scale, m->bn_shift,
                       (const PeerXchg*)m->peer_dev, epoch, m->peer_error, m->bwd_sums(), 2 * dd + 1);
            } else {
                CU(cudaMemsetAsync(m->bwd_sums(), 0, (2 * (size_t)dd + 1) * sizeof(double), m->stream));
                LAUNCH(m, col_stats_reduce_kernel, (2 * dd + 31) / 32, 256, 0, m->stat_part, nblk, 2 * dd, m->fwd_sums());
            }
        }
        phase_end(m);
        if (m->nranks > 1 && !fused_xchg(m)) {
            TRY(allreduce(m, m->fwd_sums(), 2 * (size_t)dd, true, 0));
            phase_begin(m, PH_BN_STATS);
            LAUNCH(m, bn_finalize_kernel, (dd + 127) / 128, 128, 0, m->fwd_sums(), dd, (double)m->Bglobal,
                   1e-4 /* cpp/objective.cu:114 */, m->mean, m->invstd, m->b, m->bn_scale, m->bn_shift);
            phase_end(m);
        }
    } else if (bn) {
        phase_begin(m, PH_BN_STATS);
        const int grid = grid_for(m, B, 64, 4);
        LAUNCH(m, col_stats_kernel, grid, 256, 0, m->Z, B, dd, m->fwd_sums());
        phase_end(m);
        TRY(allreduce(m, m->fwd_sums(), dd, true, 1));
        phase_begin(m, PH_BN_STATS);
        LAUNCH(m, col_var_kernel, grid, 256, 0, m->Z, B, dd, m->fwd_sums(), (double)m->Bglobal, m->var_sums());
        phase_end(m);
        TRY(allreduce(m, m->var_sums(), dd, true, 2));
        phase_begin(m, PH_BN_STATS);
        LAUNCH(m, bn_finalize2_kernel, (dd + 127) / 128, 128, 0, m->fwd_sums(), m->var_sums(), dd,
               (double)m->Bglobal, 1e-4 /* cpp/objective.cu:114 */, m->mean, m->invstd);
        phase_end(m);
    }

    if (bn && !m->use_tc) LAUNCH(m, bn_affine_kernel, (dd + 127) / 128, 128, 0, m->mean, m->invstd, m->b, dd, m->bn_scale, m->bn_shift);

    // (4) scores, loss, multipliers and d cost / d pre-activation in one pass.
    phase_begin(m, PH_SCORE);
    {
        ScoreParams sp;
        sp.Z = m->Z; sp.E = m->E; sp.ids = s->ids; sp.inst_w = s->weights;
        sp.B = B; sp.R = m->R; sp.dd = dd;
        const bool rebalance = !m->cfg.bias_negative_samples && m->z > 1;
        sp.w_scale = rebalance ? (float)(((double)(float)m->z + 1.0) / (2.0 * (double)(float)m->z)) : 1.0f;
        sp.pos_scale = rebalance ? (float)m->z : 1.0f;
        const float ef = m->cfg.clip_sigmoid ? 1e-7f : 0.0f;
        const float eb = m->cfg.clip_sigmoid ? 1e-6f : 0.0f;
        // float thresholds equivalent to the reference's float-vs-double comparisons
        auto fceil = [](double d) { float f = (float)d; if ((double)f < d) f = std::nextafter(f, INFINITY); return f; };
        auto ffloor = [](double d) { float f = (float)d; if ((double)f > d) f = std::nextafter(f, -INFINITY); return f; };
        const double slo = (double)ef, shi = 1.0 - (double)ef, dlo = (double)eb, dhi = 1.0 - (double)eb;
        sp.sig_lo_cmp = fceil(slo); sp.sig_lo_val = (float)slo;
        sp.sig_hi_cmp = ffloor(shi); sp.sig_hi_val = (float)shi;
        sp.der_lo_cmp = ffloor(dlo); sp.der_hi_cmp = fceil(dhi);
        sp.bsn = (float)std::exp(-std::log((double)m->Bglobal)) * m->text_scale;   // mixture weight folded into the multipliers
        sp.act = act_params(m, bn);
        sp.probs = m->probs; sp.mult = m->mult; sp.Gp = m->Gp; sp.Y = m->Y;
        sp.tf32_gp = (m->use_tc && !bn) ? 1 : 0;
        sp.Gp_lo = bn ? nullptr : m->Gp_lo;
        sp.loss_acc = m->loss_acc(); sp.col_sums = m->bwd_sums();
        sp.enorm = m->l2_entity ? m->enorm : nullptr;
        sp.escore = m->l2_entity ? m->escore : nullptr;
        // N > 1: the backward column sums + loss are all-reduced by the score kernel's last block (peer_sums_tail)
        // and, whenever the reduction is not left to ncclAllReduce, writes the loss into the pinned read-back ring
        sp.xchg = nullptr; sp.xchg_epoch = 0; sp.xchg_counter = nullptr; sp.xchg_error = nullptr; sp.loss_host = nullptr;
        if (fused_xchg(m)) {
            sp.xchg = m->peer_dev; sp.xchg_epoch = ++m->peer_epoch[3];
            sp.xchg_error = m->peer_error;
        }
        if (m->nranks <= 1 || fused_xchg(m)) {
            sp.xchg_counter = m->xchg_counter;
            sp.loss_host = m->loss_host + (m->forward_count % nvsm_model::kCostRing);
        }
        m->entity_prep_done = false;
        TRY(dispatch_score(m, sp));
    }
    phase_end(m);
    if (!fused_xchg(m)) TRY(allreduce(m, m->bwd_sums(), 2 * (size_t)dd + 1, true, 3));
    {
        const int slot = (int)(m->forward_count % nvsm_model::kCostRing);
        if (!(m->nranks <= 1 || fused_xchg(m)))   // (otherwise the score kernel's last block wrote it: score_sums_tail)
            CU(cudaMemcpyAsync(m->loss_host + slot, m->loss_acc(), sizeof(double), cudaMemcpyDeviceToHost, m->stream));
        CU(cudaEventRecord(m->loss_ev[slot], m->stream));
        m->loss_B[slot] = m->Bglobal;
        m->forward_count++;
    }
    m->have_forward = true;
    return 0;
}

// grad_transform's all-reduce may still be running on the communication stream (see backward()).
int join_gt_allreduce(nvsm_model* m) {
    if (m->gt_allreduce_pending) {
        CU(cudaStreamWaitEvent(m->stream, m->gt_reduced_ev, 0));
        m->gt_allreduce_pending = false;
    }
    return 0;
}

int reduce_gt_partials(nvsm_model* m) {
    if (m->gt_reduced) return 0;
    const long nT = (long)m->dw * m->dd;
    LAUNCH(m, reduce_partials_kernel, (int)((nT + 255) / 256), 256, 0, m->gT_part, m->gt_nparts, nT, m->gT);
    m->gt_reduced = true;
    return 0;
}

// ------------------------------------------------------------------------------------
// backward: Model::compute_gradients
// ------------------------------------------------------------------------------------
int backward(nvsm_model* m) {
    if (!m->have_forward) return fail("compute_gradients called without a forward result");
    TRY(join_gt_allreduce(m));
    const long B = m->B;
    const bool bn = m->cfg.batch_normalization != 0;
    const int dw = m->dw, dd = m->dd;

    phase_begin(m, PH_BN_BWD);
    const bool cols_kernel = bn && vec4_ok(dd) && 256 % (dd / 4) == 0;
    if (!cols_kernel)
        LAUNCH(m, bn_backward_prep_kernel, (dd + 127) / 128, 128, 0, m->bwd_sums(), dd, (double)m->Bglobal,
               m->score_shifted ? (const float*)m->b : (const float*)nullptr, m->gb, m->mean_dy, m->mean_dyx);
    if (bn) {
        if (cols_kernel) {
            const int grid = grid_for(m, B * dd / 4, 256 * 4, 8);
            // (not while the entity update waits on the auxiliary stream for the score kernel's event: an early-resident
            // bn_backward grid gets in front of it and the two no longer overlap, C2 +39 us)
            LAUNCH_PDL(m, (m->entity_async ? kPdlNone : kPdlBnBwd), bn_backward_cols_kernel, grid, 256, 0, m->Gp, m->Z, m->mean, m->invstd, (const double*)m->bwd_sums(),
                   (double)m->Bglobal, m->score_shifted ? (const float*)m->b : (const float*)nullptr, m->gb, m->mean_dy,
                   m->mean_dyx, B, dd, m->use_tc ? 1 : 0, m->Gp_lo);
        } else if (vec4_ok(dd)) {
            const int grid = grid_for(m, B * dd / 4, 256 * 4, 8);
            LAUNCH(m, bn_backward_kernel<4>, grid, 256, 0, m->Gp, m->Z, m->mean, m->invstd, m->mean_dy, m->mean_dyx, B, dd, m->use_tc ? 1 : 0, m->Gp_lo);
        } else {
            const int grid = grid_for(m, B * dd, 256 * 4, 8);
            LAUNCH(m, bn_backward_kernel<1>, grid, 256, 0, m->Gp, m->Z, m->mean, m->invstd, m->mean_dy, m->mean_dyx, B, dd, m->use_tc ? 1 : 0, m->Gp_lo);
        }
    }
    phase_end(m);

    // grad_transform[dw, dd] = P^T . dX  (K = B: split-K partials + deterministic reduce)
    auto run_gt = [&]() -> int {
        phase_begin(m, PH_GEMM_GT);
        const int tiles = ((dw + GEMM_BM - 1) / GEMM_BM) * ((dd + GEMM_BN - 1) / GEMM_BN);
        int splits = std::max(1, std::min(m->gt_splits, (2 * m->num_sms + tiles - 1) / tiles));
        splits = (int)std::min<long>(splits, std::max<long>(1, B / 64));
        int kps = (int)((B + splits - 1) / splits);
        kps = (kps + GEMM_BK - 1) / GEMM_BK * GEMM_BK;
        const int nz = (int)((B + kps - 1) / kps);
        const long nT = (long)dw * dd;
        int nparts = nz;
        if (m->use_tc) {
            const int mtiles = (dw + tc::kBlockM - 1) / tc::kBlockM;
            // split-K over the batch: one CTA per (M tile, split), but at least 8 k-blocks of 32 rows per split -- a small
            // batch otherwise becomes 128 one-k-block CTAs whose partials the projection update then has to sum (C5: 22 us)
            const int want = std::max(1, std::min(std::min(m->gt_splits, m->num_sms / mtiles), (int)(B / 256)));
            TRY(run_gemm_tc(m, true, dw, dd, (int)B, m->P, m->ldP, m->Gp, dd, m->gT_part, dd, want, nT, 1.0f, nullptr, &nparts,
                            m->P_lo, m->Gp_lo, nullptr, nullptr, kPdlGemmGt));
        } else {
            TRY((run_sgemm<true, false>(m, dw, dd, (int)B, m->P, m->ldP, m->Gp, dd, m->gT_part, dd, nz, 1.0f, nullptr)));
        }
        m->gt_nparts = nparts;
        m->gt_reduced = false;
        if ((m->nranks > 1 && !m->skip_gt_reduce) || m->no_fused_reduce) TRY(reduce_gt_partials(m));   // the all-reduce needs gT itself
        phase_end(m);
        return 0;
    };
    // grad_phrase[B, dw] = dX . T^T, scaled by 1/n (cpp/objective.cu:453-476)
    auto run_gp = [&]() -> int {
        phase_begin(m, PH_GEMM_GP);
        const float inv_n = (float)std::exp(-std::log((double)m->n));
        if (m->use_tc)
            TRY(run_gemm_tc(m, false, (int)B, dw, dd, m->Gp, dd, m->Tr, dd, m->gP, dw, 1, 0, inv_n, nullptr, nullptr, m->Gp_lo, m->Tr_lo,
                            nullptr, nullptr, kPdlGemmGp));
        else
            TRY((run_sgemm<false, true>(m, (int)B, dw, dd, m->Gp, dd, m->T, dd, m->gP, dw, 1, inv_n, nullptr)));
        if (m->l2_phrase)   // Normalizer::backward on grad_phrase (cpp/objective.cu:461-468); linear, so 1/n commutes
            LAUNCH(m, row_l2_normalize_backward_kernel, grid_for(m, B, 8, 8), 256, 0, m->gP, dw, m->P, m->P_lo, m->ldP,
                   m->p_norms, B, dw);
        phase_end(m);
        return 0;
    };
    auto nccl_gt = [&](cudaStream_t on) -> int {
        const int arc = nccl_api().AllReduce(m->gT, m->gT, (size_t)dw * dd, kNcclFloat32, kNcclSum, m->comm, (void*)on);
        if (arc != 0) return fail("ncclAllReduce: %s", nccl_api().GetErrorString(arc));
        return 0;
    };

    if (m->gt_side && m->in_fused_step && !m->profiling) {
        // Fused steps: grad_transform is only consumed by the projection update at the very end of the step, while
        // grad_phrase feeds the word update. The grad_phrase GEMM is issued first (critical path); the grad_transform
        // GEMM (+ its all-reduce at N > 1) goes to a side stream and shares the SMs with the L2-bound word update
        // instead of delaying it. Joined in update_transform.
        CU(cudaEventRecord(m->dx_ready, m->stream));
        TRY(run_gp());
        CU(cudaStreamWaitEvent(m->gt_stream, m->dx_ready, 0));
        cudaStream_t main_stream = m->stream;
        m->stream = m->gt_stream;
        int rc = run_gt();
        if (rc == 0 && m->nranks > 1) rc = nccl_gt(m->gt_stream);
        m->stream = main_stream;
        if (rc) return rc;
        CU(cudaEventRecord(m->gt_reduced_ev, m->gt_stream));
        m->gt_allreduce_pending = true;
    } else {
        const bool gt_push = fused_xchg(m) && m->in_fused_step && !m->no_fused_gt && m->comm_stream != nullptr;
        m->skip_gt_reduce = gt_push;
        const int grc = run_gt();
        m->skip_gt_reduce = false;
        if (grc) return grc;
        if (gt_push) {
            // Fused steps, NVLink peer exchange: no NCCL in the step. The split-K partial reduction pushes this rank's
            // grad_transform into every rank's inbox (gt_reduce_push_kernel, on the communication stream under the
            // grad_phrase GEMM and the word update); transform_update_kernel waits for the peers' flags and sums the
            // inbox in rank order.
            const long nT = (long)dw * dd;
            const int nblk = (int)std::min<long>(kPeerFlagStride, (nT + 255) / 256);
            const int chunk = (int)((nT + nblk - 1) / nblk);
            const unsigned long long epoch = ++m->peer_epoch[kPeerKindGt];
            cudaStream_t main_stream = m->stream;
            if (!m->profiling && m->gt_push_side) {
                CU(cudaEventRecord(m->gt_ready, m->stream));
                CU(cudaStreamWaitEvent(m->comm_stream, m->gt_ready, 0));
                m->stream = m->comm_stream;
            }
            phase_begin(m, PH_ALLREDUCE);
            gt_reduce_push_kernel<<<(int)((nT + chunk - 1) / chunk), 256, 0, m->stream>>>(m->peer_dev, m->gT_part, m->gt_nparts, nT, chunk, epoch);
            m->launches++;
            phase_end(m);
            const cudaError_t le = cudaPeekAtLastError();
            if (m->stream != main_stream) {
                CU(cudaEventRecord(m->gt_reduced_ev, m->comm_stream));
                m->gt_allreduce_pending = true;
                m->stream = main_stream;
            }
            if (le != cudaSuccess) return fail("launch gt_reduce_push_kernel: %s", cudaGetErrorString(le));
            m->gt_xchg_pending = true;
            m->gt_xchg_chunk = chunk;
        } else if (m->nranks > 1) {
            // grad_transform is only consumed by the projection update at the very end of the step: its all-reduce runs on
            // the communication stream under the grad_phrase GEMM and the word update (joined in update_transform).
            if (m->comm_stream && !m->profiling) {
                CU(cudaEventRecord(m->gt_ready, m->stream));
                CU(cudaStreamWaitEvent(m->comm_stream, m->gt_ready, 0));
                TRY(nccl_gt(m->comm_stream));
                CU(cudaEventRecord(m->gt_reduced_ev, m->comm_stream));
                m->gt_allreduce_pending = true;
            } else {
                TRY(allreduce(m, m->gT, (size_t)dw * dd, false));
            }
        }
        TRY(run_gp());
    }
    m->have_gradients = true;
    return 0;
}

// ------------------------------------------------------------------------------------
// update: Model::update — entities, words, transform (cpp/model.cu:187-220)
// ------------------------------------------------------------------------------------
struct AdamConsts {
    float b1, b2, lr1, lr2, s1, s2, eps;
};

AdamConsts adam_consts() {
    AdamConsts c;
    c.b1 = 0.9f; c.b2 = 0.999f;  // include/cuNVSM/updates.h:205-211
    c.lr1 = (float)(1.0 - (double)c.b1);
    c.lr2 = (float)(1.0 - (double)c.b2);
    c.s1 = (float)(1.0 - (double)(1.0f * c.lr1));  // storage decay with lambda = 1, lr = 1 - beta
    c.s2 = (float)(1.0 - (double)(1.0f * c.lr2));
    c.eps = 1e-6f;  // DEFAULT_EPSILON, include/cuNVSM/updates.h:21
    return c;
}

float adam_bias_correction(const AdamConsts& c, unsigned long t) {
    return (float)(std::sqrt(1.0 - std::pow((double)c.b2, (double)t)) / (1.0 - std::pow((double)c.b1, (double)t)));
}

int scale_table(nvsm_model* m, float* x, long count, float s) {
    const int grid = grid_for(m, count / 4 + 1, 256, 8);
    LAUNCH(m, scale_kernel, grid, 256, 0, x, count, s);
    return 0;
}

int scatter_entities(nvsm_model* m, float* target, float scale, const float* acc, float eps) {
    EntityScatterParams p;
    p.Z = m->Z; p.act = act_params(m, m->cfg.batch_normalization != 0);
    p.ids = m->cur->ids; p.mult = m->mult; p.B = m->B; p.R = m->R; p.dd = m->dd;
    p.target = target; p.scale = scale; p.acc = acc; p.eps = eps;
    const int grid = grid_for(m, m->B, 8, 8);
    if (vec4_ok(m->dd)) LAUNCH(m, entity_scatter_kernel<4>, grid, 256, 0, p);
    else LAUNCH(m, entity_scatter_kernel<1>, grid, 256, 0, p);
    return 0;
}

int scatter_words(nvsm_model* m, float* target, float scale, const float* acc, float eps) {
    WordScatterParams p;
    p.G = m->gP; p.ids = m->cur->features; p.fw = m->cur->fweights; p.B = m->B; p.n = m->n; p.dw = m->dw;
    p.target = target; p.scale = scale; p.acc = acc; p.eps = eps;
    const int grid = grid_for(m, m->B, 8, 8);
    if (vec4_ok(m->dw)) LAUNCH(m, word_scatter_kernel<4>, grid, 256, 0, p);
    else LAUNCH(m, word_scatter_kernel<1>, grid, 256, 0, p);
    return 0;
}

// acc_E[id] += scale * mean_k grad_entity[k, c]^2
int scatter_entity_meansq(nvsm_model* m, float* acc, float scale) {
    const float inv_dim = (float)std::exp(-std::log((double)m->dd));
    float* const tmp = m->entity_async_running ? m->rowtmp_e : m->rowtmp;
    LAUNCH(m, row_meansq_act_kernel, grid_for(m, m->B, 8, 8), 256, 0, m->Z,
           act_params(m, m->cfg.batch_normalization != 0), m->B, m->dd, inv_dim, tmp);
    const long total = m->B * m->R;
    LAUNCH(m, entity_scalar_scatter_kernel, (int)((total + kAggThreads - 1) / kAggThreads), kAggThreads, 0, m->cur->ids, m->mult, tmp,
           total, m->R, scale, acc, m->l2_entity ? (const float*)m->escore : (const float*)nullptr, inv_dim);
    return 0;
}

int scatter_word_meansq(nvsm_model* m, float* acc, float scale) {
    const float inv_dim = (float)std::exp(-std::log((double)m->dw));
    LAUNCH(m, row_meansq_kernel, grid_for(m, m->B, 8, 8), 256, 0, m->gP, m->B, m->dw, inv_dim, m->rowtmp);
    const long total = m->B * m->n;
    LAUNCH(m, word_scalar_scatter_kernel, (int)((total + kAggThreads - 1) / kAggThreads), kAggThreads, 0, m->cur->features, m->cur->fweights,
           m->rowtmp, total, m->n, scale, acc);
    return 0;
}


// ------------------------------------------------------------------------------------
// pull-style full Adam (pull_update.cuh)
// ------------------------------------------------------------------------------------
// Work list of pull_heavy_kernel for `references` bucketed references of rows `dim` wide (grown on demand: the
// all-gather sparse mode buckets the global batch).
int ensure_heavy(nvsm_model* m, HeavyWork* hw, HeavyWork** hw_dev, long references, int dim) {
    const int capacity = (int)(2 * references / kHeavyRefs + 1);
    const int ld = ((dim + 3) / 4) * 4 + 4;
    if (hw->items && hw->capacity >= capacity && hw->ld == ld) return 0;
    CU(cudaDeviceSynchronize());
    if (hw->items) cudaFree(hw->items);
    if (hw->count) cudaFree(hw->count);
    if (hw->part) cudaFree(hw->part);
    if (hw->arrivals) cudaFree(hw->arrivals);
    *hw = HeavyWork{};
    TRY(dev_alloc(&hw->items, (size_t)capacity, false));
    TRY(dev_alloc(&hw->count, 1));
    TRY(dev_alloc(&hw->part, (size_t)capacity * ld, false));
    TRY(dev_alloc(&hw->arrivals, (size_t)capacity));
    hw->capacity = capacity;
    hw->ld = ld;
    if (!*hw_dev) CU(cudaMalloc((void**)hw_dev, sizeof(HeavyWork)));
    CU(cudaMemcpy(*hw_dev, hw, sizeof(HeavyWork), cudaMemcpyHostToDevice));
    return 0;
}

void free_heavy(HeavyWork* hw) {
    if (hw->items) cudaFree(hw->items);
    if (hw->count) cudaFree(hw->count);
    if (hw->part) cudaFree(hw->part);
    if (hw->arrivals) cudaFree(hw->arrivals);
    *hw = HeavyWork{};
}

int heavy_grid(const nvsm_model* m, const HeavyWork& hw) {
    return std::max(1, std::min(m->num_sms * 4, (hw.capacity + 7) / 8));
}

// Both tables' reference buckets of a batch of B instances (local or gathered); the scan also resets the heavy-row
// work lists of the pull kernels.
int build_all_buckets(nvsm_model* m, const BatchSlot* s, long B);

int build_buckets(nvsm_model* m, const idx_t* ids, long total, long num_rows, int* counts, int* offsets, int* refs,
                  const HeavyWork* heavy, const HeavyWork* heavy_dev) {
    CU(cudaMemsetAsync(counts, 0, sizeof(int) * (num_rows + 1), m->stream));
    const int g = (int)((total + kAggThreads - 1) / kAggThreads);
    LAUNCH(m, ref_count_kernel, g, kAggThreads, 0, ids, total, counts);
    const int nb = (int)((num_rows + 1023) / 1024);
    LAUNCH(m, scan_blocks_kernel, nb, 1024, 0, counts, num_rows, offsets, m->scan_tmp, heavy_dev ? heavy->count : (int*)nullptr);
    LAUNCH(m, scan_blocks_kernel, 1, 1024, 0, m->scan_tmp, (long)nb, m->scan_tmp + 1024, (int*)nullptr, (int*)nullptr);
    LAUNCH(m, scan_add_kernel, nb, 1024, 0, offsets, num_rows, m->scan_tmp + 1024, total, (const int*)counts, heavy_dev);
    LAUNCH(m, ref_fill_kernel, g, kAggThreads, 0, ids, total, offsets, counts, refs);
    return 0;
}

int build_all_buckets(nvsm_model* m, const BatchSlot* s, long B) {
    TRY(ensure_heavy(m, &m->heavy_e, &m->heavy_e_dev, B * m->R, m->dd));
    TRY(ensure_heavy(m, &m->heavy_w, &m->heavy_w_dev, B * m->n, m->dw));
    const bool on = !m->no_heavy;
    TRY(build_buckets(m, s->ids, B * m->R, m->D, m->e_counts, m->e_offsets, m->e_refs, &m->heavy_e, on ? m->heavy_e_dev : nullptr));
    return build_buckets(m, s->features, B * m->n, m->V, m->w_counts, m->w_offsets, m->w_refs, &m->heavy_w, on ? m->heavy_w_dev : nullptr);
}

// In a fused step the entity table is updated on the auxiliary stream (start_entity_update). The heavy-row kernel of
// the WORD table goes there as well, behind it: the projection update on the main stream then does not wait for it —
// with uniform ids it is an empty launch, with skewed ids tens of microseconds of work that now overlap — and
// update()'s join (entity_done, re-recorded here) orders everything after the step behind it.
struct HeavyStream {
    nvsm_model* m;
    cudaStream_t saved;
    bool cross;
};

int heavy_stream_begin(nvsm_model* m, bool entities, HeavyStream* hs) {
    hs->m = m;
    hs->saved = m->stream;
    hs->cross = !entities && m->entity_async && !m->entity_async_running && m->stream != m->aux_stream;
    if (hs->cross) {
        CU(cudaEventRecord(m->word_rows_done, m->stream));
        CU(cudaStreamWaitEvent(m->aux_stream, m->word_rows_done, 0));
        m->stream = m->aux_stream;
    }
    return 0;
}

int heavy_stream_end(HeavyStream* hs, int rc) {
    if (!hs->cross) return rc;
    hs->m->stream = hs->saved;
    if (rc) return rc;
    CU(cudaEventRecord(hs->m->entity_done, hs->m->aux_stream));
    return 0;
}

template <int VEC, int NCH>
int launch_pull(nvsm_model* m, bool entities, const AdamFullConsts& k) {
    const int grid = grid_for(m, entities ? m->D : m->V, 8, 8);
    HeavyWork* const hw = entities ? &m->heavy_e : &m->heavy_w;
    if (!hw->items) return fail("pull update without a heavy-row work list (build_all_buckets)");
    // rows above kHeavyRefs references are on the list the bucket build made: the row kernel skips them
    const int row_hw = m->no_heavy ? 0 : 1;
    const int heavy_above = row_hw ? kHeavyRefs : INT_MAX;
    if (entities) {
        const float* const self_k = m->l2_entity ? (const float*)m->kself : (const float*)nullptr;
        LAUNCH(m, (adam_full_pull_kernel<VEC, NCH, true>), grid, 256, 0, m->E, m->optE.m, m->optE.v, m->D, m->dd,
               m->e_offsets, m->e_refs, m->mult, m->Y, m->R, k, self_k, heavy_above);
        if (row_hw) LAUNCH(m, (pull_heavy_kernel<VEC, NCH, true, AdamFullApply>), heavy_grid(m, *hw), 256, 0, m->dd, m->e_offsets, m->e_refs,
               (const float*)m->mult, (const float*)m->Y, m->R, (const float*)nullptr, *hw,
               AdamFullApply{m->E, m->optE.m, m->optE.v, k, self_k});
    } else {
        LAUNCH(m, (adam_full_pull_kernel<VEC, NCH, false>), grid, 256, 0, m->W, m->optW.m, m->optW.v, m->V, m->dw,
               m->w_offsets, m->w_refs, m->cur->fweights, m->gP, m->n, k, (const float*)nullptr, heavy_above);
        if (row_hw) {
            HeavyStream hs;
            TRY(heavy_stream_begin(m, entities, &hs));
            const int rc = [&]() -> int {
                LAUNCH(m, (pull_heavy_kernel<VEC, NCH, false, AdamFullApply>), heavy_grid(m, *hw), 256, 0, m->dw, m->w_offsets,
                       m->w_refs, (const float*)m->cur->fweights, (const float*)m->gP, m->n, (const float*)nullptr, *hw,
                       AdamFullApply{m->W, m->optW.m, m->optW.v, k, (const float*)nullptr});
                return 0;
            }();
            TRY(heavy_stream_end(&hs, rc));
        }
    }
    return 0;
}

template <int VEC, int NCH>
int launch_sgd_pull(nvsm_model* m, bool entities, float decay, float lr, bool touch_all, float* acc, const float* ysq,
                    const float* word_coefs) {
    const int grid = grid_for(m, entities ? m->D : m->V, 8, 8);
    HeavyWork* const hw = entities ? &m->heavy_e : &m->heavy_w;
    if (!hw->items) return fail("pull update without a heavy-row work list (build_all_buckets)");
    // rows above kHeavyRefs references are on the list the bucket build made: the row kernel skips them
    const int row_hw = m->no_heavy ? 0 : 1;
    const int heavy_above = row_hw ? kHeavyRefs : INT_MAX;
    // sgd_pull_sparse_kernel: no dense decay and at most ~2 references per row on average (C3 entities 1.7, C5 0.13): the
    // scan and the per-row load chain are the cost. Denser tables keep the plain row loop (more warps, fewer registers).
    bool sparse = !touch_all && (double)(entities ? m->B * m->R : m->B * m->n) <= 2.0 * (double)(entities ? m->D : m->V);
    if (m->knobs.sgd_sparse >= 0) sparse = !touch_all && m->knobs.sgd_sparse != 0;
    if (entities) {
        if (!sparse)
            LAUNCH(m, (sgd_pull_kernel<VEC, NCH, true>), grid, 256, 0, m->E, m->D, m->dd, m->e_offsets, m->e_refs,
                   (const float*)m->mult, (const float*)m->Y, m->R, decay, lr, touch_all ? 1 : 0, acc, ysq, 1e-6f, heavy_above);
        else
            LAUNCH(m, (sgd_pull_sparse_kernel<VEC, NCH, true>), grid, 256, 0, m->E, m->D, m->dd, m->e_offsets, m->e_refs,
                   (const float*)m->mult, (const float*)m->Y, m->R, decay, lr, acc, ysq, 1e-6f, heavy_above);
        if (row_hw) LAUNCH(m, (pull_heavy_kernel<VEC, NCH, true, SgdApply>), heavy_grid(m, *hw), 256, 0, m->dd, m->e_offsets, m->e_refs,
               (const float*)m->mult, (const float*)m->Y, m->R, ysq, *hw, SgdApply{m->E, decay, lr, acc, 1e-6f});
    } else {
        if (!sparse)
            LAUNCH(m, (sgd_pull_kernel<VEC, NCH, false>), grid, 256, 0, m->W, m->V, m->dw, m->w_offsets, m->w_refs, word_coefs,
                   (const float*)m->gP, m->n, decay, lr, touch_all ? 1 : 0, (float*)nullptr, (const float*)nullptr, 1e-6f, heavy_above);
        else
            LAUNCH(m, (sgd_pull_sparse_kernel<VEC, NCH, false>), grid, 256, 0, m->W, m->V, m->dw, m->w_offsets, m->w_refs, word_coefs,
                   (const float*)m->gP, m->n, decay, lr, (float*)nullptr, (const float*)nullptr, 1e-6f, heavy_above);
        if (row_hw) {
            HeavyStream hs;
            TRY(heavy_stream_begin(m, entities, &hs));
            const int rc = [&]() -> int {
                LAUNCH(m, (pull_heavy_kernel<VEC, NCH, false, SgdApply>), heavy_grid(m, *hw), 256, 0, m->dw, m->w_offsets,
                       m->w_refs, word_coefs, (const float*)m->gP, m->n, (const float*)nullptr, *hw,
                       SgdApply{m->W, decay, lr, (float*)nullptr, 1e-6f});
                return 0;
            }();
            TRY(heavy_stream_end(&hs, rc));
        }
    }
    return 0;
}

// SGD / Adagrad through the reference buckets (sgd_pull_kernel).
int pull_sgd(nvsm_model* m, bool entities, float lr, float lambda) {
    if (!m->buckets_in_flight) return fail("pull update without reference buckets");
    CU(cudaStreamWaitEvent(m->stream, m->buckets_ready, 0));
    const bool adagrad = m->cfg.update_method == NVSM_ADAGRAD;
    const float decay = lambda > 0.0f ? (float)(1.0 - (double)(lambda * lr)) : 1.0f;
    // The reference's dense whole-table decay (cpp/storage.cu:65-67) multiplies by the FLOAT 1 - lambda_s * lr. For the
    // large-batch configurations that factor rounds to exactly 1.0f (C3: lambda_s lr = 2e-9, C5: 2.4e-8, both below half
    // an ulp of 1), so the pass is a bit-exact no-op: rows without references are then not touched at all.
    const bool touch_all = decay != 1.0f;
    float* acc = nullptr;
    const float* ysq = nullptr;
    const float* word_coefs = m->cur->fweights;
    if (adagrad && entities) {
        float* const tmp = m->entity_async_running ? m->rowtmp_e : m->rowtmp;
        const float inv_dim = (float)std::exp(-std::log((double)m->dd));
        LAUNCH(m, row_meansq_act_kernel, grid_for(m, m->B, 8, 8), 256, 0, m->Z, act_params(m, m->cfg.batch_normalization != 0),
               m->B, m->dd, inv_dim, tmp);
        acc = m->optE.acc; ysq = tmp;
    } else if (adagrad) {
        TRY(scatter_word_meansq(m, m->optW.acc, 1.0f));   // V scalars: atomics are fine here
        LAUNCH(m, word_adagrad_coef_kernel, (int)((m->B + 255) / 256), 256, 0, (const idx_t*)m->cur->features,
               (const float*)m->cur->fweights, (const float*)m->optW.acc, m->B, m->n, 1e-6f, m->wcoef);
        word_coefs = m->wcoef;
    }
    const int dim = entities ? m->dd : m->dw;
    if (vec4_ok(dim)) {
        const int nch = (dim / 4 + 31) / 32;
        if (nch <= 1) return launch_sgd_pull<4, 1>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
        if (nch <= 2) return launch_sgd_pull<4, 2>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
        if (nch <= 3) return launch_sgd_pull<4, 3>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
        if (nch <= 4) return launch_sgd_pull<4, 4>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
        return launch_sgd_pull<4, 8>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
    }
    const int nch = (dim + 31) / 32;
    if (nch <= 1) return launch_sgd_pull<1, 1>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
    if (nch <= 4) return launch_sgd_pull<1, 4>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
    if (nch <= 16) return launch_sgd_pull<1, 16>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
    return launch_sgd_pull<1, 32>(m, entities, decay, lr, touch_all, acc, ysq, word_coefs);
}

// The ids of a batch are known before the forward pass starts, so both bucket sets are built on
// the auxiliary stream while the forward / backward kernels run on the main stream.
int start_bucket_build(nvsm_model* m, BatchSlot* s) {
    cudaStream_t main_stream = m->stream;
    if (m->buckets_ever_consumed) CU(cudaStreamWaitEvent(m->aux_stream, m->buckets_consumed, 0));
    CU(cudaStreamWaitEvent(m->aux_stream, s->ready, 0));
    m->stream = m->aux_stream;   // LAUNCH targets m->stream
    const bool prof = m->profiling;
    m->profiling = false;
    phase_begin(m, PH_BUCKETS);   // (timeline mode only)
    int rc = build_all_buckets(m, s, s->B);
    phase_end(m);
    m->stream = main_stream;
    m->profiling = prof;
    if (rc) return rc;
    CU(cudaEventRecord(m->buckets_ready, m->aux_stream));
    m->buckets_in_flight = true;
    return 0;
}

int pull_update(nvsm_model* m, bool entities, const AdamFullConsts& k) {
    // (both tables wait: with the asynchronous entity update the two run on different streams)
    if (!m->buckets_in_flight) return fail("pull update without reference buckets");
    CU(cudaStreamWaitEvent(m->stream, m->buckets_ready, 0));
    const int dim = entities ? m->dd : m->dw;
    if (vec4_ok(dim)) {
        const int nch = (dim / 4 + 31) / 32;
        if (nch <= 1) return launch_pull<4, 1>(m, entities, k);
        if (nch <= 2) return launch_pull<4, 2>(m, entities, k);
        if (nch <= 3) return launch_pull<4, 3>(m, entities, k);
        if (nch <= 4) return launch_pull<4, 4>(m, entities, k);
        return launch_pull<4, 8>(m, entities, k);
    }
    const int nch = (dim + 31) / 32;
    if (nch <= 1) return launch_pull<1, 1>(m, entities, k);
    if (nch <= 4) return launch_pull<1, 4>(m, entities, k);
    if (nch <= 16) return launch_pull<1, 16>(m, entities, k);
    return launch_pull<1, 32>(m, entities, k);
}

int update_table(nvsm_model* m, bool entities, float lr, float lambda) {
    float* theta = entities ? m->E : m->W;
    TableOpt& opt = entities ? m->optE : m->optW;
    const long N = entities ? m->D : m->V;
    const int dim = entities ? m->dd : m->dw;
    const long count = N * dim;
    // gradient descriptors of this table (CompositeGradients::get_representations_gradient): the TextEntity one
    // and / or the RepresentationSimilarity one (window 1, no weights)
    const bool text = m->has_text;
    const bool pair = m->has_pair && m->pair_entities == entities;
    if (!text && !pair) return 0;   // "No gradient": the reference skips the table (cpp/params.cu:301-304)
    const long pairM = 2 * m->pair_N;
    auto pair_scatter = [&](float* target, float scale, const float* acc, float eps) -> int {
        const int grid = grid_for(m, pairM, 8, 8);
        if (vec4_ok(dim)) LAUNCH(m, rows_scatter_kernel<4>, grid, 256, 0, (const float*)m->pair_G, (const idx_t*)m->pair_ids, pairM, dim, target, scale, acc, eps);
        else LAUNCH(m, rows_scatter_kernel<1>, grid, 256, 0, (const float*)m->pair_G, (const idx_t*)m->pair_ids, pairM, dim, target, scale, acc, eps);
        return 0;
    };
    auto scatter = [&](float* target, float scale, const float* acc, float eps) -> int {
        if (text) TRY(entities ? scatter_entities(m, target, scale, acc, eps) : scatter_words(m, target, scale, acc, eps));
        if (pair) TRY(pair_scatter(target, scale, acc, eps));
        return 0;
    };
    auto scatter_meansq = [&](float* acc, float scale) -> int {
        if (text) TRY(entities ? scatter_entity_meansq(m, acc, scale) : scatter_word_meansq(m, acc, scale));
        if (pair) {
            const float inv_dim = (float)std::exp(-std::log((double)dim));
            LAUNCH(m, row_meansq_kernel, grid_for(m, pairM, 8, 8), 256, 0, (const float*)m->pair_G, pairM, dim, inv_dim, m->pair_msq);
            LAUNCH(m, rows_scalar_scatter_kernel, (int)((pairM + 255) / 256), 256, 0, (const idx_t*)m->pair_ids,
                   (const float*)m->pair_msq, pairM, scale, acc);
        }
        return 0;
    };
    // RepresentationsStorage::update (cpp/storage.cu:51-102): dense decay, then scatter.
    const bool self = entities && m->l2_entity;   // gradient carries - kself[d] * E_d (entity normalisation)
    auto self_axpy = [&](float* target, float coef) -> int {
        LAUNCH(m, row_self_axpy_kernel, grid_for(m, count, 256 * 4, 8), 256, 0, target, (const float*)theta, N, dim, coef,
               (const float*)m->kself);
        return 0;
    };
    auto sgd = [&](const float* acc, float eps) -> int {
        const float decay = lambda > 0.0f ? (float)(1.0 - (double)(lambda * lr)) : 1.0f;
        if (self)
            LAUNCH(m, scale_rows_self_kernel, grid_for(m, count, 256 * 4, 8), 256, 0, theta, N, dim, decay, lr,
                   (const float*)m->kself, acc, eps);
        else if (decay != 1.0f) TRY(scale_table(m, theta, count, decay));   // (x * 1.0f == x: skipping is bit-exact)
        return scatter(theta, lr, acc, eps);
    };
    const int method = m->cfg.update_method;
    if (m->pull && (method == NVSM_SGD || method == NVSM_ADAGRAD) && text && !pair && !self && N >= kPullMinRows &&
        !(exact_sparse(m) && method == NVSM_ADAGRAD))   // (the gathered batch has no word-coefficient buffer)
        return pull_sgd(m, entities, lr, lambda);
    if (method == NVSM_SGD) return sgd(nullptr, 0.f);
    if (text && pair && (method == NVSM_ADAGRAD || (method == NVSM_ADAM && m->cfg.adam_mode == NVSM_ADAM_SPARSE)))
        return fail("Adagrad / sparse Adam do not implement multiple gradients (cpp/updates_adagrad.cu:108-109, "
                    "cpp/updates_adam.cu:339-340): use sgd, dense_adam or full_adam with a mixture objective");
    if (method == NVSM_ADAGRAD) {  // cpp/updates_adagrad.cu:99-179
        TRY(scatter_meansq(opt.acc, 1.0f));
        return sgd(opt.acc, 1e-6f);
    }
    // Adam, cpp/updates_adam.cu:153-385
    const AdamConsts c = adam_consts();
    const float bc = adam_bias_correction(c, opt.t);
    opt.t += 1;
    const int mode = m->cfg.adam_mode;
    if (mode == NVSM_ADAM_DENSE_UPDATE_DENSE_VARIANCE && m->pull) {
        AdamFullConsts k;
        k.s1 = c.s1; k.lr1 = c.lr1; k.reg1 = (float)((1.0 - (double)c.b1) * (double)lambda);
        k.s2 = c.s2; k.lr2 = c.lr2; k.lambda = lambda; k.lr = lr; k.bc = bc; k.eps = c.eps;
        return pull_update(m, entities, k);
    }
    if (mode == NVSM_ADAM_DENSE_UPDATE_DENSE_VARIANCE) {
        TRY(scatter(opt.agg, 1.0f, nullptr, 0.f));
        if (self) TRY(self_axpy(opt.agg, -1.0f));
        const float reg1 = (float)((1.0 - (double)c.b1) * (double)lambda);
        const int grid = grid_for(m, count / 4 + 1, 256, 8);
        LAUNCH(m, adam_full_kernel, grid, 256, 0, theta, opt.m, opt.v, opt.agg, count, c.s1, c.lr1, reg1, c.s2,
               c.lr2, lambda, lr, bc, c.eps);
        return 0;
    }
    // SPARSE and DENSE_UPDATE share the moment updates: m dense decay + scatter, scalar v.
    TRY(scale_table(m, opt.m, count, c.s1));
    TRY(scatter(opt.m, c.lr1, nullptr, 0.f));
    if (self) TRY(self_axpy(opt.m, -c.lr1));
    TRY(scale_table(m, opt.v, N, c.s2));
    TRY(scatter_meansq(opt.v, c.lr2));
    if (mode == NVSM_ADAM_DENSE_UPDATE) {
        const int grid = grid_for(m, count, 256 * 4, 8);
        LAUNCH(m, adam_dense_update_kernel, grid, 256, 0, theta, opt.m, opt.v, N, dim,
               (float)(1.0 - (double)(lambda * lr)), lr, bc, c.eps);
        return 0;
    }
    // SPARSE: window-averaged step, applied through the SGD scatter with dense decay.
    if (lambda > 0.0f && (float)(1.0 - (double)(lambda * lr)) != 1.0f) TRY(scale_table(m, theta, count, (float)(1.0 - (double)(lambda * lr))));
    if (pair) {   // single descriptor, window 1: same per-reference step as the entity side
        const int grid = grid_for(m, pairM, 8, 8);
        if (vec4_ok(dim))
            LAUNCH(m, adam_sparse_entity_kernel<4>, grid, 256, 0, (const idx_t*)m->pair_ids, pairM, dim, opt.m, opt.v, bc, c.eps, lr, theta);
        else
            LAUNCH(m, adam_sparse_entity_kernel<1>, grid, 256, 0, (const idx_t*)m->pair_ids, pairM, dim, opt.m, opt.v, bc, c.eps, lr, theta);
        return 0;
    }
    if (entities) {
        const long total = m->B * m->R;
        const int grid = grid_for(m, total, 8, 8);
        if (vec4_ok(dim))
            LAUNCH(m, adam_sparse_entity_kernel<4>, grid, 256, 0, m->cur->ids, total, dim, opt.m, opt.v, bc, c.eps, lr, theta);
        else
            LAUNCH(m, adam_sparse_entity_kernel<1>, grid, 256, 0, m->cur->ids, total, dim, opt.m, opt.v, bc, c.eps, lr, theta);
        return 0;
    }
    {
        const int grid = grid_for(m, m->B, 8, 8);
        if (vec4_ok(dim))
            LAUNCH(m, adam_sparse_word_grad_kernel<4>, grid, 256, 0, m->cur->features, m->B, m->n, dim, opt.m, opt.v, bc, c.eps, m->gP);
        else
            LAUNCH(m, adam_sparse_word_grad_kernel<1>, grid, 256, 0, m->cur->features, m->B, m->n, dim, opt.m, opt.v, bc, c.eps, m->gP);
    }
    return scatter(theta, lr, nullptr, 0.f);
}

int update_transform(nvsm_model* m, float lr, float lambda) {
    TRY(join_gt_allreduce(m));
    TransformUpdateParams p;
    p.T = m->T; p.b = m->b; p.gT = m->gT; p.gb = m->gb;
    p.nT = (long)m->dw * m->dd; p.nb = m->dd;
    p.method = m->cfg.update_method;
    p.lr = lr; p.lambda = lambda;
    p.aT = m->T_a; p.ab = m->b_a; p.vT = m->T_v; p.vb = m->b_v;
    const AdamConsts c = adam_consts();
    p.s1 = c.s1; p.lr1 = c.lr1; p.s2 = c.s2; p.lr2 = c.lr2; p.eps = c.eps;
    p.bc = 1.0f;
    if (p.method == NVSM_ADAM) {
        p.bc = adam_bias_correction(c, m->t_transform);
        m->t_transform += 1;
    }
    p.gT_part = nullptr; p.nparts = 0; p.gT_out = m->gT;
    p.xchg = nullptr; p.xchg_epoch = 0; p.xchg_chunk = 1; p.xchg_error = nullptr;
    if (m->gt_xchg_pending) {
        // every rank's grad_transform sits (or is about to land) in this rank's inbox: sum the slots in rank order
        const unsigned long long epoch = m->peer_epoch[kPeerKindGt];
        p.gT_part = m->peer.gt_inbox[m->rank] + (size_t)(epoch & 1ull) * m->nranks * m->peer.gt_elems;
        p.nparts = m->nranks;
        p.xchg = m->peer_dev; p.xchg_epoch = epoch; p.xchg_chunk = m->gt_xchg_chunk; p.xchg_error = m->peer_error;
        m->gt_xchg_pending = false;
        m->gt_reduced = true;
    } else if (!m->gt_reduced) { p.gT_part = m->gT_part; p.nparts = m->gt_nparts; m->gt_reduced = true; }
    p.Tr = nullptr; p.Tt = nullptr; p.Tr_lo = nullptr; p.Tt_lo = nullptr; p.dd = m->dd; p.ldT = m->ldP;
    if (m->use_tc) { p.Tr = m->Tr; p.Tt = m->Tt; p.Tr_lo = m->Tr_lo; p.Tt_lo = m->Tt_lo; m->t_copies_stale = false; }
    LAUNCH(m, transform_update_kernel, (int)((p.nT + p.nb + 127) / 128), 128, 0, p);
    return 0;
}

// NVSM_SPARSE_ALLGATHER (SURVEY.md 8e "parity mode"): gather every rank's rows of (entity ids, multipliers,
// activations, word ids, word weights, grad_phrase) so that each replica applies the table updates of the whole
// global batch. Returns with the model's per-step pointers redirected to the gathered copies; `saved` restores them.
struct LocalView {
    BatchSlot* cur; long B; float *mult, *Z, *Y, *gP, *rowtmp, *escore; int *e_refs, *w_refs;
};

int gather_global_batch(nvsm_model* m, LocalView* saved) {
    const long B = m->B, W = m->nranks;
    BatchSlot* s = m->cur;
    phase_begin(m, PH_ALLREDUCE);
    TRY(allgather_bytes(m, s->ids, m->ag_slot.ids, sizeof(idx_t) * B * m->R));
    TRY(allgather_bytes(m, m->mult, m->ag_mult, sizeof(float) * B * m->R));
    TRY(allgather_bytes(m, m->pull ? m->Y : m->Z, m->ag_act, sizeof(float) * B * m->dd));
    TRY(allgather_bytes(m, s->features, m->ag_slot.features, sizeof(idx_t) * B * m->n));
    TRY(allgather_bytes(m, s->fweights, m->ag_slot.fweights, sizeof(float) * B * m->n));
    TRY(allgather_bytes(m, m->gP, m->ag_gP, sizeof(float) * B * m->dw));
    if (m->l2_entity) TRY(allgather_bytes(m, m->escore, m->ag_escore, sizeof(float) * B * m->R));
    phase_end(m);
    *saved = LocalView{m->cur, m->B, m->mult, m->Z, m->Y, m->gP, m->rowtmp, m->escore, m->e_refs, m->w_refs};
    if (m->l2_entity) m->escore = m->ag_escore;
    m->ag_slot.B = B * W;
    m->cur = &m->ag_slot; m->B = B * W;
    m->mult = m->ag_mult; m->Z = m->ag_act; m->Y = m->ag_act; m->gP = m->ag_gP; m->rowtmp = m->ag_rowtmp;
    m->e_refs = m->ag_e_refs; m->w_refs = m->ag_w_refs;
    return 0;
}

void restore_local_view(nvsm_model* m, const LocalView& v) {
    m->cur = v.cur; m->B = v.B; m->mult = v.mult; m->Z = v.Z; m->Y = v.Y; m->gP = v.gP; m->rowtmp = v.rowtmp;
    m->escore = v.escore; m->e_refs = v.e_refs; m->w_refs = v.w_refs;
}

int update(nvsm_model* m, float lr, float lambda) {
    if (!m->have_gradients) return fail("update called without gradients");
    if (m->has_pair && !m->have_pair_forward) return fail("this objective needs nvsm_similarity_compute_cost before update");
    if (!m->has_text) {   // EntityEntity / TermTerm: one table, nothing else (the reference skips gradient-less params)
        if (lr < 0.f || lambda < 0.f) return fail("learning rate and lambda must be >= 0");
        phase_begin(m, m->pair_entities ? PH_UPD_ENTITIES : PH_UPD_WORDS);
        const int prc = update_table(m, m->pair_entities, lr, lambda);
        phase_end(m);
        m->have_gradients = false;
        m->have_pair_forward = false;
        return prc;
    }
    if (lr < 0.f || lambda < 0.f) return fail("learning rate and lambda must be >= 0");
    const bool exact = exact_sparse(m);
    LocalView local{};
    float* const mult_orig = m->mult;
    if (m->l2_entity) {
        // Normalizer::backward of the entity rows, split into an effective multiplier and a per-row self term
        phase_begin(m, PH_UPD_ENTITIES);
        CU(cudaMemsetAsync(m->kself, 0, sizeof(float) * m->D, m->stream));
        const long total = m->B * m->R;
        LAUNCH(m, entity_norm_prep_kernel, (int)((total + 255) / 256), 256, 0, m->cur->ids, (const float*)m->mult,
               (const float*)m->enorm, (const float*)m->escore, total, m->R, m->mult_eff, m->kself);
        phase_end(m);
        if (exact) TRY(allreduce(m, m->kself, (size_t)m->D, false));
        m->mult = m->mult_eff;
    }
    if (exact) {
        if (int grc = gather_global_batch(m, &local)) { m->mult = mult_orig; return grc; }
        if (m->pull) {  // buckets over the gathered ids, on the main stream
            phase_begin(m, PH_UPD_ENTITIES);
            int rc = build_all_buckets(m, m->cur, m->B);
            phase_end(m);
            if (rc) { restore_local_view(m, local); m->mult = mult_orig; return rc; }
            CU(cudaEventRecord(m->buckets_ready, m->stream));
            m->buckets_in_flight = true;
        }
    }
    int rc = 0;
    if (!m->entity_async) {
        phase_begin(m, PH_UPD_ENTITIES);
        rc = update_table(m, true, lr, lambda);
        phase_end(m);
    }
    if (rc == 0) {
        phase_begin(m, PH_UPD_WORDS);
        rc = update_table(m, false, lr, lambda);
        phase_end(m);
    }
    if (exact) restore_local_view(m, local);
    m->mult = mult_orig;
    if (rc) return rc;
    if (m->pull) {
        CU(cudaEventRecord(m->buckets_consumed, m->stream));
        m->buckets_ever_consumed = true;
        m->buckets_in_flight = false;
    }
    phase_begin(m, PH_UPD_TRANSFORM);
    TRY(update_transform(m, lr, lambda));
    phase_end(m);
    if (m->entity_async) {   // join: everything after update() is ordered behind the entity update as well
        CU(cudaStreamWaitEvent(m->stream, m->entity_done, 0));
        m->entity_async = false;
    }
    // The batch slot may be overwritten once everything enqueued so far has run.
    CU(cudaEventRecord(m->cur->consumed, m->stream));
    m->cur->ever_consumed = true;
    m->cur->in_use = false;
    m->have_gradients = false;  // grad_phrase may have been overwritten (sparse Adam)
    m->have_pair_forward = false;
    return 0;
}


// Fused steps (forward, backward and update issued by one call): Model::update applies the entity gradients first
// (cpp/model.cu:200-206) and those only depend on the forward pass (multipliers, activations, ids), so the entity
// table is updated on the auxiliary stream while batch-norm backward and the grad_transform / grad_phrase GEMMs run on
// the main stream. The bandwidth-bound update and the pipeline-bound GEMMs share the SMs; nothing they touch overlaps
// (E, its moments, mult, Y | Gp, P, T, gT, gP). Same arithmetic, same order per table.
int start_entity_update(nvsm_model* m, float lr, float lambda) {
    if (m->profiling || exact_sparse(m) || m->l2_entity || m->has_pair || !m->has_text || m->knobs.no_overlap) return 0;
    if (lr < 0.f || lambda < 0.f) return 0;   // update() reports it
    CU(cudaEventRecord(m->score_done, m->stream));
    CU(cudaStreamWaitEvent(m->aux_stream, m->score_done, 0));
    cudaStream_t main_stream = m->stream;
    m->stream = m->aux_stream;
    m->entity_async_running = true;
    phase_begin(m, PH_UPD_ENTITIES);   // (timeline mode only: serial profiling never gets here)
    const int rc = update_table(m, true, lr, lambda);
    phase_end(m);
    m->entity_async_running = false;
    m->stream = main_stream;
    if (rc) return rc;
    CU(cudaEventRecord(m->entity_done, m->aux_stream));
    m->entity_async = true;
    return 0;
}

int fused_step(nvsm_model* m, BatchSlot* s, float lr) {
    TRY(forward(m, s));
    const float lambda = m->cfg.regularization_lambda / (float)m->Bglobal;
    TRY(start_entity_update(m, lr, lambda));
    m->in_fused_step = true;
    const int brc = backward(m);
    m->in_fused_step = false;
    if (brc) return brc;
    return update(m, lr, lambda);
}

// ------------------------------------------------------------------------------------
// device sampler
// ------------------------------------------------------------------------------------
// Candidates needed for N draws: N plus the expected rejections with a wide safety margin.
long sampler_candidates(long N, long D) {
    const unsigned long urngrange = 2147483645ul;
    const unsigned long scaling = urngrange / (unsigned long)D, past = (unsigned long)D * scaling;
    const double p_rej = (double)(urngrange + 1 - past) / (double)(urngrange + 1);
    const double expect = (double)N * p_rej / std::max(1e-9, 1.0 - p_rej);
    long T = N + (long)(expect * 1.5 + 8.0 * std::sqrt(expect + 1.0)) + 1024;
    return (T + kSamplerChunk - 1) / kSamplerChunk * kSamplerChunk;
}

int ensure_sampler(nvsm_model* m) {
    if (!m->rng_dev) {
        TRY(dev_alloc(&m->rng_dev, 2));
        TRY(dev_alloc(&m->smp_error, 1));
        TRY(dev_alloc(&m->smp_scan, 2 * 1024 + 2));
        TRY(dev_alloc(&m->smp_rej_items, kRejListCap));
        TRY(dev_alloc(&m->smp_rej_count, 2));
    }
    const long chunks = sampler_candidates(m->maxB * std::max(1, m->z), m->D) / kSamplerChunk;
    if (chunks > m->smp_capacity) {
        if (m->smp_counts) cudaFree(m->smp_counts);
        if (m->smp_offsets) cudaFree(m->smp_offsets);
        m->smp_counts = m->smp_offsets = nullptr;
        TRY(dev_alloc(&m->smp_counts, chunks + 1));
        TRY(dev_alloc(&m->smp_offsets, chunks + 1));
        m->smp_capacity = chunks;
    }
    return 0;
}

// ids[slot] <- labels[slot] + z device-sampled negatives per instance; advances the device engine state.
int sample_labels_device(nvsm_model* m, const idx_t* labels, idx_t* ids, long B, int z, long D) {
    if (!m->rng_seeded) return fail("device sampler used before nvsm_sampler_seed");
    if (D <= 0 || D >= 2147483646L) return fail("device sampler supports 0 < num_entities < 2^31 - 2");
    if (z == 0) {
        LAUNCH(m, sampler_copy_labels_kernel, (int)((B + 255) / 256), 256, 0, labels, B, ids);
        return 0;
    }
    if (m->smp_cdf && D == m->D) {   // skewed negatives over the model's own entity table
        const long draws = B * z;
        LAUNCH(m, sampler_cdf_kernel, (int)((draws + 255) / 256), 256, 0, m->rng_dev + m->rng_cur,
               m->rng_dev + (m->rng_cur ^ 1), labels, ids, draws, z, m->smp_cdf, D);
        m->rng_cur ^= 1;
        return 0;
    }
    SamplerParams p;
    p.state_in = m->rng_dev + m->rng_cur;
    p.state_out = m->rng_dev + (m->rng_cur ^ 1);
    p.num_draws = B * z;
    p.num_candidates = sampler_candidates(p.num_draws, D);
    p.scaling = (unsigned int)(2147483645ul / (unsigned long)D);
    p.past = (unsigned int)((unsigned long)D * p.scaling);
    p.z = z; p.R = z + 1;
    p.labels = labels; p.ids = ids;
    p.counts = m->smp_counts; p.offsets = m->smp_offsets; p.error_flag = m->smp_error;
    const long nchunks = p.num_candidates / kSamplerChunk;
    if (nchunks > 1024L * 1024L) return fail("device sampler: batch too large");
    if (nchunks > m->smp_capacity) {
        CU(cudaStreamSynchronize(m->stream));
        if (m->smp_counts) cudaFree(m->smp_counts);
        if (m->smp_offsets) cudaFree(m->smp_offsets);
        m->smp_counts = m->smp_offsets = nullptr;
        TRY(dev_alloc(&m->smp_counts, nchunks + 1));
        TRY(dev_alloc(&m->smp_offsets, nchunks + 1));
        m->smp_capacity = nchunks;
        p.counts = m->smp_counts; p.offsets = m->smp_offsets;
    }
    const int grid = (int)((nchunks + 255) / 256);
    {
        // expected rejected candidates of this call; the list path holds kRejListCap chunks (overflow -> error flag)
        const unsigned long urngrange = 2147483645ul;
        const double p_rej = (double)(urngrange + 1 - (unsigned long)p.past) / (double)(urngrange + 1);
        if (!m->knobs.sampler_scan && p_rej * (double)p.num_candidates <= 512.0) {
            RejList rl;
            rl.items = m->smp_rej_items;
            rl.count = m->smp_rej_count + (m->smp_calls & 1);
            rl.count_next = m->smp_rej_count + ((m->smp_calls + 1) & 1);
            m->smp_calls++;
            LAUNCH(m, sampler_count_list_kernel, grid, 256, 0, p, rl);
            LAUNCH(m, sampler_fill_list_kernel, std::max(grid, (int)((B + 255) / 256)), 256, 0, p, rl);
            m->rng_cur ^= 1;
            return 0;
        }
    }
    LAUNCH(m, sampler_count_kernel, grid, 256, 0, p);
    const int nb = (int)((nchunks + 1023) / 1024);
    LAUNCH(m, scan_blocks_kernel, nb, 1024, 0, m->smp_counts, nchunks, m->smp_offsets, m->smp_scan, (int*)nullptr);
    LAUNCH(m, scan_blocks_kernel, 1, 1024, 0, m->smp_scan, (long)nb, m->smp_scan + 1024, (int*)nullptr, (int*)nullptr);
    LAUNCH(m, scan_add_kernel, nb, 1024, 0, m->smp_offsets, nchunks, m->smp_scan + 1024, 0L, (const int*)nullptr, (const HeavyWork*)nullptr);
    LAUNCH(m, sampler_total_kernel, 1, 1, 0, m->smp_counts, m->smp_offsets, nchunks);
    LAUNCH(m, sampler_fill_kernel, std::max(grid, (int)((B + 255) / 256)), 256, 0, p);
    m->rng_cur ^= 1;
    return 0;
}

// ------------------------------------------------------------------------------------
// batches
// ------------------------------------------------------------------------------------
// Range-check (and clamp) the ids of a batch on the stream that uploaded them; see validate_ids_kernel.
int validate_ids(nvsm_model* m, cudaStream_t on, idx_t* words, long num_words, idx_t* entities, long num_entities) {
    cudaStream_t main_stream = m->stream;
    m->stream = on;
    const int grid = grid_for(m, num_words + num_entities, 2 * 256, 8);
    int rc = [&]() -> int {
        LAUNCH(m, validate_ids_kernel, grid, 256, 0, words, num_words, m->V, entities, num_entities, m->D, m->id_flags);
        return 0;
    }();
    m->stream = main_stream;
    return rc;
}

// Report (once) that a batch uploaded earlier carried out-of-range ids. Call after a synchronisation that covers
// the upload: the flags live in mapped host memory.
int check_id_flags(nvsm_model* m) {
    if (!m->id_flags) return 0;
    volatile int* f = m->id_flags;
    const int bad_words = f[0], bad_entities = f[1];
    if (!bad_words && !bad_entities) return 0;
    f[0] = 0; f[1] = 0;
    return fail("a batch contained %s outside [0, %ld) (clamped to 0; results of that step are meaningless)",
                bad_words ? (bad_entities ? "word and entity ids" : "word ids") : "entity ids",
                bad_words ? m->V : m->D);
}

// compute_cost without update (cuNVSMTrainModel --compute_initial_cost): the bucket build of that forward pass reads the
// slot's ids on the auxiliary stream and nothing has joined it into the main stream, so a `consumed` event recorded on
// the main stream alone would let the copy stream overwrite the ids between ref_count_kernel and ref_fill_kernel.
int join_unconsumed_buckets(nvsm_model* m) {
    if (m->buckets_in_flight) CU(cudaStreamWaitEvent(m->stream, m->buckets_ready, 0));
    return 0;
}

// Word weights / instance weights of a batch: H2D copy, or -- NULL = uniform weighting, the reference's default
// (feature_weights_ and weights_ are 1.0 unless self-information / idf weighting is selected, include/cuNVSM/data.h:465-467)
// -- a device-side fill with 1.0f that is skipped while the slot still holds ones from an earlier batch. At 8 GPUs the
// host-fed step is bound by the node's aggregate H2D bandwidth (measured r2f: 8 x 6.8 MB per 0.80 ms = 67 GB/s while the
// device-resident step takes 0.67 ms); the two weight arrays are a third of those bytes.
int upload_weights(nvsm_model* m, cudaStream_t cs, float* dst, const float* src, long count, long* ones) {
    if (src) {
        CU(cudaMemcpyAsync(dst, src, sizeof(float) * count, cudaMemcpyHostToDevice, cs));
        *ones = 0;
        return 0;
    }
    if (*ones >= count) return 0;
    cudaStream_t main_stream = m->stream;
    m->stream = cs;
    const int rc = [&]() -> int {
        LAUNCH(m, op_fill_kernel, grid_for(m, count, 1024, 8), 256, 0, dst, count, 1.0f);
        return 0;
    }();
    m->stream = main_stream;
    if (rc == 0) *ones = count;
    return rc;
}

int upload_batch(nvsm_model* m, BatchSlot* s, const long* features, const float* fw, const long* ids,
                 const float* w, long B, bool use_copy_stream) {
    if (B <= 0 || B > m->maxB) return fail("num_instances %ld outside (0, max_batch_size=%ld]", B, m->maxB);
    if (!features || !ids) return fail("null batch pointer");
    cudaStream_t cs = use_copy_stream ? m->copy_stream : m->stream;
    if (cs != m->stream) {
        if (s->in_use) {  // forward without update: order after everything enqueued so far
            TRY(join_unconsumed_buckets(m));
            CU(cudaEventRecord(s->consumed, m->stream));
            s->ever_consumed = true;
        }
        if (s->ever_consumed) CU(cudaStreamWaitEvent(cs, s->consumed, 0));
    }
    s->in_use = false;
    if (cs == m->stream) phase_begin(m, PH_H2D);
    CU(cudaMemcpyAsync(s->features, features, sizeof(long) * B * m->n, cudaMemcpyHostToDevice, cs));
    TRY(upload_weights(m, cs, s->fweights, fw, B * m->n, &s->fw_ones));
    CU(cudaMemcpyAsync(s->ids, ids, sizeof(long) * B * m->R, cudaMemcpyHostToDevice, cs));
    TRY(upload_weights(m, cs, s->weights, w, B, &s->w_ones));
    TRY(validate_ids(m, cs, s->features, B * m->n, s->ids, B * m->R));
    if (cs == m->stream) phase_end(m);
    CU(cudaEventRecord(s->ready, cs));
    s->B = B;
    return 0;
}

BatchSlot* next_live_slot(nvsm_model* m) {
    BatchSlot* s = &m->slots[m->cfg.num_batch_slots + m->next_live];
    m->next_live = (m->next_live + 1) % m->live_slots;
    return s;
}

int ensure_scratch(nvsm_model* m, size_t bytes) {
    if (m->scratch_bytes >= bytes) return 0;
    if (m->scratch) cudaFree(m->scratch);
    m->scratch = nullptr;
    m->scratch_bytes = 0;
    CU(cudaMalloc((void**)&m->scratch, bytes));
    m->scratch_bytes = bytes;
    return 0;
}

struct TensorRef {
    float* ptr = nullptr;
    long count = -1;
    int kind = 0;  // 0 plain, 1 word_projections (materialise), 2 grad_entity_repr (materialise)
};

TensorRef find_tensor(nvsm_model* m, const std::string& s) {
    TensorRef r;
    const long B = m->B;
    const bool fullE = m->cfg.update_method == NVSM_ADAM && m->cfg.adam_mode == NVSM_ADAM_DENSE_UPDATE_DENSE_VARIANCE;
    auto set = [&](float* p, long c) { r.ptr = p; r.count = p ? c : -1; };
    if (s == "word_representations-representations" || s == "W") set(m->W, m->V * m->dw);
    else if (s == "entity_representations-representations" || s == "E") set(m->E, m->D * m->dd);
    else if (s == "word_entity_mapping-transform" || s == "T") set(m->T, (long)m->dw * m->dd);
    else if (s == "word_entity_mapping-bias" || s == "b") set(m->b, m->dd);
    else if (s == "phrase_reprs") { set(m->P, B * m->dw); r.kind = 3; }
    else if (s == "pre_activation") set(m->Z, B * m->dd);
    else if (s == "word_projections") { set(m->Z, B * m->dd); r.kind = 1; }
    else if (s == "similarity_probs") set(m->probs, B * m->R);
    else if (s == "instance_multipliers") set(m->mult, B * m->R);
    else if (s == "grad_transform") set(m->gT, (long)m->dw * m->dd);
    else if (s == "grad_bias") set(m->gb, m->dd);
    else if (s == "grad_phrase_reprs") set(m->gP, B * m->dw);
    else if (s == "grad_projection") set(m->Gp, B * m->dd);
    else if (s == "grad_entity_repr") { set(m->Z, B * m->R * m->dd); r.kind = 2; }
    else if (s == "similarity_pair_probs") set(m->pair_probs, m->pair_N);
    else if (s == "similarity_multipliers") set(m->pair_mult, m->pair_N);
    else if (s == "grad_similarity") set(m->pair_G, 2 * m->pair_N * (m->pair_entities ? m->dd : m->dw));
    else if (s == "bn_mean") set(m->mean, m->dd);
    else if (s == "bn_invstd") set(m->invstd, m->dd);
    else if (s == "word_representations-m") set(m->optW.m, m->V * m->dw);
    else if (s == "word_representations-v") set(m->optW.v, fullE ? m->V * m->dw : m->V);
    else if (s == "word_representations-acc") set(m->optW.acc, m->V);
    else if (s == "entity_representations-m") set(m->optE.m, m->D * m->dd);
    else if (s == "entity_representations-v") set(m->optE.v, fullE ? m->D * m->dd : m->D);
    else if (s == "entity_representations-acc") set(m->optE.acc, m->D);
    else if (s == "word_entity_mapping-transform-m" || s == "word_entity_mapping-transform-acc") set(m->T_a, (long)m->dw * m->dd);
    else if (s == "word_entity_mapping-bias-m" || s == "word_entity_mapping-bias-acc") set(m->b_a, m->dd);
    else if (s == "word_entity_mapping-transform-v") set(m->T_v, (long)m->dw * m->dd);
    else if (s == "word_entity_mapping-bias-v") set(m->b_v, m->dd);
    return r;
}

unsigned long rng_state_of(const std::minstd_rand0& rng) {
    std::ostringstream ss;
    ss << rng;
    return std::stoul(ss.str());
}

}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

const char* nvsm_last_error(void) { return g_error.c_str(); }
int nvsm_version(void) { return 100; }

int nvsm_host_alloc(void** ptr, unsigned long bytes) {
    if (!ptr) return fail("null argument");
    CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

int nvsm_host_free(void* ptr) {
    if (ptr) CU(cudaFreeHost(ptr));
    return 0;
}

int nvsm_num_phases(void) { return PH_COUNT; }
const char* nvsm_phase_name(int phase) { return (phase >= 0 && phase < PH_COUNT) ? kPhaseNames[phase] : ""; }

void nvsm_destroy(nvsm_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->comm) nccl_api().CommDestroy(m->comm);
    for (void* p : m->peer_mapped)
        if (p) cudaIpcCloseMemHandle(p);
    if (m->peer_inbox) cudaFree(m->peer_inbox);
    if (m->peer_flags) cudaFree(m->peer_flags);
    if (m->peer_error) cudaFree(m->peer_error);
    if (m->peer_dev) cudaFree(m->peer_dev);
    if (m->xchg_counter) cudaFree(m->xchg_counter);
    if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
    if (m->gt_stream) cudaStreamDestroy(m->gt_stream);
    if (m->dx_ready) cudaEventDestroy(m->dx_ready);
    if (m->build_gate) cudaEventDestroy(m->build_gate);
    if (m->gt_ready) cudaEventDestroy(m->gt_ready);
    if (m->gt_reduced_ev) cudaEventDestroy(m->gt_reduced_ev);
    float* fl[] = {m->W, m->E, m->T, m->b, m->Tt, m->Tr, m->P_lo, m->Gp_lo, m->Tt_lo, m->Tr_lo, m->optW.m, m->optW.v, m->optW.acc, m->optW.agg, m->optE.m, m->optE.v,
                   m->optE.acc, m->optE.agg, m->T_a, m->b_a, m->T_v, m->b_v, m->P, m->Z, m->Gp, m->gP, m->probs,
                   m->mult, m->rowtmp, m->mean, m->invstd, m->mean_dy, m->mean_dyx, m->bn_scale, m->bn_shift, m->stat_part, m->gT, m->gb, m->gT_part, m->scratch};
    for (float* p : fl)
        if (p) cudaFree(p);
    if (m->dsums) cudaFree(m->dsums);
    int* il[] = {m->e_counts, m->e_offsets, m->e_refs, m->w_counts, m->w_offsets, m->w_refs, m->scan_tmp};
    for (int* p : il)
        if (p) cudaFree(p);
    if (m->Y) cudaFree(m->Y);
    if (m->wcoef) cudaFree(m->wcoef);
    void* ag[] = {m->ag_slot.ids, m->ag_slot.features, m->ag_slot.fweights, m->ag_mult, m->ag_act, m->ag_gP, m->ag_rowtmp,
                  m->ag_e_refs, m->ag_w_refs, m->ag_escore, m->p_norms, m->enorm, m->escore, m->mult_eff, m->kself};
    for (void* p : ag)
        if (p) cudaFree(p);
    m->ag_slot.ids = nullptr; m->ag_slot.features = nullptr; m->ag_slot.fweights = nullptr;
    void* pr[] = {m->pair_ids, m->pair_w, m->pair_probs, m->pair_mult, m->pair_G, m->pair_msq, m->pair_loss};
    for (void* p : pr)
        if (p) cudaFree(p);
    if (m->pair_loss_host) cudaFreeHost(m->pair_loss_host);
    if (m->rng_dev) cudaFree(m->rng_dev);
    free_heavy(&m->heavy_e); free_heavy(&m->heavy_w);
    if (m->heavy_e_dev) cudaFree(m->heavy_e_dev);
    if (m->heavy_w_dev) cudaFree(m->heavy_w_dev);
    if (m->smp_cdf) cudaFree(m->smp_cdf);
    if (m->smp_rej_items) cudaFree(m->smp_rej_items);
    int* sl[] = {m->smp_counts, m->smp_offsets, m->smp_scan, m->smp_error, m->smp_rej_count};
    for (int* p : sl)
        if (p) cudaFree(p);
    for (auto& s : m->slots) {
        if (s.features) cudaFree(s.features);
        if (s.fweights) cudaFree(s.fweights);
        if (s.ids) cudaFree(s.ids);
        if (s.labels) cudaFree(s.labels);
        if (s.weights) cudaFree(s.weights);
        if (s.ready) cudaEventDestroy(s.ready);
        if (s.consumed) cudaEventDestroy(s.consumed);
    }
    for (auto& ev : m->pending) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    for (auto e : m->ev_pool) cudaEventDestroy(e);
    if (m->tl_origin) cudaEventDestroy(m->tl_origin);
    if (m->loss_host) cudaFreeHost(m->loss_host);
    if (m->id_flags) cudaFreeHost(m->id_flags);
    for (auto e : m->loss_ev)
        if (e) cudaEventDestroy(e);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->aux_stream) cudaStreamDestroy(m->aux_stream);
    if (m->score_done) cudaEventDestroy(m->score_done);
    if (m->entity_done) cudaEventDestroy(m->entity_done);
    if (m->word_rows_done) cudaEventDestroy(m->word_rows_done);
    if (m->rowtmp_e) cudaFree(m->rowtmp_e);
    if (m->buckets_ready) cudaEventDestroy(m->buckets_ready);
    if (m->buckets_consumed) cudaEventDestroy(m->buckets_consumed);
    if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
    delete m;
}

int nvsm_create(const nvsm_config* cfg, nvsm_model** out) {
    if (!cfg || !out) return fail("null argument");
    *out = nullptr;
    if (cfg->num_words <= 0 || cfg->num_entities <= 0) return fail("num_words and num_entities must be > 0");
    if (cfg->word_repr_size <= 0 || cfg->entity_repr_size <= 0) return fail("representation sizes must be > 0");
    if (cfg->word_repr_size > 1024 || cfg->entity_repr_size > 1024) return fail("representation sizes above 1024 are not supported");
    if (cfg->entity_repr_size % 4 != 0 && cfg->entity_repr_size > 512) return fail("entity_repr_size must be a multiple of 4 above 512");
    if (cfg->nonlinearity != NVSM_TANH && cfg->nonlinearity != NVSM_HARD_TANH) return fail("nonlinearity %d not implemented.", cfg->nonlinearity);
    if (cfg->update_method < NVSM_SGD || cfg->update_method > NVSM_ADAM) return fail("invalid update_method %d", cfg->update_method);
    if (cfg->update_method == NVSM_ADAM && (cfg->adam_mode < NVSM_ADAM_SPARSE || cfg->adam_mode > NVSM_ADAM_DENSE_UPDATE_DENSE_VARIANCE)) return fail("Invalid mode configuration.");
    if (cfg->num_random_entities < 0) return fail("num_random_entities must be >= 0");
    if (cfg->max_batch_size <= 0 || cfg->window_size <= 0) return fail("max_batch_size and window_size must be > 0");
    if (cfg->gemm_mode < NVSM_GEMM_FP32 || cfg->gemm_mode > NVSM_GEMM_3XTF32) return fail("invalid gemm_mode %d", cfg->gemm_mode);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("no CUDA device: libnvsm_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("device %d out of range (%d devices)", cfg->device, ndev);
    CU(cudaSetDevice(cfg->device));

    nvsm_model* m = new nvsm_model();
    m->cfg = *cfg;
    if (m->cfg.num_batch_slots < 1) m->cfg.num_batch_slots = 1;
    m->device = cfg->device;
    m->V = cfg->num_words; m->D = cfg->num_entities;
    m->dw = cfg->word_repr_size; m->dd = cfg->entity_repr_size;
    m->n = cfg->window_size; m->z = cfg->num_random_entities; m->R = m->z + 1;
    m->maxB = cfg->max_batch_size;
    m->use_tc = cfg->gemm_mode != NVSM_GEMM_FP32 && tc_shapes_ok(m->dw, m->dd);
    m->ldP = m->use_tc ? (m->dw + 31) / 32 * 32 : m->dw;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, m->device);
    m->num_sms = prop.multiProcessorCount;

    auto build = [&]() -> int {
        TRY(set_kernel_attributes());
        CU(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
        m->own_stream = true;
        CU(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&m->gt_stream, cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&m->dx_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->gt_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->gt_reduced_ev, cudaEventDisableTiming));
        {
            auto env_int = [](const char* name, int unset) { const char* e = getenv(name); return e ? atoi(e) : unset; };
            m->knobs.tc_2cta = env_int("NVSM_TC_2CTA", -1); m->knobs.tc_stages = env_int("NVSM_TC_STAGES", 0);
            m->knobs.tc_kb = env_int("NVSM_TC_KB", 0); m->knobs.score_w = env_int("NVSM_SCORE_W", 0);
            m->knobs.score_s = env_int("NVSM_SCORE_S", 0); m->knobs.stats_bps = env_int("NVSM_STATS_BPS", 0);
            m->knobs.sgd_sparse = env_int("NVSM_SGD_SPARSE", -1); m->knobs.no_ring = getenv("NVSM_NO_RING") != nullptr;
            m->knobs.no_fused_stats = getenv("NVSM_NO_FUSED_STATS") != nullptr;
            m->knobs.no_overlap = getenv("NVSM_NO_OVERLAP") != nullptr;
            m->knobs.sampler_scan = getenv("NVSM_SAMPLER_SCAN") != nullptr;   // always the count / scan / fill path
        }
        { const char* e = getenv("NVSM_GT_SIDE"); m->gt_side = e ? atoi(e) != 0 : false; }
        m->no_fused_reduce = getenv("NVSM_NO_FUSED_REDUCE") != nullptr;
        // grad_transform through the NVLink inboxes instead of ncclAllReduce: correct (multi-GPU tests, parity_check) but
        // measured slower at N = 2 (0.656 / 0.662 ms with the push kernel on the side / main stream vs 0.645 ms with
        // ncclAllReduce on the side stream, profiles/bench/r2rst_gt_exchange.md), so it is opt-in: NVSM_FUSED_GT=1.
        { const char* e = getenv("NVSM_FUSED_GT"); m->no_fused_gt = !(e && atoi(e) != 0); }
        { const char* e = getenv("NVSM_GT_PUSH_SIDE"); m->gt_push_side = e && atoi(e) != 0; }
        { const char* e = getenv("NVSM_PDL"); m->pdl = e ? atoi(e) : 0; }
        { const char* e = getenv("NVSM_BUCKETS_AT"); m->buckets_at = e ? std::max(0, std::min(2, atoi(e))) : 0; }
        CU(cudaEventCreateWithFlags(&m->build_gate, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->score_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->entity_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->word_rows_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->buckets_ready, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&m->buckets_consumed, cudaEventDisableTiming));
        const long V = m->V, D = m->D, maxB = m->maxB;
        const int dw = m->dw, dd = m->dd;
        TRY(dev_alloc(&m->W, V * dw)); TRY(dev_alloc(&m->E, D * dd));
        TRY(dev_alloc(&m->T, (size_t)dw * dd)); TRY(dev_alloc(&m->b, dd));
        TRY(dev_alloc(&m->Tt, (size_t)m->ldP * dd)); TRY(dev_alloc(&m->Tr, (size_t)dw * dd));
        if (m->use_tc && cfg->gemm_mode == NVSM_GEMM_3XTF32) {
            TRY(dev_alloc(&m->P_lo, maxB * m->ldP)); TRY(dev_alloc(&m->Gp_lo, maxB * dd));
            TRY(dev_alloc(&m->Tt_lo, (size_t)m->ldP * dd)); TRY(dev_alloc(&m->Tr_lo, (size_t)dw * dd));
        }
        const int method = cfg->update_method;
        if (method == NVSM_ADAGRAD) {
            TRY(dev_alloc(&m->optW.acc, V)); TRY(dev_alloc(&m->optE.acc, D));
            TRY(dev_alloc(&m->T_a, (size_t)dw * dd)); TRY(dev_alloc(&m->b_a, dd));
        } else if (method == NVSM_ADAM) {
            const bool full = cfg->adam_mode == NVSM_ADAM_DENSE_UPDATE_DENSE_VARIANCE;
            TRY(dev_alloc(&m->optW.m, V * dw)); TRY(dev_alloc(&m->optE.m, D * dd));
            TRY(dev_alloc(&m->optW.v, full ? V * dw : V)); TRY(dev_alloc(&m->optE.v, full ? D * dd : D));
            const bool can_pull = full && std::max(V, D) <= 1024L * 1024L && maxB * std::max<long>(m->R, m->n) < (1L << 31) &&
                                  !getenv("NVSM_NO_PULL") &&   // test knob: exercise the scatter + dense full-Adam path
                                  cfg->objective == NVSM_OBJECTIVE_TEXT_ENTITY;   // pair gradients go through the scatter path
            if (full && !can_pull) { TRY(dev_alloc(&m->optW.agg, V * dw)); TRY(dev_alloc(&m->optE.agg, D * dd)); }
            TRY(dev_alloc(&m->T_a, (size_t)dw * dd)); TRY(dev_alloc(&m->b_a, dd));
            TRY(dev_alloc(&m->T_v, (size_t)dw * dd)); TRY(dev_alloc(&m->b_v, dd));
        }
        TRY(dev_alloc(&m->P, maxB * m->ldP)); TRY(dev_alloc(&m->Z, maxB * dd));
        TRY(dev_alloc(&m->Gp, maxB * dd)); TRY(dev_alloc(&m->gP, maxB * dw));
        TRY(dev_alloc(&m->probs, maxB * m->R)); TRY(dev_alloc(&m->mult, maxB * m->R));
        TRY(dev_alloc(&m->rowtmp, maxB)); TRY(dev_alloc(&m->rowtmp_e, maxB));
        {
            const int obj = cfg->objective;
            if (obj < NVSM_OBJECTIVE_TEXT_ENTITY || obj > NVSM_OBJECTIVE_TEXT_ENTITY_TERM_TERM) return fail("unknown objective %d", obj);
            m->has_text = obj == NVSM_OBJECTIVE_TEXT_ENTITY || obj >= NVSM_OBJECTIVE_TEXT_ENTITY_ENTITY_ENTITY;
            m->has_pair = obj != NVSM_OBJECTIVE_TEXT_ENTITY;
            m->pair_entities = obj == NVSM_OBJECTIVE_ENTITY_ENTITY || obj == NVSM_OBJECTIVE_TEXT_ENTITY_ENTITY_ENTITY;
            if (m->has_text && m->has_pair) {
                // TextEntity*::Objective CHECKs both weights != 0 (cpp/objective.cu:709-710,758-759)
                const float wt = cfg->text_entity_weight, ws = cfg->similarity_weight;
                if (wt == 0.f || ws == 0.f) return fail("mixture objectives need non-zero text_entity_weight and similarity_weight");
                m->text_scale = wt / (wt + ws);
                m->pair_scale = ws / (wt + ws);
            }
            if (m->has_pair) {
                m->maxN = cfg->max_similarity_batch_size > 0 ? cfg->max_similarity_batch_size : maxB;
                const int pdim = m->pair_entities ? dd : dw;
                TRY(dev_alloc(&m->pair_ids, 2 * m->maxN)); TRY(dev_alloc(&m->pair_w, m->maxN));
                TRY(dev_alloc(&m->pair_probs, m->maxN)); TRY(dev_alloc(&m->pair_mult, m->maxN));
                TRY(dev_alloc(&m->pair_G, 2 * m->maxN * pdim)); TRY(dev_alloc(&m->pair_msq, 2 * m->maxN));
                TRY(dev_alloc(&m->pair_loss, 1));
                CU(cudaHostAlloc((void**)&m->pair_loss_host, sizeof(double), cudaHostAllocDefault));
                *m->pair_loss_host = 0.0;
            }
        }
        m->l2_phrase = cfg->l2_normalize_phrase_reprs != 0;
        m->l2_entity = cfg->l2_normalize_entity_reprs != 0;
        if (m->l2_phrase) TRY(dev_alloc(&m->p_norms, maxB));
        if (m->l2_entity) {
            TRY(dev_alloc(&m->enorm, maxB * m->R)); TRY(dev_alloc(&m->escore, maxB * m->R));
            TRY(dev_alloc(&m->mult_eff, maxB * m->R)); TRY(dev_alloc(&m->kself, m->D));
        }
        TRY(dev_alloc(&m->mean, dd)); TRY(dev_alloc(&m->invstd, dd));
        TRY(dev_alloc(&m->mean_dy, dd)); TRY(dev_alloc(&m->mean_dyx, dd));
        TRY(dev_alloc(&m->bn_scale, dd)); TRY(dev_alloc(&m->bn_shift, dd));
        TRY(dev_alloc(&m->dsums, 5 * (size_t)dd + 1));
        TRY(dev_alloc(&m->stat_part, (size_t)8 * m->num_sms * 2 * dd));
        TRY(dev_alloc(&m->gT, (size_t)dw * dd)); TRY(dev_alloc(&m->gb, dd));
        const int tiles = ((dw + GEMM_BM - 1) / GEMM_BM) * ((dd + GEMM_BN - 1) / GEMM_BN);
        m->gt_splits = std::max(m->num_sms, (2 * m->num_sms + tiles - 1) / tiles);
        TRY(dev_alloc(&m->gT_part, (size_t)m->gt_splits * dw * dd, false));
        const bool full_adam_pull = method == NVSM_ADAM && cfg->adam_mode == NVSM_ADAM_DENSE_UPDATE_DENSE_VARIANCE &&
                                    m->optE.agg == nullptr /* see can_pull above */;
        // SGD / Adagrad: pull-style as well (no float atomics), for the plain TextEntity objective
        // (one warp per table row: only worth it when there are enough rows to fill the machine; C1's 200-row entity
        // table measured 12 -> 34 us through the pull path)
        const bool sgd_pull = (method == NVSM_SGD || method == NVSM_ADAGRAD) && cfg->objective == NVSM_OBJECTIVE_TEXT_ENTITY &&
                              !cfg->l2_normalize_entity_reprs && !getenv("NVSM_NO_PULL") && std::max(V, D) >= kPullMinRows;
        m->no_heavy = getenv("NVSM_NO_HEAVY") != nullptr;
        m->pull = (full_adam_pull || sgd_pull) &&
                  V < (1L << 30) && D < (1L << 30) && maxB * std::max<long>(m->R, m->n) < (1L << 31);
        if (m->pull) {
            TRY(dev_alloc(&m->Y, maxB * dd));
            TRY(dev_alloc(&m->e_counts, D + 1)); TRY(dev_alloc(&m->e_offsets, D + 1)); TRY(dev_alloc(&m->e_refs, maxB * m->R));
            TRY(dev_alloc(&m->w_counts, V + 1)); TRY(dev_alloc(&m->w_offsets, V + 1)); TRY(dev_alloc(&m->w_refs, maxB * m->n));
            TRY(dev_alloc(&m->scan_tmp, 2 * 1024 + 2));
            if (method == NVSM_ADAGRAD) TRY(dev_alloc(&m->wcoef, maxB * m->n));
            if (std::max(V, D) > 1024L * 1024L) m->pull = false;  // two-level scan limit
        }
        { const char* e = getenv("NVSM_LIVE_SLOTS"); if (e) m->live_slots = std::max(2, std::min(4, atoi(e))); }
        m->slots.resize(m->cfg.num_batch_slots + m->live_slots);
        for (auto& s : m->slots) {
            TRY(dev_alloc(&s.features, maxB * m->n)); TRY(dev_alloc(&s.fweights, maxB * m->n));
            TRY(dev_alloc(&s.ids, maxB * m->R)); TRY(dev_alloc(&s.weights, maxB)); TRY(dev_alloc(&s.labels, maxB));
            CU(cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s.consumed, cudaEventDisableTiming));
        }
        CU(cudaHostAlloc((void**)&m->loss_host, sizeof(double) * nvsm_model::kCostRing, cudaHostAllocMapped | cudaHostAllocPortable));
        TRY(dev_alloc(&m->xchg_counter, 1));
        CU(cudaHostAlloc((void**)&m->id_flags, 2 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
        m->id_flags[0] = m->id_flags[1] = 0;
        for (auto& e : m->loss_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        // score kernel: column-sum staging in dynamic shared memory
        CU(cudaDeviceSynchronize());
        return 0;
    };
    if (build() != 0) {
        std::string keep = g_error;
        nvsm_destroy(m);
        g_error = keep;
        return 1;
    }
    *out = m;
    return 0;
}

int nvsm_set_stream(nvsm_model* m, void* cuda_stream) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(m->stream));
    if (m->own_stream) cudaStreamDestroy(m->stream);
    m->stream = (cudaStream_t)cuda_stream;
    m->own_stream = false;
    return 0;
}

int nvsm_synchronize(nvsm_model* m) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(m->copy_stream));
    CU(cudaStreamSynchronize(m->stream));
    return check_id_flags(m);
}

int nvsm_initialize(nvsm_model* m, unsigned long* rng_state) {
    if (!m || !rng_state) return fail("null argument");
    CU(cudaSetDevice(m->device));
    std::minstd_rand0 rng;
    rng.seed(*rng_state);
    // init_matrix_glorot: rows = feature dim, cols = #objects, linear memory order; W, E, T.
    auto glorot = [&](float* dst, long rows, long cols) -> int {
        const float mx = (float)std::sqrt(6.0 / (double)(rows + cols));
        std::vector<float> h((size_t)rows * cols);
        for (size_t i = 0; i < h.size(); ++i)
            h[i] = 2 * mx * (std::generate_canonical<float, 1>(rng) - 0.5);
        CU(cudaMemcpyAsync(dst, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        CU(cudaStreamSynchronize(m->stream));
        return 0;
    };
    TRY(glorot(m->W, m->dw, m->V));
    TRY(glorot(m->E, m->dd, m->D));
    TRY(glorot(m->T, m->dd, m->dw));
    m->t_copies_stale = true;
    CU(cudaMemsetAsync(m->b, 0, sizeof(float) * m->dd, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    *rng_state = rng_state_of(rng);
    return 0;
}

long nvsm_tensor_size(nvsm_model* m, const char* name) {
    if (!m || !name) return -1;
    return find_tensor(m, name).count;
}

int nvsm_get_tensor(nvsm_model* m, const char* name, float* host_out, long n) {
    if (!m || !name || !host_out) return fail("null argument");
    CU(cudaSetDevice(m->device));
    TensorRef r = find_tensor(m, name);
    if (r.count < 0) return fail("unknown tensor '%s'", name);
    if (r.count != n) return fail("tensor '%s' has %ld elements, caller asked for %ld", name, r.count, n);
    if (r.ptr == m->gT) { TRY(reduce_gt_partials(m)); TRY(join_gt_allreduce(m)); }   // single GPU: the partial sum is otherwise fused into the update
    const float* src = r.ptr;
    if (r.kind == 3) {   // P rows are padded to ldP floats
        CU(cudaMemcpy2DAsync(host_out, sizeof(float) * m->dw, m->P, sizeof(float) * m->ldP, sizeof(float) * m->dw, m->B,
                             cudaMemcpyDeviceToHost, m->stream));
        CU(cudaStreamSynchronize(m->stream));
        if (m->P_lo) {   // 3xTF32: P = hi + lo
            std::vector<float> lo((size_t)n);
            CU(cudaMemcpy2DAsync(lo.data(), sizeof(float) * m->dw, m->P_lo, sizeof(float) * m->ldP, sizeof(float) * m->dw, m->B,
                                 cudaMemcpyDeviceToHost, m->stream));
            CU(cudaStreamSynchronize(m->stream));
            for (long i = 0; i < n; ++i) host_out[i] += lo[i];
        }
        return 0;
    }
    if (r.kind != 0) {
        if (!m->have_forward) return fail("tensor '%s' needs a forward result", name);
        TRY(ensure_scratch(m, sizeof(float) * (size_t)n));
        const ActParams act = act_params(m, m->cfg.batch_normalization != 0);
        const int grid = grid_for(m, n, 256 * 4, 8);
        if (r.kind == 1)
            LAUNCH(m, materialize_activation_kernel, grid, 256, 0, m->Z, act, m->B, m->dd, m->scratch);
        else
            LAUNCH(m, materialize_grad_entity_kernel, grid, 256, 0, m->Z, act, m->mult, m->B * m->R, m->R, m->dd, m->scratch,
                   (const float*)m->E, (const idx_t*)m->cur->ids, m->l2_entity ? (const float*)m->enorm : (const float*)nullptr,
                   m->l2_entity ? (const float*)m->escore : (const float*)nullptr);
        src = m->scratch;
    }
    CU(cudaMemcpyAsync(host_out, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

int nvsm_set_tensor(nvsm_model* m, const char* name, const float* host_in, long n) {
    if (!m || !name || !host_in) return fail("null argument");
    CU(cudaSetDevice(m->device));
    TensorRef r = find_tensor(m, name);
    if (r.count < 0 || r.kind != 0) return fail("unknown or read-only tensor '%s'", name);
    if (r.count != n) return fail("tensor '%s' has %ld elements, caller passed %ld", name, r.count, n);
    CU(cudaMemcpyAsync(r.ptr, host_in, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    m->t_copies_stale = true;   // (cheap to refresh; only T matters)
    return 0;
}

int nvsm_generate_labels(const long* labels, long num_labels, long z, long num_objects, unsigned long* rng_state,
                         long* out) {
    if (!labels || !rng_state || !out) return fail("null argument");
    if (num_labels < 0 || z < 0 || num_objects <= 0) return fail("invalid sampler arguments");
    std::minstd_rand0 rng;
    rng.seed(*rng_state);
    const long R = z + 1;
    for (long i = 0; i < num_labels; ++i) {
        out[i * R] = labels[i];
        for (long k = 1; k <= z; ++k) out[i * R + k] = std::uniform_int_distribution<long>(0, num_objects - 1)(rng);
    }
    *rng_state = rng_state_of(rng);
    return 0;
}

namespace {
int check_cdf(const double* cdf, long n) {
    if (!(cdf[0] >= 0.0)) return fail("cdf[0] must be >= 0");
    for (long k = 1; k < n; ++k)
        if (!(cdf[k] >= cdf[k - 1])) return fail("the cumulative distribution decreases at %ld", k);
    if (cdf[n - 1] != 1.0) return fail("the cumulative distribution must end at exactly 1.0");
    return 0;
}
}  // namespace

int nvsm_generate_labels_cdf(const long* labels, long num_labels, long z, const double* cdf, long num_objects,
                             unsigned long* rng_state, long* out) {
    if (!labels || !rng_state || !out || !cdf) return fail("null argument");
    if (num_labels < 0 || z < 0 || num_objects <= 0) return fail("invalid sampler arguments");
    TRY(check_cdf(cdf, num_objects));
    std::minstd_rand0 rng;
    rng.seed(*rng_state);
    const long R = z + 1;
    for (long i = 0; i < num_labels; ++i) {
        out[i * R] = labels[i];
        for (long k = 1; k <= z; ++k) {
            const double u = (double)(rng() - 1u) / 2147483646.0;
            const long id = std::upper_bound(cdf, cdf + num_objects, u) - cdf;
            out[i * R + k] = std::min(id, num_objects - 1);
        }
    }
    *rng_state = rng_state_of(rng);
    return 0;
}

int nvsm_sampler_set_cdf(nvsm_model* m, const double* cdf, long num_objects) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(m->copy_stream));
    CU(cudaStreamSynchronize(m->stream));
    if (!cdf) {                          // back to the reference's uniform negatives
        if (m->smp_cdf) cudaFree(m->smp_cdf);
        m->smp_cdf = nullptr;
        return 0;
    }
    if (num_objects != m->D) return fail("the distribution has %ld entries, the entity table %ld", num_objects, m->D);
    TRY(check_cdf(cdf, num_objects));
    if (!m->smp_cdf) TRY(dev_alloc(&m->smp_cdf, (size_t)m->D, false));
    CU(cudaMemcpy(m->smp_cdf, cdf, sizeof(double) * m->D, cudaMemcpyHostToDevice));
    return 0;
}

int nvsm_compute_cost(nvsm_model* m, const long* features, const float* fw, const long* ids, const float* w,
                      long B) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    BatchSlot* s = next_live_slot(m);
    TRY(upload_batch(m, s, features, fw, ids, w, B, false));
    return forward(m, s);
}

int nvsm_compute_gradients(nvsm_model* m) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    if (m->has_pair && !m->have_pair_forward) return fail("compute_gradients: the similarity objective has no forward result");
    if (!m->has_text) {   // the pair gradients were written by the fused forward / backward kernel
        m->have_gradients = true;
        return 0;
    }
    return backward(m);
}

int nvsm_similarity_compute_cost(nvsm_model* m, const long* pair_ids, const float* weights, long num_pairs) {
    if (!m || !pair_ids || !weights) return fail("null argument");
    if (!m->has_pair) return fail("this handle has no RepresentationSimilarity objective (nvsm_config.objective)");
    if (num_pairs <= 0 || num_pairs > m->maxN) return fail("num_pairs %ld outside (0, %ld]", num_pairs, m->maxN);
    CU(cudaSetDevice(m->device));
    const long limit = m->pair_entities ? m->D : m->V;
    for (long i = 0; i < 2 * num_pairs; ++i)
        if (pair_ids[i] < 0 || pair_ids[i] >= limit) return fail("pair id %ld out of range at %ld", pair_ids[i], i);
    CU(cudaMemcpyAsync(m->pair_ids, pair_ids, sizeof(idx_t) * 2 * num_pairs, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(m->pair_w, weights, sizeof(float) * num_pairs, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemsetAsync(m->pair_loss, 0, sizeof(double), m->stream));
    m->pair_N = num_pairs;
    PairParams p;
    p.table = m->pair_entities ? m->E : m->W;
    p.dim = m->pair_entities ? m->dd : m->dw;
    p.ids = m->pair_ids; p.w = m->pair_w; p.N = num_pairs;
    const float ef = m->cfg.clip_sigmoid ? 1e-7f : 0.0f, eb = m->cfg.clip_sigmoid ? 1e-6f : 0.0f;
    auto fceil = [](double d) { float f = (float)d; if ((double)f < d) f = std::nextafter(f, INFINITY); return f; };
    auto ffloor = [](double d) { float f = (float)d; if ((double)f > d) f = std::nextafter(f, -INFINITY); return f; };
    p.sig_lo_cmp = fceil((double)ef); p.sig_lo_val = (float)(double)ef;
    p.sig_hi_cmp = ffloor(1.0 - (double)ef); p.sig_hi_val = (float)(1.0 - (double)ef);
    p.der_lo_cmp = ffloor((double)eb); p.der_hi_cmp = fceil(1.0 - (double)eb);
    p.bsn = (float)std::exp(-std::log((double)num_pairs));
    p.scale = m->pair_scale;
    p.probs = m->pair_probs; p.mult = m->pair_mult; p.G = m->pair_G; p.loss_acc = m->pair_loss;
    const int grid = grid_for(m, num_pairs, 8, 8);
    phase_begin(m, PH_SCORE);
    if (vec4_ok(p.dim)) {
        const int nch = (p.dim / 4 + 31) / 32;
        if (nch <= 1) LAUNCH(m, (pair_forward_backward_kernel<4, 1>), grid, 256, 0, p);
        else if (nch <= 2) LAUNCH(m, (pair_forward_backward_kernel<4, 2>), grid, 256, 0, p);
        else if (nch <= 4) LAUNCH(m, (pair_forward_backward_kernel<4, 4>), grid, 256, 0, p);
        else if (nch <= 8) LAUNCH(m, (pair_forward_backward_kernel<4, 8>), grid, 256, 0, p);
        else { phase_end(m); return fail("representation size %d too large for the similarity kernel", p.dim); }
    } else {
        const int nch = (p.dim + 31) / 32;
        if (nch <= 4) LAUNCH(m, (pair_forward_backward_kernel<1, 4>), grid, 256, 0, p);
        else if (nch <= 16) LAUNCH(m, (pair_forward_backward_kernel<1, 16>), grid, 256, 0, p);
        else if (nch <= 32) LAUNCH(m, (pair_forward_backward_kernel<1, 32>), grid, 256, 0, p);
        else { phase_end(m); return fail("representation size %d too large for the similarity kernel", p.dim); }
    }
    phase_end(m);
    CU(cudaMemcpyAsync(m->pair_loss_host, m->pair_loss, sizeof(double), cudaMemcpyDeviceToHost, m->stream));
    m->have_pair_forward = true;
    m->have_gradients = false;
    return 0;
}

int nvsm_similarity_get_cost(nvsm_model* m, float* cost) {
    if (!m || !cost) return fail("null argument");
    if (!m->has_pair || m->pair_N <= 0) return fail("no similarity forward result");
    CU(cudaSetDevice(m->device));
    CU(cudaStreamSynchronize(m->stream));
    *cost = (float)(-(*m->pair_loss_host) / (double)m->pair_N);   // cpp/intermediate_results.cu:80-124
    return 0;
}

float nvsm_similarity_scaled_regularization_lambda(nvsm_model* m) {
    if (!m || m->pair_N <= 0) return 0.f;
    return m->cfg.regularization_lambda / (float)m->pair_N;
}

int nvsm_update(nvsm_model* m, float lr, float lambda) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    return update(m, lr, lambda);
}

int nvsm_read_cost(nvsm_model* m, int steps_back, float* cost) {
    if (!m || !cost) return fail("null argument");
    if (steps_back < 0 || steps_back >= nvsm_model::kCostRing - 1) return fail("steps_back %d out of range", steps_back);
    if (m->forward_count - steps_back <= 0) return fail("get_cost called without a forward result");
    CU(cudaSetDevice(m->device));
    const int slot = (int)((m->forward_count - 1 - steps_back) % nvsm_model::kCostRing);
    CU(cudaEventSynchronize(m->loss_ev[slot]));
    TRY(check_id_flags(m));
    // cost = -(sum_c mass_c) / B   (cpp/intermediate_results.cu:94-120)
    float s = (float)m->loss_host[slot];
    s /= (float)m->loss_B[slot];
    *cost = -s;
    return 0;
}

int nvsm_get_cost(nvsm_model* m, float* cost) { return nvsm_read_cost(m, 0, cost); }

// The same cost without the final float rounding: the loss is accumulated in double on the device, and a central
// difference over a float32 forward pass (gradient checking, cpp/gradient_check.cu:3-133) needs every digit of it.
int nvsm_read_cost_f64(nvsm_model* m, int steps_back, double* cost) {
    if (!m || !cost) return fail("null argument");
    if (steps_back < 0 || steps_back >= nvsm_model::kCostRing - 1) return fail("steps_back %d out of range", steps_back);
    if (m->forward_count - steps_back <= 0) return fail("get_cost called without a forward result");
    CU(cudaSetDevice(m->device));
    const int slot = (int)((m->forward_count - 1 - steps_back) % nvsm_model::kCostRing);
    CU(cudaEventSynchronize(m->loss_ev[slot]));
    TRY(check_id_flags(m));
    *cost = -m->loss_host[slot] / (double)m->loss_B[slot];
    return 0;
}

float nvsm_scaled_regularization_lambda(nvsm_model* m) {
    if (!m || m->Bglobal <= 0) return 0.f;
    return m->cfg.regularization_lambda / (float)m->Bglobal;
}

int nvsm_train_step(nvsm_model* m, const long* features, const float* fw, const long* ids, const float* w, long B,
                    float lr) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    BatchSlot* s = next_live_slot(m);
    TRY(upload_batch(m, s, features, fw, ids, w, B, true));
    return fused_step(m, s, lr);
}

// Device-resident std::minstd_rand0 state for the device sampler.
int nvsm_sampler_seed(nvsm_model* m, unsigned long state) {
    if (!m) return fail("null model");
    CU(cudaSetDevice(m->device));
    TRY(ensure_sampler(m));
    if (m->copy_stream) CU(cudaStreamSynchronize(m->copy_stream));
    unsigned int x = (unsigned int)(state % 2147483647ul);
    if (x == 0) x = 1;   // std::linear_congruential_engine::seed
    CU(cudaMemcpyAsync(m->rng_dev + m->rng_cur, &x, sizeof(x), cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemsetAsync(m->smp_error, 0, sizeof(int), m->stream));
    CU(cudaStreamSynchronize(m->stream));
    m->rng_seeded = true;
    return 0;
}

int nvsm_sampler_state(nvsm_model* m, unsigned long* state) {
    if (!m || !state) return fail("null argument");
    if (!m->rng_seeded) return fail("device sampler is not seeded");
    CU(cudaSetDevice(m->device));
    unsigned int x = 0;
    int err = 0;
    if (m->copy_stream) CU(cudaStreamSynchronize(m->copy_stream));   // nvsm_step_sampled samples there
    CU(cudaMemcpyAsync(&x, m->rng_dev + m->rng_cur, sizeof(x), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaMemcpyAsync(&err, m->smp_error, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    if (err) return fail("device sampler ran out of candidates (rejection rate far above expectation)");
    *state = x;
    return 0;
}

// Upload features / weights / positive labels, draw the negatives on the device, and (train != 0) run
// the whole step. Equivalent to nvsm_generate_labels + nvsm_train_step with the engine state kept on
// the device between steps.
int nvsm_step_sampled(nvsm_model* m, const long* features, const float* fw, const long* labels, const float* w,
                      long B, float lr, int train) {
    if (!m) return fail("null model");
    if (!features || !labels) return fail("null batch pointer");
    if (B <= 0 || B > m->maxB) return fail("num_instances %ld outside (0, max_batch_size=%ld]", B, m->maxB);
    CU(cudaSetDevice(m->device));
    TRY(ensure_sampler(m));
    BatchSlot* s = next_live_slot(m);
    cudaStream_t cs = m->copy_stream;
    if (s->in_use) { TRY(join_unconsumed_buckets(m)); CU(cudaEventRecord(s->consumed, m->stream)); s->ever_consumed = true; }
    if (s->ever_consumed) CU(cudaStreamWaitEvent(cs, s->consumed, 0));
    s->in_use = false;
    cudaEvent_t tl_begin = nullptr;   // timeline mode: the copy-stream work of this batch as an "h2d" interval
    if (m->timeline) { tl_begin = get_event(m); cudaEventRecord(tl_begin, cs); }
    CU(cudaMemcpyAsync(s->features, features, sizeof(long) * B * m->n, cudaMemcpyHostToDevice, cs));
    TRY(upload_weights(m, cs, s->fweights, fw, B * m->n, &s->fw_ones));
    CU(cudaMemcpyAsync(s->labels, labels, sizeof(long) * B, cudaMemcpyHostToDevice, cs));
    TRY(upload_weights(m, cs, s->weights, w, B, &s->w_ones));
    s->B = B;
    TRY(validate_ids(m, cs, s->features, B * m->n, s->labels, B));   // the sampled negatives are in range by construction
    {
        // The negatives of this batch are drawn on the copy stream right behind its H2D copies, i.e. under the
        // previous step's kernels on the main stream (the engine state chains from call to call on that stream).
        cudaStream_t main_stream = m->stream;
        const bool prof = m->profiling;
        m->stream = cs; m->profiling = false;
        const int src = sample_labels_device(m, s->labels, s->ids, s->B, m->z, m->D);
        m->stream = main_stream; m->profiling = prof;
        if (src) return src;
    }
    if (tl_begin) {
        cudaEvent_t tl_end = get_event(m);
        cudaEventRecord(tl_end, cs);
        m->pending.push_back({PH_H2D, tl_begin, tl_end});
    }
    CU(cudaEventRecord(s->ready, cs));   // "ready" covers the copies and the sampled ids (bucket build waits on it)
    CU(cudaStreamWaitEvent(m->stream, s->ready, 0));
    if (!train) return forward(m, s);
    return fused_step(m, s, lr);
}

// Block the host until the H2D copies (and, in device-sampler mode, the sampling) of the running forward result's
// batch have finished: after this the caller may overwrite / recycle the host buffers it passed in.
int nvsm_wait_upload(nvsm_model* m) {
    if (!m) return fail("null model");
    if (!m->cur) return 0;
    CU(cudaSetDevice(m->device));
    CU(cudaEventSynchronize(m->cur->ready));
    return check_id_flags(m);
}

// Stand-alone device sampling for arbitrary (z, num_objects): host labels in, host ids out.
int nvsm_generate_labels_device(nvsm_model* m, const long* labels, long num_labels, long z, long num_objects,
                                unsigned long* rng_state, long* out) {
    if (!m || !labels || !rng_state || !out) return fail("null argument");
    if (num_labels <= 0 || z < 0) return fail("invalid sampler arguments");
    TRY(nvsm_sampler_seed(m, *rng_state));
    idx_t *d_labels = nullptr, *d_ids = nullptr;
    const long R = z + 1;
    CU(cudaMalloc((void**)&d_labels, sizeof(long) * num_labels));
    CU(cudaMalloc((void**)&d_ids, sizeof(long) * num_labels * R));
    auto run = [&]() -> int {
        CU(cudaMemcpyAsync(d_labels, labels, sizeof(long) * num_labels, cudaMemcpyHostToDevice, m->stream));
        TRY(sample_labels_device(m, d_labels, d_ids, num_labels, (int)z, num_objects));
        CU(cudaMemcpyAsync(out, d_ids, sizeof(long) * num_labels * R, cudaMemcpyDeviceToHost, m->stream));
        return nvsm_sampler_state(m, rng_state);
    };
    const int rc = run();
    cudaFree(d_labels); cudaFree(d_ids);
    return rc;
}

// Read back the entity ids (positives + sampled negatives) of the running forward result.
int nvsm_get_entity_ids(nvsm_model* m, long* out, long n) {
    if (!m || !out) return fail("null argument");
    if (!m->cur) return fail("no forward result");
    if (n != m->B * m->R) return fail("entity id count mismatch: have %ld, asked %ld", m->B * m->R, n);
    CU(cudaSetDevice(m->device));
    CU(cudaMemcpyAsync(out, m->cur->ids, sizeof(long) * n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    return 0;
}

int nvsm_stage_batch(nvsm_model* m, int slot, const long* features, const float* fw, const long* ids,
                     const float* w, long B) {
    if (!m) return fail("null model");
    if (slot < 0 || slot >= m->cfg.num_batch_slots) return fail("slot %d out of range [0, %d)", slot, m->cfg.num_batch_slots);
    CU(cudaSetDevice(m->device));
    TRY(upload_batch(m, &m->slots[slot], features, fw, ids, w, B, false));
    CU(cudaStreamSynchronize(m->stream));
    return check_id_flags(m);
}

int nvsm_compute_cost_staged(nvsm_model* m, int slot) {
    if (!m) return fail("null model");
    if (slot < 0 || slot >= m->cfg.num_batch_slots) return fail("slot %d out of range", slot);
    CU(cudaSetDevice(m->device));
    return forward(m, &m->slots[slot]);
}

int nvsm_train_step_staged(nvsm_model* m, int slot, float lr) {
    if (!m) return fail("null model");
    if (slot < 0 || slot >= m->cfg.num_batch_slots) return fail("slot %d out of range", slot);
    CU(cudaSetDevice(m->device));
    return fused_step(m, &m->slots[slot], lr);
}

int nvsm_infer(nvsm_model* m, const long* words, long N, long window, float* out) {
    if (!m || !words || !out) return fail("null argument");
    if (N <= 0 || window <= 0) return fail("empty inference request");
    CU(cudaSetDevice(m->device));
    // Model::infer: gather-mean (no word weights), projection + bias, activation, no batch-norm.
    idx_t* d_words = nullptr;
    float *d_p = nullptr, *d_z = nullptr;
    CU(cudaMalloc((void**)&d_words, sizeof(long) * N * window));
    CU(cudaMalloc((void**)&d_p, sizeof(float) * N * m->dw));
    CU(cudaMalloc((void**)&d_z, sizeof(float) * N * m->dd));
    auto run = [&]() -> int {
        CU(cudaMemcpyAsync(d_words, words, sizeof(long) * N * window, cudaMemcpyHostToDevice, m->stream));
        const int grid = grid_for(m, N, 8, 8);
        if (vec4_ok(m->dw))
            LAUNCH(m, gather_mean_kernel<4>, grid, 256, 0, m->W, m->dw, d_words, (const float*)nullptr, N, (int)window, d_p, m->dw, 0, (float*)nullptr);
        else
            LAUNCH(m, gather_mean_kernel<1>, grid, 256, 0, m->W, m->dw, d_words, (const float*)nullptr, N, (int)window, d_p, m->dw, 0, (float*)nullptr);
        TRY((run_sgemm<false, false>(m, (int)N, m->dd, m->dw, d_p, m->dw, m->T, m->dd, d_z, m->dd, 1, 1.0f, m->b)));
        LAUNCH(m, materialize_activation_kernel, grid_for(m, N * m->dd, 1024, 8), 256, 0, d_z, act_params(m, false), N, m->dd, d_z);
        CU(cudaMemcpyAsync(out, d_z, sizeof(float) * N * m->dd, cudaMemcpyDeviceToHost, m->stream));
        CU(cudaStreamSynchronize(m->stream));
        return 0;
    };
    const int rc = run();
    cudaFree(d_words); cudaFree(d_p); cudaFree(d_z);
    return rc;
}

int nvsm_increment_parameter(nvsm_model* m, const char* name, long idx, float epsilon) {
    if (!m || !name) return fail("null argument");
    CU(cudaSetDevice(m->device));
    TensorRef r = find_tensor(m, name);
    if (r.count < 0 || r.kind != 0) return fail("unknown tensor '%s'", name);
    if (idx < 0 || idx >= r.count) return fail("index %ld out of range for '%s'", idx, name);
    LAUNCH(m, increment_kernel, 1, 1, 0, r.ptr + idx, epsilon);
    m->t_copies_stale = true;
    return 0;
}

int nvsm_set_profiling(nvsm_model* m, int enabled) {
    if (!m) return fail("null model");
    TRY(collect_phases(m));
    m->profiling = enabled == 1;
    m->timeline = enabled == 2;
    if (m->timeline) {
        CU(cudaSetDevice(m->device));
        if (!m->tl_origin) CU(cudaEventCreate(&m->tl_origin));
        m->intervals.clear();
        CU(cudaEventRecord(m->tl_origin, m->stream));
    }
    return 0;
}

int nvsm_get_timeline(nvsm_model* m, int* phases, float* start_ms, float* end_ms, int capacity) {
    // (a count, not a status: errors are -1 with the message in nvsm_last_error)
    if (!m) { fail("null model"); return -1; }
    if (cudaSetDevice(m->device) != cudaSuccess) { fail("cudaSetDevice(%d) failed", m->device); return -1; }
    if (collect_phases(m) != 0) return -1;
    const int n = (int)m->intervals.size();
    for (int i = 0; i < n && i < capacity; ++i) {
        if (phases) phases[i] = m->intervals[i].phase;
        if (start_ms) start_ms[i] = m->intervals[i].start_ms;
        if (end_ms) end_ms[i] = m->intervals[i].end_ms;
    }
    return n;
}

int nvsm_get_phase_ms(nvsm_model* m, float* ms_out, int capacity) {
    if (!m || !ms_out) return fail("null argument");
    CU(cudaSetDevice(m->device));
    TRY(collect_phases(m));
    for (int i = 0; i < capacity && i < PH_COUNT; ++i) ms_out[i] = (float)m->phase_ms[i];
    return 0;
}

int nvsm_reset_phase_ms(nvsm_model* m) {
    if (!m) return fail("null model");
    TRY(collect_phases(m));
    for (int i = 0; i < PH_COUNT; ++i) m->phase_ms[i] = 0.0;
    return 0;
}

long nvsm_kernel_launches(nvsm_model* m) { return m ? m->launches : 0; }

// Test hook for the tensor-core GEMM in isolation (tests/test_gpu_gemm.py): host in / host out.
// variant 0: A[M,K], B[N,K] (K-major); variant 1: A[K,M], B[K,N] (MN-major, split-K + reduce).
int nvsm_test_gemm_tc(nvsm_model* m, int variant, int M, int N, int K, const float* A, const float* Bh,
                      float* C, float alpha, const float* bias, int splits) {
    if (!m || !A || !Bh || !C) return fail("null argument");
    CU(cudaSetDevice(m->device));
    float *dA = nullptr, *dB = nullptr, *dC = nullptr, *dP = nullptr, *dbias = nullptr;
    const size_t na = (size_t)M * K, nb = (size_t)N * K, nc = (size_t)M * N;
    auto run = [&]() -> int {
        CU(cudaMalloc((void**)&dA, na * 4)); CU(cudaMalloc((void**)&dB, nb * 4)); CU(cudaMalloc((void**)&dC, nc * 4));
        CU(cudaMemcpy(dA, A, na * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(dB, Bh, nb * 4, cudaMemcpyHostToDevice));
        if (bias) { CU(cudaMalloc((void**)&dbias, (size_t)N * 4)); CU(cudaMemcpy(dbias, bias, (size_t)N * 4, cudaMemcpyHostToDevice)); }
        float *dAlo = nullptr, *dBlo = nullptr;
        const bool split3 = variant >= 2;
        variant &= 1;
        if (split3) {
            CU(cudaMalloc((void**)&dAlo, na * 4)); CU(cudaMalloc((void**)&dBlo, nb * 4));
            LAUNCH(m, split_tf32_kernel, 512, 256, 0, dA, dAlo, (long)na);
            LAUNCH(m, split_tf32_kernel, 512, 256, 0, dB, dBlo, (long)nb);
        }
        struct Free { float *a, *b; ~Free() { cudaFree(a); cudaFree(b); } } fr{dAlo, dBlo};
        if (variant == 0) {
            TRY(run_gemm_tc(m, false, M, N, K, dA, K, dB, K, dC, N, 1, 0, alpha, dbias, nullptr, dAlo, dBlo));
        } else {
            splits = std::max(1, splits);
            CU(cudaMalloc((void**)&dP, nc * 4 * splits));
            int nparts = 1;
            TRY(run_gemm_tc(m, true, M, N, K, dA, M, dB, N, dP, N, splits, (long)nc, alpha, nullptr, &nparts, dAlo, dBlo));
            LAUNCH(m, reduce_partials_kernel, (int)((nc + 255) / 256), 256, 0, dP, nparts, (long)nc, dC);
        }
        CU(cudaStreamSynchronize(m->stream));
        CU(cudaMemcpy(C, dC, nc * 4, cudaMemcpyDeviceToHost));
        return 0;
    };
    const int rc = run();
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dP); cudaFree(dbias);
    return rc;
}

// Device-resident timing of the tensor-core GEMM alone (scripts/bench_gemm.py).
int nvsm_bench_gemm_tc(nvsm_model* m, int variant, int M, int N, int K, int splits, int with_stats, int iters,
                       float* ms_out) {
    // with_stats: bit 0 = fused column statistics in the epilogue, bit 1 = 3xTF32 (hi / lo operand pairs)
    if (!m || !ms_out) return fail("null argument");
    CU(cudaSetDevice(m->device));
    float *dA = nullptr, *dB = nullptr, *dC = nullptr, *dAlo = nullptr, *dBlo = nullptr, *dS = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool stats = (with_stats & 1) != 0, split3 = (with_stats & 2) != 0;
    // K-major operands keep the step's 128-byte-aligned row stride (d_w = 300 -> 320 floats)
    const int ldk = (K + 31) / 32 * 32;
    const size_t na = variant ? (size_t)K * M : (size_t)M * ldk, nb = variant ? (size_t)K * N : (size_t)N * ldk;
    auto run = [&]() -> int {
        CU(cudaMalloc((void**)&dA, na * 4)); CU(cudaMalloc((void**)&dB, nb * 4));
        CU(cudaMalloc((void**)&dC, (size_t)M * N * 4 * std::max(1, splits)));
        CU(cudaMemset(dA, 0, na * 4)); CU(cudaMemset(dB, 0, nb * 4));
        if (split3) {
            CU(cudaMalloc((void**)&dAlo, na * 4)); CU(cudaMalloc((void**)&dBlo, nb * 4));
            CU(cudaMemset(dAlo, 0, na * 4)); CU(cudaMemset(dBlo, 0, nb * 4));
        }
        if (stats) CU(cudaMalloc((void**)&dS, (size_t)8 * m->num_sms * 2 * N * 4));
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        for (int it = 0; it < iters + 3; ++it) {
            if (it == 3) CU(cudaEventRecord(e0, m->stream));
            int rows = 0;
            TRY(run_gemm_tc(m, variant != 0, M, N, K, dA, variant ? M : ldk, dB, variant ? N : ldk, dC, N, splits, (long)M * N, 1.0f,
                            nullptr, nullptr, dAlo, dBlo, dS, &rows));
            if (stats && rows == 0) return fail("this GEMM shape cannot fuse the column statistics");
        }
        CU(cudaEventRecord(e1, m->stream));
        CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(ms_out, e0, e1));
        *ms_out /= iters;
        return 0;
    };
    const int rc = run();
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dS); cudaFree(dAlo); cudaFree(dBlo);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

// Roofline denominators measured in place (microbench.cuh). kind 0 / 3 / 4: gather of `rows_per_item` pseudo-random
// rows of `row_floats` floats per item out of a table of `table_bytes` (L2-resident when it fits the 126 MB L2) with
// 2 / 4 / 1 rows in flight per warp, GB/s of gathered bytes; kind 1: streaming copy of `table_bytes` (read + write
// bytes), GB/s; kind 2: L2-resident streaming read (`items` sweeps over a `table_bytes` buffer per launch), GB/s.
int nvsm_bench_memory(nvsm_model* m, int kind, long table_bytes, int row_floats, int rows_per_item, long items, int iters,
                      float* gbs_out) {
    if (!m || !gbs_out) return fail("null argument");
    if (iters <= 0 || table_bytes <= 0) return fail("invalid probe arguments");
    CU(cudaSetDevice(m->device));
    float4 *table = nullptr, *out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto run = [&]() -> int {
        CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
        float ms = 0.f;
        if (kind == 1) {
            const long n = table_bytes / 16;
            CU(cudaMalloc((void**)&table, n * 16)); CU(cudaMalloc((void**)&out, n * 16));
            CU(cudaMemsetAsync(table, 0, n * 16, m->stream));
            const int grid = m->num_sms * 8;
            for (int it = 0; it < iters + 2; ++it) {
                if (it == 2) CU(cudaEventRecord(e0, m->stream));
                LAUNCH(m, stream_copy_probe_kernel, grid, 256, 0, (const float4*)table, out, n);
            }
            CU(cudaEventRecord(e1, m->stream)); CU(cudaEventSynchronize(e1));
            CU(cudaEventElapsedTime(&ms, e0, e1));
            *gbs_out = (float)(2.0 * n * 16 * iters / (ms * 1e-3) / 1e9);
            return 0;
        }
        if (kind == 2) {   // L2-resident streaming read: `items` = sweeps over the buffer per launch
            const long n = table_bytes / 16;
            const int reps = (int)std::max<long>(1, items);
            CU(cudaMalloc((void**)&table, n * 16)); CU(cudaMalloc((void**)&out, (size_t)m->num_sms * 8 * 256 * 16));
            CU(cudaMemsetAsync(table, 0, n * 16, m->stream));
            for (int it = 0; it < iters + 2; ++it) {
                if (it == 2) CU(cudaEventRecord(e0, m->stream));
                LAUNCH(m, l2_read_probe_kernel, m->num_sms * 8, 256, 0, (const float4*)table, n, reps, out);
            }
            CU(cudaEventRecord(e1, m->stream)); CU(cudaEventSynchronize(e1));
            CU(cudaEventElapsedTime(&ms, e0, e1));
            *gbs_out = (float)((double)n * 16 * reps * iters / (ms * 1e-3) / 1e9);
            return 0;
        }
        if (kind != 0 && kind != 3 && kind != 4) return fail("unknown probe kind %d", kind);
        if (row_floats <= 0 || row_floats % 4 != 0 || row_floats > 512 || rows_per_item <= 0 || items <= 0)
            return fail("gather probe: row_floats must be a multiple of 4 up to 512");
        const int row_vec4 = row_floats / 4;
        const long num_rows = std::max<long>(1, table_bytes / (16L * row_vec4));
        CU(cudaMalloc((void**)&table, num_rows * row_vec4 * 16)); CU(cudaMalloc((void**)&out, items * row_vec4 * 16));
        CU(cudaMemsetAsync(table, 0, num_rows * row_vec4 * 16, m->stream));
        const int grid = grid_for(m, items, 8, 8);
        const int K = (row_vec4 + 31) / 32;
        for (int it = 0; it < iters + 2; ++it) {
            if (it == 2) CU(cudaEventRecord(e0, m->stream));
            const unsigned seed = 0x1234567u + (unsigned)it * 7919u;
#define NVSM_PROBE(K_, U_) LAUNCH(m, (l2_gather_probe_kernel<K_, U_>), grid, 256, 0, (const float4*)table, num_rows, row_vec4, rows_per_item, items, out, seed)
#define NVSM_PROBE_K(U_) do { if (K <= 1) NVSM_PROBE(1, U_); else if (K == 2) NVSM_PROBE(2, U_); else if (K == 3) NVSM_PROBE(3, U_); else NVSM_PROBE(4, U_); } while (0)
            if (kind == 0) NVSM_PROBE_K(2); else if (kind == 3) NVSM_PROBE_K(4); else NVSM_PROBE_K(1);   // rows in flight per warp
#undef NVSM_PROBE_K
#undef NVSM_PROBE
        }
        CU(cudaEventRecord(e1, m->stream)); CU(cudaEventSynchronize(e1));
        CU(cudaEventElapsedTime(&ms, e0, e1));
        *gbs_out = (float)((double)items * rows_per_item * row_floats * 4.0 * iters / (ms * 1e-3) / 1e9);
        return 0;
    };
    const int rc = run();
    cudaFree(table); cudaFree(out);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

int nvsm_comm_unique_id(char* id_out_128) {
    if (!id_out_128) return fail("null argument");
    const char* why = "";
    if (!nccl_api().load(&why)) return fail("NCCL unavailable: %s", why);
    NcclUniqueId id;
    const int rc = nccl_api().GetUniqueId(&id);
    if (rc != 0) return fail("ncclGetUniqueId: %s", nccl_api().GetErrorString(rc));
    memcpy(id_out_128, id.internal, 128);
    return 0;
}

int nvsm_comm_init(nvsm_model* m, const char* id_128, int num_ranks, int rank) {
    if (!m || !id_128) return fail("null argument");
    if (num_ranks < 1 || rank < 0 || rank >= num_ranks) return fail("invalid rank %d of %d", rank, num_ranks);
    if (num_ranks == 1) { m->nranks = 1; m->rank = 0; return 0; }
    const char* why = "";
    if (!nccl_api().load(&why)) return fail("NCCL unavailable: %s", why);
    CU(cudaSetDevice(m->device));
    NcclUniqueId id;
    memcpy(id.internal, id_128, 128);
    const int rc = nccl_api().CommInitRank(&m->comm, num_ranks, id, rank);
    if (rc != 0) return fail("ncclCommInitRank: %s", nccl_api().GetErrorString(rc));
    m->nranks = num_ranks;
    m->rank = rank;
    if (!m->comm_stream) CU(cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking));
    return 0;
}

// NVLink peer exchange for the small per-step reductions (peer_allreduce.cuh). Optional: without it they go through
// ncclAllReduce. Protocol: after nvsm_comm_init every rank calls nvsm_comm_peer_export (128 bytes: two CUDA IPC memory
// handles), the caller all-gathers the blobs (rank order) and hands them to nvsm_comm_peer_import on every rank.
int nvsm_comm_peer_export(nvsm_model* m, char* handles_out_128) {
    if (!m || !handles_out_128) return fail("null argument");
    if (m->nranks < 2 || m->nranks > kPeerMaxRanks) return fail("peer exchange needs 2..%d ranks", kPeerMaxRanks);
    CU(cudaSetDevice(m->device));
    if (!m->peer_inbox) {
        m->peer.nranks = m->nranks; m->peer.rank = m->rank;
        m->peer.slot_doubles = 2 * m->dd + 8;
        const size_t nslots = (size_t)kPeerKinds * 2 * m->nranks;
        m->peer.gt_elems = (long)m->dw * m->dd;
        // one allocation (one IPC handle): the double slots, then the grad_transform inbox [2][nranks][dw * dd] floats
        const size_t gt_doubles = ((size_t)2 * m->nranks * m->peer.gt_elems + 1) / 2;
        TRY(dev_alloc(&m->peer_inbox, nslots * m->peer.slot_doubles + gt_doubles));
        TRY(dev_alloc(&m->peer_flags, nslots * kPeerFlagStride));
        TRY(dev_alloc(&m->peer_error, 1));
        TRY(dev_alloc(&m->peer_dev, 1));
        m->no_fused_xchg = getenv("NVSM_NO_FUSED_XCHG") != nullptr;
        CU(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h[2];
    CU(cudaIpcGetMemHandle(&h[0], m->peer_inbox));
    CU(cudaIpcGetMemHandle(&h[1], m->peer_flags));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    memcpy(handles_out_128, h, 128);
    return 0;
}

int nvsm_comm_peer_import(nvsm_model* m, const char* all_handles) {
    if (!m || !all_handles) return fail("null argument");
    if (!m->peer_inbox) return fail("nvsm_comm_peer_import before nvsm_comm_peer_export");
    CU(cudaSetDevice(m->device));
    for (int p = 0; p < m->nranks; ++p) {
        if (p == m->rank) {
            m->peer.inbox[p] = m->peer_inbox;
            m->peer.flags[p] = m->peer_flags;
            m->peer.gt_inbox[p] = (float*)(m->peer_inbox + (size_t)kPeerKinds * 2 * m->nranks * m->peer.slot_doubles);
            continue;
        }
        cudaIpcMemHandle_t h[2];
        memcpy(h, all_handles + (size_t)p * 128, 128);
        void *inbox = nullptr, *flags = nullptr;
        CU(cudaIpcOpenMemHandle(&inbox, h[0], cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle(&flags, h[1], cudaIpcMemLazyEnablePeerAccess));
        m->peer_mapped[2 * p] = inbox; m->peer_mapped[2 * p + 1] = flags;
        m->peer.inbox[p] = (double*)inbox;
        m->peer.flags[p] = (unsigned long long*)flags;
        m->peer.gt_inbox[p] = (float*)((double*)inbox + (size_t)kPeerKinds * 2 * m->nranks * m->peer.slot_doubles);
    }
    CU(cudaMemcpy(m->peer_dev, &m->peer, sizeof(PeerXchg), cudaMemcpyHostToDevice));
    m->peer_ready = true;
    return 0;
}

int nvsm_comm_peer_disable(nvsm_model* m) {
    if (!m) return fail("null argument");
    m->peer_ready = false;
    return 0;
}

int nvsm_comm_peer_status(nvsm_model* m, int* ready, int* error) {
    if (!m || !ready || !error) return fail("null argument");
    *ready = m->peer_ready ? 1 : 0;
    *error = 0;
    if (m->peer_error) {
        CU(cudaSetDevice(m->device));
        CU(cudaStreamSynchronize(m->stream));
        CU(cudaMemcpy(error, m->peer_error, sizeof(int), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int nvsm_comm_set_sparse_mode(nvsm_model* m, int mode) {
    if (!m) return fail("null argument");
    if (mode != NVSM_SPARSE_LOCAL && mode != NVSM_SPARSE_ALLGATHER) return fail("unknown sparse mode %d", mode);
    if (mode == NVSM_SPARSE_ALLGATHER && m->has_pair) return fail("NVSM_SPARSE_ALLGATHER is implemented for the TextEntity objective only");
    if (mode == NVSM_SPARSE_ALLGATHER && m->nranks > 1 && !m->ag_mult) {
        CU(cudaSetDevice(m->device));
        const size_t G = (size_t)m->maxB * m->nranks;
        TRY(dev_alloc(&m->ag_slot.ids, G * m->R)); TRY(dev_alloc(&m->ag_slot.features, G * m->n));
        TRY(dev_alloc(&m->ag_slot.fweights, G * m->n));
        TRY(dev_alloc(&m->ag_mult, G * m->R)); TRY(dev_alloc(&m->ag_act, G * m->dd)); TRY(dev_alloc(&m->ag_gP, G * m->dw));
        TRY(dev_alloc(&m->ag_rowtmp, G));
        if (m->pull) { TRY(dev_alloc(&m->ag_e_refs, G * m->R)); TRY(dev_alloc(&m->ag_w_refs, G * m->n)); }
        if (m->l2_entity) TRY(dev_alloc(&m->ag_escore, G * m->R));
    }
    m->sparse_mode = mode;
    return 0;
}

#include "ops_host.inc"

}  // extern "C"
