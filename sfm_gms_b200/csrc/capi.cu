// capi.cu — the C ABI (include/sfmgms.h) over the CUDA kernels: context, buffers, batching.
// No CPU fallback exists anywhere in this file: every entry point either runs the sm_100a kernels or
// returns an error code.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <algorithm>
#include <vector>

#include "../../include/sfmgms.h"
#include "capi_internal.h"
#include "common.cuh"
#include "hamming_tc.cuh"

using namespace sfmgms;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {  // pinned staging
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

char g_create_error[512] = "";

}  // namespace

namespace sfmgms { thread_local KernelMarks* tl_marks = nullptr; }

struct sfmgms_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    char err[512] = "";
    int64_t launches = 0;
    int hamming_kernel = SFMGMS_HAMMING_AUTO;
    size_t gms_chunk_bytes = 64ull << 20;
    int timing = 0;
    int l2_kernel = 0;   // 0 auto (tcgen05, fp32 fallback), 1 dp4a, 2 tcgen05, 3 fp32 order-exact
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;   // pipelined host<->device copies (match_image_set)
    cudaStream_t side_stream = nullptr;                        // small kernels that run beside the next tensor-core launch
    std::vector<cudaEvent_t> side_events;
    std::vector<cudaEvent_t> events;                           // pool of timing-disabled events
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    double last_ms[3] = {0, 0, 0};

    // generic device buffers
    DevBuf d_pairs, d_results, d_key, d_mask, d_hist, d_msc, d_q, d_kp1, d_kp2, d_mq, d_mt, d_out_i32, d_pts;
    DevBuf d_set_desc, d_set_kp;
    HostBuf h_stage, h_pairs_pinned, h_results;
    TcState tc;   // tensor-core Hamming operand cache (hamming_tc.cu)
    OrbWorkspace* orb = nullptr;   // ORB pyramid + scratch (orb.cu), created on first use
    std::vector<uint8_t> orb_keypoints;   // cv::KeyPoint records of the last sfmgms_set_images_from_pixels
    bool keep_orb_keypoints = false;      // set only around that function's own sfmgms_set_images call
    std::vector<OrbWorkspace*> orb_pool;  // per-worker workspaces + streams of sfmgms_set_images_from_pixels
    std::vector<cudaStream_t> orb_streams;

    // image set
    int n_images = 0;
    std::vector<int64_t> offsets;
    std::vector<int32_t> sizes;
    const uint8_t* set_desc = nullptr;   // device
    const float* set_kp = nullptr;       // device
    uint64_t set_version = 0;

    // last batch (for sfmgms_inlier_points)
    std::vector<PairDesc> last_pairs;
    std::vector<PairResult> last_results;

    // chunked pair-list runs (sfmgms_match_pairs / _compact): two slots of per-chunk device buffers so that a chunk
    // computes while the previous one's results travel to the host; device scratch is O(chunk), not O(list)
    struct KTime { std::string name; double ms = 0; long long n = 0; };
    std::vector<KTime> kernel_times;       // SFMGMS_OPT_TIMING = 2: accumulated per-kernel device time
    KernelMarks marks;                     // single-batch paths
    void absorb(KernelMarks& m) {
        for (int i = 1; i < m.used; ++i) {
            float dt = 0.f;
            if (cudaEventElapsedTime(&dt, m.ev[i - 1], m.ev[i]) != cudaSuccess) continue;
            KTime* k = nullptr;
            for (auto& e : kernel_times) if (e.name == m.name[i]) k = &e;
            if (!k) { kernel_times.push_back(KTime{m.name[i], 0, 0}); k = &kernel_times.back(); }
            k->ms += dt; k->n++;
        }
        m.used = 0;
    }
    struct Slot {
        KernelMarks marks;
        DevBuf key, mask, out_i32, cmatch, cpts, coff;
        cudaEvent_t computed = nullptr, copied = nullptr, t0 = nullptr, t1 = nullptr, t2 = nullptr;
        bool copy_pending = false;
    } slot[2];
    DevBuf d_cbase;                 // running inlier total of a compact run (int64)
    HostBuf h_coff;
    struct Pending { bool active = false; int n_pairs = 0; bool compact = false; long long capacity = 0; } pending;   // *_async
    long long chunk_rows = 4ll << 20;   // match rows per chunk (SFMGMS_OPT_CHUNK_ROWS)
    int gms_dense = 0;              // SFMGMS_OPT_GMS_DENSE
    int overlap_resolve = 1;        // SFMGMS_OPT_OVERLAP (0: everything on one stream)
    int compact_rec_bytes = 16;     // SFMGMS_OPT_COMPACT_RECORD: 16 = cv::DMatch records, 8 = {queryIdx, trainIdx}
};

namespace {

int fail(sfmgms_ctx* c, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (c) vsnprintf(c->err, sizeof c->err, fmt, ap);
    else vsnprintf(g_create_error, sizeof g_create_error, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, SFMGMS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                       \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

__global__ void decode_keys_kernel(const uint32_t* __restrict__ key, long long n, int32_t* __restrict__ train_idx,
                                   int32_t* __restrict__ dist) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += step) {
        const uint32_t k = key[i];
        const bool none = (k == kKeyInit);
        if (train_idx) train_idx[i] = none ? -1 : (int32_t)(k & kTrainIdxMask);
        if (dist) dist[i] = none ? -1 : (int32_t)(k >> kTrainIdxBits);
    }
}

// per-pair summaries of a batch for device callers (the asynchronous pair-list calls cannot pass through the host)
__global__ void export_results_kernel(const PairResult* __restrict__ res, int n, int32_t* __restrict__ n_inliers,
                                      int32_t* __restrict__ best_hyp, int32_t* __restrict__ mask_len) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const PairResult r = res[p];
    if (n_inliers) n_inliers[p] = r.n_inliers;
    if (best_hyp) best_hyp[p] = r.best_hyp;
    if (mask_len) mask_len[p] = r.mask_len;
}

// cross-check: for each train row j the nearest query (lowest query index on ties) is rkey[j];
// keep[i] = (rkey[train(i)] & mask) == i
__global__ void crosscheck_kernel(const uint32_t* __restrict__ key, const uint32_t* __restrict__ rkey, int nq,
                                  uint8_t* __restrict__ keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint32_t k = key[i];
    keep[i] = (k != kKeyInit) && ((rkey[k & kTrainIdxMask] & kTrainIdxMask) == (uint32_t)i);
}

// cross-check on plain index arrays (L2 path): keep[i] iff the nearest query of train row fwd[i] is i
__global__ void crosscheck_idx_kernel(const int32_t* __restrict__ fwd, const int32_t* __restrict__ rev, int nq,
                                      uint8_t* __restrict__ keep) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const int j = fwd[i];
    keep[i] = j >= 0 && rev[j] == i;
}

int choose_hamming(const sfmgms_ctx* c) {
    if (c->hamming_kernel == SFMGMS_HAMMING_AUTO) return SFMGMS_HAMMING_FP4;   // fastest measured (profiles/r1_notes.md)
    return c->hamming_kernel;
}

// Copies the pair table of a whole batch to the device (async on the context stream).
int upload_pairs(sfmgms_ctx* ctx, const std::vector<PairDesc>& hp) {
    const size_t n = hp.size();
    if (n == 0) return SFMGMS_OK;
    CU(ctx->d_pairs.ensure(sizeof(PairDesc) * n));
    CU(ctx->h_pairs_pinned.ensure(sizeof(PairDesc) * n));
    CU(ctx->d_results.ensure(sizeof(PairResult) * n));
    memcpy(ctx->h_pairs_pinned.p, hp.data(), sizeof(PairDesc) * n);
    CU(cudaMemcpyAsync(ctx->d_pairs.p, ctx->h_pairs_pinned.p, sizeof(PairDesc) * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(ctx->d_results.p, 0, sizeof(PairResult) * n, ctx->stream));
    return SFMGMS_OK;
}

// Enqueues Hamming (optional) + GMS (optional) for pairs [r0, r0+rn) of an uploaded batch.  Asynchronous: nothing
// here waits for the device.  Scratch reuse across ranges is ordered by the stream.
int enqueue_range(sfmgms_ctx* ctx, const std::vector<PairDesc>& hp, int r0, int rn, bool do_hamming, bool do_gms,
                  int with_rotation, int with_scale, double factor, int* ham_launches, cudaEvent_t mid = nullptr) {
    if (rn <= 0) { if (mid) CU(cudaEventRecord(mid, ctx->stream)); return SFMGMS_OK; }
    cudaStream_t st = ctx->stream;
    const PairDesc* dp = static_cast<const PairDesc*>(ctx->d_pairs.p) + r0;
    const PairDesc* hpp = hp.data() + r0;
    if (hpp[0].img1 < 0) tc_invalidate(ctx->tc);   // ad-hoc buffers: contents change between calls
    if (do_hamming) {
        const int kind = choose_hamming(ctx);
        int kMaxPerLaunch = (kind == SFMGMS_HAMMING_FP4) ? 768 : 32768;   // fp4: launch map is a kernel parameter
        // fp4, operands of a registered set resident: a larger batch goes out as up to three tensor-core launches, so that
        // the tie-resolution kernel of one launch runs on the side stream while the next launch owns the tensor cores (only
        // the last one stays exposed): +4.4 % on the all-pairs run.  Not when the operands are re-derived per launch (three
        // smaller unpack + tensor launches cost more than the hidden resolve: measured -3 % on 256-pair batches).
        bool side = false;
        // (Only with the separate tie-resolution kernel, SFMGMS_FP4_FUSED_RESOLVE=0: by default the tensor-core kernel resolves
        // its own ties and a batch is one launch.)
        if (kind == SFMGMS_HAMMING_FP4 && !fp4_fused_resolve() && rn >= 24 && !tl_marks && ctx->overlap_resolve && ctx->tc.cache_enabled &&
            ctx->tc.span_lo && hpp[0].img1 >= 0) {
            const int nsub = rn >= 96 ? 3 : 2;
            const int per = (rn + nsub - 1) / nsub;
            if (per < kMaxPerLaunch) kMaxPerLaunch = per;
            if (!ctx->side_stream) CU(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
            side = true;
        }
        size_t ev_k = 0;
        for (int c0 = 0; c0 < rn; c0 += kMaxPerLaunch) {
            const int cn = (rn - c0 < kMaxPerLaunch) ? rn - c0 : kMaxPerLaunch;
            int l;
            if (kind == SFMGMS_HAMMING_TC) {
                l = launch_hamming_tc(ctx->tc, dp + c0, hpp + c0, cn, ctx->sm_count, st);
                if (l < 0) return fail(ctx, SFMGMS_ERR_CUDA, "tensor-core Hamming launch failed: %s", tc_last_error());
            } else if (kind == SFMGMS_HAMMING_FP4) {
                cudaEvent_t ev = nullptr;
                if (side) {
                    while (ctx->side_events.size() <= ev_k) {
                        cudaEvent_t e;
                        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                        ctx->side_events.push_back(e);
                    }
                    ev = ctx->side_events[ev_k++];
                }
                l = launch_hamming_fp4(ctx->tc, dp + c0, hpp + c0, cn, ctx->sm_count, st, side ? ctx->side_stream : nullptr, ev);
                if (l < 0) return fail(ctx, SFMGMS_ERR_CUDA, "fp4 tensor-core Hamming launch failed: %s", fp4_last_error());
            } else {
                l = launch_hamming_popc(dp + c0, hpp + c0, cn, ctx->sm_count, st);
            }
            ctx->launches += l;
            if (ham_launches) *ham_launches += l;
        }
        if (side) {   // join: everything after this point on the main stream sees the resolved keys
            while (ctx->side_events.size() <= ev_k) {
                cudaEvent_t e;
                CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->side_events.push_back(e);
            }
            CU(cudaEventRecord(ctx->side_events[ev_k], ctx->side_stream));
            CU(cudaStreamWaitEvent(st, ctx->side_events[ev_k], 0));
        }
        CU(cudaGetLastError());
    }
    if (mid) CU(cudaEventRecord(mid, st));   // between the Hamming and the GMS stage
    if (do_gms) {
        const int n_scales = with_scale ? kNumScales : 1;
        const size_t per_pair = gms_scratch_bytes_per_pair(n_scales);
        size_t budget = ctx->gms_chunk_bytes;
        if (budget < per_pair) budget = per_pair;
        if (budget > per_pair * (size_t)rn) budget = per_pair * (size_t)rn;
        CU(ctx->d_hist.ensure(budget));
        CU(ctx->d_msc.ensure(gms_match_scratch_bytes(gms_match_rows(hpp, rn), n_scales)));
        int l = launch_gms(dp, hpp, rn, with_rotation, with_scale, factor,
                           static_cast<PairResult*>(ctx->d_results.p) + r0, ctx->d_hist.p, budget, ctx->d_msc.p, st,
                           ctx->gms_dense);
        if (l < 0) return fail(ctx, SFMGMS_ERR_CUDA, "GMS scratch too small");
        ctx->launches += l;
        CU(cudaGetLastError());
    }
    return SFMGMS_OK;
}

// Waits for the batch, downloads the per-pair results and turns device-side status flags into error codes.
int finish_batch(sfmgms_ctx* ctx, int n, bool do_gms) {
    cudaStream_t st = ctx->stream;
    ctx->last_results.assign(n, PairResult{0, -1, 0, 0});
    if (do_gms && n > 0) {
        CU(ctx->h_results.ensure(sizeof(PairResult) * n));
        CU(cudaMemcpyAsync(ctx->h_results.p, ctx->d_results.p, sizeof(PairResult) * n, cudaMemcpyDeviceToHost, st));
    }
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[2], st));
    CU(cudaStreamSynchronize(st));
    tc_reset_arena(ctx->tc);
    if (do_gms && n > 0) {
        memcpy(ctx->last_results.data(), ctx->h_results.p, sizeof(PairResult) * n);
        for (int p = 0; p < n; ++p) {
            const int s = ctx->last_results[p].status;
            if (s == 4) return fail(ctx, SFMGMS_ERR_INDEX, "pair %d: queryIdx/trainIdx out of range", p);
            if (s == 3) return fail(ctx, SFMGMS_ERR_DOMAIN, "pair %d: matched keypoint outside [0,w)x[0,h)", p);
        }
    }
    return SFMGMS_OK;
}

// One synchronous batch: upload table, Hamming + GMS over all pairs, results back.
// hp[].key / hp[].mask must already point into device memory; keys must be initialised to kKeyInit.
int run_batch(sfmgms_ctx* ctx, std::vector<PairDesc>& hp, bool do_hamming, bool do_gms, int with_rotation,
              int with_scale, double factor) {
    const int n = (int)hp.size();
    ctx->last_results.assign(n, PairResult{0, -1, 0, 0});
    if (n == 0) return SFMGMS_OK;
    int rc = upload_pairs(ctx, hp);
    if (rc) return rc;
    int ham_launches = 0;
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (ctx->timing >= 2) { ctx->marks.used = 0; tl_marks = &ctx->marks; kmark("begin", ctx->stream); }
    rc = enqueue_range(ctx, hp, 0, n, do_hamming, do_gms, with_rotation, with_scale, factor, &ham_launches,
                       ctx->timing ? ctx->ev[1] : nullptr);
    tl_marks = nullptr;
    if (rc) { cudaStreamSynchronize(ctx->stream); tc_reset_arena(ctx->tc); return rc; }
    rc = finish_batch(ctx, n, do_gms);
    if (rc) return rc;
    if (ctx->timing >= 2) ctx->absorb(ctx->marks);
    if (ctx->timing) {
        float a = 0.f, b = 0.f;
        CU(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
        CU(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
        ctx->last_ms[0] = a; ctx->last_ms[1] = b; ctx->last_ms[2] = ham_launches;
    }
    return SFMGMS_OK;
}

// gather (x,y) float pairs at a byte stride into a packed device buffer via pinned staging
int upload_xy(sfmgms_ctx* ctx, const void* src, int n, int stride_bytes, DevBuf& dst, size_t stage_off) {
    CU(dst.ensure((size_t)(n > 0 ? n : 1) * 8));
    if (n <= 0) return SFMGMS_OK;
    if (stride_bytes == 8) {
        CU(cudaMemcpyAsync(dst.p, src, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
        return SFMGMS_OK;
    }
    float* stg = reinterpret_cast<float*>(static_cast<char*>(ctx->h_stage.p) + stage_off);
    const char* s = static_cast<const char*>(src);
    for (int i = 0; i < n; ++i) memcpy(stg + 2 * (size_t)i, s + (size_t)i * stride_bytes, 8);
    CU(cudaMemcpyAsync(dst.p, stg, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    return SFMGMS_OK;
}

int upload_idx(sfmgms_ctx* ctx, const int32_t* src, int n, int stride_bytes, DevBuf& dst, size_t stage_off) {
    CU(dst.ensure((size_t)(n > 0 ? n : 1) * 4));
    if (n <= 0) return SFMGMS_OK;
    if (stride_bytes == 4) {
        CU(cudaMemcpyAsync(dst.p, src, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        return SFMGMS_OK;
    }
    int32_t* stg = reinterpret_cast<int32_t*>(static_cast<char*>(ctx->h_stage.p) + stage_off);
    const char* s = reinterpret_cast<const char*>(src);
    for (int i = 0; i < n; ++i) memcpy(stg + i, s + (size_t)i * stride_bytes, 4);
    CU(cudaMemcpyAsync(dst.p, stg, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    return SFMGMS_OK;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

#define GUARD_BEGIN                  \
    if (!ctx) return SFMGMS_ERR_ARG; \
    DeviceGuard guard__(ctx->device); \
    cudaGetLastError(); /* a stale non-sticky error of someone else's call must not be blamed on ours */ \
    try {
#define GUARD_END                                                        \
    }                                                                    \
    catch (const std::bad_alloc&) {                                      \
        return fail(ctx, SFMGMS_ERR_ARG, "host allocation failed");      \
    }                                                                    \
    catch (...) {                                                        \
        return fail(ctx, SFMGMS_ERR_ARG, "unexpected C++ exception");    \
    }

extern "C" {

int sfmgms_version(void) { return 100; }

int sfmgms_create(sfmgms_ctx** out, int device) {
    if (!out) return fail(nullptr, SFMGMS_ERR_ARG, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0)
        return fail(nullptr, SFMGMS_ERR_CUDA, "no CUDA device (%s); this library has no CPU fallback",
                    cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(nullptr, SFMGMS_ERR_ARG, "device %d out of range [0,%d)", device, count);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(nullptr, SFMGMS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SFMGMS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    sfmgms_ctx* c = new (std::nothrow) sfmgms_ctx();
    if (!c) return fail(nullptr, SFMGMS_ERR_ARG, "host allocation failed");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    DeviceGuard g(device);
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete c;
        return fail(nullptr, SFMGMS_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    *out = c;
    return SFMGMS_OK;
}

void sfmgms_destroy(sfmgms_ctx* ctx) {
    if (!ctx) return;
    DeviceGuard g(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_pairs, &ctx->d_results, &ctx->d_key, &ctx->d_mask, &ctx->d_hist, &ctx->d_msc, &ctx->d_q,
                      &ctx->d_kp1, &ctx->d_kp2, &ctx->d_mq, &ctx->d_mt, &ctx->d_out_i32, &ctx->d_pts,
                      &ctx->d_set_desc, &ctx->d_set_kp, &ctx->d_cbase};
    for (DevBuf* b : bufs) b->release();
    for (auto& sl : ctx->slot) {
        DevBuf* sb[] = {&sl.key, &sl.mask, &sl.out_i32, &sl.cmatch, &sl.cpts, &sl.coff};
        for (DevBuf* b : sb) b->release();
        cudaEvent_t* ev[] = {&sl.computed, &sl.copied, &sl.t0, &sl.t1, &sl.t2};
        for (cudaEvent_t* e : ev) if (*e) cudaEventDestroy(*e);
    }
    ctx->h_coff.release();
    for (int i = 0; i < ctx->marks.created; ++i) cudaEventDestroy(ctx->marks.ev[i]);
    for (auto& sl : ctx->slot) for (int i = 0; i < sl.marks.created; ++i) cudaEventDestroy(sl.marks.ev[i]);
    tc_release(ctx->tc);
    if (ctx->orb) orb_ws_destroy(ctx->orb);
    for (OrbWorkspace* w : ctx->orb_pool) orb_ws_destroy(w);
    for (cudaStream_t q : ctx->orb_streams) cudaStreamDestroy(q);
    ctx->h_stage.release(); ctx->h_pairs_pinned.release(); ctx->h_results.release();
    for (int k = 0; k < 3; ++k) if (ctx->ev[k]) cudaEventDestroy(ctx->ev[k]);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->side_events) cudaEventDestroy(e);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->h2d_stream) cudaStreamDestroy(ctx->h2d_stream);
    if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sfmgms_last_error(const sfmgms_ctx* ctx) { return ctx ? ctx->err : g_create_error; }

int sfmgms_set_option(sfmgms_ctx* ctx, int key, int64_t value) {
    if (!ctx) return SFMGMS_ERR_ARG;
    if (key == SFMGMS_OPT_HAMMING_KERNEL) {
        if (value < 0 || value > 3) return fail(ctx, SFMGMS_ERR_ARG, "bad hamming kernel %lld", (long long)value);
        if (value == SFMGMS_HAMMING_TC && !tc_available())
            return fail(ctx, SFMGMS_ERR_ARG, "tensor-core Hamming kernel not available in this build");
        ctx->hamming_kernel = (int)value;
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_GMS_CHUNK_BYTES) {
        if (value < (1 << 20)) return fail(ctx, SFMGMS_ERR_ARG, "chunk budget too small");
        ctx->gms_chunk_bytes = (size_t)value;
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_CHUNK_ROWS) {
        if (value < 1) return fail(ctx, SFMGMS_ERR_ARG, "chunk rows must be positive");
        ctx->chunk_rows = (long long)value;
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_GMS_DENSE) {
        ctx->gms_dense = value ? 1 : 0;
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_COMPACT_RECORD) {
        if (value != 0 && value != 1) return fail(ctx, SFMGMS_ERR_ARG, "bad compact record type %lld", (long long)value);
        ctx->compact_rec_bytes = value ? 8 : 16;
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_OVERLAP) {
        ctx->overlap_resolve = value ? 1 : 0;
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_L2_KERNEL) {
        if (value < 0 || value > 3) return fail(ctx, SFMGMS_ERR_ARG, "bad L2 kernel %lld", (long long)value);
        ctx->l2_kernel = (int)value;
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_TC_OPERAND_CACHE) {
        ctx->tc.cache_enabled = value != 0;
        tc_invalidate(ctx->tc);
        return SFMGMS_OK;
    }
    if (key == SFMGMS_OPT_TIMING) {
        DeviceGuard g(ctx->device);
        if (value && !ctx->ev[0])
            for (int k = 0; k < 3; ++k)
                if (cudaEventCreate(&ctx->ev[k]) != cudaSuccess) return fail(ctx, SFMGMS_ERR_CUDA, "cudaEventCreate failed");
        ctx->timing = value >= 2 ? 2 : (value ? 1 : 0);
        if (ctx->timing < 2) ctx->kernel_times.clear();
        return SFMGMS_OK;
    }
    return fail(ctx, SFMGMS_ERR_ARG, "unknown option %d", key);
}

int sfmgms_last_timing(sfmgms_ctx* ctx, double* out_ms) {
    if (!ctx || !out_ms) return SFMGMS_ERR_ARG;
    if (!ctx->timing) return fail(ctx, SFMGMS_ERR_STATE, "SFMGMS_OPT_TIMING is off");
    for (int k = 0; k < 3; ++k) out_ms[k] = ctx->last_ms[k];
    return SFMGMS_OK;
}

int64_t sfmgms_kernel_launches(const sfmgms_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* sfmgms_stream(sfmgms_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---------------------------------------------------------------------------------------------------
static int bf_common(sfmgms_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt, int desc_bytes) {
    if (nq < 0 || nt < 0) return fail(ctx, SFMGMS_ERR_ARG, "negative row count");
    if (desc_bytes != kDescBytes) return fail(ctx, SFMGMS_ERR_ARG, "desc_bytes must be 32 (256-bit descriptors), got %d", desc_bytes);
    if ((nq > 0 && !query) || (nt > 0 && !train)) return fail(ctx, SFMGMS_ERR_ARG, "null descriptor pointer");
    if (nt >= SFMGMS_MAX_TRAIN_ROWS)
        return fail(ctx, SFMGMS_ERR_TRAIN_ROWS, "train rows %d >= 2^18 (OpenCV BFMatcher: rows < IMGIDX_ONE)", nt);
    return SFMGMS_OK;
}

int sfmgms_bf_hamming(sfmgms_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt, int desc_bytes,
                      int32_t* train_idx, int32_t* dist, int* n_matches) {
    GUARD_BEGIN
    int rc = bf_common(ctx, query, nq, train, nt, desc_bytes);
    if (rc) return rc;
    if (n_matches) *n_matches = (nt == 0) ? 0 : nq;
    if (nt == 0 || nq == 0) return SFMGMS_OK;
    cudaStream_t st = ctx->stream;
    CU(ctx->d_q.ensure((size_t)(nq + nt) * 32));   // q|t contiguous: one operand array (tensor-core path)
    uint8_t* d_train = (uint8_t*)ctx->d_q.p + (size_t)nq * 32;
    CU(ctx->d_key.ensure((size_t)nq * 4)); CU(ctx->d_out_i32.ensure((size_t)nq * 8));
    CU(cudaMemcpyAsync(ctx->d_q.p, query, (size_t)nq * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_train, train, (size_t)nt * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(ctx->d_key.p, 0xFF, (size_t)nq * 4, st));
    std::vector<PairDesc> hp(1);
    PairDesc& p = hp[0];
    memset(&p, 0, sizeof p);
    p.desc1 = (const uint8_t*)ctx->d_q.p; p.desc2 = d_train;
    p.key = (uint32_t*)ctx->d_key.p; p.n1 = nq; p.n2 = nt; p.n_matches = nq; p.img1 = p.img2 = -1;
    rc = run_batch(ctx, hp, true, false, 0, 0, 0.0);
    if (rc) return rc;
    int32_t* o = (int32_t*)ctx->d_out_i32.p;
    decode_keys_kernel<<<(nq + 255) / 256, 256, 0, st>>>(p.key, nq, o, o + nq);
    ctx->launches++;
    if (train_idx) CU(cudaMemcpyAsync(train_idx, o, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (dist) CU(cudaMemcpyAsync(dist, o + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_bf_hamming_crosscheck(sfmgms_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                                 int desc_bytes, int32_t* train_idx, int32_t* dist, uint8_t* keep) {
    GUARD_BEGIN
    int rc = bf_common(ctx, query, nq, train, nt, desc_bytes);
    if (rc) return rc;
    if (nq >= SFMGMS_MAX_TRAIN_ROWS) return fail(ctx, SFMGMS_ERR_TRAIN_ROWS, "query rows %d >= 2^18", nq);
    if (nt == 0 || nq == 0) return SFMGMS_OK;
    cudaStream_t st = ctx->stream;
    CU(ctx->d_q.ensure((size_t)(nq + nt) * 32));
    uint8_t* d_train = (uint8_t*)ctx->d_q.p + (size_t)nq * 32;
    CU(ctx->d_key.ensure((size_t)(nq + nt) * 4)); CU(ctx->d_out_i32.ensure((size_t)nq * 8)); CU(ctx->d_mask.ensure(nq));
    CU(cudaMemcpyAsync(ctx->d_q.p, query, (size_t)nq * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_train, train, (size_t)nt * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(ctx->d_key.p, 0xFF, (size_t)(nq + nt) * 4, st));
    std::vector<PairDesc> hp(2);
    memset(hp.data(), 0, sizeof(PairDesc) * 2);
    hp[0].desc1 = (const uint8_t*)ctx->d_q.p; hp[0].desc2 = d_train;
    hp[0].key = (uint32_t*)ctx->d_key.p; hp[0].n1 = nq; hp[0].n2 = nt; hp[0].n_matches = nq;
    hp[1].desc1 = hp[0].desc2; hp[1].desc2 = hp[0].desc1;      // roles swapped: one extra "pair"
    hp[1].key = hp[0].key + nq; hp[1].n1 = nt; hp[1].n2 = nq; hp[1].n_matches = nt; hp[1].match_base = nq;
    hp[0].img1 = hp[0].img2 = hp[1].img1 = hp[1].img2 = -1;
    rc = run_batch(ctx, hp, true, false, 0, 0, 0.0);
    if (rc) return rc;
    int32_t* o = (int32_t*)ctx->d_out_i32.p;
    decode_keys_kernel<<<(nq + 255) / 256, 256, 0, st>>>(hp[0].key, nq, o, o + nq);
    crosscheck_kernel<<<(nq + 255) / 256, 256, 0, st>>>(hp[0].key, hp[1].key, nq, (uint8_t*)ctx->d_mask.p);
    ctx->launches += 2;
    if (train_idx) CU(cudaMemcpyAsync(train_idx, o, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (dist) CU(cudaMemcpyAsync(dist, o + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (keep) CU(cudaMemcpyAsync(keep, ctx->d_mask.p, (size_t)nq, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SFMGMS_OK;
    GUARD_END
}

// Nearest train row of every row of a[na] in b[nb] -> out = idx[na] | dist[na] (device, on ctx->stream).
// dim 128 first tries the integer tensor-core / DP4A path (OpenCV SIFT data: exact in any order); rows that are not
// integer-valued in [0,255] raise its flag and the pass is redone by the order-exact fp32 kernel (l2_f32.cu).
// Synchronises the stream when it has to read the flag.
static int l2_nn(sfmgms_ctx* ctx, const float* a, int na, const float* b, int nb, int dim, int32_t* out, int* d_bad,
                 int* n_launches) {
    cudaStream_t st = ctx->stream;
    const int mode = ctx->l2_kernel;            // 0 auto, 1 dp4a, 2 tcgen05, 3 fp32 order-exact
    bool need_f32 = (dim != 128) || mode == 3;
    if (!need_f32) {
        const int l = mode == 1 ? launch_l2_dp4a(a, na, b, nb, ctx->d_hist.p, out, (float*)(out + na), d_bad, ctx->sm_count, st)
                                : launch_l2_tc(a, na, b, nb, ctx->d_hist.p, out, (float*)(out + na), d_bad, ctx->sm_count, st);
        if (l < 0) return fail(ctx, SFMGMS_ERR_CUDA, "L2 tensor-core launch setup failed");
        *n_launches += l;
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        need_f32 = bad != 0;
    }
    if (need_f32) {
        const int l = launch_l2_f32(a, na, b, nb, dim, ctx->d_hist.p, out, (float*)(out + na), st);
        if (l < 0) return fail(ctx, SFMGMS_ERR_CUDA, "L2 fp32 launch setup failed");
        *n_launches += l;
    }
    return SFMGMS_OK;
}

static size_t l2_scratch_need(sfmgms_ctx* ctx, int na, int nb) {
    const size_t x = ctx->l2_kernel == 1 ? l2_scratch_bytes(na, nb) : l2_tc_scratch_bytes(na, nb);
    const size_t y = l2_f32_scratch_bytes(na);
    return x > y ? x : y;
}

static int l2_args(sfmgms_ctx* ctx, const float* query, int nq, const float* train, int nt, int dim, bool cross) {
    if (nq < 0 || nt < 0) return fail(ctx, SFMGMS_ERR_ARG, "negative row count");
    if (dim < 1 || dim > l2_f32_max_dim()) return fail(ctx, SFMGMS_ERR_ARG, "dim must be in [1, %d], got %d", l2_f32_max_dim(), dim);
    if ((nq > 0 && !query) || (nt > 0 && !train)) return fail(ctx, SFMGMS_ERR_ARG, "null descriptor pointer");
    if (nt >= SFMGMS_MAX_TRAIN_ROWS || (cross && nq >= SFMGMS_MAX_TRAIN_ROWS))
        return fail(ctx, SFMGMS_ERR_TRAIN_ROWS, "rows %d / %d >= 2^18 (OpenCV BFMatcher: rows < IMGIDX_ONE)", nq, nt);
    return SFMGMS_OK;
}

int sfmgms_bf_l2(sfmgms_ctx* ctx, const float* query, int nq, const float* train, int nt, int dim, int32_t* train_idx,
                 float* dist, int* n_matches) {
    GUARD_BEGIN
    int rc = l2_args(ctx, query, nq, train, nt, dim, false);
    if (rc) return rc;
    if (n_matches) *n_matches = (nt == 0) ? 0 : nq;
    if (nt == 0 || nq == 0) return SFMGMS_OK;
    cudaStream_t st = ctx->stream;
    const size_t rowb = (size_t)dim * 4;
    CU(ctx->d_q.ensure((size_t)(nq + nt) * rowb));
    float* dq = (float*)ctx->d_q.p;
    float* dt = dq + (size_t)nq * dim;
    CU(ctx->d_hist.ensure(l2_scratch_need(ctx, nq, nt)));
    CU(ctx->d_out_i32.ensure((size_t)nq * 8 + 16));
    CU(cudaMemcpyAsync(dq, query, (size_t)nq * rowb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dt, train, (size_t)nt * rowb, cudaMemcpyHostToDevice, st));
    int32_t* o = (int32_t*)ctx->d_out_i32.p;
    int* d_bad = (int*)(o + 2 * (size_t)nq);
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[0], st));
    int nl = 0;
    rc = l2_nn(ctx, dq, nq, dt, nt, dim, o, d_bad, &nl);
    if (rc) return rc;
    ctx->launches += nl;
    if (ctx->timing) { CU(cudaEventRecord(ctx->ev[1], st)); CU(cudaEventRecord(ctx->ev[2], st)); }
    CU(cudaGetLastError());
    if (train_idx) CU(cudaMemcpyAsync(train_idx, o, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (dist) CU(cudaMemcpyAsync(dist, o + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ctx->timing) {
        float a = 0.f;
        CU(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
        ctx->last_ms[0] = a; ctx->last_ms[1] = 0; ctx->last_ms[2] = nl;
    }
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_bf_l2_crosscheck(sfmgms_ctx* ctx, const float* query, int nq, const float* train, int nt, int dim,
                            int32_t* train_idx, float* dist, uint8_t* keep) {
    GUARD_BEGIN
    int rc = l2_args(ctx, query, nq, train, nt, dim, true);
    if (rc) return rc;
    if (nt == 0 || nq == 0) return SFMGMS_OK;
    cudaStream_t st = ctx->stream;
    const size_t rowb = (size_t)dim * 4;
    CU(ctx->d_q.ensure((size_t)(nq + nt) * rowb));
    float* dq = (float*)ctx->d_q.p;
    float* dt = dq + (size_t)nq * dim;
    const size_t s_fwd = l2_scratch_need(ctx, nq, nt), s_rev = l2_scratch_need(ctx, nt, nq);
    CU(ctx->d_hist.ensure(s_fwd > s_rev ? s_fwd : s_rev));
    CU(ctx->d_out_i32.ensure(((size_t)nq + nt) * 8 + 16));
    CU(ctx->d_mask.ensure(nq));
    CU(cudaMemcpyAsync(dq, query, (size_t)nq * rowb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(dt, train, (size_t)nt * rowb, cudaMemcpyHostToDevice, st));
    int32_t* o = (int32_t*)ctx->d_out_i32.p;          // forward: idx[nq], dist[nq]
    int32_t* r = o + 2 * (size_t)nq;                  // reverse: idx[nt], dist[nt]
    int* d_bad = (int*)(r + 2 * (size_t)nt);
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[0], st));
    int nl = 0;
    rc = l2_nn(ctx, dq, nq, dt, nt, dim, o, d_bad, &nl);
    if (rc) return rc;
    rc = l2_nn(ctx, dt, nt, dq, nq, dim, r, d_bad, &nl);
    if (rc) return rc;
    crosscheck_idx_kernel<<<(nq + 255) / 256, 256, 0, st>>>(o, r, nq, (uint8_t*)ctx->d_mask.p);
    ctx->launches += nl + 1;
    if (ctx->timing) { CU(cudaEventRecord(ctx->ev[1], st)); CU(cudaEventRecord(ctx->ev[2], st)); }
    CU(cudaGetLastError());
    if (train_idx) CU(cudaMemcpyAsync(train_idx, o, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (dist) CU(cudaMemcpyAsync(dist, o + nq, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    if (keep) CU(cudaMemcpyAsync(keep, ctx->d_mask.p, (size_t)nq, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ctx->timing) {
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
        ctx->last_ms[0] = t; ctx->last_ms[1] = 0; ctx->last_ms[2] = nl;
    }
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_brute_force_match(sfmgms_ctx* ctx, int norm_type, int cross_check, const void* query, int nq,
                             const void* train, int nt, int width, double distance_coef, int max_matching_size,
                             int32_t* query_idx, int32_t* train_idx, float* dist, int capacity, int* n_out) {
    GUARD_BEGIN
    if (n_out) *n_out = 0;
    if (norm_type != SFMGMS_NORM_L2 && norm_type != SFMGMS_NORM_HAMMING)
        return fail(ctx, SFMGMS_ERR_ARG, "norm_type must be SFMGMS_NORM_L2 (4) or SFMGMS_NORM_HAMMING (6), got %d", norm_type);
    if (nq < 0 || nt < 0 || capacity < 0 || max_matching_size < 0) return fail(ctx, SFMGMS_ERR_ARG, "negative size");
    // stage 1 on the device: (cross-checked) nearest neighbours in query order
    std::vector<int32_t> idx((size_t)nq), idist;
    std::vector<float> fdist;
    std::vector<uint8_t> keep((size_t)nq, 1);
    int n = 0, rc;
    if (norm_type == SFMGMS_NORM_L2) {
        fdist.resize((size_t)nq);
        rc = cross_check ? sfmgms_bf_l2_crosscheck(ctx, (const float*)query, nq, (const float*)train, nt, width, idx.data(), fdist.data(), keep.data())
                         : sfmgms_bf_l2(ctx, (const float*)query, nq, (const float*)train, nt, width, idx.data(), fdist.data(), &n);
    } else {
        idist.resize((size_t)nq);
        rc = cross_check ? sfmgms_bf_hamming_crosscheck(ctx, (const uint8_t*)query, nq, (const uint8_t*)train, nt, width, idx.data(), idist.data(), keep.data())
                         : sfmgms_bf_hamming(ctx, (const uint8_t*)query, nq, (const uint8_t*)train, nt, width, idx.data(), idist.data(), &n);
    }
    if (rc) return rc;
    if (nq == 0 || nt == 0) return SFMGMS_OK;
    // the tail of FeatureMatchUtil.cpp:24-30 on the host, as in the reference: sort, ratio prune, cap
    struct M { float d; int32_t q, t; };
    std::vector<M> m;
    m.reserve((size_t)nq);
    for (int i = 0; i < nq; ++i)
        if (keep[i]) m.push_back({norm_type == SFMGMS_NORM_L2 ? fdist[i] : (float)idist[i], i, idx[i]});
    std::stable_sort(m.begin(), m.end(), [](const M& a, const M& b) { return a.d < b.d; });   // DMatch::operator<
    while (!m.empty() && (double)m.front().d * distance_coef < (double)m.back().d) m.pop_back();
    if (m.size() > (size_t)max_matching_size) m.resize((size_t)max_matching_size);
    if (m.size() > (size_t)capacity) return fail(ctx, SFMGMS_ERR_ARG, "capacity %d < %zu surviving matches", capacity, m.size());
    for (size_t k = 0; k < m.size(); ++k) {
        if (query_idx) query_idx[k] = m[k].q;
        if (train_idx) train_idx[k] = m[k].t;
        if (dist) dist[k] = m[k].d;
    }
    if (n_out) *n_out = (int)m.size();
    return SFMGMS_OK;
    GUARD_END
}

// ---- (§8f-4) cv::ORB: compute on provided keypoints, detectAndCompute -----------------------------------
static int orb_image_args(sfmgms_ctx* ctx, const uint8_t* image, int width, int height, int channels, int stride_bytes) {
    if (!image || width <= 0 || height <= 0) return fail(ctx, SFMGMS_ERR_ARG, "empty image");
    if (channels != 1 && channels != 3) return fail(ctx, SFMGMS_ERR_ARG, "channels must be 1 (gray) or 3 (BGR), got %d", channels);
    if (stride_bytes < width * channels) return fail(ctx, SFMGMS_ERR_ARG, "row stride %d < %d", stride_bytes, width * channels);
    if (!ctx->orb) ctx->orb = orb_ws_create();
    return SFMGMS_OK;
}

int sfmgms_orb_compute(sfmgms_ctx* ctx, const uint8_t* image, int width, int height, int channels, int stride_bytes,
                       const void* keypoints, int n_keypoints, int kp_stride_bytes, int angle_offset_bytes,
                       int octave_offset_bytes, int32_t* kept_index, uint8_t* descriptors, int* n_kept) {
    GUARD_BEGIN
    if (n_kept) *n_kept = 0;
    int rc = orb_image_args(ctx, image, width, height, channels, stride_bytes);
    if (rc) return rc;
    if (n_keypoints < 0 || (n_keypoints > 0 && !keypoints) || kp_stride_bytes < 8)
        return fail(ctx, SFMGMS_ERR_ARG, "bad keypoint array");
    if (n_keypoints == 0) return SFMGMS_OK;
    // KeyPointsFilter::runByImageBorder(keypoints, image.size(), edgeThreshold = 31): a host loop, as in OpenCV.
    // Rect(31, 31, w-62, h-62).contains(Point(pt)): Point2f -> Point rounds with cvRound (half to even).
    const int kEdge = 31;
    struct K { float x, y, ang; int32_t oct; int32_t idx; };
    std::vector<K> kept;
    kept.reserve((size_t)n_keypoints);
    const char* base = (const char*)keypoints;
    int max_oct = 0;
    int32_t prev_oct = 0;
    bool sorted = true;
    for (int i = 0; i < n_keypoints; ++i) {
        const char* kp = base + (size_t)i * kp_stride_bytes;
        K k; k.ang = -1.f; k.oct = 0; k.idx = i;
        memcpy(&k.x, kp, 4); memcpy(&k.y, kp + 4, 4);
        if (angle_offset_bytes >= 0) memcpy(&k.ang, kp + angle_offset_bytes, 4);
        if (octave_offset_bytes >= 0) memcpy(&k.oct, kp + octave_offset_bytes, 4);
        if (k.oct < 0 || k.oct > 15) return fail(ctx, SFMGMS_ERR_ARG, "keypoint %d has octave %d (supported: 0..15)", i, k.oct);
        // the level count and the sortedness test look at ALL keypoints, before the border filter (orb.cpp)
        if (i > 0 && k.oct < prev_oct) sorted = false;
        prev_oct = k.oct;
        if (k.oct > max_oct) max_oct = k.oct;
        const long cx = lrintf(k.x), cy = lrintf(k.y);
        if (!(cx >= kEdge && cx < width - kEdge && cy >= kEdge && cy < height - kEdge)) continue;
        kept.push_back(k);
    }
    // keypoints not sorted by level are regrouped level by level, input order kept inside a level (orb.cpp)
    if (!sorted) std::stable_sort(kept.begin(), kept.end(), [](const K& a, const K& b) { return a.oct < b.oct; });
    const int n = (int)kept.size();
    if (n_kept) *n_kept = n;
    if (n == 0) return SFMGMS_OK;
    std::vector<float> xyao((size_t)n * 4);
    for (int i = 0; i < n; ++i) {
        xyao[4 * (size_t)i] = kept[(size_t)i].x; xyao[4 * (size_t)i + 1] = kept[(size_t)i].y; xyao[4 * (size_t)i + 2] = kept[(size_t)i].ang;
        memcpy(&xyao[4 * (size_t)i + 3], &kept[(size_t)i].oct, 4);
        if (kept_index) kept_index[i] = kept[(size_t)i].idx;
    }
    cudaStream_t st = ctx->stream;
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[0], st));
    int nl = 0;
    const int got = orb_compute_provided(ctx->orb, image, width, height, channels, stride_bytes, xyao.data(), n, max_oct + 1, descriptors,
                                         ctx->sm_count, st, &nl);
    ctx->launches += nl;
    if (got < 0) return fail(ctx, SFMGMS_ERR_CUDA, "%s", orb_ws_error(ctx->orb));
    if (ctx->timing) {
        CU(cudaEventRecord(ctx->ev[1], st)); CU(cudaEventRecord(ctx->ev[2], st));
        CU(cudaStreamSynchronize(st));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
        ctx->last_ms[0] = t; ctx->last_ms[1] = 0; ctx->last_ms[2] = nl;
    }
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_orb_detect_and_compute_ex(sfmgms_ctx* ctx, const uint8_t* image, int width, int height, int channels, int stride_bytes,
                                     const sfmgms_orb_params* prm, void* keypoints, uint8_t* descriptors, int capacity,
                                     int* n_keypoints) {
    GUARD_BEGIN
    if (n_keypoints) *n_keypoints = 0;
    int rc = orb_image_args(ctx, image, width, height, channels, stride_bytes);
    if (rc) return rc;
    if (!prm) return fail(ctx, SFMGMS_ERR_ARG, "null parameter block");
    if (prm->nfeatures < 0 || capacity < 0 || prm->fast_threshold < 0 || prm->fast_threshold > 255)
        return fail(ctx, SFMGMS_ERR_ARG, "bad nfeatures / capacity / fast_threshold");
    if (!(prm->scale_factor > 1.0f) || prm->nlevels < 1 || prm->nlevels > 16)
        return fail(ctx, SFMGMS_ERR_ARG, "scale_factor must be > 1 and nlevels in 1..16");
    if (prm->first_level != 0) return fail(ctx, SFMGMS_ERR_ARG, "firstLevel %d: only 0 is implemented", prm->first_level);
    if (prm->wta_k < 2 || prm->wta_k > 4 || prm->patch_size < 2 || prm->patch_size > 63 || (prm->score_type != 0 && prm->score_type != 1) ||
        prm->edge_threshold < 0)
        return fail(ctx, SFMGMS_ERR_ARG, "need WTA_K in 2..4, patchSize in 2..63, scoreType 0 (HARRIS_SCORE) or 1 (FAST_SCORE), edgeThreshold >= 0");
    cudaStream_t st = ctx->stream;
    if (ctx->timing) CU(cudaEventRecord(ctx->ev[0], st));
    int nl = 0, needed = 0;
    const int got = orb_detect_and_compute(ctx->orb, image, width, height, channels, stride_bytes, prm->nfeatures, prm->fast_threshold,
                                           prm->nlevels, prm->scale_factor, prm->edge_threshold, prm->score_type, prm->wta_k, prm->patch_size,
                                           keypoints, descriptors,
                                           capacity, &needed, ctx->sm_count, st, &nl);
    ctx->launches += nl;
    if (got == -2) { if (n_keypoints) *n_keypoints = needed; return fail(ctx, SFMGMS_ERR_ARG, "capacity %d < %d keypoints", capacity, needed); }
    if (got < 0) return fail(ctx, SFMGMS_ERR_CUDA, "%s", orb_ws_error(ctx->orb));
    if (n_keypoints) *n_keypoints = got;
    if (ctx->timing) {
        CU(cudaEventRecord(ctx->ev[1], st)); CU(cudaEventRecord(ctx->ev[2], st));
        CU(cudaStreamSynchronize(st));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
        ctx->last_ms[0] = t; ctx->last_ms[1] = 0; ctx->last_ms[2] = nl;
    }
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_orb_detect_and_compute(sfmgms_ctx* ctx, const uint8_t* image, int width, int height, int channels, int stride_bytes,
                                  int nfeatures, int fast_threshold, void* keypoints, uint8_t* descriptors, int capacity,
                                  int* n_keypoints) {
    const sfmgms_orb_params prm = {nfeatures, 1.2f, 8, 31, 0, 2, 0, 31, fast_threshold};   // ORB::create() defaults
    return sfmgms_orb_detect_and_compute_ex(ctx, image, width, height, channels, stride_bytes, &prm, keypoints, descriptors, capacity,
                                            n_keypoints);
}

static int gms_args(sfmgms_ctx* ctx, int w1, int h1, int w2, int h2, int n1, int n2, int s1, int s2) {
    if (w1 <= 0 || h1 <= 0 || w2 <= 0 || h2 <= 0) return fail(ctx, SFMGMS_ERR_ARG, "image sizes must be positive");
    if (n1 < 0 || n2 < 0) return fail(ctx, SFMGMS_ERR_ARG, "negative keypoint count");
    if (s1 < 8 || s2 < 8) return fail(ctx, SFMGMS_ERR_ARG, "keypoint stride must be >= 8 bytes");
    return SFMGMS_OK;
}

int sfmgms_gms(sfmgms_ctx* ctx, int w1, int h1, int w2, int h2, const void* kp1, int n1, int kp1_stride_bytes,
               const void* kp2, int n2, int kp2_stride_bytes, const int32_t* query_idx, const int32_t* train_idx,
               int idx_stride_bytes, int n_matches, int with_rotation, int with_scale, double threshold_factor,
               uint8_t* mask, int* mask_len, int* n_inliers, int* best_hyp) {
    GUARD_BEGIN
    int rc = gms_args(ctx, w1, h1, w2, h2, n1, n2, kp1_stride_bytes, kp2_stride_bytes);
    if (rc) return rc;
    if (n_matches < 0 || idx_stride_bytes < 4) return fail(ctx, SFMGMS_ERR_ARG, "bad match count/stride");
    if ((n1 > 0 && !kp1) || (n2 > 0 && !kp2) || (n_matches > 0 && (!query_idx || !train_idx)))
        return fail(ctx, SFMGMS_ERR_ARG, "null input pointer");
    cudaStream_t st = ctx->stream;
    const size_t o1 = 0, o2 = align256((size_t)n1 * 8), o3 = o2 + align256((size_t)n2 * 8),
                 o4 = o3 + align256((size_t)n_matches * 4);
    CU(ctx->h_stage.ensure(o4 + align256((size_t)n_matches * 4)));
    if ((rc = upload_xy(ctx, kp1, n1, kp1_stride_bytes, ctx->d_kp1, o1))) return rc;
    if ((rc = upload_xy(ctx, kp2, n2, kp2_stride_bytes, ctx->d_kp2, o2))) return rc;
    if ((rc = upload_idx(ctx, query_idx, n_matches, idx_stride_bytes, ctx->d_mq, o3))) return rc;
    if ((rc = upload_idx(ctx, train_idx, n_matches, idx_stride_bytes, ctx->d_mt, o4))) return rc;
    CU(ctx->d_mask.ensure((size_t)(n_matches > 0 ? n_matches : 1)));
    std::vector<PairDesc> hp(1);
    PairDesc& p = hp[0];
    memset(&p, 0, sizeof p);
    p.kp1 = (const float*)ctx->d_kp1.p; p.kp2 = (const float*)ctx->d_kp2.p;
    p.mq = (const int32_t*)ctx->d_mq.p; p.mt = (const int32_t*)ctx->d_mt.p;
    p.mask = (uint8_t*)ctx->d_mask.p;
    p.n1 = n1; p.n2 = n2; p.n_matches = n_matches; p.w1 = w1; p.h1 = h1; p.w2 = w2; p.h2 = h2; p.img1 = p.img2 = -1;
    rc = run_batch(ctx, hp, false, true, with_rotation, with_scale, threshold_factor);
    ctx->last_pairs = hp;
    if (rc) return rc;
    const PairResult& r = ctx->last_results[0];
    if (mask && r.mask_len > 0) {
        CU(cudaMemcpyAsync(mask, p.mask, (size_t)r.mask_len, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    if (mask_len) *mask_len = r.mask_len;
    if (n_inliers) *n_inliers = r.n_inliers;
    if (best_hyp) *best_hyp = r.best_hyp;
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_match_pair(sfmgms_ctx* ctx, const uint8_t* desc1, int n1, const uint8_t* desc2, int n2, int desc_bytes,
                      const void* kp1, int kp1_stride_bytes, const void* kp2, int kp2_stride_bytes, int w1, int h1,
                      int w2, int h2, int with_rotation, int with_scale, double threshold_factor, int32_t* train_idx,
                      int32_t* dist, uint8_t* mask, int* mask_len, int* n_inliers, int* best_hyp) {
    GUARD_BEGIN
    int rc = bf_common(ctx, desc1, n1, desc2, n2, desc_bytes);
    if (rc) return rc;
    if ((rc = gms_args(ctx, w1, h1, w2, h2, n1, n2, kp1_stride_bytes, kp2_stride_bytes))) return rc;
    if ((n1 > 0 && !kp1) || (n2 > 0 && !kp2)) return fail(ctx, SFMGMS_ERR_ARG, "null keypoint pointer");
    cudaStream_t st = ctx->stream;
    const int nm = (n2 == 0) ? 0 : n1;
    const size_t o2 = align256((size_t)n1 * 8);
    CU(ctx->h_stage.ensure(o2 + align256((size_t)n2 * 8)));
    CU(ctx->d_q.ensure((size_t)(n1 + n2 + 1) * 32));
    uint8_t* d_train = (uint8_t*)ctx->d_q.p + (size_t)n1 * 32;
    CU(ctx->d_key.ensure((size_t)(n1 + 1) * 4)); CU(ctx->d_out_i32.ensure((size_t)(n1 + 1) * 8));
    CU(ctx->d_mask.ensure((size_t)n1 + 1));
    if (n1) CU(cudaMemcpyAsync(ctx->d_q.p, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, st));
    if (n2) CU(cudaMemcpyAsync(d_train, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, st));
    if ((rc = upload_xy(ctx, kp1, n1, kp1_stride_bytes, ctx->d_kp1, 0))) return rc;
    if ((rc = upload_xy(ctx, kp2, n2, kp2_stride_bytes, ctx->d_kp2, o2))) return rc;
    CU(cudaMemsetAsync(ctx->d_key.p, 0xFF, (size_t)(n1 + 1) * 4, st));
    std::vector<PairDesc> hp(1);
    PairDesc& p = hp[0];
    memset(&p, 0, sizeof p);
    p.desc1 = (const uint8_t*)ctx->d_q.p; p.desc2 = d_train;
    p.kp1 = (const float*)ctx->d_kp1.p; p.kp2 = (const float*)ctx->d_kp2.p;
    p.key = (uint32_t*)ctx->d_key.p; p.mask = (uint8_t*)ctx->d_mask.p;
    p.n1 = n1; p.n2 = n2; p.n_matches = nm; p.w1 = w1; p.h1 = h1; p.w2 = w2; p.h2 = h2; p.img1 = p.img2 = -1;
    rc = run_batch(ctx, hp, nm > 0, true, with_rotation, with_scale, threshold_factor);
    ctx->last_pairs = hp;
    if (rc) return rc;
    const PairResult& r = ctx->last_results[0];
    if (nm > 0 && (train_idx || dist)) {
        int32_t* o = (int32_t*)ctx->d_out_i32.p;
        decode_keys_kernel<<<(nm + 255) / 256, 256, 0, st>>>(p.key, nm, o, o + nm);
        ctx->launches++;
        if (train_idx) CU(cudaMemcpyAsync(train_idx, o, (size_t)nm * 4, cudaMemcpyDeviceToHost, st));
        if (dist) CU(cudaMemcpyAsync(dist, o + nm, (size_t)nm * 4, cudaMemcpyDeviceToHost, st));
    }
    if (mask && r.mask_len > 0) CU(cudaMemcpyAsync(mask, p.mask, (size_t)r.mask_len, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (mask_len) *mask_len = r.mask_len;
    if (n_inliers) *n_inliers = r.n_inliers;
    if (best_hyp) *best_hyp = r.best_hyp;
    return SFMGMS_OK;
    GUARD_END
}

// ---------------------------------------------------------------------------------------------------
int sfmgms_set_images(sfmgms_ctx* ctx, int n_images, const int64_t* kp_offsets, const uint8_t* desc,
                      const float* kp_xy, const int32_t* sizes_wh, int location) {
    GUARD_BEGIN
    if (n_images < 0 || !kp_offsets || !sizes_wh) return fail(ctx, SFMGMS_ERR_ARG, "bad image-set arguments");
    if (kp_offsets[0] != 0) return fail(ctx, SFMGMS_ERR_ARG, "kp_offsets[0] must be 0");
    for (int i = 0; i < n_images; ++i) {
        const int64_t n = kp_offsets[i + 1] - kp_offsets[i];
        if (n < 0) return fail(ctx, SFMGMS_ERR_ARG, "kp_offsets not monotone at image %d", i);
        if (n >= SFMGMS_MAX_TRAIN_ROWS) return fail(ctx, SFMGMS_ERR_TRAIN_ROWS, "image %d has %lld rows >= 2^18", i, (long long)n);
        if (sizes_wh[2 * i] <= 0 || sizes_wh[2 * i + 1] <= 0) return fail(ctx, SFMGMS_ERR_ARG, "image %d has a non-positive size", i);
    }
    const int64_t total = kp_offsets[n_images];
    if (total > 0 && (!desc || !kp_xy)) return fail(ctx, SFMGMS_ERR_ARG, "null descriptor/keypoint pointer");
    if (location == SFMGMS_HOST) {
        CU(ctx->d_set_desc.ensure((size_t)total * 32 + 32)); CU(ctx->d_set_kp.ensure((size_t)total * 8 + 8));
        if (total) {
            CU(cudaMemcpyAsync(ctx->d_set_desc.p, desc, (size_t)total * 32, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaMemcpyAsync(ctx->d_set_kp.p, kp_xy, (size_t)total * 8, cudaMemcpyHostToDevice, ctx->stream));
        }
        ctx->set_desc = (const uint8_t*)ctx->d_set_desc.p; ctx->set_kp = (const float*)ctx->d_set_kp.p;
    } else if (location == SFMGMS_DEVICE) {
        if (((uintptr_t)desc & 15) || ((uintptr_t)kp_xy & 7)) return fail(ctx, SFMGMS_ERR_ARG, "device buffers must be 16-byte (desc) / 8-byte (kp) aligned");
        ctx->set_desc = desc; ctx->set_kp = kp_xy;
    } else {
        return fail(ctx, SFMGMS_ERR_ARG, "bad location %d", location);
    }
    ctx->n_images = n_images;
    if (!ctx->keep_orb_keypoints) ctx->orb_keypoints.clear();   // only set_images_from_pixels keeps its own keypoints
    ctx->offsets.assign(kp_offsets, kp_offsets + n_images + 1);
    ctx->sizes.assign(sizes_wh, sizes_wh + 2 * (size_t)n_images);
    ctx->set_version++;
    tc_invalidate(ctx->tc);
    tc_pin_span(ctx->tc, ctx->set_desc, total);   // the set's +-1 operands are derived once and serve every later launch
    ctx->last_pairs.clear();
    return SFMGMS_OK;
    GUARD_END
}

// pixels -> ORB -> image set: the caller side of the path (DisparityUtil.cpp:139-149 per pair; an SfM sequence does
// it once per image).  Keypoints stay available through sfmgms_get_image_keypoints.
int sfmgms_set_images_from_pixels(sfmgms_ctx* ctx, int n_images, const uint8_t* const* images, const int32_t* widths,
                                  const int32_t* heights, const int32_t* channels, const int32_t* strides,
                                  const sfmgms_orb_params* prm, int64_t* kp_offsets_out) {
    GUARD_BEGIN
    if (!prm) return fail(ctx, SFMGMS_ERR_ARG, "null parameter block");
    if (n_images < 0 || (n_images > 0 && (!images || !widths || !heights || !channels)))
        return fail(ctx, SFMGMS_ERR_ARG, "bad image list");
    {   // argument checks once, through the single-image entry point's rules
        sfmgms_orb_params p = *prm;
        if (p.nfeatures < 0 || p.fast_threshold < 0 || p.fast_threshold > 255 || !(p.scale_factor > 1.0f) || p.nlevels < 1 || p.nlevels > 16 ||
            p.first_level != 0 || p.wta_k < 2 || p.wta_k > 4 || p.patch_size < 2 || p.patch_size > 63 || (p.score_type != 0 && p.score_type != 1) ||
            p.edge_threshold < 0)
            return fail(ctx, SFMGMS_ERR_ARG, "bad ORB parameters (see sfmgms_orb_params)");
        for (int i = 0; i < n_images; ++i)
            if (!images[i] || widths[i] <= 0 || heights[i] <= 0 || (channels[i] != 1 && channels[i] != 3) ||
                (strides && strides[i] < widths[i] * channels[i]))
                return fail(ctx, SFMGMS_ERR_ARG, "image %d: bad pointer / size / channels / stride", i);
    }
    // ORB per image is dominated by its host half (retainBest = std::nth_element over every FAST corner, as OpenCV
    // does it), so images are spread over a few host threads, each with its own workspace and stream: one image's
    // host work overlaps the others' kernels.
    struct PerImage { std::vector<uint8_t> kp, desc; int n = 0; };
    std::vector<PerImage> res((size_t)n_images);
    const int n_workers = n_images < 4 ? (n_images > 0 ? n_images : 1) : 4;
    std::vector<std::string> errs((size_t)n_workers);
    std::vector<int> launches((size_t)n_workers, 0);
    const int device = ctx->device, sm_count = ctx->sm_count;
    while ((int)ctx->orb_pool.size() < n_workers) {            // created once per context, reused by later calls
        cudaStream_t q = nullptr;
        CU(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
        ctx->orb_streams.push_back(q);
        ctx->orb_pool.push_back(orb_ws_create());
    }
    std::vector<int> thread_err((size_t)n_workers, 0);   // set when even the error string could not be built
    auto work_body = [&](int t) {
        if (cudaSetDevice(device) != cudaSuccess) { errs[(size_t)t] = "cudaSetDevice failed"; return; }
        cudaStream_t st = ctx->orb_streams[(size_t)t];
        OrbWorkspace* ws = ctx->orb_pool[(size_t)t];
        for (int i = t; i < n_images && errs[(size_t)t].empty(); i += n_workers) {
            PerImage& r = res[(size_t)i];
            int cap = 2 * prm->nfeatures + 64;
            for (int attempt = 0; attempt < 2; ++attempt) {
                r.kp.resize((size_t)cap * 28); r.desc.resize((size_t)cap * 32);
                int needed = 0;
                const int got = orb_detect_and_compute(ws, images[i], widths[i], heights[i], channels[i],
                                                       strides ? strides[i] : widths[i] * channels[i], prm->nfeatures, prm->fast_threshold,
                                                       prm->nlevels, prm->scale_factor, prm->edge_threshold, prm->score_type, prm->wta_k,
                                                       prm->patch_size, r.kp.data(), r.desc.data(), cap, &needed, sm_count, st, &launches[(size_t)t]);
                if (got == -2 && attempt == 0) { cap = needed; continue; }       // ties at a level's cut
                if (got < 0) { errs[(size_t)t] = std::string("image ") + std::to_string(i) + ": " + orb_ws_error(ws); break; }
                r.n = got;
                break;
            }
        }
    };
    auto work = [&](int t) {   // nothing may escape a worker thread (std::terminate) or the extern "C" boundary
        try { work_body(t); } catch (...) { thread_err[(size_t)t] = 1; }
    };
    {
        std::vector<std::thread> th;
        for (int t = 1; t < n_workers; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < n_workers; ++t) {
        ctx->launches += launches[(size_t)t];
        if (thread_err[(size_t)t]) return fail(ctx, SFMGMS_ERR_ARG, "host allocation failed in an ORB worker");
        if (!errs[(size_t)t].empty()) return fail(ctx, SFMGMS_ERR_CUDA, "%s", errs[(size_t)t].c_str());
    }
    std::vector<uint8_t> desc;
    std::vector<float> xy;
    std::vector<int64_t> off((size_t)n_images + 1, 0);
    std::vector<int32_t> sizes((size_t)n_images * 2);
    ctx->orb_keypoints.clear();
    for (int i = 0; i < n_images; ++i) {
        const PerImage& r = res[(size_t)i];
        if (r.n >= SFMGMS_MAX_TRAIN_ROWS) return fail(ctx, SFMGMS_ERR_TRAIN_ROWS, "image %d has %d keypoints >= 2^18", i, r.n);
        off[(size_t)i + 1] = off[(size_t)i] + r.n;
        sizes[2 * (size_t)i] = widths[i]; sizes[2 * (size_t)i + 1] = heights[i];
        desc.insert(desc.end(), r.desc.begin(), r.desc.begin() + (size_t)r.n * 32);
        ctx->orb_keypoints.insert(ctx->orb_keypoints.end(), r.kp.begin(), r.kp.begin() + (size_t)r.n * 28);
        for (int k = 0; k < r.n; ++k) {
            float p[2];
            memcpy(p, r.kp.data() + (size_t)k * 28, 8);
            xy.push_back(p[0]); xy.push_back(p[1]);
        }
    }
    if (kp_offsets_out) memcpy(kp_offsets_out, off.data(), off.size() * sizeof(int64_t));
    static const uint8_t dummy_desc[32] = {0};
    static const float dummy_xy[2] = {0.f, 0.f};
    ctx->keep_orb_keypoints = true;
    const int rc = sfmgms_set_images(ctx, n_images, off.data(), desc.empty() ? dummy_desc : desc.data(), xy.empty() ? dummy_xy : xy.data(),
                                     sizes.data(), SFMGMS_HOST);
    ctx->keep_orb_keypoints = false;
    if (rc) { ctx->orb_keypoints.clear(); return rc; }
    return cudaStreamSynchronize(ctx->stream) == cudaSuccess ? SFMGMS_OK : fail(ctx, SFMGMS_ERR_CUDA, "upload of the image set failed");
    GUARD_END
}

int sfmgms_get_image_keypoints(sfmgms_ctx* ctx, int image, void* keypoints, int capacity, int* n_out) {
    if (!ctx) return SFMGMS_ERR_ARG;
    if (n_out) *n_out = 0;
    if (image < 0 || image >= ctx->n_images || ctx->orb_keypoints.size() != (size_t)ctx->offsets[(size_t)ctx->n_images] * 28)
        return fail(ctx, SFMGMS_ERR_STATE, "no ORB keypoints for image %d (call sfmgms_set_images_from_pixels first)", image);
    const int64_t a = ctx->offsets[(size_t)image], b = ctx->offsets[(size_t)image + 1];
    if (n_out) *n_out = (int)(b - a);
    if (b - a > capacity) return fail(ctx, SFMGMS_ERR_ARG, "capacity %d < %lld keypoints", capacity, (long long)(b - a));
    if (keypoints && b > a) memcpy(keypoints, ctx->orb_keypoints.data() + (size_t)a * 28, (size_t)(b - a) * 28);
    return SFMGMS_OK;
}

int sfmgms_match_offsets(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int64_t* match_offsets) {
    GUARD_BEGIN
    if (n_pairs < 0 || (n_pairs > 0 && !pairs) || !match_offsets) return fail(ctx, SFMGMS_ERR_ARG, "bad arguments");
    int64_t o = 0;
    for (int p = 0; p < n_pairs; ++p) {
        const int a = pairs[2 * p];
        if (a < 0 || a >= ctx->n_images || pairs[2 * p + 1] < 0 || pairs[2 * p + 1] >= ctx->n_images)
            return fail(ctx, SFMGMS_ERR_ARG, "pair %d references an image outside the set", p);
        match_offsets[p] = o;
        o += ctx->offsets[a + 1] - ctx->offsets[a];
    }
    match_offsets[n_pairs] = o;
    return SFMGMS_OK;
    GUARD_END
}

// Builds the PairDesc table of a pair list over the registered image set (device pointers, match offsets).
static int build_pair_table(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, std::vector<PairDesc>& hp, int64_t* total_out) {
    hp.assign((size_t)n_pairs, PairDesc());
    int64_t total = 0;
    for (int p = 0; p < n_pairs; ++p) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        if (a < 0 || a >= ctx->n_images || b < 0 || b >= ctx->n_images)
            return fail(ctx, SFMGMS_ERR_ARG, "pair %d references an image outside the set", p);
        PairDesc& d = hp[p];
        memset(&d, 0, sizeof d);
        d.n1 = (int)(ctx->offsets[a + 1] - ctx->offsets[a]);
        d.n2 = (int)(ctx->offsets[b + 1] - ctx->offsets[b]);
        d.desc1 = ctx->set_desc + ctx->offsets[a] * 32; d.desc2 = ctx->set_desc + ctx->offsets[b] * 32;
        d.kp1 = ctx->set_kp + ctx->offsets[a] * 2; d.kp2 = ctx->set_kp + ctx->offsets[b] * 2;
        d.w1 = ctx->sizes[2 * a]; d.h1 = ctx->sizes[2 * a + 1]; d.w2 = ctx->sizes[2 * b]; d.h2 = ctx->sizes[2 * b + 1];
        d.n_matches = d.n2 == 0 ? 0 : d.n1;
        d.match_base = total;
        d.pair_index = p; d.img1 = a; d.img2 = b;
        total += d.n1;
    }
    *total_out = total;
    return SFMGMS_OK;
}

}  // extern "C" (the runner below is C++; the entry points after it re-open the block)

// ---- chunked pair-list runner -------------------------------------------------------------------------
// One implementation behind sfmgms_match_pairs and sfmgms_match_pairs_compact.  The pair list is walked in chunks
// of at most ctx->chunk_rows match rows (and kMaxChunkPairs pairs); all per-match device scratch (keys, masks, GMS
// cell indices, decoded / compacted outputs) is sized for ONE chunk and lives in two slots, so chunk c+1 computes
// while chunk c's results are copied out.  The reference handles one pair at a time (FeatureMatchUtil.cpp:66-69),
// i.e. has no list-size limit either.
namespace {

struct PairsJob {
    const int32_t* pairs = nullptr; int n_pairs = 0;
    int with_rotation = 0, with_scale = 0; double factor = 6.0;
    int out_location = SFMGMS_HOST;
    int32_t *n_inliers = nullptr, *best_hyp = nullptr, *mask_len = nullptr;   // per pair
    int32_t *train_idx = nullptr, *dist = nullptr; uint8_t* mask = nullptr;   // per match row, list order
    bool compact = false;                                                     // compacted inliers
    int64_t* inlier_offsets = nullptr; void* matches = nullptr; float *pts1 = nullptr, *pts2 = nullptr;
    int64_t capacity = 0; int64_t* n_total = nullptr;
    // multi-GPU (host outputs only): several contexts append their chunks to ONE caller buffer through a shared
    // cursor (atomic reservation per chunk); inlier_offsets[p] then is the absolute first row of pair p (n entries)
    int64_t* shared_cursor = nullptr;
    bool async = false;   // device outputs only: enqueue everything on the context stream and return (sfmgms_wait finishes)
};

constexpr int kMaxChunkPairs = 4096;

int ensure_slot_events(sfmgms_ctx* ctx) {
    for (auto& sl : ctx->slot) {
        if (!sl.computed) CU(cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming));
        if (!sl.copied) CU(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming));
        if (ctx->timing && !sl.t0) { CU(cudaEventCreate(&sl.t0)); CU(cudaEventCreate(&sl.t1)); CU(cudaEventCreate(&sl.t2)); }
    }
    return SFMGMS_OK;
}

int run_pairs_job(sfmgms_ctx* ctx, const PairsJob& J) {
    const int n_pairs = J.n_pairs;
    if (n_pairs < 0 || (n_pairs > 0 && !J.pairs)) return fail(ctx, SFMGMS_ERR_ARG, "bad pair list");
    if (J.out_location != SFMGMS_HOST && J.out_location != SFMGMS_DEVICE) return fail(ctx, SFMGMS_ERR_ARG, "bad out_location");
    if (n_pairs > 0 && ctx->n_images == 0) return fail(ctx, SFMGMS_ERR_STATE, "sfmgms_set_images has not been called");
    if (J.compact && (J.capacity < 0 || (J.capacity > 0 && !J.matches && !J.pts1 && !J.pts2)))
        return fail(ctx, SFMGMS_ERR_ARG, "compact output: capacity > 0 needs at least one of matches / pts1 / pts2");
    if (J.n_total) *J.n_total = 0;
    if (ctx->pending.active) return fail(ctx, SFMGMS_ERR_STATE, "an asynchronous pair list is in flight: call sfmgms_wait first");
    cudaStream_t st = ctx->stream;
    const bool dev_out = (J.out_location == SFMGMS_DEVICE);
    if (J.async && !dev_out) return fail(ctx, SFMGMS_ERR_ARG, "asynchronous calls take device outputs only");
    std::vector<PairDesc> hp;
    int64_t total = 0;
    int rc = build_pair_table(ctx, J.pairs, n_pairs, hp, &total);
    if (rc) return rc;
    ctx->last_pairs.clear();
    ctx->last_results.assign((size_t)n_pairs, PairResult{0, -1, 0, 0});
    if (J.inlier_offsets && !dev_out) J.inlier_offsets[0] = 0;
    if (n_pairs == 0) {
        if (J.inlier_offsets && dev_out) CU(cudaMemsetAsync(J.inlier_offsets, 0, 8, st));
        CU(cudaStreamSynchronize(st));
        return SFMGMS_OK;
    }
    if (!ctx->d2h_stream) CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    if ((rc = ensure_slot_events(ctx))) return rc;

    // ---- chunk boundaries -------------------------------------------------------------------------------
    std::vector<int> cbeg;
    {
        int p = 0;
        while (p < n_pairs) {
            cbeg.push_back(p);
            long long rows = 0;
            int q = p;
            while (q < n_pairs && q - p < kMaxChunkPairs && (q == p || rows + hp[q].n1 <= ctx->chunk_rows)) rows += hp[q++].n1;
            p = q;
        }
        cbeg.push_back(n_pairs);
    }
    const int n_chunks = (int)cbeg.size() - 1;
    long long max_rows = 0;
    int max_cp = 0;
    for (int c = 0; c < n_chunks; ++c) {
        const long long rows = (cbeg[c + 1] < n_pairs ? hp[cbeg[c + 1]].match_base : total) - hp[cbeg[c]].match_base;
        if (rows > max_rows) max_rows = rows;
        if (cbeg[c + 1] - cbeg[c] > max_cp) max_cp = cbeg[c + 1] - cbeg[c];
    }
    // ---- buffers: one pair table / result table for the list (96 + 16 B per pair), per-chunk scratch in 2 slots ----
    const int n_slots = n_chunks > 1 ? 2 : 1;
    const bool want_idx = J.train_idx || J.dist;
    const bool stage_mask = !(dev_out && J.mask);                 // device output: kernels write the caller's mask directly
    const bool stage_idx = want_idx && !dev_out;
    const bool stage_compact = J.compact && !dev_out;
    for (int b = 0; b < n_slots; ++b) {
        auto& sl = ctx->slot[b];
        CU(sl.key.ensure((size_t)max_rows * 4 + 4));
        if (stage_mask) CU(sl.mask.ensure((size_t)max_rows + 1));
        if (stage_idx) CU(sl.out_i32.ensure((size_t)max_rows * 8 + 8));
        if (J.compact) CU(sl.coff.ensure((size_t)(max_cp + 1) * 8));
        if (stage_compact && J.matches) CU(sl.cmatch.ensure((size_t)max_rows * 16 + 16));
        if (stage_compact && (J.pts1 || J.pts2)) CU(sl.cpts.ensure((size_t)max_rows * 16 + 16));
    }
    if (J.compact) { CU(ctx->d_cbase.ensure(8)); CU(cudaMemsetAsync(ctx->d_cbase.p, 0, 8, st)); }
    for (int c = 0; c < n_chunks; ++c) {
        auto& sl = ctx->slot[c % n_slots];
        const int64_t cb = hp[cbeg[c]].match_base;
        for (int p = cbeg[c]; p < cbeg[c + 1]; ++p) {
            hp[p].key = (uint32_t*)sl.key.p + (hp[p].match_base - cb);
            hp[p].mask = stage_mask ? (uint8_t*)sl.mask.p + (hp[p].match_base - cb) : J.mask + hp[p].match_base;
            hp[p].match_base -= cb;                               // chunk-local row offset from here on
        }
    }
    std::vector<int64_t> chunk_row0((size_t)n_chunks + 1);
    {
        int64_t o = 0;
        for (int c = 0; c < n_chunks; ++c) {
            chunk_row0[c] = o;
            for (int p = cbeg[c]; p < cbeg[c + 1]; ++p) o += hp[p].n1;
        }
        chunk_row0[n_chunks] = o;
    }
    if ((rc = upload_pairs(ctx, hp))) return rc;
    CU(ctx->h_results.ensure(sizeof(PairResult) * (size_t)n_pairs));
    if (J.compact) CU(ctx->h_coff.ensure((size_t)(max_cp + 1) * 8 * 2));
    double t_ham = 0, t_gms = 0;
    int ham_launches = 0;
    const int timing = J.async ? 0 : ctx->timing;      // per-chunk events are read back by the host: not in async mode
    int64_t inl_total = 0;           // inliers of finished chunks (host outputs: also the write cursor)
    int64_t produced = 0;            // inliers this call produced (== inl_total unless a shared cursor places the chunks)
    bool overflow = false;
    int err_rc = SFMGMS_OK;

    // finalize(c): wait for chunk c, read its per-pair results, start the copies of its outputs to the caller
    auto finalize = [&](int c) -> int {
        auto& sl = ctx->slot[c % n_slots];
        const int p0 = cbeg[c], pn = cbeg[c + 1] - cbeg[c];
        CU(cudaEventSynchronize(sl.computed));
        if (timing) {
            float a = 0.f, b = 0.f;
            CU(cudaEventElapsedTime(&a, sl.t0, sl.t1)); CU(cudaEventElapsedTime(&b, sl.t1, sl.t2));
            t_ham += a; t_gms += b;
            if (timing >= 2) ctx->absorb(sl.marks);
        }
        const PairResult* hr = static_cast<const PairResult*>(ctx->h_results.p) + p0;
        int64_t chunk_inl = 0;
        if (J.shared_cursor) {     // reserve this chunk's rows in the shared output
            int64_t need = 0;
            for (int p = 0; p < pn; ++p) need += hr[p].mask_len > 0 ? hr[p].n_inliers : 0;
            inl_total = __atomic_fetch_add(J.shared_cursor, need, __ATOMIC_RELAXED);
        }
        for (int p = 0; p < pn; ++p) {
            ctx->last_results[(size_t)p0 + p] = hr[p];
            if (hr[p].status == 4 && !err_rc) err_rc = fail(ctx, SFMGMS_ERR_INDEX, "pair %d: queryIdx/trainIdx out of range", p0 + p);
            if (hr[p].status == 3 && !err_rc) err_rc = fail(ctx, SFMGMS_ERR_DOMAIN, "pair %d: matched keypoint outside [0,w)x[0,h)", p0 + p);
            if (!dev_out) {
                if (J.n_inliers) J.n_inliers[p0 + p] = hr[p].n_inliers;
                if (J.best_hyp) J.best_hyp[p0 + p] = hr[p].best_hyp;
                if (J.mask_len) J.mask_len[p0 + p] = hr[p].mask_len;
                if (J.inlier_offsets) J.inlier_offsets[p0 + p] = inl_total + chunk_inl;
            }
            chunk_inl += hr[p].mask_len > 0 ? hr[p].n_inliers : 0;
        }
        if (!dev_out) {
            const int64_t r0 = chunk_row0[c], rows = chunk_row0[c + 1] - chunk_row0[c];
            if (rows > 0) {
                int32_t* dti = (int32_t*)sl.out_i32.p;
                if (J.train_idx) CU(cudaMemcpyAsync(J.train_idx + r0, dti, (size_t)rows * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                if (J.dist) CU(cudaMemcpyAsync(J.dist + r0, dti + rows, (size_t)rows * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                if (J.mask) CU(cudaMemcpyAsync(J.mask + r0, sl.mask.p, (size_t)rows, cudaMemcpyDeviceToHost, ctx->d2h_stream));
            }
            if (J.compact && chunk_inl > 0) {
                int64_t n = chunk_inl;
                if (inl_total + n > J.capacity) { overflow = true; n = J.capacity > inl_total ? J.capacity - inl_total : 0; }
                if (n > 0) {
                    if (J.matches) CU(cudaMemcpyAsync((char*)J.matches + inl_total * ctx->compact_rec_bytes, sl.cmatch.p, (size_t)n * ctx->compact_rec_bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                    if (J.pts1) CU(cudaMemcpyAsync(J.pts1 + inl_total * 2, sl.cpts.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                    if (J.pts2) CU(cudaMemcpyAsync(J.pts2 + inl_total * 2, (char*)sl.cpts.p + (size_t)max_rows * 8, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->d2h_stream));
                }
            }
            CU(cudaEventRecord(sl.copied, ctx->d2h_stream));
            sl.copy_pending = true;
        }
        inl_total += chunk_inl;
        produced += chunk_inl;
        return SFMGMS_OK;
    };

    for (int c = 0; c < n_chunks && !err_rc; ++c) {
        auto& sl = ctx->slot[c % n_slots];
        const int p0 = cbeg[c], pn = cbeg[c + 1] - cbeg[c];
        const int64_t rows = chunk_row0[c + 1] - chunk_row0[c];
        if (sl.copy_pending) { CU(cudaStreamWaitEvent(st, sl.copied, 0)); sl.copy_pending = false; }
        if (rows > 0) {
            CU(cudaMemsetAsync(sl.key.p, 0xFF, (size_t)rows * 4, st));
            CU(cudaMemsetAsync(stage_mask ? (uint8_t*)sl.mask.p : J.mask + chunk_row0[c], 0, (size_t)rows, st));
        }
        if (timing) CU(cudaEventRecord(sl.t0, st));
        if (timing >= 2) { sl.marks.used = 0; tl_marks = &sl.marks; kmark("begin", st); }
        rc = enqueue_range(ctx, hp, p0, pn, true, true, J.with_rotation, J.with_scale, J.factor, &ham_launches,
                           timing ? sl.t1 : nullptr);
        if (rc) { tl_marks = nullptr; err_rc = rc; break; }
        if (timing) CU(cudaEventRecord(sl.t2, st));
        if (rows > 0 && want_idx) {
            int32_t* dti = dev_out ? (J.train_idx ? J.train_idx + chunk_row0[c] : nullptr) : (J.train_idx ? (int32_t*)sl.out_i32.p : nullptr);
            int32_t* ddi = dev_out ? (J.dist ? J.dist + chunk_row0[c] : nullptr) : (J.dist ? (int32_t*)sl.out_i32.p + rows : nullptr);
            decode_keys_kernel<<<(unsigned)((rows + 255) / 256 > 8192 ? 8192 : (rows + 255) / 256), 256, 0, st>>>(
                (const uint32_t*)sl.key.p, rows, dti, ddi);
            ctx->launches++; kmark("decode_keys", st);
        }
        if (J.compact) {
            // host output: chunk-local offsets into the slot's staging (base reset per chunk); device output: the
            // caller's buffers and offsets array directly, one running base
            if (!dev_out) CU(cudaMemsetAsync(ctx->d_cbase.p, 0, 8, st));
            long long* doff = dev_out && J.inlier_offsets ? (long long*)J.inlier_offsets + p0 : (long long*)sl.coff.p;
            DMatchRec* dm = dev_out ? (DMatchRec*)J.matches : (J.matches ? (DMatchRec*)sl.cmatch.p : nullptr);
            float* dp1 = dev_out ? J.pts1 : (J.pts1 ? (float*)sl.cpts.p : nullptr);
            float* dp2 = dev_out ? J.pts2 : (J.pts2 ? (float*)((char*)sl.cpts.p + (size_t)max_rows * 8) : nullptr);
            ctx->launches += launch_gms_compact(static_cast<const PairDesc*>(ctx->d_pairs.p) + p0, static_cast<const PairResult*>(ctx->d_results.p) + p0,
                                                pn, (long long*)ctx->d_cbase.p, doff, dev_out ? J.capacity : rows, dm, dp1, dp2, st,
                                                ctx->compact_rec_bytes == 8);
            CU(cudaGetLastError());
        }
        tl_marks = nullptr;
        CU(cudaMemcpyAsync(static_cast<PairResult*>(ctx->h_results.p) + p0, static_cast<const PairResult*>(ctx->d_results.p) + p0,
                           sizeof(PairResult) * (size_t)pn, cudaMemcpyDeviceToHost, st));
        CU(cudaEventRecord(sl.computed, st));
        if (!J.async && c > 0 && (rc = finalize(c - 1))) { err_rc = rc; break; }
    }
    if (J.async) {
        if (err_rc) { cudaStreamSynchronize(st); tc_reset_arena(ctx->tc); return err_rc; }
        export_results_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(static_cast<const PairResult*>(ctx->d_results.p), n_pairs, J.n_inliers,
                                                                    J.best_hyp, J.mask_len);
        ctx->launches++;
        CU(cudaMemcpyAsync(ctx->h_results.p, ctx->d_results.p, sizeof(PairResult) * (size_t)n_pairs, cudaMemcpyDeviceToHost, st));
        CU(cudaGetLastError());
        ctx->pending.active = true; ctx->pending.n_pairs = n_pairs; ctx->pending.compact = J.compact; ctx->pending.capacity = J.capacity;
        if (n_chunks == 1) ctx->last_pairs = hp;
        return SFMGMS_OK;
    }
    if (!err_rc && (rc = finalize(n_chunks - 1))) err_rc = rc;
    cudaStreamSynchronize(st);
    cudaStreamSynchronize(ctx->d2h_stream);
    for (auto& sl : ctx->slot) sl.copy_pending = false;
    tc_reset_arena(ctx->tc);
    if (ctx->timing) { ctx->last_ms[0] = t_ham; ctx->last_ms[1] = t_gms; ctx->last_ms[2] = ham_launches; }
    if (err_rc) return err_rc;
    if (dev_out) {   // per-pair summaries for device callers (small; after the run)
        std::vector<int32_t> tmp((size_t)n_pairs);
        int32_t* outs[3] = {J.n_inliers, J.best_hyp, J.mask_len};
        for (int k = 0; k < 3; ++k) {
            if (!outs[k]) continue;
            for (int p = 0; p < n_pairs; ++p)
                tmp[p] = k == 0 ? ctx->last_results[p].n_inliers : k == 1 ? ctx->last_results[p].best_hyp : ctx->last_results[p].mask_len;
            CU(cudaMemcpy(outs[k], tmp.data(), (size_t)n_pairs * 4, cudaMemcpyHostToDevice));
        }
    } else if (J.inlier_offsets && !J.shared_cursor) {
        J.inlier_offsets[n_pairs] = inl_total;
    }
    if (J.n_total) *J.n_total = produced;
    if (n_chunks == 1) ctx->last_pairs = hp;     // sfmgms_inlier_points: only while the chunk's buffers are intact
    if (J.compact && (overflow || (dev_out && inl_total > J.capacity)))
        return fail(ctx, SFMGMS_ERR_CAPACITY, "compact output needs %lld rows, capacity is %lld",
                    (long long)(J.shared_cursor ? __atomic_load_n(J.shared_cursor, __ATOMIC_RELAXED) : inl_total), (long long)J.capacity);
    return SFMGMS_OK;
}

}  // namespace

// used by multi.cu: one shard of a multi-GPU compact run (host outputs appended through the shared cursor)
namespace sfmgms {
int match_pairs_compact_shared(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                               double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int64_t* inlier_begin,
                               void* matches, float* pts1, float* pts2, int64_t capacity, int64_t* shared_cursor) {
    GUARD_BEGIN
    PairsJob J;
    J.pairs = pairs; J.n_pairs = n_pairs; J.with_rotation = with_rotation; J.with_scale = with_scale; J.factor = threshold_factor;
    J.out_location = SFMGMS_HOST; J.n_inliers = n_inliers; J.best_hyp = best_hyp;
    J.compact = true; J.inlier_offsets = inlier_begin; J.matches = matches; J.pts1 = pts1; J.pts2 = pts2;
    J.capacity = capacity; J.shared_cursor = shared_cursor;
    return run_pairs_job(ctx, J);
    GUARD_END
}
}  // namespace sfmgms

extern "C" {

int sfmgms_match_pairs(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                       double threshold_factor, int out_location, int32_t* n_inliers, int32_t* best_hyp,
                       int32_t* mask_len, int32_t* train_idx, int32_t* dist, uint8_t* mask) {
    GUARD_BEGIN
    PairsJob J;
    J.pairs = pairs; J.n_pairs = n_pairs; J.with_rotation = with_rotation; J.with_scale = with_scale; J.factor = threshold_factor;
    J.out_location = out_location; J.n_inliers = n_inliers; J.best_hyp = best_hyp; J.mask_len = mask_len;
    J.train_idx = train_idx; J.dist = dist; J.mask = mask;
    return run_pairs_job(ctx, J);
    GUARD_END
}

int sfmgms_match_pairs_async(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                             double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int32_t* mask_len, int32_t* train_idx,
                             int32_t* dist, uint8_t* mask) {
    GUARD_BEGIN
    PairsJob J;
    J.pairs = pairs; J.n_pairs = n_pairs; J.with_rotation = with_rotation; J.with_scale = with_scale; J.factor = threshold_factor;
    J.out_location = SFMGMS_DEVICE; J.n_inliers = n_inliers; J.best_hyp = best_hyp; J.mask_len = mask_len;
    J.train_idx = train_idx; J.dist = dist; J.mask = mask; J.async = true;
    return run_pairs_job(ctx, J);
    GUARD_END
}

int sfmgms_match_pairs_compact_async(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                                     double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int64_t* inlier_offsets,
                                     void* matches, float* pts1, float* pts2, int64_t capacity) {
    GUARD_BEGIN
    PairsJob J;
    J.pairs = pairs; J.n_pairs = n_pairs; J.with_rotation = with_rotation; J.with_scale = with_scale; J.factor = threshold_factor;
    J.out_location = SFMGMS_DEVICE; J.n_inliers = n_inliers; J.best_hyp = best_hyp;
    J.compact = true; J.inlier_offsets = inlier_offsets; J.matches = matches; J.pts1 = pts1; J.pts2 = pts2; J.capacity = capacity;
    J.async = true;
    return run_pairs_job(ctx, J);
    GUARD_END
}

// Completes an asynchronous pair list: waits for the context stream, turns device-side status flags into error codes.
int sfmgms_wait(sfmgms_ctx* ctx, int64_t* n_total) {
    GUARD_BEGIN
    if (n_total) *n_total = 0;
    CU(cudaStreamSynchronize(ctx->stream));
    if (!ctx->pending.active) return SFMGMS_OK;
    ctx->pending.active = false;
    tc_reset_arena(ctx->tc);
    const int n = ctx->pending.n_pairs;
    const PairResult* hr = static_cast<const PairResult*>(ctx->h_results.p);
    ctx->last_results.assign(hr, hr + n);
    long long total = 0;
    for (int p = 0; p < n; ++p) {
        if (hr[p].status == 4) return fail(ctx, SFMGMS_ERR_INDEX, "pair %d: queryIdx/trainIdx out of range", p);
        if (hr[p].status == 3) return fail(ctx, SFMGMS_ERR_DOMAIN, "pair %d: matched keypoint outside [0,w)x[0,h)", p);
        total += hr[p].mask_len > 0 ? hr[p].n_inliers : 0;
    }
    if (n_total) *n_total = total;
    if (ctx->pending.compact && total > ctx->pending.capacity)
        return fail(ctx, SFMGMS_ERR_CAPACITY, "compact output needs %lld rows, capacity is %lld", total, ctx->pending.capacity);
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_match_pairs_compact(sfmgms_ctx* ctx, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                               double threshold_factor, int out_location, int32_t* n_inliers, int32_t* best_hyp,
                               int64_t* inlier_offsets, void* matches, float* pts1, float* pts2, int64_t capacity,
                               int64_t* n_total) {
    GUARD_BEGIN
    PairsJob J;
    J.pairs = pairs; J.n_pairs = n_pairs; J.with_rotation = with_rotation; J.with_scale = with_scale; J.factor = threshold_factor;
    J.out_location = out_location; J.n_inliers = n_inliers; J.best_hyp = best_hyp;
    J.compact = true; J.inlier_offsets = inlier_offsets; J.matches = matches; J.pts1 = pts1; J.pts2 = pts2;
    J.capacity = capacity; J.n_total = n_total;
    return run_pairs_job(ctx, J);
    GUARD_END
}

static int get_event(sfmgms_ctx* ctx, size_t k, cudaEvent_t* out) {
    while (ctx->events.size() <= k) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->events.push_back(e);
    }
    *out = ctx->events[k];
    return SFMGMS_OK;
}

int sfmgms_match_image_set(sfmgms_ctx* ctx, int n_images, const int64_t* kp_offsets, const uint8_t* desc,
                           const float* kp_xy, const int32_t* sizes_wh, const int32_t* pairs, int n_pairs,
                           int with_rotation, int with_scale, double threshold_factor, int32_t* n_inliers,
                           int32_t* best_hyp, int32_t* mask_len, int32_t* train_idx, int32_t* dist, uint8_t* mask) {
    GUARD_BEGIN
    if (n_images < 0 || !kp_offsets || !sizes_wh) return fail(ctx, SFMGMS_ERR_ARG, "bad image-set arguments");
    if (n_pairs < 0 || (n_pairs > 0 && !pairs)) return fail(ctx, SFMGMS_ERR_ARG, "bad pair list");
    if (kp_offsets[0] != 0) return fail(ctx, SFMGMS_ERR_ARG, "kp_offsets[0] must be 0");
    for (int i = 0; i < n_images; ++i) {
        const int64_t n = kp_offsets[i + 1] - kp_offsets[i];
        if (n < 0) return fail(ctx, SFMGMS_ERR_ARG, "kp_offsets not monotone at image %d", i);
        if (n >= SFMGMS_MAX_TRAIN_ROWS) return fail(ctx, SFMGMS_ERR_TRAIN_ROWS, "image %d has %lld rows >= 2^18", i, (long long)n);
        if (sizes_wh[2 * i] <= 0 || sizes_wh[2 * i + 1] <= 0) return fail(ctx, SFMGMS_ERR_ARG, "image %d has a non-positive size", i);
    }
    const int64_t total_rows = kp_offsets[n_images];
    if (total_rows > 0 && (!desc || !kp_xy)) return fail(ctx, SFMGMS_ERR_ARG, "null descriptor/keypoint pointer");
    cudaStream_t st = ctx->stream;
    if (!ctx->h2d_stream) CU(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    if (!ctx->d2h_stream) CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    CU(cudaStreamSynchronize(st));   // nothing of an earlier call may still read the buffers re-used below
    CU(ctx->d_set_desc.ensure((size_t)total_rows * 32 + 32));
    CU(ctx->d_set_kp.ensure((size_t)total_rows * 8 + 8));
    ctx->set_desc = (const uint8_t*)ctx->d_set_desc.p; ctx->set_kp = (const float*)ctx->d_set_kp.p;
    ctx->n_images = n_images;
    ctx->orb_keypoints.clear();
    ctx->offsets.assign(kp_offsets, kp_offsets + n_images + 1);
    ctx->sizes.assign(sizes_wh, sizes_wh + 2 * (size_t)n_images);
    ctx->set_version++;
    tc_invalidate(ctx->tc);
    tc_pin_span(ctx->tc, nullptr, 0);   // the set arrives chunk by chunk here: operands are derived per sub-batch span
    ctx->last_pairs.clear();
    // every early return below must leave no copy in flight into the caller's buffers
    auto drain = [&]() { cudaStreamSynchronize(ctx->h2d_stream); cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->d2h_stream); tc_reset_arena(ctx->tc); };

    static const bool trace = getenv("SFMGMS_TRACE") != nullptr;   // developer aid: print the stream timeline
    auto now_ms = []() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_enter = now_ms();
    double t_h2d0 = 0, t_enq = 0, t_fin = 0;
    std::vector<cudaEvent_t> tr_h2d, tr_cmp;
    std::vector<std::pair<cudaEvent_t, const char*>> tr_marks;
    auto mark = [&](const char* label) { if (trace) { cudaEvent_t t; cudaEventCreate(&t); cudaEventRecord(t, st); tr_marks.push_back({t, label}); } };
    cudaEvent_t tr0 = nullptr;
    if (trace) { cudaEventCreate(&tr0); cudaEventRecord(tr0, ctx->h2d_stream); t_h2d0 = now_ms(); }
    // ---- image chunks (H2D on the copy stream, one event per chunk).  Chunks are ENQUEUED lazily, a little
    // ahead of the pairs that need them: the copy engine serves H2D requests of all streams in FIFO order, so the
    // small table uploads of a sub-batch (pair table, work units) must not queue behind the whole image set. ----
    const size_t kChunkBytes = 12u << 20;
    const int kLookahead = 2;
    std::vector<int> chunk_of_image((size_t)n_images, 0);
    std::vector<int> chunk_first;   // first image of each chunk; chunk c = images [chunk_first[c], chunk_first[c+1])
    {
        int i0 = 0;
        while (i0 < n_images) {
            int i1 = i0;
            size_t bytes = 0;
            while (i1 < n_images && (bytes == 0 || bytes + (size_t)(kp_offsets[i1 + 1] - kp_offsets[i1]) * 40 <= kChunkBytes)) {
                bytes += (size_t)(kp_offsets[i1 + 1] - kp_offsets[i1]) * 40;
                chunk_of_image[i1] = (int)chunk_first.size();
                ++i1;
            }
            chunk_first.push_back(i0);
            i0 = i1;
        }
        chunk_first.push_back(n_images);
    }
    const int n_chunks = (int)chunk_first.size() - 1;
    for (int c = 0; c < n_chunks; ++c) { cudaEvent_t e; int rc0 = get_event(ctx, (size_t)c, &e); if (rc0) return rc0; }
    int next_chunk = 0;
    auto issue_chunks_upto = [&](int c_last) -> int {
        for (; next_chunk <= c_last && next_chunk < n_chunks; ++next_chunk) {
            const int64_t r0 = kp_offsets[chunk_first[next_chunk]], r1 = kp_offsets[chunk_first[next_chunk + 1]];
            if (r1 > r0) {
                CU(cudaMemcpyAsync((uint8_t*)ctx->d_set_desc.p + r0 * 32, desc + r0 * 32, (size_t)(r1 - r0) * 32, cudaMemcpyHostToDevice, ctx->h2d_stream));
                CU(cudaMemcpyAsync((float*)ctx->d_set_kp.p + r0 * 2, kp_xy + r0 * 2, (size_t)(r1 - r0) * 8, cudaMemcpyHostToDevice, ctx->h2d_stream));
            }
            CU(cudaEventRecord(ctx->events[(size_t)next_chunk], ctx->h2d_stream));
            if (trace) { cudaEvent_t t; cudaEventCreate(&t); cudaEventRecord(t, ctx->h2d_stream); tr_h2d.push_back(t); }
        }
        return SFMGMS_OK;
    };
    // ---- pair table ----
    std::vector<PairDesc> hp;
    int64_t total = 0;
    int rc = build_pair_table(ctx, pairs, n_pairs, hp, &total);
    if (rc) { cudaStreamSynchronize(ctx->h2d_stream); return rc; }
    CU(ctx->d_key.ensure((size_t)total * 4 + 4));
    CU(ctx->d_mask.ensure((size_t)total + 1));
    CU(ctx->d_out_i32.ensure((size_t)total * 8 + 8));
    uint8_t* dmask = (uint8_t*)ctx->d_mask.p;
    int32_t* dti = (int32_t*)ctx->d_out_i32.p;
    int32_t* ddi = dti + total;
    for (int p = 0; p < n_pairs; ++p) {
        hp[p].key = (uint32_t*)ctx->d_key.p + hp[p].match_base;
        hp[p].mask = dmask + hp[p].match_base;
    }
    if (total) {
        CU(cudaMemsetAsync(ctx->d_key.p, 0xFF, (size_t)total * 4, st));
        CU(cudaMemsetAsync(dmask, 0, (size_t)total, st));
    }
    mark("memsets-queued");
    rc = upload_pairs(ctx, hp);
    if (rc) { cudaStreamSynchronize(ctx->h2d_stream); return rc; }
    mark("pairs-uploaded");
    if (ctx->timing) { CU(cudaEventRecord(ctx->ev[0], st)); CU(cudaEventRecord(ctx->ev[1], st)); }
    // ---- sub-batches in list order: wait for the chunk(s) they need, compute, hand results to the D2H stream ----
    // Sub-batch sizes DECREASE along the list: compute is faster than the H2D transfer, so every sub-batch ends up
    // waiting for its images and the time after the last byte has arrived is one sub-batch of compute — keep that
    // one small (1/32 of the list); the early ones are large to amortise launch overheads.
    std::vector<int> range_begin;
    {
        static const int kNum[] = {8, 8, 6, 4, 3, 2, 1};   // 32nds of the pair list
        int acc = 0;
        range_begin.push_back(0);
        if (n_pairs >= 64) {
            for (int k = 0; k < 6; ++k) {
                acc += kNum[k];
                const int b = (int)((long long)n_pairs * acc / 32);
                if (b > range_begin.back() && b < n_pairs) range_begin.push_back(b);
            }
        } else if (n_pairs >= 8) {
            for (int k = 1; k < 4; ++k) range_begin.push_back(n_pairs * k / 4);
        }
        range_begin.push_back(n_pairs);
    }
    int ham_launches = 0;
    size_t ev_k = (size_t)n_chunks;
    for (size_t ri = 0; ri + 1 < range_begin.size(); ++ri) {
        const int r0 = range_begin[ri], rn = range_begin[ri + 1] - r0;
        if (rn <= 0) continue;
        int need = 0;
        for (int p = r0; p < r0 + rn; ++p) {
            const int c = chunk_of_image[hp[p].img1] > chunk_of_image[hp[p].img2] ? chunk_of_image[hp[p].img1] : chunk_of_image[hp[p].img2];
            if (c > need) need = c;
        }
        if (trace) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); fprintf(stderr, "[sfmgms trace] host t=%.3f ms: range at pair %d needs chunk %d\n", ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6, r0, need); }
        if ((rc = issue_chunks_upto(need + kLookahead))) { drain(); return rc; }
        mark("before-wait");
        CU(cudaStreamWaitEvent(st, ctx->events[(size_t)need], 0));
        mark("after-wait");
        rc = enqueue_range(ctx, hp, r0, rn, true, true, with_rotation, with_scale, threshold_factor, &ham_launches);
        if (rc) { drain(); return rc; }
        const int64_t m0 = hp[r0].match_base;
        const int64_t m1 = (r0 + rn < n_pairs) ? hp[r0 + rn].match_base : total;
        if (m1 > m0) {
            if (train_idx || dist) {
                decode_keys_kernel<<<(unsigned)(((m1 - m0) + 255) / 256 > 4096 ? 4096 : ((m1 - m0) + 255) / 256), 256, 0, st>>>(
                    (const uint32_t*)ctx->d_key.p + m0, m1 - m0, train_idx ? dti + m0 : nullptr, dist ? ddi + m0 : nullptr);
                ctx->launches++;
            }
            cudaEvent_t e;
            rc = get_event(ctx, ev_k++, &e);
            if (rc) { drain(); return rc; }
            CU(cudaEventRecord(e, st));
            if (trace) { cudaEvent_t t; cudaEventCreate(&t); cudaEventRecord(t, st); tr_cmp.push_back(t); }
            CU(cudaStreamWaitEvent(ctx->d2h_stream, e, 0));
            if (train_idx) CU(cudaMemcpyAsync(train_idx + m0, dti + m0, (size_t)(m1 - m0) * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
            if (dist) CU(cudaMemcpyAsync(dist + m0, ddi + m0, (size_t)(m1 - m0) * 4, cudaMemcpyDeviceToHost, ctx->d2h_stream));
            if (mask) CU(cudaMemcpyAsync(mask + m0, dmask + m0, (size_t)(m1 - m0), cudaMemcpyDeviceToHost, ctx->d2h_stream));
        }
    }
    {
        const int rc2 = issue_chunks_upto(n_chunks - 1);   // images no pair referenced still belong to the registered set
        if (rc2) { drain(); return rc2; }
    }
    t_enq = now_ms();
    rc = finish_batch(ctx, n_pairs, true);
    t_fin = now_ms();
    drain();
    ctx->last_pairs = hp;
    if (trace) {
        cudaEvent_t tend; cudaEventCreate(&tend); cudaEventRecord(tend, ctx->d2h_stream); cudaEventSynchronize(tend);
        float ms;
        fprintf(stderr, "[sfmgms trace] h2d chunks done at ms:");
        for (cudaEvent_t t : tr_h2d) { cudaEventElapsedTime(&ms, tr0, t); fprintf(stderr, " %.2f", ms); cudaEventDestroy(t); }
        fprintf(stderr, "\n[sfmgms trace] marks:");
        for (auto& m : tr_marks) { cudaEventElapsedTime(&ms, tr0, m.first); fprintf(stderr, " %s=%.2f", m.second, ms); cudaEventDestroy(m.first); }
        fprintf(stderr, "\n[sfmgms trace] compute ranges done at ms:");
        for (cudaEvent_t t : tr_cmp) { cudaEventElapsedTime(&ms, tr0, t); fprintf(stderr, " %.2f", ms); cudaEventDestroy(t); }
        cudaEventElapsedTime(&ms, tr0, tend);
        fprintf(stderr, "\n[sfmgms trace] d2h done at ms: %.2f\n", ms);
        fprintf(stderr, "[sfmgms trace] host: enter->first h2d %.2f ms, ->all enqueued %.2f, ->finish_batch done %.2f, ->now %.2f\n",
                t_h2d0 - t_enter, t_enq - t_enter, t_fin - t_enter, now_ms() - t_enter);
        cudaEventDestroy(tend); cudaEventDestroy(tr0);
    }
    if (rc) return rc;
    if (ctx->timing) { ctx->last_ms[0] = 0; ctx->last_ms[1] = 0; ctx->last_ms[2] = ham_launches; }
    for (int p = 0; p < n_pairs; ++p) {
        if (n_inliers) n_inliers[p] = ctx->last_results[p].n_inliers;
        if (best_hyp) best_hyp[p] = ctx->last_results[p].best_hyp;
        if (mask_len) mask_len[p] = ctx->last_results[p].mask_len;
    }
    return SFMGMS_OK;
    GUARD_END
}

int sfmgms_inlier_points(sfmgms_ctx* ctx, int pair_index, float* pts1, float* pts2, int capacity, int* n_out) {
    GUARD_BEGIN
    if (capacity < 0 || !n_out || (capacity > 0 && (!pts1 || !pts2))) return fail(ctx, SFMGMS_ERR_ARG, "bad arguments");
    if (ctx->last_pairs.empty() && !ctx->last_results.empty())
        return fail(ctx, SFMGMS_ERR_STATE, "the last pair list ran in several chunks; use sfmgms_match_pairs_compact for its coordinates");
    if (pair_index < 0 || pair_index >= (int)ctx->last_pairs.size()) return fail(ctx, SFMGMS_ERR_STATE, "no such pair in the last batch");
    const PairDesc& pd = ctx->last_pairs[pair_index];
    const PairResult& r = ctx->last_results[pair_index];
    if (r.mask_len == 0 || pd.n_matches == 0 || r.n_inliers == 0) { *n_out = 0; return SFMGMS_OK; }
    // a one-pair view of the batched compaction (same kernels as sfmgms_match_pairs_compact)
    cudaStream_t st = ctx->stream;
    CU(ctx->d_pts.ensure((size_t)capacity * 16 + 64));
    float* d1 = (float*)ctx->d_pts.p;
    float* d2 = d1 + 2 * (size_t)capacity;
    long long* dbase = (long long*)(d2 + 2 * (size_t)capacity);
    CU(cudaMemsetAsync(dbase, 0, 8, st));
    ctx->launches += launch_gms_compact(static_cast<const PairDesc*>(ctx->d_pairs.p) + pair_index,
                                        static_cast<const PairResult*>(ctx->d_results.p) + pair_index, 1, dbase, dbase + 1,
                                        capacity, nullptr, d1, d2, st);
    const int m = r.n_inliers < capacity ? r.n_inliers : capacity;
    if (m > 0) {
        CU(cudaMemcpyAsync(pts1, d1, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(pts2, d2, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    *n_out = r.n_inliers;
    return SFMGMS_OK;
    GUARD_END
}

// All 40 hypotheses of one pair (diagnostic / parity entry point): the inlier count GMSMatcher::run returns for
// every (scale, rotation), scale-major — what getInlierMask compares (DLL @VA 0x180047dc0).
int sfmgms_gms_hypotheses(sfmgms_ctx* ctx, int w1, int h1, int w2, int h2, const void* kp1, int n1, int kp1_stride_bytes,
                          const void* kp2, int n2, int kp2_stride_bytes, const int32_t* query_idx, const int32_t* train_idx,
                          int idx_stride_bytes, int n_matches, double threshold_factor, int32_t* counts) {
    GUARD_BEGIN
    if (!counts) return fail(ctx, SFMGMS_ERR_ARG, "counts is null");
    int rc = sfmgms_gms(ctx, w1, h1, w2, h2, kp1, n1, kp1_stride_bytes, kp2, n2, kp2_stride_bytes, query_idx, train_idx,
                        idx_stride_bytes, n_matches, 1, 1, threshold_factor, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    const size_t off = gms_counts_offset_words(kNumScales, kNumRot, n_matches, ctx->gms_dense);
    CU(cudaMemcpy(counts, static_cast<const int32_t*>(ctx->d_hist.p) + off, kMaxHyp * 4, cudaMemcpyDeviceToHost));
    return SFMGMS_OK;
    GUARD_END
}

// "name:total_ms:launches;..." of every kernel timed since SFMGMS_OPT_TIMING was set to 2 (resets the accumulators)
int sfmgms_kernel_times(sfmgms_ctx* ctx, char* buf, int buf_len) {
    GUARD_BEGIN
    if (!buf || buf_len <= 0) return fail(ctx, SFMGMS_ERR_ARG, "bad buffer");
    std::string o;
    for (const auto& k : ctx->kernel_times) {
        char tmp[160];
        snprintf(tmp, sizeof tmp, "%s:%.6f:%lld;", k.name.c_str(), k.ms, k.n);
        o += tmp;
    }
    if ((int)o.size() + 1 > buf_len) return fail(ctx, SFMGMS_ERR_CAPACITY, "buffer of %d bytes too small (%zu needed)", buf_len, o.size() + 1);
    memcpy(buf, o.c_str(), o.size() + 1);
    ctx->kernel_times.clear();
    return SFMGMS_OK;
    GUARD_END
}

int64_t sfmgms_device_bytes(const sfmgms_ctx* ctx) {
    if (!ctx) return 0;
    size_t t = 0;
    const DevBuf* bufs[] = {&ctx->d_pairs, &ctx->d_results, &ctx->d_key, &ctx->d_mask, &ctx->d_hist, &ctx->d_msc, &ctx->d_q,
                            &ctx->d_kp1, &ctx->d_kp2, &ctx->d_mq, &ctx->d_mt, &ctx->d_out_i32, &ctx->d_pts,
                            &ctx->d_set_desc, &ctx->d_set_kp, &ctx->d_cbase};
    for (const DevBuf* b : bufs) t += b->cap;
    for (const auto& sl : ctx->slot) t += sl.key.cap + sl.mask.cap + sl.out_i32.cap + sl.cmatch.cap + sl.cpts.cap + sl.coff.cap;
    t += ctx->tc.ops_cap + ctx->tc.work_cap;
    return (int64_t)t;
}

int sfmgms_host_alloc(size_t bytes, void** out) {
    if (!out) return SFMGMS_ERR_ARG;
    *out = nullptr;
    if (bytes == 0) return SFMGMS_OK;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return SFMGMS_ERR_CUDA; }
    return SFMGMS_OK;
}

int sfmgms_host_free(void* p) {
    if (!p) return SFMGMS_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) { cudaGetLastError(); return SFMGMS_ERR_CUDA; }
    return SFMGMS_OK;
}

}  // extern "C"
