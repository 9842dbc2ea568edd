// multi.cu — multi-GPU entry points of the C ABI (include/sfmgms.h, "multi-GPU" section).
//
// ONE process, one host thread per GPU, `ncclCommInitAll` over the GPUs of one box; the shared
// descriptor + keypoint set goes host -> GPU 0 once and from there to every other GPU with ONE ncclBroadcast per array
// (NVLink 5 / NVSwitch); image pairs are independent, so the pair list is cut into contiguous shards (one per GPU) and
// there is NO collective on the result path: every GPU's results go straight into the caller's host buffers.
// The reference loop this serves is the per-pair call of FeatureMatchUtil.cpp:66-69 driven over an image sequence
// (main.cpp:32,39,47; SfMUtil.cpp:16-18).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the single-GPU library keeps no link-time dependency on it, and
// inside a process that already carries an NCCL (e.g. PyTorch's) the same copy is reused.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sfmgms.h"
#include "capi_internal.h"

namespace {

typedef void* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclUint8 = 1;   // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;
    bool load(std::string& why) {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { why = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        Broadcast = reinterpret_cast<decltype(Broadcast)>(dlsym(lib, "ncclBroadcast"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        GetVersion = reinterpret_cast<decltype(GetVersion)>(dlsym(lib, "ncclGetVersion"));
        if (!CommInitAll || !CommDestroy || !Broadcast || !GroupStart || !GroupEnd || !GetErrorString) {
            why = "libnccl.so.2 lacks a required symbol";
            return false;
        }
        return true;
    }
};

char g_multi_create_error[512] = "";

}  // namespace

struct sfmgms_multi {
    int n = 0;
    std::vector<int> devices;
    std::vector<sfmgms_ctx*> ctx;
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;          // broadcast streams, one per device
    std::vector<void*> d_desc, d_kp;
    std::vector<size_t> cap_desc, cap_kp;
    NcclApi nccl;
    char err[512] = "";
    int n_images = 0;
    std::vector<int64_t> offsets;
    double last_broadcast_ms = 0;
    std::vector<int> last_shard_begin;          // pair index where each GPU's shard starts (n + 1 entries)
};

namespace {

int mfail(sfmgms_multi* m, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    if (m) vsnprintf(m->err, sizeof m->err, fmt, ap);
    else vsnprintf(g_multi_create_error, sizeof g_multi_create_error, fmt, ap);
    va_end(ap);
    return code;
}

#define MCU(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e__ = (call);                                                                          \
        if (e__ != cudaSuccess)                                                                            \
            return mfail(m, SFMGMS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define MNCCL(call)                                                                                        \
    do {                                                                                                   \
        ncclResult_t r__ = (call);                                                                         \
        if (r__ != 0) return mfail(m, SFMGMS_ERR_NCCL, "%s failed: %s", #call, m->nccl.GetErrorString(r__)); \
    } while (0)

struct DevSwitch {
    int prev = -1;
    DevSwitch() { cudaGetDevice(&prev); }
    ~DevSwitch() { if (prev >= 0) cudaSetDevice(prev); }
};

// contiguous shards with (nearly) equal work: work of a pair = n1 * n2 distance evaluations (+ its rows for GMS)
void cut_shards(const sfmgms_multi* m, const int32_t* pairs, int n_pairs, std::vector<int>& begin) {
    std::vector<double> w((size_t)n_pairs + 1, 0.0);
    for (int p = 0; p < n_pairs; ++p) {
        const int a = pairs[2 * p], b = pairs[2 * p + 1];
        double n1 = 0, n2 = 0;
        if (a >= 0 && a < m->n_images && b >= 0 && b < m->n_images) {
            n1 = (double)(m->offsets[(size_t)a + 1] - m->offsets[(size_t)a]);
            n2 = (double)(m->offsets[(size_t)b + 1] - m->offsets[(size_t)b]);
        }
        w[(size_t)p + 1] = w[(size_t)p] + n1 * n2 + 64.0 * n1 + 1.0;
    }
    begin.assign((size_t)m->n + 1, n_pairs);
    begin[0] = 0;
    int p = 0;
    for (int g = 1; g < m->n; ++g) {
        const double target = w[(size_t)n_pairs] * g / m->n;
        while (p < n_pairs && w[(size_t)p] < target) ++p;
        begin[(size_t)g] = p;
    }
}

}  // namespace

extern "C" {

int sfmgms_multi_create(sfmgms_multi** out, const int* devices, int n_devices) {
    if (!out) return mfail(nullptr, SFMGMS_ERR_ARG, "out is null");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
        return mfail(nullptr, SFMGMS_ERR_CUDA, "no CUDA device; this library has no CPU fallback");
    if (n_devices <= 0) n_devices = count;                     // all visible GPUs
    if (n_devices > count && !devices) return mfail(nullptr, SFMGMS_ERR_ARG, "%d devices requested, %d visible", n_devices, count);
    sfmgms_multi* m = new (std::nothrow) sfmgms_multi();
    if (!m) return mfail(nullptr, SFMGMS_ERR_ARG, "host allocation failed");
    DevSwitch keep;
    try {
        m->n = n_devices;
        for (int i = 0; i < n_devices; ++i) {
            const int d = devices ? devices[i] : i;
            if (d < 0 || d >= count) { delete m; return mfail(nullptr, SFMGMS_ERR_ARG, "device %d out of range [0,%d)", d, count); }
            for (int j = 0; j < i; ++j)
                if (m->devices[(size_t)j] == d) { delete m; return mfail(nullptr, SFMGMS_ERR_ARG, "device %d listed twice", d); }
            m->devices.push_back(d);
        }
        m->ctx.assign((size_t)n_devices, nullptr);
        m->streams.assign((size_t)n_devices, nullptr);
        m->d_desc.assign((size_t)n_devices, nullptr); m->d_kp.assign((size_t)n_devices, nullptr);
        m->cap_desc.assign((size_t)n_devices, 0); m->cap_kp.assign((size_t)n_devices, 0);
        for (int i = 0; i < n_devices; ++i) {
            const int rc = sfmgms_create(&m->ctx[(size_t)i], m->devices[(size_t)i]);
            if (rc) {
                mfail(nullptr, rc, "device %d: %s", m->devices[(size_t)i], sfmgms_last_error(nullptr));
                sfmgms_multi_destroy(m);
                return rc;
            }
            cudaSetDevice(m->devices[(size_t)i]);
            if (cudaStreamCreateWithFlags(&m->streams[(size_t)i], cudaStreamNonBlocking) != cudaSuccess) {
                sfmgms_multi_destroy(m);
                return mfail(nullptr, SFMGMS_ERR_CUDA, "cudaStreamCreate failed on device %d", m->devices[(size_t)i]);
            }
        }
        if (n_devices > 1) {
            std::string why;
            if (!m->nccl.load(why)) { sfmgms_multi_destroy(m); return mfail(nullptr, SFMGMS_ERR_NCCL, "%s", why.c_str()); }
            m->comms.assign((size_t)n_devices, nullptr);
            const ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), n_devices, m->devices.data());
            if (r != 0) {
                mfail(nullptr, SFMGMS_ERR_NCCL, "ncclCommInitAll failed: %s", m->nccl.GetErrorString(r));
                m->comms.clear();
                sfmgms_multi_destroy(m);
                return SFMGMS_ERR_NCCL;
            }
        }
    } catch (...) {
        sfmgms_multi_destroy(m);
        return mfail(nullptr, SFMGMS_ERR_ARG, "host allocation failed");
    }
    *out = m;
    return SFMGMS_OK;
}

void sfmgms_multi_destroy(sfmgms_multi* m) {
    if (!m) return;
    DevSwitch keep;
    for (size_t i = 0; i < m->comms.size(); ++i)
        if (m->comms[i]) m->nccl.CommDestroy(m->comms[i]);
    for (size_t i = 0; i < m->ctx.size(); ++i) {
        if (!m->ctx[i]) continue;                      // never created: nothing of this device to release
        sfmgms_destroy(m->ctx[i]);                     // before the buffers it adopted are freed
        if (cudaSetDevice(m->devices[i]) != cudaSuccess) continue;
        if (m->d_desc[i]) cudaFree(m->d_desc[i]);
        if (m->d_kp[i]) cudaFree(m->d_kp[i]);
        if (m->streams[i]) cudaStreamDestroy(m->streams[i]);
    }
    cudaGetLastError();                                // leave no stale (non-sticky) error behind for the next launch check
    delete m;
}

const char* sfmgms_multi_last_error(const sfmgms_multi* m) { return m ? m->err : g_multi_create_error; }
int sfmgms_multi_device_count(const sfmgms_multi* m) { return m ? m->n : 0; }
sfmgms_ctx* sfmgms_multi_context(sfmgms_multi* m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[(size_t)i] : nullptr; }
double sfmgms_multi_last_broadcast_ms(const sfmgms_multi* m) { return m ? m->last_broadcast_ms : 0.0; }

int sfmgms_multi_set_images(sfmgms_multi* m, int n_images, const int64_t* kp_offsets, const uint8_t* desc, const float* kp_xy,
                            const int32_t* sizes_wh) {
    if (!m) return SFMGMS_ERR_ARG;
    if (n_images < 0 || !kp_offsets || !sizes_wh || kp_offsets[0] != 0) return mfail(m, SFMGMS_ERR_ARG, "bad image-set arguments");
    for (int i = 0; i < n_images; ++i)
        if (kp_offsets[i + 1] < kp_offsets[i]) return mfail(m, SFMGMS_ERR_ARG, "kp_offsets not monotone at image %d", i);
    const int64_t total = kp_offsets[n_images];
    if (total > 0 && (!desc || !kp_xy)) return mfail(m, SFMGMS_ERR_ARG, "null descriptor/keypoint pointer");
    DevSwitch keep;
    try {
        const size_t bd = (size_t)total * 32, bk = (size_t)total * 8;
        for (int i = 0; i < m->n; ++i) {
            MCU(cudaSetDevice(m->devices[(size_t)i]));
            if (bd + 32 > m->cap_desc[(size_t)i]) {
                if (m->d_desc[(size_t)i]) cudaFree(m->d_desc[(size_t)i]);
                m->d_desc[(size_t)i] = nullptr; m->cap_desc[(size_t)i] = 0;
                MCU(cudaMalloc(&m->d_desc[(size_t)i], bd + bd / 8 + 256));
                m->cap_desc[(size_t)i] = bd + bd / 8 + 256;
            }
            if (bk + 8 > m->cap_kp[(size_t)i]) {
                if (m->d_kp[(size_t)i]) cudaFree(m->d_kp[(size_t)i]);
                m->d_kp[(size_t)i] = nullptr; m->cap_kp[(size_t)i] = 0;
                MCU(cudaMalloc(&m->d_kp[(size_t)i], bk + bk / 8 + 256));
                m->cap_kp[(size_t)i] = bk + bk / 8 + 256;
            }
        }
        // host -> GPU 0 once ...
        MCU(cudaSetDevice(m->devices[0]));
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        MCU(cudaEventCreate(&e0)); MCU(cudaEventCreate(&e1));
        if (total) {
            MCU(cudaMemcpyAsync(m->d_desc[0], desc, bd, cudaMemcpyHostToDevice, m->streams[0]));
            MCU(cudaMemcpyAsync(m->d_kp[0], kp_xy, bk, cudaMemcpyHostToDevice, m->streams[0]));
        }
        MCU(cudaEventRecord(e0, m->streams[0]));
        // ... then ONE broadcast per array to every other GPU (single thread: the calls sit in one NCCL group)
        if (m->n > 1 && total) {
            MNCCL(m->nccl.GroupStart());
            for (int i = 0; i < m->n; ++i)
                MNCCL(m->nccl.Broadcast(m->d_desc[0], m->d_desc[(size_t)i], bd, kNcclUint8, 0, m->comms[(size_t)i], m->streams[(size_t)i]));
            MNCCL(m->nccl.GroupEnd());
            MNCCL(m->nccl.GroupStart());
            for (int i = 0; i < m->n; ++i)
                MNCCL(m->nccl.Broadcast(m->d_kp[0], m->d_kp[(size_t)i], bk, kNcclUint8, 0, m->comms[(size_t)i], m->streams[(size_t)i]));
            MNCCL(m->nccl.GroupEnd());
        }
        MCU(cudaEventRecord(e1, m->streams[0]));
        for (int i = 0; i < m->n; ++i) {
            MCU(cudaSetDevice(m->devices[(size_t)i]));
            MCU(cudaStreamSynchronize(m->streams[(size_t)i]));
        }
        float ms = 0.f;
        cudaSetDevice(m->devices[0]);
        cudaEventElapsedTime(&ms, e0, e1);
        m->last_broadcast_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        for (int i = 0; i < m->n; ++i) {
            const int rc = sfmgms_set_images(m->ctx[(size_t)i], n_images, kp_offsets, (const uint8_t*)m->d_desc[(size_t)i],
                                             (const float*)m->d_kp[(size_t)i], sizes_wh, SFMGMS_DEVICE);
            if (rc) return mfail(m, rc, "GPU %d: %s", m->devices[(size_t)i], sfmgms_last_error(m->ctx[(size_t)i]));
        }
        m->n_images = n_images;
        m->offsets.assign(kp_offsets, kp_offsets + n_images + 1);
    } catch (...) {
        return mfail(m, SFMGMS_ERR_ARG, "host allocation failed");
    }
    return SFMGMS_OK;
}

// shared body: kind 0 = full per-match outputs, 1 = compact
static int multi_run(sfmgms_multi* m, int kind, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                     double factor, int32_t* n_inliers, int32_t* best_hyp, int32_t* mask_len, int32_t* train_idx, int32_t* dist,
                     uint8_t* mask, int64_t* inlier_begin, void* matches, float* pts1, float* pts2, int64_t capacity,
                     int64_t* n_total) {
    if (!m) return SFMGMS_ERR_ARG;
    if (n_pairs < 0 || (n_pairs > 0 && !pairs)) return mfail(m, SFMGMS_ERR_ARG, "bad pair list");
    if (n_total) *n_total = 0;
    try {
        std::vector<int> begin;
        cut_shards(m, pairs, n_pairs, begin);
        m->last_shard_begin = begin;
        std::vector<int64_t> row0((size_t)m->n, 0);          // first match row of each shard (full outputs)
        if (kind == 0) {
            int64_t o = 0;
            int g = 0;
            for (int p = 0; p <= n_pairs; ++p) {
                while (g < m->n && begin[(size_t)g] == p) row0[(size_t)g++] = o;
                if (p == n_pairs) break;
                const int a = pairs[2 * p];
                if (a < 0 || a >= m->n_images) return mfail(m, SFMGMS_ERR_ARG, "pair %d references an image outside the set", p);
                o += m->offsets[(size_t)a + 1] - m->offsets[(size_t)a];
            }
        }
        int64_t cursor = 0;
        std::vector<int> rcs((size_t)m->n, 0);
        auto work = [&](int g) {
            try {
                const int p0 = begin[(size_t)g], pn = begin[(size_t)g + 1] - p0;
                if (pn <= 0) return;
                sfmgms_ctx* c = m->ctx[(size_t)g];
                auto off = [&](auto* base, int64_t k) { return base ? base + k : base; };
                if (kind == 0)
                    rcs[(size_t)g] = sfmgms_match_pairs(c, pairs + 2 * (size_t)p0, pn, with_rotation, with_scale, factor, SFMGMS_HOST,
                                                        off(n_inliers, p0), off(best_hyp, p0), off(mask_len, p0),
                                                        off(train_idx, row0[(size_t)g]), off(dist, row0[(size_t)g]), off(mask, row0[(size_t)g]));
                else
                    rcs[(size_t)g] = sfmgms::match_pairs_compact_shared(c, pairs + 2 * (size_t)p0, pn, with_rotation, with_scale, factor,
                                                                        off(n_inliers, p0), off(best_hyp, p0), off(inlier_begin, p0),
                                                                        matches, pts1, pts2, capacity, &cursor);
            } catch (...) {
                rcs[(size_t)g] = SFMGMS_ERR_ARG;
            }
        };
        std::vector<std::thread> th;                          // one host thread per GPU
        for (int g = 1; g < m->n; ++g) th.emplace_back(work, g);
        work(0);
        for (auto& t : th) t.join();
        if (n_total) *n_total = cursor;
        int worst = SFMGMS_OK;
        for (int g = 0; g < m->n; ++g)
            if (rcs[(size_t)g] && (!worst || rcs[(size_t)g] != SFMGMS_ERR_CAPACITY)) {
                worst = rcs[(size_t)g];
                if (worst == SFMGMS_ERR_CAPACITY)
                    mfail(m, worst, "compact output needs %lld rows, capacity is %lld", (long long)cursor, (long long)capacity);
                else
                    mfail(m, worst, "GPU %d (pairs from %d): %s", m->devices[(size_t)g], begin[(size_t)g], sfmgms_last_error(m->ctx[(size_t)g]));
            }
        return worst;
    } catch (...) {
        return mfail(m, SFMGMS_ERR_ARG, "host allocation failed");
    }
}

int sfmgms_multi_match_pairs(sfmgms_multi* m, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                             double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int32_t* mask_len,
                             int32_t* train_idx, int32_t* dist, uint8_t* mask) {
    return multi_run(m, 0, pairs, n_pairs, with_rotation, with_scale, threshold_factor, n_inliers, best_hyp, mask_len, train_idx, dist,
                     mask, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
}

int sfmgms_multi_match_pairs_compact(sfmgms_multi* m, const int32_t* pairs, int n_pairs, int with_rotation, int with_scale,
                                     double threshold_factor, int32_t* n_inliers, int32_t* best_hyp, int64_t* inlier_begin,
                                     void* matches, float* pts1, float* pts2, int64_t capacity, int64_t* n_total) {
    if (m && (capacity < 0 || (capacity > 0 && !matches && !pts1 && !pts2)))
        return mfail(m, SFMGMS_ERR_ARG, "compact output: capacity > 0 needs at least one of matches / pts1 / pts2");
    return multi_run(m, 1, pairs, n_pairs, with_rotation, with_scale, threshold_factor, n_inliers, best_hyp, nullptr, nullptr, nullptr,
                     nullptr, inlier_begin, matches, pts1, pts2, capacity, n_total);
}

}  // extern "C"
