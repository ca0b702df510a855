// common.cuh -- context, error handling and host/device buffer staging shared by the
// C-ABI translation units of libpycd_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <stdexcept>

#include "../../include/pycd_b200.h"

namespace pycd {

void set_error(const char *fmt, ...);

struct Error : std::runtime_error {
    explicit Error(const std::string &m) : std::runtime_error(m) {}
};

#define PYCD_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess)                                                             \
            throw ::pycd::Error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + \
                                " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")");    \
    } while (0)

#define PYCD_REQUIRE(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) throw ::pycd::Error(std::string("invalid argument: ") + (msg)); \
    } while (0)

enum KernelClass { KC_EWALD_FOURIER = 0, KC_EWALD_FINISH = 1, KC_EWALD_EXPAND = 2,
                   KC_KMC_STEP = 3, KC_MSD = 4, KC_VLAT = 5, KC_COUNT = 6 };

}  // namespace pycd

struct pycd_ctx {
    int device = 0;
    int n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[pycd::KC_COUNT][2] = {};
    int64_t launches = 0;
    double last_ms[pycd::KC_COUNT] = {0, 0, 0, 0, 0, 0};
    double total_ms[pycd::KC_COUNT] = {0, 0, 0, 0, 0, 0};
    int64_t class_launches[pycd::KC_COUNT] = {0, 0, 0, 0, 0, 0};
    void *flush_buf = nullptr;
    int flush_value = 0;
    cudaStream_t copy_stream = nullptr;   // pipelined read-back (pycd_kmc_read_begin / _end)
    void *arena = nullptr;                // work buffers of the synchronous entry points (grow-only, see Arena)
    size_t arena_bytes = 0;
    // kernel timers of launches that were issued without a host synchronisation (pycd_kmc_advance_async):
    // their event pairs wait here until collect_timers() adds them to the totals
    struct Pending { cudaEvent_t a, b; int cls, n_launch; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> ev_pool;
};

namespace pycd {

// RAII: make the context's device current.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(const pycd_ctx *ctx) {
        cudaGetDevice(&prev);
        if (prev != ctx->device) PYCD_CUDA(cudaSetDevice(ctx->device));
    }
    ~DeviceGuard() {
        int cur = -1;
        cudaGetDevice(&cur);
        if (prev >= 0 && cur != prev) cudaSetDevice(prev);
    }
};

bool is_device_pointer(const void *p);

// Owning device buffer.
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        release();
        if (count == 0) return;
        PYCD_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
        n = count;
    }
    void zero(cudaStream_t s) {
        if (p) PYCD_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
};

// Work buffers of one synchronous entry point, carved out of the context's grow-only arena: no cudaMalloc /
// cudaFree pair per call (each costs 0.1-0.5 ms, more than the kernels of a small job).  One user at a time:
// the entry points that take it synchronise the stream before they return.
struct Arena {
    pycd_ctx *ctx;
    size_t need = 0, used = 0;
    explicit Arena(pycd_ctx *c) : ctx(c) {}
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    void reserve(size_t bytes) { need += pad(bytes); }
    void commit() {
        if (need > ctx->arena_bytes) {
            if (ctx->arena) cudaFree(ctx->arena);
            ctx->arena = nullptr;
            ctx->arena_bytes = 0;
            PYCD_CUDA(cudaMalloc(&ctx->arena, need));
            ctx->arena_bytes = need;
        }
    }
    template <typename T>
    T *take(size_t count) {
        T *p = reinterpret_cast<T *>(static_cast<char *>(ctx->arena) + used);
        used += pad(count * sizeof(T));
        if (used > ctx->arena_bytes) throw Error("arena overrun");
        return p;
    }
};

// Read-only input that may live on the host or on the device: device pointers are
// used in place, host pointers are copied into an owned buffer.
template <typename T>
struct InBuf {
    const T *p = nullptr;
    DevBuf<T> own;
    void bind(const T *src, size_t count, cudaStream_t s) {
        if (src == nullptr || count == 0) { p = nullptr; return; }
        if (is_device_pointer(src)) { p = src; return; }
        own.alloc(count);
        PYCD_CUDA(cudaMemcpyAsync(own.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
        p = own.p;
    }
};

// Output that may live on the host or on the device: kernels write dev(); finish()
// copies back to the host if needed.
template <typename T>
struct OutBuf {
    T *user = nullptr;
    T *d = nullptr;
    size_t n = 0;
    DevBuf<T> own;
    void bind(T *dst, size_t count) {
        user = dst; n = count;
        if (!dst || count == 0) { d = nullptr; return; }
        if (is_device_pointer(dst)) { d = dst; return; }
        own.alloc(count);
        d = own.p;
    }
    T *dev() const { return d; }
    void finish(cudaStream_t s) {
        if (user && d && d != user)
            PYCD_CUDA(cudaMemcpyAsync(user, d, n * sizeof(T), cudaMemcpyDeviceToHost, s));
    }
};

inline cudaEvent_t pooled_event(pycd_ctx *ctx) {
    cudaEvent_t e = nullptr;
    if (!ctx->ev_pool.empty()) {
        e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
    } else {
        PYCD_CUDA(cudaEventCreate(&e));
    }
    return e;
}

// adds the deferred timers to the totals (blocks until their launches have finished)
inline void collect_timers(pycd_ctx *ctx) {
    for (const pycd_ctx::Pending &p : ctx->pending) {
        float ms = 0.f;
        cudaEventSynchronize(p.b);
        cudaEventElapsedTime(&ms, p.a, p.b);
        ctx->last_ms[p.cls] = ms;
        ctx->total_ms[p.cls] += ms;
        ctx->class_launches[p.cls] += p.n_launch;
        ctx->ev_pool.push_back(p.a);
        ctx->ev_pool.push_back(p.b);
    }
    ctx->pending.clear();
}

struct KernelTimer {
    pycd_ctx *ctx;
    int cls;
    int n_launch;
    cudaEvent_t a, b;
    // deferred = true: own event pair from the context's pool, so that several launches can be in flight
    KernelTimer(pycd_ctx *c, int k, bool deferred = false) : ctx(c), cls(k), n_launch(0) {
        a = deferred ? pooled_event(c) : c->ev[k][0];
        b = deferred ? pooled_event(c) : c->ev[k][1];
        cudaEventRecord(a, ctx->stream);
    }
    // stop() right after the launches (asynchronous); read() blocks on the stop event
    void stop(int launches_timed = 1) {
        n_launch = launches_timed;
        cudaEventRecord(b, ctx->stream);
    }
    void read() {
        float ms = 0.f;
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        ctx->last_ms[cls] = ms;
        ctx->total_ms[cls] += ms;
        ctx->class_launches[cls] += n_launch;
    }
    // instead of read(): leave the pair with the context (collect_timers)
    void defer() { ctx->pending.push_back({a, b, cls, n_launch}); }
};

// NVTX range around a library call (visible to nsys / ncu --nvtx; a no-op without a tool attached)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

inline void check_launch(pycd_ctx *ctx, const char *name) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw Error(std::string("launch of ") + name + " failed: " + cudaGetErrorString(e));
    ctx->launches += 1;
}

// wraps an ABI body: exceptions -> error code + message
template <typename F>
int guarded(F &&f) {
    try {
        f();
        return 0;
    } catch (const std::exception &e) {
        set_error("%s", e.what());
        return 1;
    } catch (...) {
        set_error("unknown error");
        return 2;
    }
}

}  // namespace pycd
