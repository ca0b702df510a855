// ctx.cu -- context lifetime, error reporting, pointer classification.
#include "common.cuh"

namespace pycd {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

bool is_device_pointer(const void *p) {
    if (!p) return false;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear: plain host memory on old drivers
        return false;
    }
    return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

}  // namespace pycd

using namespace pycd;

extern "C" {

int pycd_abi_version(void) { return PYCD_ABI_VERSION; }

const char *pycd_last_error(void) { return g_last_error.c_str(); }

int pycd_ctx_create(int device, pycd_ctx **out) {
    return guarded([&] {
        PYCD_REQUIRE(out != nullptr, "out is NULL");
        *out = nullptr;
        int n_dev = 0;
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e != cudaSuccess || n_dev == 0)
            throw Error(std::string("no CUDA device visible (") + cudaGetErrorString(e) +
                        "); libpycd_b200 has no CPU fallback");
        PYCD_REQUIRE(device >= 0 && device < n_dev, "device ordinal out of range");
        cudaDeviceProp prop;
        PYCD_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            throw Error(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                        std::to_string(prop.minor) + "; this library is built for sm_100a (B200) only");
        pycd_ctx *ctx = new pycd_ctx();
        ctx->device = device;
        ctx->n_sm = prop.multiProcessorCount;
        PYCD_CUDA(cudaSetDevice(device));
        PYCD_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        PYCD_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < KC_COUNT; ++k)
            for (int j = 0; j < 2; ++j) PYCD_CUDA(cudaEventCreate(&ctx->ev[k][j]));
        *out = ctx;
    });
}

int pycd_ctx_destroy(pycd_ctx *ctx) {
    return guarded([&] {
        if (!ctx) return;
        DeviceGuard g(ctx);
        cudaStreamSynchronize(ctx->stream);
        for (int k = 0; k < KC_COUNT; ++k)
            for (int j = 0; j < 2; ++j)
                if (ctx->ev[k][j]) cudaEventDestroy(ctx->ev[k][j]);
        collect_timers(ctx);
        for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
        if (ctx->flush_buf) cudaFree(ctx->flush_buf);
        if (ctx->arena) cudaFree(ctx->arena);
        cudaStreamSynchronize(ctx->copy_stream);
        cudaStreamDestroy(ctx->copy_stream);
        cudaStreamDestroy(ctx->stream);
        delete ctx;
    });
}

int pycd_ctx_info(pycd_ctx *ctx, int32_t *n_sm, int64_t *free_bytes, int64_t *total_bytes) {
    return guarded([&] {
        PYCD_REQUIRE(ctx != nullptr, "ctx is NULL");
        DeviceGuard g(ctx);
        size_t f = 0, t = 0;
        PYCD_CUDA(cudaMemGetInfo(&f, &t));
        if (n_sm) *n_sm = ctx->n_sm;
        if (free_bytes) *free_bytes = (int64_t)f;
        if (total_bytes) *total_bytes = (int64_t)t;
    });
}

int64_t pycd_ctx_launch_count(pycd_ctx *ctx) { return ctx ? ctx->launches : -1; }

// launches issued by pycd_kmc_advance_async are accounted when their events are collected
static void settle(pycd_ctx *ctx) {
    if (ctx && !ctx->pending.empty()) {
        DeviceGuard g(ctx);
        collect_timers(ctx);
    }
}

double pycd_ctx_last_kernel_ms(pycd_ctx *ctx, int32_t k) {
    settle(ctx);
    return (ctx && k >= 0 && k < KC_COUNT) ? ctx->last_ms[k] : -1.0;
}

double pycd_ctx_total_kernel_ms(pycd_ctx *ctx, int32_t k) {
    settle(ctx);
    return (ctx && k >= 0 && k < KC_COUNT) ? ctx->total_ms[k] : -1.0;
}

int64_t pycd_ctx_class_launches(pycd_ctx *ctx, int32_t k) {
    settle(ctx);
    return (ctx && k >= 0 && k < KC_COUNT) ? ctx->class_launches[k] : -1;
}

int pycd_ctx_flush_l2(pycd_ctx *ctx) {
    return guarded([&] {
        PYCD_REQUIRE(ctx != nullptr, "ctx is NULL");
        DeviceGuard g(ctx);
        const size_t bytes = (size_t)512 << 20;  // 4x the 126 MB L2
        if (!ctx->flush_buf) PYCD_CUDA(cudaMalloc(&ctx->flush_buf, bytes));
        ctx->flush_value ^= 0xff;
        PYCD_CUDA(cudaMemsetAsync(ctx->flush_buf, ctx->flush_value, bytes, ctx->stream));   // stream-ordered
    });
}

int pycd_host_alloc(int64_t bytes, void **out) {
    return guarded([&] {
        PYCD_REQUIRE(out && bytes > 0, "bad pinned allocation request");
        PYCD_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    });
}

int pycd_host_free(void *ptr) {
    return guarded([&] {
        if (ptr) PYCD_CUDA(cudaFreeHost(ptr));
    });
}

int pycd_nvtx_push(const char *name) {
    nvtxRangePushA(name ? name : "pycd");
    return 0;
}

int pycd_nvtx_pop(void) {
    nvtxRangePop();
    return 0;
}

int pycd_ctx_reset_timers(pycd_ctx *ctx) {
    if (!ctx) return 1;
    settle(ctx);
    for (int k = 0; k < KC_COUNT; ++k) {
        ctx->total_ms[k] = 0;
        ctx->class_launches[k] = 0;
    }
    return 0;
}

}  // extern "C"
