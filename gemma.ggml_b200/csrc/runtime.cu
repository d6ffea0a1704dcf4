// runtime.cu — device runtime of the C-ABI layer: context, buffers, lookup tables, cudaEvent
// profiling and CUDA-graph capture.  Replaces the reference's OpenCL wrapper
// (ref: src/opencl.cpp:19-34 global state, :161-245 setup, :300-348 enqueue/buffers) and its
// wall-clock profiler (ref: src/profiling.cpp:15-134).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <math.h>

#include "common.cuh"

namespace {

struct ProfEntry {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
    cudaEvent_t open = nullptr;
    double total_ms = 0.0;
    int64_t calls = 0;
};

struct Context {
    bool init = false;
    int device = -1;
    cudaDeviceProp prop{};
    cudaStream_t stream = nullptr;
    uint16_t *exp_table = nullptr, *gelu_table = nullptr;
    float *rope_cos = nullptr, *rope_sin = nullptr;
    int rope_n_ctx = 0, rope_n_dims = 0;
    float rope_base = 0.f;
    int64_t launches = 0;
    bool capturing = false;
    int64_t capture_launches = 0;
    bool prof_on = false;
    std::map<std::string, ProfEntry> prof;
};

Context g_ctx;
thread_local std::string g_err;
std::string g_err_global;
std::mutex g_err_mu;

} // namespace

void ggb_set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err_global = buf;
}

bool ggb_is_init() { return g_ctx.init; }
cudaStream_t ggb_cuda_stream() { return g_ctx.stream; }
void ggb_count_launch(int n) {
    if (g_ctx.capturing) g_ctx.capture_launches += n;
    else g_ctx.launches += n;
}
const uint16_t *ggb_exp_table_dev() { return g_ctx.exp_table; }
const uint16_t *ggb_gelu_table_dev() { return g_ctx.gelu_table; }
size_t ggb_tables_bytes() { return 2 * 65536 * 2; } // exp then gelu, contiguous from ggb_exp_table_dev()
const float *ggb_rope_cos_dev() { return g_ctx.rope_cos; }
const float *ggb_rope_sin_dev() { return g_ctx.rope_sin; }
int ggb_rope_n_ctx() { return g_ctx.rope_n_ctx; }
int ggb_rope_n_dims() { return g_ctx.rope_n_dims; }

struct ggb_graph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int n_kernels = 0;
};

// fp16 lookup tables as ggml_init builds them ([UPSTREAM] ggml.c: table[i] = FP32_TO_FP16(f(FP16_TO_FP32(i))));
// computed with the host libm so they are identical to what a CPU ggml on this box would hold.
static inline float host_gelu_f32(float x) {
    const float GELU_COEF_A = 0.044715f;
    const float SQRT_2_OVER_PI = 0.79788456080286535587989211986876f;
    return 0.5f * x * (1.0f + tanhf(SQRT_2_OVER_PI * x * (1.0f + GELU_COEF_A * x * x)));
}

static int build_tables() {
    std::vector<uint16_t> e(65536), g(65536);
    for (uint32_t i = 0; i < 65536; i++) {
        const float f = __half2float(__ushort_as_half((unsigned short)i));
        e[i] = __half_as_ushort(__float2half_rn(expf(f)));
        g[i] = __half_as_ushort(__float2half_rn(host_gelu_f32(f)));
    }
    // one allocation for both tables: the decode kernel pins it in L2 with an access-policy window
    GGB_CUDA(cudaMalloc(&g_ctx.exp_table, 2 * 65536 * 2));
    g_ctx.gelu_table = g_ctx.exp_table + 65536;
    // set-aside L2 for persisting lines (the tables; a lookup that misses under the weight stream costs ~1 us)
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 4u << 20);
    cudaGetLastError();
    GGB_CUDA(cudaMemcpy(g_ctx.exp_table, e.data(), 65536 * 2, cudaMemcpyHostToDevice));
    GGB_CUDA(cudaMemcpy(g_ctx.gelu_table, g.data(), 65536 * 2, cudaMemcpyHostToDevice));
    return GGB_OK;
}

extern "C" {

int ggb_init(int device) {
    if (g_ctx.init) return GGB_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        ggb_set_error("no CUDA device (%s); libggb_cuda has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return GGB_ERR_CUDA;
    }
    GGB_CHECK_ARG(device >= 0 && device < n, "device %d out of range [0,%d)", device, n);
    GGB_CUDA(cudaSetDevice(device));
    GGB_CUDA(cudaGetDeviceProperties(&g_ctx.prop, device));
    if (g_ctx.prop.major != 10) {
        ggb_set_error("device %s is sm_%d%d; this library is built for sm_100a only", g_ctx.prop.name,
                      g_ctx.prop.major, g_ctx.prop.minor);
        return GGB_ERR_UNSUPPORTED;
    }
    GGB_CUDA(cudaStreamCreateWithFlags(&g_ctx.stream, cudaStreamNonBlocking));
    g_ctx.device = device;
    int rc = build_tables();
    if (rc) return rc;
    g_ctx.init = true;
    return GGB_OK;
}

int ggb_shutdown(void) {
    if (!g_ctx.init) return GGB_OK;
    cudaStreamSynchronize(g_ctx.stream);
    ggb_comm_destroy();
    cudaFree(g_ctx.exp_table);
    cudaFree(g_ctx.rope_cos);
    cudaFree(g_ctx.rope_sin);
    for (auto &kv : g_ctx.prof) {
        for (auto &p : kv.second.pending) {
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
    }
    g_ctx.prof.clear();
    cudaStreamDestroy(g_ctx.stream);
    g_ctx = Context();
    return GGB_OK;
}

const char *ggb_last_error(void) {
    if (!g_err.empty()) return g_err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_err = g_err_global;
    return g_err.c_str();
}

int ggb_device_info(char *buf, size_t buflen) {
    GGB_REQUIRE_INIT();
    snprintf(buf, buflen, "%s sm_%d%d SMs=%d mem=%.1fGiB L2=%.0fMiB smem/block=%zuKiB", g_ctx.prop.name,
             g_ctx.prop.major, g_ctx.prop.minor, g_ctx.prop.multiProcessorCount,
             (double)g_ctx.prop.totalGlobalMem / (1 << 30), (double)g_ctx.prop.l2CacheSize / (1 << 20),
             g_ctx.prop.sharedMemPerBlockOptin >> 10);
    return GGB_OK;
}

int ggb_sm_count(void) { return g_ctx.init ? g_ctx.prop.multiProcessorCount : 0; }

int ggb_sync(void) {
    GGB_REQUIRE_INIT();
    GGB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return GGB_OK;
}

void *ggb_stream(void) { return (void *)g_ctx.stream; }
int64_t ggb_launch_count(void) { return g_ctx.launches; }

int ggb_buffer_alloc(size_t bytes, void **dptr) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(dptr != nullptr, "dptr is null");
    GGB_CUDA(cudaMalloc(dptr, bytes ? bytes : 16));
    return GGB_OK;
}
int ggb_buffer_free(void *dptr) {
    GGB_REQUIRE_INIT();
    GGB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    GGB_CUDA(cudaFree(dptr));
    return GGB_OK;
}
int ggb_buffer_memset(void *dptr, int value, size_t bytes) {
    GGB_REQUIRE_INIT();
    GGB_CUDA(cudaMemsetAsync(dptr, value, bytes, g_ctx.stream));
    return GGB_OK;
}
int ggb_upload(void *dst_dev, const void *src_host, size_t bytes) {
    GGB_REQUIRE_INIT();
    GGB_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
    return GGB_OK;
}
int ggb_download(void *dst_host, const void *src_dev, size_t bytes) {
    GGB_REQUIRE_INIT();
    GGB_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
    GGB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    return GGB_OK;
}
int ggb_host_alloc(size_t bytes, void **hptr) {
    GGB_REQUIRE_INIT();
    GGB_CUDA(cudaMallocHost(hptr, bytes ? bytes : 16));
    return GGB_OK;
}
int ggb_host_free(void *hptr) {
    GGB_REQUIRE_INIT();
    GGB_CUDA(cudaFreeHost(hptr));
    return GGB_OK;
}

// ---- rope tables: [UPSTREAM] ggml rope NEOX recurrence (theta *= theta_scale), host libm
int ggb_rope_table_init(int n_ctx, int n_dims, float freq_base) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(n_ctx > 0 && n_dims > 0 && n_dims % 2 == 0, "bad rope table shape %d x %d", n_ctx, n_dims);
    if (g_ctx.rope_cos && g_ctx.rope_n_ctx >= n_ctx && g_ctx.rope_n_dims == n_dims && g_ctx.rope_base == freq_base)
        return GGB_OK;
    GGB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    cudaFree(g_ctx.rope_cos);
    cudaFree(g_ctx.rope_sin);
    g_ctx.rope_cos = g_ctx.rope_sin = nullptr;
    const int half = n_dims / 2;
    std::vector<float> c((size_t)n_ctx * half), s((size_t)n_ctx * half);
    const float theta_scale = powf(freq_base, -2.0f / n_dims);
    for (int p = 0; p < n_ctx; p++) {
        float theta = (float)p;
        for (int i = 0; i < half; i++) {
            c[(size_t)p * half + i] = cosf(theta);
            s[(size_t)p * half + i] = sinf(theta);
            theta *= theta_scale;
        }
    }
    GGB_CUDA(cudaMalloc(&g_ctx.rope_cos, c.size() * 4));
    GGB_CUDA(cudaMalloc(&g_ctx.rope_sin, s.size() * 4));
    GGB_CUDA(cudaMemcpy(g_ctx.rope_cos, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
    GGB_CUDA(cudaMemcpy(g_ctx.rope_sin, s.data(), s.size() * 4, cudaMemcpyHostToDevice));
    g_ctx.rope_n_ctx = n_ctx;
    g_ctx.rope_n_dims = n_dims;
    g_ctx.rope_base = freq_base;
    return GGB_OK;
}

// ---- profiling: same name-keyed accumulate/count interface as src/profiling.h, on cudaEvents
int ggb_prof_enable(int on) {
    g_ctx.prof_on = on != 0;
    return GGB_OK;
}
int ggb_prof_start(const char *name) {
    if (!g_ctx.prof_on || g_ctx.capturing) return GGB_OK;
    GGB_REQUIRE_INIT();
    ProfEntry &pe = g_ctx.prof[name];
    cudaEvent_t ev;
    GGB_CUDA(cudaEventCreate(&ev));
    GGB_CUDA(cudaEventRecord(ev, g_ctx.stream));
    pe.open = ev;
    return GGB_OK;
}
int ggb_prof_stop(const char *name) {
    if (!g_ctx.prof_on || g_ctx.capturing) return GGB_OK;
    GGB_REQUIRE_INIT();
    auto it = g_ctx.prof.find(name);
    GGB_CHECK_ARG(it != g_ctx.prof.end() && it->second.open, "prof_stop('%s') without start", name);
    cudaEvent_t ev;
    GGB_CUDA(cudaEventCreate(&ev));
    GGB_CUDA(cudaEventRecord(ev, g_ctx.stream));
    it->second.pending.emplace_back(it->second.open, ev);
    it->second.open = nullptr;
    return GGB_OK;
}
static int prof_resolve() {
    GGB_CUDA(cudaStreamSynchronize(g_ctx.stream));
    for (auto &kv : g_ctx.prof) {
        for (auto &p : kv.second.pending) {
            float ms = 0.f;
            GGB_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
            kv.second.total_ms += ms;
            kv.second.calls++;
            cudaEventDestroy(p.first);
            cudaEventDestroy(p.second);
        }
        kv.second.pending.clear();
    }
    return GGB_OK;
}
int ggb_prof_reset(void) {
    GGB_REQUIRE_INIT();
    int rc = prof_resolve();
    g_ctx.prof.clear();
    return rc;
}
int ggb_prof_get(const char *name, double *total_ms, int64_t *calls) {
    GGB_REQUIRE_INIT();
    int rc = prof_resolve();
    if (rc) return rc;
    auto it = g_ctx.prof.find(name);
    if (total_ms) *total_ms = it == g_ctx.prof.end() ? 0.0 : it->second.total_ms;
    if (calls) *calls = it == g_ctx.prof.end() ? 0 : it->second.calls;
    return GGB_OK;
}
int ggb_prof_report(char *buf, size_t buflen) {
    GGB_REQUIRE_INIT();
    int rc = prof_resolve();
    if (rc) return rc;
    double total = 0;
    for (auto &kv : g_ctx.prof) total += kv.second.total_ms;
    size_t off = 0;
    for (auto &kv : g_ctx.prof) {
        int w = snprintf(buf + off, off < buflen ? buflen - off : 0, "%s: %.3f ms, %.4f%%, calls %lld\n",
                         kv.first.c_str(), kv.second.total_ms, total > 0 ? 100.0 * kv.second.total_ms / total : 0.0,
                         (long long)kv.second.calls);
        if (w < 0 || off + (size_t)w >= buflen) break;
        off += (size_t)w;
    }
    return GGB_OK;
}

// ---- CUDA graph capture
int ggb_capture_begin(void) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(!g_ctx.capturing, "capture already in progress");
    GGB_CUDA(cudaStreamBeginCapture(g_ctx.stream, cudaStreamCaptureModeThreadLocal));
    g_ctx.capturing = true;
    g_ctx.capture_launches = 0;
    return GGB_OK;
}
int ggb_capture_end(ggb_graph **out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(g_ctx.capturing, "no capture in progress");
    g_ctx.capturing = false;
    ggb_graph *g = new ggb_graph();
    cudaError_t e = cudaStreamEndCapture(g_ctx.stream, &g->graph);
    if (e != cudaSuccess) {
        ggb_set_error("cudaStreamEndCapture -> %s", cudaGetErrorString(e));
        delete g;
        return GGB_ERR_CUDA;
    }
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) {
        ggb_set_error("cudaGraphInstantiate -> %s", cudaGetErrorString(e));
        cudaGraphDestroy(g->graph);
        delete g;
        return GGB_ERR_CUDA;
    }
    g->n_kernels = (int)g_ctx.capture_launches;
    *out = g;
    return GGB_OK;
}
int ggb_graph_launch(ggb_graph *g) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(g && g->exec, "null graph");
    GGB_CUDA(cudaGraphLaunch(g->exec, g_ctx.stream));
    g_ctx.launches += g->n_kernels;
    return GGB_OK;
}
int ggb_graph_free(ggb_graph *g) {
    if (!g) return GGB_OK;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return GGB_OK;
}
int ggb_graph_kernel_count(const ggb_graph *g) { return g ? g->n_kernels : 0; }

} // extern "C"
