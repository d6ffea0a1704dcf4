// common.cuh — shared device helpers and the library context for libggb_cuda.so (sm_100a).
// Compiled with -fmad=false: every fp32 expression rounds once per operator unless it
// spells fmaf()/__fmaf_rn() explicitly, so the kernels' arithmetic matches the oracle's.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/ggml_b200.h"

#define GGB_QK 32

// ---- error plumbing -------------------------------------------------------
void ggb_set_error(const char *fmt, ...);
bool ggb_is_init();
cudaStream_t ggb_cuda_stream();
void ggb_count_launch(int n = 1);

#define GGB_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t err__ = (call);                                                             \
        if (err__ != cudaSuccess) {                                                             \
            ggb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
            return GGB_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define GGB_REQUIRE_INIT()                                  \
    do {                                                    \
        if (!ggb_is_init()) {                               \
            ggb_set_error("ggb_init has not succeeded");    \
            return GGB_ERR_NOT_INIT;                        \
        }                                                   \
    } while (0)

#define GGB_CHECK_ARG(cond, ...)        \
    do {                                \
        if (!(cond)) {                  \
            ggb_set_error(__VA_ARGS__); \
            return GGB_ERR_ARG;         \
        }                               \
    } while (0)

#define GGB_LAUNCH_CHECK()                                                                       \
    do {                                                                                         \
        cudaError_t err__ = cudaGetLastError();                                                  \
        if (err__ != cudaSuccess) {                                                              \
            ggb_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(err__)); \
            return GGB_ERR_CUDA;                                                                 \
        }                                                                                        \
        ggb_count_launch();                                                                      \
    } while (0)

// ---- device weight layout -------------------------------------------------
// SoA repack of ggml AoS blocks: payload plane [rows][nb][16|32 B], scale plane [rows][nb] fp16.
struct ggb_weight {
    int type;
    int64_t rows, cols, nb; // nb = cols / 32
    uint8_t *qs;            // payload plane
    __half *d;              // scale plane
    void *base;             // owning allocation (null for slices)
    size_t bytes;
};

static inline int ggb_payload_bytes(int type) { return type == GGB_TYPE_Q4_0 ? 16 : 32; }
static inline int ggb_block_bytes(int type) { return type == GGB_TYPE_Q4_0 ? 18 : 34; }

// ---- device helpers -------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
__device__ __forceinline__ float warp_max_f32(float v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// 128-bit streaming load (weights are read once per token: do not allocate in L1)
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float half_bits_to_float(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t float_to_half_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }

// quantize_row_q8_0 for one block spread over a warp: lane i holds element i.
// returns the int8 value; d_out (all lanes) is the fp32 scale before fp16 rounding.
__device__ __forceinline__ int quantize_q8_lane(float x, float &d_out) {
    const float amax = warp_max_f32(fabsf(x));
    const float d = __fdiv_rn(amax, 127.0f);
    const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
    d_out = d;
    return (int)roundf(__fmul_rn(x, id));
}

// counter-based synthetic weights; identical to oracle/ggml_oracle.c orc_synth_value
__host__ __device__ __forceinline__ uint64_t ggb_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t ggb_synth_key(uint64_t seed, uint64_t tensor_id) {
    return ggb_splitmix64(seed + tensor_id * 0xD1B54A32D192ED03ull);
}
__device__ __forceinline__ float ggb_synth_value(uint64_t key, uint64_t index, float unit) {
    const uint64_t r = ggb_splitmix64(key + index);
    const int sum = (int)(r & 0xffff) + (int)((r >> 16) & 0xffff) + (int)((r >> 32) & 0xffff) +
                    (int)((r >> 48) & 0xffff) - 131070;
    return __fmul_rn((float)sum, unit);
}

#endif // __CUDACC__

// host: sigma * (1/37837.2274f) with the same two roundings the oracle performs
static inline float ggb_synth_unit(float sigma) {
    volatile float inv = 1.0f / 37837.2274f;
    volatile float u = sigma * inv;
    return u;
}

// tables shared by several translation units (runtime.cu owns them)
const uint16_t *ggb_exp_table_dev();  // 64 Ki fp16: exp(f16)
const uint16_t *ggb_gelu_table_dev(); // 64 Ki fp16: gelu(f16)
size_t ggb_tables_bytes();            // both tables are one allocation starting at ggb_exp_table_dev()
const float *ggb_rope_cos_dev();      // [n_ctx][n_dims/2]
const float *ggb_rope_sin_dev();
int ggb_rope_n_ctx();
int ggb_rope_n_dims();
