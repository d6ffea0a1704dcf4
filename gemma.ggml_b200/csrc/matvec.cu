// matvec.cu — K2/K3: Q4_0 / Q8_0 weight rows · Q8_0 activations (decode mat-vec), the hot loop
// of the reference (ref: src/hpc.cpp:15-41 mul_mat_sub -> vec_dot, 99.98 % of mul_mat time,
// notes.md:87).  Not a port of src/kernals.cl (one work-item per output, serial K loop):
//
//   * a warp owns a row; lane l computes the per-block term of blocks l, l + 32, ... with dp4a
//     (one coalesced 128-bit payload load per lane per block, nibbles unpacked in registers, the
//     "-8" of Q4_0 folded into -8 * sum of the activation block);
//   * the per-block terms are the scalar reference's ((float)sumi*dx)*dy and are ADDED IN THE
//     REFERENCE'S ORDER: one fp32 accumulator, blocks 0, 1, 2, ... (the warp passes the 32 terms of a
//     round through a shuffle chain), so every output is bit-identical to ggml's scalar vec_dot.
//     This per-operator form is the fallback / T > 1 lowering; the persistent decode kernel
//     (decoder.cu) produces the same bits with a dedicated chain warp.
#include "dot.cuh"

namespace {

constexpr int kMaxMats = 3;

struct MatvecMat {
    const uint8_t *qs;
    const __half *d;
    float *dst;
    const float *resid;
    int64_t rows;
};

struct MatvecArgs {
    MatvecMat m[kMaxMats];
    int n_mats;
    int64_t total_rows;
    int64_t nb;
    const uint8_t *act; // planar q8 rows
    int64_t act_stride; // bytes between activation rows (T > 1)
    int64_t dst_stride; // floats between dst rows
    int64_t cols;
};

constexpr int kMvWarps = 8;

template <int TYPE>
__global__ void __launch_bounds__(kMvWarps * 32) matvec_kernel(const MatvecArgs a) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint8_t *act = a.act + (int64_t)blockIdx.y * a.act_stride;
    const int64_t gr = (int64_t)blockIdx.x * kMvWarps + warp;
    if (gr >= a.total_rows) return;
    int mi = 0;
    int64_t r = gr;
    while (mi + 1 < a.n_mats && r >= a.m[mi].rows) {
        r -= a.m[mi].rows;
        mi++;
    }
    float v = row_dot_ref<TYPE>(a.m[mi].qs, a.m[mi].d, r, a.nb, act, a.cols, lane);
    if (lane == 0) {
        float *dst = a.m[mi].dst + (int64_t)blockIdx.y * a.dst_stride;
        if (a.m[mi].resid) v = __fadd_rn(v, a.m[mi].resid[(int64_t)blockIdx.y * a.dst_stride + r]);
        dst[r] = v;
    }
}

// FFN front half: one CTA per 32 output rows: gate and up dots, GeGLU through the fp16 gelu
// table ([UPSTREAM] ggml_vec_gelu_f32), then quantize_row_q8_0 of the 32 products — the
// activation of the down projection never exists in fp32 in HBM (unless act_f32 asks for a dump).
struct GateUpArgs {
    const uint8_t *gate_qs, *up_qs;
    const __half *gate_d, *up_d;
    int64_t nb, cols, rows;
    const uint8_t *act;
    const uint16_t *gelu_table;
    int8_t *out_qs;
    __half *out_d;
    float *act_f32;
};

template <int TYPE>
__global__ void __launch_bounds__(256) gate_up_kernel(const GateUpArgs a) {
    __shared__ float s_g[32], s_u[32];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t tile_row0 = (int64_t)blockIdx.x * 32;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        const int64_t r = tile_row0 + warp * 4 + i;
        const float gs = row_dot_ref<TYPE>(a.gate_qs, a.gate_d, r, a.nb, a.act, a.cols, lane);
        const float us = row_dot_ref<TYPE>(a.up_qs, a.up_d, r, a.nb, a.act, a.cols, lane);
        if (lane == 0) {
            s_g[warp * 4 + i] = gs;
            s_u[warp * 4 + i] = us;
        }
    }
    __syncthreads();
    if (warp == 0) {
        const float gv = s_g[lane];
        const float gl = half_bits_to_float(a.gelu_table[float_to_half_bits(gv)]);
        const float x = __fmul_rn(gl, s_u[lane]);
        if (a.act_f32) a.act_f32[tile_row0 + lane] = x;
        float d;
        const int q = quantize_q8_lane(x, d);
        a.out_qs[tile_row0 + lane] = (int8_t)q;
        if (lane == 0) a.out_d[blockIdx.x] = __float2half_rn(d);
    }
}

int launch_matvec(int n, const ggb_weight *const *w, const void *q8, int64_t T, float *const *dst,
                  int64_t dst_stride, const float *const *resid) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(n >= 1 && n <= kMaxMats, "n_mats %d out of range", n);
    GGB_CHECK_ARG(q8 != nullptr && T >= 1, "null activations");
    MatvecArgs a{};
    a.n_mats = n;
    a.total_rows = 0;
    for (int i = 0; i < n; i++) {
        GGB_CHECK_ARG(w[i] && dst[i], "null weight/dst %d", i);
        GGB_CHECK_ARG(w[i]->type == w[0]->type && w[i]->cols == w[0]->cols, "mixed types/cols in one launch");
        a.m[i].qs = w[i]->qs;
        a.m[i].d = w[i]->d;
        a.m[i].dst = dst[i];
        a.m[i].resid = resid ? resid[i] : nullptr;
        a.m[i].rows = w[i]->rows;
        a.total_rows += w[i]->rows;
    }
    a.nb = w[0]->nb;
    a.cols = w[0]->cols;
    a.act = (const uint8_t *)q8;
    a.act_stride = (int64_t)ggb_q8_row_bytes(a.cols);
    a.dst_stride = dst_stride;
    const dim3 grid((unsigned)((a.total_rows + kMvWarps - 1) / kMvWarps), (unsigned)T);
    const int warps = kMvWarps;
    const size_t smem = 0;
    if (w[0]->type == GGB_TYPE_Q4_0)
        matvec_kernel<GGB_TYPE_Q4_0><<<grid, warps * 32, smem, ggb_cuda_stream()>>>(a);
    else
        matvec_kernel<GGB_TYPE_Q8_0><<<grid, warps * 32, smem, ggb_cuda_stream()>>>(a);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

} // namespace

extern "C" {

size_t ggb_q8_row_bytes(int64_t cols) { return (size_t)((cols + cols / 16 + 15) & ~(int64_t)15); }

int ggb_op_mul_mat_q8(const ggb_weight *w, const void *q8, int64_t T, float *dst, int64_t dst_stride,
                      const float *residual) {
    const ggb_weight *ws[1] = {w};
    float *ds[1] = {dst};
    const float *rs[1] = {residual};
    return launch_matvec(1, ws, q8, T, ds, dst_stride, rs);
}

int ggb_op_mul_mat_q8_multi(int n, const ggb_weight *const *w, const void *q8, float *const *dst) {
    return launch_matvec(n, w, q8, 1, dst, 0, nullptr);
}

int ggb_op_ffn_gate_up(const ggb_weight *gate, const ggb_weight *up, const void *q8, void *q8_out, float *act_f32) {
    GGB_CHECK_ARG(gate && q8_out, "null pointer");
    return ggb_op_ffn_gate_up_shard(gate, up, q8, q8_out, (uint8_t *)q8_out + gate->rows, act_f32);
}

int ggb_op_ffn_gate_up_shard(const ggb_weight *gate, const ggb_weight *up, const void *q8, void *out_qs, void *out_d,
                             float *act_f32) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(gate && up && q8 && out_qs && out_d, "null pointer");
    GGB_CHECK_ARG(gate->type == up->type && gate->rows == up->rows && gate->cols == up->cols, "gate/up mismatch");
    GGB_CHECK_ARG(gate->rows % 32 == 0, "n_ff %lld not a multiple of 32", (long long)gate->rows);
    GateUpArgs a{};
    a.gate_qs = gate->qs; a.up_qs = up->qs; a.gate_d = gate->d; a.up_d = up->d;
    a.nb = gate->nb; a.cols = gate->cols; a.rows = gate->rows;
    a.act = (const uint8_t *)q8;
    a.gelu_table = ggb_gelu_table_dev();
    a.out_qs = (int8_t *)out_qs;
    a.out_d = (__half *)out_d;
    a.act_f32 = act_f32;
    const unsigned grid = (unsigned)(gate->rows / 32);
    if (gate->type == GGB_TYPE_Q4_0) gate_up_kernel<GGB_TYPE_Q4_0><<<grid, 256, 0, ggb_cuda_stream()>>>(a);
    else gate_up_kernel<GGB_TYPE_Q8_0><<<grid, 256, 0, ggb_cuda_stream()>>>(a);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

} // extern "C"
