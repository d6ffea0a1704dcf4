// matvec.cu — K2/K3: Q4_0 / Q8_0 weight rows · Q8_0 activations (decode mat-vec), the hot loop
// of the reference (ref: src/hpc.cpp:15-41 mul_mat_sub -> vec_dot, 99.98 % of mul_mat time,
// notes.md:87).  Not a port of src/kernals.cl (one work-item per output, serial K loop):
//
//   * a warp owns a row (or, for K > 2048, one 64-block segment of a row); lane l owns blocks
//     l, l+32 of the segment, so the activation blocks a lane needs never change
//     from row to row and live in REGISTERS for the whole kernel;
//   * weight payloads stream with one coalesced 128-bit load per lane per block (SoA planes,
//     L1 no-allocate), nibbles are unpacked in registers and dotted with dp4a; the "-8" of
//     Q4_0 is folded into a per-lane constant (-8 * sum of the activation block);
//   * per-block terms are the scalar reference's ((float)sumi*dx)*dy, summed in the DEV order
//     documented in include/ggml_b200.h; reduction is warp shuffles (+ smem across segments).
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
    int ksegs;         // ceil(nb / 128)
    int rows_per_iter; // warps per CTA / ksegs
    int64_t rows_per_cta;
};

template <int TYPE>
__global__ void __launch_bounds__(384) matvec_kernel(const MatvecArgs a) {
    extern __shared__ float s_part[]; // [rows_per_iter][ksegs]
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int kseg = warp % a.ksegs;
    const int rsub = warp / a.ksegs;
    const uint8_t *act = a.act + (int64_t)blockIdx.y * a.act_stride;

    ActBlock ab[kTPB];
    bool valid[kTPB];
    const int64_t b0 = (int64_t)kseg * GGB_DEV_SEG_BLOCKS + lane;
#pragma unroll
    for (int t = 0; t < kTPB; t++) {
        valid[t] = b0 + 32 * t < a.nb;
        load_act(ab[t], act, a.cols, b0 + 32 * t, valid[t]);
    }

    const int64_t row_begin = (int64_t)blockIdx.x * a.rows_per_cta;
    int64_t row_end = row_begin + a.rows_per_cta;
    if (row_end > a.total_rows) row_end = a.total_rows;
    for (int64_t base = row_begin; base < row_end; base += a.rows_per_iter) {
        const int64_t gr = base + rsub;
        const bool live = gr < row_end;
        // locate the matrix this global row belongs to
        int mi = 0;
        int64_t r = gr;
        if (live) {
            while (mi + 1 < a.n_mats && r >= a.m[mi].rows) {
                r -= a.m[mi].rows;
                mi++;
            }
        }
        float seg = 0.0f;
        if (live) seg = segment_dot<TYPE>(a.m[mi].qs, a.m[mi].d, r, a.nb, b0, ab, valid);
        if (a.ksegs == 1) {
            if (live && lane == 0) {
                float *dst = a.m[mi].dst + (int64_t)blockIdx.y * a.dst_stride;
                float v = seg;
                if (a.m[mi].resid) v = __fadd_rn(v, a.m[mi].resid[(int64_t)blockIdx.y * a.dst_stride + r]);
                dst[r] = v;
            }
        } else {
            if (lane == 0) s_part[rsub * a.ksegs + kseg] = seg;
            __syncthreads();
            if (live && kseg == 0 && lane == 0) {
                float total = s_part[rsub * a.ksegs];
                for (int s = 1; s < a.ksegs; s++) total = __fadd_rn(total, s_part[rsub * a.ksegs + s]);
                float *dst = a.m[mi].dst + (int64_t)blockIdx.y * a.dst_stride;
                if (a.m[mi].resid) total = __fadd_rn(total, a.m[mi].resid[(int64_t)blockIdx.y * a.dst_stride + r]);
                dst[r] = total;
            }
            __syncthreads();
        }
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
    const int ksegs = (int)((a.nb + GGB_DEV_SEG_BLOCKS - 1) / GGB_DEV_SEG_BLOCKS);
    float g[4], u[4];
    // segments in order (DEV): the lane's activation blocks are reloaded per segment
    for (int seg = 0; seg < ksegs; seg++) {
        ActBlock ab[kTPB];
        bool valid[kTPB];
        const int64_t b0 = (int64_t)seg * GGB_DEV_SEG_BLOCKS + lane;
#pragma unroll
        for (int t = 0; t < kTPB; t++) {
            valid[t] = b0 + 32 * t < a.nb;
            load_act(ab[t], a.act, a.cols, b0 + 32 * t, valid[t]);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int64_t r = tile_row0 + warp * 4 + i;
            const float gs = segment_dot<TYPE>(a.gate_qs, a.gate_d, r, a.nb, b0, ab, valid);
            const float us = segment_dot<TYPE>(a.up_qs, a.up_d, r, a.nb, b0, ab, valid);
            g[i] = seg == 0 ? gs : __fadd_rn(g[i], gs);
            u[i] = seg == 0 ? us : __fadd_rn(u[i], us);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            s_g[warp * 4 + i] = g[i];
            s_u[warp * 4 + i] = u[i];
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
    a.ksegs = (int)((a.nb + GGB_DEV_SEG_BLOCKS - 1) / GGB_DEV_SEG_BLOCKS);
    GGB_CHECK_ARG(a.ksegs <= 12, "K = %lld too large for one CTA (max 24576)", (long long)a.cols);
    a.rows_per_iter = a.ksegs >= 8 ? 1 : 8 / a.ksegs;
    const int warps = a.rows_per_iter * a.ksegs;
    const int64_t target_ctas = (int64_t)ggb_sm_count() * 8;
    int64_t rpc = (a.total_rows + target_ctas - 1) / target_ctas;
    rpc = ((rpc + a.rows_per_iter - 1) / a.rows_per_iter) * a.rows_per_iter;
    if (rpc < a.rows_per_iter) rpc = a.rows_per_iter;
    a.rows_per_cta = rpc;
    const unsigned grid_x = (unsigned)((a.total_rows + rpc - 1) / rpc);
    const dim3 grid(grid_x, (unsigned)T);
    const size_t smem = sizeof(float) * (size_t)warps;
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
