// dot.cuh — the per-block arithmetic of the per-operator mat-vec kernels (matvec.cu): one Q4_0 / Q8_0
// weight block dotted with dp4a against a Q8_0 activation block, the scalar reference's per-block term
//   Q4_0: (float)sumi * d_x * d_y   (left to right)      Q8_0: (float)sumi * (d_x * d_y)
// ([UPSTREAM] ggml_vec_dot_q4_0_q8_0 / ggml_vec_dot_q8_0_q8_0 scalar branches; called through the
// vec_dot pointer at ref: src/hpc.cpp:35), and a row dot that adds the terms in the reference's order
// (one fp32 accumulator, blocks in order).
#pragma once

#include "common.cuh"

// activation block held by a lane
struct ActBlock {
    int a[8];  // 32 int8
    float d;   // fp16 scale widened
    int bias;  // -8 * sum(a): folds Q4_0's "- 8" out of the inner loop
};

// act: planar Q8_0 row (cols int8, then cols/32 fp16 scales)
__device__ __forceinline__ void load_act(ActBlock &ab, const uint8_t *act, int64_t cols, int64_t b) {
    const uint4 lo = *reinterpret_cast<const uint4 *>(act + b * 32);
    const uint4 hi = *reinterpret_cast<const uint4 *>(act + b * 32 + 16);
    ab.a[0] = lo.x; ab.a[1] = lo.y; ab.a[2] = lo.z; ab.a[3] = lo.w;
    ab.a[4] = hi.x; ab.a[5] = hi.y; ab.a[6] = hi.z; ab.a[7] = hi.w;
    ab.d = __half2float(reinterpret_cast<const __half *>(act + cols)[b]);
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s = __dp4a(ab.a[i], 0x01010101, s);
    ab.bias = -8 * s;
}

__device__ __forceinline__ float q4_term(const uint4 w, float dx, const ActBlock &ab) {
    int sumi = ab.bias;
    const uint32_t q[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int lo = (int)(q[i] & 0x0F0F0F0Fu);
        const int hi = (int)((q[i] >> 4) & 0x0F0F0F0Fu);
        sumi = __dp4a(lo, ab.a[i], sumi);
        sumi = __dp4a(hi, ab.a[4 + i], sumi);
    }
    return __fmul_rn(__fmul_rn((float)sumi, dx), ab.d);
}

__device__ __forceinline__ float q8_term(const uint4 w0, const uint4 w1, float dx, const ActBlock &ab) {
    int sumi = 0;
    sumi = __dp4a((int)w0.x, ab.a[0], sumi);
    sumi = __dp4a((int)w0.y, ab.a[1], sumi);
    sumi = __dp4a((int)w0.z, ab.a[2], sumi);
    sumi = __dp4a((int)w0.w, ab.a[3], sumi);
    sumi = __dp4a((int)w1.x, ab.a[4], sumi);
    sumi = __dp4a((int)w1.y, ab.a[5], sumi);
    sumi = __dp4a((int)w1.z, ab.a[6], sumi);
    sumi = __dp4a((int)w1.w, ab.a[7], sumi);
    return __fmul_rn((float)sumi, __fmul_rn(dx, ab.d));
}

// One row (planar weights in global memory) · one planar Q8_0 activation row, for a whole warp:
// per round of 32 blocks lane l computes the term of block base + l, then the 32 terms are added to the ONE
// accumulator in block order (a shuffle chain every lane runs redundantly): exactly
//   float sumf = 0; for (i = 0; i < nb; i++) sumf += term(i);
// of the scalar reference.  Returns the sum in every lane.
template <int TYPE>
__device__ __forceinline__ float row_dot_ref(const uint8_t *qs, const __half *d, int64_t row, int64_t nb,
                                             const uint8_t *act, int64_t cols, int lane) {
    constexpr int PB = TYPE == GGB_TYPE_Q4_0 ? 16 : 32;
    float acc = 0.0f;
#pragma unroll 1
    for (int64_t base = 0; base < nb; base += 32) {
        const int64_t b = base + lane;
        float term = 0.0f;
        if (b < nb) {
            ActBlock ab;
            load_act(ab, act, cols, b);
            const int64_t blk = row * nb + b;
            const float dx = __half2float(d[blk]);
            const uint4 w0 = ldg_stream_u4(qs + blk * PB);
            if (TYPE == GGB_TYPE_Q4_0) {
                term = q4_term(w0, dx, ab);
            } else {
                const uint4 w1 = ldg_stream_u4(qs + blk * PB + 16);
                term = q8_term(w0, w1, dx, ab);
            }
        }
        const int cnt = (int)(nb - base < 32 ? nb - base : 32);
        if (cnt == 32) {
#pragma unroll
            for (int i = 0; i < 32; i++) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, term, i));
        } else {
            for (int i = 0; i < cnt; i++) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, term, i));
        }
    }
    return acc;
}
