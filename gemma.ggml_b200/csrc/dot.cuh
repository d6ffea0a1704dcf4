// dot.cuh — the per-block arithmetic shared by the mat-vec kernels (matvec.cu) and the persistent
// decode kernel (decoder.cu): a lane's Q8_0 activation blocks in registers, one Q4_0 / Q8_0 weight
// block dotted against them with dp4a, and the scalar reference's per-block term
//   Q4_0: (float)sumi * d_x * d_y   (left to right)      Q8_0: (float)sumi * (d_x * d_y)
// ([UPSTREAM] ggml_vec_dot_q4_0_q8_0 / ggml_vec_dot_q8_0_q8_0 scalar branches; called through the
// vec_dot pointer at ref: src/hpc.cpp:35).
#pragma once

#include "common.cuh"

constexpr int kTPB = GGB_DEV_SEG_BLOCKS / 32; // blocks per lane per 64-block segment = 2

// activation block held by a lane
struct ActBlock {
    int a[8];  // 32 int8
    float d;   // fp16 scale widened
    int bias;  // -8 * sum(a): folds Q4_0's "- 8" out of the inner loop
};

// act: planar Q8_0 row (cols int8, then cols/32 fp16 scales); generic pointer (global or shared)
__device__ __forceinline__ void load_act(ActBlock &ab, const uint8_t *act, int64_t cols, int64_t b, bool valid) {
    if (valid) {
        const uint4 lo = *reinterpret_cast<const uint4 *>(act + b * 32);
        const uint4 hi = *reinterpret_cast<const uint4 *>(act + b * 32 + 16);
        ab.a[0] = lo.x; ab.a[1] = lo.y; ab.a[2] = lo.z; ab.a[3] = lo.w;
        ab.a[4] = hi.x; ab.a[5] = hi.y; ab.a[6] = hi.z; ab.a[7] = hi.w;
        ab.d = __half2float(reinterpret_cast<const __half *>(act + cols)[b]);
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s = __dp4a(ab.a[i], 0x01010101, s);
        ab.bias = -8 * s;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) ab.a[i] = 0;
        ab.d = 0.f;
        ab.bias = 0;
    }
}

// same, reading through L2 (data produced by another CTA of the same kernel)
__device__ __forceinline__ void load_act_cg(ActBlock &ab, const uint8_t *act, int64_t cols, int64_t b, bool valid) {
    if (valid) {
        const uint4 lo = __ldcg(reinterpret_cast<const uint4 *>(act + b * 32));
        const uint4 hi = __ldcg(reinterpret_cast<const uint4 *>(act + b * 32 + 16));
        ab.a[0] = lo.x; ab.a[1] = lo.y; ab.a[2] = lo.z; ab.a[3] = lo.w;
        ab.a[4] = hi.x; ab.a[5] = hi.y; ab.a[6] = hi.z; ab.a[7] = hi.w;
        ab.d = __half2float(__ushort_as_half(__ldcg(reinterpret_cast<const unsigned short *>(act + cols) + b)));
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) s = __dp4a(ab.a[i], 0x01010101, s);
        ab.bias = -8 * s;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) ab.a[i] = 0;
        ab.d = 0.f;
        ab.bias = 0;
    }
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

// weight payload of one block, loaded from global (streaming) or shared memory
template <int TYPE> struct Payload;
template <> struct Payload<GGB_TYPE_Q4_0> {
    uint4 v;
    __device__ __forceinline__ void load(const uint8_t *qs, int64_t blk) { v = ldg_stream_u4(qs + blk * 16); }
    __device__ __forceinline__ void load_smem(const uint8_t *qs, int blk) {
        v = *reinterpret_cast<const uint4 *>(qs + blk * 16);
    }
    __device__ __forceinline__ float term(float dx, const ActBlock &ab) const { return q4_term(v, dx, ab); }
};
template <> struct Payload<GGB_TYPE_Q8_0> {
    uint4 v0, v1;
    __device__ __forceinline__ void load(const uint8_t *qs, int64_t blk) {
        v0 = ldg_stream_u4(qs + blk * 32);
        v1 = ldg_stream_u4(qs + blk * 32 + 16);
    }
    __device__ __forceinline__ void load_smem(const uint8_t *qs, int blk) {
        v0 = *reinterpret_cast<const uint4 *>(qs + blk * 32);
        v1 = *reinterpret_cast<const uint4 *>(qs + blk * 32 + 16);
    }
    __device__ __forceinline__ float term(float dx, const ActBlock &ab) const { return q8_term(v0, v1, dx, ab); }
};

// dot of one 128-block segment of one row, weights in GLOBAL memory: lane partial (blocks in order,
// from 0.0f) + xor-butterfly.  b0 = first block of this lane (segment start + lane).
template <int TYPE>
__device__ __forceinline__ float segment_dot(const uint8_t *qs, const __half *d, int64_t row, int64_t nb, int64_t b0,
                                             const ActBlock (&ab)[kTPB], const bool (&valid)[kTPB]) {
    Payload<TYPE> w[kTPB];
    float dx[kTPB];
#pragma unroll
    for (int t = 0; t < kTPB; t++) {
        if (valid[t]) {
            const int64_t blk = row * nb + b0 + 32 * t;
            w[t].load(qs, blk);
            dx[t] = __half2float(d[blk]);
        }
    }
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < kTPB; t++) {
        if (valid[t]) acc = __fadd_rn(acc, w[t].term(dx[t], ab[t]));
    }
    return warp_sum_f32(acc);
}

// ---- explicit shared-space loads (32-bit shared addresses).  Going through generic pointers made nvcc
// emit LD.E + local-memory bounces in the persistent kernel; these are guaranteed LDS.
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float lds_half_as_float(uint32_t addr) {
    unsigned short h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __half2float(__ushort_as_half(h));
}

// Branch-free dot of one segment of one row staged in shared memory.  pay = shared address of this lane's
// first block payload, scl = shared address of its first fp16 scale; block t of the lane sits 32 blocks
// further.  Blocks beyond the row (live[t] false) are still loaded (any in-bounds shared address) and
// dropped with a select, so the loop has no control flow and every value stays in registers.
template <int TYPE, int BPL>
__device__ __forceinline__ float row_dot_lds(uint32_t pay, uint32_t scl, const ActBlock (&ab)[BPL],
                                             const bool (&live)[BPL]) {
    constexpr int PB = TYPE == GGB_TYPE_Q4_0 ? 16 : 32;
    uint4 w0[BPL], w1[BPL];
    float dx[BPL];
#pragma unroll
    for (int t = 0; t < BPL; t++) {
        w0[t] = lds_u4(pay + t * 32 * PB);
        if (TYPE == GGB_TYPE_Q8_0) w1[t] = lds_u4(pay + t * 32 * PB + 16);
        dx[t] = lds_half_as_float(scl + t * 64);
    }
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < BPL; t++) {
        const float term = TYPE == GGB_TYPE_Q4_0 ? q4_term(w0[t], dx[t], ab[t]) : q8_term(w0[t], w1[t], dx[t], ab[t]);
        const float next = __fadd_rn(acc, term);
        acc = live[t] ? next : acc;
    }
    return warp_sum_f32(acc);
}

// The four block terms of one pass, with the four dp4a chains advanced in lock step: a warp issues in order
// and the persistent kernel runs only 3 consumer warps per scheduler, so instruction-level parallelism inside
// the warp is what hides the 5-cycle dp4a / 4-cycle ALU latencies.  Same per-block arithmetic as q4_term /
// q8_term (integer dots are exact; the fp32 part is per block).
// NA = number of distinct activation blocks held: 4 (slot t uses ab[t]) or 2 (two rows share the
// same two blocks: slot t uses ab[t & 1]).
template <int TYPE, int NA>
__device__ __forceinline__ void pass_terms(const uint4 (&w0)[4], const uint4 (&w1)[4], const float (&dx)[4],
                                           const ActBlock (&abx)[NA], float (&term)[4]) {
    int sumi[4];
#define ab(t) abx[NA == 4 ? (t) : ((t) & 1)]
    if (TYPE == GGB_TYPE_Q4_0) {
#pragma unroll
        for (int t = 0; t < 4; t++) sumi[t] = ab(t).bias;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t q[4];
#pragma unroll
            for (int t = 0; t < 4; t++) q[t] = i == 0 ? w0[t].x : (i == 1 ? w0[t].y : (i == 2 ? w0[t].z : w0[t].w));
#pragma unroll
            for (int t = 0; t < 4; t++) sumi[t] = __dp4a((int)(q[t] & 0x0F0F0F0Fu), ab(t).a[i], sumi[t]);
#pragma unroll
            for (int t = 0; t < 4; t++) sumi[t] = __dp4a((int)((q[t] >> 4) & 0x0F0F0F0Fu), ab(t).a[4 + i], sumi[t]);
        }
#pragma unroll
        for (int t = 0; t < 4; t++) term[t] = __fmul_rn(__fmul_rn((float)sumi[t], dx[t]), ab(t).d);
    } else {
#pragma unroll
        for (int t = 0; t < 4; t++) sumi[t] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const uint4 &v = i < 4 ? w0[t] : w1[t];
                const int k = i & 3;
                const uint32_t q = k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
                sumi[t] = __dp4a((int)q, ab(t).a[i], sumi[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < 4; t++) term[t] = __fmul_rn((float)sumi[t], __fmul_rn(dx[t], ab(t).d));
    }
#undef ab
}

// shared-memory variant with a compile-time number of blocks per lane (BPL = ceil(segment blocks / 32)):
// K = 2048 needs 2, so only 2 activation blocks occupy registers and no predicated dead work is issued.
// `live[t]` masks the ragged tail (block index >= nb) and idle warps.
template <int TYPE, int BPL>
__device__ __forceinline__ float segment_dot_smem_n(const uint8_t *qs_row, const __half *d_row, int b0,
                                                    const ActBlock (&ab)[BPL], const bool (&live)[BPL]) {
    Payload<TYPE> w[BPL];
    float dx[BPL];
#pragma unroll
    for (int t = 0; t < BPL; t++) {
        if (live[t]) {
            w[t].load_smem(qs_row, b0 + 32 * t);
            dx[t] = __half2float(d_row[b0 + 32 * t]);
        }
    }
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < BPL; t++) {
        if (live[t]) acc = __fadd_rn(acc, w[t].term(dx[t], ab[t]));
    }
    return warp_sum_f32(acc);
}

// same with the row's payload / scale planes staged in SHARED memory (row-local block index)
template <int TYPE>
__device__ __forceinline__ float segment_dot_smem(const uint8_t *qs_row, const __half *d_row, int b0,
                                                  const ActBlock (&ab)[kTPB], const bool (&valid)[kTPB]) {
    Payload<TYPE> w[kTPB];
    float dx[kTPB];
#pragma unroll
    for (int t = 0; t < kTPB; t++) {
        if (valid[t]) {
            w[t].load_smem(qs_row, b0 + 32 * t);
            dx[t] = __half2float(d_row[b0 + 32 * t]);
        }
    }
    float acc = 0.0f;
#pragma unroll
    for (int t = 0; t < kTPB; t++) {
        if (valid[t]) acc = __fadd_rn(acc, w[t].term(dx[t], ab[t]));
    }
    return warp_sum_f32(acc);
}
