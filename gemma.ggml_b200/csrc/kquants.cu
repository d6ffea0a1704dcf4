// kquants.cu — SURVEY.md 8f-2: the formats the reference's own offload kernels target.
//   Q4_K x Q8_K   ref: src/kernals.cl:48-111  (matrix_multiply_q4_K :166-214)
//   Q6_K x Q8_K   ref: src/kernals.cl:114-156 ([UPSTREAM] ggml's QK_K = 256 layout; the reference's copy indexes
//                 the block as QK_K = 64 and reads an uninitialised aux8 tail - its host never launches it,
//                 src/hpc.cpp:254-257 - so the pinned behaviour is ggml's scalar vec_dot)
//   F16 x F16     ref: src/kernals.cl:158-164 (double accumulator)
//   quantize_row_q8_K  [UPSTREAM] ggml INIT stage of mul_mat for K-quant weights (block at src/kernals.cl:29-34)
//
// The reference runs one work-item per output with a private aux8[256] and a serial loop.  Here a warp owns an
// output row and a super-block is spread over the warp: lane L holds the 8 consecutive aux8 elements 8L..8L+7, so
// element l of the lane IS the reference's lane-sum index (aux32[l], sums[l], l = e mod 8).  The integer part
// (aux32[l] = sum over the super-block of scale * q8 * a, sumi = sum bsums * mins) is exact in any order and is
// reduced with a folding butterfly (9 shuffles for the eight sums); the fp32 part keeps the reference's order: eight
// chains sums[l] += d * aux32[l] advanced super-block by super-block, sumf -= dmin * sumi, then sumf += sums[0..7].
// Results are bit-identical to the scalar code (tests compare bits with oracle/_ref, the reference's own source).
//
// Weights and activations stay in ggml's AoS blocks (144 / 210 / 292 B): at 4.5 / 6.56 bits per weight a row of a
// super-block is one 128 B + one 16 B line, so the warp's loads are already contiguous and a repack buys nothing.
#include "common.cuh"

namespace {

constexpr int kQK = 256;
constexpr int kQ4K = 144, kQ6K = 210, kQ8K = 292;
constexpr int kWarps = 8;

__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }
__device__ __forceinline__ uint32_t ld_u16(const uint8_t *p) { return *reinterpret_cast<const uint16_t *>(p); }
// 8 bytes from a 2-byte aligned address
__device__ __forceinline__ void ld8_a2(const uint8_t *p, uint32_t &lo, uint32_t &hi) {
    lo = ld_u16(p) | (ld_u16(p + 2) << 16);
    hi = ld_u16(p + 4) | (ld_u16(p + 6) << 16);
}
__device__ __forceinline__ int sx8(uint32_t w, int i) { return (int)(int8_t)(w >> (8 * i)); }

// v[0..7] per lane -> sum over the warp of v[l], returned in every lane of the group {4l .. 4l+3}
__device__ __forceinline__ int fold8(int (&v)[8], int lane) {
    int w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int send = b4 ? v[k] : v[4 + k];
        const int keep = b4 ? v[4 + k] : v[k];
        w4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int send = b3 ? w4[k] : w4[2 + k];
        const int keep = b3 ? w4[2 + k] : w4[k];
        w2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    int w = (b2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? w2[0] : w2[1], 4);
    w += __shfl_xor_sync(0xffffffffu, w, 2);
    w += __shfl_xor_sync(0xffffffffu, w, 1);
    return w; // l = 4*bit4 + 2*bit3 + bit2 = lane >> 2
}

__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// the eight chains live in lane groups; the final "for l: sumf += sums[l]" reads them in order
__device__ __forceinline__ float finish_lanes(float sumf, float sums) {
#pragma unroll
    for (int l = 0; l < 8; l++) sumf = __fadd_rn(sumf, __shfl_sync(0xffffffffu, sums, 4 * l));
    return sumf;
}

// the same with the eight chains in lanes 0..7
__device__ __forceinline__ float finish_lanes8(float sumf, float sums) {
#pragma unroll
    for (int l = 0; l < 8; l++) sumf = __fadd_rn(sumf, __shfl_sync(0xffffffffu, sums, l));
    return sumf;
}

__device__ __forceinline__ float dot_q4_K(const uint8_t *x, const uint8_t *y, int nb, int lane) {
    float sums = 0.0f, sumf = 0.0f;
    const int c = lane >> 3, hi = lane & 4, j = lane >> 2;
#pragma unroll 1
    for (int i = 0; i < nb; i++, x += kQ4K, y += kQ8K) {
        const uint32_t dd = ld_u32(x);
        uint32_t u0 = ld_u32(x + 4), u1 = ld_u32(x + 8), u2 = ld_u32(x + 12);
        // 6-bit scales / mins (src/kernals.cl:75-80)
        const uint32_t u3 = ((u2 >> 4) & 0x0f0f0f0fu) | (((u1 >> 6) & 0x03030303u) << 4);
        const uint32_t uaux = u1 & 0x3f3f3f3fu;
        u1 = (u2 & 0x0f0f0f0fu) | (((u0 >> 6) & 0x03030303u) << 4);
        u2 = uaux;
        u0 &= 0x3f3f3f3fu;
        const int scale = (int)(((j < 4 ? u0 : u1) >> (8 * (j & 3))) & 0xff);
        const uint8_t *q4 = x + 16 + 32 * c + 8 * (lane & 3);
        uint32_t w0 = ld_u32(q4), w1 = ld_u32(q4 + 4);
        if (hi) { w0 >>= 4; w1 >>= 4; }
        w0 &= 0x0f0f0f0fu;
        w1 &= 0x0f0f0f0fu;
        const uint32_t a0 = ld_u32(y + 4 + 8 * lane), a1 = ld_u32(y + 8 + 8 * lane);
        int v[8];
#pragma unroll
        for (int l = 0; l < 4; l++) {
            v[l] = scale * (sx8(a0, l) * (int)((w0 >> (8 * l)) & 0xff));
            v[4 + l] = scale * (sx8(a1, l) * (int)((w1 >> (8 * l)) & 0xff));
        }
        const int aux = fold8(v, lane);
        // sumi = sum_j bsums[j] * mins[j / 2]
        int sm = 0;
        if (lane < 16) {
            const int bs = (int)(int16_t)ld_u16(y + 260 + 2 * lane);
            const int mj = lane >> 1;
            sm = bs * (int)(((mj < 4 ? u2 : u3) >> (8 * (mj & 3))) & 0xff);
        }
        const int sumi = warp_sum_i32(sm);
        const float yd = __uint_as_float(ld_u32(y));
        const float d = __fmul_rn(half_bits_to_float((uint16_t)(dd & 0xffff)), yd);
        const float dmin = __fmul_rn(half_bits_to_float((uint16_t)(dd >> 16)), yd);
        sums = __fadd_rn(sums, __fmul_rn(d, (float)aux));
        sumf = __fsub_rn(sumf, __fmul_rn(dmin, (float)sumi));
    }
    return finish_lanes(sumf, sums);
}

__device__ __forceinline__ float dot_q6_K(const uint8_t *x, const uint8_t *y, int nb, int lane) {
    float sums = 0.0f;
    const int c = lane >> 4, quarter = (lane >> 2) & 3, l0 = 8 * (lane & 3);
#pragma unroll 1
    for (int i = 0; i < nb; i++, x += kQ6K, y += kQ8K) {
        uint32_t q0, q1, h0, h1;
        ld8_a2(x + 64 * c + 32 * (quarter & 1) + l0, q0, q1);
        ld8_a2(x + 128 + 32 * c + l0, h0, h1);
        if (quarter & 2) { q0 >>= 4; q1 >>= 4; }
        q0 &= 0x0f0f0f0fu;
        q1 &= 0x0f0f0f0fu;
        h0 = ((h0 >> (2 * quarter)) & 0x03030303u) << 4;
        h1 = ((h1 >> (2 * quarter)) & 0x03030303u) << 4;
        q0 |= h0;
        q1 |= h1;
        const int scale = (int)(int8_t)x[192 + (lane >> 1)];
        const uint32_t a0 = ld_u32(y + 4 + 8 * lane), a1 = ld_u32(y + 8 + 8 * lane);
        int v[8];
#pragma unroll
        for (int l = 0; l < 4; l++) {
            v[l] = scale * (sx8(a0, l) * ((int)((q0 >> (8 * l)) & 0xff) - 32));
            v[4 + l] = scale * (sx8(a1, l) * ((int)((q1 >> (8 * l)) & 0xff) - 32));
        }
        const int aux = fold8(v, lane);
        const float d = __fmul_rn(half_bits_to_float((uint16_t)ld_u16(x + 208)), __uint_as_float(ld_u32(y)));
        sums = __fadd_rn(sums, __fmul_rn(d, (float)aux));
    }
    return finish_lanes(0.0f, sums);
}

// fp16 x fp16 products are exact in fp32 and in double; the double sum is insensitive to its order at the
// precision the final rounding to fp32 keeps (same argument as attention.cu)
__device__ __forceinline__ float dot_f16(const uint8_t *x, const uint8_t *y, int64_t n, int lane) {
    const __half2 *x2 = reinterpret_cast<const __half2 *>(x), *y2 = reinterpret_cast<const __half2 *>(y);
    const __half *xh = reinterpret_cast<const __half *>(x), *yh = reinterpret_cast<const __half *>(y);
    double acc = 0.0;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 3) == 0) {
        for (int64_t i = lane; i < n / 2; i += 32) {
            const float2 a = __half22float2(x2[i]), b = __half22float2(y2[i]);
            acc = __fma_rn((double)a.x, (double)b.x, acc);
            acc = __fma_rn((double)a.y, (double)b.y, acc);
        }
        if ((n & 1) && lane == 0)
            acc = __fma_rn((double)__half2float(xh[n - 1]), (double)__half2float(yh[n - 1]), acc);
    } else { // odd row pitch: 2-byte loads
        for (int64_t i = lane; i < n; i += 32)
            acc = __fma_rn((double)__half2float(xh[i]), (double)__half2float(yh[i]), acc);
    }
    return __double2float_rn(warp_sum_f64(acc));
}

// grid (ceil(rows / 8), T)
template <int TYPE>
__global__ void __launch_bounds__(kWarps * 32)
raw_matvec_kernel(const uint8_t *__restrict__ w, int64_t rows, int64_t nb01, int64_t K, const uint8_t *__restrict__ act,
                  int64_t row_size, float *__restrict__ dst, int64_t dst_stride) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const uint8_t *x = w + row * nb01;
    const uint8_t *y = act + (int64_t)blockIdx.y * row_size;
    float s;
    if (TYPE == GGB_TYPE_Q4_K) s = dot_q4_K(x, y, (int)(K / kQK), lane);
    else if (TYPE == GGB_TYPE_Q6_K) s = dot_q6_K(x, y, (int)(K / kQK), lane);
    else s = dot_f16(x, y, K, lane);
    if (lane == 0) dst[(int64_t)blockIdx.y * dst_stride + row] = s;
}

// [UPSTREAM] quantize_row_q8_K_reference for one 256-element block spread over a warp: lane L holds elements 8L..8L+7.
// Returns the lane's eight codes and the block's d (0 for an all-zero block); shared by the stand-alone quantiser and by
// the mat-vec kernels that quantise their fp32 activation row while they stage it.
__device__ __forceinline__ float q8k_quantize_lane(const float (&v)[8], int lane, int (&q)[8]) {
    // first element attaining max |x| (strict > in the scalar loop)
    float amax = 0.0f, mx = 0.0f;
    int idx = 0;
#pragma unroll
    for (int l = 0; l < 8; l++) {
        const float ax = fabsf(v[l]);
        if (ax > amax) { amax = ax; mx = v[l]; idx = 8 * lane + l; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const float oa = __shfl_xor_sync(0xffffffffu, amax, off), om = __shfl_xor_sync(0xffffffffu, mx, off);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
        if (oa > amax || (oa == amax && oi < idx)) { amax = oa; mx = om; idx = oi; }
    }
    if (!(amax > 0.0f)) { // "if (!amax)": all zero (a NaN block cannot reach amax > 0 either; ggml leaves it at 0)
#pragma unroll
        for (int l = 0; l < 8; l++) q[l] = 0;
        return 0.0f;
    }
    const float iscale = __fdiv_rn(-128.0f, mx);
#pragma unroll
    for (int l = 0; l < 8; l++) {
        const float val = __fadd_rn(__fmul_rn(iscale, v[l]), 12582912.0f); // nearest_int: round half to even
        q[l] = min(127, (__float_as_int(val) & 0x007fffff) - 0x00400000);
    }
    return __fdiv_rn(1.0f, iscale);
}

// one warp per 256-element block
__global__ void __launch_bounds__(kWarps * 32)
quantize_q8_K_kernel(const float *__restrict__ x, int64_t n_blocks, uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    const float4 f0 = *reinterpret_cast<const float4 *>(x + b * kQK + 8 * lane);
    const float4 f1 = *reinterpret_cast<const float4 *>(x + b * kQK + 8 * lane + 4);
    const float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
    int q[8];
    const float d = q8k_quantize_lane(v, lane, q);
    uint8_t *o = out + b * kQ8K;
    int s = 0;
#pragma unroll
    for (int l = 0; l < 8; l++) s += q[l];
    const uint32_t p0 = (q[0] & 0xff) | ((q[1] & 0xff) << 8) | ((q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    const uint32_t p1 = (q[4] & 0xff) | ((q[5] & 0xff) << 8) | ((q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    *reinterpret_cast<uint32_t *>(o + 4 + 8 * lane) = p0;
    *reinterpret_cast<uint32_t *>(o + 8 + 8 * lane) = p1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (!(lane & 1)) *reinterpret_cast<int16_t *>(o + 260 + lane) = (int16_t)s; // bsums[lane / 2]
    if (lane == 0) *reinterpret_cast<float *>(o) = d;
}


// ---------------------------------------------------------------- Q4_K, packed form (resident weights)
// The AoS kernel above spends 14 shuffles and ~60 instructions on a 144-byte super-block; it reaches 0.8-0.9 TB/s.  For
// weights that stay on the device (ggb_mul_mat's cache) the super-block is re-laid out ONCE so that the four nibbles a
// lane-sum index l meets inside a 32-element sub-block - elements l, 8 + l, 16 + l, 24 + l - are adjacent:
//   block (144 B, same size) = d | dmin | scales[12] | X[8 sub-blocks][4 words], word p of sub-block j:
//   bits 4g .. 4g+3 = nibble of element 32 j + 8 g + 2 p, bits 16 + 4g .. = nibble of element 32 j + 8 g + 2 p + 1.
// The activation row is transposed the same way when a CTA stages it in shared memory (once per CTA, for all of its
// rows): qT[j][l] = bytes q8[32 j + 8 g + l], g = 0..3.  Then lane (s, j) - super-block s of four, sub-block j - turns
// its 16 bytes into eight dp4a (one per l), multiplies by its 6-bit scale, and a three-step folded butterfly inside
// the 8-lane group leaves aux32[l = j] of super-block s in lane (s, j).  The fp32 part is the scalar code's, in its
// order: the products d * aux32[l] of the four super-blocks are made in parallel, then lanes 0..7 add them to
// sums[l] one by one; lane 0 does the same for sumf -= dmin * sumi.  4 shuffles per super-block instead of 14.
constexpr int kActSB = 304; // staged activation super-block: qT 256 B | pair sums int[8] | d float | pad

__device__ __forceinline__ uint32_t ldg_u32(const uint8_t *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }

__global__ void __launch_bounds__(256) q4k_repack_kernel(const uint8_t *__restrict__ raw, int64_t rows, int64_t nb01, int nb,
                                                        uint8_t *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; // (row, super-block, sub-block j)
    if (idx >= rows * nb * 8) return;
    const int j = (int)(idx & 7);
    const int64_t rb = idx >> 3, row = rb / nb;
    const int sb = (int)(rb - row * nb);
    const uint8_t *x = raw + row * nb01 + (int64_t)sb * kQ4K;
    uint8_t *o = out + (row * nb + sb) * kQ4K;
    if (j < 4) *reinterpret_cast<uint32_t *>(o + 4 * j) = *reinterpret_cast<const uint32_t *>(x + 4 * j); // d, dmin, scales
    const uint8_t *q = x + 16 + 32 * (j >> 1);
    const int sh = (j & 1) * 4;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int g = 0; g < 4; g++)
        for (int l = 0; l < 8; l++) {
            const uint32_t nib = (q[8 * g + l] >> sh) & 0xFu;
            w[l >> 1] |= nib << (16 * (l & 1) + 4 * g);
        }
    *reinterpret_cast<uint4 *>(o + 16 + 16 * j) = make_uint4(w[0], w[1], w[2], w[3]);
}

// grid (x: row groups, strided; y: activation row); 8 warps; dynamic shared memory nb * kActSB
// F32ACT: `act` is the fp32 activation row (pitch row_size bytes); the CTA quantises it to Q8_K itself while staging
// (ggml's INIT stage of the mul_mat node, fused: the stand-alone quantiser launch per node goes away)
template <bool F32ACT>
__global__ void __launch_bounds__(256)
q4k_packed_kernel(const uint8_t *__restrict__ w, int64_t rows, int nb, const uint8_t *__restrict__ act, int64_t row_size,
                  float *__restrict__ dst, int64_t dst_stride) {
    extern __shared__ __align__(16) uint8_t s_act[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *y = act + (int64_t)blockIdx.y * row_size;
    // ---- stage the activation row, transposed
    if (F32ACT) {
        const float *xf = reinterpret_cast<const float *>(y);
        for (int sb = warp; sb < nb; sb += kWarps) { // a warp quantises one super-block at a time: lane = elements 8L..8L+7
            const float4 f0 = *reinterpret_cast<const float4 *>(xf + sb * kQK + 8 * lane);
            const float4 f1 = *reinterpret_cast<const float4 *>(xf + sb * kQK + 8 * lane + 4);
            const float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            int q[8];
            const float d = q8k_quantize_lane(v, lane, q);
            // element 8L + l = sub-block j = L / 4, group g = L % 4, lane-sum index l: byte g of qT[j][l]
            const int j = lane >> 2, g = lane & 3;
            uint8_t *base = s_act + sb * kActSB + j * 16 + g;
            int ssum = 0;
#pragma unroll
            for (int l = 0; l < 8; l++) {
                base[(l >> 2) * 128 + (l & 3) * 4] = (uint8_t)q[l];
                ssum += q[l];
            }
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
            if (g == 0) *reinterpret_cast<int *>(s_act + sb * kActSB + 256 + 4 * j) = ssum; // bsums[2j] + bsums[2j + 1]
            if (lane == 0) *reinterpret_cast<float *>(s_act + sb * kActSB + 288) = d;
        }
    } else {
    for (int i = tid; i < nb * 64; i += 256) { // word (sb, j, l)
        const int sb = i >> 6, j = (i >> 3) & 7, l = i & 7;
        const uint8_t *q8 = y + (int64_t)sb * kQ8K + 4 + 32 * j + l;
        const uint32_t v = (uint32_t)q8[0] | ((uint32_t)q8[8] << 8) | ((uint32_t)q8[16] << 16) | ((uint32_t)q8[24] << 24);
        *reinterpret_cast<uint32_t *>(s_act + sb * kActSB + (l >> 2) * 128 + j * 16 + (l & 3) * 4) = v;
    }
    for (int i = tid; i < nb * 9; i += 256) {
        const int sb = i / 9, k = i - sb * 9;
        const uint8_t *yb = y + (int64_t)sb * kQ8K;
        if (k < 8) {
            const int16_t *bs = reinterpret_cast<const int16_t *>(yb + 260);
            *reinterpret_cast<int *>(s_act + sb * kActSB + 256 + 4 * k) = (int)bs[2 * k] + (int)bs[2 * k + 1];
        } else {
            *reinterpret_cast<uint32_t *>(s_act + sb * kActSB + 288) = ld_u32(yb);
        }
    }
    }
    __syncthreads();
    // lane (s, c): super-block s of EIGHT, sub-blocks 2c and 2c + 1 (64 elements per lane: the fold runs over four lanes
    // only, and the scalar bookkeeping of an iteration is paid per eight super-blocks)
    const int s = lane >> 2, c = lane & 3;
    const bool b1 = lane & 2, b0 = lane & 1;
    for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < rows; row += (int64_t)gridDim.x * kWarps) {
        const uint8_t *xrow = w + row * (int64_t)nb * kQ4K;
        float sumsA = 0.0f, sumsB = 0.0f, sumf = 0.0f; // lane (0, c): sums[2c], sums[2c + 1]; lane 0: sumf
#pragma unroll 1
        for (int sb0 = 0; sb0 < nb; sb0 += 8) {
            const int sb = sb0 + s;
            const bool valid = sb < nb;
            const uint8_t *blk = xrow + (int64_t)(valid ? sb : 0) * kQ4K;
            const uint4 X0 = __ldg(reinterpret_cast<const uint4 *>(blk + 16 + 32 * c));
            const uint4 X1 = __ldg(reinterpret_cast<const uint4 *>(blk + 32 + 32 * c));
            const uint32_t dd = ldg_u32(blk);
            uint32_t u0 = ldg_u32(blk + 4), u1 = ldg_u32(blk + 8), u2 = ldg_u32(blk + 12);
            // 6-bit scales / mins (src/kernals.cl:75-80)
            const uint32_t u3 = ((u2 >> 4) & 0x0f0f0f0fu) | (((u1 >> 6) & 0x03030303u) << 4);
            const uint32_t uaux = u1 & 0x3f3f3f3fu;
            u1 = (u2 & 0x0f0f0f0fu) | (((u0 >> 6) & 0x03030303u) << 4);
            u2 = uaux;
            u0 &= 0x3f3f3f3fu;
            const uint32_t sc2 = ((c < 2 ? u0 : u1) >> (16 * (c & 1))) & 0xffffu; // scales of sub-blocks 2c, 2c + 1
            const uint32_t mn2 = ((c < 2 ? u2 : u3) >> (16 * (c & 1))) & 0xffffu;
            const uint8_t *ya = s_act + (valid ? sb : 0) * kActSB;
            int v[8];
#pragma unroll
            for (int h = 0; h < 2; h++) { // sub-block j = 2c + h
                const uint4 X = h ? X1 : X0;
                const int scale = (int)((sc2 >> (8 * h)) & 0xff);
                const uint4 qa = *reinterpret_cast<const uint4 *>(ya + 32 * c + 16 * h);       // qT[j][0..3]
                const uint4 qb = *reinterpret_cast<const uint4 *>(ya + 128 + 32 * c + 16 * h); // qT[j][4..7]
                const uint32_t xs[4] = {X.x, X.y, X.z, X.w};
                const uint32_t qs[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const uint32_t t1 = xs[p] & 0x0f0f0f0fu, t2 = (xs[p] >> 4) & 0x0f0f0f0fu;
                    const int pe = __dp4a((int)__byte_perm(t1, t2, 0x5140), (int)qs[2 * p], 0);
                    const int po = __dp4a((int)__byte_perm(t1, t2, 0x7362), (int)qs[2 * p + 1], 0);
                    v[2 * p] = h ? v[2 * p] + scale * pe : scale * pe;
                    v[2 * p + 1] = h ? v[2 * p + 1] + scale * po : scale * po;
                }
            }
            // fold inside the 4-lane group: lane (s, c) ends with aux32[2c] and aux32[2c + 1]
            int w4[4];
#pragma unroll
            for (int k = 0; k < 4; k++) w4[k] = (b1 ? v[4 + k] : v[k]) + __shfl_xor_sync(0xffffffffu, b1 ? v[k] : v[4 + k], 2);
            const int aux0 = (b0 ? w4[2] : w4[0]) + __shfl_xor_sync(0xffffffffu, b0 ? w4[0] : w4[2], 1);
            const int aux1 = (b0 ? w4[3] : w4[1]) + __shfl_xor_sync(0xffffffffu, b0 ? w4[1] : w4[3], 1);
            const int *sp = reinterpret_cast<const int *>(ya + 256);
            int sumi = (int)(mn2 & 0xff) * sp[2 * c] + (int)(mn2 >> 8) * sp[2 * c + 1];
            sumi += __shfl_xor_sync(0xffffffffu, sumi, 1);
            sumi += __shfl_xor_sync(0xffffffffu, sumi, 2);
            const float yd = *reinterpret_cast<const float *>(ya + 288);
            const float d = __fmul_rn(half_bits_to_float((uint16_t)(dd & 0xffff)), yd);
            const float dmin = __fmul_rn(half_bits_to_float((uint16_t)(dd >> 16)), yd);
            const float tA = valid ? __fmul_rn(d, (float)aux0) : 0.0f; // x + 0.0f == x: a missing super-block adds nothing
            const float tB = valid ? __fmul_rn(d, (float)aux1) : 0.0f;
            const float tm = valid ? __fmul_rn(dmin, (float)sumi) : 0.0f;
            // the scalar order: the eight super-blocks one after the other
            sumsA = __fadd_rn(sumsA, tA);
            sumsB = __fadd_rn(sumsB, tB);
            sumf = __fsub_rn(sumf, tm);
#pragma unroll
            for (int k = 1; k < 8; k++) {
                sumsA = __fadd_rn(sumsA, __shfl_sync(0xffffffffu, tA, c + 4 * k)); // (meaningful in lanes 0..3)
                sumsB = __fadd_rn(sumsB, __shfl_sync(0xffffffffu, tB, c + 4 * k));
                sumf = __fsub_rn(sumf, __shfl_sync(0xffffffffu, tm, 4 * k));       // (meaningful in lane 0)
            }
        }
#pragma unroll
        for (int l = 0; l < 8; l++) {
            const float a = __shfl_sync(0xffffffffu, (l & 1) ? sumsB : sumsA, l >> 1);
            sumf = __fadd_rn(sumf, a);
        }
        if (lane == 0) dst[(int64_t)blockIdx.y * dst_stride + row] = sumf;
    }
}

// ---------------------------------------------------------------- Q6_K, packed form (resident weights)
// Same idea for Q6_K.  The 6-bit codes are expanded to int8 (q - 32) once - a row becomes W[nb][16 sub-blocks][16 B] |
// scales[nb][16] | d[nb] (274 instead of 210 bytes per super-block: the kernel is bound by instruction issue, not by
// HBM) - with word p of sub-block j = [a(g0, l = 2p), a(g1, 2p), a(g0, 2p + 1), a(g1, 2p + 1)], element = 16 j + 8 g + l,
// and the activation words are staged in the same arrangement: two masked dp4a per word give the two lane sums.
// Lane (s, c): super-block s of four, sub-blocks 2c and 2c + 1 of sixteen; the fold runs over the 8-lane group.
constexpr int kActSB6 = 272; // staged activation super-block: qT 256 B | d float | pad

__host__ __device__ inline int64_t q6k_packed_pitch(int nb) { return ((int64_t)nb * 274 + 15) & ~(int64_t)15; }

__global__ void __launch_bounds__(256) q6k_repack_kernel(const uint8_t *__restrict__ raw, int64_t rows, int64_t nb01, int nb,
                                                        uint8_t *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; // (row, super-block, sub-block j)
    if (idx >= rows * nb * 16) return;
    const int j = (int)(idx & 15);
    const int64_t rb = idx >> 4, row = rb / nb;
    const int sb = (int)(rb - row * nb);
    const uint8_t *x = raw + row * nb01 + (int64_t)sb * kQ6K;
    uint8_t *orow = out + row * q6k_packed_pitch(nb);
    // element e of the scalar code's aux8 (ggml QK_K = 256 layout): chunk c = e / 128, quarter = (e % 128) / 32, i = e % 32
    auto code = [&](int e) {
        const int c = e >> 7, quarter = (e >> 5) & 3, i = e & 31;
        const uint8_t ql = x[64 * c + 32 * (quarter & 1) + i], qh = x[128 + 32 * c + i];
        const int lo = (quarter & 2) ? (ql >> 4) : (ql & 0xF);
        return (int)(int8_t)(lo | (((qh >> (2 * quarter)) & 3) << 4)) - 32;
    };
    uint32_t w[4];
    for (int p = 0; p < 4; p++) {
        const int e0 = 16 * j + 2 * p;
        w[p] = (uint32_t)(uint8_t)code(e0) | ((uint32_t)(uint8_t)code(e0 + 8) << 8) | ((uint32_t)(uint8_t)code(e0 + 1) << 16) |
               ((uint32_t)(uint8_t)code(e0 + 9) << 24);
    }
    *reinterpret_cast<uint4 *>(orow + ((int64_t)sb * 16 + j) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    orow[(int64_t)nb * 256 + sb * 16 + j] = x[192 + j];
    if (j == 0) *reinterpret_cast<uint16_t *>(orow + (int64_t)nb * 272 + 2 * sb) = *reinterpret_cast<const uint16_t *>(x + 208);
}

template <bool F32ACT>
__global__ void __launch_bounds__(256)
q6k_packed_kernel(const uint8_t *__restrict__ w, int64_t rows, int nb, const uint8_t *__restrict__ act, int64_t row_size,
                  float *__restrict__ dst, int64_t dst_stride) {
    extern __shared__ __align__(16) uint8_t s_act[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t *y = act + (int64_t)blockIdx.y * row_size;
    if (F32ACT) {
        const float *xf = reinterpret_cast<const float *>(y);
        for (int sb = warp; sb < nb; sb += kWarps) {
            const float4 f0 = *reinterpret_cast<const float4 *>(xf + sb * kQK + 8 * lane);
            const float4 f1 = *reinterpret_cast<const float4 *>(xf + sb * kQK + 8 * lane + 4);
            const float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
            int q[8];
            const float d = q8k_quantize_lane(v, lane, q);
            // element 8L + l = sub-block j = L / 2 (of sixteen), group g = L % 2: word l / 2, byte 2 (l % 2) + g
            uint8_t *base = s_act + sb * kActSB6 + (lane >> 1) * 16 + (lane & 1);
#pragma unroll
            for (int l = 0; l < 8; l++) base[(l >> 1) * 4 + (l & 1) * 2] = (uint8_t)q[l];
            if (lane == 0) *reinterpret_cast<float *>(s_act + sb * kActSB6 + 256) = d;
        }
    } else {
    for (int i = tid; i < nb * 64; i += 256) { // word (sb, j, p)
        const int sb = i >> 6, j = (i >> 2) & 15, p = i & 3;
        const uint8_t *q8 = y + (int64_t)sb * kQ8K + 4 + 16 * j + 2 * p;
        const uint32_t v = (uint32_t)q8[0] | ((uint32_t)q8[8] << 8) | ((uint32_t)q8[1] << 16) | ((uint32_t)q8[9] << 24);
        *reinterpret_cast<uint32_t *>(s_act + sb * kActSB6 + j * 16 + p * 4) = v;
    }
    for (int sb = tid; sb < nb; sb += 256) *reinterpret_cast<uint32_t *>(s_act + sb * kActSB6 + 256) = ld_u32(y + (int64_t)sb * kQ8K);
    }
    __syncthreads();
    // lane (s, c): super-block s of four, sub-blocks 2c and 2c + 1 (32 elements per lane); the fold runs over the 8-lane group
    const int s = lane >> 3, c = lane & 7;
    const bool b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
    const int64_t pitch = q6k_packed_pitch(nb);
    for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < rows; row += (int64_t)gridDim.x * kWarps) {
        const uint8_t *xrow = w + row * pitch;
        float sums = 0.0f; // lane (0, l): sums[l]
#pragma unroll 1
        for (int sb0 = 0; sb0 < nb; sb0 += 4) {
            const int sb = sb0 + s;
            const bool valid = sb < nb;
            const int sbc = valid ? sb : 0;
            const uint8_t *xw = xrow + ((int64_t)sbc * 16 + 2 * c) * 16;
            const uint4 X0 = __ldg(reinterpret_cast<const uint4 *>(xw)), X1 = __ldg(reinterpret_cast<const uint4 *>(xw + 16));
            const uint32_t sc2 = __ldg(reinterpret_cast<const uint16_t *>(xrow + (int64_t)nb * 256 + sbc * 16 + 2 * c));
            const uint16_t dh = __ldg(reinterpret_cast<const uint16_t *>(xrow + (int64_t)nb * 272 + 2 * sbc));
            const uint8_t *ya = s_act + sbc * kActSB6;
            const uint4 q0 = *reinterpret_cast<const uint4 *>(ya + 32 * c), q1 = *reinterpret_cast<const uint4 *>(ya + 32 * c + 16);
            int v[8];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint4 X = h ? X1 : X0, q = h ? q1 : q0;
                const int scale = (int)(int8_t)(sc2 >> (8 * h));
                const uint32_t xs[4] = {X.x, X.y, X.z, X.w}, qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int p = 0; p < 4; p++) {
                    const int pe = __dp4a((int)(xs[p] & 0x0000ffffu), (int)qs[p], 0);
                    const int po = __dp4a((int)(xs[p] & 0xffff0000u), (int)qs[p], 0);
                    v[2 * p] = h ? v[2 * p] + scale * pe : scale * pe;
                    v[2 * p + 1] = h ? v[2 * p + 1] + scale * po : scale * po;
                }
            }
            int w4[4], w2[2];
#pragma unroll
            for (int k = 0; k < 4; k++) w4[k] = (b2 ? v[4 + k] : v[k]) + __shfl_xor_sync(0xffffffffu, b2 ? v[k] : v[4 + k], 4);
#pragma unroll
            for (int k = 0; k < 2; k++) w2[k] = (b1 ? w4[2 + k] : w4[k]) + __shfl_xor_sync(0xffffffffu, b1 ? w4[k] : w4[2 + k], 2);
            const int aux = (b0 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, b0 ? w2[0] : w2[1], 1); // aux32[l = c]
            const float d = __fmul_rn(half_bits_to_float(dh), *reinterpret_cast<const float *>(ya + 256));
            const float t = valid ? __fmul_rn(d, (float)aux) : 0.0f;
            sums = __fadd_rn(sums, t); // the scalar order: the four super-blocks one after the other (lanes 0..7)
#pragma unroll
            for (int k = 1; k < 4; k++) sums = __fadd_rn(sums, __shfl_sync(0xffffffffu, t, c + 8 * k));
        }
        float sumf = 0.0f;
#pragma unroll
        for (int l = 0; l < 8; l++) sumf = __fadd_rn(sumf, __shfl_sync(0xffffffffu, sums, l));
        if (lane == 0) dst[(int64_t)blockIdx.y * dst_stride + row] = sumf;
    }
}


// ---------------------------------------------------------------- get_rows on K-quant blocks (embedding lookup)
// [UPSTREAM] dequantize_row_q4_K / dequantize_row_q6_K for the rows ids[t], times `scale` (ggml_scale): one thread per
// element, the scalar code's roundings: Q4_K y = (d * sc) * q - (dmin * m); Q6_K y = (d * scale) * (q - 32).
__global__ void __launch_bounds__(256)
get_rows_blocks_kernel(int type, const uint8_t *__restrict__ raw, int64_t n_rows, int64_t nb01, int64_t K,
                       const int32_t *__restrict__ ids, float scale, float *__restrict__ y) {
    const int64_t t = blockIdx.y;
    int64_t row = ids[t];
    row = row < 0 ? 0 : (row >= n_rows ? n_rows - 1 : row);
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K) return;
    const int64_t sb = e >> 8;
    const int i = (int)(e & 255);
    float val;
    if (type == GGB_TYPE_Q4_K) {
        const uint8_t *x = raw + row * nb01 + sb * kQ4K;
        const int j = i >> 5; // sub-block
        uint32_t u0 = ld_u32(x + 4), u1 = ld_u32(x + 8), u2 = ld_u32(x + 12);
        const uint32_t u3 = ((u2 >> 4) & 0x0f0f0f0fu) | (((u1 >> 6) & 0x03030303u) << 4);
        const uint32_t uaux = u1 & 0x3f3f3f3fu;
        u1 = (u2 & 0x0f0f0f0fu) | (((u0 >> 6) & 0x03030303u) << 4);
        u2 = uaux;
        u0 &= 0x3f3f3f3fu;
        const int sc = (int)(((j < 4 ? u0 : u1) >> (8 * (j & 3))) & 0xff), mn = (int)(((j < 4 ? u2 : u3) >> (8 * (j & 3))) & 0xff);
        const uint32_t dd = ld_u32(x);
        const float d1 = __fmul_rn(half_bits_to_float((uint16_t)(dd & 0xffff)), (float)sc);
        const float m1 = __fmul_rn(half_bits_to_float((uint16_t)(dd >> 16)), (float)mn);
        const uint8_t q = x[16 + 32 * (j >> 1) + (i & 31)];
        const int nib = (j & 1) ? (q >> 4) : (q & 0xF);
        val = __fsub_rn(__fmul_rn(d1, (float)nib), m1);
    } else {
        const uint8_t *x = raw + row * nb01 + sb * kQ6K;
        const int c = i >> 7, quarter = (i >> 5) & 3, l = i & 31;
        const uint8_t ql = x[64 * c + 32 * (quarter & 1) + l], qh = x[128 + 32 * c + l];
        const int lo = (quarter & 2) ? (ql >> 4) : (ql & 0xF);
        const int a = (int)(int8_t)(lo | (((qh >> (2 * quarter)) & 3) << 4)) - 32;
        const float d = half_bits_to_float((uint16_t)ld_u16(x + 208));
        val = __fmul_rn(__fmul_rn(d, (float)(int8_t)x[192 + (i >> 4)]), (float)a);
    }
    y[t * K + e] = __fmul_rn(val, scale);
}

} // namespace

extern "C" {

size_t ggb_type_row_bytes(int type, int64_t cols) {
    switch (type) {
    case GGB_TYPE_F32: return (size_t)cols * 4;
    case GGB_TYPE_F16: return (size_t)cols * 2;
    case GGB_TYPE_Q4_0: return (size_t)(cols / 32) * 18;
    case GGB_TYPE_Q8_0: return (size_t)(cols / 32) * 34;
    case GGB_TYPE_Q4_K: return (size_t)(cols / kQK) * kQ4K;
    case GGB_TYPE_Q6_K: return (size_t)(cols / kQK) * kQ6K;
    case GGB_TYPE_Q8_K: return (size_t)(cols / kQK) * kQ8K;
    default: return 0;
    }
}

int ggb_op_quantize_q8_K(const float *x, int64_t cols, int64_t rows, void *out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(x && out, "null pointer");
    GGB_CHECK_ARG(cols > 0 && cols % kQK == 0 && rows > 0, "quantize_q8_K: cols %lld not a multiple of 256", (long long)cols);
    const int64_t n_blocks = rows * (cols / kQK);
    quantize_q8_K_kernel<<<(unsigned)((n_blocks + kWarps - 1) / kWarps), kWarps * 32, 0, ggb_cuda_stream()>>>(
        x, n_blocks, (uint8_t *)out);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_q4k_repack(const void *raw, int64_t rows, int64_t nb01, int64_t K, void *packed) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(raw && packed && rows > 0 && K > 0 && K % kQK == 0 && nb01 >= K / kQK * kQ4K && nb01 % 4 == 0,
                  "q4k_repack: bad arguments");
    const int nb = (int)(K / kQK);
    const int64_t n = rows * nb * 8;
    q4k_repack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ggb_cuda_stream()>>>((const uint8_t *)raw, rows, nb01, nb,
                                                                                 (uint8_t *)packed);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

static int launch_packed(int type, bool f32act, const void *packed, int64_t rows, int64_t K, const void *act, int64_t row_size,
                         int64_t T, float *dst, int64_t dst_stride) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(packed && act && dst && rows > 0 && T > 0 && K > 0 && K % kQK == 0, "mul_mat_packed: bad arguments");
    const int nb = (int)(K / kQK);
    const int64_t need = f32act ? K * 4 : (int64_t)nb * kQ8K;
    GGB_CHECK_ARG(row_size >= need && row_size % (f32act ? 16 : 4) == 0, "mul_mat_packed: activation row pitch %lld", (long long)row_size);
    const bool q4 = type == GGB_TYPE_Q4_K;
    const size_t smem = (size_t)nb * (q4 ? kActSB : kActSB6);
    GGB_CHECK_ARG(smem <= 200 * 1024, "mul_mat_packed: K %lld too long for the staged activation row", (long long)K);
    auto kern = q4 ? (f32act ? q4k_packed_kernel<true> : q4k_packed_kernel<false>)
                   : (f32act ? q6k_packed_kernel<true> : q6k_packed_kernel<false>);
    if (smem > 48 * 1024) GGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t groups = (rows + kWarps - 1) / kWarps;
    const unsigned gx = (unsigned)(groups < 148 * 8 ? groups : 148 * 8);
    for (int64_t t0 = 0; t0 < T; t0 += 65535) {
        const dim3 grid(gx, (unsigned)(T - t0 < 65535 ? T - t0 : 65535));
        kern<<<grid, 256, smem, ggb_cuda_stream()>>>((const uint8_t *)packed, rows, nb, (const uint8_t *)act + t0 * row_size,
                                                     row_size, dst + t0 * dst_stride, dst_stride);
        GGB_LAUNCH_CHECK();
    }
    return GGB_OK;
}

int ggb_op_mul_mat_q4k_packed(const void *packed, int64_t rows, int64_t K, const void *act, int64_t row_size, int64_t T,
                              float *dst, int64_t dst_stride) {
    return launch_packed(GGB_TYPE_Q4_K, false, packed, rows, K, act, row_size, T, dst, dst_stride);
}

int ggb_op_mul_mat_q6k_packed(const void *packed, int64_t rows, int64_t K, const void *act, int64_t row_size, int64_t T,
                              float *dst, int64_t dst_stride) {
    return launch_packed(GGB_TYPE_Q6_K, false, packed, rows, K, act, row_size, T, dst, dst_stride);
}

int ggb_op_mul_mat_kq_packed_f32(int type, const void *packed, int64_t rows, int64_t K, const float *x, int64_t T, float *dst,
                                 int64_t dst_stride) {
    GGB_CHECK_ARG(type == GGB_TYPE_Q4_K || type == GGB_TYPE_Q6_K, "mul_mat_kq_packed_f32: type %d", type);
    return launch_packed(type, true, packed, rows, K, x, K * 4, T, dst, dst_stride);
}

size_t ggb_q6k_packed_bytes(int64_t rows, int64_t K) { return (size_t)rows * (size_t)q6k_packed_pitch((int)(K / kQK)); }

int ggb_op_q6k_repack(const void *raw, int64_t rows, int64_t nb01, int64_t K, void *packed) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(raw && packed && rows > 0 && K > 0 && K % kQK == 0 && nb01 >= K / kQK * kQ6K && nb01 % 2 == 0,
                  "q6k_repack: bad arguments");
    const int nb = (int)(K / kQK);
    const int64_t n = rows * nb * 16;
    q6k_repack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ggb_cuda_stream()>>>((const uint8_t *)raw, rows, nb01, nb,
                                                                                 (uint8_t *)packed);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_get_rows_blocks(int type, const void *raw, int64_t n_rows, int64_t nb01, int64_t K, const int32_t *ids, int64_t T,
                           float scale, float *y) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(raw && ids && y && T > 0 && T <= 65535 && n_rows > 0, "get_rows_blocks: bad arguments");
    GGB_CHECK_ARG((type == GGB_TYPE_Q4_K || type == GGB_TYPE_Q6_K) && K > 0 && K % kQK == 0, "get_rows_blocks: type %d / K %lld", type, (long long)K);
    GGB_CHECK_ARG(nb01 >= (int64_t)ggb_type_row_bytes(type, K) && nb01 % (type == GGB_TYPE_Q4_K ? 4 : 2) == 0, "get_rows_blocks: row pitch %lld", (long long)nb01);
    const dim3 grid((unsigned)((K + 255) / 256), (unsigned)T);
    get_rows_blocks_kernel<<<grid, 256, 0, ggb_cuda_stream()>>>(type, (const uint8_t *)raw, n_rows, nb01, K, ids, scale, y);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

int ggb_op_mul_mat_blocks(int type, const void *w, int64_t rows, int64_t nb01, int64_t K, const void *act,
                          int64_t row_size, int64_t T, float *dst, int64_t dst_stride) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(w && act && dst, "null pointer");
    GGB_CHECK_ARG(rows > 0 && T > 0 && K > 0, "mul_mat_blocks: bad extents");
    const bool kq = type == GGB_TYPE_Q4_K || type == GGB_TYPE_Q6_K;
    if (!kq && type != GGB_TYPE_F16) {
        ggb_set_error("mul_mat_blocks: no kernel for type %d", type);
        return GGB_ERR_UNSUPPORTED;
    }
    GGB_CHECK_ARG(!kq || K % kQK == 0, "mul_mat_blocks: K %lld not a multiple of 256", (long long)K);
    GGB_CHECK_ARG(nb01 >= (int64_t)ggb_type_row_bytes(type, K) && nb01 % (type == GGB_TYPE_Q4_K ? 4 : 2) == 0,
                  "mul_mat_blocks: src0 row pitch %lld", (long long)nb01);
    const int64_t act_row = (int64_t)ggb_type_row_bytes(kq ? GGB_TYPE_Q8_K : GGB_TYPE_F16, K);
    GGB_CHECK_ARG(row_size >= act_row && row_size % (kq ? 4 : 2) == 0, "mul_mat_blocks: activation row pitch %lld",
                  (long long)row_size);
    const uint8_t *wb = (const uint8_t *)w, *ab = (const uint8_t *)act;
    cudaStream_t st = ggb_cuda_stream();
    for (int64_t t0 = 0; t0 < T; t0 += 65535) { // grid.y carries the activation row
        const dim3 grid((unsigned)((rows + kWarps - 1) / kWarps), (unsigned)(T - t0 < 65535 ? T - t0 : 65535));
        const uint8_t *a0 = ab + t0 * row_size;
        float *d0 = dst + t0 * dst_stride;
        if (type == GGB_TYPE_Q4_K)
            raw_matvec_kernel<GGB_TYPE_Q4_K><<<grid, kWarps * 32, 0, st>>>(wb, rows, nb01, K, a0, row_size, d0, dst_stride);
        else if (type == GGB_TYPE_Q6_K)
            raw_matvec_kernel<GGB_TYPE_Q6_K><<<grid, kWarps * 32, 0, st>>>(wb, rows, nb01, K, a0, row_size, d0, dst_stride);
        else
            raw_matvec_kernel<GGB_TYPE_F16><<<grid, kWarps * 32, 0, st>>>(wb, rows, nb01, K, a0, row_size, d0, dst_stride);
        GGB_LAUNCH_CHECK();
    }
    return GGB_OK;
}

} // extern "C"
