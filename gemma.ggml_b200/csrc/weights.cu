// weights.cu — resident quantised matrices: one-time upload + repack of ggml AoS blocks into the
// device SoA layout, the on-device synthetic generator/quantiser, row slices for TP shards.
// The reference re-wraps host weights per mul_mat call (ref: src/hpc.cpp:168-170); on B200 they
// are uploaded once and stay in HBM.
#include <vector>

#include "common.cuh"

// ggml AoS block -> planes.  One thread per block (one-time cost; byte loads because 18/34-byte
// blocks are not 4-byte aligned).
__global__ void repack_kernel(const uint8_t *__restrict__ src, int64_t nb01, int type, int64_t rows, int64_t nb,
                              uint8_t *__restrict__ qs, __half *__restrict__ d) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nb) return;
    const int64_t r = idx / nb, b = idx % nb;
    const int pb = type == GGB_TYPE_Q4_0 ? 16 : 32;
    const uint8_t *blk = src + r * nb01 + b * (pb + 2);
    const uint16_t dbits = (uint16_t)blk[0] | ((uint16_t)blk[1] << 8);
    d[idx] = __ushort_as_half(dbits);
    uint8_t *out = qs + idx * pb;
    for (int i = 0; i < pb; i++) out[i] = blk[2 + i];
}

__global__ void unpack_kernel(const uint8_t *__restrict__ qs, const __half *__restrict__ d, int type, int64_t rows,
                              int64_t nb, uint8_t *__restrict__ dst) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * nb) return;
    const int pb = type == GGB_TYPE_Q4_0 ? 16 : 32;
    uint8_t *blk = dst + idx * (pb + 2);
    const uint16_t dbits = __half_as_ushort(d[idx]);
    blk[0] = (uint8_t)(dbits & 0xff);
    blk[1] = (uint8_t)(dbits >> 8);
    const uint8_t *in = qs + idx * pb;
    for (int i = 0; i < pb; i++) blk[2 + i] = in[i];
}

// Synthetic matrix: one warp per 32-element block, lane = element.
// Q4_0 follows [UPSTREAM] quantize_row_q4_0_reference, Q8_0 quantize_row_q8_0_reference.
__global__ void synth_kernel(int type, int64_t rows, int64_t nb, uint64_t key, float unit, uint8_t *__restrict__ qs,
                             __half *__restrict__ d) {
    const int lane = threadIdx.x & 31;
    const int64_t blk = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (blk >= rows * nb) return;
    const float v = ggb_synth_value(key, (uint64_t)(blk * 32 + lane), unit);
    if (type == GGB_TYPE_Q4_0) {
        // first element with the largest |v| (strict '<' scan in the reference)
        float amax = fabsf(v);
        int who = lane;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const float oa = __shfl_xor_sync(0xffffffffu, amax, off);
            const int ow = __shfl_xor_sync(0xffffffffu, who, off);
            if (oa > amax || (oa == amax && ow < who)) {
                amax = oa;
                who = ow;
            }
        }
        const float mx = __shfl_sync(0xffffffffu, v, who);
        const float dd = __fdiv_rn(mx, -8.0f);
        const float id = dd != 0.0f ? __fdiv_rn(1.0f, dd) : 0.0f;
        const float x0 = __fmul_rn(v, id);
        int q = (int)(int8_t)(__fadd_rn(x0, 8.5f));
        q = q < 15 ? q : 15;
        const int qhi = __shfl_down_sync(0xffffffffu, q, 16);
        if (lane < 16) qs[blk * 16 + lane] = (uint8_t)(q | (qhi << 4));
        if (lane == 0) d[blk] = __float2half_rn(dd);
    } else {
        float dd;
        const int q = quantize_q8_lane(v, dd);
        ((int8_t *)qs)[blk * 32 + lane] = (int8_t)q;
        if (lane == 0) d[blk] = __float2half_rn(dd);
    }
}

__global__ void vector_synth_kernel(float *dst, int64_t n, uint64_t key, float unit, float bias) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __fadd_rn(bias, ggb_synth_value(key, (uint64_t)i, unit));
}

static int alloc_weight(int type, int64_t rows, int64_t cols, ggb_weight **out) {
    GGB_CHECK_ARG(type == GGB_TYPE_Q4_0 || type == GGB_TYPE_Q8_0, "weight type %d has no kernel", type);
    GGB_CHECK_ARG(rows > 0 && cols > 0 && cols % GGB_QK == 0, "bad weight shape %lld x %lld", (long long)rows,
                  (long long)cols);
    ggb_weight *w = new ggb_weight();
    w->type = type;
    w->rows = rows;
    w->cols = cols;
    w->nb = cols / GGB_QK;
    const size_t pbytes = (size_t)rows * w->nb * ggb_payload_bytes(type);
    const size_t dbytes = ((size_t)rows * w->nb * 2 + 255) & ~(size_t)255;
    w->bytes = pbytes + dbytes + 256;
    cudaError_t e = cudaMalloc(&w->base, w->bytes);
    if (e != cudaSuccess) {
        ggb_set_error("cudaMalloc(%zu) for weight -> %s", w->bytes, cudaGetErrorString(e));
        delete w;
        return GGB_ERR_CUDA;
    }
    w->qs = (uint8_t *)w->base;
    w->d = (__half *)((uint8_t *)w->base + ((pbytes + 255) & ~(size_t)255));
    *out = w;
    return GGB_OK;
}

extern "C" {

int ggb_weight_upload(int type, int64_t rows, int64_t cols, const void *host_blocks, int64_t nb01,
                      ggb_weight **out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(host_blocks && out, "null pointer");
    ggb_weight *w = nullptr;
    int rc = alloc_weight(type, rows, cols, &w);
    if (rc) return rc;
    const int64_t min_stride = w->nb * ggb_block_bytes(type);
    if (nb01 <= 0) nb01 = min_stride;
    if (nb01 < min_stride) {
        ggb_weight_free(w);
        ggb_set_error("nb01 %lld smaller than a row (%lld)", (long long)nb01, (long long)min_stride);
        return GGB_ERR_ARG;
    }
    cudaStream_t st = ggb_cuda_stream();
    void *staging = nullptr;
    const size_t raw = (size_t)(rows - 1) * nb01 + min_stride;
    // (every failure below releases the weight and the staging buffer)
    cudaError_t e = cudaMalloc(&staging, raw);
    if (e == cudaSuccess) e = cudaMemcpyAsync(staging, host_blocks, raw, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        const int64_t total = rows * w->nb;
        repack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint8_t *)staging, nb01, type, rows, w->nb,
                                                                      w->qs, w->d);
        e = cudaGetLastError();
        ggb_count_launch();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(staging);
    if (e != cudaSuccess) {
        ggb_weight_free(w);
        ggb_set_error("ggb_weight_upload -> %s", cudaGetErrorString(e));
        return GGB_ERR_CUDA;
    }
    *out = w;
    return GGB_OK;
}

int ggb_weight_synth(int type, int64_t rows, int64_t cols, uint64_t seed, uint64_t tensor_id, float sigma,
                     ggb_weight **out) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(out, "null pointer");
    ggb_weight *w = nullptr;
    int rc = alloc_weight(type, rows, cols, &w);
    if (rc) return rc;
    const int64_t threads = rows * w->nb * 32;
    synth_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, ggb_cuda_stream()>>>(
        type, rows, w->nb, ggb_synth_key(seed, tensor_id), ggb_synth_unit(sigma), w->qs, w->d);
    GGB_LAUNCH_CHECK();
    *out = w;
    return GGB_OK;
}

int ggb_weight_slice(const ggb_weight *w, int64_t row0, int64_t nrows, ggb_weight **out) {
    GGB_CHECK_ARG(w && out && row0 >= 0 && nrows > 0 && row0 + nrows <= w->rows, "bad slice [%lld,+%lld)",
                  (long long)row0, (long long)nrows);
    ggb_weight *s = new ggb_weight(*w);
    s->rows = nrows;
    s->qs = w->qs + (size_t)row0 * w->nb * ggb_payload_bytes(w->type);
    s->d = w->d + (size_t)row0 * w->nb;
    s->base = nullptr;
    s->bytes = 0;
    *out = s;
    return GGB_OK;
}

int ggb_weight_download(const ggb_weight *w, void *host_blocks) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(w && host_blocks, "null pointer");
    const size_t raw = (size_t)w->rows * w->nb * ggb_block_bytes(w->type);
    void *staging = nullptr;
    GGB_CUDA(cudaMalloc(&staging, raw));
    const int64_t total = w->rows * w->nb;
    unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ggb_cuda_stream()>>>(w->qs, w->d, w->type, w->rows, w->nb,
                                                                                (uint8_t *)staging);
    GGB_LAUNCH_CHECK();
    GGB_CUDA(cudaMemcpyAsync(host_blocks, staging, raw, cudaMemcpyDeviceToHost, ggb_cuda_stream()));
    GGB_CUDA(cudaStreamSynchronize(ggb_cuda_stream()));
    GGB_CUDA(cudaFree(staging));
    return GGB_OK;
}

int ggb_weight_free(ggb_weight *w) {
    if (!w) return GGB_OK;
    if (w->base) {
        cudaStreamSynchronize(ggb_cuda_stream());
        cudaFree(w->base);
    }
    delete w;
    return GGB_OK;
}

int64_t ggb_weight_rows(const ggb_weight *w) { return w ? w->rows : 0; }
int64_t ggb_weight_cols(const ggb_weight *w) { return w ? w->cols : 0; }
int ggb_weight_type(const ggb_weight *w) { return w ? w->type : -1; }
size_t ggb_weight_algorithmic_bytes(const ggb_weight *w) {
    return w ? (size_t)w->rows * w->nb * ggb_block_bytes(w->type) : 0;
}

int ggb_vector_synth(float *dst_dev, int64_t n, uint64_t seed, uint64_t tensor_id, float sigma, float bias) {
    GGB_REQUIRE_INIT();
    GGB_CHECK_ARG(dst_dev && n > 0, "bad vector");
    vector_synth_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ggb_cuda_stream()>>>(
        dst_dev, n, ggb_synth_key(seed, tensor_id), ggb_synth_unit(sigma), bias);
    GGB_LAUNCH_CHECK();
    return GGB_OK;
}

} // extern "C"
