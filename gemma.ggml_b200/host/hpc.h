// hpc.h — the reference's offload entry point, rewired to the C-ABI CUDA layer.
// Mirrors ref: src/hpc.h:16-36 (same name, same 15 parameters, C linkage so ggml.c can include it).
#ifndef GGB_HOST_HPC_H
#define GGB_HOST_HPC_H

#ifdef __cplusplus
#include <cstddef>
#include <cstdint>
extern "C" {
#else
#include <stddef.h>
#include <stdint.h>
#endif

#ifdef GGB_HAVE_GGML_H
#include "ggml.h"
#else
/* ggml.h is not vendored by the reference (lib/ is absent): the three things this entry point needs.
 * Building inside the reference tree defines GGB_HAVE_GGML_H and uses the real declarations. */
struct ggml_tensor {
    void *data;
};
typedef void (*ggml_vec_dot_t)(int n, float *s, size_t bs, const void *x, size_t bx, const void *y, size_t by, int nrc);
enum ggml_type { GGML_TYPE_F32 = 0, GGML_TYPE_F16 = 1, GGML_TYPE_Q4_0 = 2, GGML_TYPE_Q8_0 = 8, GGML_TYPE_Q4_K = 12,
                 GGML_TYPE_Q6_K = 14, GGML_TYPE_Q8_K = 15 };
#endif

/* ref: src/hpc.h:22-32.  The device kernel is selected by src0_type (Q4_0, Q8_0, Q4_K, Q6_K, F16); for any other
 * type the rows are computed with the caller's vec_dot, as the reference's mul_mat_sub does (src/hpc.cpp:15-41).
 * As in the reference there is no return value; a device failure prints the library's error text and exits
 * (ref: src/hpc.cpp:163-166, src/opencl.cpp:13-17). */
void mul_mat(int64_t ne01, int64_t ne11, int64_t ne12,
             int64_t nb01,
             int64_t ne1,
             int64_t nb1, int64_t nb2,
             size_t row_size,
             int64_t shared_edge,
             struct ggml_tensor *src0, struct ggml_tensor *src1, struct ggml_tensor *dst,
             ggml_vec_dot_t vec_dot,
             enum ggml_type src0_type,
             const char *wdata);

/* ref: init_opencl() called from app::run (src/app.cpp:34) */
int init_cuda_offload(int device);

#ifdef __cplusplus
}
#endif
#endif
