// hpc.cpp — rewired HPC dispatch (ref: src/hpc.cpp:216-277).  The reference splits src0's rows
// between a 4-thread pool and an OpenCL queue and waits for both; here the whole mat-mul goes to the
// B200 through the C-ABI layer and the call returns when dst is written.
#include "hpc.h"

#include <cstdio>
#include <cstdlib>

#include "ggml_b200.h"

extern "C" {

int init_cuda_offload(int device) { return ggb_init(device); }

void mul_mat(int64_t ne01, int64_t ne11, int64_t ne12, int64_t nb01, int64_t ne1, int64_t nb1, int64_t nb2,
             size_t row_size, int64_t shared_edge, struct ggml_tensor *src0, struct ggml_tensor *src1,
             struct ggml_tensor *dst, ggml_vec_dot_t vec_dot, enum ggml_type src0_type, const char *wdata) {
    (void)src1;
    const int rc = ggb_mul_mat(ne01, ne11, ne12, nb01, ne1, nb1, nb2, row_size, shared_edge, src0->data, dst->data,
                               (int)src0_type, wdata);
    if (rc == GGB_ERR_UNSUPPORTED && vec_dot) {
        // A src0 type without a device kernel (the graph's F32 x F32 products, any quantization other than
        // Q4_0 / Q8_0 / Q4_K / Q6_K / F16): the reference computes such rows with the caller's vec_dot in
        // mul_mat_sub (ref: src/hpc.cpp:15-41, its CPU share of every call); same loop, same addressing.
        const int64_t col_num = ne11 * ne12;
        for (int64_t col = 0; col < col_num; col++) {
            float *dst_col = (float *)((char *)dst->data + (col % ne1) * nb1 + (col / ne1) * nb2);
            for (int64_t row = 0; row < ne01; row++)
                vec_dot((int)shared_edge, &dst_col[row], 0, (const char *)src0->data + row * nb01, 0,
                        wdata + col * row_size, 0, 1);
        }
        return;
    }
    if (rc != GGB_OK) {
        fprintf(stderr, "mul_mat: %s\n", ggb_last_error());
        exit(EXIT_FAILURE);
    }
}

} // extern "C"
