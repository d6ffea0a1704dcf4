// kernals_shim.cc — builds the REFERENCE's own scalar dot products as host code, from the reference's file where it
// lies (-DKERNALS_CL="<path>/src/kernals.cl", never copied into this repo): vec_dot_q4_K_q8_K (src/kernals.cl:48-111),
// vec_dot_q6_K_q8_K (:114-156) and vec_dot_f16 (:158-164) are plain C apart from the OpenCL address-space qualifiers.
// Compiled as C++ with -fpermissive (the file converts void* implicitly and sizes struct arrays with a `const int`).
// TEST INFRASTRUCTURE: pins oracle/ggml_oracle.c's restatements to the reference (tests/test_kquants_cpu.py).
//
// NOTE (reference quirk): the reference's vec_dot_q6_K_q8_K unpacks only the first 64 of a block's 256 weights (it is
// ggml's QK_K == 64 branch pasted under QK_K = 256) and then reads the rest of aux8 uninitialised; it is exported
// here for completeness, but the oracle's Q6_K follows ggml's QK_K = 256 layout, checked against gguf-py instead.
#define __global
#define __constant static
#define __kernel static
#define restrict __restrict__
typedef unsigned char uchar;
typedef unsigned short ushort;
typedef unsigned int uint;
typedef unsigned long size_t;
typedef _Float16 half;
#define memset ref_cl_memset
#define memcpy ref_cl_memcpy
extern "C" int printf(const char *, ...);
static inline int get_global_id(int) { return 0; }
#include KERNALS_CL
#undef memset
#undef memcpy

extern "C" {
void ref_vec_dot_q4_K_q8_K(int n, float *s, const void *vx, const void *vy) { vec_dot_q4_K_q8_K(n, s, 0, vx, 0, vy, 0); }
void ref_vec_dot_q6_K_q8_K(int n, float *s, const void *vx, const void *vy) { vec_dot_q6_K_q8_K(n, s, 0, vx, 0, vy, 0); }
void ref_vec_dot_f16(int n, float *s, const void *vx, const void *vy) { vec_dot_f16(n, s, 0, (const char *)vx, 0, (const char *)vy, 0); }
int ref_sizeof_block_q4_K(void) { return (int)sizeof(block_q4_K); }
int ref_sizeof_block_q6_K(void) { return (int)sizeof(block_q6_K); }
int ref_sizeof_block_q8_K(void) { return (int)sizeof(block_q8_K); }
}
