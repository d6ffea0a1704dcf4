// micro-benchmark: cost of a grid barrier among 148 persistent CTAs (512 threads each) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_bar tools/ubench_bar.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// V0: one contended counter, thread 0 polls
__device__ __forceinline__ void bar_counter(unsigned long long *cnt, unsigned long long target) {
    cbar();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(cnt), "l"(1ull) : "memory");
        unsigned long long v;
        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(cnt) : "memory"); } while (v < target);
    }
    cbar();
}
// V1: flag per CTA (stride words), threads 0..G-1 poll one flag each
template <int STRIDE>
__device__ __forceinline__ void bar_flags(uint32_t *flags, uint32_t epoch, int G) {
    cbar();
    const int tid = threadIdx.x;
    if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * STRIDE), "r"(epoch) : "memory");
    if (tid < G) {
        uint32_t v;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + tid * STRIDE) : "memory"); } while ((int)(v - epoch) < 0);
    }
    cbar();
}
// V2: flag per CTA, ONE warp polls (lane covers flags lane, lane+32, ...)
__device__ __forceinline__ void bar_flags_1warp(uint32_t *flags, uint32_t epoch, int G) {
    cbar();
    const int tid = threadIdx.x;
    if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
    if (tid < 32) {
        for (int i = tid; i < G; i += 32) {
            uint32_t v;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory"); } while ((int)(v - epoch) < 0);
        }
    }
    cbar();
}
// V3: two-level: CTA arrives on one of NG group counters; thread 0 polls a global "generation" word that the last
// arriver of the last group bumps -> few contended atomics per address
template <int NG>
__device__ __forceinline__ void bar_tree(unsigned int *grp, unsigned int *root, unsigned int *gen, uint32_t epoch, int G) {
    cbar();
    if (threadIdx.x == 0) {
        const int g = blockIdx.x % NG;
        const int members = (G - g + NG - 1) / NG;
        unsigned int old;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(old) : "l"(grp + 32 * g) : "memory");
        if (old + 1 == members * epoch) {
            unsigned int o2;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(o2) : "l"(root) : "memory");
            if (o2 + 1 == NG * epoch) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(gen), "r"(epoch) : "memory");
        }
        uint32_t v;
        do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gen) : "memory"); } while ((int)(v - epoch) < 0);
    }
    cbar();
}
// V4: counter with relaxed red + explicit fence (what does the fence cost?)
__device__ __forceinline__ void bar_counter_nofence(unsigned long long *cnt, unsigned long long target) {
    cbar();
    if (threadIdx.x == 0) {
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(cnt), "l"(1ull) : "memory");
        unsigned long long v;
        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(cnt) : "memory"); } while (v < target);
    }
    cbar();
}

template <int V>
__global__ void __launch_bounds__(544, 1) k(unsigned long long *cnt, uint32_t *flags, unsigned int *tree, int iters, long long *cyc, float *sink) {
    const int G = gridDim.x;
    if (threadIdx.x >= 512) return;
    long long t0 = clock64();
    for (int it = 1; it <= iters; it++) {
        sink[blockIdx.x * 512 + threadIdx.x] = (float)it; // a global write per thread per phase, like a real epilogue
        if (V == 0) bar_counter(cnt, (unsigned long long)G * it);
        if (V == 1) bar_flags<1>(flags, it, G);
        if (V == 2) bar_flags_1warp(flags, it, G);
        if (V == 3) bar_tree<8>(tree, tree + 32 * 16, tree + 32 * 17, it, G);
        if (V == 4) bar_counter_nofence(cnt, (unsigned long long)G * it);
        if (V == 5) bar_flags<32>(flags, it, G);
        if (V == 6) bar_tree<16>(tree, tree + 32 * 16, tree + 32 * 17, it, G);
        if (V == 7) bar_tree<4>(tree, tree + 32 * 16, tree + 32 * 17, it, G);
    }
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

template <int V> void run(const char *name) {
    unsigned long long *cnt; uint32_t *flags; unsigned int *tree; long long *cyc; float *sink;
    CK(cudaMalloc(&cnt, 64)); CK(cudaMalloc(&flags, 148 * 128 + 4096)); CK(cudaMalloc(&tree, 32 * 18 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaMalloc(&sink, 148 * 512 * 4));
    CK(cudaMemset(cnt, 0, 64)); CK(cudaMemset(flags, 0, 148 * 128 + 4096)); CK(cudaMemset(tree, 0, 32 * 18 * 4));
    const int iters = 2000;
    void *args[] = {&cnt, &flags, &tree, (void *)&iters, &cyc, &sink};
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    CK(cudaLaunchCooperativeKernel((const void *)k<V>, dim3(148), dim3(544), args, 0, 0));
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-40s %.3f us per barrier\n", name, ms * 1e3 / iters);
    cudaFree(cnt); cudaFree(flags); cudaFree(tree); cudaFree(cyc); cudaFree(sink);
}

int main() {
    run<0>("counter red.release + 1 poller");
    run<4>("counter red.relaxed (no fence) + 1 poller");
    run<1>("flags packed, 148 pollers");
    run<5>("flags 128B apart, 148 pollers");
    run<2>("flags packed, 1 warp polls");
    run<7>("tree 4 groups");
    run<3>("tree 8 groups");
    run<6>("tree 16 groups");
    return 0;
}
