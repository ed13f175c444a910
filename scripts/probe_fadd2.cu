// FADD vs FADD2 (add.f32x2, sm_100a) issue/pipe-rate microbenchmark: lane-ops per second of each form, and of a mix with
// shared-memory loads in the shadow of the packed instructions (what the TransE tile loop looks like).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__global__ void __launch_bounds__(256) k_fadd(float *out, int iters, float c) {
    float a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
    float s = 0; for (int i = 0; i < 16; i++) s += a[i];
    if (s == 1234.5f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_fadd2(float *out, int iters, float c) {
    u64 a[16];
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("{.reg .b64 t; mov.b64 t, {%1, %1}; add.f32x2 %0, %0, t;}" : "+l"(a[i]) : "f"(c));
    u64 s = 0; for (int i = 0; i < 16; i++) s += a[i];
    if (s == 12345) out[0] = 1;
}
// sub + add|.| pairs as in the tile loop, packed
__global__ void __launch_bounds__(256) k_l1_2(float *out, int iters, float c) {
    u64 a[16], q[4];
    for (int i = 0; i < 16; i++) a[i] = 0;
    for (int i = 0; i < 4; i++) q[i] = threadIdx.x * 4 + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) {
                u64 d;
                asm volatile("{.reg .b64 t; mov.b64 t, {%2, %2}; sub.f32x2 %0, %1, t;}" : "=l"(d) : "l"(q[i & 3]), "f"(c));
                asm volatile("{.reg .b32 lo, hi; mov.b64 {lo, hi}, %1; abs.f32 lo, lo; abs.f32 hi, hi; mov.b64 %1, {lo, hi}; add.f32x2 %0, %0, %1;}" : "+l"(a[i]), "+l"(d));
            }
    u64 s = 0; for (int i = 0; i < 16; i++) s += a[i];
    if (s == 12345) out[0] = 1;
}
__global__ void __launch_bounds__(256) k_l1_1(float *out, int iters, float c) {
    float a[32], q[4];
    for (int i = 0; i < 32; i++) a[i] = 0;
    for (int i = 0; i < 4; i++) q[i] = threadIdx.x * 4 + i;
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int i = 0; i < 32; i++) {
                float d;
                asm volatile("sub.f32 %0, %1, %2;" : "=f"(d) : "f"(q[i & 3]), "f"(c));
                asm volatile("{abs.f32 %1, %1; add.f32 %0, %0, %1;}" : "+f"(a[i]), "+f"(d));
            }
    float s = 0; for (int i = 0; i < 32; i++) s += a[i];
    if (s == 1234.5f) out[0] = s;
}
template <class K> double run(K k, double ops_per_thread_iter, int iters) {
    float *out; cudaMalloc(&out, 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = 148 * 8; double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); k<<<blocks, 256>>>(out, iters, 1e-7f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double r = (double)blocks * 256 * iters * ops_per_thread_iter / (ms * 1e-3);
        if (rep && r > best) best = r;
    }
    return best;
}
int main() {
    printf("FADD   : %.3e lane-op/s\n", run(k_fadd, 64, 4096));
    printf("FADD2  : %.3e lane-op/s\n", run(k_fadd2, 128, 4096));
    printf("L1 x1  : %.3e lane-op/s (sub + add|.|)\n", run(k_l1_1, 128, 4096));
    printf("L1 x2  : %.3e lane-op/s (sub.f32x2 + add.f32x2|.|)\n", run(k_l1_2, 256, 4096));
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
