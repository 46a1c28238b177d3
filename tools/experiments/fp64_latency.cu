// dependent-chain latency of DADD / DFMA / IMAD / FADD on one warp (cycles per operation)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/experiments/fp64_latency tools/experiments/fp64_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double *out, long long *cyc, double x, int n)
{
    double s = out[0];
    unsigned u = (unsigned)x;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; i++) {
        if (OP == 0) s = __dadd_rn(s, x);
        else if (OP == 1) s = __fma_rn(x, x, s);
        else if (OP == 2) u = u * 3u + 1u;
        else asm volatile("add.f32 %0, %0, %1;" : "+f"(*(float *)&u) : "f"(1.0f));
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[1] = s + u; }
}
int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 16); cudaMalloc(&cyc, 8); cudaMemset(out, 0, 16);
    const int n = 1 << 14;
    const char *names[4] = {"DADD", "DFMA", "IMAD", "FADD"};
    for (int rep = 0; rep < 2; rep++)
        for (int op = 0; op < 4; op++) {
            if (op == 0) chain<0><<<1, 32>>>(out, cyc, 1.5, n);
            if (op == 1) chain<1><<<1, 32>>>(out, cyc, 1.5, n);
            if (op == 2) chain<2><<<1, 32>>>(out, cyc, 1.5, n);
            if (op == 3) chain<3><<<1, 32>>>(out, cyc, 1.5, n);
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("%s dependent chain: %.2f cycles/op\n", names[op], (double)h / n);
        }
    // 4 warps on one SM (one per scheduler) and 8 warps: does DFMA issue contention show?
    return 0;
}
