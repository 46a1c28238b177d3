// Dependent DADD chain on the current GPU: cycles per add for one lane, for a full warp, and for W
// warps of one CTA running their own chains at the same time (FP64 pipe occupancy).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(double *out, const double *in, int n, int all_lanes)
{
    if (!all_lanes && (threadIdx.x & 31) != 0) return;
    double s = in[0];
    const double v = in[1];
    long long t0 = clock64();
    for (int i = 0; i < n; i++) s += v;          // dependent chain (v not a compile-time constant)
    long long t1 = clock64();
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) out[1024] = (double)(t1 - t0) / n;
}
int main()
{
    double *d_in, *d_out, h[2] = {1.0, 1e-9}, r;
    cudaMalloc(&d_in, 16); cudaMalloc(&d_out, 1025 * 8);
    cudaMemcpy(d_in, h, 16, cudaMemcpyHostToDevice);
    for (int all = 0; all < 2; all++) {
        chain<<<1, 32>>>(d_out, d_in, 1 << 20, all);
        cudaMemcpy(&r, d_out + 1024, 8, cudaMemcpyDeviceToHost);
        printf("dependent DADD, %s: %.2f cycles per add\n", all ? "32 lanes" : "1 lane", r);
    }
    for (int warps = 2; warps <= 32; warps *= 2) {
        chain<<<1, 32 * warps>>>(d_out, d_in, 1 << 18, 1);
        cudaMemcpy(&r, d_out + 1024, 8, cudaMemcpyDeviceToHost);
        printf("%2d warps on one SM, each its own chain (32 lanes): %.2f cycles per add per warp -> %.1f lanes/clk/SM\n",
               warps, r, 32.0 * warps / r);
    }
    return 0;
}
