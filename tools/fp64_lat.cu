// Dependent-issue latency of FP64 add / fma and of a shared-memory load on this GPU
// (one warp, clock64 around a dependent chain).  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double a, double b, int n)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 1023);
    __syncthreads();
    double x = a;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) x = x + b;            // DADD chain
    long long t1 = clock64();
    double y = a;
    for (int i = 0; i < n; i++) y = fma(y, b, a);      // DFMA chain
    long long t2 = clock64();
    int idx = threadIdx.x;
    for (int i = 0; i < n; i++) idx = (int)sm[idx & 1023];  // LDS.64 + F2I chain
    long long t3 = clock64();
    float f = (float)a;
    for (int i = 0; i < n; i++) f = fmaf(f, (float)b, (float)a);
    long long t4 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
    out[threadIdx.x] = x + y + idx + f;
}
int main()
{
    double *o; long long *c, h[4];
    cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 32);
    for (int warps : {1, 4, 8}) {
        k<<<1, 32 * warps>>>(o, c, 1.0, 1.0000001, 4096);
        cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
        printf("%d warp(s): DADD %.1f  DFMA %.1f  LDS.64+cvt %.1f  FFMA %.1f cycles per dependent op\n", warps,
               h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0);
    }
    return 0;
}
