// Micro-benchmark: what does a WRITE-dominated stream reach on this GPU, next to a copy?
// The evaluator's traffic is ~99 % writes (cfg4: 7 MB read, 688 MB written per launch), while
// MEASURED_PEAKS.json's HBM figure is a copy (half reads, half writes).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_bw tools/write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_write8(double *out, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = v;
}
__global__ void k_write8_cs(double *out, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) __stcs(out + i, v);
}
__global__ void k_write16_cs(double2 *out, size_t n, double v)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) __stcs(out + i, make_double2(v, v));
}
// the evaluator's pattern: each thread owns one of 64 breakpoints and walks rows with stride 64 doubles
__global__ void k_write_rows(double *out, size_t nprob, int rows, double v)
{
    const int bp = threadIdx.x & 63, pl = threadIdx.x >> 6;
    for (size_t p = (size_t)blockIdx.x * 4 + pl; p < nprob; p += (size_t)gridDim.x * 4) {
        double *o = out + p * rows * 64 + bp;
#pragma unroll 4
        for (int r = 0; r < rows; r++) __stcs(o + (size_t)r * 64, v + r);
    }
}
__global__ void k_copy16(const double2 *in, double2 *out, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = in[i];
}

template <class F> static float timeit(F f, int iters)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main()
{
    const size_t bytes = 65536ull * 11496;  // cfg4's output footprint
    const size_t n = bytes / 8;
    double *a, *b;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
    cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int it = 30;
    float ms;
    ms = timeit([&] { cudaMemsetAsync(a, 0, bytes); }, it);
    printf("cudaMemset            %8.1f us %8.1f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    for (int bps : {2, 4, 8, 16}) {
        ms = timeit([&] { k_write8<<<sms * bps, 256>>>(a, n, 1.0); }, it);
        printf("write8  default %2d/SM %8.1f us %8.1f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&] { k_write8_cs<<<sms * bps, 256>>>(a, n, 1.0); }, it);
        printf("write8  .cs     %2d/SM %8.1f us %8.1f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
        ms = timeit([&] { k_write16_cs<<<sms * bps, 256>>>((double2 *)a, n / 2, 1.0); }, it);
        printf("write16 .cs     %2d/SM %8.1f us %8.1f GB/s\n", bps, ms * 1e3, bytes / ms / 1e6);
    }
    const int rows = 11496 / 8 / 64;  // 22 rows of 64 doubles per problem ~ the same bytes
    for (int bps : {2, 4, 8}) {
        ms = timeit([&] { k_write_rows<<<sms * bps, 256>>>(a, 65536, rows, 1.0); }, it);
        printf("rows of 64 .cs  %2d/SM %8.1f us %8.1f GB/s\n", bps, ms * 1e3, 65536.0 * rows * 512 / ms / 1e6);
    }
    for (int bps : {4, 8, 16}) {
        ms = timeit([&] { k_copy16<<<sms * bps, 256>>>((const double2 *)a, (double2 *)b, n / 2); }, it);
        printf("copy16 (r+w)    %2d/SM %8.1f us %8.1f GB/s (read+write)\n", bps, ms * 1e3, 2.0 * bytes / ms / 1e6);
    }
    ms = timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, it);
    printf("cudaMemcpy D2D        %8.1f us %8.1f GB/s (read+write)\n", ms * 1e3, 2.0 * bytes / ms / 1e6);
    return 0;
}
