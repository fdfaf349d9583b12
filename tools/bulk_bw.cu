// Micro-benchmark: what does a cp.async.bulk shared->global store stream reach, as a function of
// row size, destination alignment and the number of lanes issuing?  (K1c/H drains Jacobian rows of
// bpc*8 = 1600..1608 bytes whose start is only 16-byte aligned after the head fix-up.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulk_bw tools/bulk_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_store(void *g, unsigned s, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

// one warp per CTA; `lanes` lanes each issue one row per step; rows of `row_bytes` at pitch `pitch_bytes`
// starting `skew` bytes into the buffer
template <int DEPTH>
__global__ void k_bulk(char *out, size_t rows_total, int row_bytes, size_t pitch_bytes, int skew, int lanes)
{
    extern __shared__ __align__(128) char sm[];
    const int lane = threadIdx.x;
    for (int i = threadIdx.x; i < 32 * 1024 / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = 1.0;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    const unsigned s = (unsigned)__cvta_generic_to_shared(sm) + (unsigned)(lane % 8) * 2048u;
    for (size_t r = (size_t)blockIdx.x * lanes + lane; r < rows_total + lanes; r += (size_t)gridDim.x * lanes) {
        if (lane < lanes && r < rows_total) bulk_store(out + skew + r * pitch_bytes, s, (unsigned)row_bytes);
        bulk_commit();
        bulk_wait_read<DEPTH>();
    }
    bulk_wait_all();
}

__global__ void k_rows_stg(double *out, size_t rows_total, int row_elems, size_t pitch_elems, int skew_elems)
{
    // the old K1c pattern: a CTA of 224 threads writes row after row, lanes = consecutive elements
    for (size_t r = blockIdx.x; r < rows_total; r += gridDim.x) {
        double *o = out + skew_elems + r * pitch_elems;
        for (int i = threadIdx.x; i < row_elems; i += blockDim.x) __stcs(o + i, 1.0);
    }
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
    const size_t bytes = 2600ull << 20;  // ~cfg5 / 4096 problems
    char *a;
    cudaMalloc(&a, bytes + 4096);
    cudaMemset(a, 0, bytes + 4096);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(k_bulk<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int it = 10;
    struct Case { int row, skew; size_t pitch; const char *what; };
    const Case cases[] = {
        {2048, 0, 2048, "2048 B rows, 128 B aligned, contiguous"},
        {1600, 0, 1600, "1600 B rows, contiguous (64 B aligned)"},
        {1600, 16, 1600, "1600 B rows, contiguous, +16 skew"},
        {1600, 0, 3216, "1600 B rows at pitch 3216 (~K1c/H rank 0)"},
        {1600, 16, 3216, "1600 B rows at pitch 3216, +16 skew"},
        {256, 0, 256, "256 B rows, contiguous"},
        {16384, 0, 16384, "16 KB rows, contiguous"},
    };
    for (const Case &c : cases) {
        const size_t rows = bytes / c.pitch;
        for (int lanes : {1, 8, 32}) {
            float ms = timeit([&] { k_bulk<1><<<sms, 32, 40 * 1024>>>(a, rows, c.row, c.pitch, c.skew, lanes); }, it);
            float ms4 = timeit([&] { k_bulk<4><<<sms, 32, 40 * 1024>>>(a, rows, c.row, c.pitch, c.skew, lanes); }, it);
            printf("%-46s lanes %2d: depth1 %8.1f us %7.1f GB/s | depth4 %8.1f us %7.1f GB/s\n", c.what, lanes, ms * 1e3,
                   (double)rows * c.row / ms / 1e6, ms4 * 1e3, (double)rows * c.row / ms4 / 1e6);
        }
    }
    {
        const size_t rows = bytes / 3216;
        for (int bps : {1, 2, 4}) {
            float ms = timeit([&] { k_rows_stg<<<sms * bps, 224>>>((double *)a, rows, 200, 402, 0); }, it);
            printf("STG.cs rows of 200 doubles at pitch 402, %d CTA/SM: %8.1f us %7.1f GB/s\n", bps, ms * 1e3,
                   (double)rows * 1600 / ms / 1e6);
        }
    }
    return 0;
}
