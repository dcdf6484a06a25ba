// Microbenchmark: how many bytes per second ONE SM can pull (1 CTA of 1024 threads, 8 x 128-bit loads in flight per
// thread), from DRAM (buffer >> L2) and from L2 (4 MB buffer re-read), vs all 148 SMs together.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sm_ingest sm_ingest.cu && ./sm_ingest
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) pull(const uint4 *__restrict__ src, size_t n16_per_cta, int reps, unsigned *out)
{
    const uint4 *p = src + size_t(blockIdx.x) * n16_per_cta;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < reps; ++r) {
        for (size_t i = threadIdx.x; i + 7 * 1024 < n16_per_cta; i += 8 * 1024) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                v[u] = __ldcg(p + i + u * 1024);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc.x ^= v[u].x;
                acc.y ^= v[u].y;
                acc.z ^= v[u].z;
                acc.w ^= v[u].w;
            }
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u)
        out[0] = 1;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    const size_t big = size_t(148) * (64u << 20); // 64 MB per CTA
    uint4 *d;
    unsigned *o;
    cudaMalloc(&d, big);
    cudaMalloc(&o, 4);
    cudaMemset(d, 1, big);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](const char *name, int ctas, size_t bytes_per_cta, int reps) {
        pull<<<ctas, 1024>>>(d, bytes_per_cta / 16, 1, o);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        pull<<<ctas, 1024>>>(d, bytes_per_cta / 16, reps, o);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gb = double(bytes_per_cta) * reps * ctas / 1e9;
        printf("%-34s %8.3f ms  %8.1f GB/s total  %7.1f GB/s per CTA\n", name, ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / ctas);
    };
    run("1 CTA, DRAM (64 MB once)", 1, 64u << 20, 1);
    run("1 CTA, L2 (4 MB x 16)", 1, 4u << 20, 16);
    run("8 CTAs, DRAM", 8, 64u << 20, 1);
    run("148 CTAs, DRAM (64 MB each)", nsm, 64u << 20, 1);
    run("148 CTAs, L2 (0.5 MB each x 64)", nsm, 512u << 10, 64);
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
