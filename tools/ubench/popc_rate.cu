// Microbenchmark: cost of POPC relative to LOP3/IADD on sm_100a (lanes per clock per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o popc_rate popc_rate.cu && ./popc_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int POPCS>  // POPC instructions per (LOP3 + IADD) pair: 0, 1, 2
__global__ void __launch_bounds__(1024) k(unsigned *out, unsigned seed, int iters)
{
    unsigned a[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed * 2654435761u + threadIdx.x * 8 + i;
        acc[i] = 0;
    }
    unsigned c = seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            c = c * 1664525u + 1013904223u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                unsigned x = a[i] ^ c;       // LOP3
                if (POPCS >= 1) x = __popc(x);
                if (POPCS >= 2) x = __popc(x ^ a[i]) ; // second POPC (+1 LOP3)
                acc[i] += x;                 // IADD
            }
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        s ^= acc[i];
    if (s == 0x12345678u)
        out[0] = s;
}

int main()
{
    unsigned *d;
    cudaMalloc(&d, 4);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](const char *name, auto kern) {
        kern<<<p.multiProcessorCount, 1024>>>(d, 1, 16);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        kern<<<p.multiProcessorCount, 1024>>>(d, 1, iters);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("%-28s %.3f ms  (%.2f ns per 64-element inner block per thread; 1024 threads/SM)\n", name, ms,
               ms * 1e6 / iters);
        return ms;
    };
    const float t0 = run("lop3+iadd", k<0>);
    const float t1 = run("lop3+popc+iadd", k<1>);
    const float t2 = run("lop3+popc+lop3+popc+iadd", k<2>);
    // lanes of POPC per clock per SM, from the marginal time of the added POPCs (clock: assume 1.9 GHz)
    const double popcs = 1024.0 * 64 * iters;
    printf("marginal POPC rate: %.1f lanes/us/SM (1 popc), %.1f lanes/us/SM (2nd popc+lop3)\n", popcs / ((t1 - t0) * 1e3),
           popcs / ((t2 - t1) * 1e3));
    printf("lop3+iadd pair rate: %.1f pairs/us/SM\n", popcs / (t0 * 1e3));
    printf("error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
