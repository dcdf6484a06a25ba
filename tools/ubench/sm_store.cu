// Microbenchmark: store bandwidth of ONE SM (1 CTA of 1024 threads) and of all SMs: st.global 128-bit per lane
// (coalesced, default and .cs) vs cp.async.bulk shared -> global (TMA) stores of 1 KB per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sm_store sm_store.cu && ./sm_store
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024) push(uint4 *dst, size_t n16_per_cta, int reps)
{
    __shared__ __align__(128) uint4 stage[2][32 * 32]; // 2 x 16 KB: 512 B per warp and buffer
    uint4 *p = dst + size_t(blockIdx.x) * n16_per_cta;
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 0xFF00FF00u, 0x00FF00FFu);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = 0; r < reps; ++r) {
        if (MODE < 2) {
            for (size_t i = threadIdx.x; i < n16_per_cta; i += 1024) {
                if (MODE == 0)
                    p[i] = v;
                else
                    __stcs(p + i, v);
            }
        } else {
            // each warp: 32 chunks of 16 B = 512 B per bulk store, double buffered
            int k = 0;
            for (size_t i = size_t(warp) * 32; i + 32 <= n16_per_cta; i += 32 * 32, ++k) {
                uint4 *s = &stage[k & 1][warp * 32];
                if (k >= 2)
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                s[lane] = v;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(p + i),
                                 "r"(uint32_t(__cvta_generic_to_shared(s)))
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (lane == 0)
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
        }
    }
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    const size_t big = size_t(nsm) * (32u << 20);
    uint4 *d;
    cudaMalloc(&d, big);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto run = [&](const char *name, auto kern, int ctas, size_t bytes_per_cta, int reps) {
        kern<<<ctas, 1024>>>(d, bytes_per_cta / 16, 1);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        kern<<<ctas, 1024>>>(d, bytes_per_cta / 16, reps);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gb = double(bytes_per_cta) * reps * ctas / 1e9;
        printf("%-44s %8.3f ms  %8.1f GB/s total  %7.1f GB/s per CTA  (%s)\n", name, ms, gb / (ms * 1e-3), gb / (ms * 1e-3) / ctas,
               cudaGetErrorString(cudaGetLastError()));
    };
    run("1 CTA st.global.v4, 32 MB", push<0>, 1, 32u << 20, 1);
    run("1 CTA st.global.cs.v4, 32 MB", push<1>, 1, 32u << 20, 1);
    run("1 CTA cp.async.bulk 512 B/warp, 32 MB", push<2>, 1, 32u << 20, 1);
    run("1 CTA st.global.v4, 2 MB x 16 (L2)", push<0>, 1, 2u << 20, 16);
    run("1 CTA cp.async.bulk, 2 MB x 16 (L2)", push<2>, 1, 2u << 20, 16);
    run("148 CTAs st.global.v4, 32 MB each", push<0>, nsm, 32u << 20, 1);
    run("148 CTAs cp.async.bulk, 32 MB each", push<2>, nsm, 32u << 20, 1);
    return 0;
}
