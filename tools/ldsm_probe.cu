// Probe: fragment layout and throughput of ldmatrix.m16n16.trans.b8 on sm_100a.  Build and run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/ldsm_probe tools/ldsm_probe.cu && /tmp/ldsm_probe
// (result: profiles/r2_ldsm_probe.txt)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
    __shared__ __align__(128) uint8_t tile[2 * 16 * 16];
    for (int i = threadIdx.x; i < 512; i += 32) tile[i] = (uint8_t)i;  // byte value = row*16 + col (second tile same)
    __syncwarp();
    uint32_t addr = (uint32_t)__cvta_generic_to_shared(tile + (threadIdx.x & 15) * 16);
    uint32_t r0, r1;
    asm volatile("ldmatrix.sync.aligned.m16n16.x1.trans.shared.b8 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
    out[threadIdx.x * 2] = r0;
    out[threadIdx.x * 2 + 1] = r1;
    // x2: lanes 0-15 address tile 0 rows, lanes 16-31 tile 1 rows; mark tile 1 by value + 0 (same) -> use distinct rows order
    uint32_t addr2 = (uint32_t)__cvta_generic_to_shared(tile + (threadIdx.x >> 4) * 256 + (15 - (threadIdx.x & 15)) * 16);
    uint32_t q0, q1, q2, q3;
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                 : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(addr2));
    out[64 + threadIdx.x * 4 + 0] = q0; out[64 + threadIdx.x * 4 + 1] = q1;
    out[64 + threadIdx.x * 4 + 2] = q2; out[64 + threadIdx.x * 4 + 3] = q3;
}
template <int KIND>
__global__ void __launch_bounds__(512) thr(uint32_t* out, int iters, long long* clk) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = i * 2654435761u;
    __syncthreads();
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t base = (uint32_t)__cvta_generic_to_shared(sm + warp * 4096);
    uint32_t acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (KIND != 1) {   // 8 x ldmatrix.x2 = 4096 B per warp out of a 32-row x 128-B box
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int r = lane & 15, m = r >> 2, i = r & 3, p = lane >> 4;
                int frame = KIND == 0 ? 4 * k + i : 8 * (k >> 1) + 4 * ((k + m) & 1) + i;
                int chunk = 4 * p + m;
                if (KIND == 2) chunk ^= frame & 7;
                uint32_t a = base + frame * 128 + chunk * 16;
                uint32_t q0, q1, q2, q3;
                asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0, %1, %2, %3}, [%4];"
                             : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(a));
                acc ^= q0 ^ q1 ^ q2 ^ q3;
            }
        } else {           // 32 x LDS.32 conflict-free = 4096 B per warp
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                uint32_t v;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + k * 128 + lane * 4));
                acc ^= v;
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[KIND] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    uint32_t* d; cudaMalloc(&d, 1 << 22);
    long long* clk; cudaMalloc(&clk, 32);
    probe<<<1, 32>>>(d);
    uint32_t h[192]; cudaError_t e = cudaMemcpy(h, d, 768, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(e));
    for (int t = 0; t < 32; ++t) {
        printf("lane %2d: x1", t);
        for (int r = 0; r < 2; ++r) for (int b = 0; b < 4; ++b) { int v = (h[t*2+r] >> (8*b)) & 255; printf(" (%2d,%2d)", v >> 4, v & 15); }
        printf(" | x2");
        for (int r = 0; r < 4; ++r) for (int b = 0; b < 4; ++b) { int v = (h[64+t*4+r] >> (8*b)) & 255; printf(" (%2d,%2d)", v >> 4, v & 15); }
        printf("\n");
    }
    int iters = 2000;
    cudaFuncSetAttribute(thr<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(thr<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(thr<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rep = 0; rep < 2; ++rep) {
        thr<0><<<148, 512, 65536>>>(d, iters, clk);
        thr<1><<<148, 512, 65536>>>(d, iters, clk);
        thr<2><<<148, 512, 65536>>>(d, iters, clk);
    }
    long long hc[3]; e = cudaMemcpy(hc, clk, 24, cudaMemcpyDeviceToHost);
    printf("err=%s\n", cudaGetErrorString(e));
    // 16 warps x 4096 B per iteration per SM
    for (int k = 0; k < 3; ++k)
        printf("%s: %lld clk for %d iters -> %.1f B/clk/SM\n", k == 1 ? "LDS.32 x32" : k ? "LDSM.8.MT1616 x2 x8 swizzled rows" : "LDSM.8.MT1616 x2 x8 plain rows", hc[k], iters,
               16.0 * 4096 * iters / hc[k]);
    return 0;
}
