// Development microbenchmark (not part of the product library): how fast can one B200 stream a frame stack
// through TMA with the median kernel's access pattern, as a function of ring depth and box shape?
//   mode 0: "tile" pattern  -- box = P bytes x (4096/P) frames, frame stride = frame_bytes (DRAM page miss per row)
//   mode 1: "linear" pattern -- box = 4096 contiguous bytes of one frame (sequential)
// Consumers only wait for the stage and release it (one word read per lane), so this is the memory-system ceiling.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cvvidproc_b200/csrc -o gpurun_out/tma_bench tools/tma_pattern_bench.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ptx_helpers.cuh"

using namespace cvvp;

constexpr int kStageBytes = 4096;

__global__ void __launch_bounds__(32 * 17, 1)
    stream_kernel(const __grid_constant__ CUtensorMap tmap, int mode, uint32_t P, uint32_t nframes, uint32_t nelem,
                  uint32_t nring, uint32_t ncons, unsigned long long *sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + size_t(nring) * kStageBytes);
    uint64_t *empty_bar = full_bar + nring;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (uint32_t i = 0; i < nring; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const uint32_t rows_per_box = kStageBytes / P;
    // total stages for this CTA
    uint64_t total;
    uint32_t nst = 0, ntiles = 0;
    if (mode == 0) {
        nst = (nframes + rows_per_box - 1) / rows_per_box;
        ntiles = (nelem + P - 1) / P;
        const uint32_t mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
        total = uint64_t(mine) * nst;
    } else {
        const uint64_t boxes = (uint64_t(nframes) * nelem) / kStageBytes;
        total = blockIdx.x < boxes ? (boxes - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    }
    if (warp == 16) {
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (uint64_t g = 0; g < total; ++g) {
                int32_t x, y;
                if (mode == 0) {
                    const uint32_t ti = uint32_t(g / nst), st = uint32_t(g % nst);
                    x = int32_t((blockIdx.x + ti * gridDim.x) * P);
                    y = int32_t(st * rows_per_box);
                } else {
                    const uint64_t box = blockIdx.x + g * gridDim.x; // consecutive CTAs read consecutive 4 KB
                    x = 0;
                    y = int32_t(box * 16); // linear view: rows of 256 bytes, 16 rows = 4096 contiguous bytes
                }
                mbar_wait(&empty_bar[slot], phase ^ 1u);
                mbar_arrive_expect_tx(&full_bar[slot], kStageBytes);
                tma_load_2d(ring + size_t(slot) * 1024, &tmap, &full_bar[slot], x, y, kL2EvictFirst);
                if (++slot == nring) {
                    slot = 0;
                    phase ^= 1u;
                }
            }
        }
        return;
    }
    if (warp >= ncons)
        return;
    // consumer warp w takes stages g = w, w + ncons, ... ; nring is a multiple of ncons so slots are private
    unsigned long long acc = 0;
    uint32_t k = 0;
    const uint32_t R = nring / ncons;
    for (uint64_t g = warp; g < total; g += ncons, ++k) {
        const uint32_t slot = warp + ncons * (k % R);
        mbar_wait(&full_bar[slot], (k / R) & 1u);
        acc += ring[size_t(slot) * 1024 + lane];
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&empty_bar[slot]);
    }
    if (acc == 0x123456789ull)
        *sink = acc;
}

int main(int argc, char **argv)
{
    const uint32_t W = 1920, H = 1080, N = argc > 1 ? atoi(argv[1]) : 1000;
    const size_t nelem = size_t(W) * H;
    uint8_t *d = nullptr;
    unsigned long long *sink = nullptr;
    cudaMalloc(&d, nelem * N);
    cudaMalloc(&sink, 8);
    cudaMemset(d, 1, nelem * N);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = reinterpret_cast<CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                                const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                                CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                                CUtensorMapFloatOOBfill)>(fn);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    struct Cfg { int mode; uint32_t P, nring, ncons; };
    std::vector<Cfg> cfgs;
    for (uint32_t nring : {8u, 16u, 24u, 32u, 48u})
        cfgs.push_back({0, 128, nring, 8});
    cfgs.push_back({0, 128, 48, 16});
    cfgs.push_back({0, 256, 48, 16});
    cfgs.push_back({0, 64, 48, 16});
    for (uint32_t nring : {8u, 16u, 24u, 48u})
        cfgs.push_back({1, 4096, nring, 8});
    for (const Cfg &c : cfgs) {
        CUtensorMap tmap;
        const cuuint64_t gdim[2] = {nelem, N};
        const cuuint64_t gstride[1] = {nelem};
        cuuint32_t box[2];
        if (c.mode == 0) {
            box[0] = c.P;
            box[1] = kStageBytes / c.P;
        } else {
            box[0] = 256; // 2-D box of 256 x 16 over a [rows of 256 B] view is not contiguous; use a 1-row trick below
            box[1] = 1;
        }
        const cuuint32_t es[2] = {1, 1};
        CUresult cr;
        if (c.mode == 0) {
            cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        } else {
            // linear view: the whole stack as rows of 256 bytes; a 256 x 16 box = 4096 contiguous bytes
            const cuuint64_t ld[2] = {256, nelem * N / 256};
            const cuuint64_t ls[1] = {256};
            const cuuint32_t lb[2] = {256, 16};
            cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, ld, ls, lb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        if (cr != CUDA_SUCCESS) {
            printf("encode failed %d\n", int(cr));
            continue;
        }
        const size_t smem = size_t(c.nring) * kStageBytes + c.nring * 16;
        cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (c.mode == 0)
                stream_kernel<<<148, 32 * 17, smem>>>(tmap, 0, c.P, N, uint32_t(nelem), c.nring, c.ncons, sink);
            else
                stream_kernel<<<148, 32 * 17, smem>>>(tmap, 1, 256, N, uint32_t(nelem), c.nring, c.ncons, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best)
                best = ms;
        }
        cudaError_t err = cudaGetLastError();
        printf("mode %d P %4u ring %2u (%3u KB) cons %2u : %.3f ms  %.0f GB/s  %s\n", c.mode, c.P, c.nring,
               c.nring * 4, c.ncons, best, double(nelem) * N / best / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
    }
    return 0;
}
