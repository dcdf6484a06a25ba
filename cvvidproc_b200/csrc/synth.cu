// Deterministic synthetic uint8 frames generated in device memory.
// Integer-only, so the host generator (cvvidproc_b200/synth.py) produces identical bytes.
//
//   mix32 = murmur3 fmix32
//   r(f,y,x) = mix32(seed ^ mix32(f*0x9E3779B1 + (y*W + x)))
//   B(y,x)   = 140 + (x*20)/W - (y*10)/H ;  noise = (r & 7) - 3
//   disk k   : a = mix32(seed*1000003 + k), b = mix32(a), c = mix32(b)
//              cx = (a % W + 2f) % W, cy = b % H, rad = 3 + c % 30, depth = 10 + (c>>8) % 50
//              core (rad > 8): radius rad/3, adds back 10 + (c>>16) % 50
//   frame    = clamp(B + noise - sum(depth inside disks) + sum(core add-back), 0, 255)
#include "context.hpp"

namespace cvvp
{
namespace
{
constexpr int kMaxDisks = 64;

__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

struct Disk {
    int cx, cy, r2, depth, core_r2, core_add;
};

// grid = (ceil(W/1024), nrows, nframes); 256 threads, 4 consecutive pixels each
__global__ void __launch_bounds__(256) synth_kernel(uint8_t *__restrict__ frames, size_t frame_stride, int W, int H,
                                                    int row0, long long first_frame, uint32_t seed, int ndisks)
{
    __shared__ Disk disks[kMaxDisks];
    const long long f = first_frame + blockIdx.z;
    const int y = row0 + int(blockIdx.y);
    if (threadIdx.x < unsigned(ndisks)) {
        const uint32_t k = threadIdx.x;
        const uint32_t a = mix32(seed * 1000003u + k);
        const uint32_t b = mix32(a);
        const uint32_t c = mix32(b);
        Disk d;
        d.cx = int((uint64_t(a % uint32_t(W)) + 2ull * uint64_t(f)) % uint64_t(W));
        d.cy = int(b % uint32_t(H));
        const int rad = 3 + int(c % 30u);
        d.r2 = rad * rad;
        d.depth = 10 + int((c >> 8) % 50u);
        if (rad > 8) {
            const int cr = rad / 3;
            d.core_r2 = cr * cr;
            d.core_add = 10 + int((c >> 16) % 50u);
        } else {
            d.core_r2 = -1;
            d.core_add = 0;
        }
        disks[k] = d;
    }
    __syncthreads();

    const int x0 = (int(blockIdx.x) * 256 + int(threadIdx.x)) * 4;
    if (x0 >= W)
        return;
    uint8_t *dst = frames + size_t(blockIdx.z) * frame_stride + size_t(blockIdx.y) * size_t(W) + size_t(x0);
    const uint32_t fterm = uint32_t(uint64_t(f)) * 0x9E3779B1u;
    int v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int x = x0 + i;
        const uint32_t r = mix32(seed ^ mix32(fterm + uint32_t(y) * uint32_t(W) + uint32_t(x)));
        v[i] = 140 + (x * 20) / W - (y * 10) / H + int(r & 7u) - 3;
    }
    for (int k = 0; k < ndisks; ++k) {
        const Disk d = disks[k];
        const int dy = y - d.cy;
        const int dy2 = dy * dy;
        if (dy2 > d.r2)
            continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int dx = x0 + i - d.cx;
            const int q = dx * dx + dy2;
            if (q <= d.r2)
                v[i] -= d.depth;
            if (q <= d.core_r2)
                v[i] += d.core_add;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (x0 + i < W)
            dst[i] = uint8_t(min(max(v[i], 0), 255));
    }
}
} // namespace

int synth_launch(cvvp_ctx *ctx, uint8_t *d_frames, size_t frame_stride, int width, int height, int row0, int nrows,
                 long long first_frame, long long nframes, uint32_t seed, int ndisks, cudaStream_t stream)
{
    if (!d_frames || width <= 0 || height <= 0 || row0 < 0 || nrows <= 0 || row0 + nrows > height || nframes <= 0 ||
        first_frame < 0)
        return fail(ctx, CVVP_ERR_INVALID, "synth: bad geometry");
    if (ndisks < 0 || ndisks > kMaxDisks)
        return fail(ctx, CVVP_ERR_INVALID, "synth: ndisks must be in [0, %d]", kMaxDisks);
    if (nrows > 65535)
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "synth: more than 65535 rows per launch");
    // the band [row0,row0+nrows) of frame i is stored densely at d_frames + i*frame_stride
    for (long long done = 0; done < nframes;) {
        const long long chunk = (nframes - done) < 32768 ? (nframes - done) : 32768;
        dim3 grid(unsigned((width + 1023) / 1024), unsigned(nrows), unsigned(chunk));
        synth_kernel<<<grid, 256, 0, stream>>>(d_frames + size_t(done) * frame_stride, frame_stride, width, height, row0,
                                               first_frame + done, seed, ndisks);
        CVVP_CUDA_OK(ctx, cudaGetLastError());
        ctx->launches++;
        done += chunk;
    }
    return CVVP_OK;
}
} // namespace cvvp
