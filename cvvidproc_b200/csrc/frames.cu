// Frame source on the device: decoded frames -> the single-plane (or as-is) tokens both operators consume.
//
// Replaces the per-frame host work of CvVidFramesGeneratorAlgo::GetTokenSet
// (/root/reference/Sources/ProcessorTokenHandlers/cv_vid_frames_generator_algo.h:140-156):
//     frame = frame(crop_rectangle)                           :141
//     vid_is_grayscale     -> cv::extractChannel(frame, 0)    :149-151
//     convert_to_grayscale -> cv::cvtColor(COLOR_RGB2GRAY)    :152-154   (on whatever channel order the decoder gave,
//                                                                          i.e. channel 0 takes the "R" weight)
//     else the frame as is                                    :155-156
// The decoded frames (rows x cols x C interleaved bytes, C in 1..4) are uploaded once; one kernel crops and reduces
// them straight into the device frame stack of the median job or the input batch of the highlight job.
//
// RGB2GRAY arithmetic (OpenCV is not vendored in the reference; pinned here against cv2 4.13 on all 2^24 triples,
// tests/test_oracle_frames.py):  y = (c0*9798 + c1*19235 + c2*3735 + 16384) >> 15.
//
// HBM-bound byte work: a CTA takes up to 1024 output elements of four consecutive rows; their source bytes are fetched
// as aligned 128-bit words into shared memory (coalesced whatever the crop offset), every thread reduces four elements
// per row from shared memory and stores one 32-bit word per row.  The grey conversion reads its four pixels as 32-bit
// words (thread stride 3 words for C = 3: conflict-free) and takes two IDP.2A per pixel (gray4); channel extraction and
// the as-is copy read bytes at immediate offsets.  Measured at 1080p: 5.6 TB/s RGB2GRAY, 6.4 TB/s channel 0
// (profiles/r1_frames_probe.txt).
#include "context.hpp"

namespace cvvp
{
namespace
{
constexpr int kTile = 1024;   // output elements per CTA
constexpr int kThreads = 256; // 4 elements per thread and row
constexpr int kRows = 4;      // rows per CTA
constexpr int kTileWords = (kTile * 4 + 32) / 16; // 128-bit words of one staged row segment (4 channels + alignment slack)

struct PrepArgs {
    const uint8_t *src;
    size_t src_stride; // bytes between source frames
    size_t src_bytes;  // bytes readable from src (all frames)
    uint8_t *dst;
    size_t dst_stride;
    uint32_t src_row_bytes; // source row pitch in bytes (src_width * C)
    uint32_t first_byte;    // byte offset of element 0 of output row 0 inside a source frame
    uint32_t row_elems;     // output elements per row
    uint32_t step;          // source bytes per output element (C, or 1 for the as-is copy)
    uint32_t tiles_per_row;
    uint32_t rows;
    uint32_t gray;    // 1: RGB2GRAY over 3 bytes, 0: take byte 0
    uint32_t aligned; // src is 16-byte aligned: 128-bit loads allowed
};

template <bool GRAY>
__device__ __forceinline__ uint32_t reduce_element(const uint8_t *p)
{
    if (GRAY)
        return (uint32_t(p[0]) * 9798u + uint32_t(p[1]) * 19235u + uint32_t(p[2]) * 3735u + 16384u) >> 15;
    return p[0];
}

// RGB2GRAY of four pixels held in STEP aligned words (x[0] byte 0 = first channel of the first pixel) -> four grey bytes.
// IDP.2A: c + A.lo16 * B.byte{0|2} + A.hi16 * B.byte{1|3}; the coefficient pairs put each pixel's three bytes against
// 9798 / 19235 / 3735 wherever they fall in the words.  Exactly the closed form of reduce_element.
template <int STEP>
__device__ __forceinline__ uint32_t gray4(const uint32_t *x)
{
    constexpr uint32_t C0 = 9798u, C1 = 19235u, C2 = 3735u, RND = 16384u;
    uint32_t y0, y1, y2, y3;
    if (STEP == 3) {
        y0 = __dp2a_hi(C2, x[0], __dp2a_lo(C0 | C1 << 16, x[0], RND));             // bytes 0 1 2 of x0
        y1 = __dp2a_lo(C1 | C2 << 16, x[1], __dp2a_hi(C0 << 16, x[0], RND));       // byte 3 of x0, bytes 0 1 of x1
        y2 = __dp2a_lo(C2, x[2], __dp2a_hi(C0 | C1 << 16, x[1], RND));             // bytes 2 3 of x1, byte 0 of x2
        y3 = __dp2a_hi(C1 | C2 << 16, x[2], __dp2a_lo(C0 << 16, x[2], RND));       // bytes 1 2 3 of x2
    } else { // 4 channels: one pixel per word, the fourth byte is ignored
        y0 = __dp2a_hi(C2, x[0], __dp2a_lo(C0 | C1 << 16, x[0], RND));
        y1 = __dp2a_hi(C2, x[1], __dp2a_lo(C0 | C1 << 16, x[1], RND));
        y2 = __dp2a_hi(C2, x[2], __dp2a_lo(C0 | C1 << 16, x[2], RND));
        y3 = __dp2a_hi(C2, x[3], __dp2a_lo(C0 | C1 << 16, x[3], RND));
    }
    return (y0 >> 15) | (y1 >> 15) << 8 | (y2 >> 15) << 16 | (y3 >> 15) << 24;
}

// 16 source bytes at offset `at` (a multiple of 16 relative to a.src)
__device__ __forceinline__ uint4 load_chunk(const PrepArgs &a, size_t at)
{
    if (a.aligned && at + 16 <= a.src_bytes)
        return __ldcs(reinterpret_cast<const uint4 *>(a.src + at)); // read once: streaming
    uint32_t w[4] = {0, 0, 0, 0};
    for (int b = 0; b < 16; ++b)
        if (at + b < a.src_bytes)
            w[b >> 2] |= uint32_t(a.src[at + b]) << (8 * (b & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// A CTA takes the same tile of kRows consecutive rows.  Every thread first issues its 128-bit load of EVERY row and
// only then stores them to shared memory, so kRows independent loads per thread (~12 KB per CTA, ~100 KB per SM) are
// in flight; a load -> store loop per row serialised one DRAM latency per row and left the kernel at 3.4 TB/s.
// STEP (source bytes per output element) and GRAY are template parameters: with them the twelve byte reads of a
// thread's four pixels are immediate offsets from one shared-memory address (ncu: the first version issued 84 % of
// all cycles, most of it address arithmetic with a run-time step).
template <int STEP, bool GRAY>
__global__ void __launch_bounds__(kThreads) frames_prepare_kernel(const PrepArgs a)
{
    __shared__ uint4 tile[kRows][kTileWords];
    const uint32_t rg = blockIdx.x / a.tiles_per_row;
    const uint32_t x0 = (blockIdx.x % a.tiles_per_row) * kTile;
    const uint32_t f = blockIdx.y;
    const uint32_t n = min(uint32_t(kTile), a.row_elems - x0);
    const uint32_t row0 = rg * kRows;
    const uint32_t nrows = min(uint32_t(kRows), a.rows - row0);
    const uint32_t seg_len = n * STEP;

    // source bytes of row r's tile: [seg, seg + seg_len) relative to a.src, staged from the aligned address a0 = seg - off
    uint32_t off[kRows];
    size_t a0[kRows];
    uint4 v[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const size_t seg = size_t(f) * a.src_stride + size_t(a.first_byte) + size_t(row0 + r) * a.src_row_bytes + size_t(x0) * STEP;
        a0[r] = seg & ~size_t(15);
        off[r] = uint32_t(seg - a0[r]);
        if (uint32_t(r) < nrows && threadIdx.x < ((off[r] + seg_len + 15u) >> 4))
            v[r] = load_chunk(a, a0[r] + size_t(threadIdx.x) * 16);
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        if (uint32_t(r) < nrows) {
            const uint32_t nchunks = (off[r] + seg_len + 15u) >> 4;
            if (threadIdx.x < nchunks)
                tile[r][threadIdx.x] = v[r];
            for (uint32_t k = threadIdx.x + kThreads; k < nchunks; k += kThreads) // only 4-channel tiles get here
                tile[r][k] = load_chunk(a, a0[r] + size_t(k) * 16);
        }
    }
    __syncthreads();

    const uint32_t e0 = threadIdx.x * 4;
    if (e0 >= n)
        return;
    uint8_t *out0 = a.dst + size_t(f) * a.dst_stride + size_t(row0) * a.row_elems + x0 + e0;
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        if (uint32_t(r) >= nrows)
            break;
        const uint32_t at = off[r] + e0 * STEP; // byte offset of this thread's first element in the staged segment
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(tile[r]) + at;
        uint8_t *out = out0 + size_t(r) * a.row_elems;
        if (e0 + 4 <= n && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
            uint32_t w;
            if (GRAY) {
                // four pixels = STEP words: 32-bit reads from the word below `at` (thread stride STEP words: conflict-
                // free for STEP = 3), funnel-shifted into place, then two 16-bit x 8-bit dot products per pixel
                uint32_t x[STEP + 1];
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile[r]) + (at >> 2);
#pragma unroll
                for (int i = 0; i <= STEP; ++i)
                    x[i] = wp[i];
                const uint32_t sh = (at & 3u) * 8u;
#pragma unroll
                for (int i = 0; i < STEP; ++i)
                    x[i] = __funnelshift_r(x[i], x[i + 1], sh);
                w = gray4<STEP>(x);
            } else {
                w = uint32_t(bytes[0]) | uint32_t(bytes[STEP]) << 8 | uint32_t(bytes[2 * STEP]) << 16 | uint32_t(bytes[3 * STEP]) << 24;
            }
            __stcs(reinterpret_cast<uint32_t *>(out), w);
        } else {
            for (uint32_t i = 0; i < 4 && e0 + i < n; ++i)
                out[i] = uint8_t(reduce_element<GRAY>(bytes + i * STEP));
        }
    }
}
} // namespace

size_t frames_out_bytes(const cvvp_frame_format &f)
{
    const size_t oc = f.mode == CVVP_FRAMES_AS_IS ? size_t(f.src_channels) : 1;
    return size_t(f.crop_width) * size_t(f.crop_height) * oc;
}

int frames_check_format(cvvp_ctx *ctx, const cvvp_frame_format *f)
{
    if (!f)
        return fail(ctx, CVVP_ERR_INVALID, "frame format is NULL");
    if (f->src_width <= 0 || f->src_height <= 0 || f->src_channels < 1 || f->src_channels > 4)
        return fail(ctx, CVVP_ERR_INVALID, "frame format: source must be rows x cols x 1..4 channels");
    if (f->crop_x < 0 || f->crop_y < 0 || f->crop_width <= 0 || f->crop_height <= 0 ||
        (long long)f->crop_x + f->crop_width > f->src_width || (long long)f->crop_y + f->crop_height > f->src_height)
        return fail(ctx, CVVP_ERR_INVALID, "frame format: crop rectangle outside the frame");
    if (f->mode != CVVP_FRAMES_AS_IS && f->mode != CVVP_FRAMES_CHANNEL0 && f->mode != CVVP_FRAMES_RGB2GRAY)
        return fail(ctx, CVVP_ERR_INVALID, "frame format: unknown mode %d", f->mode);
    if (f->mode == CVVP_FRAMES_RGB2GRAY && f->src_channels < 3)
        return fail(ctx, CVVP_ERR_INVALID, "frame format: RGB2GRAY needs 3 or 4 channels (cv::cvtColor asserts scn == 3 || scn == 4)");
    if (size_t(f->src_width) * size_t(f->src_height) * size_t(f->src_channels) >= (size_t(1) << 31))
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "frame format: source frame of 2 GiB or more");
    return CVVP_OK;
}

// rows [band_row0, band_row0 + band_rows) of every source frame are present at d_src + i*src_stride (band_row0 = 0,
// band_rows = src_height for whole frames; the host-buffer entry points upload only the crop's rows)
int frames_prepare_launch(cvvp_ctx *ctx, const uint8_t *d_src, long long n, size_t src_stride, size_t src_bytes,
                          const cvvp_frame_format &f, int band_row0, uint8_t *d_dst, size_t dst_stride, cudaStream_t stream)
{
    const uint32_t C = uint32_t(f.src_channels);
    const bool as_is = f.mode == CVVP_FRAMES_AS_IS;
    PrepArgs a{};
    a.src = d_src;
    a.src_stride = src_stride;
    a.src_bytes = src_bytes;
    a.dst = d_dst;
    a.dst_stride = dst_stride;
    a.src_row_bytes = uint32_t(f.src_width) * C;
    a.first_byte = (uint32_t(f.crop_y - band_row0) * uint32_t(f.src_width) + uint32_t(f.crop_x)) * C;
    a.row_elems = uint32_t(f.crop_width) * (as_is ? C : 1u);
    a.step = as_is ? 1u : C;
    a.tiles_per_row = (a.row_elems + kTile - 1) / kTile;
    a.rows = uint32_t(f.crop_height);
    a.gray = f.mode == CVVP_FRAMES_RGB2GRAY ? 1u : 0u;
    a.aligned = (reinterpret_cast<uintptr_t>(d_src) & 15) == 0 ? 1u : 0u;
    const unsigned long long ctas = (unsigned long long)a.tiles_per_row * ((a.rows + kRows - 1) / kRows);
    if (ctas > 0x7fffffffull)
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "frames: frame too large for one launch");
    for (long long done = 0; done < n;) {
        const long long chunk = (n - done) < 65535 ? (n - done) : 65535;
        PrepArgs b = a;
        b.src = d_src + size_t(done) * src_stride;
        b.src_bytes = src_bytes - size_t(done) * src_stride;
        b.aligned = (reinterpret_cast<uintptr_t>(b.src) & 15) == 0 ? 1u : 0u;
        b.dst = d_dst + size_t(done) * dst_stride;
        const dim3 grid{unsigned(ctas), unsigned(chunk), 1u};
        if (b.gray && b.step == 3)
            frames_prepare_kernel<3, true><<<grid, kThreads, 0, stream>>>(b);
        else if (b.gray)
            frames_prepare_kernel<4, true><<<grid, kThreads, 0, stream>>>(b);
        else if (b.step == 1)
            frames_prepare_kernel<1, false><<<grid, kThreads, 0, stream>>>(b);
        else if (b.step == 2)
            frames_prepare_kernel<2, false><<<grid, kThreads, 0, stream>>>(b);
        else if (b.step == 3)
            frames_prepare_kernel<3, false><<<grid, kThreads, 0, stream>>>(b);
        else
            frames_prepare_kernel<4, false><<<grid, kThreads, 0, stream>>>(b);
        CVVP_CUDA_OK(ctx, cudaGetLastError());
        ctx->launches++;
        done += chunk;
    }
    return CVVP_OK;
}
} // namespace cvvp
