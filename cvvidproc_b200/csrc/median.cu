// Temporal median of a uint8 frame stack on sm_100a -- replaces HistogramMedianAlgo<T>
// (/root/reference/Sources/ProcessorAlgos/histogram_median_algo.h:116-193).
//
// The reference builds a 256-bin histogram per element and returns the first bin whose
// cumulative count exceeds N/2, i.e. sorted[N/2].  On B200 a shared-memory histogram costs one
// (atomic) read-modify-write per input byte and cannot keep up with HBM, so this kernel computes
// the same order statistic with an on-chip BIT-SLICED RADIX SELECT:
//
//   tile      = P consecutive elements (P = 128 >> LOG2S) x all N frames, resident on one SM;
//   producer  = one thread streaming the tile through a ring of 4 KB TMA boxes
//               (P bytes x 4096/P frames each, zero-filled out of bounds);
//   transpose = each consumer warp takes one 4 KB stage, every lane reads 32 words (4 elements x
//               32 frame slots) and bit-transposes them in registers (32x32 bit matrix:
//               2 PRMT stages + 3 shift/LOP3 stages) into 4 elements x 8 bit planes, one word =
//               32 frames of one bit of one element; planes go to shared memory (swizzled so both
//               the stores and the select loads are bank-conflict free);
//   select    = 8 passes, MSB first: count = popc(alive & ~plane) summed over the element's frame
//               blocks (4..32 threads per element, warp shuffles), compare with the remaining
//               rank, keep the matching half (alive &= plane or ~plane).
//
// Frame order inside a plane word is irrelevant to an order statistic, and zero-filled padding
// frames are accounted for by raising the wanted rank by the number of pad slots, so no masks
// are needed.  HBM traffic is the algorithmic minimum: every input byte is read exactly once.
#include "context.hpp"

#include <cstdlib>
#include "ptx_helpers.cuh"
#include "median_common.cuh"

namespace cvvp
{
namespace
{
constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;
constexpr int kThreads = kConsumerThreads + 32; // + 1 producer warp
constexpr int kStageBytes = 4096;
constexpr int kStageWords = kStageBytes / 4;
constexpr int kMaxStagesPerTile = 40; // planes of one tile: nst * 4 KB
// An mbarrier parity wait is only sound if the previous fill of that slot has completed, so every ring slot is
// private to one consumer warp (1 or 2 slots per warp).
constexpr int kMinRing = kConsumerWarps;
static_assert(kMaxStagesPerTile == 40, "dispatch_jt covers JT = 1..10");

// LOG2S: log2 of the number of 32-frame sub-blocks one 4 KB stage holds per element.
//   P (elements per tile)   = 128 >> LOG2S
//   frame slots per stage   = 32 << LOG2S
// JT   : stages per select thread = ceil(nst / 4), a compile-time constant so that the per-thread
//        alive[]/plane-word arrays live in registers with no guards.
template <int LOG2S, int JT>
__global__ void __launch_bounds__(kThreads, 1)
    median_bitslice_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t *__restrict__ out, const uint32_t nelem,
                           const uint32_t nframes, const uint32_t nst, const uint32_t nring, const uint32_t ntiles)
{
    constexpr int S = 1 << LOG2S;
    constexpr int P = 128 >> LOG2S;
    constexpr int kSlotsPerStage = 32 * S;
    constexpr int kCloBits = 3 - LOG2S; // low bits of the word-column index kept in the lane id

    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);                           // nring x 4 KB
    constexpr uint32_t nst4 = 4u * JT; // plane rows are padded to a multiple of 4 stages
    uint32_t *planes = reinterpret_cast<uint32_t *>(smem + size_t(nring) * kStageBytes); // [8][nst4][4][32] words
    uint8_t *outstage = smem + size_t(nring + nst4) * kStageBytes;                 // 128 B
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(outstage + 128);             // nring
    uint64_t *empty_bar = full_bar + nring;                                        // nring

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;

    if (tid == 0) {
        prefetch_tmap(&tmap);
        for (uint32_t i = 0; i < nring; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ===== TMA producer: one thread, runs ahead of the consumers by up to nring stages,
        // including across tile boundaries (the next tile streams in during the select phase).
        if (lane == 0) {
            // each consumer warp owns R = nring/16 PRIVATE ring slots (slot = w + 16*(k % R) for its k-th stage):
            // TMA completions are unordered across slots, so a slot shared between warps could alias a parity wait
            const uint32_t R = nring / kConsumerWarps;
            uint32_t seq[kConsumerWarps];
#pragma unroll
            for (int i = 0; i < kConsumerWarps; ++i)
                seq[i] = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int32_t x = int32_t(tile * P);
                for (uint32_t st = 0; st < nst; ++st) {
                    const uint32_t w = st % kConsumerWarps;
                    uint32_t k = 0;
#pragma unroll
                    for (int i = 0; i < kConsumerWarps; ++i) {
                        if (uint32_t(i) == w) {
                            k = seq[i];
                            seq[i] = k + 1;
                        }
                    }
                    const uint32_t slot = w + kConsumerWarps * (k % R);
                    const uint32_t fill = k / R;
                    mbar_wait(&empty_bar[slot], (fill & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&full_bar[slot], kStageBytes);
                    tma_load_2d(ring + size_t(slot) * kStageWords, &tmap, &full_bar[slot], x,
                                int32_t(st * kSlotsPerStage), kL2EvictFirst);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    // transposer coordinates of this lane inside a stage: word column c of the P-byte row,
    // sub-block s (frame slots of the stage with slot % S == s)
    const uint32_t t_c = lane & ((32u >> LOG2S) - 1u);
    const uint32_t t_s = lane >> (5 - LOG2S);
    const uint32_t t_chi = t_c >> kCloBits;
    const uint32_t t_low3 = ((t_c & ((1u << kCloBits) - 1u)) << LOG2S) | t_s;

    // select coordinates: warp = (byte p of the word, c_hi); lane = (g, c_lo, s)
    const uint32_t s_p = warp >> 2;
    const uint32_t s_chi = warp & 3u;
    const uint32_t s_g = lane >> 3;
    const uint32_t s_low3 = lane & 7u;
    const uint32_t s_col = ((s_chi ^ s_g) << 3) | s_low3; // swizzled column for stages st == g (mod 4)
    const uint32_t s_c = (s_chi << kCloBits) | (s_low3 >> LOG2S);
    const uint32_t s_elem = 4u * s_c + s_p; // element index inside the tile
    const bool s_writer = (s_g == 0u) && ((s_low3 & (S - 1u)) == 0u);

    constexpr uint32_t plane_stride = nst4 * 128u; // words per bit plane
    const uint32_t k0 = nframes / 2u + (nst * kSlotsPerStage - nframes); // rank incl. zero pad slots

    const uint32_t R = nring / kConsumerWarps;
    uint32_t kseq = 0; // stages consumed by this warp
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // ---- transpose phase: warp w takes stages w, w+16, ...
        for (uint32_t st = warp; st < nst; st += kConsumerWarps, ++kseq) {
            const uint32_t slot = warp + kConsumerWarps * (kseq % R);
            const uint32_t phase = (kseq / R) & 1u;
            mbar_wait(&full_bar[slot], phase);
            const uint32_t *src = ring + size_t(slot) * kStageWords + lane;
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i)
                r[i] = src[i * 32];
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&empty_bar[slot]);
            transpose32(r);
            // r[8*p + b] = bit plane b of element 4*c+p over this lane's 32 frame slots
            uint32_t *dst = planes + st * 128u + (((t_chi ^ (st & 3u)) << 3) | t_low3);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
#pragma unroll
                for (int p = 0; p < 4; ++p)
                    dst[b * plane_stride + p * 32] = r[8 * p + b];
            }
        }
        named_bar_sync(1, kConsumerThreads);

        // ---- select phase: this thread owns stages st = 4*j + g of element s_elem, sub-block s
        {
            const uint32_t *base = planes + s_g * 128u + s_p * 32u + s_col;
            uint32_t alive[JT];
            uint32_t w[JT];
#pragma unroll
            for (int j = 0; j < JT; ++j)
                alive[j] = (4u * j + s_g) < nst ? 0xFFFFFFFFu : 0u;
            uint32_t k = k0;
            uint32_t med = 0;
#pragma unroll 1
            for (int b = 7; b >= 0; --b) {
                const uint32_t *pb = base + uint32_t(b) * plane_stride;
                uint32_t cnt = 0;
#pragma unroll
                for (int j = 0; j < JT; ++j) {
                    w[j] = pb[j * 512]; // rows >= nst hold garbage; alive[j] == 0 masks them
                    cnt += __popc(alive[j] & ~w[j]);
                }
                if (LOG2S >= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 1);
                if (LOG2S >= 2) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 2);
                if (LOG2S >= 3) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 4);
                cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 8);
                cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 16);
                const bool one = k >= cnt; // fewer than k+1 candidates have a 0 here -> the bit is 1
                if (one) {
                    k -= cnt;
                    med |= 1u << b;
                }
                const uint32_t flip = one ? 0u : 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < JT; ++j)
                    alive[j] &= (w[j] ^ flip);
            }
            if (s_writer)
                outstage[s_elem] = uint8_t(med);
        }
        named_bar_sync(1, kConsumerThreads);

        // ---- coalesced store of the tile's P result bytes
        if (warp == 0) {
            const size_t e0 = size_t(tile) * P;
            if (e0 + P <= nelem && (reinterpret_cast<uintptr_t>(out + e0) & 3u) == 0) {
                if (lane < P / 4)
                    reinterpret_cast<uint32_t *>(out + e0)[lane] = reinterpret_cast<const uint32_t *>(outstage)[lane];
            } else {
                for (uint32_t i = lane; i < P; i += 32)
                    if (e0 + i < nelem)
                        out[e0 + i] = outstage[i];
            }
        }
    }
}

template <int LOG2S, int JT>
int launch_variant(cvvp_ctx *ctx, const CUtensorMap &tmap, uint8_t *d_out, uint32_t nelem, uint32_t nframes,
                   uint32_t nst, cudaStream_t stream)
{
    constexpr int P = 128 >> LOG2S;
    constexpr uint32_t nst4 = 4u * JT;
    const uint32_t ntiles = (nelem + P - 1) / P;
    const size_t fixed = 128 + 1024 /*alignment slack*/;
    const size_t avail_stages = (ctx->smem_optin - fixed) / (kStageBytes + 16);
    if (avail_stages < nst4 + kMinRing)
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "median: tile does not fit shared memory");
    // a multiple of the consumer-warp count: every ring slot is private to one warp
    const uint32_t nring = (avail_stages - nst4) >= 2u * kConsumerWarps ? 2u * kConsumerWarps : uint32_t(kConsumerWarps);
    const size_t smem_bytes = size_t(nring + nst4) * kStageBytes + 128 + size_t(nring) * 16;
    auto kern = median_bitslice_kernel<LOG2S, JT>;
    CVVP_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
    const uint32_t grid = ntiles < uint32_t(ctx->sm_count) ? ntiles : uint32_t(ctx->sm_count);
    kern<<<grid, kThreads, smem_bytes, stream>>>(tmap, d_out, nelem, nframes, nst, nring, ntiles);
    CVVP_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches++;
    return CVVP_OK;
}

template <int LOG2S>
int dispatch_jt(cvvp_ctx *ctx, const CUtensorMap &tmap, uint8_t *d_out, uint32_t nelem, uint32_t nframes, uint32_t nst,
                cudaStream_t stream)
{
    switch ((nst + 3u) / 4u) {
#define CVVP_JT_CASE(J) \
    case J: return launch_variant<LOG2S, J>(ctx, tmap, d_out, nelem, nframes, nst, stream);
        CVVP_JT_CASE(1)
        CVVP_JT_CASE(2)
        CVVP_JT_CASE(3)
        CVVP_JT_CASE(4)
        CVVP_JT_CASE(5)
        CVVP_JT_CASE(6)
        CVVP_JT_CASE(7)
        CVVP_JT_CASE(8)
        CVVP_JT_CASE(9)
        CVVP_JT_CASE(10)
#undef CVVP_JT_CASE
    default: return fail(ctx, CVVP_ERR_UNSUPPORTED, "median: unsupported stage count %u", nst);
    }
}
bool force_single_buffer()
{
    // development switch: CVVP_MEDIAN_KERNEL=single forces the single-buffer kernel for every frame count
    const char *e = getenv("CVVP_MEDIAN_KERNEL");
    return e && e[0] == 's';
}
} // namespace

int median_launch(cvvp_ctx *ctx, const uint8_t *d_frames, long long nframes, size_t nelem, size_t frame_stride,
                  uint8_t *d_out, cudaStream_t stream)
{
    if (!d_frames || !d_out || nframes <= 0 || nelem == 0)
        return fail(ctx, CVVP_ERR_INVALID, "median: null pointer or empty stack");
    if ((reinterpret_cast<uintptr_t>(d_frames) & 15u) || (frame_stride & 15u) || frame_stride < nelem)
        return fail(ctx, CVVP_ERR_INVALID,
                    "median: device stack must be 16-byte aligned with a frame stride that is a multiple of 16 and >= nelem");
    if (nelem >= (1ull << 31) || nframes >= (1ll << 31))
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "median: stack dimensions exceed the tensor-map limits");

    // smallest LOG2S (widest tile) whose planes fit on chip; the pipelined kernel (double-buffered planes,
    // <= 16 stages per tile) is preferred, the single-buffer kernel takes the larger frame counts
    const bool pipelined = nframes <= median_pipe_max_frames() && !force_single_buffer();
    const long long max_stages = pipelined ? 16 : kMaxStagesPerTile;
    int log2s = -1;
    uint32_t nst = 0;
    for (int l = 0; l <= 3; ++l) {
        const long long slots = 32ll << l;
        const long long need = (nframes + slots - 1) / slots;
        if (need <= max_stages) {
            log2s = l;
            nst = uint32_t(need);
            break;
        }
    }
    if (log2s < 0)
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "median: %lld frames exceed the on-chip select capacity (%d)", nframes,
                    kMaxStagesPerTile * 256);

    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {cuuint64_t(nelem), cuuint64_t(nframes)};
    const cuuint64_t gstride[1] = {cuuint64_t(frame_stride)};
    const cuuint32_t box[2] = {cuuint32_t(128 >> log2s), cuuint32_t(32 << log2s)};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult cr = ctx->encode_tiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(d_frames), gdim,
                                          gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS)
        return fail(ctx, CVVP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(cr));

    if (pipelined)
        return median_pipe_launch(ctx, tmap, log2s, d_out, uint32_t(nelem), uint32_t(nframes), nst, stream);
    switch (log2s) {
    case 0: return dispatch_jt<0>(ctx, tmap, d_out, uint32_t(nelem), uint32_t(nframes), nst, stream);
    case 1: return dispatch_jt<1>(ctx, tmap, d_out, uint32_t(nelem), uint32_t(nframes), nst, stream);
    case 2: return dispatch_jt<2>(ctx, tmap, d_out, uint32_t(nelem), uint32_t(nframes), nst, stream);
    default: return dispatch_jt<3>(ctx, tmap, d_out, uint32_t(nelem), uint32_t(nframes), nst, stream);
    }
}
} // namespace cvvp
