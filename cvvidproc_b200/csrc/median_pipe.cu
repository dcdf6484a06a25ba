// Pipelined temporal-median kernel (the fast path for N <= 4096 frames) -- same bit-sliced radix select as
// median.cu, restructured so that HBM streaming, bit transposition and rank selection all run concurrently on
// every SM instead of in alternating phases:
//
//   warp  0..7   SELECT     (256 threads: 4..32 threads per element) -- tile t   , plane buffer t & 1
//   warp  8..19  TRANSPOSE  (12 warps, one 4 KB stage at a time)     -- tile t+1 , plane buffer (t+1) & 1
//   warp  20     PRODUCER   (one thread issuing TMA boxes, up to 24 stages ahead, across tiles)
//
// Shared memory: 24-stage TMA ring (96 KB) + 2 plane buffers of up to 16 stages (2 x 64 KB).
//
// Synchronisation is all mbarrier based and every parity wait is provably at most one phase away:
//   * each transposer warp owns two PRIVATE ring slots (slot = warp + 12*(k&1) for its k-th stage), so the fill it
//     waits for is always the direct successor of a fill it consumed itself (TMA completions are unordered across
//     slots, which makes shared slots unsafe);
//   * every transposer warp and every select warp walks through EVERY tile of its CTA in order (warps without a
//     stage in a tile still do the buffer handshake), so planes_full/planes_empty waits are one phase away by
//     induction.
//
// Plane layout per buffer: [stage][byte p][nibble bh][column phi][4 words = bits 4bh..4bh+3], so a transposer lane
// stores uint4 and a select thread loads the four planes of a nibble with one LDS.128.  phi = (transposer lane id)
// ^ (stage & 1): the select thread that owns stages of parity g for logical lane L reads column L ^ g, which equals
// (its own lane id) ^ (warp & 1).  Both patterns touch 8 distinct 16-byte columns per quarter-warp: conflict free.
#include "context.hpp"
#include "median_common.cuh"
#include "ptx_helpers.cuh"

namespace cvvp
{
namespace
{
constexpr int kSelWarps = 8;
constexpr int kTrWarps = 12;
constexpr int kSelThreads = kSelWarps * 32;
constexpr int kPipeThreads = (kSelWarps + kTrWarps + 1) * 32;
constexpr int kStageBytes = 4096;
constexpr int kStageWords = kStageBytes / 4;
constexpr int kRing = 2 * kTrWarps; // two private slots per transposer warp
constexpr int kMaxStages = 16;      // stages per tile (per plane buffer)

// LOG2S: log2(32-frame sub-blocks per stage per element); P = 128 >> LOG2S elements per tile, 32 << LOG2S frame
//        slots per stage.   JT: stages per select thread = ceil(nst / 2) (compile time).
template <int LOG2S, int JT>
__global__ void __launch_bounds__(kPipeThreads, 1)
    median_pipe_kernel(const __grid_constant__ CUtensorMap tmap, uint8_t *__restrict__ out, const uint32_t nelem,
                       const uint32_t nframes, const uint32_t nst, const uint32_t ntiles, const uint64_t l2_policy)
{
    constexpr int S = 1 << LOG2S;
    constexpr int P = 128 >> LOG2S;
    constexpr int kSlotsPerStage = 32 * S;
    constexpr int kColBits = 5 - LOG2S;             // word-column bits of a transposer lane id (the rest: sub-block)
    constexpr uint32_t kBufWords = 2u * JT * 1024u; // one plane buffer: 2*JT stages x 4 KB

    extern __shared__ __align__(1024) uint8_t smem[];
    uint32_t *ring = reinterpret_cast<uint32_t *>(smem);                         // kRing x 4 KB
    uint32_t *planes = reinterpret_cast<uint32_t *>(smem + kRing * kStageBytes); // 2 buffers
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kRing * kStageBytes + 2u * kBufWords * 4u);
    uint64_t *ring_full = bars;
    uint64_t *ring_empty = bars + kRing;
    uint64_t *planes_full = bars + 2 * kRing;
    uint64_t *planes_empty = planes_full + 2;

    const uint32_t tid = threadIdx.x;
    const uint32_t warp = tid >> 5;
    const uint32_t lane = tid & 31;

    if (tid == 0) {
        prefetch_tmap(&tmap);
        for (int i = 0; i < kRing; ++i) {
            mbar_init(&ring_full[i], 1);
            mbar_init(&ring_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&planes_full[i], kTrWarps);
            mbar_init(&planes_empty[i], kSelWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    // tiles of this CTA: adjacent tile PAIRS (2m, 2m+1), so that for P < 128 both halves of a 128-byte line are
    // requested by the same SM back to back (the second one hits L2)
    const uint32_t npairs = (ntiles + 1) / 2;
    const uint32_t my_pairs = blockIdx.x < npairs ? (npairs - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t my_tiles = 2 * my_pairs; // the odd tail tile (>= ntiles) is skipped below
    auto tile_of = [&](uint32_t it) { return 2u * (blockIdx.x + (it >> 1) * gridDim.x) + (it & 1u); };

    if (warp == kSelWarps + kTrWarps) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t seq[kTrWarps]; // stages issued so far per transposer warp
#pragma unroll
            for (int i = 0; i < kTrWarps; ++i)
                seq[i] = 0;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t tile = tile_of(it);
                if (tile >= ntiles)
                    continue;
                const int32_t x = int32_t(tile * P);
                for (uint32_t st = 0; st < nst; ++st) {
                    const uint32_t w = st % kTrWarps;
                    uint32_t k = 0;
#pragma unroll
                    for (int i = 0; i < kTrWarps; ++i) { // register array with a dynamic index: select by compare
                        if (uint32_t(i) == w) {
                            k = seq[i];
                            seq[i] = k + 1;
                        }
                    }
                    const uint32_t slot = w + kTrWarps * (k & 1u);
                    const uint32_t fill = k >> 1;
                    mbar_wait(&ring_empty[slot], (fill & 1u) ^ 1u);
                    mbar_arrive_expect_tx(&ring_full[slot], kStageBytes);
                    tma_load_2d(ring + size_t(slot) * kStageWords, &tmap, &ring_full[slot], x,
                                int32_t(st * kSlotsPerStage), l2_policy);
                }
            }
        }
        return;
    }

    if (warp >= kSelWarps) {
        // ===================== transposers =====================
        const uint32_t w = warp - kSelWarps;
        // lane L of a stage holds word column c = L & (W-1) of the P-byte row and sub-block s = L >> kColBits
        uint32_t k = 0; // stages consumed by this warp
        for (uint32_t it = 0; it < my_tiles; ++it) {
            if (tile_of(it) >= ntiles)
                continue;
            const uint32_t buf = it & 1u;
            const uint32_t q = it >> 1;
            mbar_wait(&planes_empty[buf], (q & 1u) ^ 1u); // selectors are done with the previous tile in this buffer
            uint32_t *pbuf = planes + buf * kBufWords;
            for (uint32_t st = w; st < nst; st += kTrWarps, ++k) {
                const uint32_t slot = w + kTrWarps * (k & 1u);
                mbar_wait(&ring_full[slot], (k >> 1) & 1u);
                const uint32_t *src = ring + size_t(slot) * kStageWords + lane;
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    r[i] = src[i * 32];
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(&ring_empty[slot]);
                transpose32(r);
                // r[8*p + b] = bit plane b of element 4*c+p over this lane's 32 frame slots
                const uint32_t col = lane ^ (st & 1u);
                uint4 *dst = reinterpret_cast<uint4 *>(pbuf + st * 1024u) + col;
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    dst[(p * 2 + 0) * 32] = make_uint4(r[8 * p + 0], r[8 * p + 1], r[8 * p + 2], r[8 * p + 3]);
                    dst[(p * 2 + 1) * 32] = make_uint4(r[8 * p + 4], r[8 * p + 5], r[8 * p + 6], r[8 * p + 7]);
                }
            }
            __syncwarp();
            if (lane == 0)
                mbar_arrive(&planes_full[buf]);
        }
        return;
    }

    // ===================== selectors =====================
    // warp = (byte p of the word, cx = lowest column bit); lane = logical transposer lane id with bit 0 replaced
    // by g; this thread owns stages st = 2j + g of element 4c + p, sub-block s
    const uint32_t s_p = warp >> 1;
    const uint32_t s_cx = warp & 1u;
    const uint32_t s_g = lane & 1u;
    const uint32_t s_col = lane ^ s_cx;
    const uint32_t s_L = (lane & ~1u) | s_cx;
    const uint32_t s_c = s_L & ((32u >> LOG2S) - 1u);
    const uint32_t s_elem = 4u * s_c + s_p;
    const bool s_writer = (s_g == 0u) && ((lane >> kColBits) == 0u);
    const uint32_t k0 = nframes / 2u + (nst * kSlotsPerStage - nframes); // wanted rank incl. zero pad slots

    for (uint32_t it = 0; it < my_tiles; ++it) {
        const uint32_t tile = tile_of(it);
        if (tile >= ntiles)
            continue;
        const uint32_t buf = it & 1u;
        const uint32_t q = it >> 1;
        mbar_wait(&planes_full[buf], q & 1u);
        const uint4 *base = reinterpret_cast<const uint4 *>(planes + buf * kBufWords + s_g * 1024u) + s_p * 64u + s_col;
        uint32_t alive[JT];
#pragma unroll
        for (int j = 0; j < JT; ++j)
            alive[j] = (2u * j + s_g) < nst ? 0xFFFFFFFFu : 0u; // rows >= nst hold garbage; alive == 0 masks them
        uint32_t k = k0;
        uint32_t med = 0;
#pragma unroll
        for (int bh = 1; bh >= 0; --bh) {
            uint4 w4[JT];
#pragma unroll
            for (int j = 0; j < JT; ++j)
                w4[j] = base[j * 512 + bh * 32]; // stage 2j+g, nibble bh: 4 planes in one LDS.128
#pragma unroll
            for (int bl = 3; bl >= 0; --bl) {
                uint32_t cnt = 0;
#pragma unroll
                for (int j = 0; j < JT; ++j) {
                    const uint32_t wv = bl == 0 ? w4[j].x : bl == 1 ? w4[j].y : bl == 2 ? w4[j].z : w4[j].w;
                    cnt += __popc(alive[j] & ~wv);
                }
                cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 1); // the two stage parities
                if (LOG2S >= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 16); // sub-blocks
                if (LOG2S >= 2) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 8);
                if (LOG2S >= 3) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, 4);
                const bool one = k >= cnt; // fewer than k+1 candidates have a 0 here -> the bit is 1
                if (one) {
                    k -= cnt;
                    med |= 1u << (4 * bh + bl);
                }
                const uint32_t flip = one ? 0u : 0xFFFFFFFFu;
#pragma unroll
                for (int j = 0; j < JT; ++j) {
                    const uint32_t wv = bl == 0 ? w4[j].x : bl == 1 ? w4[j].y : bl == 2 ? w4[j].z : w4[j].w;
                    alive[j] &= (wv ^ flip);
                }
            }
        }
        __syncwarp();
        if (lane == 0)
            mbar_arrive(&planes_empty[buf]);
        const size_t e = size_t(tile) * P + s_elem;
        if (s_writer && e < nelem)
            out[e] = uint8_t(med);
    }
}

template <int LOG2S, int JT>
int launch_pipe_variant(cvvp_ctx *ctx, const CUtensorMap &tmap, uint8_t *d_out, uint32_t nelem, uint32_t nframes,
                        uint32_t nst, cudaStream_t stream)
{
    constexpr int P = 128 >> LOG2S;
    const uint32_t ntiles = (nelem + P - 1) / P;
    const size_t smem_bytes = size_t(kRing) * kStageBytes + 2u * (2u * JT * 4096u) + size_t(2 * kRing + 4) * 8;
    if (smem_bytes > ctx->smem_optin)
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "median: pipelined tile does not fit shared memory");
    auto kern = median_pipe_kernel<LOG2S, JT>;
    CVVP_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes)));
    const uint32_t npairs = (ntiles + 1) / 2;
    const uint32_t grid = npairs < uint32_t(ctx->sm_count) ? npairs : uint32_t(ctx->sm_count);
    // P == 128: every line is read once -> evict-first.  P < 128: the sibling tile re-reads the line from L2.
    const uint64_t policy = (LOG2S == 0) ? kL2EvictFirst : kL2EvictNormal;
    kern<<<grid, kPipeThreads, smem_bytes, stream>>>(tmap, d_out, nelem, nframes, nst, ntiles, policy);
    CVVP_CUDA_OK(ctx, cudaGetLastError());
    ctx->launches++;
    return CVVP_OK;
}

template <int LOG2S>
int dispatch_pipe(cvvp_ctx *ctx, const CUtensorMap &tmap, uint8_t *d_out, uint32_t nelem, uint32_t nframes, uint32_t nst,
                  cudaStream_t stream)
{
    switch ((nst + 1u) / 2u) {
#define CVVP_JT_CASE(J) \
    case J: return launch_pipe_variant<LOG2S, J>(ctx, tmap, d_out, nelem, nframes, nst, stream);
        CVVP_JT_CASE(1)
        CVVP_JT_CASE(2)
        CVVP_JT_CASE(3)
        CVVP_JT_CASE(4)
        CVVP_JT_CASE(5)
        CVVP_JT_CASE(6)
        CVVP_JT_CASE(7)
        CVVP_JT_CASE(8)
#undef CVVP_JT_CASE
    default: return fail(ctx, CVVP_ERR_UNSUPPORTED, "median: unsupported stage count %u", nst);
    }
}
} // namespace

// Largest frame count the pipelined kernel takes (16 stages x 256 frame slots at P = 16).
long long median_pipe_max_frames()
{
    return 16ll * 256ll;
}

int median_pipe_launch(cvvp_ctx *ctx, const CUtensorMap &tmap, int log2s, uint8_t *d_out, uint32_t nelem,
                       uint32_t nframes, uint32_t nst, cudaStream_t stream)
{
    if (nst == 0 || nst > kMaxStages)
        return fail(ctx, CVVP_ERR_UNSUPPORTED, "median: %u stages exceed the pipelined kernel's plane buffers", nst);
    switch (log2s) {
    case 0: return dispatch_pipe<0>(ctx, tmap, d_out, nelem, nframes, nst, stream);
    case 1: return dispatch_pipe<1>(ctx, tmap, d_out, nelem, nframes, nst, stream);
    case 2: return dispatch_pipe<2>(ctx, tmap, d_out, nelem, nframes, nst, stream);
    case 3: return dispatch_pipe<3>(ctx, tmap, d_out, nelem, nframes, nst, stream);
    default: return fail(ctx, CVVP_ERR_INVALID, "median: bad tile variant");
    }
}
} // namespace cvvp
