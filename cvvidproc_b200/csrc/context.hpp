// Opaque context behind the C ABI (include/cvvp.h): one CUDA device, its streams/events, the
// device frame stack of the running median job and the pinned staging ring.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/cvvp.h"

namespace cvvp
{
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct StagingBuf {
    uint8_t *host{nullptr};
    size_t bytes{0};
    cudaEvent_t done{nullptr};
    bool in_flight{false};
};

struct MedianJob {
    bool active{false};
    size_t nelem{0};
    size_t stride{0}; // device frame pitch in bytes (multiple of 128)
    long long capacity{0};
    long long count{0};
    uint8_t *d_stack{nullptr};
    uint8_t *d_out{nullptr};
    size_t d_out_bytes{0};
    size_t d_stack_bytes{0};
    // constant-memory form (median_hist.cu): frames folded into value histograms when the stack cannot grow
    uint32_t *d_hist{nullptr}; // [256][stride] counts, or nullptr
    size_t d_hist_bytes{0};
    long long folded{0};       // frames that live in d_hist only (count = the frames resident in d_stack)
    // after the first fold the stack is used as two halves: one is filled while the other is being folded
    bool split{false};
    int half{0};
    long long half_cap{0};     // frames per half
    long long base{0};         // first frame slot of the region being filled (0 or half_cap)
    cudaEvent_t ev_fold[2]{nullptr, nullptr}; // the fold that last read half h is complete
};
struct HighlightState; // highlight_state.hpp

// Where the counting rounds of a frame-sharded median push their per-element nibble counts (median_shard.cu):
// element e belongs to rank e / slice, whose receive area for THIS rank's counts starts at dst[e / slice]
// (8 words = 16 u16 counts per element, peer memory over NVLink or local).
constexpr int kMaxShardRanks = 16;
struct ShardPush {
    uint32_t *dst[kMaxShardRanks];
    const uint32_t *sel;   // round 2: per element, the globally selected high nibble in bits 0..3
    const uint32_t *accum; // counts of this rank's earlier frame chunks (8 words per element), or nullptr
    uint32_t slice;
    uint32_t stage; // 1: round 1 stages a tile's count vectors in shared memory and stores them 512 B per warp (peer owners)
    // not NULL: the launch only takes the *tile_count tiles listed in tile_list (128 elements each) instead of all of
    // them -- the two counting rounds behind a window pass recount only the tiles that hold an undecided element
    const uint32_t *tile_list;
    const uint32_t *tile_count;
    // window counting (MODE 3): a source's frame count goes to one header word per owner, not into every record
    uint32_t *hdr[kMaxShardRanks];
    uint32_t nranks;
};
struct MedianShard; // median_shard.cu

// device staging of decoded frames waiting for the preparation kernel (frames.cu), double-buffered
struct RawStage {
    uint8_t *d[2]{nullptr, nullptr};
    size_t cap{0}; // bytes per half
    cudaEvent_t up[2]{nullptr, nullptr};       // H2D of this half complete
    cudaEvent_t consumed[2]{nullptr, nullptr}; // the kernel that read this half is complete
    bool used[2]{false, false};
    int next{0};
    uint8_t *d_out{nullptr}; // result staging of the host-buffer form (cvvp_frames_prepare)
    size_t out_cap{0};
};
struct HighlightQueue; // frame_pipeline.cu
} // namespace cvvp

struct cvvp_ctx {
    int device{0};
    int sm_count{0};
    int cc_major{0};
    int cc_minor{0};
    size_t smem_optin{0};
    cudaStream_t compute{nullptr};
    cudaStream_t copy{nullptr};     // host -> device
    cudaStream_t copy_out{nullptr}; // device -> host (so both directions of the link overlap)
    cudaEvent_t ev_start{nullptr};
    cudaEvent_t ev_stop{nullptr};
    cudaEvent_t ev_copy{nullptr};
    bool have_kernel_time{false};
    long long launches{0};
    std::string err;
    cvvp::EncodeTiledFn encode_tiled{nullptr};
    cvvp::MedianJob med;
    cvvp::HighlightState *hl{nullptr};
    cvvp::MedianShard *shard{nullptr};
    cvvp::MedianShard *big{nullptr}; // internal one-rank job of the two-pass path for long stacks (median.cu)
    std::vector<cvvp::StagingBuf> staging;
    size_t staging_next{0};
    cvvp::RawStage raw;
    cvvp::HighlightQueue *hq{nullptr};
};

namespace cvvp
{
void set_global_error(const char *fmt, ...);
const char *global_error();

int fail(cvvp_ctx *ctx, int code, const char *fmt, ...);

#define CVVP_CUDA_OK(ctx, expr)                                                                                       \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess)                                                                                         \
            return ::cvvp::fail((ctx), CVVP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),          \
                                __FILE__, __LINE__);                                                                   \
    } while (0)

// RAII device switch: every ABI call makes its context's device current and restores the caller's.
struct DeviceGuard {
    int prev{-1};
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess)
            prev = -1;
        if (prev != dev)
            cudaSetDevice(dev);
        else
            prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};

// median.cu
int median_launch(cvvp_ctx *ctx, const uint8_t *d_frames, long long nframes, size_t nelem, size_t frame_stride,
                  uint8_t *d_out, cudaStream_t stream);
int median_launch_mode(cvvp_ctx *ctx, const uint8_t *d_frames, long long nframes, size_t nelem, size_t frame_stride,
                       uint8_t *d_out, int mode, const ShardPush &push, cudaStream_t stream);
// median_pipe.cu
long long median_max_frames();
int median_pipe_launch(cvvp_ctx *ctx, const CUtensorMap &tmap, int log2s, uint8_t *d_out, uint32_t nelem, uint32_t nframes,
                       uint32_t nst, int mode, const ShardPush &push, cudaStream_t stream);
// median_hist.cu
size_t median_hist_bytes(size_t frame_stride);
int median_hist_fold(cvvp_ctx *ctx, const uint8_t *d_stack, long long nframes, size_t frame_stride, uint32_t *d_hist,
                     cudaStream_t stream);
int median_hist_select(cvvp_ctx *ctx, const uint32_t *d_hist, size_t frame_stride, size_t nelem, long long total, uint8_t *d_out,
                       cudaStream_t stream);
// median_shard.cu
void median_shard_release(cvvp_ctx *ctx);
int median_long_stack(cvvp_ctx *ctx, const uint8_t *d_frames, long long nframes, size_t nelem, size_t frame_stride,
                      uint8_t *d_out, cudaStream_t stream, int window);
long long median_two_pass_max_frames();
// highlight.cu
void highlight_release(cvvp_ctx *ctx);
int highlight_begin(cvvp_ctx *ctx, const uint8_t *background, int width, int height, const uint8_t *selem, int kw, int kh,
                    int threshold, int threshold_lo, int threshold_hi, int min_size_hyst, int min_size_threshold);
int highlight_device(cvvp_ctx *ctx, const uint8_t *d_frames, long long n, size_t frame_stride, uint8_t *d_out,
                     size_t out_stride, cudaStream_t stream);
int highlight_device_cc(cvvp_ctx *ctx, const uint8_t *d_frames, long long n, size_t frame_stride, uint8_t *d_out,
                        size_t out_stride, cvvp_component *d_comps, int max_comps, int *d_ncomps, int32_t *d_labels,
                        size_t labels_stride, cudaStream_t stream);
int highlight_frames_host_cc(cvvp_ctx *ctx, const uint8_t *frames, long long n, size_t frame_stride, uint8_t *masks_out,
                             size_t out_stride, cvvp_component *comps_out, int max_comps, int *ncomps_out,
                             int32_t *labels_out, size_t labels_stride);
int highlight_set_path(cvvp_ctx *ctx, int path);
int highlight_frames_in_flight(cvvp_ctx *ctx, int *out_frames);
int highlight_frames_host(cvvp_ctx *ctx, const uint8_t *frames, long long n, size_t frame_stride, uint8_t *masks_out,
                          size_t out_stride);
// frames.cu
size_t frames_out_bytes(const cvvp_frame_format &f);
int frames_check_format(cvvp_ctx *ctx, const cvvp_frame_format *f);
int frames_prepare_launch(cvvp_ctx *ctx, const uint8_t *d_src, long long n, size_t src_stride, size_t src_bytes,
                          const cvvp_frame_format &f, int band_row0, uint8_t *d_dst, size_t dst_stride, cudaStream_t stream);
// frame_pipeline.cu
void raw_stage_release(cvvp_ctx *ctx);
int raw_stage_ensure(cvvp_ctx *ctx, size_t bytes_per_half);
int frames_prepare_host(cvvp_ctx *ctx, const uint8_t *frames, long long n, size_t frame_stride, const cvvp_frame_format &f,
                        uint8_t *out, size_t out_stride);
int frames_upload_prepare(cvvp_ctx *ctx, const uint8_t *frames, long long n, size_t frame_stride, const cvvp_frame_format &f,
                          uint8_t *d_dst, size_t dst_stride);
void highlight_queue_release(cvvp_ctx *ctx);
int highlight_queue_begin(cvvp_ctx *ctx, int depth, long long max_batch, const cvvp_frame_format *fmt, int max_comps);
int highlight_queue_pending(const cvvp_ctx *ctx);
int highlight_queue_submit(cvvp_ctx *ctx, const uint8_t *frames, long long n, size_t frame_stride);
int highlight_queue_ready(cvvp_ctx *ctx);
int highlight_queue_next(cvvp_ctx *ctx, uint8_t *masks_out, size_t out_stride, long long *n_out, cvvp_component *comps_out,
                         int *ncomps_out);
int highlight_slot_acquire(cvvp_ctx *ctx, uint8_t **h_frames, size_t *frame_pitch, long long *max_frames);
int highlight_slot_commit(cvvp_ctx *ctx, long long n);
int highlight_queue_next_view(cvvp_ctx *ctx, const uint8_t **h_masks, size_t *mask_pitch, long long *n_out,
                              const cvvp_component **comps, const int **ncomps);
int highlight_queue_view_release(cvvp_ctx *ctx);
// highlight.cu: geometry of the running job (false: no job)
bool highlight_geometry(const cvvp_ctx *ctx, int *width, int *height);
// synth.cu
int synth_launch(cvvp_ctx *ctx, uint8_t *d_frames, size_t frame_stride, int width, int height, int row0, int nrows,
                 long long first_frame, long long nframes, uint32_t seed, int ndisks, cudaStream_t stream);
} // namespace cvvp
