// State of a highlight job (parameters of TokenProcessorPack<HighlightObjectsAlgo>,
// /root/reference/Sources/ProcessorAlgos/highlight_objects_algo.h:21-32, plus device scratch).  Shared by the two
// device implementations: highlight_fused.cu (the product path: one persistent CTA per frame) and highlight.cu
// (per-pixel kernels: the first implementation, kept as an on-device cross-check and for frames whose geometry the
// run records cannot encode).
#pragma once
#include "context.hpp"

#include <vector>

namespace cvvp
{
struct HlGeom {
    int W, H;
    uint32_t npix;    // W * H
    uint32_t lstride; // label-array stride per frame (npix + 1, padded to 4)
    uint32_t mstride; // mask stride per frame (npix padded to 16)
};

enum HighlightPath { kPathFused = 0, kPathPixel = 1 };

struct FusedScratch {
    int slots{0};            // resident CTAs the scratch was sized for
    uint32_t *bits{nullptr};   // [slots][kFusedImages][nwords]
    uint32_t *runs{nullptr};   // [slots][kFusedRunArrays][cap]
    uint32_t *rowoff{nullptr}; // [slots][2][rstride]
    unsigned *queue{nullptr};  // frame queue counters (one per launch in flight, ring of kQueueRing)
    unsigned queue_next{0};
    size_t bits_bytes{0}, runs_bytes{0}, rowoff_bytes{0}; // sizes of the pooled allocations above (pool.hpp)
};

// optional component output of the call in progress (cvvp_highlight_device_cc); comps == nullptr: off
struct CcOut {
    cvvp_component *comps{nullptr};
    int max_comps{0};
    int *ncomps{nullptr};
    int32_t *labels{nullptr};
    size_t labels_stride{0};
};

struct HighlightState {
    HlGeom g{};
    int th{}, lo{}, hi{}, min_hyst{}, min_th{};
    int noffs{0};
    int dy_min{0}, dy_max{0}; // row range of the structuring element's taps (offsets are sorted by row, then column)
    int path{kPathFused};
    uint8_t *d_bg{nullptr};
    short2 *d_offs{nullptr};
    std::vector<short2> h_offs; // host copy of the taps
    // per-pixel path: per-batch work buffers
    int batch_cap{0};
    uint8_t *m_a{nullptr}, *m_u{nullptr}, *m_l{nullptr}, *m_t{nullptr}, *m_out{nullptr};
    uint32_t *lab0{nullptr}, *lab1{nullptr};
    int *st_s{nullptr}, *st_e{nullptr}, *st_x{nullptr};
    // Otsu (threshold == -1): per-frame thresholds of the batch
    int th_cap{0};
    int *d_th{nullptr};
    unsigned int *d_hist{nullptr};
    // fused path
    int fused_variant{-1}; // which build of the fused kernel the job uses: 0 = 1024-thread CTAs, 1 = 256-thread CTAs, -1 = not chosen yet
    FusedScratch fs;
    CcOut cc;
    // staging of the host-buffer component entry point
    cvvp_component *d_comps{nullptr};
    int *d_ncomps{nullptr};
    int32_t *d_labels{nullptr};
    size_t comps_cap{0}, ncomps_cap{0}, labels_cap{0};
    // staging for the host-buffer entry point
    uint8_t *d_in{nullptr}, *d_res{nullptr};
    size_t in_bytes{0};
};

// highlight.cu
int highlight_thresholds(cvvp_ctx *ctx, HighlightState *st, const uint8_t *in, size_t frame_stride, unsigned nb,
                         cudaStream_t stream);
// highlight_fused.cu
bool fused_supports(const HighlightState *st);
void fused_release(HighlightState *st);
int fused_frames_in_flight(cvvp_ctx *ctx, HighlightState *st);
int highlight_fused_batch(cvvp_ctx *ctx, HighlightState *st, const uint8_t *in, size_t frame_stride, unsigned nb,
                          uint8_t *d_out, size_t out_stride, cudaStream_t stream);
} // namespace cvvp
