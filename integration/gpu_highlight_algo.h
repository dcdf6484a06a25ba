// Reference-side binding of the highlight stage: a TokenProcessorAlgo over the C ABI of include/cvvp.h.
// This is the file a CvVidProc maintainer would add next to Sources/ProcessorAlgos/highlight_objects_algo.h (whose
// class it replaces, :36-98); the pack keeps the reference's fields (:21-32) and gains `device`.  INTEGRATION.md quotes
// it; tests/test_reference_binding_gpu.py compiles it against the reference's headers (oracle/Makefile target
// ref_binding) and runs it beside the reference's own HighlightObjectsAlgo on the same tokens.
#ifndef CVVP_GPU_HIGHLIGHT_ALGO_H
#define CVVP_GPU_HIGHLIGHT_ALGO_H

#include "token_processor_algo.h"

#include <opencv2/opencv.hpp>

#include "cvvp.h"

#include <memory>
#include <stdexcept>
#include <utility>

class GpuHighlightAlgo;
template <>
struct TokenProcessorPack<GpuHighlightAlgo> final {
    cv::Mat background{};
    cv::Mat struct_element{};
    const int threshold{};
    const int threshold_lo{};
    const int threshold_hi{};
    const int min_size_hyst{};
    const int min_size_threshold{};
    const int width_border{};
    const int device{0};
};

class GpuHighlightAlgo final : public TokenProcessorAlgo<GpuHighlightAlgo, cv::Mat, cv::Mat>
{
public:
    GpuHighlightAlgo() = delete;
    GpuHighlightAlgo(TokenProcessorPack<GpuHighlightAlgo> p) : TokenProcessorAlgo{std::move(p)}
    {
        if (cvvp_ctx_create(m_pack.device, &m_ctx) != CVVP_OK)
            throw std::runtime_error(cvvp_last_error(nullptr));
        cv::Mat se;
        m_pack.struct_element.convertTo(se, CV_8U); // morphologyEx treats every non-zero entry as set
        check(cvvp_highlight_begin(m_ctx, m_pack.background.data, m_pack.background.cols, m_pack.background.rows, se.data,
                                   se.cols, se.rows, m_pack.threshold, m_pack.threshold_lo, m_pack.threshold_hi,
                                   m_pack.min_size_hyst, m_pack.min_size_threshold, m_pack.width_border));
    }
    GpuHighlightAlgo(const GpuHighlightAlgo &) = delete;
    GpuHighlightAlgo &operator=(const GpuHighlightAlgo &) = delete;
    ~GpuHighlightAlgo() override { cvvp_ctx_destroy(m_ctx); }

    void Insert(std::unique_ptr<cv::Mat> f) override // highlight_objects_algo.h:60-69: in place, the token is the result
    {
        if (!f || !f->data || f->empty())
            return;
        if (!f->isContinuous())
            *f = f->clone();
        check(cvvp_highlight_frames(m_ctx, f->data, 1, f->total(), f->data, f->total()));
        m_result = std::move(f);
    }
    std::unique_ptr<cv::Mat> TryGetResult() override { return std::move(m_result); } // :72-79
    void NotifyNoMoreTokens() override {}                                            // :82-85
    bool HasResults() override { return static_cast<bool>(m_result); }               // :88-91

private:
    void check(int rc)
    {
        if (rc != CVVP_OK)
            throw std::runtime_error(cvvp_last_error(m_ctx));
    }
    cvvp_ctx *m_ctx{nullptr};
    std::unique_ptr<cv::Mat> m_result{};
};

#endif
