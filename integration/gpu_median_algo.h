// Reference-side binding of the median: a TokenProcessorAlgo over the C ABI of include/cvvp.h.
// This is the file a CvVidProc maintainer would add next to Sources/ProcessorAlgos/histogram_median_algo.h (whose
// class it replaces, :56-113); it includes the reference's own token_processor_algo.h and OpenCV, nothing of this
// repo but cvvp.h, and links with -lcvvp_cuda.  INTEGRATION.md quotes it; tests/test_reference_binding_gpu.py
// compiles it against the reference's headers (oracle/Makefile target ref_binding) and runs it beside the
// reference's own classes.
#ifndef CVVP_GPU_MEDIAN_ALGO_H
#define CVVP_GPU_MEDIAN_ALGO_H

#include "token_processor_algo.h"

#include <opencv2/opencv.hpp>

#include "cvvp.h"

#include <memory>
#include <stdexcept>
#include <utility>

class GpuMedianAlgo;
template <>
struct TokenProcessorPack<GpuMedianAlgo> final {
    int device{0};
    long long frames_hint{-1};
};

class GpuMedianAlgo final : public TokenProcessorAlgo<GpuMedianAlgo, cv::Mat, cv::Mat>
{
public:
    GpuMedianAlgo() = delete;
    GpuMedianAlgo(TokenProcessorPack<GpuMedianAlgo> pack) : TokenProcessorAlgo{std::move(pack)}
    {
        if (cvvp_ctx_create(m_pack.device, &m_ctx) != CVVP_OK)
            throw std::runtime_error(cvvp_last_error(nullptr)); // the EXCEPTION_ASSERT convention: a runtime_error
    }
    GpuMedianAlgo(const GpuMedianAlgo &) = delete;
    GpuMedianAlgo &operator=(const GpuMedianAlgo &) = delete;
    ~GpuMedianAlgo() override { cvvp_ctx_destroy(m_ctx); }

    void Insert(std::unique_ptr<cv::Mat> m) override // histogram_median_algo.h:66-87
    {
        if (!m || !m->data || m->empty())
            return; // same guard as :69-70
        if (!m_started) {
            m_rows = m->rows; // geometry of the first token (:73-77)
            m_cols = m->cols;
            m_type = m->type();
            check(cvvp_median_begin(m_ctx, m->total() * m->channels(), m_pack.frames_hint));
            m_started = true;
        }
        cv::Mat c = m->isContinuous() ? *m : m->clone();
        check(cvvp_median_push(m_ctx, c.data, 1, c.total() * c.channels()));
    }
    void NotifyNoMoreTokens() override // :101-108
    {
        if (!m_started)
            return;
        cv::Mat out(m_rows, m_cols, m_type); // continuous: rows x cols x channels bytes
        check(cvvp_median_finish(m_ctx, out.data));
        m_result = std::make_unique<cv::Mat>(std::move(out));
        m_started = false;
    }
    std::unique_ptr<cv::Mat> TryGetResult() override { return std::move(m_result); } // :90-98
    bool HasResults() override { return static_cast<bool>(m_result); }               // :110-113

private:
    void check(int rc)
    {
        if (rc != CVVP_OK)
            throw std::runtime_error(cvvp_last_error(m_ctx));
    }
    cvvp_ctx *m_ctx{nullptr};
    bool m_started{false};
    int m_rows{0}, m_cols{0}, m_type{0};
    std::unique_ptr<cv::Mat> m_result{};
};

#endif
