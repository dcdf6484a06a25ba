// C++ host layer over the C ABI (include/cvvp.h): the two hot-path operators with the reference's own plugin shape.
//
// The reference's operator interface is TokenProcessorAlgo<Algo, TokenT, ResultT>
// (/root/reference/Sources/AsyncTokens/token_processor_algo.h:52-61):
//     Insert(unique_ptr<TokenT>), TryGetResult() -> unique_ptr<ResultT> (null = none), NotifyNoMoreTokens(), HasResults()
// with the parameter pack moved in by value (:37).  GpuMedianAlgo / GpuHighlightAlgo keep exactly that shape; the only
// change is that a token is a BATCH of frames (FrameBatch) instead of one cv::Mat, because a GPU wants many frames
// per submission.  No OpenCV types: OpenCV C++ is not part of this build (decode is done by the caller).
//
// Error convention (Sources/Utility/exception_assert.cpp:22-33): validation failures throw std::runtime_error with a
// "cvvidproc(<ver>) <file>:<line>: assert failed in function '<f>()'" text; C-ABI failures are converted the same way.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "cvvp.h"

#define CVVP_HOST_VERSION "b200-0.1"

namespace cvvp_host
{
[[noreturn]] inline void assert_fail(const char *expr, const char *func, const char *file, int line, const char *msg)
{
    std::ostringstream os;
    os << "cvvidproc(" << CVVP_HOST_VERSION << ") " << file << ':' << line << ": assert failed in function '" << func
       << "()'\n" << expr;
    if (msg && *msg)
        os << "\nassert msg: " << msg;
    throw std::runtime_error(os.str());
}
#define CVVP_ASSERT(expr) \
    do { if (!(expr)) ::cvvp_host::assert_fail(#expr, __func__, __FILE__, __LINE__, ""); } while (0)
#define CVVP_ASSERT_MSG(expr, msg) \
    do { if (!(expr)) ::cvvp_host::assert_fail(#expr, __func__, __FILE__, __LINE__, msg); } while (0)

// owning 8-bit image / image batch: n frames of rows x cols x channels bytes, contiguous
struct FrameBatch {
    int n{0}, rows{0}, cols{0}, channels{1};
    std::vector<std::uint8_t> data;
    std::size_t frame_bytes() const { return std::size_t(rows) * cols * channels; }
    bool empty() const { return n == 0 || data.empty(); }
};

class Context
{
public:
    explicit Context(int device = -1)
    {
        if (cvvp_ctx_create(device, &m_ctx) != CVVP_OK)
            throw std::runtime_error(std::string("cvvidproc(" CVVP_HOST_VERSION "): ") + cvvp_last_error(nullptr));
    }
    ~Context() { cvvp_ctx_destroy(m_ctx); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    cvvp_ctx *get() const { return m_ctx; }
    void check(int rc) const
    {
        if (rc != CVVP_OK)
            throw std::runtime_error(std::string("cvvidproc(" CVVP_HOST_VERSION "): ") + cvvp_last_error(m_ctx));
    }

private:
    cvvp_ctx *m_ctx{nullptr};
};

// ---------------------------------------------------------------------------------------------------------------------
// temporal median: replaces HistogramMedianAlgo<T> (Sources/ProcessorAlgos/histogram_median_algo.h:41-222)
// ---------------------------------------------------------------------------------------------------------------------
struct GpuMedianPack {
    int device{-1};
    long long frames_hint{-1}; // frames_to_analyze of GetVideoBackground (cv_vid_bg_helpers.cpp:226-229)
};

class GpuMedianAlgo
{
public:
    using token_type = FrameBatch;
    using result_type = FrameBatch;
    explicit GpuMedianAlgo(GpuMedianPack pack) : m_pack{pack}, m_ctx{pack.device} {}

    // histogram_median_algo.h:66-87 -- null / empty tokens are skipped without counting
    void Insert(std::unique_ptr<FrameBatch> batch)
    {
        if (!batch || batch->empty())
            return;
        if (!m_started) {
            m_rows = batch->rows;
            m_cols = batch->cols;
            m_channels = batch->channels;
            m_ctx.check(cvvp_median_begin(m_ctx.get(), batch->frame_bytes(), m_pack.frames_hint));
            m_started = true;
        }
        CVVP_ASSERT_MSG(batch->rows == m_rows && batch->cols == m_cols && batch->channels == m_channels,
                        "all frames must have the geometry of the first one");
        m_ctx.check(cvvp_median_push(m_ctx.get(), batch->data.data(), batch->n, batch->frame_bytes()));
    }
    // raw form for callers that already hold the bytes (no extra copy)
    void InsertRaw(const std::uint8_t *frames, long long n, int rows, int cols, int channels, std::size_t stride)
    {
        if (!frames || n <= 0)
            return;
        if (!m_started) {
            m_rows = rows;
            m_cols = cols;
            m_channels = channels;
            m_ctx.check(cvvp_median_begin(m_ctx.get(), std::size_t(rows) * cols * channels, m_pack.frames_hint));
            m_started = true;
        }
        CVVP_ASSERT_MSG(rows == m_rows && cols == m_cols && channels == m_channels,
                        "all frames must have the geometry of the first one");
        m_ctx.check(cvvp_median_push(m_ctx.get(), frames, n, stride));
    }
    // :101-108 -- end of stream: publish the result once, reset
    void NotifyNoMoreTokens()
    {
        if (!m_started)
            return;
        auto out = std::make_unique<FrameBatch>();
        out->n = 1;
        out->rows = m_rows;
        out->cols = m_cols;
        out->channels = m_channels;
        out->data.resize(out->frame_bytes());
        m_started = false;
        m_ctx.check(cvvp_median_finish(m_ctx.get(), out->data.data()));
        m_result = std::move(out);
    }
    std::unique_ptr<FrameBatch> TryGetResult() { return std::move(m_result); } // :90-98
    bool HasResults() const { return static_cast<bool>(m_result); }            // :110-113
    long long FramesInserted() const { return cvvp_median_count(m_ctx.get()); }

private:
    GpuMedianPack m_pack;
    Context m_ctx;
    bool m_started{false};
    int m_rows{0}, m_cols{0}, m_channels{1};
    std::unique_ptr<FrameBatch> m_result{};
};

// ---------------------------------------------------------------------------------------------------------------------
// highlight: replaces HighlightObjectsAlgo (Sources/ProcessorAlgos/highlight_objects_algo.h:21-107)
// ---------------------------------------------------------------------------------------------------------------------
struct GpuHighlightPack { // TokenProcessorPack<HighlightObjectsAlgo> :21-32
    int device{-1};
    FrameBatch background{};     // 1 frame, 1 channel
    FrameBatch struct_element{}; // 1 "frame" of kh x kw bytes
    int threshold{}, threshold_lo{}, threshold_hi{}, min_size_hyst{}, min_size_threshold{}, width_border{};
    // opt-in extras (no counterpart in the reference's pack): 8-connected components of every mask for the tracker
    bool components{false};
    bool labels{false};
    int max_components{1024};
};

class GpuHighlightAlgo
{
public:
    using token_type = FrameBatch;
    using result_type = FrameBatch;
    explicit GpuHighlightAlgo(GpuHighlightPack pack) : m_pack{std::move(pack)}, m_ctx{m_pack.device}
    {
        CVVP_ASSERT(!m_pack.background.empty());
        CVVP_ASSERT(!m_pack.struct_element.empty());
        CVVP_ASSERT_MSG(m_pack.background.channels == 1, "the highlight stage works on single-channel frames");
        m_ctx.check(cvvp_highlight_begin(m_ctx.get(), m_pack.background.data.data(), m_pack.background.cols,
                                         m_pack.background.rows, m_pack.struct_element.data.data(),
                                         m_pack.struct_element.cols, m_pack.struct_element.rows, m_pack.threshold,
                                         m_pack.threshold_lo, m_pack.threshold_hi, m_pack.min_size_hyst,
                                         m_pack.min_size_threshold, m_pack.width_border));
    }
    // highlight_objects_algo.h:60-69 -- the token is processed in place and becomes the result
    void Insert(std::unique_ptr<FrameBatch> batch)
    {
        if (!batch || batch->empty())
            return;
        CVVP_ASSERT_MSG(batch->channels == 1 && batch->rows == m_pack.background.rows && batch->cols == m_pack.background.cols,
                        "frame geometry must match the background");
        if (m_pack.components) {
            m_comps.assign(std::size_t(batch->n) * m_pack.max_components, cvvp_component{});
            m_ncomps.assign(std::size_t(batch->n), 0);
            if (m_pack.labels)
                m_labels.resize(std::size_t(batch->n) * batch->frame_bytes());
            m_ctx.check(cvvp_highlight_frames_cc(m_ctx.get(), batch->data.data(), batch->n, batch->frame_bytes(),
                                                 batch->data.data(), batch->frame_bytes(), m_comps.data(),
                                                 m_pack.max_components, m_ncomps.data(),
                                                 m_pack.labels ? m_labels.data() : nullptr, batch->frame_bytes()));
        } else {
            m_ctx.check(cvvp_highlight_frames(m_ctx.get(), batch->data.data(), batch->n, batch->frame_bytes(),
                                              batch->data.data(), batch->frame_bytes()));
        }
        m_result = std::move(batch);
    }
    // components of frame i of the last result batch (valid until the next Insert)
    int component_count(int i) const { return m_ncomps.empty() ? 0 : m_ncomps[std::size_t(i)]; }
    const cvvp_component *components(int i) const { return m_comps.data() + std::size_t(i) * m_pack.max_components; }
    const std::int32_t *labels(int i, std::size_t npix) const { return m_labels.empty() ? nullptr : m_labels.data() + std::size_t(i) * npix; }
    int max_components() const { return m_pack.max_components; }
    std::unique_ptr<FrameBatch> TryGetResult() { return std::move(m_result); } // :72-79
    void NotifyNoMoreTokens() {}                                               // :82-85 (tokens are independent)
    bool HasResults() const { return static_cast<bool>(m_result); }            // :88-91

private:
    GpuHighlightPack m_pack;
    Context m_ctx;
    std::unique_ptr<FrameBatch> m_result{};
    std::vector<cvvp_component> m_comps{};
    std::vector<int> m_ncomps{};
    std::vector<std::int32_t> m_labels{};
};
} // namespace cvvp_host
