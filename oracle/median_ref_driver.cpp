// TEST INFRASTRUCTURE ONLY.
// Drives the reference's own HistogramMedianAlgo<T> class, compiled unmodified from
// /root/reference/Sources/ProcessorAlgos/histogram_median_algo.h against oracle/shim, through
// its plugin interface exactly as the reference's worker thread does
// (Sources/AsyncTokens/token_processing_unit.h:293 Insert, :334 NotifyNoMoreTokens,
//  :307 TryGetResult).  Built only where /root/reference is mounted; output goes to
// oracle/_ref/libcvvp_median_ref.so (git-ignored, travels to the GPU box with the snapshot).
//
// No reference SOURCE is copied here: the header is #included from where it lies.
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <thread>
#include <vector>

#include <opencv2/opencv.hpp> // oracle/shim

// The reference header calls these two helpers, which the reference declares in
// Sources/Utility/cv_util.h (:46-49) and defines in cv_util.cpp (:243-284) on top of real
// OpenCV.  Own equivalents for the shim Mat (byte copy in, byte copy + reshape out).
static bool cv_mat_to_std_vector_uchar(const cv::Mat &mat_input, std::vector<unsigned char> &vec_output)
{
    if (!mat_input.data || mat_input.empty())
        return false;
    const std::size_t n = mat_input.total() * static_cast<std::size_t>(mat_input.channels());
    vec_output.assign(mat_input.data, mat_input.data + n);
    return true;
}

static bool cv_mat_from_std_vector_uchar(cv::Mat &mat_output, const std::vector<unsigned char> &vec_input, const int rows,
                                         const int channels)
{
    mat_output = cv::Mat{vec_input, true}.reshape(channels, rows);
    return true;
}

#include "histogram_median_algo.h" // from /root/reference (include path set by oracle/Makefile)

namespace
{
template <typename T>
void run_strip(const std::uint8_t *frames, std::size_t nframes, std::size_t frame_pitch, std::size_t e0, std::size_t e1,
               std::uint8_t *out)
{
    const std::size_t n = e1 - e0;
    if (n == 0)
        return;
    HistogramMedianAlgo<T> algo{TokenProcessorPack<HistogramMedianAlgo<T>>{}};
    for (std::size_t f = 0; f < nframes; ++f) {
        // one strip token: n x 1, 8UC1
        auto token = std::make_unique<cv::Mat>(static_cast<int>(n), 1, CV_8UC1);
        std::memcpy(token->data, frames + f * frame_pitch + e0, n);
        algo.Insert(std::move(token));
    }
    algo.NotifyNoMoreTokens();
    std::unique_ptr<cv::Mat> result = algo.TryGetResult();
    if (!result || result->total() * result->channels() != n)
        throw std::runtime_error("reference median returned no result");
    std::memcpy(out + e0, result->data, n);
}
} // namespace

extern "C" __attribute__((visibility("default"))) int cvvp_ref_median(const std::uint8_t *frames, std::size_t nframes,
                                                                      std::size_t nelem, std::size_t frame_pitch,
                                                                      int bin_bytes, int nthreads, std::uint8_t *out)
{
    try {
        if (!frames || !out || nelem == 0)
            return -3;
        if (bin_bytes == 0)
            bin_bytes = nframes <= 255 ? 1 : (nframes <= 65535 ? 2 : 4);
        if (nthreads < 1)
            nthreads = 1;
        if (static_cast<std::size_t>(nthreads) > nelem)
            nthreads = static_cast<int>(nelem);
        std::vector<std::thread> workers;
        std::vector<int> rcs(static_cast<std::size_t>(nthreads), 0);
        const std::size_t base = nelem / static_cast<std::size_t>(nthreads);
        for (int t = 0; t < nthreads; ++t) {
            const std::size_t e0 = base * static_cast<std::size_t>(t);
            const std::size_t e1 = (t == nthreads - 1) ? nelem : base * static_cast<std::size_t>(t + 1);
            workers.emplace_back([=, &rcs]() {
                try {
                    switch (bin_bytes) {
                    case 1: run_strip<unsigned char>(frames, nframes, frame_pitch, e0, e1, out); break;
                    case 2: run_strip<std::uint16_t>(frames, nframes, frame_pitch, e0, e1, out); break;
                    case 4: run_strip<std::uint32_t>(frames, nframes, frame_pitch, e0, e1, out); break;
                    default: rcs[static_cast<std::size_t>(t)] = -2; break;
                    }
                } catch (...) {
                    rcs[static_cast<std::size_t>(t)] = -1;
                }
            });
        }
        for (auto &w : workers)
            w.join();
        for (int rc : rcs)
            if (rc)
                return rc;
        return 0;
    } catch (...) {
        return -1;
    }
}
