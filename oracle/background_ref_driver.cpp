// TEST INFRASTRUCTURE ONLY.
// Python module `cvvp_background_ref`: the reference's own GetVideoBackground -- the ENTRY POINT, with everything behind
// it -- compiled UNMODIFIED from /root/reference where the sources lie:
//     Sources/cv_vid_bg_helpers.cpp          GetVideoBackground, VidBackgroundWithAlgo, GetCroppedFrameDims
//     Sources/Utility/cv_util.cpp            cv_mat_to_chunks / cv_mat_from_chunks / vector conversions
//     Sources/AsyncTokens/*.h                generator and worker threads, queues
//     Sources/ProcessorTokenHandlers/*.h     CvVidFramesGeneratorAlgo, CvVidFragmentConsumer
//     Sources/ProcessorAlgos/histogram_median_algo.h   HistogramMedianAlgo8/16/32
// against oracle/shim_cv2 (cv::Mat over numpy, cv::VideoCapture / cvtColor / extractChannel forwarded to the cv2 wheel;
// every shim call takes the GIL itself because the reference runs them on its own threads).  oracle/Makefile passes
// the two .cpp files to the compiler from where they lie; no reference source is copied here.
//
// It pins what no single class does: the crop rule (GetCroppedFrameDims :39-60 with its :56 quirk), frame_limit
// (:226-229), the bin-width dispatch (:232-253), the per-generator frame ranges (:84-120), the strip split and
// re-assembly (cv_util.cpp), batch sizes from max_threads (:166-193).  tests/test_oracle_background.py holds the oracles
// to it; tests/test_reference_chain_gpu.py holds the drop-in module's GetVideoBackground to it, pack for pack.
//
// Built only where /root/reference is mounted, into oracle/_ref/ (git-ignored, travels with the snapshot).
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>

#include <opencv2/opencv.hpp> // oracle/shim_cv2

#include "cv_vid_bg_helpers.h" // from /root/reference (include path set by oracle/Makefile)

namespace py = pybind11;

// Definitions the two .cpp files link against and that live in reference files which cannot be compiled here:
// Sources/Utility/exception_assert.cpp needs the CMake-generated project_config.h, Sources/main.cpp is the demo CLI
// (cv::CommandLineParser, config::videos_dir).  Own equivalents: the assert throws std::runtime_error with the same
// fields; GetAdditionalThreads restates main.cpp:36-54 (threads available above min_threads, capped by max_threads).
void exception_assert(std::string expr, std::string func, std::string file, int line, std::string msg)
{
    std::string text = file + ":" + std::to_string(line) + ": assert failed in function '" + func + "()'\n" + expr;
    if (!msg.empty())
        text += "\nassert msg: " + msg;
    throw std::runtime_error(text);
}
void exception_assert(std::string expr, std::string func, std::string file, int line)
{
    exception_assert(std::move(expr), std::move(func), std::move(file), line, "");
}
int GetAdditionalThreads(int min_threads, int extra_threads, int max_threads)
{
    if (extra_threads < 0)
        extra_threads = 0;
    if (min_threads < 0)
        min_threads = 0;
    const int supported = int(std::thread::hardware_concurrency());
    if (max_threads < 1 || (supported && max_threads > supported + extra_threads))
        max_threads = supported + extra_threads;
    return max_threads > min_threads ? max_threads - min_threads : 0;
}

PYBIND11_MODULE(cvvp_background_ref, m)
{
    m.doc() = "TEST INFRASTRUCTURE ONLY: the reference's GetVideoBackground compiled unmodified against a cv2-forwarding shim";
    m.def(
        "GetVideoBackground",
        [](const std::string &vid_path, const std::string &bg_algo, int max_threads, long long frame_limit, bool grayscale,
           bool vid_is_grayscale, int crop_x, int crop_y, int crop_width, int crop_height, int token_storage_limit,
           bool print_timing_report) -> py::object {
            const VidBgPack pack{vid_path, bg_algo, max_threads, frame_limit, grayscale, vid_is_grayscale, crop_x, crop_y,
                                 crop_width, crop_height, token_storage_limit, print_timing_report};
            cv::Mat bg;
            {
                py::gil_scoped_release nogil; // the reference's threads call back into the shim, which takes the GIL
                bg = GetVideoBackground(pack);
            }
            std::cout.flush(); // the reference ends its report lines with '\n'; tests read them back at once
            if (!bg.has_array() || bg.empty())
                return py::none();
            return bg.array();
        },
        py::arg("vid_path"), py::arg("bg_algo") = "hist", py::arg("max_threads") = -1, py::arg("frame_limit") = -1,
        py::arg("grayscale") = false, py::arg("vid_is_grayscale") = false, py::arg("crop_x") = 0, py::arg("crop_y") = 0,
        py::arg("crop_width") = 0, py::arg("crop_height") = 0, py::arg("token_storage_limit") = -1,
        py::arg("print_timing_report") = false);
    m.def("GetCroppedFrameDims", [](int x, int y, int width, int height, int hor_pixels, int vert_pixels) {
        const cv::Rect r = GetCroppedFrameDims(x, y, width, height, hor_pixels, vert_pixels);
        return py::make_tuple(r.x, r.y, r.width, r.height);
    });
}
