// TEST INFRASTRUCTURE ONLY.
// Python module `cvvp_frames_ref`: the reference's own CvVidFramesGeneratorAlgo, compiled UNMODIFIED from
//     /root/reference/Sources/ProcessorTokenHandlers/cv_vid_frames_generator_algo.h
// against oracle/shim_cv2 (cv::VideoCapture, cv::extractChannel, cv::cvtColor forward to the `cv2` wheel).  The class is
// driven the way the reference's generator thread drives it: GetTokenSet() until it returns an empty set
// (Sources/AsyncTokens/token_batch_generator.h, async_token_batch_generator.h).  It pins oracle/frames_oracle.py: the
// frame range, the crop, and the three channel modes (:130-156).
//
// Built only where /root/reference is mounted, into oracle/_ref/ (git-ignored, travels with the snapshot).  Only tests/
// and tests/golden/make_frames_golden.py import it.  No reference source is copied here.
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp> // oracle/shim_cv2

#include "cv_vid_frames_generator_algo.h" // from /root/reference (include path set by oracle/Makefile)

namespace py = pybind11;

// The header calls these; the reference defines them in Sources/Utility/exception_assert.cpp (which needs the
// CMake-generated project_config.h) and Sources/Utility/cv_util.cpp (strip splitting on real OpenCV; SURVEY section 2
// marks it out of scope, and with one chunk per frame the generator never calls it).  Own equivalents.
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
bool cv_mat_to_chunks(const cv::Mat &, std::vector<std::unique_ptr<cv::Mat>> &, const int, const int, int, int)
{
    throw std::logic_error("cv_mat_to_chunks: strip splitting is not part of this driver (chunks_per_frame is 1)");
}

namespace
{
class RefFrames
{
public:
    RefFrames(const std::string &vid_path, long long start_frame, long long last_frame, int crop_x, int crop_y, int crop_w,
              int crop_h, bool convert_to_grayscale, bool vid_is_grayscale, int frames_in_batch)
        : m_algo{TokenGeneratorPack<CvVidFramesGeneratorAlgo>{frames_in_batch, frames_in_batch, 1, vid_path, start_frame,
                                                             last_frame, cv::Rect{crop_x, crop_y, crop_w, crop_h},
                                                             convert_to_grayscale, vid_is_grayscale, 0, 0}}
    {
    }

    // one batch: the tokens of up to frames_in_batch frames (own copies), [] at the end of the range / stream
    py::list get_token_set()
    {
        py::list out;
        for (auto &token : m_algo.GetTokenSet())
            if (token && token->has_array())
                out.append(py::module_::import("numpy").attr("ascontiguousarray")(token->array()).attr("copy")());
        return out;
    }

private:
    CvVidFramesGeneratorAlgo m_algo;
};
} // namespace

PYBIND11_MODULE(cvvp_frames_ref, m)
{
    m.doc() = "TEST INFRASTRUCTURE ONLY: the reference's CvVidFramesGeneratorAlgo compiled unmodified against a cv2-forwarding shim";
    py::module_ cv2 = py::module_::import("cv2");
    auto same = [&](const char *name, int v) {
        if (cv2.attr(name).cast<int>() != v)
            throw std::runtime_error(std::string("opencv shim: cv2.") + name + " differs from the shim's value");
    };
    same("CAP_PROP_POS_FRAMES", cv::CAP_PROP_POS_FRAMES);
    same("CAP_PROP_FRAME_WIDTH", cv::CAP_PROP_FRAME_WIDTH);
    same("CAP_PROP_FRAME_HEIGHT", cv::CAP_PROP_FRAME_HEIGHT);
    same("CAP_PROP_FRAME_COUNT", cv::CAP_PROP_FRAME_COUNT);
    same("CAP_PROP_FORMAT", cv::CAP_PROP_FORMAT);
    same("CAP_PROP_CONVERT_RGB", cv::CAP_PROP_CONVERT_RGB);
    same("COLOR_RGB2GRAY", cv::COLOR_RGB2GRAY);
    same("CV_8UC1", CV_8UC1);
    same("CV_8UC3", CV_8UC3);
    same("CV_8UC4", CV_8UC4);
    m.attr("opencv_version") = cv2.attr("__version__");

    py::class_<RefFrames>(m, "RefFrames")
        .def(py::init<const std::string &, long long, long long, int, int, int, int, bool, bool, int>(), py::arg("vid_path"),
             py::arg("start_frame"), py::arg("last_frame"), py::arg("crop_x"), py::arg("crop_y"), py::arg("crop_w"),
             py::arg("crop_h"), py::arg("convert_to_grayscale"), py::arg("vid_is_grayscale"), py::arg("frames_in_batch") = 1)
        .def("get_token_set", &RefFrames::get_token_set);
}
