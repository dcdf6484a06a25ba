// TEST INFRASTRUCTURE ONLY.
// Python module `cvvp_highlight_ref`: the reference's own HighlightObjectsAlgo, compiled UNMODIFIED from
//     /root/reference/Sources/ProcessorAlgos/highlight_objects_algo.{h,cpp}
// (oracle/Makefile passes the .cpp to the compiler where it lies; no reference source is copied here) against
// oracle/shim_cv2, whose cv:: functions forward to the `cv2` wheel.  The class is driven through its plugin interface
// exactly as the reference's worker thread drives it (Sources/AsyncTokens/token_processing_unit.h:293 Insert,
// :307 TryGetResult); its public stage functions are exposed one by one as well, so that every function of
// oracle/highlight_oracle.py can be held to the reference function it restates.
//
// Built only where /root/reference is mounted, into oracle/_ref/ (git-ignored, travels to the GPU box with the
// snapshot).  Only tests/ and tests/golden/make_highlight_golden.py import it.
#include <memory>
#include <stdexcept>
#include <utility>

#include <opencv2/opencv.hpp> // oracle/shim_cv2

#include "highlight_objects_algo.h" // from /root/reference (include path set by oracle/Makefile)

namespace py = pybind11;

namespace
{
cv::Mat mat_copy_of(const py::array &a)
{
    // the token owns its pixels: an own contiguous copy, never the caller's array
    return cv::Mat{py::module_::import("numpy").attr("array")(a, py::arg("copy") = true, py::arg("order") = "C").cast<py::array>()};
}

class RefHighlight
{
public:
    RefHighlight(const py::array &background, const py::array &struct_element, int threshold, int threshold_lo,
                 int threshold_hi, int min_size_hyst, int min_size_threshold, int width_border)
        : m_algo{TokenProcessorPack<HighlightObjectsAlgo>{mat_copy_of(background), mat_copy_of(struct_element), threshold,
                                                          threshold_lo, threshold_hi, min_size_hyst, min_size_threshold,
                                                          width_border}}
    {
    }

    // one token through Insert / HasResults / TryGetResult; None when the operator produced no result (empty token)
    py::object insert(const py::object &frame)
    {
        std::unique_ptr<cv::Mat> token{};
        if (!frame.is_none())
            token = std::make_unique<cv::Mat>(mat_copy_of(frame.cast<py::array>()));
        m_algo.Insert(std::move(token));
        if (!m_algo.HasResults())
            return py::none();
        std::unique_ptr<cv::Mat> result = m_algo.TryGetResult();
        if (!result)
            throw std::runtime_error("HasResults() was true but TryGetResult() returned nothing");
        return result->array();
    }
    void notify_no_more_tokens() { m_algo.NotifyNoMoreTokens(); }
    bool has_results() { return m_algo.HasResults(); }

    py::array threshold_image(const py::array &image, int threshold)
    {
        cv::Mat im = mat_copy_of(image);
        return m_algo.ThresholdImage(im, threshold).array();
    }
    py::array threshold_image_with_hysteresis(const py::array &image, int lo, int hi)
    {
        cv::Mat im = mat_copy_of(image);
        return m_algo.ThresholdImageWithHysteresis(im, lo, hi).array();
    }
    py::array remove_small_objects(const py::array &image, int min_size)
    {
        cv::Mat im = mat_copy_of(image);
        m_algo.RemoveSmallObjects(im, min_size);
        return im.array();
    }
    py::array fill_holes(const py::array &image)
    {
        cv::Mat im = mat_copy_of(image);
        m_algo.FillHoles(im);
        return im.array();
    }
    py::array frame_and_fill(const py::array &image, int width_border) // dead code upstream (.cpp:71), kept callable
    {
        cv::Mat im = mat_copy_of(image);
        m_algo.FrameAndFill(im, width_border);
        return im.array();
    }

private:
    HighlightObjectsAlgo m_algo;
};
} // namespace

PYBIND11_MODULE(cvvp_highlight_ref, m)
{
    m.doc() = "TEST INFRASTRUCTURE ONLY: the reference's HighlightObjectsAlgo compiled unmodified against a cv2-forwarding shim";
    // the shim's enum values are compile-time copies of OpenCV's: hold them to the cv2 that will execute the calls
    py::module_ cv2 = py::module_::import("cv2");
    auto same = [&](const char *name, int v) {
        if (cv2.attr(name).cast<int>() != v)
            throw std::runtime_error(std::string("opencv shim: cv2.") + name + " differs from the shim's value");
    };
    same("THRESH_BINARY", cv::THRESH_BINARY);
    same("THRESH_OTSU", cv::THRESH_OTSU);
    same("MORPH_OPEN", cv::MORPH_OPEN);
    same("RETR_EXTERNAL", cv::RETR_EXTERNAL);
    same("RETR_TREE", cv::RETR_TREE);
    same("CHAIN_APPROX_NONE", cv::CHAIN_APPROX_NONE);
    same("CHAIN_APPROX_SIMPLE", cv::CHAIN_APPROX_SIMPLE);
    same("FLOODFILL_FIXED_RANGE", cv::FLOODFILL_FIXED_RANGE);
    same("LINE_8", cv::LINE_8);
    same("CV_8U", CV_8U);
    same("CV_16S", CV_16S);
    m.attr("opencv_version") = cv2.attr("__version__");

    py::class_<RefHighlight>(m, "RefHighlight")
        .def(py::init<const py::array &, const py::array &, int, int, int, int, int, int>(), py::arg("background"),
             py::arg("struct_element"), py::arg("threshold"), py::arg("threshold_lo"), py::arg("threshold_hi"),
             py::arg("min_size_hyst"), py::arg("min_size_threshold"), py::arg("width_border"))
        .def("insert", &RefHighlight::insert)
        .def("notify_no_more_tokens", &RefHighlight::notify_no_more_tokens)
        .def("has_results", &RefHighlight::has_results)
        .def("threshold_image", &RefHighlight::threshold_image)
        .def("threshold_image_with_hysteresis", &RefHighlight::threshold_image_with_hysteresis)
        .def("remove_small_objects", &RefHighlight::remove_small_objects)
        .def("fill_holes", &RefHighlight::fill_holes)
        .def("frame_and_fill", &RefHighlight::frame_and_fill);
}
