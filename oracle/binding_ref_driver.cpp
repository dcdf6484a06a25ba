// TEST INFRASTRUCTURE ONLY.
// Python module `cvvp_binding_ref`: the reference-side bindings of INTEGRATION.md (integration/gpu_median_algo.h,
// integration/gpu_highlight_algo.h) compiled against the REFERENCE's own token_processor_algo.h
// (/root/reference/Sources/AsyncTokens) and the cv2-forwarding cv::Mat of oracle/shim_cv2, linked with
// libcvvp_cuda.so.  Every call below goes through a TokenProcessorAlgo<..., cv::Mat, cv::Mat> base reference -- the
// reference's plugin interface -- the way its worker thread calls an algo
// (Sources/AsyncTokens/token_processing_unit.h:293 Insert, :307 TryGetResult, :334 NotifyNoMoreTokens).
// It proves that the drop-in boundary is what INTEGRATION.md says; it is not an oracle and computes nothing itself.
// Built only where /root/reference is mounted, into oracle/_ref/ (git-ignored, travels with the snapshot).
#include <memory>
#include <stdexcept>
#include <utility>

#include <opencv2/opencv.hpp> // oracle/shim_cv2

#include "gpu_highlight_algo.h" // integration/
#include "gpu_median_algo.h"

namespace py = pybind11;

namespace
{
cv::Mat mat_copy_of(const py::array &a)
{
    return cv::Mat{py::module_::import("numpy").attr("array")(a, py::arg("copy") = true, py::arg("order") = "C").cast<py::array>()};
}

// what the reference's processing unit does with any algo: tokens in, results out, through the base interface
template <class AlgoT>
py::object feed(TokenProcessorAlgo<AlgoT, cv::Mat, cv::Mat> &algo, const py::object &token)
{
    std::unique_ptr<cv::Mat> t{};
    if (!token.is_none())
        t = std::make_unique<cv::Mat>(mat_copy_of(token.cast<py::array>()));
    algo.Insert(std::move(t));
    if (!algo.HasResults())
        return py::none();
    std::unique_ptr<cv::Mat> r = algo.TryGetResult();
    if (!r)
        throw std::runtime_error("HasResults() was true but TryGetResult() returned nothing");
    return r->array();
}

class BoundMedian
{
public:
    BoundMedian(int device, long long frames_hint) : m_algo{TokenProcessorPack<GpuMedianAlgo>{device, frames_hint}} {}
    py::object insert(const py::object &token) { return feed<GpuMedianAlgo>(m_algo, token); }
    py::object finish()
    {
        TokenProcessorAlgo<GpuMedianAlgo, cv::Mat, cv::Mat> &algo = m_algo;
        algo.NotifyNoMoreTokens();
        if (!algo.HasResults())
            return py::none();
        return algo.TryGetResult()->array();
    }

private:
    GpuMedianAlgo m_algo;
};

class BoundHighlight
{
public:
    BoundHighlight(const py::array &background, const py::array &struct_element, int threshold, int threshold_lo,
                   int threshold_hi, int min_size_hyst, int min_size_threshold, int width_border, int device)
        : m_algo{TokenProcessorPack<GpuHighlightAlgo>{mat_copy_of(background), mat_copy_of(struct_element), threshold,
                                                      threshold_lo, threshold_hi, min_size_hyst, min_size_threshold,
                                                      width_border, device}}
    {
    }
    py::object insert(const py::object &token) { return feed<GpuHighlightAlgo>(m_algo, token); }

private:
    GpuHighlightAlgo m_algo;
};
} // namespace

PYBIND11_MODULE(cvvp_binding_ref, m)
{
    m.doc() = "TEST INFRASTRUCTURE ONLY: INTEGRATION.md's TokenProcessorAlgo bindings compiled against the reference's headers";
    py::class_<BoundMedian>(m, "GpuMedianAlgo")
        .def(py::init<int, long long>(), py::arg("device") = 0, py::arg("frames_hint") = -1)
        .def("insert", &BoundMedian::insert)
        .def("finish", &BoundMedian::finish);
    py::class_<BoundHighlight>(m, "GpuHighlightAlgo")
        .def(py::init<const py::array &, const py::array &, int, int, int, int, int, int, int>(), py::arg("background"),
             py::arg("struct_element"), py::arg("threshold"), py::arg("threshold_lo"), py::arg("threshold_hi"),
             py::arg("min_size_hyst"), py::arg("min_size_threshold"), py::arg("width_border"), py::arg("device") = 0)
        .def("insert", &BoundHighlight::insert);
}
