// TEST INFRASTRUCTURE ONLY.
// Stand-in for <opencv2/opencv.hpp> that lets the reference's own
//     /root/reference/Sources/ProcessorAlgos/highlight_objects_algo.{h,cpp}
//     /root/reference/Sources/ProcessorTokenHandlers/cv_vid_frames_generator_algo.h
//     /root/reference/Sources/cv_vid_bg_helpers.cpp, Sources/Utility/cv_util.cpp (+ the AsyncTokens headers and
//     histogram_median_algo.h they drive)
// be compiled UNMODIFIED where OpenCV's C++ headers and libraries are absent (this image): every cv:: function the
// file calls is forwarded to the function of the same name in the Python wheel `cv2` (OpenCV 4.13), which IS
// installed and runs OpenCV's real core/imgproc code.  cv::Mat wraps a numpy array; `rows`, `cols` and `data` are
// kept as public fields because the reference reads them directly (highlight_objects_algo.h:63, .cpp:26, :201-208).
//
// This is not OpenCV and not reference code.  What it must get right, and how:
//   * `Mat = Mat - Mat` (.cpp:27): a MatExpr assigned to a Mat evaluates cv::subtract(a, b, dst) with dtype -1, i.e.
//     the destination takes the operands' depth (CV_8U, saturating) whatever it was declared as  -> cv2.subtract.
//   * OutputArray: Mat::create() keeps the buffer when size and type already match, else allocates  -> shim::output().
//   * a Mat copy shares its pixels (reference-counted header)  -> the copies share one numpy array.
//   * cv::drawContours with an empty contour list is a no-op in C++; cv2's binding rejects the empty list -> skipped.
//   * cv::VideoCapture is cv2.VideoCapture (same FFmpeg backend, same property numbers); `vid >> frame` leaves an
//     empty Mat at the end of the stream.
//   * threads: the reference's pipelines (AsyncTokens) create, copy and drop Mats on worker threads.  Every function here
//     that talks to Python takes the GIL itself, and a Mat holds its array through shim::PyRef, whose copies and
//     destructor take it too; the fields `rows`, `cols`, `step`, `data` are plain and need none.  A driver that calls
//     into threaded reference code releases the GIL around the call.
// Built into oracle/_ref/cvvp_{highlight,frames,background,binding}_ref*.so by oracle/Makefile (targets ref_highlight,
// ref_frames, ref_background, ref_binding); used by tests/ and by the tests/golden/make_*_golden.py scripts only.
#ifndef CVVP_ORACLE_OPENCV_CV2_SHIM_HPP
#define CVVP_ORACLE_OPENCV_CV2_SHIM_HPP

#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cassert>
#include <climits>
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream> // the real header brings it in; cv_vid_frames_generator_algo.h:164 relies on that
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(type) ((type) & ((1 << CV_CN_SHIFT) - 1))
#define CV_MAT_CN(type) (((type) >> CV_CN_SHIFT) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)

namespace cv
{
namespace py = pybind11;
typedef unsigned char uchar;
typedef std::string String;
class CommandLineParser; // only named in a declaration of the reference's main.h

enum ThresholdTypes { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3, THRESH_TOZERO_INV = 4,
                      THRESH_OTSU = 8 };
enum MorphTypes { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
enum RetrievalModes { RETR_EXTERNAL = 0, RETR_LIST = 1, RETR_CCOMP = 2, RETR_TREE = 3 };
enum ContourApproximationModes { CHAIN_APPROX_NONE = 1, CHAIN_APPROX_SIMPLE = 2 };
enum FloodFillFlags { FLOODFILL_FIXED_RANGE = 1 << 16, FLOODFILL_MASK_ONLY = 1 << 17 };
enum LineTypes { FILLED = -1, LINE_4 = 4, LINE_8 = 8 };
enum VideoCaptureProperties { CAP_PROP_POS_MSEC = 0, CAP_PROP_POS_FRAMES = 1, CAP_PROP_FRAME_WIDTH = 3, CAP_PROP_FRAME_HEIGHT = 4,
                              CAP_PROP_FPS = 5, CAP_PROP_FOURCC = 6, CAP_PROP_FRAME_COUNT = 7, CAP_PROP_FORMAT = 8,
                              CAP_PROP_CONVERT_RGB = 16 };
enum ColorConversionCodes { COLOR_BGR2GRAY = 6, COLOR_RGB2GRAY = 7 };

struct Size {
    int width{0}, height{0};
    Size() = default;
    Size(int w, int h) : width{w}, height{h} {}
};
struct Point {
    int x{0}, y{0};
    Point() = default;
    Point(int x_, int y_) : x{x_}, y{y_} {}
};
struct Rect {
    int x{0}, y{0}, width{0}, height{0};
    Rect() = default;
    Rect(int x_, int y_, int w, int h) : x{x_}, y{y_}, width{w}, height{h} {}
};
struct Scalar {
    double val[4]{0, 0, 0, 0};
    Scalar() = default;
    Scalar(double v0) : val{v0, 0, 0, 0} {}
    Scalar(double v0, double v1, double v2 = 0, double v3 = 0) : val{v0, v1, v2, v3} {}
};

namespace shim
{
// Takes the GIL for the current scope; a no-op where the caller already holds it.  A thread the reference created keeps
// ONE Python thread state for its whole life (made on its first call, dropped when the thread ends) and only hands the
// GIL back between calls -- creating and deleting a thread state around every call, as PyGILState_Ensure / Release or
// pybind11's gil_scoped_acquire do on such threads, is what the reference's worker threads would otherwise do thousands
// of times per video.
class Gil
{
public:
    Gil()
    {
        if (PyGILState_Check())
            return;
        Slot &t = slot();
        if (!t.made) {
            t.state = PyGILState_Ensure();
            t.made = true;
        } else {
            PyEval_RestoreThread(t.saved);
        }
        m_taken = true;
    }
    ~Gil()
    {
        if (m_taken)
            slot().saved = PyEval_SaveThread();
    }
    Gil(const Gil &) = delete;
    Gil &operator=(const Gil &) = delete;

private:
    struct Slot {
        bool made{false};
        PyGILState_STATE state{};
        PyThreadState *saved{nullptr};
        ~Slot()
        {
            if (made && saved && Py_IsInitialized()) {
                PyEval_RestoreThread(saved);
                PyGILState_Release(state);
            }
        }
    };
    static Slot &slot()
    {
        thread_local Slot s;
        return s;
    }
    bool m_taken{false};
};

// A Python reference that may be copied, moved and dropped on any thread.
class PyRef
{
public:
    PyRef() = default;
    explicit PyRef(py::object o)
    {
        Gil gil;
        m_p = o.release().ptr();
    }
    PyRef(const PyRef &o)
    {
        if (o.m_p) {
            Gil gil;
            m_p = o.m_p;
            Py_INCREF(m_p);
        }
    }
    PyRef(PyRef &&o) noexcept : m_p{o.m_p} { o.m_p = nullptr; }
    PyRef &operator=(const PyRef &o)
    {
        if (this != &o) {
            PyRef tmp{o};
            std::swap(m_p, tmp.m_p);
        }
        return *this;
    }
    PyRef &operator=(PyRef &&o) noexcept
    {
        std::swap(m_p, o.m_p);
        return *this;
    }
    ~PyRef()
    {
        if (m_p) {
            Gil gil;
            Py_DECREF(m_p);
        }
    }
    bool empty() const { return m_p == nullptr; }
    py::object get() const { return m_p ? py::reinterpret_borrow<py::object>(m_p) : py::object(py::none()); } // GIL held

private:
    PyObject *m_p{nullptr};
};

inline py::module_ cv2() { return py::module_::import("cv2"); }
inline py::module_ np() { return py::module_::import("numpy"); }
inline py::tuple tup(const Scalar &s) { return py::make_tuple(s.val[0], s.val[1], s.val[2], s.val[3]); }
inline py::tuple tup(const Point &p) { return py::make_tuple(p.x, p.y); }
inline const char *dtype_name(int depth)
{
    static const char *names[] = {"uint8", "int8", "uint16", "int16", "int32", "float32", "float64"};
    if (depth < 0 || depth > CV_64F)
        throw std::invalid_argument("opencv shim: unsupported depth");
    return names[depth];
}
inline int depth_of(const py::array &a)
{
    const std::string n = py::str(a.dtype().attr("name"));
    for (int d = 0; d <= CV_64F; ++d)
        if (n == dtype_name(d))
            return d;
    throw std::invalid_argument("opencv shim: unsupported dtype " + n);
}
} // namespace shim

class Mat;
struct MatExpr { // only `a - b` is needed (.cpp:27)
    const Mat &a;
    const Mat &b;
};

class Mat
{
public:
    int rows{0};
    int cols{0};
    std::size_t step{0}; // bytes from one row to the next
    uchar *data{nullptr};

    Mat() = default;
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int rows_, int cols_, int type) { create(rows_, cols_, type); }
    Mat(int rows_, int cols_, int type, const Scalar &s)
    {
        create(rows_, cols_, type);
        *this = s;
    }
    explicit Mat(py::array a) { adopt(std::move(a)); }
    // a column vector over the bytes of a std::vector (cv_util.cpp:291; the shim always copies)
    Mat(const std::vector<uchar> &vec, bool copy_data)
    {
        (void)copy_data;
        shim::Gil gil;
        py::array_t<uchar> a({py::ssize_t(vec.size()), py::ssize_t(1)});
        if (!vec.empty())
            std::memcpy(a.mutable_data(), vec.data(), vec.size());
        adopt(std::move(a));
    }
    Mat(const Mat &) = default; // shares the pixels, like a reference-counted cv::Mat header
    Mat(Mat &&o) noexcept : rows{o.rows}, cols{o.cols}, step{o.step}, data{o.data}, m_arr{std::move(o.m_arr)} { o.forget(); }
    Mat &operator=(const Mat &) = default;
    Mat &operator=(Mat &&o) noexcept
    {
        if (this != &o) {
            rows = o.rows;
            cols = o.cols;
            step = o.step;
            data = o.data;
            m_arr = std::move(o.m_arr);
            o.m_arr = shim::PyRef{};
            o.forget();
        }
        return *this;
    }
    inline Mat &operator=(const MatExpr &e);
    Mat &operator=(const Scalar &s) // setTo over every channel
    {
        if (has_array()) {
            shim::Gil gil;
            if (channels() == 1)
                array().attr("fill")(s.val[0]);
            else
                array()[py::ellipsis()] = shim::tup(s)[py::slice(0, channels(), 1)];
        }
        return *this;
    }

    void create(int rows_, int cols_, int type)
    {
        shim::Gil gil;
        const int cn = CV_MAT_CN(type);
        py::tuple shape = cn == 1 ? py::tuple(py::make_tuple(rows_, cols_)) : py::tuple(py::make_tuple(rows_, cols_, cn));
        adopt(shim::np().attr("zeros")(shape, shim::dtype_name(CV_MAT_DEPTH(type))).cast<py::array>());
    }
    void adopt(py::array a)
    {
        shim::Gil gil;
        if (a.ndim() != 2 && a.ndim() != 3)
            throw std::invalid_argument("opencv shim: a Mat wraps a 2-D or 3-D array");
        rows = int(a.shape(0));
        cols = int(a.shape(1));
        step = std::size_t(a.strides(0));
        data = static_cast<uchar *>(a.mutable_data());
        m_arr = shim::PyRef{std::move(a)};
    }
    py::array array() const { return has_array() ? m_arr.get().cast<py::array>() : py::array(); } // GIL held by the caller
    bool has_array() const { return !m_arr.empty(); }

    int channels() const
    {
        if (!has_array())
            return 1;
        shim::Gil gil;
        py::array a = array();
        return a.ndim() == 3 ? int(a.shape(2)) : 1;
    }
    int depth() const
    {
        if (!has_array())
            return CV_8U;
        shim::Gil gil;
        return shim::depth_of(array());
    }
    int type() const { return CV_MAKETYPE(depth(), channels()); }
    bool empty() const { return data == nullptr || total() == 0; }
    bool isContinuous() const
    {
        if (!has_array())
            return true;
        shim::Gil gil;
        return array().attr("flags").attr("c_contiguous").cast<bool>();
    }
    std::size_t total() const { return std::size_t(rows) * std::size_t(cols); }
    Size size() const { return Size{cols, rows}; }
    template <typename T>
    T *ptr(int r)
    {
        return reinterpret_cast<T *>(data + std::size_t(r) * step);
    }
    template <typename T>
    const T *ptr(int r) const
    {
        return reinterpret_cast<const T *>(data + std::size_t(r) * step);
    }

    Mat clone() const
    {
        if (!has_array())
            return Mat{};
        shim::Gil gil;
        return Mat{array().attr("copy")().cast<py::array>()};
    }
    // same bytes, new channel count and row count (columns inferred), cv_util.cpp:291
    Mat reshape(int cn, int new_rows = 0) const
    {
        if (!has_array())
            return Mat{};
        shim::Gil gil;
        py::array a = array();
        const int old_cn = a.ndim() == 3 ? int(a.shape(2)) : 1;
        if (cn <= 0)
            cn = old_cn;
        if (new_rows <= 0)
            new_rows = rows;
        const std::size_t elems = total() * std::size_t(old_cn);
        if (elems % (std::size_t(cn) * std::size_t(new_rows)) != 0)
            throw std::invalid_argument("opencv shim: reshape does not divide the elements");
        const py::ssize_t new_cols = py::ssize_t(elems / (std::size_t(cn) * std::size_t(new_rows)));
        py::tuple shape = cn == 1 ? py::tuple(py::make_tuple(new_rows, new_cols)) : py::tuple(py::make_tuple(new_rows, new_cols, cn));
        return Mat{a.attr("reshape")(shape).cast<py::array>()};
    }
    inline void convertTo(Mat &dst, int rtype) const;
    void copyTo(const Mat &dst) const // destination of the same size (a region of interest): pixels are copied into it
    {
        if (!dst.has_array() || dst.rows != rows || dst.cols != cols)
            throw std::invalid_argument("opencv shim: copyTo needs an allocated destination of the same size");
        shim::Gil gil;
        shim::np().attr("copyto")(dst.array(), array());
    }
    Mat operator()(const Rect &r) const // a view of the same pixels
    {
        if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > cols || r.y + r.height > rows)
            throw std::out_of_range("opencv shim: region of interest outside the Mat");
        shim::Gil gil;
        return Mat{array()[py::make_tuple(py::slice(r.y, r.y + r.height, 1), py::slice(r.x, r.x + r.width, 1))].cast<py::array>()};
    }

private:
    void forget()
    {
        rows = cols = 0;
        step = 0;
        data = nullptr;
    }
    shim::PyRef m_arr{};
};

namespace shim
{
// what an OutputArray does with a result: Mat::create() keeps a buffer of the right size and type and the function
// writes into it; otherwise the destination header gets a new buffer  (GIL held by the caller)
inline void output(Mat &dst, py::object result)
{
    py::array r = result.cast<py::array>();
    if (dst.has_array()) {
        py::array d = dst.array();
        if (d.is(r))
            return;
        bool same = d.ndim() == r.ndim() && d.dtype().is(r.dtype());
        for (py::ssize_t k = 0; same && k < d.ndim(); ++k)
            same = d.shape(k) == r.shape(k);
        if (same) {
            np().attr("copyto")(d, r);
            return;
        }
    }
    dst.adopt(std::move(r));
}
inline py::array contour_array(const std::vector<Point> &c)
{
    py::array_t<std::int32_t> a({py::ssize_t(c.size()), py::ssize_t(1), py::ssize_t(2)});
    auto w = a.mutable_unchecked<3>();
    for (std::size_t i = 0; i < c.size(); ++i) {
        w(py::ssize_t(i), 0, 0) = c[i].x;
        w(py::ssize_t(i), 0, 1) = c[i].y;
    }
    return std::move(a);
}
} // namespace shim

inline MatExpr operator-(const Mat &a, const Mat &b) { return MatExpr{a, b}; }

inline Mat &Mat::operator=(const MatExpr &e)
{
    shim::Gil gil;
    shim::output(*this, shim::cv2().attr("subtract")(e.a.array(), e.b.array())); // dtype = -1: the operands' depth
    return *this;
}

inline void Mat::convertTo(Mat &dst, int rtype) const
{
    shim::Gil gil;
    const int d = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
    py::object src = array();
    py::object out;
    if (d == depth()) {
        out = src.attr("copy")();
    } else if (d >= CV_32F) {
        out = src.attr("astype")(shim::dtype_name(d));
    } else { // saturate_cast: round to nearest even, clamp to the target's range
        py::module_ np = shim::np();
        py::object info = np.attr("iinfo")(shim::dtype_name(d));
        out = np.attr("clip")(np.attr("rint")(src), info.attr("min"), info.attr("max")).attr("astype")(shim::dtype_name(d));
    }
    shim::output(dst, out);
}

inline std::string format(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    return std::string{buf};
}

inline double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type)
{
    shim::Gil gil;
    py::tuple r = shim::cv2().attr("threshold")(src.array(), thresh, maxval, type);
    shim::output(dst, r[1]);
    return r[0].cast<double>();
}

inline void morphologyEx(const Mat &src, Mat &dst, int op, const Mat &kernel)
{
    // defaults of the C++ signature: anchor (-1,-1), 1 iteration, BORDER_CONSTANT with morphologyDefaultBorderValue()
    shim::Gil gil;
    shim::output(dst, shim::cv2().attr("morphologyEx")(src.array(), op, kernel.array()));
}

inline void findContours(const Mat &image, std::vector<std::vector<Point>> &contours, int mode, int method)
{
    shim::Gil gil;
    py::tuple r = shim::cv2().attr("findContours")(image.array(), mode, method);
    contours.clear();
    for (py::handle h : r[0]) {
        py::array_t<std::int32_t, py::array::c_style | py::array::forcecast> a = py::reinterpret_borrow<py::object>(h);
        auto v = a.unchecked<3>();
        std::vector<Point> c;
        c.reserve(std::size_t(v.shape(0)));
        for (py::ssize_t i = 0; i < v.shape(0); ++i)
            c.emplace_back(v(i, 0, 0), v(i, 0, 1));
        contours.push_back(std::move(c));
    }
}

inline double contourArea(const std::vector<Point> &contour, bool oriented = false)
{
    shim::Gil gil;
    return shim::cv2().attr("contourArea")(shim::contour_array(contour), oriented).cast<double>();
}

inline void drawContours(Mat &image, const std::vector<std::vector<Point>> &contours, int contourIdx, const Scalar &color,
                         int thickness = 1, int lineType = LINE_8)
{
    if (contours.empty())
        return; // C++ loops over nothing; the Python binding refuses an empty list
    shim::Gil gil;
    py::list cs;
    for (const auto &c : contours)
        cs.append(shim::contour_array(c));
    shim::output(image, shim::cv2().attr("drawContours")(image.array(), cs, contourIdx, shim::tup(color), thickness, lineType));
}

inline int floodFill(Mat &image, Point seedPoint, Scalar newVal, Rect *rect = nullptr, Scalar loDiff = Scalar(),
                     Scalar upDiff = Scalar(), int flags = 4)
{
    shim::Gil gil;
    py::tuple r = shim::cv2().attr("floodFill")(image.array(), py::none(), shim::tup(seedPoint), shim::tup(newVal),
                                                 shim::tup(loDiff), shim::tup(upDiff), flags);
    shim::output(image, r[1]);
    if (rect) {
        py::tuple b = r[3];
        *rect = Rect{b[0].cast<int>(), b[1].cast<int>(), b[2].cast<int>(), b[3].cast<int>()};
    }
    return r[0].cast<int>();
}

inline void extractChannel(const Mat &src, Mat &dst, int coi)
{
    shim::Gil gil;
    shim::output(dst, shim::cv2().attr("extractChannel")(src.array(), coi));
}

inline void cvtColor(const Mat &src, Mat &dst, int code)
{
    shim::Gil gil;
    shim::output(dst, shim::cv2().attr("cvtColor")(src.array(), code));
}

class VideoCapture
{
public:
    VideoCapture() = default;
    explicit VideoCapture(const std::string &filename)
    {
        shim::Gil gil;
        m_cap = shim::PyRef{shim::cv2().attr("VideoCapture")(filename)};
    }
    bool isOpened() const
    {
        if (m_cap.empty())
            return false;
        shim::Gil gil;
        return m_cap.get().attr("isOpened")().cast<bool>();
    }
    double get(int prop) const
    {
        if (m_cap.empty())
            return 0.0;
        shim::Gil gil;
        return m_cap.get().attr("get")(prop).cast<double>();
    }
    bool set(int prop, double value)
    {
        if (m_cap.empty())
            return false;
        shim::Gil gil;
        return m_cap.get().attr("set")(prop, value).cast<bool>();
    }
    bool read(Mat &image)
    {
        image = Mat{}; // a failed read releases the destination
        if (m_cap.empty())
            return false;
        shim::Gil gil;
        py::tuple r = m_cap.get().attr("read")();
        if (!r[0].cast<bool>() || r[1].is_none())
            return false;
        image = Mat{r[1].cast<py::array>()};
        return true;
    }
    VideoCapture &operator>>(Mat &image)
    {
        read(image);
        return *this;
    }

private:
    shim::PyRef m_cap{};
};

inline void bitwise_not(const Mat &src, Mat &dst)
{
    shim::Gil gil;
    shim::output(dst, shim::cv2().attr("bitwise_not")(src.array()));
}

inline void bitwise_or(const Mat &a, const Mat &b, Mat &dst)
{
    shim::Gil gil;
    shim::output(dst, shim::cv2().attr("bitwise_or")(a.array(), b.array()));
}
} // namespace cv

#endif
