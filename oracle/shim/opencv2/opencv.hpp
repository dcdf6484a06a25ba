// TEST INFRASTRUCTURE ONLY.
// Minimal stand-in for <opencv2/opencv.hpp> so that the reference's own
// Sources/ProcessorAlgos/histogram_median_algo.h can be compiled UNMODIFIED from
// /root/reference (OpenCV C++ is not installed in this image; the header uses cv::Mat only as
// a byte container).  This is not OpenCV and not reference code: it implements just the
// members that header (and oracle/median_ref_driver.cpp) touch.
#ifndef CVVP_ORACLE_OPENCV_SHIM_HPP
#define CVVP_ORACLE_OPENCV_SHIM_HPP

#include <cassert>
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)

namespace cv
{
class Mat
{
public:
    int rows{0};
    int cols{0};
    unsigned char *data{nullptr};

    Mat() = default;
    Mat(int rows_, int cols_, int type_) : rows{rows_}, cols{cols_}, m_cn{(type_ >> CV_CN_SHIFT) + 1}
    {
        m_store = std::make_shared<std::vector<unsigned char>>(
            static_cast<std::size_t>(rows) * static_cast<std::size_t>(cols) * static_cast<std::size_t>(m_cn));
        data = m_store->data();
    }
    // column-vector Mat from a std::vector<uchar> (the only overload the reference uses)
    Mat(const std::vector<unsigned char> &vec, bool copy_data) : rows{static_cast<int>(vec.size())}, cols{1}, m_cn{1}
    {
        (void)copy_data; // the shim always copies
        m_store = std::make_shared<std::vector<unsigned char>>(vec);
        data = m_store->data();
    }

    int channels() const { return m_cn; }
    int type() const { return CV_MAKETYPE(CV_8U, m_cn); }
    bool empty() const { return data == nullptr || total() == 0; }
    bool isContinuous() const { return true; }
    std::size_t total() const { return static_cast<std::size_t>(rows) * static_cast<std::size_t>(cols); }

    Mat clone() const
    {
        Mat out{};
        out.rows = rows;
        out.cols = cols;
        out.m_cn = m_cn;
        if (m_store) {
            out.m_store = std::make_shared<std::vector<unsigned char>>(*m_store);
            out.data = out.m_store->data();
        }
        return out;
    }

    // same bytes, new channel count / row count (columns inferred)
    Mat reshape(int cn, int new_rows) const
    {
        Mat out{*this};
        const std::size_t bytes = total() * static_cast<std::size_t>(m_cn);
        if (cn <= 0)
            cn = m_cn;
        if (new_rows <= 0)
            new_rows = rows;
        assert(bytes % (static_cast<std::size_t>(cn) * static_cast<std::size_t>(new_rows)) == 0);
        out.m_cn = cn;
        out.rows = new_rows;
        out.cols = static_cast<int>(bytes / (static_cast<std::size_t>(cn) * static_cast<std::size_t>(new_rows)));
        return out;
    }

    template <typename T>
    T *ptr(int r)
    {
        return reinterpret_cast<T *>(data + static_cast<std::size_t>(r) * cols * m_cn);
    }
    template <typename T>
    const T *ptr(int r) const
    {
        return reinterpret_cast<const T *>(data + static_cast<std::size_t>(r) * cols * m_cn);
    }

private:
    int m_cn{1};
    std::shared_ptr<std::vector<unsigned char>> m_store{};
};
} // namespace cv

#endif
