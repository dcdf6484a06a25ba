// see pool.hpp
#include "pool.hpp"

#include <mutex>
#include <vector>

namespace cvvp
{
namespace
{
struct Parked {
    void *p;
    size_t bytes;
    int device; // -1: pinned host memory
};
std::mutex g_mu;
std::vector<Parked> g_parked;
size_t g_host_bytes = 0, g_dev_bytes = 0;
constexpr size_t kHostCap = size_t(3) << 30;  // parked pinned memory
constexpr size_t kDevCap = size_t(24) << 30;  // parked device memory (all devices together)
constexpr size_t kMinBytes = size_t(1) << 20; // smaller buffers are not worth parking

void *take(size_t bytes, int device)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (size_t i = 0; i < g_parked.size(); ++i) {
        if (g_parked[i].bytes == bytes && g_parked[i].device == device) {
            void *p = g_parked[i].p;
            (device < 0 ? g_host_bytes : g_dev_bytes) -= bytes;
            g_parked[i] = g_parked.back();
            g_parked.pop_back();
            return p;
        }
    }
    return nullptr;
}

bool park(void *p, size_t bytes, int device)
{
    if (bytes < kMinBytes)
        return false;
    std::lock_guard<std::mutex> lk(g_mu);
    size_t &total = device < 0 ? g_host_bytes : g_dev_bytes;
    if (total + bytes > (device < 0 ? kHostCap : kDevCap))
        return false;
    total += bytes;
    g_parked.push_back(Parked{p, bytes, device});
    return true;
}
} // namespace

cudaError_t pool_host_alloc(void **p, size_t bytes)
{
    if ((*p = take(bytes, -1)) != nullptr)
        return cudaSuccess;
    cudaError_t e = cudaMallocHost(p, bytes);
    if (e != cudaSuccess && pool_trim() > 0) { // parked buffers may be what is in the way
        cudaGetLastError();
        e = cudaMallocHost(p, bytes);
    }
    return e;
}

void pool_host_free(void *p, size_t bytes)
{
    if (p && !park(p, bytes, -1))
        cudaFreeHost(p);
}

cudaError_t pool_dev_alloc(void **p, size_t bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if ((*p = take(bytes, dev)) != nullptr)
        return cudaSuccess;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess && pool_trim() > 0) {
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    return e;
}

void pool_dev_free(void *p, size_t bytes)
{
    if (!p)
        return;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!park(p, bytes, dev))
        cudaFree(p);
}

size_t pool_trim()
{
    std::vector<Parked> all;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        all.swap(g_parked);
        g_host_bytes = g_dev_bytes = 0;
    }
    int cur = 0;
    cudaGetDevice(&cur);
    size_t freed = 0;
    for (const Parked &b : all) {
        if (b.device < 0) {
            cudaFreeHost(b.p);
        } else {
            cudaSetDevice(b.device);
            cudaFree(b.p);
        }
        freed += b.bytes;
    }
    cudaSetDevice(cur);
    cudaGetLastError();
    return freed;
}
} // namespace cvvp
