// Process-wide reuse of the large buffers a job allocates and frees: the queue's pinned slots and device buffers, the
// fused highlight kernel's scratch.  Page-locking a few hundred MB costs 150-250 ms and unlocking it 70 ms and more
// (measured: tools/probe_setup_cost.py), which is what a TrackObjects call on a short clip spent most of its time on;
// a drop-in module is called once per video, many times per process.  Freed buffers are parked (exact size, same
// device) up to a cap and handed to the next job; cvvp_pool_trim() releases everything that is parked.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace cvvp
{
// pinned host memory (usable from every device of the process)
cudaError_t pool_host_alloc(void **p, size_t bytes);
void pool_host_free(void *p, size_t bytes);
// device memory of the CURRENT device
cudaError_t pool_dev_alloc(void **p, size_t bytes);
void pool_dev_free(void *p, size_t bytes);
// release everything that is parked; returns the bytes released (host + device)
size_t pool_trim();
} // namespace cvvp
