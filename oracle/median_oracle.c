/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the temporal-median hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (cvvidproc_b200/csrc, libcvvp_cuda.so)
 * never links, imports or calls it.
 *
 * This is a plain-C restatement of the reference's HistogramMedianAlgo<T>
 *   /root/reference/Sources/ProcessorAlgos/histogram_median_algo.h
 *     ConsumeVector           :116-141   (saturating per-element histogram increment)
 *     MedianFromHistograms    :144-193   (cumulative scan, "> cap/2" rule, backtrack loop)
 *   bin-type choice           /root/reference/Sources/cv_vid_bg_helpers.cpp:232-253
 *   strip-per-worker threading /root/reference/Sources/cv_vid_bg_helpers.cpp:105-117,
 *                              ProcessorTokenHandlers/cv_vid_frames_generator_algo.h:159-164
 *
 * Parity pin: the reference ships no golden vectors, so this restatement
 * is pinned against the reference's own class compiled from /root/reference
 * (oracle/_ref/libcvvp_median_ref.so, recipe in oracle/Makefile) by
 * tests/test_oracle_median.py, and against the committed fixtures under tests/golden/.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CVVP_ORACLE_EXPORT __attribute__((visibility("default")))

/* One worker = one HistogramMedianAlgo<T> instance consuming a contiguous element range
 * [e0, e1) of every frame (the reference hands each worker one strip of every frame;
 * because the algorithm is element-wise the result is independent of the strip geometry,
 * histogram_median_algo.h:133-140). */
typedef struct {
    const uint8_t *frames;
    size_t nframes;
    size_t frame_pitch;
    size_t e0, e1;
    int bin_bytes;
    uint8_t *out;
    int rc;
} strip_job;

/* histogram layout [bin][element], exactly as m_histograms (histogram_median_algo.h:123-126) */
#define DEFINE_STRIP(T, NAME)                                                                   \
    static int NAME(const strip_job *job)                                                       \
    {                                                                                           \
        const size_t n = job->e1 - job->e0;                                                     \
        if (n == 0)                                                                             \
            return 0;                                                                           \
        T *hist = (T *)calloc((size_t)256 * n, sizeof(T));                                      \
        if (!hist)                                                                              \
            return -1;                                                                          \
        const T tmax = (T)(-1);                                                                 \
        /* ConsumeVector :133-140 -- only increment if it will not roll over */                 \
        for (size_t f = 0; f < job->nframes; ++f) {                                             \
            const uint8_t *src = job->frames + f * job->frame_pitch + job->e0;                  \
            for (size_t e = 0; e < n; ++e) {                                                    \
                T *slot = &hist[(size_t)src[e] * n + e];                                        \
                if (*slot != tmax)                                                              \
                    (*slot)++;                                                                  \
            }                                                                                   \
        }                                                                                       \
        /* MedianFromHistograms :152-190 */                                                     \
        const unsigned long cap = (unsigned long)job->nframes;                                  \
        for (size_t e = 0; e < n; ++e) {                                                        \
            unsigned long acc = 0;                                                              \
            size_t halfway = 255;                                                               \
            for (size_t b = 0; b < 256; ++b) {                                                  \
                acc += (unsigned long)hist[b * n + e];                                          \
                if (halfway == 255 && acc > cap / 2)                                            \
                    halfway = b;                                                                \
            }                                                                                   \
            if (acc != cap) { /* :169-184 saturation backtrack */                               \
                const unsigned long temp_cap = acc;                                             \
                for (size_t b = halfway; b != (size_t)-1; --b) {                                \
                    acc -= (unsigned long)hist[b * n + e];                                      \
                    if (acc < temp_cap / 2)                                                     \
                        break;                                                                  \
                    halfway--;                                                                  \
                }                                                                               \
            }                                                                                   \
            job->out[job->e0 + e] = (uint8_t)halfway;                                           \
        }                                                                                       \
        free(hist);                                                                             \
        return 0;                                                                               \
    }

DEFINE_STRIP(uint8_t, strip_u8)
DEFINE_STRIP(uint16_t, strip_u16)
DEFINE_STRIP(uint32_t, strip_u32)

static void *strip_thread(void *arg)
{
    strip_job *job = (strip_job *)arg;
    switch (job->bin_bytes) {
    case 1: job->rc = strip_u8(job); break;
    case 2: job->rc = strip_u16(job); break;
    case 4: job->rc = strip_u32(job); break;
    default: job->rc = -2; break;
    }
    return NULL;
}

/* bin-type choice of GetVideoBackground (cv_vid_bg_helpers.cpp:237-247) */
CVVP_ORACLE_EXPORT int cvvp_oracle_bin_bytes_for(long long nframes)
{
    if (nframes <= 255)
        return 1;
    if (nframes <= 65535)
        return 2;
    if (nframes <= 4294967295LL)
        return 4;
    return 0;
}

/*
 * frames      : nframes frames, frame f starts at frames + f*frame_pitch, nelem bytes each
 *               (nelem = rows*cols*channels; the median is element-wise, cv_util.cpp:251-254)
 * bin_bytes   : 1, 2 or 4 (HistogramMedianAlgo8/16/32); 0 = choose like GetVideoBackground
 * nthreads    : worker count; each worker owns one contiguous element range of every frame
 * returns 0 on success.
 */
CVVP_ORACLE_EXPORT int cvvp_oracle_median(const uint8_t *frames, size_t nframes, size_t nelem,
                                          size_t frame_pitch, int bin_bytes, int nthreads,
                                          uint8_t *out)
{
    if (!frames || !out || nelem == 0)
        return -3;
    if (bin_bytes == 0)
        bin_bytes = cvvp_oracle_bin_bytes_for((long long)nframes);
    if (bin_bytes != 1 && bin_bytes != 2 && bin_bytes != 4)
        return -2;
    if (nthreads < 1)
        nthreads = 1;
    if ((size_t)nthreads > nelem)
        nthreads = (int)nelem;

    strip_job *jobs = (strip_job *)calloc((size_t)nthreads, sizeof(strip_job));
    pthread_t *tids = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    if (!jobs || !tids) {
        free(jobs);
        free(tids);
        return -1;
    }
    /* like get_bordered_chunks (cv_util.cpp:56-134): equal ranges, the last takes the remainder */
    const size_t base = nelem / (size_t)nthreads;
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].frames = frames;
        jobs[t].nframes = nframes;
        jobs[t].frame_pitch = frame_pitch;
        jobs[t].e0 = base * (size_t)t;
        jobs[t].e1 = (t == nthreads - 1) ? nelem : base * (size_t)(t + 1);
        jobs[t].bin_bytes = bin_bytes;
        jobs[t].out = out;
    }
    int rc = 0;
    if (nthreads == 1) {
        strip_thread(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t)
            pthread_create(&tids[t], NULL, strip_thread, &jobs[t]);
        for (int t = 0; t < nthreads; ++t)
            pthread_join(tids[t], NULL);
    }
    for (int t = 0; t < nthreads; ++t)
        if (jobs[t].rc)
            rc = jobs[t].rc;
    free(jobs);
    free(tids);
    return rc;
}
