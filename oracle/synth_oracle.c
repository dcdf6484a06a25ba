/*
 * TEST INFRASTRUCTURE ONLY -- multi-threaded host generator of the synthetic input stream.
 *
 * Same integer arithmetic as cvvidproc_b200/synth.py (which stays the readable definition and is what
 * tests/test_oracle_median.py holds this file to).  It exists so that bench.py's CPU legs can fill the WHOLE
 * 1920x1080x1000 workload in a few seconds on the GPU box's host cores instead of a row band in minutes; like the rest
 * of oracle/ it is never loaded by the product path.
 *
 *   r(f,y,x) = mix32(seed ^ mix32(f*0x9E3779B1 + (y*W + x)))          mix32 = murmur3 fmix32
 *   B(y,x)   = 140 + (x*20)/W - (y*10)/H ;  noise = (r & 7) - 3
 *   disk k   : a = mix32(seed*1000003 + k), b = mix32(a), c = mix32(b)
 *              cx = (a % W + 2f) % W, cy = b % H, rad = 3 + c % 30, depth = 10 + (c>>8) % 50
 *              core (rad > 8): radius rad/3, adds back 10 + (c>>16) % 50
 *   frame    = clamp(B + noise - sum(depth inside disks) + sum(core add-back), 0, 255)
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>

#define CVVP_ORACLE_EXPORT __attribute__((visibility("default")))
#define MAX_DISKS 64

static inline uint32_t mix32(uint32_t h)
{
    h ^= h >> 16;
    h *= 0x85ebca6bu;
    h ^= h >> 13;
    h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

typedef struct {
    uint8_t *out;
    size_t frame_pitch;
    int width, height, row0, nrows;
    long long first_frame, f0, f1; /* this worker's frames [f0, f1) of the call */
    uint32_t seed;
    int ndisks;
} synth_job;

static void *synth_worker(void *arg)
{
    const synth_job *j = (const synth_job *)arg;
    const int W = j->width, H = j->height;
    int *row = (int *)malloc(sizeof(int) * (size_t)W);
    if (!row)
        return (void *)1;
    for (long long i = j->f0; i < j->f1; ++i) {
        const long long f = j->first_frame + i;
        int cx[MAX_DISKS], cy[MAX_DISKS], rad[MAX_DISKS], depth[MAX_DISKS], core_r[MAX_DISKS], core_add[MAX_DISKS];
        for (int k = 0; k < j->ndisks; ++k) {
            const uint32_t a = mix32(j->seed * 1000003u + (uint32_t)k);
            const uint32_t b = mix32(a);
            const uint32_t c = mix32(b);
            cx[k] = (int)(((uint64_t)(a % (uint32_t)W) + 2ull * (uint64_t)f) % (uint64_t)W);
            cy[k] = (int)(b % (uint32_t)H);
            rad[k] = 3 + (int)(c % 30u);
            depth[k] = 10 + (int)((c >> 8) % 50u);
            if (rad[k] > 8) {
                core_r[k] = rad[k] / 3;
                core_add[k] = 10 + (int)((c >> 16) % 50u);
            } else {
                core_r[k] = -1;
                core_add[k] = 0;
            }
        }
        const uint32_t fterm = (uint32_t)((uint64_t)f) * 0x9E3779B1u;
        for (int ry = 0; ry < j->nrows; ++ry) {
            const int y = j->row0 + ry;
            const int yterm = 140 - (y * 10) / H - 3;
            const uint32_t lin0 = fterm + (uint32_t)y * (uint32_t)W;
            for (int x = 0; x < W; ++x) {
                const uint32_t r = mix32(j->seed ^ mix32(lin0 + (uint32_t)x));
                row[x] = yterm + (x * 20) / W + (int)(r & 7u);
            }
            for (int k = 0; k < j->ndisks; ++k) {
                const int dy = y - cy[k];
                if (dy > rad[k] || dy < -rad[k])
                    continue;
                const int r2 = rad[k] * rad[k], c2 = core_r[k] >= 0 ? core_r[k] * core_r[k] : -1;
                const int x_lo = cx[k] - rad[k] < 0 ? 0 : cx[k] - rad[k];
                const int x_hi = cx[k] + rad[k] > W - 1 ? W - 1 : cx[k] + rad[k];
                for (int x = x_lo; x <= x_hi; ++x) {
                    const int dx = x - cx[k];
                    const int q = dx * dx + dy * dy;
                    if (q <= r2)
                        row[x] -= depth[k];
                    if (q <= c2)
                        row[x] += core_add[k];
                }
            }
            uint8_t *dst = j->out + (size_t)i * j->frame_pitch + (size_t)ry * (size_t)W;
            for (int x = 0; x < W; ++x)
                dst[x] = (uint8_t)(row[x] < 0 ? 0 : (row[x] > 255 ? 255 : row[x]));
        }
    }
    free(row);
    return NULL;
}

/* Rows [row0, row0 + nrows) of frames first_frame .. first_frame + nframes - 1, frame i at out + i * frame_pitch
 * (rows dense, `width` bytes each).  Returns 0, -1 on bad arguments, -2 when a worker failed. */
CVVP_ORACLE_EXPORT int cvvp_oracle_synth_frames(uint8_t *out, size_t frame_pitch, int width, int height, int row0, int nrows,
                                                long long first_frame, long long nframes, uint32_t seed, int ndisks,
                                                int nthreads)
{
    if (!out || width <= 0 || height <= 0 || row0 < 0 || nrows <= 0 || row0 + nrows > height || nframes < 0 ||
        first_frame < 0 || ndisks < 0 || ndisks > MAX_DISKS || frame_pitch < (size_t)width * (size_t)nrows)
        return -1;
    if (nthreads < 1)
        nthreads = 1;
    if (nthreads > 256)
        nthreads = 256;
    if ((long long)nthreads > nframes)
        nthreads = nframes > 0 ? (int)nframes : 1;
    pthread_t tid[256];
    synth_job jobs[256];
    int started = 0, rc = 0;
    for (int t = 0; t < nthreads; ++t) {
        synth_job *j = &jobs[t];
        j->out = out;
        j->frame_pitch = frame_pitch;
        j->width = width;
        j->height = height;
        j->row0 = row0;
        j->nrows = nrows;
        j->first_frame = first_frame;
        j->f0 = nframes * t / nthreads;
        j->f1 = nframes * (t + 1) / nthreads;
        j->seed = seed;
        j->ndisks = ndisks;
        if (pthread_create(&tid[t], NULL, synth_worker, j) != 0) {
            rc = -2;
            break;
        }
        ++started;
    }
    for (int t = 0; t < started; ++t) {
        void *res = NULL;
        pthread_join(tid[t], &res);
        if (res)
            rc = -2;
    }
    return rc;
}
