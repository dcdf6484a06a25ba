// Fused per-frame highlight kernel for sm_100a -- the product path of the highlight stage.
// Replaces HighlightObjectsAlgo::HighlightObjects (/root/reference/Sources/ProcessorAlgos/highlight_objects_algo.cpp:17-221);
// semantics and reference citations per step: highlight.cu's header and oracle/highlight_model.py.
//
// ONE kernel launch per batch.  A persistent CTA pulls frames from a queue and takes each frame through every step of
// the reference function with block barriers in between, so a frame costs one read of its bytes and one write of its
// mask in HBM (the algorithmic traffic) plus L2-resident scratch:
//
//   bits     frame, background -> three bit images  A = d > th, U = d > hi, L = d > lo   (d = saturating bg - frame)
//            32 pixels per thread from 128-bit loads; the byte compares are SWAR (four pixels per instruction)
//   open     erode + dilate on bit rows: one funnel shift + AND/OR per structuring-element tap per 32 pixels
//   label    the nodes of the union-find are the horizontal RUNS of equal bits (a few thousand per 1080p frame, not two
//            million pixels), numbered in raster order by a two-pass block scan, so "smallest id of a component" is
//            still its raster-first pixel; foreground 8-/4-connected, background 4-connected, optional FRAME node
//   hyst     seeds = first runs of the external hi components; marked lo regions are scattered back as bits
//   rso      contour statistics from run geometry: run ends give the horizontal cracks and the convex corners, overlaps
//            with the runs of the neighbouring rows give the vertical cracks; the polygon-area rule, the edge rule and
//            the even-odd nesting parity of the single filled drawContours call decide which bits are cleared
//   fill     background runs that are not connected to the seed corner are set (in place)
//   expand   (A | B) -> 0/255 bytes, 128-bit stores
//
// Scratch is per RESIDENT CTA (a "slot"), not per frame of the batch: 4 bit images and 8 run arrays.  Run-level arrays
// that are touched by atomics are only ever read with ld.global.cg (L2), so a stale L1 line can never be observed.
//
// run record: xinfo[id] = x_start | row << 16 | value << 31   (W <= 65535, H <= 32767; larger frames use highlight.cu)
#include "highlight_state.hpp"

#include <climits>
#include <cstdlib>

namespace cvvp
{
namespace
{
constexpr int NT = 512; // threads per CTA
constexpr int NW = NT / 32;
constexpr int kImages = 4;    // A, U (later B), L, Tm
constexpr int kRunArrays = 8; // xinfo0, parent0, xinfo1, parent1, link, st_s, st_e, st_x
constexpr uint32_t kNoLabel = 0xFFFFFFFFu;
constexpr uint32_t kH = 0x80808080u;

struct ThreshSpec {
    uint32_t t4;       // threshold replicated into four bytes (clamped to 0..254)
    uint32_t force_or; // all-ones when the threshold is negative (every pixel passes)
    uint32_t force_and; // zero when the threshold is >= 255 (no pixel passes)
};

struct FusedArgs {
    const uint8_t *frames;
    size_t frame_stride;
    const uint8_t *bg;
    uint8_t *out;
    size_t out_stride;
    const int *th_a; // per-frame threshold of branch A (Otsu), or nullptr -> th
    int th, lo, hi, min_th, min_hyst;
    const short2 *offs;
    int noffs;
    int W, H, WW, WWp;
    int wwp_shift;    // log2(WWp) when WWp is a power of two, else -1
    uint32_t nwords;  // H * WWp
    uint32_t cap;     // run-array capacity per slot
    uint32_t rstride; // rowoff stride
    uint32_t *bits;
    uint32_t *runs;
    uint32_t *rowoff;
    unsigned *queue; // [0] next frame, [1] CTAs finished
    unsigned nframes;
    int fast_io; // W % 32 == 0 and 16-byte aligned pointers / strides
    int debug_stage; // 0 = off; k > 0: write the bit image of intermediate stage k instead of the result (tools/)
};

struct RunSet {
    uint32_t *xinfo;
    uint32_t *parent;
    uint32_t *rowoff; // [H + 1]; rowoff[H] = T
};

struct Shared {
    unsigned frame;
    uint32_t wsum[NW];
    uint32_t T[2];
    int white[2];
};

// ------------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld(const uint32_t *p) { return __ldcg(p); }
__device__ __forceinline__ int ld(const int *p) { return __ldcg(p); }
__device__ __forceinline__ void st(uint32_t *p, uint32_t v) { __stcg(p, v); }
__device__ __forceinline__ void st(int *p, int v) { __stcg(p, v); }

__device__ __forceinline__ uint32_t run_x(uint32_t xi) { return xi & 0xFFFFu; }
__device__ __forceinline__ uint32_t run_y(uint32_t xi) { return (xi >> 16) & 0x7FFFu; }
__device__ __forceinline__ uint32_t run_v(uint32_t xi) { return xi >> 31; }

__device__ __forceinline__ uint32_t valid_mask(const FusedArgs &P, int wx)
{
    const int rem = P.W - 32 * wx;
    return rem >= 32 ? 0xFFFFFFFFu : (rem <= 0 ? 0u : (0xFFFFFFFFu >> (32 - rem)));
}

__device__ __forceinline__ void split(const FusedArgs &P, uint32_t i, int &y, int &wx)
{
    if (P.wwp_shift >= 0) {
        y = int(i >> P.wwp_shift);
        wx = int(i & uint32_t(P.WWp - 1));
    } else {
        y = int(i / uint32_t(P.WWp));
        wx = int(i - uint32_t(y) * uint32_t(P.WWp));
    }
}

__device__ __forceinline__ uint32_t range_mask(int lo, int hi) // bits lo..hi (0 <= lo <= hi <= 31)
{
    return (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
}

__device__ __forceinline__ uint32_t uf_find(const uint32_t *lab, uint32_t x)
{
    uint32_t p = ld(lab + x);
    while (p != x) {
        x = p;
        p = ld(lab + x);
    }
    return x;
}

__device__ __forceinline__ void uf_union(uint32_t *lab, uint32_t a, uint32_t b)
{
    bool done;
    do {
        a = uf_find(lab, a);
        b = uf_find(lab, b);
        if (a < b) {
            const uint32_t old = atomicMin(&lab[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const uint32_t old = atomicMin(&lab[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// largest j in [a, b) with x_start(j) <= x   (a row's first run starts at 0, so it always exists)
__device__ __forceinline__ uint32_t run_at(const uint32_t *xinfo, uint32_t a, uint32_t b, uint32_t x)
{
    uint32_t lo = a, hi = b - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (run_x(xinfo[mid]) <= x)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ uint32_t run_end(const uint32_t *xinfo, uint32_t r, uint32_t row_end, int W)
{
    return (r + 1 < row_end) ? run_x(xinfo[r + 1]) - 1u : uint32_t(W - 1);
}

__device__ __forceinline__ bool bit_at(const uint32_t *img, const FusedArgs &P, int x, int y)
{
    if (x < 0 || y < 0 || x >= P.W || y >= P.H)
        return false;
    return (ld(img + size_t(y) * P.WWp + (x >> 5)) >> (x & 31)) & 1u;
}

// canonical id of the background region of run j: the FRAME region -> T
__device__ __forceinline__ uint32_t bg_region(const uint32_t *parent, uint32_t j, uint32_t T, uint32_t frame_root)
{
    const uint32_t r = ld(parent + j);
    return r == frame_root ? T : r;
}

__device__ __forceinline__ void clear_range(uint32_t *row, int x0, int x1)
{
    for (int w = x0 >> 5; w <= (x1 >> 5); ++w)
        atomicAnd(&row[w], ~range_mask(max(x0, 32 * w) - 32 * w, min(x1, 32 * w + 31) - 32 * w));
}

__device__ __forceinline__ void set_range(uint32_t *row, int x0, int x1)
{
    for (int w = x0 >> 5; w <= (x1 >> 5); ++w)
        atomicOr(&row[w], range_mask(max(x0, 32 * w) - 32 * w, min(x1, 32 * w + 31) - 32 * w));
}

// ------------------------------------------------------------------------------------------------------------------
// bits: d = max(bg - frame, 0); three thresholds -> three bit images          (highlight_objects_algo.cpp:27-29, :100,
//                                                                               :117-122)
// ------------------------------------------------------------------------------------------------------------------
// per-byte saturating x - y (SWAR: no carries between bytes)
__device__ __forceinline__ uint32_t sub_sat_u8x4(uint32_t x, uint32_t y)
{
    const uint32_t t = (x | kH) - (y & ~kH);
    const uint32_t ge = (x & ~y) | (~(x ^ y) & t); // byte MSB: x >= y
    const uint32_t diff = t ^ ((x ^ ~y) & kH);
    uint32_t m; // byte MSB -> 0xFF / 0x00 (prmt's sign-replicate mode; __byte_perm masks that selector bit away)
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(m) : "r"(ge), "r"(0u), "r"(0xba98u));
    return diff & m;
}

// four-bit mask of d > t per byte; d7 = d & 0x7f7f7f7f, th = t4 | 0x80808080
__device__ __forceinline__ uint32_t gt_nibble_top(uint32_t d, uint32_t d7, uint32_t t4, uint32_t th)
{
    const uint32_t t = th - d7;
    const uint32_t ge = (t4 & ~d) | (~(t4 ^ d) & t); // byte MSB: t4 >= d
    return (~ge & kH) * 0x00204081u;                 // bits 28..31 = the four MSBs (lower bits: junk)
}

__device__ __forceinline__ ThreshSpec make_thresh(int t)
{
    ThreshSpec s;
    const int c = t < 0 ? 0 : (t > 254 ? 254 : t);
    s.t4 = uint32_t(c) * 0x01010101u;
    s.force_or = t < 0 ? 0xFFFFFFFFu : 0u;
    s.force_and = t >= 255 ? 0u : 0xFFFFFFFFu;
    return s;
}

__device__ __forceinline__ void bits_of_32(const uint4 &f0, const uint4 &f1, const uint4 &b0, const uint4 &b1,
                                           const ThreshSpec &ta, const ThreshSpec &tu, const ThreshSpec &tl, uint32_t &wa,
                                           uint32_t &wu, uint32_t &wl)
{
    const uint32_t fr[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
    const uint32_t bk[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const uint32_t tha = ta.t4 | kH, thu = tu.t4 | kH, thl = tl.t4 | kH;
    wa = wu = wl = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t d = sub_sat_u8x4(bk[k], fr[k]);
        const uint32_t d7 = d & ~kH;
        wa = (wa >> 4) | (gt_nibble_top(d, d7, ta.t4, tha) & 0xF0000000u);
        wu = (wu >> 4) | (gt_nibble_top(d, d7, tu.t4, thu) & 0xF0000000u);
        wl = (wl >> 4) | (gt_nibble_top(d, d7, tl.t4, thl) & 0xF0000000u);
    }
    wa = (wa | ta.force_or) & ta.force_and;
    wu = (wu | tu.force_or) & tu.force_and;
    wl = (wl | tl.force_or) & tl.force_and;
}

__device__ void bits_phase(const FusedArgs &P, unsigned f, uint32_t *A, uint32_t *U, uint32_t *L)
{
    const int tid = threadIdx.x;
    const uint8_t *fr = P.frames + size_t(f) * P.frame_stride;
    const int th_a = P.th_a ? __ldg(P.th_a + f) : P.th;
    if (P.fast_io) {
        const ThreshSpec ta = make_thresh(th_a), tu = make_thresh(P.hi), tl = make_thresh(P.lo);
        // two words per thread per iteration: eight 128-bit loads in flight
        for (uint32_t i0 = tid; i0 < P.nwords; i0 += 2 * NT) {
            const uint32_t i1 = i0 + NT;
            int y0, wx0, y1 = 0, wx1 = P.WW;
            split(P, i0, y0, wx0);
            if (i1 < P.nwords)
                split(P, i1, y1, wx1);
            const bool v0 = wx0 < P.WW, v1 = wx1 < P.WW;
            uint4 fa0, fa1, ba0, ba1, fb0, fb1, bb0, bb1;
            fa0 = fa1 = ba0 = ba1 = fb0 = fb1 = bb0 = bb1 = make_uint4(0, 0, 0, 0);
            if (v0) {
                const size_t o = size_t(y0) * P.W + 32u * wx0;
                fa0 = __ldcs(reinterpret_cast<const uint4 *>(fr + o)); // frames are read once: streaming
                fa1 = __ldcs(reinterpret_cast<const uint4 *>(fr + o + 16));
                ba0 = __ldg(reinterpret_cast<const uint4 *>(P.bg + o));
                ba1 = __ldg(reinterpret_cast<const uint4 *>(P.bg + o + 16));
            }
            if (v1) {
                const size_t o = size_t(y1) * P.W + 32u * wx1;
                fb0 = __ldcs(reinterpret_cast<const uint4 *>(fr + o));
                fb1 = __ldcs(reinterpret_cast<const uint4 *>(fr + o + 16));
                bb0 = __ldg(reinterpret_cast<const uint4 *>(P.bg + o));
                bb1 = __ldg(reinterpret_cast<const uint4 *>(P.bg + o + 16));
            }
            uint32_t wa = 0, wu = 0, wl = 0;
            if (v0)
                bits_of_32(fa0, fa1, ba0, ba1, ta, tu, tl, wa, wu, wl);
            A[i0] = wa;
            U[i0] = wu;
            L[i0] = wl;
            if (i1 < P.nwords) {
                wa = wu = wl = 0;
                if (v1)
                    bits_of_32(fb0, fb1, bb0, bb1, ta, tu, tl, wa, wu, wl);
                A[i1] = wa;
                U[i1] = wu;
                L[i1] = wl;
            }
        }
    } else {
        for (uint32_t i = tid; i < P.nwords; i += NT) {
            int y, wx;
            split(P, i, y, wx);
            uint32_t wa = 0, wu = 0, wl = 0;
            if (wx < P.WW) {
                const size_t o = size_t(y) * P.W;
                const int x1 = min(32 * wx + 32, P.W);
                for (int x = 32 * wx; x < x1; ++x) {
                    const int d = max(int(__ldg(P.bg + o + x)) - int(__ldg(fr + o + x)), 0);
                    const uint32_t b = 1u << (x & 31);
                    if (d > th_a)
                        wa |= b;
                    if (d > P.hi)
                        wu |= b;
                    if (d > P.lo)
                        wl |= b;
                }
            }
            A[i] = wa;
            U[i] = wu;
            L[i] = wl;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// opening on bit rows (cv::morphologyEx(MORPH_OPEN) :39, :61): taps sorted by (dy, dx); out-of-image samples never win
// ------------------------------------------------------------------------------------------------------------------
template <bool ERODE, bool CG>
__device__ void morph_phase(const FusedArgs &P, const uint32_t *src, uint32_t *dst)
{
    const uint32_t fill = ERODE ? 0xFFFFFFFFu : 0u;
    for (uint32_t i = threadIdx.x; i < P.nwords; i += NT) {
        int y, wx;
        split(P, i, y, wx);
        uint32_t res = 0;
        if (wx < P.WW) {
            auto word = [&](const uint32_t *row, int w) -> uint32_t {
                if (w < 0 || w >= P.WW)
                    return fill;
                uint32_t v = CG ? __ldcg(row + w) : row[w];
                if (ERODE)
                    v |= ~valid_mask(P, w);
                return v;
            };
            uint32_t acc = fill;
            int cur_dy = INT_MIN, cur_q = INT_MIN;
            const uint32_t *row = nullptr;
            uint32_t w0 = fill, w1 = fill;
            for (int k = 0; k < P.noffs; ++k) {
                const short2 o = __ldg(&P.offs[k]);
                if (o.y != cur_dy) {
                    cur_dy = o.y;
                    cur_q = INT_MIN;
                    const int yy = y + o.y;
                    row = (yy >= 0 && yy < P.H) ? src + size_t(yy) * P.WWp : nullptr;
                }
                if (!row)
                    continue;
                const int q = int(o.x) >> 5, r = int(o.x) & 31; // floor division: columns 32*wx + dx .. + 31
                if (q != cur_q) {
                    cur_q = q;
                    w0 = word(row, wx + q);
                    w1 = word(row, wx + q + 1);
                }
                const uint32_t v = r ? __funnelshift_r(w0, w1, r) : w0;
                acc = ERODE ? (acc & v) : (acc | v);
                if (ERODE && acc == 0)
                    break;
            }
            res = acc & valid_mask(P, wx);
        }
        dst[i] = res;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// run extraction.  Warp w owns a contiguous range of 128-bit quads of the bit image; pass 1 counts run starts, a scan
// over the warp totals gives each warp its first id, pass 2 writes the run records.  Returns T (the run count);
// rowoff[H] = T and parent[T] = T (the FRAME node).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t quad_transitions(const FusedArgs &P, const uint4 &w, uint32_t prev_msb, int wx0, uint32_t t[4])
{
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    const bool row_start = wx0 == 0;
    const uint32_t any = w.x | w.y | w.z | w.w, all = w.x & w.y & w.z & w.w;
    if (!row_start && ((any == 0 && prev_msb == 0) || (all == 0xFFFFFFFFu && prev_msb == 1 && 32 * (wx0 + 4) <= P.W))) {
        t[0] = t[1] = t[2] = t[3] = 0;
        return 0;
    }
    uint32_t p = row_start ? (~w.x & 1u) : prev_msb; // x = 0 always starts a run
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        t[k] = (ww[k] ^ ((ww[k] << 1) | p)) & valid_mask(P, wx0 + k);
        p = ww[k] >> 31;
        c += __popc(t[k]);
    }
    return c;
}

__device__ uint32_t extract_runs(const FusedArgs &P, Shared &sh, const uint32_t *img, const RunSet &rs)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nq = P.nwords >> 2;
    const uint32_t per_warp = ((nq + NW * 32 - 1) / (NW * 32)) * 32;
    const uint32_t q0 = min(uint32_t(warp) * per_warp, nq), q1 = min(q0 + per_warp, nq);
    const uint32_t first_carry = (q0 < q1 && q0 > 0) ? (ld(img + 4 * size_t(q0) - 1) >> 31) : 0u;
    // pass 1: count
    {
        uint32_t cnt = 0, carry = first_carry;
        for (uint32_t qb = q0; qb < q1; qb += 32) {
            const uint32_t q = qb + lane;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (q < q1)
                w = __ldcg(reinterpret_cast<const uint4 *>(img) + q);
            uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, w.w, 1) >> 31;
            if (lane == 0)
                prev = carry;
            carry = __shfl_sync(0xFFFFFFFFu, w.w, 31) >> 31;
            if (q < q1) {
                int y, wx0;
                split(P, 4 * q, y, wx0);
                uint32_t t[4];
                cnt += quad_transitions(P, w, prev, wx0, t);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
            cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        if (lane == 0)
            sh.wsum[warp] = cnt;
    }
    __syncthreads();
    uint32_t base = 0, T = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const uint32_t c = sh.wsum[w];
        if (w < warp)
            base += c;
        T += c;
    }
    // pass 2: fill
    {
        uint32_t carry = first_carry;
        for (uint32_t qb = q0; qb < q1; qb += 32) {
            const uint32_t q = qb + lane;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (q < q1)
                w = __ldcg(reinterpret_cast<const uint4 *>(img) + q);
            uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, w.w, 1) >> 31;
            if (lane == 0)
                prev = carry;
            carry = __shfl_sync(0xFFFFFFFFu, w.w, 31) >> 31;
            uint32_t t[4] = {0, 0, 0, 0};
            uint32_t c = 0;
            int y = 0, wx0 = 0;
            if (q < q1) {
                split(P, 4 * q, y, wx0);
                c = quad_transitions(P, w, prev, wx0, t);
            }
            if (__ballot_sync(0xFFFFFFFFu, c != 0) == 0)
                continue;
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d)
                    inc += n;
            }
            uint32_t id = base + inc - c;
            base += __shfl_sync(0xFFFFFFFFu, inc, 31);
            if (c) {
                if (wx0 == 0)
                    rs.rowoff[y] = id;
                const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t tt = t[k];
                    while (tt) {
                        const int b = __ffs(tt) - 1;
                        tt &= tt - 1;
                        rs.xinfo[id] = uint32_t(32 * (wx0 + k) + b) | (uint32_t(y) << 16) | (((ww[k] >> b) & 1u) << 31);
                        st(rs.parent + id, id);
                        ++id;
                    }
                }
            }
        }
    }
    if (threadIdx.x == 0) {
        rs.rowoff[P.H] = T;
        st(rs.parent + T, T);
    }
    __syncthreads();
    return T;
}

// ------------------------------------------------------------------------------------------------------------------
// labelling of runs
// ------------------------------------------------------------------------------------------------------------------
template <bool FG8, bool FRAME, bool MERGE_FG>
__device__ void merge_phase(const FusedArgs &P, const RunSet &rs, uint32_t T)
{
    for (uint32_t r = threadIdx.x; r < T; r += NT) {
        const uint32_t xi = rs.xinfo[r];
        const uint32_t s = run_x(xi), y = run_y(xi), v = run_v(xi);
        if (!MERGE_FG && v)
            continue;
        const uint32_t e = run_end(rs.xinfo, r, rs.rowoff[y + 1], P.W);
        if (y > 0) {
            const uint32_t d = (FG8 && v) ? 1u : 0u; // 8-connected runs may touch diagonally
            const uint32_t a = rs.rowoff[y - 1], b = rs.rowoff[y];
            const uint32_t c0 = s >= d ? s - d : 0u;
            const uint32_t c1 = min(e + d, uint32_t(P.W - 1));
            for (uint32_t j = run_at(rs.xinfo, a, b, c0); j < b; ++j) {
                const uint32_t xj = rs.xinfo[j];
                if (run_x(xj) > c1)
                    break;
                if (run_v(xj) == v)
                    uf_union(rs.parent, r, j);
            }
        }
        if (FRAME && !v && (y == 0 || y == uint32_t(P.H - 1) || s == 0 || e == uint32_t(P.W - 1)))
            uf_union(rs.parent, r, T);
    }
}

__device__ void flatten_phase(const RunSet &rs, uint32_t T)
{
    for (uint32_t r = threadIdx.x; r <= T; r += NT)
        st(rs.parent + r, uf_find(rs.parent, r));
}

template <bool FG8, bool FRAME, bool MERGE_FG>
__device__ uint32_t label_runs(const FusedArgs &P, Shared &sh, const uint32_t *img, const RunSet &rs)
{
    const uint32_t T = extract_runs(P, sh, img, rs);
    merge_phase<FG8, FRAME, MERGE_FG>(P, rs, T);
    __syncthreads();
    flatten_phase(rs, T);
    __syncthreads();
    return T;
}

// ------------------------------------------------------------------------------------------------------------------
// hysteresis (ThresholdImageWithHysteresis :107-144)
// ------------------------------------------------------------------------------------------------------------------
__device__ void hysteresis_phase(const FusedArgs &P, const RunSet &ru, uint32_t Tu, const RunSet &rl, uint32_t Tl, int *marks,
                                 uint32_t *out)
{
    for (uint32_t r = threadIdx.x; r <= Tl; r += NT)
        st(marks + r, 0);
    for (uint32_t i = threadIdx.x; i < (P.nwords >> 2); i += NT)
        reinterpret_cast<uint4 *>(out)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const uint32_t frame_root = ld(ru.parent + Tu);
    for (uint32_t r = threadIdx.x; r < Tu; r += NT) {
        const uint32_t xi = ru.xinfo[r];
        if (!run_v(xi) || ld(ru.parent + r) != r)
            continue; // seeds are the raster-first pixels of the hi components (contour[0])
        const uint32_t s = run_x(xi), y = run_y(xi);
        // RETR_EXTERNAL: the region left of the first pixel (the component's outer background) must be FRAME
        const bool external = (s == 0) || (ld(ru.parent + r - 1) == frame_root);
        if (external) {
            const uint32_t j = run_at(rl.xinfo, rl.rowoff[y], rl.rowoff[y + 1], s);
            st(marks + ld(rl.parent + j), 1);
        }
    }
    __syncthreads();
    // runs of the lower mask whose region holds a seed (both values: the lo > hi quirk)
    for (uint32_t r = threadIdx.x; r < Tl; r += NT) {
        if (!ld(marks + ld(rl.parent + r)))
            continue;
        const uint32_t xi = rl.xinfo[r];
        const uint32_t y = run_y(xi);
        set_range(out + size_t(y) * P.WWp, int(run_x(xi)), int(run_end(rl.xinfo, r, rl.rowoff[y + 1], P.W)));
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// remove small objects (RemoveSmallObjects :146-181), in place on img
// ------------------------------------------------------------------------------------------------------------------
__device__ void rso_phase(const FusedArgs &P, uint32_t *img, const RunSet &rs, uint32_t T, uint32_t *link, int *st_s, int *st_e,
                          int *st_x, int min_size)
{
    const uint32_t frame_root = ld(rs.parent + T);
    // roots: zero the statistics; link = outer background region (components) / parent component (holes)
    for (uint32_t r = threadIdx.x; r <= T; r += NT) {
        if (r < T && ld(rs.parent + r) != r)
            continue;
        st(st_s + r, 0);
        st(st_e + r, 0);
        st(st_x + r, 0);
        if (r == T)
            continue;
        const uint32_t xi = rs.xinfo[r];
        const uint32_t s = run_x(xi);
        if (run_v(xi))
            st(link + r, s == 0 ? T : bg_region(rs.parent, r - 1, T, frame_root)); // region left of the first pixel
        else
            st(link + r, (r == frame_root || s == 0) ? kNoLabel : ld(rs.parent + r - 1)); // component left of the hole
    }
    __syncthreads();
    // contour statistics, accumulated on the contour's owner: the component for its outer contour, the hole's
    // background region for a hole contour
    for (uint32_t r = threadIdx.x; r < T; r += NT) {
        const uint32_t xi = rs.xinfo[r];
        if (!run_v(xi))
            continue;
        const int s = int(run_x(xi)), y = int(run_y(xi));
        const int e = int(run_end(rs.xinfo, r, rs.rowoff[y + 1], P.W));
        const uint32_t C = ld(rs.parent + r);
        const uint32_t bout = ld(link + C);
        auto owner = [&](uint32_t b) { return b == bout ? C : b; };
        const uint32_t own_l = owner(s == 0 ? T : bg_region(rs.parent, r - 1, T, frame_root));
        const uint32_t own_r = owner(e == P.W - 1 ? T : bg_region(rs.parent, r + 1, T, frame_root));
        // horizontal cracks at the two run ends, and the convex corners there: 2x2 blocks in which a run end is the
        // only foreground pixel
        int xl = 0, xr = 0;
        if (!bit_at(img, P, s, y - 1) && !bit_at(img, P, s - 1, y - 1))
            ++xl;
        if (!bit_at(img, P, s, y + 1) && !bit_at(img, P, s - 1, y + 1))
            ++xl;
        if (!bit_at(img, P, e, y - 1) && !bit_at(img, P, e + 1, y - 1))
            ++xr;
        if (!bit_at(img, P, e, y + 1) && !bit_at(img, P, e + 1, y + 1))
            ++xr;
        if (own_l == own_r) {
            atomicAdd(&st_s[own_l], e + 1 - s);
            atomicAdd(&st_e[own_l], 2);
            if (xl + xr)
                atomicAdd(&st_x[own_l], xl + xr);
        } else {
            atomicAdd(&st_s[own_l], -s);
            atomicAdd(&st_e[own_l], 1);
            atomicAdd(&st_s[own_r], e + 1);
            atomicAdd(&st_e[own_r], 1);
            if (xl)
                atomicAdd(&st_x[own_l], xl);
            if (xr)
                atomicAdd(&st_x[own_r], xr);
        }
        // vertical cracks: columns of this run whose neighbour in the adjacent row is background
        for (int dy = -1; dy <= 1; dy += 2) {
            const int yy = y + dy;
            if (yy < 0 || yy >= P.H) {
                atomicAdd(&st_e[owner(T)], e - s + 1);
                continue;
            }
            const uint32_t a = rs.rowoff[yy], b = rs.rowoff[yy + 1];
            for (uint32_t j = run_at(rs.xinfo, a, b, uint32_t(s)); j < b; ++j) {
                const uint32_t xj = rs.xinfo[j];
                if (int(run_x(xj)) > e)
                    break;
                if (run_v(xj))
                    continue;
                const int js = max(int(run_x(xj)), s), je = min(int(run_end(rs.xinfo, j, b, P.W)), e);
                atomicAdd(&st_e[owner(bg_region(rs.parent, j, T, frame_root))], je - js + 1);
            }
        }
    }
    __syncthreads();
    // per root: st_e <- small flag (2*area < 2*min_size; contourArea(c) < min_size, :171)
    for (uint32_t r = threadIdx.x; r < T; r += NT) {
        if (ld(rs.parent + r) != r)
            continue;
        const long long s = ld(st_s + r);
        const int cracks = ld(st_e + r);
        const long long len = (long long)cracks - ld(st_x + r);
        const long long two_a = run_v(rs.xinfo[r]) ? 2 * s - len - 2 : 2 * (s < 0 ? -s : s) + len - 2;
        st(st_e + r, (cracks > 0 && two_a < 2ll * min_size) ? 1 : 0);
    }
    __syncthreads();
    // per component: st_x <- parity of the number of consecutive small contours up the nesting chain
    for (uint32_t r = threadIdx.x; r < T; r += NT) {
        if (ld(rs.parent + r) != r || !run_v(rs.xinfo[r]))
            continue;
        int count = 0;
        uint32_t cur = r;
        for (;;) {
            if (!ld(st_e + cur))
                break;
            ++count;
            const uint32_t b = ld(link + cur);
            if (b == T)
                break;
            if (!ld(st_e + b))
                break;
            ++count;
            const uint32_t par = ld(link + b);
            if (par == kNoLabel)
                break;
            cur = par;
        }
        st(st_x + r, count & 1);
    }
    __syncthreads();
    // clear the pixels the single filled drawContours call erases (:178)
    for (uint32_t r = threadIdx.x; r < T; r += NT) {
        const uint32_t xi = rs.xinfo[r];
        if (!run_v(xi))
            continue;
        const int s = int(run_x(xi)), y = int(run_y(xi));
        const int e = int(run_end(rs.xinfo, r, rs.rowoff[y + 1], P.W));
        const uint32_t C = ld(rs.parent + r);
        uint32_t *row = img + size_t(y) * P.WWp;
        if (ld(st_x + C)) {
            clear_range(row, s, e);
            continue;
        }
        const uint32_t bout = ld(link + C);
        auto is_small = [&](uint32_t b) { return ld(st_e + (b == bout ? C : b)) != 0; };
        if (is_small(s == 0 ? T : bg_region(rs.parent, r - 1, T, frame_root)))
            clear_range(row, s, s);
        if (is_small(e == P.W - 1 ? T : bg_region(rs.parent, r + 1, T, frame_root)))
            clear_range(row, e, e);
        for (int dy = -1; dy <= 1; dy += 2) {
            const int yy = y + dy;
            if (yy < 0 || yy >= P.H) {
                if (is_small(T))
                    clear_range(row, s, e);
                continue;
            }
            const uint32_t a = rs.rowoff[yy], b = rs.rowoff[yy + 1];
            for (uint32_t j = run_at(rs.xinfo, a, b, uint32_t(s)); j < b; ++j) {
                const uint32_t xj = rs.xinfo[j];
                if (int(run_x(xj)) > e)
                    break;
                if (run_v(xj) || !is_small(bg_region(rs.parent, j, T, frame_root)))
                    continue;
                clear_range(row, max(int(run_x(xj)), s), min(int(run_end(rs.xinfo, j, b, P.W)), e));
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// hole fill (FillHoles :183-221), in place: set every background run that is not connected to the seed corner.
// white <- 1 when the seed pixel itself is set (the flood fill is then a no-op and the result is all 255).
// ------------------------------------------------------------------------------------------------------------------
__device__ void fill_phase(const FusedArgs &P, uint32_t *img, const RunSet &rs, uint32_t T, int *white)
{
    // seed = (0,0) if that pixel is set, else the bottom-right corner (follow the code :201-209, not its comment)
    const uint32_t seed = run_v(rs.xinfo[0]) ? 0u : T - 1u;
    if (run_v(rs.xinfo[seed])) {
        if (threadIdx.x == 0)
            *white = 1;
    } else {
        const uint32_t seed_root = ld(rs.parent + seed);
        for (uint32_t r = threadIdx.x; r < T; r += NT) {
            const uint32_t xi = rs.xinfo[r];
            if (run_v(xi) || ld(rs.parent + r) == seed_root)
                continue;
            const uint32_t y = run_y(xi);
            set_range(img + size_t(y) * P.WWp, int(run_x(xi)), int(run_end(rs.xinfo, r, rs.rowoff[y + 1], P.W)));
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// expand: (A | B) -> 0 / 255 bytes (:77)
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t n) // bit j -> byte j = 0xFF
{
    return (((n & 0xFu) * 0x00204081u) & 0x01010101u) * 0xFFu;
}

__device__ void expand_phase(const FusedArgs &P, unsigned f, const uint32_t *A, const uint32_t *B, bool white)
{
    uint8_t *dst = P.out + size_t(f) * P.out_stride;
    for (uint32_t i = threadIdx.x; i < P.nwords; i += NT) {
        int y, wx;
        split(P, i, y, wx);
        if (wx >= P.WW)
            continue;
        const uint32_t w = white ? 0xFFFFFFFFu : (ld(A + i) | ld(B + i));
        uint8_t *o = dst + size_t(y) * P.W + 32u * wx;
        if (P.fast_io) {
            uint4 lo4, hi4;
            lo4.x = nibble_to_bytes(w);
            lo4.y = nibble_to_bytes(w >> 4);
            lo4.z = nibble_to_bytes(w >> 8);
            lo4.w = nibble_to_bytes(w >> 12);
            hi4.x = nibble_to_bytes(w >> 16);
            hi4.y = nibble_to_bytes(w >> 20);
            hi4.z = nibble_to_bytes(w >> 24);
            hi4.w = nibble_to_bytes(w >> 28);
            __stcs(reinterpret_cast<uint4 *>(o), lo4); // masks are written once: streaming
            __stcs(reinterpret_cast<uint4 *>(o) + 1, hi4);
        } else {
            const int n = min(32, P.W - 32 * wx);
            for (int k = 0; k < n; ++k)
                o[k] = ((w >> k) & 1u) ? 255 : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2) highlight_fused_kernel(const FusedArgs P)
{
    __shared__ Shared sh;
    const size_t slot = blockIdx.x;
    uint32_t *A = P.bits + slot * kImages * size_t(P.nwords);
    uint32_t *U = A + P.nwords, *L = U + P.nwords, *Tm = L + P.nwords;
    uint32_t *rbase = P.runs + slot * kRunArrays * size_t(P.cap);
    uint32_t *ro = P.rowoff + slot * 2 * size_t(P.rstride);
    const RunSet ra{rbase, rbase + P.cap, ro};
    const RunSet rb{rbase + 2 * size_t(P.cap), rbase + 3 * size_t(P.cap), ro + P.rstride};
    uint32_t *link = rbase + 4 * size_t(P.cap);
    int *st_s = reinterpret_cast<int *>(rbase + 5 * size_t(P.cap));
    int *st_e = reinterpret_cast<int *>(rbase + 6 * size_t(P.cap));
    int *st_x = reinterpret_cast<int *>(rbase + 7 * size_t(P.cap));

    for (;;) {
        if (threadIdx.x == 0) {
            sh.frame = atomicAdd(&P.queue[0], 1u);
            sh.white[0] = sh.white[1] = 0;
        }
        __syncthreads();
        const unsigned f = sh.frame;
        if (f >= P.nframes)
            break;
        bits_phase(P, f, A, U, L);
        __syncthreads();
        // debug_stage (CVVP_HL_DEBUG_STAGE): 1 A=d>th, 2 U=d>hi, 3 L=d>lo, 4 A opened, 5 A small removed, 6 A filled,
        // 7 B=hysteresis, 8 B opened, 9 B small removed, 10 B filled
#define CVVP_DEBUG_STAGE(k, img, w)                                                                                    \
    if (P.debug_stage == (k)) {                                                                                        \
        expand_phase(P, f, (img), (img), (w));                                                                         \
        __syncthreads();                                                                                               \
        continue;                                                                                                      \
    }
        CVVP_DEBUG_STAGE(1, A, false)
        CVVP_DEBUG_STAGE(2, U, false)
        CVVP_DEBUG_STAGE(3, L, false)
        // ---- branch A: threshold -> open -> remove small -> fill holes                                  (:35-47)
        morph_phase<true, false>(P, A, Tm);
        __syncthreads();
        morph_phase<false, false>(P, Tm, A);
        __syncthreads();
        CVVP_DEBUG_STAGE(4, A, false)
        uint32_t T = label_runs<true, true, true>(P, sh, A, ra);
        rso_phase(P, A, ra, T, link, st_s, st_e, st_x, P.min_th);
        CVVP_DEBUG_STAGE(5, A, false)
        T = label_runs<false, false, false>(P, sh, A, ra);
        fill_phase(P, A, ra, T, &sh.white[0]);
        CVVP_DEBUG_STAGE(6, A, sh.white[0] != 0)
        // ---- branch B: hysteresis -> open -> remove small -> fill holes                                 (:54-73)
        const uint32_t Tu = label_runs<true, true, true>(P, sh, U, ra);
        const uint32_t Tl = label_runs<false, false, true>(P, sh, L, rb);
        hysteresis_phase(P, ra, Tu, rb, Tl, st_x, U); // U's bits are no longer needed: it now holds the result
        CVVP_DEBUG_STAGE(7, U, false)
        morph_phase<true, true>(P, U, Tm);
        __syncthreads();
        morph_phase<false, false>(P, Tm, U);
        __syncthreads();
        CVVP_DEBUG_STAGE(8, U, false)
        T = label_runs<true, true, true>(P, sh, U, ra);
        rso_phase(P, U, ra, T, link, st_s, st_e, st_x, P.min_hyst);
        CVVP_DEBUG_STAGE(9, U, false)
        T = label_runs<false, false, false>(P, sh, U, ra);
        fill_phase(P, U, ra, T, &sh.white[1]);
        CVVP_DEBUG_STAGE(10, U, sh.white[1] != 0)
#undef CVVP_DEBUG_STAGE
        // ---- out = 255 * (A | B)                                                                         (:77)
        expand_phase(P, f, A, U, sh.white[0] || sh.white[1]);
        __syncthreads();
    }
    // the last CTA to leave re-arms the queue for the next launch
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(&P.queue[1], 1u);
        if (done == gridDim.x - 1) {
            P.queue[0] = 0;
            P.queue[1] = 0;
            __threadfence();
        }
    }
}

size_t round_up_sz(size_t v, size_t a)
{
    return (v + a - 1) / a * a;
}

struct FusedGeom {
    int WW, WWp, wwp_shift;
    uint32_t nwords, cap, rstride;
};

FusedGeom fused_geom(const HlGeom &g)
{
    FusedGeom fg;
    fg.WW = (g.W + 31) / 32;
    int wwp = (fg.WW + 3) & ~3;
    int p2 = 4;
    while (p2 < fg.WW)
        p2 <<= 1;
    fg.wwp_shift = -1;
    if (p2 - fg.WW <= fg.WW / 8) { // a power-of-two row pitch replaces the divisions by shifts (<= 12.5 % padding)
        wwp = p2;
        fg.wwp_shift = 0;
        while ((1 << fg.wwp_shift) < p2)
            ++fg.wwp_shift;
    }
    fg.WWp = wwp;
    fg.nwords = uint32_t(g.H) * uint32_t(wwp);
    fg.cap = uint32_t(round_up_sz(size_t(g.npix) + 2, 4));
    fg.rstride = uint32_t(round_up_sz(size_t(g.H) + 2, 4));
    return fg;
}

size_t slot_bytes(const FusedGeom &fg)
{
    return sizeof(uint32_t) * (size_t(kImages) * fg.nwords + size_t(kRunArrays) * fg.cap + 2 * size_t(fg.rstride));
}
} // namespace

bool fused_supports(const HighlightState *st)
{
    return st->g.W <= 65535 && st->g.H <= 32767;
}

void fused_release(HighlightState *st)
{
    FusedScratch &fs = st->fs;
    void *ptrs[] = {fs.bits, fs.runs, fs.rowoff, fs.queue};
    for (void *p : ptrs)
        if (p)
            cudaFree(p);
    fs = FusedScratch();
}

// number of frames the kernel keeps in flight = resident CTAs, bounded by a scratch budget
int fused_frames_in_flight(cvvp_ctx *ctx, HighlightState *st)
{
    if (st->fs.slots > 0)
        return st->fs.slots;
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, highlight_fused_kernel, NT, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 1;
    }
    const FusedGeom fg = fused_geom(st->g);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        free_b = size_t(8) << 30;
    }
    size_t budget = free_b / 4;
    if (budget > (size_t(24) << 30))
        budget = size_t(24) << 30;
    long long slots = (long long)per_sm * ctx->sm_count;
    const long long by_mem = (long long)(budget / slot_bytes(fg));
    if (slots > by_mem)
        slots = by_mem;
    if (slots < 1)
        slots = 1;
    return int(slots);
}

static int ensure_fused(cvvp_ctx *ctx, HighlightState *st)
{
    FusedScratch &fs = st->fs;
    if (fs.slots > 0)
        return CVVP_OK;
    const int slots = fused_frames_in_flight(ctx, st);
    const FusedGeom fg = fused_geom(st->g);
    const size_t nb = sizeof(uint32_t) * size_t(slots) * kImages * fg.nwords;
    const size_t nr = sizeof(uint32_t) * size_t(slots) * kRunArrays * fg.cap;
    const size_t no = sizeof(uint32_t) * size_t(slots) * 2 * fg.rstride;
    if (cudaMalloc(reinterpret_cast<void **>(&fs.bits), nb) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&fs.runs), nr) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&fs.rowoff), no) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void **>(&fs.queue), 2 * sizeof(unsigned)) != cudaSuccess) {
        cudaGetLastError();
        fused_release(st);
        return fail(ctx, CVVP_ERR_NOMEM, "highlight: cudaMalloc of %zu bytes of scratch failed", nb + nr + no);
    }
    if (cudaMemset(fs.queue, 0, 2 * sizeof(unsigned)) != cudaSuccess) {
        fused_release(st);
        return fail(ctx, CVVP_ERR_CUDA, "highlight: queue initialisation failed");
    }
    fs.slots = slots;
    return CVVP_OK;
}

// One batch of nb frames (device pointers), one kernel launch.  Work in flight on a context's highlight scratch must
// be on one stream at a time.
int highlight_fused_batch(cvvp_ctx *ctx, HighlightState *st, const uint8_t *in, size_t frame_stride, unsigned nb,
                          uint8_t *d_out, size_t out_stride, cudaStream_t stream)
{
    int rc = ensure_fused(ctx, st);
    if (rc != CVVP_OK)
        return rc;
    const FusedGeom fg = fused_geom(st->g);
    FusedArgs P;
    P.frames = in;
    P.frame_stride = frame_stride;
    P.bg = st->d_bg;
    P.out = d_out;
    P.out_stride = out_stride;
    P.th_a = st->th == -1 ? st->d_th : nullptr;
    P.th = st->th;
    P.lo = st->lo;
    P.hi = st->hi;
    P.min_th = st->min_th;
    P.min_hyst = st->min_hyst;
    P.offs = st->d_offs;
    P.noffs = st->noffs;
    P.W = st->g.W;
    P.H = st->g.H;
    P.WW = fg.WW;
    P.WWp = fg.WWp;
    P.wwp_shift = fg.wwp_shift;
    P.nwords = fg.nwords;
    P.cap = fg.cap;
    P.rstride = fg.rstride;
    P.bits = st->fs.bits;
    P.runs = st->fs.runs;
    P.rowoff = st->fs.rowoff;
    P.queue = st->fs.queue;
    P.nframes = nb;
    auto aligned16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    P.fast_io = (P.W % 32 == 0) && aligned16(in) && aligned16(d_out) && aligned16(st->d_bg) && frame_stride % 16 == 0 &&
                out_stride % 16 == 0;
    const char *dbg = getenv("CVVP_HL_DEBUG_STAGE");
    P.debug_stage = dbg ? atoi(dbg) : 0;
    const unsigned grid = nb < unsigned(st->fs.slots) ? nb : unsigned(st->fs.slots);
    highlight_fused_kernel<<<grid, NT, 0, stream>>>(P);
    ctx->launches += 1;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(ctx, CVVP_ERR_CUDA, "highlight: kernel launch failed: %s", cudaGetErrorString(e));
    return CVVP_OK;
}
} // namespace cvvp
