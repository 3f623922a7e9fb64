// k_rects.cuh - per-contour minAreaRect + length/width filter + boxPoints, and the fillPoly rasteriser
// that builds box_img (as a bit mask).
//
// Reference: fit_minAreaRect, /root/reference/lfd/detecttrails/processfield.py:248-261
//   rect = cv2.minAreaRect(cnt); length/width filter; box = int32(cv2.boxPoints(rect));
//   cv2.fillPoly(box_img, [box], 255); detection = True
#pragma once
#include "common.cuh"
#include "geom.cuh"
#include "k_ccl.cuh"

struct RectBuf {
    lfd_rect* rects;     // [2*maxcomp]  same indexing as CompBuf entries
    int* passing;        // [2*maxcomp]  entry indices of rectangles that passed
    lfdgeom::Pt* hull;   // [hullcap]
    float* hullf;        // [3*hullcap]
};

// one thread per contour
__global__ void __launch_bounds__(128)
k_rects(CompBuf* __restrict__ comps, RectBuf* __restrict__ rbufs, const CclBuf* __restrict__ ccl0,
        const CclBuf* __restrict__ ccl1, FrameCtl* __restrict__ ctl, int pass,
        Dims d, double minLen, double lwTresh)
{
    int f = blockIdx.y;
    if (!ctl[f].active[pass]) return;
    CompBuf cb = comps[f];
    RectBuf rb = rbufs[f];
    int n0 = min(ctl[f].ncomp[0], cb.maxcomp), n1 = min(ctl[f].ncomp[1], cb.maxcomp);
    int total = n0 + n1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int kind = i >= n0;
        int e = kind ? cb.maxcomp + (i - n0) : i;
        int hh = cb.h[e], slot = cb.slot[e], ho = cb.hulloff[e], y0 = cb.y0[e];
        lfdgeom::Pt* st = rb.hull + ho;
        float* vect = rb.hullf + 3 * (size_t)ho;
        int start;
        int n = lfdgeom::hull_from_rows(cb.rowmin + slot, cb.rowmax + slot, hh, y0, st, &start);
        lfdgeom::Rect r;
        lfdgeom::min_area_rect(st, n, start, vect, vect + 2 * (size_t)(2 * hh + 2), &r);
        lfd_rect o;
        o.cx = r.cx; o.cy = r.cy; o.w = r.w; o.h = r.h; o.angle = r.angle;
        o.kind = kind;
        Run rr = (kind ? ccl1[f] : ccl0[f]).runs[cb.root[e]];   // raster-first run of the component / hole
        o.key = (int)rr.y * d.W + (int)rr.xs;
        float length = r.w > r.h ? r.w : r.h, width = r.w > r.h ? r.h : r.w;
        int passed = 0;
        if ((double)length > minLen && (double)width > minLen)
            if ((double)length / (double)width > lwTresh) passed = 1;
        o.passed = passed;
        float f8[8];
        lfdgeom::box_points(r, f8, o.box);
        rb.rects[e] = o;
        if (passed) {
            ctl[f].hough[pass] = 1;
            int pi = atomicAdd(&ctl[f].npass, 1);
            rb.passing[pi] = e;
        }
    }
}

// ---- warp-per-contour variant (production) ---------------------------------------------------------
// The convex hull is built in two levels: every lane runs the monotone chain over its own slice of rows
// (both directions), then lane 0 / lane 1 re-run the chain over the concatenated partial chains, which is
// the same strict hull in the same vertex order as the one-thread scan (hull of hulls).  Edge vectors,
// inverse lengths and the four extreme vertices are computed by all lanes; only the rotating loop itself
// is serial.  Contours of up to RECT_SMALL_PTS/2 rows keep everything in shared memory, taller ones use
// the per-frame global scratch.
#define RECT_WARPS 4
#define RECT_SMALL_PTS 128

__device__ __forceinline__ u32 pt_pack(int x, int y) { return (u32)(x & 0xffff) | ((u32)y << 16); }
__device__ __forceinline__ lfdgeom::Pt pt_unpack(u32 v) { lfdgeom::Pt p; p.x = (int)(v & 0xffffu); p.y = (int)(v >> 16); return p; }

// Monotone-chain stack with its two top entries cached in registers: the common case (no pop) touches
// memory only for the store.
struct Chain {
    u32* st; int k; lfdgeom::Pt t1, t2;       // t1 = st[k-1], t2 = st[k-2]
    __device__ __forceinline__ void init(u32* s) { st = s; k = 0; t1.x = t1.y = t2.x = t2.y = 0; }
    __device__ __forceinline__ void push(u32 pv)
    {
        lfdgeom::Pt p = pt_unpack(pv);
        while (k >= 2) {
            int cr = (t1.x - t2.x) * (p.y - t2.y) - (t1.y - t2.y) * (p.x - t2.x);
            if (cr > 0) break;
            k--;
            t1 = t2;
            if (k >= 2) t2 = pt_unpack(st[k - 2]);
        }
        st[k++] = pv;
        t2 = t1; t1 = p;
    }
};

struct PackedHull {
    const u32* a; int n, start;
    __device__ __forceinline__ lfdgeom::Pt operator()(int i) const { int j = i + start; if (j >= n) j -= n; return pt_unpack(a[j]); }
};

__device__ __forceinline__ u64 warp_max64(u64 v)
{
    for (int o = 16; o; o >>= 1) { u64 t = __shfl_xor_sync(FULLMASK, v, o); if (t > v) v = t; }
    return v;
}
__device__ __forceinline__ u64 warp_min64(u64 v)
{
    for (int o = 16; o; o >>= 1) { u64 t = __shfl_xor_sync(FULLMASK, v, o); if (t < v) v = t; }
    return v;
}

__global__ void __launch_bounds__(RECT_WARPS * 32)
k_rects_warp(CompBuf* __restrict__ comps, RectBuf* __restrict__ rbufs, const CclBuf* __restrict__ ccl0,
             const CclBuf* __restrict__ ccl1, FrameCtl* __restrict__ ctl, int pass,
             Dims d, double minLen, double lwTresh)
{
    int f = blockIdx.y;
    if (!ctl[f].active[pass]) return;
    CompBuf cb = comps[f];
    RectBuf rb = rbufs[f];
    int n0 = min(ctl[f].ncomp[0], cb.maxcomp), n1 = min(ctl[f].ncomp[1], cb.maxcomp);
    int total = n0 + n1;
    __shared__ u32 sA[RECT_WARPS][RECT_SMALL_PTS], sB[RECT_WARPS][RECT_SMALL_PTS];
    __shared__ float sV[RECT_WARPS][2 * RECT_SMALL_PTS + 4], sI[RECT_WARPS][RECT_SMALL_PTS + 2];
    __shared__ int sRow[RECT_WARPS][RECT_SMALL_PTS];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    for (int i = blockIdx.x * RECT_WARPS + warp; i < total; i += gridDim.x * RECT_WARPS) {
        int kind = i >= n0;
        int e = kind ? cb.maxcomp + (i - n0) : i;
        int hh = cb.h[e], slot = cb.slot[e], ho = cb.hulloff[e], y0 = cb.y0[e];
        const int* rmin = cb.rowmin + slot;
        const int* rmax = cb.rowmax + slot;
        u32 *A, *B;
        float *vect, *inv;
        __syncwarp();
        if (2 * hh <= RECT_SMALL_PTS) {
            A = sA[warp]; B = sB[warp]; vect = sV[warp]; inv = sI[warp];
            for (int r = lane; r < hh; r += 32) { sRow[warp][r] = rmin[r]; sRow[warp][RECT_SMALL_PTS / 2 + r] = rmax[r]; }
            __syncwarp();
            rmin = sRow[warp]; rmax = sRow[warp] + RECT_SMALL_PTS / 2;
        } else {
            A = reinterpret_cast<u32*>(rb.hull + ho); B = A + 2 * hh;
            vect = rb.hullf + 3 * (size_t)ho; inv = vect + 2 * (size_t)(2 * hh + 2);
        }
        // level 1: per-lane partial chains; B regions are mirrored so that level 2 can run in place
        const int c = (hh + 31) >> 5;
        const int r0 = min(lane * c, hh), r1 = min(r0 + c, hh);
        int ka, kb;
        {
            Chain ca; ca.init(A + 2 * r0);
            for (int r = r0; r < r1; r++) {
                int a = rmin[r], b = rmax[r];
                if (a > b) continue;
                ca.push(pt_pack(a, y0 + r));
                if (b != a) ca.push(pt_pack(b, y0 + r));
            }
            ka = ca.k;
            Chain cbk; cbk.init(B + 2 * (hh - r1));
            for (int r = r1 - 1; r >= r0; r--) {
                int a = rmin[r], b = rmax[r];
                if (a > b) continue;
                cbk.push(pt_pack(b, y0 + r));
                if (a != b) cbk.push(pt_pack(a, y0 + r));
            }
            kb = cbk.k;
        }
        __syncwarp();
        // level 2: lane 0 merges the ascending chains, lane 1 the descending ones (same instruction stream)
        Chain cm; cm.init(lane == 1 ? B : A);
        for (int t = 0; t < 32; t++) {
            int na_t = __shfl_sync(FULLMASK, ka, t), nb_t = __shfl_sync(FULLMASK, kb, 31 - t);
            if (lane < 2) {
                int L = lane ? 31 - t : t;
                int q0 = min(L * c, hh), q1 = min(q0 + c, hh);
                const u32* src = lane ? B + 2 * (hh - q1) : A + 2 * q0;
                int cnt = lane ? nb_t : na_t;
                for (int j = 0; j < cnt; j++) cm.push(src[j]);
            }
        }
        const int KA = __shfl_sync(FULLMASK, cm.k, 0), KB = __shfl_sync(FULLMASK, cm.k, 1);
        const int n = (KA <= 1) ? KA : KA + KB - 2;
        __syncwarp();
        // gather the hull (A chain, then the interior of the B chain); small hulls of tall contours move to
        // shared memory so that the serial caliper loop never waits on global memory
        if (2 * hh > RECT_SMALL_PTS && n <= RECT_SMALL_PTS) {
            u32* H2 = sA[warp];
            for (int j = lane; j < n; j += 32) H2[j] = j < KA ? A[j] : B[1 + j - KA];
            A = H2; vect = sV[warp]; inv = sI[warp];
        } else {
            for (int j = lane; j < KB - 2; j += 32) A[KA + j] = B[1 + j];
        }
        __syncwarp();
        // caliper start vertex: lexicographic maximum (x, then y)
        u64 bk = 0;
        for (int j = lane; j < n; j += 32) {
            u32 v = A[j];
            u64 k64 = ((u64)(((v & 0xffffu) << 16) | (v >> 16)) << 32) | (u32)j;
            if (k64 > bk) bk = k64;
        }
        bk = warp_max64(bk);
        PackedHull hp; hp.a = A; hp.n = n; hp.start = (int)(bk & 0xffffffffu);
        lfdgeom::Rect r;
        if (n <= 2) {
            lfdgeom::Pt z; z.x = 0; z.y = 0;
            lfdgeom::min_area_rect_small(n, n > 0 ? hp(0) : z, n > 1 ? hp(1) : z, &r);
        } else {
            u64 kl = ~0ull, kbm = ~0ull, kr = 0, kt = 0;       // first index of min x / min y / max x / max y
            for (int j = lane; j < n; j += 32) {
                lfdgeom::Pt p = hp(j), q = hp(j + 1 < n ? j + 1 : 0);
                lfdgeom::hull_edge(p, q, &vect[2 * j], &vect[2 * j + 1], &inv[j]);
                u64 lo = (u32)j, hi = 0xffffffffu - (u32)j;
                u64 a = ((u64)(u32)p.x << 32) | lo; if (a < kl) kl = a;
                a = ((u64)(u32)p.y << 32) | lo; if (a < kbm) kbm = a;
                a = ((u64)(u32)p.x << 32) | hi; if (a > kr) kr = a;
                a = ((u64)(u32)p.y << 32) | hi; if (a > kt) kt = a;
            }
            kl = warp_min64(kl); kbm = warp_min64(kbm); kr = warp_max64(kr); kt = warp_max64(kt);
            __syncwarp();
            if (lane == 0)
                lfdgeom::min_area_rect_core(hp, n, vect, inv, (int)(kl & 0xffffffffu), (int)(kbm & 0xffffffffu),
                                            (int)(0xffffffffu - (u32)(kr & 0xffffffffu)), (int)(0xffffffffu - (u32)(kt & 0xffffffffu)), &r);
        }
        if (lane == 0) {
            lfd_rect o;
            o.cx = r.cx; o.cy = r.cy; o.w = r.w; o.h = r.h; o.angle = r.angle;
            o.kind = kind;
            Run rr = (kind ? ccl1[f] : ccl0[f]).runs[cb.root[e]];   // raster-first run of the component / hole
            o.key = (int)rr.y * d.W + (int)rr.xs;
            float length = r.w > r.h ? r.w : r.h, width = r.w > r.h ? r.h : r.w;
            int passed = 0;
            if ((double)length > minLen && (double)width > minLen)
                if ((double)length / (double)width > lwTresh) passed = 1;
            o.passed = passed;
            float f8[8];
            lfdgeom::box_points(r, f8, o.box);
            rb.rects[e] = o;
            if (passed) {
                ctl[f].hough[pass] = 1;
                int pi = atomicAdd(&ctl[f].npass, 1);
                rb.passing[pi] = e;
            }
        }
    }
}

// cv2.fillPoly(box_img, [box], 255) for every passing rectangle; one warp per rectangle.
__global__ void __launch_bounds__(128)
k_fill_boxes(RectBuf* __restrict__ rbufs, u32* __restrict__ box, const FrameCtl* __restrict__ ctl, int pass, Dims d)
{
    int f = blockIdx.y;
    if (!ctl[f].active[pass] || !ctl[f].hough[pass]) return;
    RectBuf rb = rbufs[f];
    u32* bm = box + (size_t)f * d.NW;
    int np = ctl[f].npass;
    int warps = (gridDim.x * blockDim.x) >> 5;
    int lane = lane_id();
    __shared__ lfdgeom::Edge sedge[4][4];
    int wslot = threadIdx.x >> 5;
    for (int pi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pi < np; pi += warps) {
        const lfd_rect& R = rb.rects[rb.passing[pi]];
        __syncwarp();
        if (lane < 4) {
            int i = lane, j = (lane + 3) & 3;
            int draw, x0, y0, x1, y1;
            lfdgeom::Edge e;
            lfdgeom::poly_edge(d.W, d.H, R.box[2 * j], R.box[2 * j + 1], R.box[2 * i], R.box[2 * i + 1], &e,
                               &draw, &x0, &y0, &x1, &y1);
            sedge[wslot][lane] = e;
            if (draw) {
                lfdgeom::LineIt it;
                it.init(x0, y0, x1, y1);
                for (int s = 0; s < it.count; s++) {
                    atomicOr(&bm[(size_t)it.y * d.WW + (it.x >> 5)], 1u << (it.x & 31));
                    it.next();
                }
            }
        }
        __syncwarp();
        lfdgeom::Edge e4[4];
        int ymin = 0x7fffffff, ymax = -0x7fffffff, nv = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            e4[i] = sedge[wslot][i];
            if (e4[i].valid) { nv++; ymin = min(ymin, e4[i].y0); ymax = max(ymax, e4[i].y1); }
        }
        if (nv >= 2) {
            ymin = max(ymin, 0);
            ymax = min(ymax, d.H);
            for (int y = ymin + lane; y < ymax; y += 32) {
                int xs[4];
                int ns = lfdgeom::row_spans(e4, 4, y, d.W, xs);
                for (int s = 0; s < ns; s++) {
                    int x1 = xs[2 * s], x2 = xs[2 * s + 1];
                    for (int w = x1 >> 5; w <= (x2 >> 5); w++) {
                        int blo = max(x1 - (w << 5), 0), bhi = min(x2 - (w << 5), 31);
                        atomicOr(&bm[(size_t)y * d.WW + w], bit_range(blo, bhi));
                    }
                }
            }
        }
    }
}
