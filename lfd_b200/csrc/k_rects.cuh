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
