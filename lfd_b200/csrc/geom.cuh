// geom.cuh - per-contour geometry that runs one-thread-per-contour on the device:
// convex hull from per-row extremes, cv2-4.13-compatible minAreaRect (rotating calipers),
// boxPoints, and the clipLine / fixed-point polygon-edge setup used by the box rasteriser.
//
// Behaviour follows what the reference gets from cv2 at
// /root/reference/lfd/detecttrails/processfield.py:249 (minAreaRect), :259-260 (boxPoints + int32
// truncation) and :261 (fillPoly).  Everything is float32 with explicit single roundings
// (the library is compiled with -fmad=false; cv2's own build is SSE3-baseline, no FMA).
//
// The functions are __host__ __device__ so tests can exercise the same source on the CPU
// (tests/hostgeom.cpp, test-only); the shipped library only ever runs them on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define LFD_HD __host__ __device__ __forceinline__
#else
#define LFD_HD inline
#endif

namespace lfdgeom {

struct Pt { int x, y; };

LFD_HD long long cross3(Pt o, Pt a, Pt b)
{
    return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x);
}

// Convex hull of the points {(rowmin[i], y0+i), (rowmax[i], y0+i)} (rows with rowmin > rowmax are
// empty).  Points are visited in (y, x) order; turns with cross > 0 are kept (strict hull), which
// yields the same cyclic vertex order as cv2.convexHull(clockwise=False).  `st` needs 2*h+2 slots.
// Returns the vertex count; *start is the index of the lexicographic maximum (max x, then max y),
// the vertex the calipers start from (DESIGN.md "hull start vertex").
LFD_HD int hull_from_rows(const int* rowmin, const int* rowmax, int h, int y0, Pt* st, int* start)
{
    int k = 0;
    // pass 1: ascending (y, x)
    for (int i = 0; i < h; i++) {
        int a = rowmin[i], b = rowmax[i];
        if (a > b) continue;
        for (int t = 0; t < 2; t++) {
            if (t == 1 && b == a) break;
            Pt p; p.x = t ? b : a; p.y = y0 + i;
            while (k >= 2 && cross3(st[k - 2], st[k - 1], p) <= 0) k--;
            st[k++] = p;
        }
    }
    if (k <= 1) { *start = 0; return k; }
    // pass 2: descending (y, x), skipping the last point of pass 1
    int lo = k + 1;
    bool first = true;
    for (int i = h - 1; i >= 0; i--) {
        int a = rowmin[i], b = rowmax[i];
        if (a > b) continue;
        for (int t = 0; t < 2; t++) {
            if (t == 1 && b == a) break;
            Pt p; p.x = t ? a : b; p.y = y0 + i;
            if (first) { first = false; continue; }   // the global (y,x) maximum is already on the stack
            while (k >= lo && cross3(st[k - 2], st[k - 1], p) <= 0) k--;
            st[k++] = p;
        }
    }
    k--;  // last point equals st[0]
    if (k == 2 && st[0].x == st[1].x && st[0].y == st[1].y) k = 1;
    int s = 0;
    for (int i = 1; i < k; i++)
        if (st[i].x > st[s].x || (st[i].x == st[s].x && st[i].y > st[s].y)) s = i;
    *start = s;
    return k;
}

struct Rect { float cx, cy, w, h, angle; };

LFD_HD void normalise_angle(double ang, float w, float h, Rect* r)
{
    while (ang >= 0) { float t = w; w = h; h = t; ang -= 90.0; }
    while (ang < -90) { float t = w; w = h; h = t; ang += 90.0; }
    r->w = w; r->h = h; r->angle = (float)ang;
}

// Rotating calipers of cv2.minAreaRect (4.13) on a strict convex hull of n >= 3 vertices.  `hp(i)` returns
// vertex i (0 <= i < n) in caliper order; vect[2i..2i+1] = edge i -> i+1 as floats, inv[i] = 1/|edge i|;
// left/bottom/right/top = FIRST index of min x / min y / max x / max y (cv2 scans with strict compares).
// edge vector / inverse length of hull edge p -> q as cv2 computes them
LFD_HD void hull_edge(Pt p, Pt q, float* vx, float* vy, float* inv)
{
    double dx = (double)((float)q.x - (float)p.x), dy = (double)((float)q.y - (float)p.y);
    *vx = (float)dx; *vy = (float)dy;
    *inv = (float)(1. / sqrt(dx * dx + dy * dy));
}

// edge provider backed by precomputed arrays
struct ArrEdges {
    const float* vect; const float* inv;
    LFD_HD void vec(int i, float* vx, float* vy) const { *vx = vect[2 * i]; *vy = vect[2 * i + 1]; }
    LFD_HD float invlen(int i) const { return inv[i]; }
};

// edge provider that recomputes edge i from the hull vertices (same float results, no scratch arrays)
template <class HP>
struct FlyEdges {
    HP hp; int n;
    LFD_HD void vec(int i, float* vx, float* vy) const
    {
        Pt p = hp(i), q = hp(i + 1 < n ? i + 1 : 0);
        *vx = (float)q.x - (float)p.x; *vy = (float)q.y - (float)p.y;
    }
    LFD_HD float invlen(int i) const
    {
        float vx, vy, iv;
        hull_edge(hp(i), hp(i + 1 < n ? i + 1 : 0), &vx, &vy, &iv);
        return iv;
    }
};

template <class HP, class EV>
LFD_HD void min_area_rect_core(const HP& hp, int n, const EV& ev, int left, int bottom,
                               int right, int top, Rect* out)
{
    const double RAD2DEG = 180.0 / 3.14159265358979323846;
    float orientation = 0.f;
    {
        float fx, fy;
        ev.vec(n - 1, &fx, &fy);
        double ax = fx, ay = fy;
        for (int i = 0; i < n; i++) {
            ev.vec(i, &fx, &fy);
            double bx = fx, by = fy;
            double c = ax * by - ay * bx;
            if (c != 0) { orientation = c > 0 ? 1.f : -1.f; break; }
            ax = bx; ay = by;
        }
    }
    float base_a = orientation, base_b = 0.f;
    int seq[4] = {bottom, right, top, left};
    float minarea = 3.402823466e+38f, bA = 0, bB = 0, bW = 0, bH = 0;
    int bL = 0, bBt = 0;
    for (int k = 0; k < n; k++) {
        // edge with the smallest rotation: exact cross products of the rotated edge vectors
        float rvx[4], rvy[4], ex[4], ey[4];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
        for (int i = 0; i < 4; i++) ev.vec(seq[i], &ex[i], &ey[i]);
        rvx[0] = ex[0];  rvy[0] = ey[0];
        rvx[1] = ey[1];  rvy[1] = -ex[1];
        rvx[2] = -ex[2]; rvy[2] = -ey[2];
        rvx[3] = -ey[3]; rvy[3] = ex[3];
        int me = 0;
        for (int i = 1; i < 4; i++) {
            float tx = rvy[i], ty = -rvx[i];
            float t1 = tx * rvx[me], t2 = ty * rvy[me];
            if (t1 + t2 < 0) me = i;
        }
        int pi = seq[me];
        float il = ev.invlen(pi);
        float lx = ex[me] * il, ly = ey[me] * il;
        switch (me) {
        case 0: base_a = lx; base_b = ly; break;
        case 1: base_a = ly; base_b = -lx; break;
        case 2: base_a = -lx; base_b = -ly; break;
        default: base_a = -ly; base_b = lx; break;
        }
        seq[me] += 1; if (seq[me] == n) seq[me] = 0;
        Pt q1 = hp(seq[1]), q3 = hp(seq[3]), q2 = hp(seq[2]), q0 = hp(seq[0]);
        float dx = (float)q1.x - (float)q3.x, dy = (float)q1.y - (float)q3.y;
        float w1 = dx * base_a, w2 = dy * base_b;
        float width = w1 + w2;
        dx = (float)q2.x - (float)q0.x; dy = (float)q2.y - (float)q0.y;
        float h1 = -dx * base_b, h2 = dy * base_a;
        float height = h1 + h2;
        float area = width * height;
        if (area <= minarea) { minarea = area; bL = seq[3]; bA = base_a; bW = width; bB = base_b; bH = height; bBt = seq[0]; }
    }
    float A1 = bA, B1 = bB, A2 = -bB, B2 = bA;
    Pt pl = hp(bL), pb = hp(bBt);
    float lxp = (float)pl.x, lyp = (float)pl.y, bxp = (float)pb.x, byp = (float)pb.y;
    float c1a = A1 * lxp, c1b = lyp * B1; float C1 = c1a + c1b;
    float c2a = A2 * bxp, c2b = byp * B2; float C2 = c2a + c2b;
    float d1 = A1 * B2, d2 = A2 * B1;
    float idet = 1.f / (d1 - d2);
    float n1 = C1 * B2, n2 = C2 * B1; float px = (n1 - n2) * idet;
    float n3 = A1 * C2, n4 = A2 * C1; float py = (n3 - n4) * idet;
    float o2 = A1 * bW, o3 = B1 * bW, o4 = A2 * bH, o5 = B2 * bH;
    out->cx = px + (o2 + o4) * 0.5f;
    out->cy = py + (o3 + o5) * 0.5f;
    double o2d = o2, o3d = o3, o4d = o4, o5d = o5;
    float w = (float)sqrt(o2d * o2d + o3d * o3d);
    float h = (float)sqrt(o4d * o4d + o5d * o5d);
    normalise_angle(atan2(o3d, o2d) * RAD2DEG, w, h, out);
}

// the n == 1 and n == 2 cases of cv2.minAreaRect
LFD_HD void min_area_rect_small(int n, Pt a, Pt b, Rect* out)
{
    const double RAD2DEG = 180.0 / 3.14159265358979323846;
    out->cx = out->cy = out->w = out->h = out->angle = 0.f;
    if (n <= 0) return;
    if (n == 1) { out->cx = (float)a.x; out->cy = (float)a.y; out->angle = -90.f; return; }
    out->cx = ((float)a.x + (float)b.x) * 0.5f;
    out->cy = ((float)a.y + (float)b.y) * 0.5f;
    double dx = (double)((float)b.x - (float)a.x), dy = (double)((float)b.y - (float)a.y);
    float w = (float)sqrt(dx * dx + dy * dy);
    normalise_angle(atan2(dy, dx) * RAD2DEG, w, 0.f, out);
}

struct RotHull {
    const Pt* st; int n, start;
    LFD_HD Pt operator()(int i) const { int j = i + start; if (j >= n) j -= n; return st[j]; }
};

// cv2.minAreaRect of a strict convex hull st[0..n) taken in the order start, start+1, ... (mod n).
// vect/inv scratch: 2n and n floats.
LFD_HD void min_area_rect(const Pt* st, int n, int start, float* vect, float* inv, Rect* out)
{
    RotHull hp; hp.st = st; hp.n = n; hp.start = start;
    if (n <= 2) { min_area_rect_small(n, n > 0 ? hp(0) : Pt{0, 0}, n > 1 ? hp(1) : Pt{0, 0}, out); return; }
    int left = 0, bottom = 0, right = 0, top = 0;
    Pt p0 = hp(0);
    float left_x = (float)p0.x, right_x = left_x, top_y = (float)p0.y, bottom_y = top_y;
    for (int i = 0; i < n; i++) {
        Pt p = hp(i), q = hp(i + 1 < n ? i + 1 : 0);
        float px = (float)p.x, py = (float)p.y;
        if (px < left_x) { left_x = px; left = i; }
        if (px > right_x) { right_x = px; right = i; }
        if (py > top_y) { top_y = py; top = i; }
        if (py < bottom_y) { bottom_y = py; bottom = i; }
        hull_edge(p, q, &vect[2 * i], &vect[2 * i + 1], &inv[i]);
    }
    ArrEdges ev; ev.vect = vect; ev.inv = inv;
    min_area_rect_core(hp, n, ev, left, bottom, right, top, out);
}

// cv2.boxPoints followed by the reference's np.asarray(..., int32) truncation toward zero.
LFD_HD void box_points(const Rect& r, float* f8, int* i8)
{
    double ang = (double)r.angle * 3.14159265358979323846 / 180.;
    float b = (float)cos(ang) * 0.5f, a = (float)sin(ang) * 0.5f;
    float ah = a * r.h, bw = b * r.w, bh = b * r.h, aw = a * r.w;
    f8[0] = r.cx - ah - bw; f8[1] = r.cy + bh - aw;
    f8[2] = r.cx + ah - bw; f8[3] = r.cy - bh - aw;
    f8[4] = 2 * r.cx - f8[0]; f8[5] = 2 * r.cy - f8[1];
    f8[6] = 2 * r.cx - f8[2]; f8[7] = 2 * r.cy - f8[3];
    for (int i = 0; i < 8; i++) i8[i] = (int)f8[i];
}

// cv2.clipLine on int64 points; returns 1 if something of the segment is inside.  Endpoints are updated
// even when the segment is rejected (fillPoly's edge setup relies on that).
LFD_HD int clip_line(long long W, long long H, long long* x1, long long* y1, long long* x2, long long* y2)
{
    long long right = W - 1, bottom = H - 1;
    if (W <= 0 || H <= 0) return 0;
    int c1 = (*x1 < 0) + (*x1 > right) * 2 + (*y1 < 0) * 4 + (*y1 > bottom) * 8;
    int c2 = (*x2 < 0) + (*x2 > right) * 2 + (*y2 < 0) * 4 + (*y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            *x1 += (long long)((double)(a - *y1) * (double)(*x2 - *x1) / (double)(*y2 - *y1));
            *y1 = a;
            c1 = (*x1 < 0) + (*x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            *x2 += (long long)((double)(a - *y2) * (double)(*x2 - *x1) / (double)(*y2 - *y1));
            *y2 = a;
            c2 = (*x2 < 0) + (*x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                *y1 += (long long)((double)(a - *x1) * (double)(*y2 - *y1) / (double)(*x2 - *x1));
                *x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                *y2 += (long long)((double)(a - *x2) * (double)(*y2 - *y1) / (double)(*x2 - *x1));
                *x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// One polygon edge of cv2.fillPoly's scan conversion (16.16 fixed point).
struct Edge { int y0, y1; long long x, dx; int valid; };

// Edge p0 -> p1 of an integer polygon on a W x H image.  Also returns the LINE_8 segment to draw
// (lx0..ly1; draw = 0 if it is completely outside).
LFD_HD void poly_edge(int W, int H, int p0x_, int p0y_, int p1x_, int p1y_, Edge* e,
                      int* draw, int* lx0, int* ly0, int* lx1, int* ly1)
{
    const int XY_SHIFT = 16;
    long long p0x = (long long)p0x_ << XY_SHIFT, p0y = p0y_, p1x = (long long)p1x_ << XY_SHIFT, p1y = p1y_;
    long long t0x = p0x_, t0y = p0y_, t1x = p1x_, t1y = p1y_;
    long long c0x = p0x, c0y = p0y, c1x = p1x, c1y = p1y;
    bool outside = (unsigned long long)t0x >= (unsigned long long)W || (unsigned long long)t1x >= (unsigned long long)W ||
                   (unsigned long long)t0y >= (unsigned long long)H || (unsigned long long)t1y >= (unsigned long long)H;
    *draw = 1;
    if (outside) {
        int ok = clip_line(W, H, &t0x, &t0y, &t1x, &t1y);
        *draw = ok;
        if (t0y != t1y) { c0y = t0y; c1y = t1y; }
        c0x = t0x << XY_SHIFT; c1x = t1x << XY_SHIFT;
    }
    *lx0 = (int)t0x; *ly0 = (int)t0y; *lx1 = (int)t1x; *ly1 = (int)t1y;
    e->valid = 0;
    if (p0y == p1y) return;
    e->dx = (c1x - c0x) / (c1y - c0y);
    if (p0y < p1y) { e->y0 = (int)p0y; e->y1 = (int)p1y; e->x = c0x + (p0y - c0y) * e->dx; }
    else { e->y0 = (int)p1y; e->y1 = (int)p0y; e->x = c1x + (p1y - c1y) * e->dx; }
    e->valid = 1;
}

// Spans cv2.fillPoly paints on row y for up to 4 valid edges: active edges sorted by their current x,
// paired; left end rounds up, right end rounds down (16.16), clipped to [0, W-1].
// Returns the number of spans written to xs as (x1, x2) pairs.
LFD_HD int row_spans(const Edge* e, int ne, int y, int W, int* xs)
{
    long long x[4];
    int na = 0;
    for (int i = 0; i < ne; i++)
        if (e[i].valid && e[i].y0 <= y && y < e[i].y1) x[na++] = e[i].x + (long long)(y - e[i].y0) * e[i].dx;
    for (int i = 1; i < na; i++) {
        long long t = x[i]; int j = i - 1;
        while (j >= 0 && x[j] > t) { x[j + 1] = x[j]; j--; }
        x[j + 1] = t;
    }
    int ns = 0;
    for (int i = 0; i + 1 < na; i += 2) {
        int x1 = (int)((x[i] + 65535) >> 16), x2 = (int)(x[i + 1] >> 16);
        if (x1 < W && x2 >= 0) {
            if (x1 < 0) x1 = 0;
            if (x2 >= W) x2 = W - 1;
            if (x1 <= x2) { xs[2 * ns] = x1; xs[2 * ns + 1] = x2; ns++; }
        }
    }
    return ns;
}

// cv2.line(LINE_8) stepping (already clipped endpoints): left-to-right Bresenham.
struct LineIt {
    int x, y, err, plusDelta, minusDelta, sx, sy, vert, count;
    LFD_HD void init(int x1, int y1, int x2, int y2)
    {
        int dx = x2 - x1, dy = y2 - y1;
        sx = 1; sy = 1; x = x1; y = y1;
        if (dx < 0) { dx = -dx; dy = -dy; x = x2; y = y2; }
        if (dy < 0) { dy = -dy; sy = -1; }
        vert = dy > dx;
        if (vert) { int t = dx; dx = dy; dy = t; }
        err = dx - (dy + dy); plusDelta = dx + dx; minusDelta = -(dy + dy); count = dx + 1;
    }
    LFD_HD void next()
    {
        int mask = err < 0 ? -1 : 0;
        err += minusDelta + (plusDelta & mask);
        if (!vert) { x += sx; y += sy & mask; }
        else { y += sy; x += sx & mask; }
    }
};

}  // namespace lfdgeom
