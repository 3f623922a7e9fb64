/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the OpenCV routines that lfd.detecttrails' per-frame path calls
 * (/root/reference/lfd/detecttrails/processfield.py:236 Canny, :241-246 findContours,
 * :249 minAreaRect, :259 boxPoints, :261 fillPoly, :370-371/:488-489 HoughLines).
 * OpenCV itself is a third-party dependency that is not under /root/reference
 * (setup.py:18-28 "opencv-python", unpinned); the parity target is the cv2 4.13.0 binary of
 * this image.  Every function here is pinned bit-for-bit against that binary by
 * tests/test_oracle_cv.py; what cv2 does not expose (Hough accumulator, NMS classes,
 * component/hole structure) is validated through the outputs that depend on it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may load this library.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, like cv2's SSE3 baseline)
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

static inline int cv_round(double v) { return (int)lrint(v); }       /* round half to even */
static inline int cv_roundf(float v) { return (int)lrintf(v); }

/* ------------------------------------------------------------------------------------------ */
/* Canny(img, low, high), aperture 3, L1 gradient.  cls: 0 none, 1 weak candidate, 2 strong.    */
/* ------------------------------------------------------------------------------------------ */
EXPORT void orc_canny_classes(const uint8_t* img, int H, int W, int low, int high, uint8_t* cls,
                              int32_t* mag_out /* may be NULL */)
{
    int32_t* mag = (int32_t*)calloc((size_t)(H + 2) * (W + 2), sizeof(int32_t));
    int16_t* gx = (int16_t*)malloc((size_t)H * W * sizeof(int16_t));
    int16_t* gy = (int16_t*)malloc((size_t)H * W * sizeof(int16_t));
    const int MS = W + 2;
#define PX(y, x) img[(size_t)((y) < 0 ? 0 : (y) >= H ? H - 1 : (y)) * W + ((x) < 0 ? 0 : (x) >= W ? W - 1 : (x))]
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int dx = (PX(y - 1, x + 1) + 2 * PX(y, x + 1) + PX(y + 1, x + 1)) -
                     (PX(y - 1, x - 1) + 2 * PX(y, x - 1) + PX(y + 1, x - 1));
            int dy = (PX(y + 1, x - 1) + 2 * PX(y + 1, x) + PX(y + 1, x + 1)) -
                     (PX(y - 1, x - 1) + 2 * PX(y - 1, x) + PX(y - 1, x + 1));
            gx[(size_t)y * W + x] = (int16_t)dx;
            gy[(size_t)y * W + x] = (int16_t)dy;
            mag[(size_t)(y + 1) * MS + x + 1] = abs(dx) + abs(dy);
        }
#undef PX
    const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int32_t* m = mag + (size_t)(y + 1) * MS + x + 1;
            int mm = m[0];
            uint8_t c = 0;
            if (mm > low) {
                int dx = gx[(size_t)y * W + x], dy = gy[(size_t)y * W + x];
                int ax = abs(dx), ay = abs(dy) << 15;
                int tg22x = ax * TG22;
                int keep;
                if (ay < tg22x)
                    keep = mm > m[-1] && mm >= m[1];
                else {
                    int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x)
                        keep = mm > m[-MS] && mm >= m[MS];
                    else {
                        int s = (dx ^ dy) < 0 ? -1 : 1;
                        keep = mm > m[-MS - s] && mm > m[MS + s];
                    }
                }
                if (keep) c = mm > high ? 2 : 1;
            }
            cls[(size_t)y * W + x] = c;
        }
    if (mag_out)
        for (int y = 0; y < H; y++)
            memcpy(mag_out + (size_t)y * W, mag + (size_t)(y + 1) * MS + 1, W * sizeof(int32_t));
    free(mag); free(gx); free(gy);
}

/* hysteresis: keep strong pixels and weak pixels 8-connected to them through candidates */
EXPORT void orc_canny_hysteresis(const uint8_t* cls, int H, int W, uint8_t* edges)
{
    int32_t* stack = (int32_t*)malloc((size_t)H * W * sizeof(int32_t));
    size_t sp = 0;
    memset(edges, 0, (size_t)H * W);
    for (int i = 0; i < H * W; i++)
        if (cls[i] == 2) { edges[i] = 255; stack[sp++] = i; }
    while (sp) {
        int i = stack[--sp];
        int y = i / W, x = i % W;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++) {
                int yy = y + dy, xx = x + dx;
                if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                int j = yy * W + xx;
                if (cls[j] == 1 && !edges[j]) { edges[j] = 255; stack[sp++] = j; }
            }
    }
    free(stack);
}

/* ------------------------------------------------------------------------------------------ */
/* connected components: fg 8-connected / bg 4-connected (flood fill).  Labels are raster-first  */
/* pixel indices; bg components that touch the frame border get label -2 ("outside"), fg px -1   */
/* in the bg map and bg px -1 in the fg map.                                                    */
/* ------------------------------------------------------------------------------------------ */
static void flood(const uint8_t* img, int H, int W, int fg, int conn8, int32_t* lab)
{
    int32_t* stack = (int32_t*)malloc((size_t)H * W * sizeof(int32_t));
    for (int i = 0; i < H * W; i++) lab[i] = -1;
    for (int s = 0; s < H * W; s++) {
        if (((img[s] != 0) != fg) || lab[s] != -1) continue;
        size_t sp = 0;
        stack[sp++] = s; lab[s] = s;
        while (sp) {
            int i = stack[--sp];
            int y = i / W, x = i % W;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    if (!dx && !dy) continue;
                    if (!conn8 && dx && dy) continue;
                    int yy = y + dy, xx = x + dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    int j = yy * W + xx;
                    if (((img[j] != 0) == fg) && lab[j] == -1) { lab[j] = s; stack[sp++] = j; }
                }
        }
    }
    free(stack);
}

EXPORT void orc_label_fg8(const uint8_t* img, int H, int W, int32_t* lab) { flood(img, H, W, 1, 1, lab); }

EXPORT void orc_label_bg4(const uint8_t* img, int H, int W, int32_t* lab)
{
    flood(img, H, W, 0, 0, lab);
    /* mark components that reach the border as outside (-2) */
    uint8_t* out = (uint8_t*)calloc((size_t)H * W, 1);
    for (int x = 0; x < W; x++) {
        if (lab[x] >= 0) out[lab[x]] = 1;
        if (lab[(size_t)(H - 1) * W + x] >= 0) out[lab[(size_t)(H - 1) * W + x]] = 1;
    }
    for (int y = 0; y < H; y++) {
        if (lab[(size_t)y * W] >= 0) out[lab[(size_t)y * W]] = 1;
        if (lab[(size_t)y * W + W - 1] >= 0) out[lab[(size_t)y * W + W - 1]] = 1;
    }
    for (int i = 0; i < H * W; i++)
        if (lab[i] >= 0 && out[lab[i]]) lab[i] = -2;
    free(out);
}

/* ------------------------------------------------------------------------------------------ */
/* convexHull of an integer point set, in the vertex order cv2.convexHull(clockwise=False) uses  */
/* up to its cyclic shift: strict turns only, start at the lexicographic maximum (max x, max y). */
/* pts: n x 2 int32 (any order, duplicates allowed).  Returns hull size, hull: m x 2 float.      */
/* ------------------------------------------------------------------------------------------ */
static int cmp_pt(const void* a, const void* b)
{
    const int32_t* p = (const int32_t*)a; const int32_t* q = (const int32_t*)b;
    if (p[0] != q[0]) return p[0] < q[0] ? -1 : 1;
    if (p[1] != q[1]) return p[1] < q[1] ? -1 : 1;
    return 0;
}
static inline int64_t cross3(const int32_t* o, const int32_t* a, const int32_t* b)
{
    return (int64_t)(a[0] - o[0]) * (b[1] - o[1]) - (int64_t)(a[1] - o[1]) * (b[0] - o[0]);
}
EXPORT int orc_hull(const int32_t* pts_in, int n, float* hull)
{
    int32_t* pts = (int32_t*)malloc((size_t)n * 2 * sizeof(int32_t));
    memcpy(pts, pts_in, (size_t)n * 2 * sizeof(int32_t));
    qsort(pts, n, 2 * sizeof(int32_t), cmp_pt);
    int m = 0;
    for (int i = 0; i < n; i++)
        if (!m || cmp_pt(pts + 2 * i, pts + 2 * (m - 1))) { pts[2 * m] = pts[2 * i]; pts[2 * m + 1] = pts[2 * i + 1]; m++; }
    if (m <= 2) {
        /* cv2 order for 2 points: (max) then (min) */
        for (int i = 0; i < m; i++) { hull[2 * i] = (float)pts[2 * (m - 1 - i)]; hull[2 * i + 1] = (float)pts[2 * (m - 1 - i) + 1]; }
        free(pts);
        return m;
    }
    int32_t* st = (int32_t*)malloc((size_t)(2 * m + 2) * sizeof(int32_t));
    int k = 0;
    for (int i = 0; i < m; i++) {            /* lower chain (y-down image coordinates) */
        while (k >= 2 && cross3(pts + 2 * st[k - 2], pts + 2 * st[k - 1], pts + 2 * i) <= 0) k--;
        st[k++] = i;
    }
    int lo = k + 1;
    for (int i = m - 2; i >= 0; i--) {       /* upper chain */
        while (k >= lo && cross3(pts + 2 * st[k - 2], pts + 2 * st[k - 1], pts + 2 * i) <= 0) k--;
        st[k++] = i;
    }
    k--;                                      /* last == first */
    /* rotate so the lexicographic maximum (index m-1 of the sorted set) comes first */
    int s = 0;
    for (int i = 0; i < k; i++) if (st[i] == m - 1) s = i;
    for (int i = 0; i < k; i++) {
        int j = st[(s + i) % k];
        hull[2 * i] = (float)pts[2 * j]; hull[2 * i + 1] = (float)pts[2 * j + 1];
    }
    free(st); free(pts);
    return k;
}

/* ------------------------------------------------------------------------------------------ */
/* minAreaRect of a hull given in cv2.convexHull(clockwise=False) order.  out5 = cx,cy,w,h,angle */
/* Rotating calipers, cv2 4.13 flavour: the edge with the smallest angle is chosen by exact      */
/* cross products of the (rotated) edge vectors, not by comparing cosines.                      */
/* ------------------------------------------------------------------------------------------ */
EXPORT void orc_min_area_rect(const float* pts, int n, float* out5)
{
    out5[0] = out5[1] = out5[2] = out5[3] = out5[4] = 0.f;
    if (n == 1) { out5[0] = pts[0]; out5[1] = pts[1]; out5[4] = -90.f; return; }  /* 0 deg -> normalised to -90 */
    if (n == 2) {
        out5[0] = (pts[0] + pts[2]) * 0.5f; out5[1] = (pts[1] + pts[3]) * 0.5f;
        double dx = pts[2] - pts[0], dy = pts[3] - pts[1];
        float w = (float)sqrt(dx * dx + dy * dy), h = 0.f;
        double ang = atan2(dy, dx) * 180 / M_PI;
        /* same normalisation as below (verified against cv2 in tests) */
        while (ang >= 0) { float t = w; w = h; h = t; ang -= 90.0; }
        while (ang < -90) { float t = w; w = h; h = t; ang += 90.0; }
        out5[2] = w; out5[3] = h; out5[4] = (float)ang;
        return;
    }
    if (n < 1) return;
    float* inv = (float*)malloc(sizeof(float) * n * 3);
    float* vect = inv + n;
    int left = 0, bottom = 0, right = 0, top = 0, seq[4];
    float orientation = 0, base_a, base_b = 0;
    float p0x = pts[0], p0y = pts[1];
    float left_x = p0x, right_x = p0x, top_y = p0y, bottom_y = p0y;
    for (int i = 0; i < n; i++) {
        if (p0x < left_x) { left_x = p0x; left = i; }
        if (p0x > right_x) { right_x = p0x; right = i; }
        if (p0y > top_y) { top_y = p0y; top = i; }
        if (p0y < bottom_y) { bottom_y = p0y; bottom = i; }
        int j = (i + 1 < n) ? i + 1 : 0;
        float px = pts[2 * j], py = pts[2 * j + 1];
        double dx = px - p0x, dy = py - p0y;
        vect[2 * i] = (float)dx; vect[2 * i + 1] = (float)dy;
        inv[i] = (float)(1. / sqrt(dx * dx + dy * dy));
        p0x = px; p0y = py;
    }
    {
        double ax = vect[2 * (n - 1)], ay = vect[2 * (n - 1) + 1];
        for (int i = 0; i < n; i++) {
            double bx = vect[2 * i], by = vect[2 * i + 1];
            double c = ax * by - ay * bx;
            if (c != 0) { orientation = c > 0 ? 1.f : -1.f; break; }
            ax = bx; ay = by;
        }
    }
    base_a = orientation;
    seq[0] = bottom; seq[1] = right; seq[2] = top; seq[3] = left;
    float minarea = FLT_MAX, bA = 0, bB = 0, bW = 0, bH = 0; int bL = 0, bBt = 0;
    for (int k = 0; k < n; k++) {
        int me = 0;
        float rv[4][2];
        rv[0][0] = vect[2 * seq[0]];      rv[0][1] = vect[2 * seq[0] + 1];
        rv[1][0] = vect[2 * seq[1] + 1];  rv[1][1] = -vect[2 * seq[1]];
        rv[2][0] = -vect[2 * seq[2]];     rv[2][1] = -vect[2 * seq[2] + 1];
        rv[3][0] = -vect[2 * seq[3] + 1]; rv[3][1] = vect[2 * seq[3]];
        for (int i = 1; i < 4; i++) {
            float tx = rv[i][1], ty = -rv[i][0];
            if (tx * rv[me][0] + ty * rv[me][1] < 0) me = i;
        }
        int pi = seq[me];
        float lx = vect[2 * pi] * inv[pi], ly = vect[2 * pi + 1] * inv[pi];
        switch (me) {
        case 0: base_a = lx; base_b = ly; break;
        case 1: base_a = ly; base_b = -lx; break;
        case 2: base_a = -lx; base_b = -ly; break;
        default: base_a = -ly; base_b = lx; break;
        }
        seq[me] += 1; if (seq[me] == n) seq[me] = 0;
        float dx = pts[2 * seq[1]] - pts[2 * seq[3]], dy = pts[2 * seq[1] + 1] - pts[2 * seq[3] + 1];
        float width = dx * base_a + dy * base_b;
        dx = pts[2 * seq[2]] - pts[2 * seq[0]]; dy = pts[2 * seq[2] + 1] - pts[2 * seq[0] + 1];
        float height = -dx * base_b + dy * base_a;
        float area = width * height;
        if (area <= minarea) { minarea = area; bL = seq[3]; bA = base_a; bW = width; bB = base_b; bH = height; bBt = seq[0]; }
    }
    float A1 = bA, B1 = bB, A2 = -bB, B2 = bA;
    float C1 = A1 * pts[2 * bL] + pts[2 * bL + 1] * B1;
    float C2 = A2 * pts[2 * bBt] + pts[2 * bBt + 1] * B2;
    float idet = 1.f / (A1 * B2 - A2 * B1);
    float px = (C1 * B2 - C2 * B1) * idet, py = (A1 * C2 - A2 * C1) * idet;
    float o2 = A1 * bW, o3 = B1 * bW, o4 = A2 * bH, o5 = B2 * bH;
    out5[0] = px + (o2 + o4) * 0.5f;
    out5[1] = py + (o3 + o5) * 0.5f;
    float w = (float)sqrt((double)o2 * o2 + (double)o3 * o3);
    float h = (float)sqrt((double)o4 * o4 + (double)o5 * o5);
    double ang = atan2((double)o3, (double)o2) * 180 / M_PI;
    while (ang >= 0) { float t = w; w = h; h = t; ang -= 90.0; }
    while (ang < -90) { float t = w; w = h; h = t; ang += 90.0; }
    out5[2] = w; out5[3] = h; out5[4] = (float)ang;
    free(inv);
}

/* boxPoints(rect): 4 corners, x0,y0,...,x3,y3 */
EXPORT void orc_box_points(const float* r5, float* out8)
{
    double ang = r5[4] * M_PI / 180.;
    float b = (float)cos(ang) * 0.5f, a = (float)sin(ang) * 0.5f;
    float cx = r5[0], cy = r5[1], w = r5[2], h = r5[3];
    out8[0] = cx - a * h - b * w; out8[1] = cy + b * h - a * w;
    out8[2] = cx + a * h - b * w; out8[3] = cy - b * h - a * w;
    out8[4] = 2 * cx - out8[0];   out8[5] = 2 * cy - out8[1];
    out8[6] = 2 * cx - out8[2];   out8[7] = 2 * cy - out8[3];
}

/* ------------------------------------------------------------------------------------------ */
/* clipLine / line(LINE_8) / fillPoly on a uint8 image                                          */
/* ------------------------------------------------------------------------------------------ */
static int clip_line(int64_t W, int64_t H, int64_t* x1, int64_t* y1, int64_t* x2, int64_t* y2)
{
    int64_t right = W - 1, bottom = H - 1;
    if (W <= 0 || H <= 0) return 0;
    int c1 = (*x1 < 0) + (*x1 > right) * 2 + (*y1 < 0) * 4 + (*y1 > bottom) * 8;
    int c2 = (*x2 < 0) + (*x2 > right) * 2 + (*y2 < 0) * 4 + (*y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            *x1 += (int64_t)((double)(a - *y1) * (*x2 - *x1) / (*y2 - *y1));
            *y1 = a;
            c1 = (*x1 < 0) + (*x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            *x2 += (int64_t)((double)(a - *y2) * (*x2 - *x1) / (*y2 - *y1));
            *y2 = a;
            c2 = (*x2 < 0) + (*x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                *y1 += (int64_t)((double)(a - *x1) * (*y2 - *y1) / (*x2 - *x1));
                *x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                *y2 += (int64_t)((double)(a - *x2) * (*y2 - *y1) / (*x2 - *x1));
                *x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

EXPORT int orc_clip_line(int W, int H, int64_t* p) { return clip_line(W, H, p, p + 1, p + 2, p + 3); }

EXPORT void orc_line8(uint8_t* img, int H, int W, int x1_, int y1_, int x2_, int y2_, int val)
{
    int64_t x1 = x1_, y1 = y1_, x2 = x2_, y2 = y2_;
    if ((uint64_t)x1 >= (uint64_t)W || (uint64_t)x2 >= (uint64_t)W || (uint64_t)y1 >= (uint64_t)H || (uint64_t)y2 >= (uint64_t)H)
        if (!clip_line(W, H, &x1, &y1, &x2, &y2)) return;
    int dx = (int)(x2 - x1), dy = (int)(y2 - y1), sx = 1, sy = 1;
    int px = (int)x1, py = (int)y1;
    if (dx < 0) { dx = -dx; dy = -dy; px = (int)x2; py = (int)y2; }
    if (dy < 0) { dy = -dy; sy = -1; }
    int vert = dy > dx;
    if (vert) { int t = dx; dx = dy; dy = t; }
    int err = dx - (dy + dy), plusDelta = dx + dx, minusDelta = -(dy + dy), count = dx + 1;
    for (int i = 0; i < count; i++) {
        img[(size_t)py * W + px] = (uint8_t)val;
        int mask = err < 0 ? -1 : 0;
        err += minusDelta + (plusDelta & mask);
        if (!vert) { px += sx; py += sy & mask; }
        else { py += sy; px += sx & mask; }
    }
}

typedef struct { int y0, y1; int64_t x, dx; } PolyEdge;

static int cmp_edge(const void* a, const void* b)
{
    const PolyEdge* e1 = (const PolyEdge*)a; const PolyEdge* e2 = (const PolyEdge*)b;
    if (e1->y0 != e2->y0) return e1->y0 < e2->y0 ? -1 : 1;
    if (e1->x != e2->x) return e1->x < e2->x ? -1 : 1;
    if (e1->dx != e2->dx) return e1->dx < e2->dx ? -1 : 1;
    return 0;
}

/* fillPoly(img, [poly], val), LINE_8, shift 0.  poly: n x 2 int32 */
EXPORT void orc_fill_poly(uint8_t* img, int H, int W, const int32_t* v, int n, int val)
{
    enum { XY_SHIFT = 16 };
    const int64_t XY_ONE = 1 << XY_SHIFT;
    PolyEdge* edges = (PolyEdge*)malloc(sizeof(PolyEdge) * (n + 1));
    int ne = 0;
    int64_t p0x = (int64_t)v[2 * (n - 1)] << XY_SHIFT, p0y = v[2 * (n - 1) + 1];
    for (int i = 0; i < n; i++) {
        int64_t p1x = (int64_t)v[2 * i] << XY_SHIFT, p1y = v[2 * i + 1];
        int64_t t0x = (p0x + (XY_ONE >> 1)) >> XY_SHIFT, t0y = p0y;
        int64_t t1x = (p1x + (XY_ONE >> 1)) >> XY_SHIFT, t1y = p1y;
        int64_t c0x = p0x, c0y = p0y, c1x = p1x, c1y = p1y;
        orc_line8(img, H, W, (int)t0x, (int)t0y, (int)t1x, (int)t1y, val);
        if ((uint64_t)t0x >= (uint64_t)W || (uint64_t)t1x >= (uint64_t)W ||
            (uint64_t)t0y >= (uint64_t)H || (uint64_t)t1y >= (uint64_t)H) {
            clip_line(W, H, &t0x, &t0y, &t1x, &t1y);
            if (t0y != t1y) { c0y = t0y; c1y = t1y; }
            c0x = t0x << XY_SHIFT; c1x = t1x << XY_SHIFT;
        }
        if (p0y != p1y) {
            PolyEdge e;
            e.dx = (c1x - c0x) / (c1y - c0y);
            if (p0y < p1y) { e.y0 = (int)p0y; e.y1 = (int)p1y; e.x = c0x + (p0y - c0y) * e.dx; }
            else { e.y0 = (int)p1y; e.y1 = (int)p0y; e.x = c1x + (p1y - c1y) * e.dx; }
            edges[ne++] = e;
        }
        p0x = p1x; p0y = p1y;
    }
    if (ne >= 2) {
        int ymin = INT32_MAX, ymax = INT32_MIN;
        for (int i = 0; i < ne; i++) { if (edges[i].y0 < ymin) ymin = edges[i].y0; if (edges[i].y1 > ymax) ymax = edges[i].y1; }
        qsort(edges, ne, sizeof(PolyEdge), cmp_edge);
        if (ymax > H) ymax = H;
        int64_t xs[16];
        for (int y = ymin; y < ymax; y++) {
            int na = 0;
            for (int i = 0; i < ne && na < 16; i++)
                if (edges[i].y0 <= y && y < edges[i].y1) xs[na++] = edges[i].x + (int64_t)(y - edges[i].y0) * edges[i].dx;
            for (int i = 1; i < na; i++) { int64_t t = xs[i]; int j = i - 1; while (j >= 0 && xs[j] > t) { xs[j + 1] = xs[j]; j--; } xs[j + 1] = t; }
            if (y < 0) continue;
            for (int i = 0; i + 1 < na; i += 2) {
                int x1 = (int)((xs[i] + XY_ONE - 1) >> XY_SHIFT), x2 = (int)(xs[i + 1] >> XY_SHIFT);
                if (x1 < W && x2 >= 0) {
                    if (x1 < 0) x1 = 0;
                    if (x2 >= W) x2 = W - 1;
                    for (int x = x1; x <= x2; x++) img[(size_t)y * W + x] = (uint8_t)val;
                }
            }
        }
    }
    free(edges);
}

/* ------------------------------------------------------------------------------------------ */
/* HoughLines (standard), accumulator exposed.  Returns number of lines written (<= max_lines).  */
/* accum must hold (numangle+2)*(numrho+2) int32 (query sizes with orc_hough_dims first).        */
/* ------------------------------------------------------------------------------------------ */
EXPORT void orc_hough_dims(int H, int W, double rho_, double theta_, int* numangle, int* numrho)
{
    float rho = (float)rho_, theta = (float)theta_;
    int max_rho = W + H, min_rho = -max_rho;
    double min_theta = 0, max_theta = M_PI;
    int na = (int)floor((max_theta - min_theta) / theta) + 1;
    if (na > 1 && fabs(M_PI - (na - 1) * theta) < theta / 2) --na;
    *numangle = na;
    *numrho = cv_round(((max_rho - min_rho) + 1) / rho);
}

EXPORT void orc_hough_tables(int numangle, double rho_, double theta_, float* tabSin, float* tabCos)
{
    float rho = (float)rho_, theta = (float)theta_;
    float irho = 1 / rho;
    float ang = 0.f;
    for (int n = 0; n < numangle; ang += theta, n++) {
        tabSin[n] = (float)(sin((double)ang) * irho);
        tabCos[n] = (float)(cos((double)ang) * irho);
    }
}

typedef struct { int32_t votes; int32_t idx; } Peak;
static int cmp_peak(const void* a, const void* b)
{
    const Peak* p = (const Peak*)a; const Peak* q = (const Peak*)b;
    if (p->votes != q->votes) return p->votes > q->votes ? -1 : 1;
    return p->idx < q->idx ? -1 : (p->idx > q->idx);
}

EXPORT int orc_hough_lines(const uint8_t* img, int H, int W, double rho_, double theta_, int threshold,
                           int32_t* accum, float* lines /* 2*max_lines */, int32_t* line_votes, int max_lines)
{
    float rho = (float)rho_, theta = (float)theta_;
    int numangle, numrho;
    orc_hough_dims(H, W, rho_, theta_, &numangle, &numrho);
    float* tabSin = (float*)malloc(sizeof(float) * numangle * 2);
    float* tabCos = tabSin + numangle;
    orc_hough_tables(numangle, rho_, theta_, tabSin, tabCos);
    const int RS = numrho + 2;
    memset(accum, 0, sizeof(int32_t) * (size_t)(numangle + 2) * RS);
    for (int i = 0; i < H; i++)
        for (int j = 0; j < W; j++)
            if (img[(size_t)i * W + j] != 0)
                for (int n = 0; n < numangle; n++) {
                    float a = (float)j * tabCos[n];
                    float b = (float)i * tabSin[n];
                    int r = cv_roundf(a + b);
                    r += (numrho - 1) / 2;
                    accum[(size_t)(n + 1) * RS + r + 1]++;
                }
    Peak* pk = (Peak*)malloc(sizeof(Peak) * (size_t)numangle * numrho);
    int np_ = 0;
    for (int r = 0; r < numrho; r++)
        for (int n = 0; n < numangle; n++) {
            int base = (n + 1) * RS + r + 1;
            if (accum[base] > threshold && accum[base] > accum[base - 1] && accum[base] >= accum[base + 1] &&
                accum[base] > accum[base - RS] && accum[base] >= accum[base + RS]) {
                pk[np_].votes = accum[base]; pk[np_].idx = base; np_++;
            }
        }
    qsort(pk, np_, sizeof(Peak), cmp_peak);
    int nl = np_ < max_lines ? np_ : max_lines;
    double scale = 1. / RS;
    for (int i = 0; i < nl; i++) {
        int idx = pk[i].idx;
        int n = (int)floor(idx * scale) - 1;
        int r = idx - (n + 1) * RS - 1;
        lines[2 * i] = (r - (numrho - 1) / 2) * rho;
        lines[2 * i + 1] = n * theta;
        if (line_votes) line_votes[i] = pk[i].votes;
    }
    free(pk); free(tabSin);
    return np_;
}
