// TEST-ONLY: compiles lfd_b200/csrc/geom.cuh for the host so the CPU suite can check the device
// geometry source against cv2 without a GPU.  Never loaded by the product.
#include <stdlib.h>
#include <string.h>
#include "../lfd_b200/csrc/geom.cuh"
using namespace lfdgeom;
extern "C" {
// points: n x 2 int32 (x, y) -> per-row extremes -> hull -> rect.  out5 = cx,cy,w,h,angle; returns hull size
int hg_rect(const int* pts, int n, float* out5, int* box8)
{
    int ymin = 1 << 30, ymax = -(1 << 30);
    for (int i = 0; i < n; i++) { if (pts[2*i+1] < ymin) ymin = pts[2*i+1]; if (pts[2*i+1] > ymax) ymax = pts[2*i+1]; }
    int h = ymax - ymin + 1;
    int* rmin = (int*)malloc(sizeof(int) * h * 2); int* rmax = rmin + h;
    for (int i = 0; i < h; i++) { rmin[i] = 1 << 30; rmax[i] = -1; }
    for (int i = 0; i < n; i++) { int r = pts[2*i+1] - ymin; if (pts[2*i] < rmin[r]) rmin[r] = pts[2*i]; if (pts[2*i] > rmax[r]) rmax[r] = pts[2*i]; }
    Pt* st = (Pt*)malloc(sizeof(Pt) * (2 * h + 2));
    float* vect = (float*)malloc(sizeof(float) * (2 * h + 2) * 3);
    int start; int k = hull_from_rows(rmin, rmax, h, ymin, st, &start);
    Rect r; min_area_rect(st, k, start, vect, vect + 2 * (2 * h + 2), &r);
    out5[0] = r.cx; out5[1] = r.cy; out5[2] = r.w; out5[3] = r.h; out5[4] = r.angle;
    float f8[8]; box_points(r, f8, box8);
    free(rmin); free(st); free(vect);
    return k;
}
void hg_fill_quad(unsigned char* img, int H, int W, const int* q)
{
    Edge e[4];
    for (int i = 0; i < 4; i++) {
        int j = (i + 3) & 3, draw, x0, y0, x1, y1;
        poly_edge(W, H, q[2*j], q[2*j+1], q[2*i], q[2*i+1], &e[i], &draw, &x0, &y0, &x1, &y1);
        if (draw) { LineIt it; it.init(x0, y0, x1, y1); for (int s = 0; s < it.count; s++) { img[(size_t)it.y * W + it.x] = 255; it.next(); } }
    }
    int ymin = 1 << 30, ymax = -(1 << 30), nv = 0;
    for (int i = 0; i < 4; i++) if (e[i].valid) { nv++; if (e[i].y0 < ymin) ymin = e[i].y0; if (e[i].y1 > ymax) ymax = e[i].y1; }
    if (nv < 2) return;
    if (ymin < 0) ymin = 0;
    if (ymax > H) ymax = H;
    for (int y = ymin; y < ymax; y++) {
        int xs[4]; int ns = row_spans(e, 4, y, W, xs);
        for (int s = 0; s < ns; s++) for (int x = xs[2*s]; x <= xs[2*s+1]; x++) img[(size_t)y * W + x] = 255;
    }
}
}
