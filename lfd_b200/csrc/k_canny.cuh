// k_canny.cuh - cv2.Canny(img, 0, 255) front half: 3x3 Sobel, L1 magnitude and non-maximum
// suppression fused per tile, producing two bit masks (candidate, strong).  Hysteresis is done by the
// run-based connected-component pass in k_ccl.cuh (an edge component survives iff it holds a strong px).
//
// Reference call: cv2.Canny(img, 0, 255) at /root/reference/lfd/detecttrails/processfield.py:236.
// Semantics restated in oracle/c/cvrestate.c::orc_canny_classes (pinned against cv2 4.13):
// Sobel with BORDER_REPLICATE, magnitude 0 outside the frame, OpenCV's integer tan(22.5) sectors and
// its asymmetric > / >= tie rules.
#pragma once
#include "common.cuh"

#define CANNY_TW 128
#define CANNY_TH 16

__global__ void __launch_bounds__(256)
k_canny_nms(const u8* __restrict__ img, const u32* __restrict__ nz, u32* __restrict__ cand, u32* __restrict__ strong,
            u8* __restrict__ nms_tap, const FrameCtl* __restrict__ ctl, int pass, Dims d, int low, int high)
{
    int f = blockIdx.z;
    if (!ctl[f].active[pass]) return;
    // pixels: rows ty0-2 .. ty0+TH+1, cols tx0-2 .. tx0+TW+1
    __shared__ u8 P[CANNY_TH + 4][CANNY_TW + 8];
    __shared__ u16 M[CANNY_TH + 2][CANNY_TW + 4];     // magnitude for rows ty0-1.., cols tx0-1..
    const int tx0 = blockIdx.x * CANNY_TW, ty0 = blockIdx.y * CANNY_TH;
    const u8* src = img + (size_t)f * d.N;

    // A tile whose input window (tile + 2 px) holds no non-zero pixel has zero gradient everywhere:
    // consult the 1-bit/px non-zero mask the morphology kernel wrote (120 words) instead of loading
    // 2.6 KB of pixels.  Sky-subtracted frames are mostly such tiles.
    {
        int any = 0;
        if (threadIdx.x < 20 * 6) {
            int r = threadIdx.x / 6, c = threadIdx.x - r * 6;
            int y = min(max(ty0 - 2 + r, 0), d.H - 1);
            int w = (tx0 >> 5) - 1 + c;
            if (w >= 0 && w < d.WW) any = nz[(size_t)f * d.NW + (size_t)y * d.WW + w] != 0;
        }
        if (!__syncthreads_or(any)) {
            for (int i = threadIdx.x; i < CANNY_TH * (CANNY_TW / 32); i += blockDim.x) {
                int r = i / (CANNY_TW / 32), c = i - r * (CANNY_TW / 32);
                int y = ty0 + r, w = (tx0 >> 5) + c;
                if (y < d.H && w < d.WW) {
                    size_t o = (size_t)f * d.NW + (size_t)y * d.WW + w;
                    cand[o] = 0u; strong[o] = 0u;
                }
            }
            if (nms_tap)
                for (int i = threadIdx.x; i < CANNY_TH * CANNY_TW; i += blockDim.x) {
                    int y = ty0 + i / CANNY_TW, x = tx0 + i % CANNY_TW;
                    if (y < d.H && x < d.W) nms_tap[(size_t)f * d.N + (size_t)y * d.W + x] = 0;
                }
            return;
        }
    }

    for (int i = threadIdx.x; i < (CANNY_TH + 4) * (CANNY_TW + 4); i += blockDim.x) {
        int r = i / (CANNY_TW + 4), c = i - r * (CANNY_TW + 4);
        int y = min(max(ty0 - 2 + r, 0), d.H - 1), x = min(max(tx0 - 2 + c, 0), d.W - 1);   // BORDER_REPLICATE
        P[r][c] = src[(size_t)y * d.W + x];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (CANNY_TH + 2) * (CANNY_TW + 2); i += blockDim.x) {
        int r = i / (CANNY_TW + 2), c = i - r * (CANNY_TW + 2);
        int y = ty0 - 1 + r, x = tx0 - 1 + c;
        int m = 0;
        if (y >= 0 && y < d.H && x >= 0 && x < d.W) {
            // P index of (y, x) is [r+1][c+1]
            int a00 = P[r][c], a01 = P[r][c + 1], a02 = P[r][c + 2];
            int a10 = P[r + 1][c], a12 = P[r + 1][c + 2];
            int a20 = P[r + 2][c], a21 = P[r + 2][c + 1], a22 = P[r + 2][c + 2];
            int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
            int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
            m = abs(dx) + abs(dy);
        }
        M[r][c] = (u16)m;
    }
    __syncthreads();
    const int TG22 = 13573;   // (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5)
    // one warp per 32 consecutive pixels of a row: 4 warps per row, 2 rows per iteration
    for (int i = threadIdx.x; i < CANNY_TH * CANNY_TW; i += blockDim.x) {
        int r = i / CANNY_TW, c = i - r * CANNY_TW;
        int y = ty0 + r, x = tx0 + c;
        int cls = 0;
        if (y < d.H && x < d.W) {
            int mm = M[r + 1][c + 1];
            if (mm > low) {
                int a00 = P[r + 1][c + 1], a01 = P[r + 1][c + 2], a02 = P[r + 1][c + 3];
                int a10 = P[r + 2][c + 1], a12 = P[r + 2][c + 3];
                int a20 = P[r + 3][c + 1], a21 = P[r + 3][c + 2], a22 = P[r + 3][c + 3];
                int dx = (a02 + 2 * a12 + a22) - (a00 + 2 * a10 + a20);
                int dy = (a20 + 2 * a21 + a22) - (a00 + 2 * a01 + a02);
                int ax = abs(dx), ay = abs(dy) << 15;
                int tg22x = ax * TG22;
                bool keep;
                if (ay < tg22x) {
                    keep = mm > M[r + 1][c] && mm >= M[r + 1][c + 2];
                } else {
                    int tg67x = tg22x + (ax << 16);
                    if (ay > tg67x) {
                        keep = mm > M[r][c + 1] && mm >= M[r + 2][c + 1];
                    } else {
                        int s = ((dx ^ dy) < 0) ? -1 : 1;
                        keep = mm > M[r][c + 1 - s] && mm > M[r + 2][c + 1 + s];
                    }
                }
                if (keep) cls = mm > high ? 2 : 1;
            }
            if (nms_tap) nms_tap[(size_t)f * d.N + (size_t)y * d.W + x] = (u8)cls;
        }
        u32 bc = __ballot_sync(FULLMASK, cls != 0);
        u32 bs = __ballot_sync(FULLMASK, cls == 2);
        if (lane_id() == 0 && y < d.H && x < d.W) {
            size_t o = (size_t)f * d.NW + (size_t)y * d.WW + (x >> 5);
            cand[o] = bc;
            strong[o] = bs;
        }
    }
}
