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

// ================================================================================================
// Register-marching variant (production path when W % 8 == 0), same strip geometry as k_morph_march:
// a warp owns 64 4-pixel words (lanes 2..29 produce output: 224 px = 7 mask words) and marches down
// NMS_R rows.  Sobel runs on u16x2 pairs (VIADD.16x2 / VIMNMX.S16x2): per arriving image row the
// horizontal [1,2,1] sum and the horizontal difference are formed once and kept in 3-row register
// rings, dx/dy/magnitude are their vertical combinations; the direction test of OpenCV (32-bit
// integer tan 22.5) is scalar per pixel and only runs on rows where some pixel has a gradient.
// Empty strip chunks (no non-zero input pixel, from the morphology kernel's 1-bit mask) are skipped.
// ================================================================================================
#include "k_morph.cuh"

#ifndef NMS_R
#define NMS_R 32
#endif
#ifndef NMS_WPC
#define NMS_WPC 1                // warps (= strip units) per CTA: units differ a lot in work (skips), so small CTAs free their slot sooner
#endif

__device__ __forceinline__ u32 vneg2(u32 a) { return __vadd2(~a, 0x00010001u); }
__device__ __forceinline__ u32 vsub2(u32 a, u32 b) { return __vadd2(a, vneg2(b)); }
__device__ __forceinline__ u32 vabs2s(u32 a) { return __vmaxs2(a, vneg2(a)); }

// cand / strong must be zero on entry (the caller memsets them): only non-zero mask words are stored.
template <bool TAP>
__global__ void __launch_bounds__(NMS_WPC * 32)
k_nms_march(const u8* __restrict__ img, const u32* __restrict__ nz, u32* __restrict__ cand, u32* __restrict__ strong,
            u8* __restrict__ nms_tap, const FrameCtl* __restrict__ ctl, int pass, Dims d, int nstrips, int nunits,
            int low, int high)
{
    const int f = blockIdx.y;
    if (!ctl[f].active[pass]) return;
    const int unit = blockIdx.x * NMS_WPC + (threadIdx.x >> 5);
    if (unit >= nunits) return;
    // per warp: 3-row rings of magnitude / dx / dy as 16-bit planes of the 256-px strip.  The SIMD stage
    // writes them in "8 px per lane" layout, the decision stage reads them in "1 px per lane" layout, so
    // a 10-px run of gradient pixels occupies 10 lanes for one iteration instead of 2 lanes for 8.
    __shared__ __align__(16) u16 sM[NMS_WPC][3][256];
    __shared__ __align__(16) short sDX[NMS_WPC][3][256];
    __shared__ __align__(16) short sDY[NMS_WPC][3][256];
    const int wid = threadIdx.x >> 5;
    const int chunk = unit / nstrips, s = unit - chunk * nstrips;
    const int lane = lane_id();
    const int Ww = d.W >> 2;
    const int wx = s * MARCH_UW - MARCH_HW + 2 * lane;
    const bool col_in = wx >= 0 && wx < Ww;
    const int y0 = chunk * NMS_R, y1 = min(y0 + NMS_R, d.H);
    const int mw0 = s * (MARCH_UW / 8);                     // first mask word of the strip
    u32* candf = cand + (size_t)f * d.NW;
    u32* strongf = strong + (size_t)f * d.NW;
    const int nwords = min(MARCH_UW / 8, d.WW - mw0);        // mask words this strip owns (<= 7)

    // any non-zero input pixel in the window of this unit?  mask words 7s-1 .. 7s+7, rows y0-2 .. y1+1
    {
        const int ra = max(y0 - 2, 0), rb = min(y1 + 1, d.H - 1);
        const int wa = max(mw0 - 1, 0), wb = min(mw0 + 7, d.WW - 1);
        const int nw = wb - wa + 1, tot = (rb - ra + 1) * nw;
        u32 any = 0;
        const u32* nzf = nz + (size_t)f * d.NW + wa;
        if (nw == 9) {                                      // interior strips: constant divisor (mul-shift, no division)
            for (int i = lane; i < tot; i += 32) any |= nzf[(ra + i / 9) * d.WW + i % 9];
        } else {
            for (int i = lane; i < tot; i += 32) any |= nzf[(ra + i / nw) * d.WW + i % nw];
        }
        if (!__any_sync(FULLMASK, any != 0)) {
            // nothing to do: the candidate / strong masks are zeroed by the caller (cudaMemsetAsync) before the launch,
            // this kernel only ever stores non-zero mask words (zero rows were 20 % of its executed instructions)
            if (TAP)
                for (int y = y0; y < y1; y++)
                    for (int x = lane; x < 32 * nwords; x += 32)
                        if (32 * mw0 + x < d.W) nms_tap[(size_t)f * d.N + (size_t)y * d.W + 32 * mw0 + x] = 0;
            return;
        }
    }

    const uint2* g = reinterpret_cast<const uint2*>(img + (size_t)f * d.N);
    u32 HD[3][4], H3[3][4];
    bool nzf[3] = {false, false, false};          // this lane has a gradient pixel in the row of slot k
    bool nzf_any[3] = {false, false, false};      // ... some lane of the warp has (the smem row of slot k is not all-zero)
    bool rz[3] = {false, false, false};           // input row of slot k is all-zero across the strip
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int c = 0; c < 4; c++) { HD[k][c] = 0; H3[k][c] = 0; }
    for (int i = lane; i < 3 * 256 / 2; i += 32) reinterpret_cast<u32*>(&sM[wid][0][0])[i] = 0u;
    __syncwarp();
    const int TG22 = 13573;
    // rows are requested NMS_PF steps ahead of their use (software pipeline; the kernel is latency-bound)
    // every lane issues the same kind of load (lanes left / right of the frame read the edge word pair), so the
    // prefetch covers them too; BORDER_REPLICATE is applied when the row is consumed (fix_row)
    const int wxc = min(max(wx, 0), Ww - 2);
    auto load_row = [&](int yi) -> uint2 {
        const int yl = min(max(yi, 0), d.H - 1);                     // BORDER_REPLICATE rows
        return __ldg(g + ((yl * Ww + wxc) >> 1));         // per-frame offsets fit 32 bits (H * W <= 2^28)
    };
    auto fix_row = [&](uint2 v) -> uint2 {
        if (!col_in) {                                               // BORDER_REPLICATE columns
            u32 e = (wx < 0) ? (v.x & 0xffu) : (v.y >> 24);
            v.x = v.y = e * 0x01010101u;
        }
        return v;
    };
    uint2 nxt[3];
#pragma unroll
    for (int u = 0; u < 3; u++) nxt[u] = load_row(y0 - 2 + u);
    for (int yb = y0 - 2; yb <= y1 + 1; yb += 3) {
        uint2 cur[3];
#pragma unroll
        for (int u = 0; u < 3; u++) cur[u] = nxt[u];
        if (yb + 3 <= y1 + 1) {
#pragma unroll
            for (int u = 0; u < 3; u++) nxt[u] = load_row(yb + 3 + u);
        }
#pragma unroll
        for (int u = 0; u < 3; u++) {
            const int yi = yb + u;
            const uint2 v = fix_row(cur[u]);
            const int sa = (u + 1) % 3, sb = (u + 2) % 3, sc = u % 3;   // ring slots of rows yi-2, yi-1, yi
            rz[sc] = !__any_sync(FULLMASK, (v.x | v.y) != 0u);
            if (rz[0] && rz[1] && rz[2]) {
                // three all-zero input rows across the strip: zero gradient on the centre row, nothing to suppress
#pragma unroll
                for (int c = 0; c < 4; c++) { HD[sc][c] = 0; H3[sc][c] = 0; }
                if (nzf_any[sb]) reinterpret_cast<uint4*>(&sM[wid][sb][0])[lane] = make_uint4(0u, 0u, 0u, 0u);
                nzf[sb] = false; nzf_any[sb] = false;
                const int yn = yi - 2;
                __syncwarp();
                if (yn >= y0 && yn < y1) {
                    if (!nzf_any[sa]) {
                        if (TAP)
                            for (int x = lane; x < 32 * nwords; x += 32)
                                if (32 * mw0 + x < d.W) nms_tap[(size_t)f * d.N + (size_t)yn * d.W + 32 * mw0 + x] = 0;
                        continue;
                    }
                } else continue;
            }
            u32 p[4];
            p[0] = __byte_perm(v.x, 0, 0x4140); p[1] = __byte_perm(v.x, 0, 0x4342);
            p[2] = __byte_perm(v.y, 0, 0x4140); p[3] = __byte_perm(v.y, 0, 0x4342);
            const u32 eL = __shfl_up_sync(FULLMASK, p[3], 1), eR = __shfl_down_sync(FULLMASK, p[0], 1);
            u32 S[5];                                                // S[j] = (px 2j-1, px 2j) of the lane's 8 pixels
            S[0] = __byte_perm(eL, p[0], 0x5432);
            S[1] = __byte_perm(p[0], p[1], 0x5432);
            S[2] = __byte_perm(p[1], p[2], 0x5432);
            S[3] = __byte_perm(p[2], p[3], 0x5432);
            S[4] = __byte_perm(p[3], eR, 0x5432);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                HD[sc][c] = vsub2(S[c + 1], S[c]);                                  // p(x+1) - p(x-1)
                H3[sc][c] = __vadd2(__vadd2(S[c], S[c + 1]), __vadd2(p[c], p[c]));  // p(x-1) + 2p(x) + p(x+1)
            }
            // gradient of the centre row yc = yi - 1 -> shared-memory slot sb
            const int yc = yi - 1;
            const bool cin = yc >= 0 && yc < d.H && col_in;
            u32 mg[4], dxv[4], dyv[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                dxv[c] = __vadd2(__vadd2(HD[sa][c], HD[sc][c]), __vadd2(HD[sb][c], HD[sb][c]));
                dyv[c] = vsub2(H3[sc][c], H3[sa][c]);
                mg[c] = cin ? __vadd2(vabs2s(dxv[c]), vabs2s(dyv[c])) : 0u;
            }
            nzf[sb] = (mg[0] | mg[1] | mg[2] | mg[3]) != 0u;
            nzf_any[sb] = __any_sync(FULLMASK, nzf[sb]);
            reinterpret_cast<uint4*>(&sM[wid][sb][0])[lane] = make_uint4(mg[0], mg[1], mg[2], mg[3]);
            if (nzf[sb]) {
                reinterpret_cast<uint4*>(&sDX[wid][sb][0])[lane] = make_uint4(dxv[0], dxv[1], dxv[2], dxv[3]);
                reinterpret_cast<uint4*>(&sDY[wid][sb][0])[lane] = make_uint4(dyv[0], dyv[1], dyv[2], dyv[3]);
            }
            __syncwarp();
            // non-maximum suppression of row yn = yi - 2 (rows yn-1, yn, yn+1 in slots sc, sa, sb)
            const int yn = yi - 2;
            if (yn >= y0 && yn < y1) {                               // warp-uniform
                const u32 bal = __ballot_sync(FULLMASK, nzf[sa]);
                u32 myc = 0, mys = 0;                                // lane g keeps mask word g
                const u16* Mu = sM[wid][sc]; const u16* Mc = sM[wid][sa]; const u16* Md = sM[wid][sb];
                // bit 4*gi of gm: some pixel of 32-px group gi has a gradient (its four lanes 2+4gi .. 5+4gi)
                u32 gm = bal >> 2;
                gm = (gm | (gm >> 1) | (gm >> 2) | (gm >> 3)) & 0x1111111u;
                if (TAP) gm = 0x1111111u;                            // the tap writes every pixel of the row
#pragma unroll 1
                while (gm) {
                    const int gi = (__ffs(gm) - 1) >> 2;
                    gm &= gm - 1;
                    int cls = 0;
                    if ((bal >> (2 + 4 * gi)) & 0xfu) {              // (always true unless TAP)
                        const int x = 16 + 32 * gi + lane;           // pixel index inside the 256-px strip
                        const int mm = Mc[x];
                        if (mm > low) {
                            const int dx = sDX[wid][sa][x], dy = sDY[wid][sa][x];
                            const int ax = abs(dx), ay = abs(dy) << 15;
                            const int tg22x = ax * TG22;
                            bool keep;
                            if (ay < tg22x) keep = mm > Mc[x - 1] && mm >= Mc[x + 1];
                            else {
                                const int tg67x = tg22x + (ax << 16);
                                if (ay > tg67x) keep = mm > Mu[x] && mm >= Md[x];
                                else {
                                    const int sg = ((dx ^ dy) < 0) ? -1 : 1;
                                    keep = mm > Mu[x - sg] && mm > Md[x + sg];
                                }
                            }
                            if (keep) cls = mm > high ? 2 : 1;
                        }
                        const u32 bc = __ballot_sync(FULLMASK, cls != 0), bs = __ballot_sync(FULLMASK, cls == 2);
                        if (lane == gi) { myc = bc; mys = bs; }
                    }
                    if (TAP) {
                        const int xg = 32 * (mw0 + gi) + lane;
                        if (gi < nwords && xg < d.W) nms_tap[(size_t)f * d.N + (size_t)yn * d.W + xg] = (u8)cls;
                    }
                }
                if (lane < nwords && myc) { const int o = yn * d.WW + mw0 + lane; candf[o] = myc; if (mys) strongf[o] = mys; }
            }
            __syncwarp();
        }
    }
}
