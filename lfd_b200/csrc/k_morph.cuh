// k_morph.cuh - erode / dilate with all-ones rectangular kernels, fused with the equalisation LUT.
//
// Reference: cv2.erode(equ, erodeKernel) processfield.py:464 ; cv2.dilate(., dilateKernel) :354, :471
// (paths under /root/reference/lfd/detecttrails/).  cv2 semantics: anchor = (kw/2, kh/2), window rows
// y-ay .. y-ay+kh-1, out-of-frame samples ignored.
//
// equalizeHist's LUT is monotone non-decreasing, so min/max filters commute with it:
// dilate(erode(lut[gray])) == lut[dilate(erode(gray))].  The kernel therefore filters the raw `gray`
// plane in shared memory and applies the LUT once on the way out; the equalised plane itself is only
// materialised for the debug tap.
//
// Tile: 128 x 16 output pixels per 256-thread CTA, 4 pixels (one 32-bit word) per thread and row,
// byte-wise SIMD min/max (__vminu4/__vmaxu4) with funnel shifts for the horizontal window.
#pragma once
#include "common.cuh"

#define MORPH_TW 128                // tile width in pixels
#define MORPH_TWW (MORPH_TW / 4)    // in 4-pixel words
#define MORPH_TH 16
#define MORPH_PADW 4                // halo words each side (16 px)
#define MORPH_PADH 16               // halo rows each side
#define MORPH_SW (MORPH_TWW + 2 * MORPH_PADW)
#define MORPH_SH (MORPH_TH + 2 * MORPH_PADH)

struct MorphCfg {
    int eh, ew;     // erode kernel (0 = none)
    int dh, dw;     // dilate kernel (>= 1)
};

// horizontal window op over bytes: out byte x = op over x+lo .. x+hi of row `s` (word array with PADW halo)
template <bool IS_MAX>
__device__ __forceinline__ u32 hwin(const u32* s, int w, int lo, int hi)
{
    u32 acc = IS_MAX ? 0u : 0xffffffffu;
    for (int o = lo; o <= hi; o++) {
        int wi = w + (o >> 2);          // arithmetic shift = floor
        int b = o & 3;
        u32 v = __funnelshift_r(s[wi], s[wi + 1], 8 * b);
        acc = IS_MAX ? __vmaxu4(acc, v) : __vminu4(acc, v);
    }
    return acc;
}

// gray -> [erode] -> dilate -> lut -> morph (uint8) + nz (bit mask of morph != 0) [+ eroded tap (lut applied)]
__global__ void __launch_bounds__(256)
k_morph(const u8* __restrict__ gray, const u8* __restrict__ lut, u8* __restrict__ morph,
        u32* __restrict__ nz, u8* __restrict__ eroded_tap, const FrameCtl* __restrict__ ctl, int pass,
        Dims d, MorphCfg mc)
{
    int f = blockIdx.z;
    if (!ctl[f].active[pass]) return;
    __shared__ u32 A[MORPH_SH][MORPH_SW + 1];
    __shared__ u32 Bf[MORPH_SH][MORPH_SW + 1];
    __shared__ u8 slut[256];
    const int tx0w = blockIdx.x * MORPH_TWW;      // tile origin in words
    const int ty0 = blockIdx.y * MORPH_TH;
    const int Ww = d.W >> 2;
    const u32* g = reinterpret_cast<const u32*>(gray + (size_t)f * d.N);
    slut[threadIdx.x] = lut[(size_t)f * 256 + threadIdx.x];

    const bool has_e = mc.eh > 0;
    // window extents (relative offsets, inclusive)
    const int e_t = has_e ? -(mc.eh / 2) : 0, e_b = has_e ? mc.eh - 1 - mc.eh / 2 : 0;
    const int e_l = has_e ? -(mc.ew / 2) : 0, e_r = has_e ? mc.ew - 1 - mc.ew / 2 : 0;
    const int d_t = -(mc.dh / 2), d_b = mc.dh - 1 - mc.dh / 2;
    const int d_l = -(mc.dw / 2), d_r = mc.dw - 1 - mc.dw / 2;

    // 1. load rows [ty0 - PADH, ty0 + TH + PADH) x words [tx0w - PADW, tx0w + TWW + PADW); outside the
    //    frame -> identity of the FIRST filter
    const u32 ident_first = has_e ? 0xffffffffu : 0u;
    for (int i = threadIdx.x; i < MORPH_SH * MORPH_SW; i += blockDim.x) {
        int r = i / MORPH_SW, c = i - r * MORPH_SW;
        int y = ty0 - MORPH_PADH + r, w = tx0w - MORPH_PADW + c;
        u32 v = ident_first;
        if (y >= 0 && y < d.H && w >= 0 && w < Ww) v = g[(size_t)y * Ww + w];
        A[r][c] = v;
    }
    __syncthreads();

    if (has_e) {
        // 2a. vertical min: rows needed by the dilate = [ty0 + d_t, ty0 + TH - 1 + d_b]
        int r_lo = MORPH_PADH + d_t, r_hi = MORPH_PADH + MORPH_TH - 1 + d_b;
        int nrows = r_hi - r_lo + 1;
        for (int i = threadIdx.x; i < nrows * MORPH_SW; i += blockDim.x) {
            int r = r_lo + i / MORPH_SW, c = i % MORPH_SW;
            u32 acc = 0xffffffffu;
            for (int o = e_t; o <= e_b; o++) acc = __vminu4(acc, A[r + o][c]);
            Bf[r][c] = acc;
        }
        __syncthreads();
        // 2b. horizontal min into A; positions outside the frame become 0 (ignored by the dilate)
        // columns the dilate will read (host validates that the combined halo fits the padding)
        int c_lo = MORPH_PADW + (d_l >> 2), c_hi = MORPH_PADW + MORPH_TWW - 1 + ((d_r + 3) >> 2);
        for (int i = threadIdx.x; i < nrows * (c_hi - c_lo + 1); i += blockDim.x) {
            int r = r_lo + i / (c_hi - c_lo + 1), c = c_lo + i % (c_hi - c_lo + 1);
            int y = ty0 - MORPH_PADH + r, w = tx0w - MORPH_PADW + c;
            u32 v = 0u;
            if (y >= 0 && y < d.H && w >= 0 && w < Ww) v = hwin<false>(&Bf[r][0], c, e_l, e_r);
            A[r][c] = v;
        }
        __syncthreads();
        if (eroded_tap) {
            for (int i = threadIdx.x; i < MORPH_TH * MORPH_TWW; i += blockDim.x) {
                int r = i / MORPH_TWW, c = i % MORPH_TWW;
                int y = ty0 + r, w = tx0w + c;
                if (y < d.H && w < Ww) {
                    u32 v = A[MORPH_PADH + r][MORPH_PADW + c];
                    u32 o = slut[v & 255] | (slut[(v >> 8) & 255] << 8) | (slut[(v >> 16) & 255] << 16) | (slut[v >> 24] << 24);
                    reinterpret_cast<u32*>(eroded_tap + (size_t)f * d.N)[(size_t)y * Ww + w] = o;
                }
            }
        }
    }

    // 3a. vertical max over the output rows (all columns incl. halo)
    for (int i = threadIdx.x; i < MORPH_TH * MORPH_SW; i += blockDim.x) {
        int r = MORPH_PADH + i / MORPH_SW, c = i % MORPH_SW;
        u32 acc = 0u;
        for (int o = d_t; o <= d_b; o++) acc = __vmaxu4(acc, A[r + o][c]);
        Bf[r][c] = acc;
    }
    __syncthreads();
    // 3b. horizontal max + LUT + outputs.  One warp covers one tile row (32 words = 128 px).
    for (int i = threadIdx.x; i < MORPH_TH * MORPH_TWW; i += blockDim.x) {
        int r = i / MORPH_TWW, c = i % MORPH_TWW;   // c == lane
        int y = ty0 + r, w = tx0w + c;
        bool in = (y < d.H && w < Ww);
        u32 o = 0u;
        if (in) {
            u32 v = hwin<true>(&Bf[MORPH_PADH + r][0], MORPH_PADW + c, d_l, d_r);
            o = slut[v & 255] | (slut[(v >> 8) & 255] << 8) | (slut[(v >> 16) & 255] << 16) | (slut[v >> 24] << 24);
            reinterpret_cast<u32*>(morph + (size_t)f * d.N)[(size_t)y * Ww + w] = o;
        }
        // non-zero bit mask: 8 lanes x 4 px = one 32-bit mask word
        u32 b0 = __ballot_sync(FULLMASK, (o & 0x000000ffu) != 0);
        u32 b1 = __ballot_sync(FULLMASK, (o & 0x0000ff00u) != 0);
        u32 b2 = __ballot_sync(FULLMASK, (o & 0x00ff0000u) != 0);
        u32 b3 = __ballot_sync(FULLMASK, (o & 0xff000000u) != 0);
        int lane = lane_id();
        if (lane < 4) {
            u32 word = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                int l = lane * 8 + k;
                word |= ((b0 >> l) & 1u) << (4 * k);
                word |= ((b1 >> l) & 1u) << (4 * k + 1);
                word |= ((b2 >> l) & 1u) << (4 * k + 2);
                word |= ((b3 >> l) & 1u) << (4 * k + 3);
            }
            int mw = (tx0w * 4) / 32 + lane;      // mask word index in the row
            if (y < d.H && mw < d.WW) nz[(size_t)f * d.NW + (size_t)y * d.WW + mw] = word;
        }
    }
}
