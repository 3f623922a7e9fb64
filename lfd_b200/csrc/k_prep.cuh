// k_prep.cuh - star mask rasterisation, fused blot+flip+clip+convertScaleAbs+histogram, equalisation LUT.
//
// Reference behaviour (paths under /root/reference/lfd/detecttrails/):
//   removestars.py:231      img[x-dxy:x+dxy, y-dxy:y+dxy].fill(0.0)  -> k_star_mask (+ applied in k_prep)
//   detecttrails.py:124     cv2.flip(img, 0)                          -> source row H-1-y in k_prep
//   processfield.py:342     img[img < 0] = 0                          -> bright clip
//   processfield.py:453-454 img[img < minFlux] = 0; img[img > 0] += addFlux
//   processfield.py:346,456 cv2.convertScaleAbs                       -> csa()
//   processfield.py:347,457 cv2.equalizeHist                          -> histogram here, LUT in k_lut
#pragma once
#include "common.cuh"

// rects: (r0, r1, c0, c1) half-open on the UN-flipped image; rect_off[n+1] indexes them per frame.
// mask: [n][H][WW] bit per pixel in the FLIPPED orientation (row H-1-r).
// One warp per rectangle over the flat list of ALL frames (the frame is found by bisection of rect_off),
// so the grid is proportional to the number of stars, not to frames x the densest frame.
__global__ void __launch_bounds__(128)
k_star_mask(const int4* __restrict__ rects, const int* __restrict__ rect_off, int nframes, int total,
            u32* __restrict__ mask, Dims d)
{
    const int ri = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ri >= total) return;
    int flo = 0, fhi = nframes;                     // largest f with rect_off[f] <= ri
    while (fhi - flo > 1) { int mid = (flo + fhi) >> 1; if (rect_off[mid] <= ri) flo = mid; else fhi = mid; }
    const int4 r = rects[ri];
    int r0 = max(r.x, 0), r1 = min(r.y, d.H), c0 = max(r.z, 0), c1 = min(r.w, d.W);
    if (r0 >= r1 || c0 >= c1) return;
    int w0 = c0 >> 5, w1 = (c1 - 1) >> 5;
    int nw = w1 - w0 + 1;
    int totalw = (r1 - r0) * nw;
    u32* m = mask + (size_t)flo * d.NW;
    for (int i = lane_id(); i < totalw; i += 32) {
        int row = r0 + i / nw, w = w0 + i % nw;
        int lo = max(c0 - (w << 5), 0), hi = min(c1 - 1 - (w << 5), 31);
        atomicOr(&m[(size_t)(d.H - 1 - row) * d.WW + w], bit_range(lo, hi));
    }
}

// cv2.convertScaleAbs for alpha=1, beta=0: saturate_u8(rint(|v|)); NaN, inf and |v| >= 2^31 give 0
// (x86 cvtps2dq "integer indefinite" is negative and saturates to 0).
__device__ __forceinline__ u32 csa(float v)
{
    float a = fabsf(v);
    if (!(a < 2147483648.0f)) return 0u;
    int r = __float2int_rn(a);
    return (u32)min(r, 255);
}

__device__ __forceinline__ float bswapf(float v)
{
    return __uint_as_float(__byte_perm(__float_as_uint(v), 0, 0x0123));
}

// ---- generic variant: also emits the clipped float image (standalone pass functions with write-back).
// mode 0: whole-frame pipeline: mask + flip + bright clip -> gray0 ; + dim threshold/offset -> gray1
// mode 1: standalone bright on an already flipped frame (no mask, no flip) -> gray0
// mode 2: standalone dim on an already flipped frame (no bright clip)     -> gray1
// Each thread converts 4 consecutive pixels (W % 4 == 0).  hist: [2][n][256].
// clipped (optional): float image after the in-place clip of the selected standalone pass.
__global__ void __launch_bounds__(256)
k_prep_generic(const float* __restrict__ in, const u32* __restrict__ mask, u8* __restrict__ gray0,
       u8* __restrict__ gray1, u32* __restrict__ hist0, u32* __restrict__ hist1, float* __restrict__ clipped,
       Dims d, int mode, int bigendian, float minFlux, float addFlux)
{
    __shared__ u32 sh[2][256];
    int f = blockIdx.y;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();

    const float* src = in + (size_t)f * d.N;
    int quads = d.N >> 2;
    int wq = d.W >> 2;
    // bins 0..2 hold almost every pixel of a sky-subtracted frame (dim: |v + addFlux| rounds to 1 or 2);
    // they are counted in registers and merged once per warp, the rest goes to shared-memory atomics
    u32 z0[3] = {0, 0, 0}, z1[3] = {0, 0, 0};
    const int stride = gridDim.x * blockDim.x;
    for (int q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < quads; q0 += 2 * stride) {
        // two independent 16-byte loads in flight per thread
        float4 vv[2];
        int qq[2] = {q0, q0 + stride};
        int yy[2], xx[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            int q = qq[u];
            if (q < quads) {
                yy[u] = q / wq; xx[u] = (q - yy[u] * wq) << 2;
                int sy = (mode == 0) ? (d.H - 1 - yy[u]) : yy[u];
                vv[u] = __ldcs(reinterpret_cast<const float4*>(src + (size_t)sy * d.W + xx[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            int q = qq[u];
            if (q >= quads) continue;
            int y = yy[u], x = xx[u];
            float4 v = vv[u];
            if (bigendian) { v.x = bswapf(v.x); v.y = bswapf(v.y); v.z = bswapf(v.z); v.w = bswapf(v.w); }
            float a[4] = {v.x, v.y, v.z, v.w};
            u32 mbits = 0;
            if (mode == 0) mbits = (mask[(size_t)f * d.NW + (size_t)y * d.WW + (x >> 5)] >> (x & 31)) & 0xf;
            u32 g0 = 0, g1 = 0;
            float cl[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float t = a[k];
                if (mbits & (1u << k)) t = 0.0f;
                if (mode != 2) { if (t < 0.0f) t = 0.0f; }          // bright clip (NaN stays NaN)
                float b = t;
                if (mode != 1) {
                    if (b < minFlux) b = 0.0f;
                    if (b > 0.0f) b = __fadd_rn(b, addFlux);
                }
                u32 c0 = csa(t), c1 = csa(b);
                g0 |= c0 << (8 * k);
                g1 |= c1 << (8 * k);
                if (mode != 2) {
                    if (c0 < 3) { z0[0] += (c0 == 0); z0[1] += (c0 == 1); z0[2] += (c0 == 2); }
                    else atomicAdd(&sh[0][c0], 1u);
                }
                if (mode != 1) {
                    if (c1 < 3) { z1[0] += (c1 == 0); z1[1] += (c1 == 1); z1[2] += (c1 == 2); }
                    else atomicAdd(&sh[1][c1], 1u);
                }
                cl[k] = (mode == 2) ? b : t;
            }
            size_t o = (size_t)f * d.N + (size_t)y * d.W + x;
            if (mode != 2) *reinterpret_cast<u32*>(gray0 + o) = g0;
            if (mode != 1) *reinterpret_cast<u32*>(gray1 + o) = g1;
            if (clipped) *reinterpret_cast<float4*>(clipped + o) = make_float4(cl[0], cl[1], cl[2], cl[3]);
        }
    }
    // low bins: warp-aggregate before touching shared memory
#pragma unroll
    for (int k = 0; k < 3; k++) {
        for (int o = 16; o; o >>= 1) {
            z0[k] += __shfl_xor_sync(FULLMASK, z0[k], o);
            z1[k] += __shfl_xor_sync(FULLMASK, z1[k], o);
        }
        if (lane_id() == 0) {
            if (z0[k]) atomicAdd(&sh[0][k], z0[k]);
            if (z1[k]) atomicAdd(&sh[1][k], z1[k]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        if (mode != 2 && sh[0][i]) atomicAdd(&hist0[(size_t)f * 256 + i], sh[0][i]);
        if (mode != 1 && sh[1][i]) atomicAdd(&hist1[(size_t)f * 256 + i], sh[1][i]);
    }
}


// ---- fast variant (no clipped output): the instruction-lean kernel of the batch pipeline -----------------
// saturate_u8(rint(|v|)) in three instructions: fabs (operand modifier), the >= 2^31 / NaN / inf -> 0
// select, and one cvt.rni.sat.u8.f32 (round-half-even + saturation in hardware; NaN converts to 0).
__device__ __forceinline__ u32 csa_fast(float v)
{
    float a = fabsf(v);
    a = (a < 2147483648.0f) ? a : 0.0f;
    u32 r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(a));
    return r;
}

// rint to u8 for 0 <= a < 256: a + 2^23 leaves round-half-even(a) in the low mantissa byte (FADD rounds to
// nearest even) - one full-rate FP32 instruction instead of a quarter-rate F2I.  Anything that needs
// saturation or special handling (a >= 255.5, |v| >= 2^31, inf - which cv2 maps to 0, not 255) shows up as a
// non-zero bit in 0x00ffff00 of the sum's bit pattern and is redone exactly by the caller.
__device__ __forceinline__ u32 rint_bits(float a) { return __float_as_uint(__fadd_rn(a, 8388608.0f)); }
#define RINT_BITS_OVERFLOW 0x00ffff00u

__device__ __forceinline__ u32 pack_low_bytes(u32 u0, u32 u1, u32 u2, u32 u3)
{
    return __byte_perm(__byte_perm(u0, u1, 0x0040), __byte_perm(u2, u3, 0x0040), 0x5410);
}

// histogram of four packed u8 values.  Sky-subtracted frames are almost entirely 0 (not counted at all: the zero
// bin is the remainder) and 1 (one popc); anything else goes to shared-memory atomics.
__device__ __forceinline__ void hist4_fast(u32 g, u32& n1, u32& nbig, u32* sh)
{
    n1 += __popc(g);                               // exact when every byte is 0 or 1 ...
    if (g & 0xfefefefeu) {                         // ... otherwise (rare) take it back and count byte by byte
        n1 -= __popc(g);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            u32 c = (g >> (8 * k)) & 0xffu;
            if (c == 1u) n1++;
            else if (c >= 2u) { atomicAdd(&sh[c], 1u); nbig++; }
        }
    }
}

// one float4: mask, clips, both conversions.  t[] are the un-flipped-source pixels of 4 consecutive columns.
template <int MODE>
__device__ __forceinline__ void prep4(float4 v, u32 mb, int bigendian, float minFlux, float addFlux, u32& g0, u32& g1)
{
    if (bigendian) { v.x = bswapf(v.x); v.y = bswapf(v.y); v.z = bswapf(v.z); v.w = bswapf(v.w); }
    if (MODE == 0 && mb) {
        if (mb & 1u) v.x = 0.0f;
        if (mb & 2u) v.y = 0.0f;
        if (mb & 4u) v.z = 0.0f;
        if (mb & 8u) v.w = 0.0f;
    }
    float t[4] = {v.x, v.y, v.z, v.w};
    u32 a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float tt = t[k];
        if (MODE != 2) {
            tt = fmaxf(tt, 0.0f);              // v < 0 -> 0 ; NaN -> 0 (the reference keeps NaN, which converts to 0 too)
            a[k] = rint_bits(tt);
        }
        if (MODE != 1) {
            float bb = (tt < minFlux) ? 0.0f : tt;
            bb = (bb > 0.0f) ? __fadd_rn(bb, addFlux) : bb;
            if (MODE == 2) bb = fabsf(bb);     // standalone dim: negative input -> |v| ; NaN fails the overflow test below
            b[k] = rint_bits(bb);
        }
    }
    g0 = (MODE != 2) ? pack_low_bytes(a[0], a[1], a[2], a[3]) : 0u;
    g1 = (MODE != 1) ? pack_low_bytes(b[0], b[1], b[2], b[3]) : 0u;
    // exact redo of the rare pixels that need saturation / special values
    if ((a[0] | a[1] | a[2] | a[3] | b[0] | b[1] | b[2] | b[3]) & RINT_BITS_OVERFLOW) {
        g0 = 0; g1 = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float tt = t[k];
            if (MODE != 2) { tt = fmaxf(tt, 0.0f); g0 |= csa_fast(tt) << (8 * k); }
            if (MODE != 1) {
                float bb = (tt < minFlux) ? 0.0f : tt;
                bb = (bb > 0.0f) ? __fadd_rn(bb, addFlux) : bb;
                g1 |= csa_fast(bb) << (8 * k);
            }
        }
    }
}

// MODE 0: pipeline (mask + flip, both passes) ; 1: bright only ; 2: dim only (no bright clip).
// One CTA walks whole rows (no integer division); a thread converts 2 x 4 px per iteration, both 16-byte
// loads issued before either is processed (bytes in flight per SM are what bounds a streaming kernel).
template <int MODE>
__global__ void __launch_bounds__(256)
k_prep(const float* __restrict__ in, const u32* __restrict__ mask, u8* __restrict__ gray0, u8* __restrict__ gray1,
       u32* __restrict__ hist0, u32* __restrict__ hist1, Dims d, int bigendian, float minFlux, float addFlux)
{
    // grid = (rows-CTAs per frame, frames), sized by the host to one wave of resident CTAs.  A CTA walks whole
    // rows (no integer division), two rows per iteration: a thread issues four 16-byte loads before it converts
    // anything - bytes in flight per SM, not instruction issue, bound this streaming kernel.
    __shared__ u32 sh[2][256];
    const int f = blockIdx.y;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&sh[0][0])[i] = 0;
    __syncthreads();
    const float* src = in + (size_t)f * d.N;
    const int wq = d.W >> 2;
    u32 a1 = 0, abig = 0, b1 = 0, bbig = 0, nwords = 0;
    for (int ya = blockIdx.x; ya < d.H; ya += 2 * gridDim.x) {
        const int yb = ya + gridDim.x;
        const bool rowb = yb < d.H;
        const int ybc = rowb ? yb : ya;
        const int sya = (MODE == 0) ? (d.H - 1 - ya) : ya, syb = (MODE == 0) ? (d.H - 1 - ybc) : ybc;
        const float4* srowa = reinterpret_cast<const float4*>(src + (size_t)sya * d.W);
        const float4* srowb = reinterpret_cast<const float4*>(src + (size_t)syb * d.W);
        const u32* mrowa = mask + (size_t)f * d.NW + (size_t)ya * d.WW;
        const u32* mrowb = mask + (size_t)f * d.NW + (size_t)ybc * d.WW;
        u32* o0a = reinterpret_cast<u32*>(gray0 + (size_t)f * d.N + (size_t)ya * d.W);
        u32* o1a = reinterpret_cast<u32*>(gray1 + (size_t)f * d.N + (size_t)ya * d.W);
        u32* o0b = reinterpret_cast<u32*>(gray0 + (size_t)f * d.N + (size_t)ybc * d.W);
        u32* o1b = reinterpret_cast<u32*>(gray1 + (size_t)f * d.N + (size_t)ybc * d.W);
        for (int xq = threadIdx.x; xq < wq; xq += 2 * blockDim.x) {
            const int xq2 = xq + blockDim.x;
            const bool two = xq2 < wq;
            const int xq2c = two ? xq2 : xq;
            float4 v[4];
            v[0] = __ldcs(srowa + xq); v[1] = __ldcs(srowa + xq2c);
            v[2] = __ldcs(srowb + xq); v[3] = __ldcs(srowb + xq2c);
            u32 m[4] = {0, 0, 0, 0};
            if (MODE == 0) {
                m[0] = (__ldg(mrowa + (xq >> 3)) >> ((xq & 7) << 2)) & 0xfu;
                m[1] = (__ldg(mrowa + (xq2c >> 3)) >> ((xq2c & 7) << 2)) & 0xfu;
                m[2] = (__ldg(mrowb + (xq >> 3)) >> ((xq & 7) << 2)) & 0xfu;
                m[3] = (__ldg(mrowb + (xq2c >> 3)) >> ((xq2c & 7) << 2)) & 0xfu;
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const bool on = (k == 0) || (k == 1 && two) || (k == 2 && rowb) || (k == 3 && two && rowb);
                if (!on) continue;
                u32 g0, g1;
                prep4<MODE>(v[k], m[k], bigendian, minFlux, addFlux, g0, g1);
                const int xo = (k & 1) ? xq2 : xq;
                if (MODE != 2) { ((k & 2) ? o0b : o0a)[xo] = g0; hist4_fast(g0, a1, abig, sh[0]); }
                if (MODE != 1) { ((k & 2) ? o1b : o1a)[xo] = g1; hist4_fast(g1, b1, bbig, sh[1]); }
                nwords++;
            }
        }
    }
    // bins 1 and 0 from the register counters (0 = everything that was not counted elsewhere)
    u32 acc[4] = {a1, 4u * nwords - a1 - abig, b1, 4u * nwords - b1 - bbig};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u32 vsum = acc[k];
        for (int o = 16; o; o >>= 1) vsum += __shfl_xor_sync(FULLMASK, vsum, o);
        if (lane_id() == 0 && vsum) atomicAdd(&sh[k >> 1][(k & 1) ? 0 : 1], vsum);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        if (MODE != 2 && sh[0][i]) atomicAdd(&hist0[(size_t)f * 256 + i], sh[0][i]);
        if (MODE != 1 && sh[1][i]) atomicAdd(&hist1[(size_t)f * 256 + i], sh[1][i]);
    }
}

// cv2.equalizeHist's LUT from the 256-bin histogram; one 256-thread block per (frame, pass).
// grid = (n, 2): blockIdx.y selects hist/lut of pass 0/1 (pointers are [2][n][256] slabs).
__global__ void __launch_bounds__(256)
k_lut(const u32* __restrict__ hist, u8* __restrict__ lut, const FrameCtl* __restrict__ ctl, int n, int total,
      int pass_lo)
{
    int f = blockIdx.x, p = pass_lo + blockIdx.y;
    if (!ctl[f].active[p]) return;
    const u32* h = hist + ((size_t)p * n + f) * 256;
    u8* l = lut + ((size_t)p * n + f) * 256;
    __shared__ int s[256];
    __shared__ int i0s;
    int t = threadIdx.x;
    int hv = (int)h[t];
    if (t == 0) i0s = 256;
    __syncthreads();
    if (hv > 0) atomicMin(&i0s, t);
    __syncthreads();
    int i0 = i0s;
    if (i0 == 256 || (int)h[i0] == total) { l[t] = (u8)t; return; }   // constant image: dst = src
    s[t] = (t > i0) ? hv : 0;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        int v = (t >= o) ? s[t - o] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    float scale = 255.0f / (float)(total - (int)h[i0]);
    int val = 0;
    if (t > i0) {
        val = __float2int_rn(__fmul_rn((float)s[t], scale));
        val = min(max(val, 0), 255);
    }
    l[t] = (u8)val;
}

// expand a bit mask to a 0/255 uint8 plane (stage taps)
__global__ void k_expand_mask(const u32* __restrict__ mask, u8* __restrict__ out, Dims d)
{
    int f = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x) {
        int y = i / d.W, x = i - y * d.W;
        u32 w = mask[(size_t)f * d.NW + (size_t)y * d.WW + (x >> 5)];
        out[(size_t)f * d.N + i] = ((w >> (x & 31)) & 1u) ? 255 : 0;
    }
}

// LUT applied to a plane (EQU tap: equalizeHist output before morphology)
__global__ void k_apply_lut(const u8* __restrict__ in, const u8* __restrict__ lut, u8* __restrict__ out, Dims d)
{
    int f = blockIdx.y;
    const u8* l = lut + (size_t)f * 256;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d.N; i += gridDim.x * blockDim.x)
        out[(size_t)f * d.N + i] = l[in[(size_t)f * d.N + i]];
}
