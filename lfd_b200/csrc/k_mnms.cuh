// k_mnms.cuh - morphology -> Sobel -> non-maximum suppression in ONE marching kernel, fed by TMA row tiles.
//
// Replaces, for the production path, the k_morph_march -> (morph plane in HBM) -> k_nms_march pair:
//   cv2.equalizeHist LUT + cv2.erode + cv2.dilate      /root/reference/lfd/detecttrails/processfield.py:347,354 / :457,464,471
//   cv2.Canny(img, 0, 255) front half (Sobel, L1, NMS) /root/reference/lfd/detecttrails/processfield.py:236
// The arithmetic is the one of those two kernels (k_morph.cuh / k_canny.cuh: same u16x2 SIMD window code, same
// OpenCV integer direction test); what changes is the data movement and the scheduling:
//
// * The uint8 `gray` plane is the only pixel input.  A warp owns a 256-px strip (224 useful + 16 px halo per side;
//   the halo covers erode + dilate reach (<= 8 px) + Sobel (1) + NMS (1)) and marches down FZ_R output rows in
//   blocks of U = lcm(erode rows, dilate rows) input rows (the period of the vertical register rings).  Every block
//   arrives as ONE U x 256-byte box of a 3-D tensor map over gray[frame][y][x] (cp.async.bulk.tensor -> UTMALDG),
//   NSTG boxes in flight per warp, each completing on the warp's own mbarrier: no register prefetch, no
//   __syncthreads in the loop, and rows / columns outside the frame arrive zero-filled by the TMA unit.
// * The dilated + LUT-ed rows never go to HBM: the producer half of a block leaves its rows in a per-warp
//   shared-memory queue (each lane re-reads only what it wrote), the consumer half pushes them through the
//   Sobel / NMS row machine of k_nms_march.  The `morph` plane is written only for the stage tap (TAP).
// * All-zero rows are the common case on sky-subtracted frames (bright: everything but stars; dim: everything the
//   3x3 erosion removes).  A warp-uniform vote per row skips the dilation and the LUT of rows whose vertical window
//   is zero, and a block whose rows are all zero while the Sobel / NMS machine is at rest (three zero rows pushed,
//   no magnitude left in its rings) skips the consumer half altogether: its output rows are stored as zeros.
// * Four independent warps per CTA (the old kernels had single-warp CTAs and hit the 32-CTA/SM limit at 25 %
//   achieved occupancy).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "k_canny.cuh"
#include "k_morph.cuh"
#include "k_prep.cuh"      // mbarrier helpers

#define FZ_WARPS 4          // warps (= strip units) per CTA
#ifndef FZ_R
#define FZ_R 64             // output rows per unit
#endif
#ifndef FZ_NSTG_SMALL
#define FZ_NSTG_SMALL 4     // boxes in flight per warp when a box is <= 4 rows
#endif
#ifndef FZ_MINB_SMALL
#define FZ_MINB_SMALL 1     // __launch_bounds__ min CTAs per SM: blocks of <= 4 rows / <= 9 rows / larger
#endif
#ifndef FZ_MINB_MID
#define FZ_MINB_MID 1
#endif
#ifndef FZ_MINB_BIG
#define FZ_MINB_BIG 1
#endif
#define FZ_BOXW 256         // box width in pixels = strip width

__device__ __forceinline__ void tma_load_3d(u32 dst, const CUtensorMap* tm, u32 bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(dst), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

// compile-time geometry of one instantiation (also used by the host for the tensor-map box and the smem size)
template <int EH, int DH>
struct FzGeom {
    static constexpr int EHR = EH > 0 ? EH : 1;
    static constexpr int U = EHR * DH / morph_gcd(EHR, DH);        // rows per block = rows per TMA box
    static constexpr int NSTG = U <= 4 ? FZ_NSTG_SMALL : 2;         // boxes in flight per warp (sized so that 4-6 CTAs fit an SM)
    static constexpr int MINB = U <= 4 ? FZ_MINB_SMALL : (U <= 9 ? FZ_MINB_MID : FZ_MINB_BIG);
    static constexpr int RING_B = NSTG * U * FZ_BOXW;
    static constexpr int MQ_B = U * 256;                            // produced morph rows (8 px per lane, packed bytes)
    static constexpr int NMS_B = 3 * 3 * 256 * 2;                   // magnitude / dx / dy rows, 3 each, 16 bit
    static constexpr int WARP_B = (RING_B + MQ_B + NMS_B + NSTG * 8 + 127) / 128 * 128;
    static constexpr int SMEM_B = FZ_WARPS * WARP_B + 256 + 128;    // + LUT + alignment slack
};

// grid = (ceil(nunits / FZ_WARPS), frames of this launch); pointers are already offset to the launch's first frame,
// the tensor map covers the whole batch, so its frame coordinate is f0 + blockIdx.y.
template <int EH, int EW, int DH, int DW, bool TAP>
__global__ void __launch_bounds__(FZ_WARPS * 32, FzGeom<EH, DH>::MINB)
k_morph_nms(const __grid_constant__ CUtensorMap tm, int f0, const u8* __restrict__ lut, u8* __restrict__ morph_tap,
            u32* __restrict__ nz, u8* __restrict__ eroded_tap, u32* __restrict__ cand, u32* __restrict__ strong,
            u8* __restrict__ nms_tap, const FrameCtl* __restrict__ ctl, int pass, Dims d, int nstrips, int nunits,
            int low, int high)
{
    typedef FzGeom<EH, DH> G;
    constexpr bool HAS_E = EH > 0;
    constexpr int EHR = G::EHR, U = G::U, NSTG = G::NSTG;
    constexpr int E_T = HAS_E ? -(EH / 2) : 0, E_B = HAS_E ? EH - 1 - EH / 2 : 0;
    constexpr int E_L = HAS_E ? -(EW / 2) : 0, E_R = HAS_E ? EW - 1 - EW / 2 : 0;
    constexpr int D_T = -(DH / 2), D_B = DH - 1 - DH / 2, D_L = -(DW / 2), D_R = DW - 1 - DW / 2;
    static_assert(-E_L - D_L + 2 <= 16 && E_R + D_R + 2 <= 16, "erode + dilate + Sobel + NMS reach exceeds the 16-px strip halo");
    static_assert(U <= 32, "queue flags are one 32-bit word");

    const int f = blockIdx.y;
    if (!ctl[f].active[pass]) return;
    extern __shared__ unsigned char fz_dsm[];
    unsigned char* const dsm = fz_dsm + ((128u - (smem_u32(fz_dsm) & 127u)) & 127u);     // TMA destinations: 128-byte aligned
    u8* const slut = dsm + FZ_WARPS * G::WARP_B;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* const wb = dsm + wid * G::WARP_B;
    unsigned char* const ring = wb;                                                  // [NSTG][U][256]
    uint2* const mq = reinterpret_cast<uint2*>(wb + G::RING_B);                      // [U][32]
    u16* const sM = reinterpret_cast<u16*>(wb + G::RING_B + G::MQ_B);               // [3][256]
    short* const sDX = reinterpret_cast<short*>(sM + 3 * 256);                       // [3][256]
    short* const sDY = sDX + 3 * 256;                                                // [3][256]
    const u32 bar0 = smem_u32(wb + G::RING_B + G::MQ_B + G::NMS_B);                  // [NSTG] mbarriers
    for (int i = threadIdx.x; i < 256; i += blockDim.x) slut[i] = lut[(size_t)f * 256 + i];
    if (lane == 0)
        for (int s = 0; s < NSTG; s++) mbar_init(bar0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int unit = blockIdx.x * FZ_WARPS + wid;
    if (unit >= nunits) return;

    const int chunk = unit / nstrips, s = unit - chunk * nstrips;
    const int Ww = d.W >> 2;
    const int wx = s * MARCH_UW - MARCH_HW + 2 * lane;       // first of this lane's two 4-px words
    const bool col_in = wx >= 0 && wx < Ww;
    const bool lane_out = lane >= 2 && lane < 30 && col_in;
    const int y0 = chunk * FZ_R, y1 = min(y0 + FZ_R, d.H);
    const int m_first = max(y0 - 2, 0), m_last = min(y1 + 1, d.H - 1);     // morph rows the Sobel / NMS stage needs
    const int in_first = m_first + D_T + E_T, in_last = m_last + D_B + E_B; // gray rows those need
    const int nblocks = (in_last - in_first + U) / U;
    const int x0 = (s * MARCH_UW - MARCH_HW) * 4;            // strip origin in pixels (may be -16: zero-filled)
    const int fabs_ = f0 + f;
    const u32 ring0 = smem_u32(ring);

    int pb = 0, ps = 0;                                      // producer: next block to request, its ring slot
    auto issue = [&]() {
        if (pb < nblocks) {
            if (lane == 0) {
                mbar_expect_tx(bar0 + 8 * ps, U * FZ_BOXW);
                tma_load_3d(ring0 + ps * (U * FZ_BOXW), &tm, bar0 + 8 * ps, x0, in_first + pb * U, fabs_);
            }
            pb++;
            ps = (ps + 1 == NSTG) ? 0 : ps + 1;
        }
    };
#pragma unroll
    for (int st = 0; st < NSTG; st++) issue();

    // ---- consumer state: the Sobel / NMS row machine of k_nms_march with run-time ring rotation ----
    const int mw0 = s * (MARCH_UW / 8);                      // first mask word of the strip
    const int nwords = min(MARCH_UW / 8, d.WW - mw0);        // mask words this strip owns (<= 7)
    u32* const candf = cand + (size_t)f * d.NW + mw0 + lane;
    u32* const strongf = strong + (size_t)f * d.NW + mw0 + lane;
    const bool out_lane = lane < nwords;
    u32 HD0[4], HD1[4], HD2[4], H30[4], H31[4], H32[4];      // rows yi-2, yi-1, yi
#pragma unroll
    for (int c = 0; c < 4; c++) { HD0[c] = HD1[c] = HD2[c] = 0; H30[c] = H31[c] = H32[c] = 0; }
    // The machine starts "at rest": rings zero, as if three all-zero rows had been pushed.  (For the first real rows
    // that is indistinguishable from the k_nms_march start-up: the first two centre rows it produces are never used.)
    bool rz0 = true, rz1 = true, rz2 = true;                 // pushed rows yi-2, yi-1, yi were all-zero across the strip
    int slA = 0, slB = 1, slC = 2;                           // smem slots: A = mag(yi-2), B = mag(yi-1) (written now), C = mag(yi-3)
    bool anyA = false, anyB = false, anyC = false;           // slot content has a non-zero magnitude (some lane)
    bool lnA = false, lnB = false, lnC = false;              // ... in this lane's 8 pixels
    for (int i = lane; i < 3 * 256 / 2; i += 32) reinterpret_cast<u32*>(sM)[i] = 0u;
    __syncwarp();
    const int TG22 = 13573;

    auto zero_row_out = [&](int yn) {
        if (out_lane) { const int o = yn * d.WW; candf[o] = 0u; strongf[o] = 0u; }
        if (TAP)
            for (int x = lane; x < 32 * nwords; x += 32)
                if (32 * mw0 + x < d.W) nms_tap[(size_t)f * d.N + (size_t)yn * d.W + 32 * mw0 + x] = 0;
    };

    // push image row `yi` (q = its LUT-ed pixels as u16x2 pairs, BORDER_REPLICATE already applied to the columns):
    // gradient of the centre row yi-1, suppression of row yi-2
    auto push = [&](const u32 (&q)[4], bool rowzero, int yi) {
        // rotate the rings: (yi-2, yi-1, yi) <- (yi-1, yi, new)
#pragma unroll
        for (int c = 0; c < 4; c++) { HD0[c] = HD1[c]; HD1[c] = HD2[c]; H30[c] = H31[c]; H31[c] = H32[c]; }
        rz0 = rz1; rz1 = rz2; rz2 = rowzero;
        { const int t = slA; slA = slB; slB = slC; slC = t; }
        { const bool t = anyA; anyA = anyB; anyB = anyC; anyC = t; }
        { const bool t = lnA; lnA = lnB; lnB = lnC; lnC = t; }
        const int yn = yi - 2;
        const bool yn_in = yn >= y0 && yn < y1;              // warp-uniform
        if (rz0 && rz1 && rz2) {
            // three all-zero rows: zero gradient on the centre row, nothing to suppress
#pragma unroll
            for (int c = 0; c < 4; c++) { HD2[c] = 0; H32[c] = 0; }
            if (anyB) reinterpret_cast<uint4*>(sM + slB * 256)[lane] = make_uint4(0u, 0u, 0u, 0u);
            lnB = false; anyB = false;
            __syncwarp();
            if (!yn_in) return;
            if (!anyA) { zero_row_out(yn); return; }
        } else {
            const u32 eL = __shfl_up_sync(FULLMASK, q[3], 1), eR = __shfl_down_sync(FULLMASK, q[0], 1);
            u32 T[5];                                        // T[j] = (px 2j-1, px 2j) of the lane's 8 pixels
            T[0] = __byte_perm(eL, q[0], 0x5432);
            T[1] = __byte_perm(q[0], q[1], 0x5432);
            T[2] = __byte_perm(q[1], q[2], 0x5432);
            T[3] = __byte_perm(q[2], q[3], 0x5432);
            T[4] = __byte_perm(q[3], eR, 0x5432);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                HD2[c] = vsub2(T[c + 1], T[c]);                                     // p(x+1) - p(x-1)
                H32[c] = __vadd2(__vadd2(T[c], T[c + 1]), __vadd2(q[c], q[c]));     // p(x-1) + 2p(x) + p(x+1)
            }
            const int yc = yi - 1;
            const bool cin = yc >= 0 && yc < d.H && col_in;
            u32 mg[4], dxv[4], dyv[4];
#pragma unroll
            for (int c = 0; c < 4; c++) {
                dxv[c] = __vadd2(__vadd2(HD0[c], HD2[c]), __vadd2(HD1[c], HD1[c]));
                dyv[c] = vsub2(H32[c], H30[c]);
                mg[c] = cin ? __vadd2(vabs2s(dxv[c]), vabs2s(dyv[c])) : 0u;
            }
            lnB = (mg[0] | mg[1] | mg[2] | mg[3]) != 0u;
            anyB = __any_sync(FULLMASK, lnB);
            reinterpret_cast<uint4*>(sM + slB * 256)[lane] = make_uint4(mg[0], mg[1], mg[2], mg[3]);
            if (lnB) {
                reinterpret_cast<uint4*>(sDX + slB * 256)[lane] = make_uint4(dxv[0], dxv[1], dxv[2], dxv[3]);
                reinterpret_cast<uint4*>(sDY + slB * 256)[lane] = make_uint4(dyv[0], dyv[1], dyv[2], dyv[3]);
            }
            __syncwarp();
            if (!yn_in) return;
        }
        // non-maximum suppression of row yn (magnitude rows yn-1, yn, yn+1 in slots C, A, B)
        const u32 bal = __ballot_sync(FULLMASK, lnA);
        u32 myc = 0, mys = 0;                                // lane g keeps mask word g
        const u16* Mu = sM + slC * 256; const u16* Mc = sM + slA * 256; const u16* Md = sM + slB * 256;
        const short* DXc = sDX + slA * 256; const short* DYc = sDY + slA * 256;
        u32 gm = bal >> 2;                                   // bit 4*gi: 32-px group gi (lanes 2+4gi .. 5+4gi) has a gradient
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
                    const int dx = DXc[x], dy = DYc[x];
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
        if (out_lane) { const int o = yn * d.WW; candf[o] = myc; strongf[o] = mys; }
        __syncwarp();
    };

    // ---- producer state: the vertical register rings of k_morph_march ----
    uint2* const mo = TAP ? reinterpret_cast<uint2*>(morph_tap + (size_t)f * d.N) : nullptr;
    uint2* const et = (TAP && eroded_tap) ? reinterpret_cast<uint2*>(eroded_tap + (size_t)f * d.N) : nullptr;
    const bool lut0z = slut[0] == 0;                         // gray 0 -> equalised 0 (always, when 0 occurs in the frame)
    const int r = (lane - 2) & 31;                           // mask word of this lane's group: lanes 2..5 -> word 7s, ...
    const int mw = mw0 + (r >> 2);
    const int q8 = (r & 3) * 8;
    const int src1 = ((r ^ 1) + 2) & 31, src2 = ((r ^ 2) + 2) & 31;
    const bool nz_lane = (r & 3) == 0 && lane >= 2 && lane < 30 && mw < d.WW;
    u32* const nzp = nz + (size_t)f * d.NW + mw;
    u32 eR[EHR][4], dR[DH][4];
#pragma unroll
    for (int k = 0; k < EHR; k++)
#pragma unroll
        for (int c = 0; c < 4; c++) eR[k][c] = 0;
#pragma unroll
    for (int k = 0; k < DH; k++)
#pragma unroll
        for (int c = 0; c < 4; c++) dR[k][c] = 0;
    u32 ezbits = (1u << EHR) - 1u;                           // ring rows of the erosion that are all-zero across the strip
    u32 dnz = 0;                                             // ring rows of the dilation that are NOT all-zero

    int cs = 0; u32 cpar = 0;                                // consumer: ring slot of the current block, its phase parity
    for (int b = 0; b < nblocks; b++) {
        const int yb = in_first + b * U;
        mbar_wait(bar0 + 8 * cs, cpar);
        const unsigned char* const rowp = ring + cs * (U * FZ_BOXW) + lane * 8;
        int nq = 0, q_first = 0;                             // rows queued by this block: yo = q_first .. q_first + nq - 1
        u32 qzero = 0;                                       // bit j: queued row j is all-zero (not stored in the queue)
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int y = yb + u;
            if (y > in_last) break;                          // warp-uniform (last block only)
            const bool yin = (unsigned)y < (unsigned)d.H;
            uint2 v = *reinterpret_cast<const uint2*>(rowp + u * FZ_BOXW);      // pixels outside the frame arrive as zeros
            const bool in_zero = !__any_sync(FULLMASK, (v.x | v.y) != 0u);
            u32 e[4];
            bool e_zero;                                     // warp-uniform: the row entering the dilation is all-zero
            int ye = y;
            if (HAS_E) {
                ye = y - E_B;
                if (in_zero && yin) {                        // a frame row of zeros: every minimum that sees it is zero
                    if (!((ezbits >> (u % EHR)) & 1u)) {
#pragma unroll
                        for (int c = 0; c < 4; c++) eR[u % EHR][c] = 0;
                        ezbits |= 1u << (u % EHR);
                    }
                } else {
                    // what cv2 ignores (pixels outside the frame) enters as the identity of the minimum
                    if (!(yin && col_in)) v = make_uint2(0xffffffffu, 0xffffffffu);
                    u32 p[4], hm[4];
                    p[0] = __byte_perm(v.x, 0, 0x4140); p[1] = __byte_perm(v.x, 0, 0x4342);
                    p[2] = __byte_perm(v.y, 0, 0x4140); p[3] = __byte_perm(v.y, 0, 0x4342);
                    hwin16<false, E_L, E_R>(p, hm);
#pragma unroll
                    for (int c = 0; c < 4; c++) eR[u % EHR][c] = hm[c];
                    ezbits &= ~(1u << (u % EHR));
                }
                if (ezbits != 0u) {                          // a zero row in the window: the minimum is zero
#pragma unroll
                    for (int c = 0; c < 4; c++) e[c] = 0u;
                    e_zero = true;
                } else {
                    const bool ein = col_in && ye >= 0 && ye < d.H;
#pragma unroll
                    for (int c = 0; c < 4; c++) e[c] = ein ? vreduce16<false, EHR>(eR, c) : 0u;
                    e_zero = !__any_sync(FULLMASK, (e[0] | e[1] | e[2] | e[3]) != 0u);
                }
                if (TAP && et && ye >= y0 && ye < y1) {
                    u32 w0, w1;
                    lut_pack(e, slut, w0, w1);
                    if (lane_out) et[((long long)ye * Ww + wx) >> 1] = make_uint2(w0, w1);
                }
            } else {
                e_zero = in_zero;
            }
            if (e_zero) {
                if ((dnz >> (u % DH)) & 1u) {
#pragma unroll
                    for (int c = 0; c < 4; c++) dR[u % DH][c] = 0;
                    dnz &= ~(1u << (u % DH));
                }
            } else {
                if (!HAS_E) {
                    e[0] = __byte_perm(v.x, 0, 0x4140); e[1] = __byte_perm(v.x, 0, 0x4342);
                    e[2] = __byte_perm(v.y, 0, 0x4140); e[3] = __byte_perm(v.y, 0, 0x4342);
                }
                u32 hd[4];
                hwin16<true, D_L, D_R>(e, hd);
#pragma unroll
                for (int c = 0; c < 4; c++) dR[u % DH][c] = hd[c];
                dnz |= 1u << (u % DH);
            }
            const int yo = ye - D_B;
            if (yo >= m_first && yo <= m_last) {             // warp-uniform
                if (nq == 0) q_first = yo;
                const bool own = yo >= y0 && yo < y1;        // rows of this unit's chunk (the others only feed the Sobel / NMS halo)
                if (dnz == 0u && lut0z) {                    // the whole vertical window is zero: nothing to compute or queue
                    qzero |= 1u << nq;
                    if (own) {
                        if (nz_lane) nzp[yo * d.WW] = 0u;
                        if (TAP && lane_out) mo[((long long)yo * Ww + wx) >> 1] = make_uint2(0u, 0u);
                    }
                } else {
                    u32 o[4], w0, w1;
#pragma unroll
                    for (int c = 0; c < 4; c++) o[c] = vreduce16<true, DH>(dR, c);
                    lut_pack(o, slut, w0, w1);
                    const bool o_zero = !__any_sync(FULLMASK, (w0 | w1) != 0u);
                    mq[nq * 32 + lane] = make_uint2(w0, w1);
                    qzero |= (o_zero ? 1u : 0u) << nq;
                    if (own) {
                        if (TAP && lane_out) mo[((long long)yo * Ww + wx) >> 1] = make_uint2(w0, w1);
                        u32 bits = 0;
                        if (!o_zero) {
                            bits = lane_out ? ((nzbits4(w0) | (nzbits4(w1) << 4)) << q8) : 0u;
                            bits |= __shfl_sync(FULLMASK, bits, src1);
                            bits |= __shfl_sync(FULLMASK, bits, src2);
                        }
                        if (nz_lane) nzp[yo * d.WW] = bits;
                    }
                }
                nq++;
            }
        }
        // every lane holds its part of the box's last row: refill the slot
        __syncwarp();
        issue();
        cs++; if (cs == NSTG) { cs = 0; cpar ^= 1u; }
        if (nq == 0) continue;

        // ---- consumer half ----
        const int q_last = q_first + nq - 1;
        const bool at_bottom = q_last == d.H - 1 && y1 == d.H;       // the frame's last row is in this block: it is pushed 3 times
        if (!TAP && qzero == (0xffffffffu >> (32 - nq)) && rz0 && rz1 && rz2 && !anyA && !anyB && !anyC) {
            // all rows zero and the machine at rest: it stays at rest, the rows it would have suppressed are zero
            const int lo = max(q_first - 2, y0), hi = min(q_last - 2 + (at_bottom ? 2 : 0), y1 - 1);
            if (out_lane)
                for (int yn = lo; yn <= hi; yn++) { const int o = yn * d.WW; candf[o] = 0u; strongf[o] = 0u; }
            continue;
        }
#pragma unroll 1
        for (int j = 0; j < nq; j++) {
            const int yo = q_first + j;
            const bool rowzero = (qzero >> j) & 1u;
            u32 q[4] = {0u, 0u, 0u, 0u};
            if (!rowzero) {
                const uint2 w = mq[j * 32 + lane];
                q[0] = __byte_perm(w.x, 0, 0x4140); q[1] = __byte_perm(w.x, 0, 0x4342);
                q[2] = __byte_perm(w.y, 0, 0x4140); q[3] = __byte_perm(w.y, 0, 0x4342);
                // BORDER_REPLICATE columns: a lane outside the frame takes the frame's edge pixel from its neighbour
                const u32 fromR = __shfl_down_sync(FULLMASK, q[0] & 0xffffu, 1);      // first pixel of the lane to the right
                const u32 fromL = __shfl_up_sync(FULLMASK, q[3] >> 16, 1);            // last pixel of the lane to the left
                if (!col_in) {
                    const u32 ev = ((wx < 0) ? fromR : fromL) * 0x00010001u;
                    q[0] = q[1] = q[2] = q[3] = ev;
                }
            }
            // BORDER_REPLICATE rows: the frame's first row also stands in for row -1, its last row for rows H and H+1
            int yi = yo, reps = 1;
            if (yo == 0 && y0 == 0) { yi = -1; reps = 2; }
            if (yo == d.H - 1 && y1 == d.H) reps += 2;
            for (int k = 0; k < reps; k++) push(q, rowzero, yi + k);
        }
    }
}
