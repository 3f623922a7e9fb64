// k_hough.cuh - cv2.HoughLines-compatible standard Hough transform on bit masks, plus the line-set check.
//
// Reference calls (paths under /root/reference/lfd/detecttrails/):
//   processfield.py:370-371, :488-489   cv2.HoughLines(equ | box_img, houghMethod, np.pi/180, 1)
//   processfield.py:36-150              check_theta
// cv2 semantics (restated and pinned in oracle/c/cvrestate.c::orc_hough_lines): every NON-ZERO pixel votes
// r = cvRound(x*tabCos[n] + y*tabSin[n]) + (numrho-1)/2 with float32 tables (sin/cos in double, divided by
// rho, rounded to float; built on the host exactly as OpenCV does), separate roundings for both products
// and the sum (no FMA), round-half-even.  Peaks: votes > threshold, 4-neighbour maximum with OpenCV's
// > / >= asymmetry, sorted by (votes desc, accumulator index asc).
//
// Design: the voting unit is a non-zero 32-pixel mask word, not a pixel.  For one angle the 32 pixels of
// a word span at most ceil(32*|cos|/rho)+1 rho bins, so a lane (= angle) usually issues ONE shared-memory
// atomic of weight popc(word) instead of 32.  Accumulators are privatised per CTA in shared memory as
// int32 [angles-per-CTA][numrho+2] and flushed once with global atomics.
#pragma once
#include "common.cuh"

// non-zero words of the equ mask (which=0) and box mask (which=1) -> (y<<16 | w, bits)
__global__ void __launch_bounds__(256)
k_hough_compact(const u32* __restrict__ nzmask, const u32* __restrict__ boxmask, uint2* __restrict__ segs,
                FrameCtl* __restrict__ ctl, int pass, Dims d, size_t seg_stride)
{
    int f = blockIdx.y, which = blockIdx.z;
    if (!ctl[f].active[pass] || !ctl[f].hough[pass]) return;
    const u32* m = (which ? boxmask : nzmask) + (size_t)f * d.NW;
    uint2* out = segs + ((size_t)f * 2 + which) * seg_stride;
    int lane = lane_id();
    int nnz = 0;
    for (int i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; i0 < d.NW; i0 += gridDim.x * blockDim.x) {
        int i = i0 + lane;
        u32 v = (i < d.NW) ? m[i] : 0u;
        u32 bal = __ballot_sync(FULLMASK, v != 0);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&ctl[f].nseg[which], __popc(bal));
            base = __shfl_sync(FULLMASK, base, 0);
            if (v) {
                int y = i / d.WW, w = i - y * d.WW;
                out[base + __popc(bal & ((1u << lane) - 1u))] = make_uint2(((u32)y << 16) | (u32)w, v);
                nnz += __popc(v);
            }
        }
    }
    for (int o = 16; o; o >>= 1) nnz += __shfl_xor_sync(FULLMASK, nnz, o);
    if (lane == 0 && nnz) atomicAdd(&ctl[f].nnz[which], nnz);
}

#define HOUGH_THREADS 256

#define HOUGH_MIN_SEGS 512

// smem accumulator row stride: odd, so that lanes (= angles) voting for similar rho hit different banks
__host__ __device__ inline int hough_rss(int RS) { return RS | 1; }

// r = cvRound(x * cos + y * sin) exactly as OpenCV's float expression (separate roundings, round-half-even).
// MAGIC: the two conversions run on the FP32 pipe instead of the quarter-rate XU pipe (I2F / F2I), bit-identically:
//   (float)x      = as_float(0x4B000000 + x) - 2^23          exact for 0 <= x < 2^23
//   rint(v) (RNE) = as_int(v + 1.5 * 2^23) - 0x4B400000      exact for |v| < 2^22 (the sum lands in [2^23, 2^24), where the
//                                                            float grid is the integers and FADD rounds half to even)
// The host selects MAGIC when the frame and rho keep both ranges (hough_magic_ok); otherwise the cvt instructions.
template <bool MAGIC>
__device__ __forceinline__ int hough_r(int x, float c, float ys)
{
    if (MAGIC) {
        const float xf = __fsub_rn(__int_as_float(0x4B000000 + x), 8388608.0f);
        const float v = __fadd_rn(__fmul_rn(xf, c), ys);
        return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
    }
    return __float2int_rn(__fadd_rn(__fmul_rn((float)x, c), ys));
}
__host__ inline bool hough_magic_ok(int H, int W, float rho) { return W < (1 << 23) && (double)(W + H) / rho < 4194300.0; }

// grid = (chunks, ngroups, 2*n); dynamic smem = apc * hough_rss(RS) * 4 bytes
template <bool MAGIC>
__global__ void __launch_bounds__(HOUGH_THREADS)
k_hough_vote(const uint2* __restrict__ segs, int* __restrict__ accum, const float* __restrict__ tabSin,
             const float* __restrict__ tabCos, const FrameCtl* __restrict__ ctl, int pass, HoughCfg hc,
             size_t seg_stride, size_t accum_stride, unsigned long long* __restrict__ counters)
{
    extern __shared__ int acc[];
    int f = blockIdx.z >> 1, which = blockIdx.z & 1;
    if (!ctl[f].active[pass] || !ctl[f].hough[pass]) return;
    int nseg = ctl[f].nseg[which];
    // lanes = angles; when apc < 32 the warp covers 32/apc segments at once
    const int per = (hc.apc >= 32) ? 1 : (32 / hc.apc);
    // Every CTA zeroes and flushes apc x RSS accumulator cells whatever it votes, so an (image, angle group) gets only
    // as many of the gridDim.x CTAs (= privatised copies of its accumulator rows) as pay off:
    //  * >= HOUGH_MIN_SEGS segments each (the box image has ~10x fewer segments than the morphology output);
    //  * >= 8 votes per accumulator cell and copy on average.  At fine rho the rows are long (RS = 7077 at rho = 1 on a
    //    2048 x 1489 frame) and a copy that receives one or two votes per cell is flushed with as many GLOBAL atomics as
    //    it absorbed shared-memory ones: 32 copies of a 1.27 M-cell accumulator were 40 M global atomics per image.
    //  A single copy owns its rows and is flushed with plain stores.
    const int RSS = hough_rss(hc.RS);
    const int nnz = ctl[f].nnz[which];
    int nchunks = min((int)gridDim.x, max(1, (nseg + HOUGH_MIN_SEGS - 1) / HOUGH_MIN_SEGS));
    //  * but at least enough copies to put ~2 CTAs on every SM when few images are in flight (the single-image entry
    //    point lfd_hough_lines; a batch has 2 x frames images in gridDim.z and never needs this)
    {
        const int by_seg = nchunks;
        const int by_votes = max(1, nnz / (8 * RSS));
        const int by_par = (296 + hc.ngroups * (int)gridDim.z - 1) / (hc.ngroups * (int)gridDim.z);
        nchunks = min(by_seg, max(by_votes, by_par));
    }
    if ((int)blockIdx.x >= nchunks || nseg == 0) return;
    int g = blockIdx.y;
    int a0 = g * hc.apc;
    int na = min(hc.apc, hc.numangle - a0);
    for (int i = threadIdx.x; i < hc.apc * RSS; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    const uint2* sg = segs + ((size_t)f * 2 + which) * seg_stride;
    int lane = lane_id();
    int sub = lane / hc.apc, a = lane - sub * hc.apc;
    bool live = (per == 1) ? (lane < na) : (sub < per && a < na);
    if (per == 1) { sub = 0; a = lane; }
    int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int nwarps = (nchunks * blockDim.x) >> 5;
    float c = 0.f, s = 0.f;
    if (live) { c = tabCos[a0 + a]; s = tabSin[a0 + a]; }
    const float inv_c = (c != 0.f) ? __frcp_rn(c) : 0.f;      // only used to predict bin boundaries (then verified)
    const int off = (hc.numrho - 1) / 2 + 1;
    int* row = acc + a * RSS + off;
    int natom = 0;                 // shared-memory atomics this thread issued (counters[13])
#define HVOTE(bin, w) do { atomicAdd(&row[bin], w); natom++; } while (0)
#define HOUGH_R(xx) hough_r<MAGIC>((xx), c, ys)
    // The segment list is read two iterations ahead: an iteration is a dependent chain behind one 8-byte load (L2
    // latency ~600 cycles against ~100 cycles of work), so without the prefetch the kernel waits on memory.
    // When a warp covers `per` > 1 words at once (fine rho: few angles per CTA), those words are taken from `per` distant
    // slices of the raster-ordered list (word si of slice `sub`), not from `per` neighbours: vertically adjacent words of
    // a blob vote for the same bins at the same moment, and same-address shared-memory atomics serialise.
    const int S = (nseg + per - 1) / per;              // words per slice
    const int step = nwarps;
    const uint2* const sgp = sg + sub * S;             // this lane's slice ...
    const int lim = live ? min(S, nseg - sub * S) : 0; // ... and how many words of it exist
    auto fetch = [&](int si) -> uint2 { return si < lim ? __ldg(sgp + si) : make_uint2(0u, 0u); };
    uint2 q0 = fetch(warp), q1 = fetch(warp + step);
    for (int si = warp; si < S; si += step) {
        const uint2 sgv = q0;
        q0 = q1;
        q1 = fetch(si + 2 * step);
        if (sgv.y == 0u) continue;                  // lane has no segment in this iteration (mask words are never zero)
        int y = sgv.x >> 16, x0 = (sgv.x & 0xffff) << 5;
        u32 m = sgv.y;
        const float yf = MAGIC ? __fsub_rn(__int_as_float(0x4B000000 + y), 8388608.0f) : (float)y;
        float ys = __fmul_rn(yf, s);
        const float x0f = MAGIC ? __fsub_rn(__int_as_float(0x4B000000 + x0), 8388608.0f) : (float)x0;
        int fb = __ffs(m) - 1, lb = 31 - __clz(m);
        int r1 = HOUGH_R(x0 + fb);
        int r2 = HOUGH_R(x0 + lb);
        if (r1 == r2) {
            HVOTE(r1, __popc(m));
        } else if (abs(r2 - r1) <= 2 && __popc(m) > 4) {
            // r is monotone in x: locate the (at most two) bin boundaries inside the word.  The crossing of
            // r + 0.5 (towards r2) is predicted from the real-valued line x = (r +- 0.5 - y sin) / cos, the exact
            // float expression is then probed at the prediction and its neighbour (two evaluations instead of a
            // five-step bisection); whatever interval is left - the prediction can be off by one where OpenCV's
            // two float roundings matter - is finished by bisection on the exact expression.
            const float half = (r2 > r1) ? 0.5f : -0.5f;
            auto last_with = [&](int lo, int hi, int rr) -> int {       // R(lo) == rr, R(hi) != rr  ->  largest x with R(x) == rr
                // (prediction only: the conversions may be the magic ones, exact for |rr| < 2^22 and 0 <= x0 < 2^23)
                const float rrf = MAGIC ? __fsub_rn(__int_as_float(0x4B400000 + rr), 12582912.0f) : (float)rr;
                const float xg = __fmul_rn(__fsub_rn(__fadd_rn(rrf, half), ys), inv_c) - x0f;
                // (a prediction only - any integer works, it is verified below - so the conversion may be the magic one too)
                int g = MAGIC ? (__float_as_int(__fadd_rn(xg, 12582912.0f)) - 0x4B400000) : __float2int_rd(xg);
                g = min(max(g, lo), hi - 1);
                if (HOUGH_R(x0 + g) == rr) lo = g; else hi = g;
                if (hi - lo > 1) {
                    g = (lo == g) ? g + 1 : g - 1;
                    if (HOUGH_R(x0 + g) == rr) lo = g; else hi = g;
                }
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (HOUGH_R(x0 + mid) == rr) lo = mid; else hi = mid;
                }
                return lo;
            };
            const int lo = last_with(fb, lb, r1), hi = lo + 1;
            int c1 = __popc(m & bit_range(fb, lo));
            HVOTE(r1, c1);
            // two adjacent bins and a monotone r: what follows the last pixel of r1 is r2, no evaluation needed
            int rm = (abs(r2 - r1) == 1) ? r2 : HOUGH_R(x0 + hi);
            if (rm == r2) {
                HVOTE(r2, __popc(m & bit_range(hi, lb)));
            } else {
                const int lo2 = last_with(hi, lb, rm), hi2 = lo2 + 1;
                int c2 = __popc(m & bit_range(hi, lo2));
                if (c2) HVOTE(rm, c2);
                HVOTE(r2, __popc(m & bit_range(hi2, lb)));
            }
        } else {
            // more than two bin boundaries inside the word (fine rho): the bins change at most pixels, so merging equal
            // neighbours saved few atomics and cost a divergent branch per pixel (37 executed instructions per vote at
            // rho = 1, ncu profiles/r02n_hough_c4_*); every pixel votes on its own, two independent evaluations per
            // iteration to overlap their dependent FP32 chains
            natom += __popc(m);
            atomicAdd(&row[r1], 1);
            m &= m - 1;
            while (m) {
                const int b0 = __ffs(m) - 1;
                m &= m - 1;
                const int ra = HOUGH_R(x0 + b0);
                if (m) {
                    const int b1 = __ffs(m) - 1;
                    m &= m - 1;
                    atomicAdd(&row[HOUGH_R(x0 + b1)], 1);
                }
                atomicAdd(&row[ra], 1);
            }
        }
    }
#undef HOUGH_R
#undef HVOTE
    for (int o = 16; o; o >>= 1) natom += __shfl_xor_sync(FULLMASK, natom, o);
    if (lane == 0 && natom && counters) atomicAdd(&counters[13], (unsigned long long)natom);
    __syncthreads();
    int* gacc = accum + ((size_t)f * 2 + which) * accum_stride;
    for (int aa = threadIdx.x >> 5; aa < na; aa += HOUGH_THREADS / 32) {        // one warp per angle row: no division
        const int* arow = acc + aa * RSS;
        int* grow = gacc + (size_t)(a0 + aa + 1) * hc.RS;
        if (nchunks == 1) {
            for (int col = lane; col < RSS; col += 32) { const int v = arow[col]; if (v) grow[col] = v; }      // sole owner of the row
        } else {
            for (int col = lane; col < RSS; col += 32) { const int v = arow[col]; if (v) atomicAdd(&grow[col], v); }
        }
    }
}

// peaks -> 64-bit sort keys ((INT_MAX - votes) << 32 | accumulator index): ascending key order is
// OpenCV's (votes desc, index asc)
__global__ void __launch_bounds__(256)
k_hough_peaks(const int* __restrict__ accum, u64* __restrict__ keys, FrameCtl* __restrict__ ctl, int pass,
              HoughCfg hc, size_t accum_stride, size_t key_stride)
{
    int f = blockIdx.y >> 1, which = blockIdx.y & 1;
    if (!ctl[f].active[pass] || !ctl[f].hough[pass]) return;
    const int* A = accum + ((size_t)f * 2 + which) * accum_stride;
    u64* K = keys + ((size_t)f * 2 + which) * key_stride;
    int cells = hc.numangle * hc.numrho;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
        int n = i / hc.numrho, r = i - n * hc.numrho;
        int base = (n + 1) * hc.RS + r + 1;
        int v = A[base];
        if (v > hc.threshold && v > A[base - 1] && v >= A[base + 1] && v > A[base - hc.RS] && v >= A[base + hc.RS]) {
            int pos = atomicAdd(&ctl[f].npeaks[which], 1);
            K[pos] = ((u64)(u32)(0x7fffffff - v) << 32) | (u32)base;
        }
    }
}

__device__ __forceinline__ void key_to_line(u64 key, const HoughCfg& hc, float* rho, float* theta)
{
    int base = (int)(key & 0xffffffffu);
    int n = base / hc.RS - 1;
    int r = base - (n + 1) * hc.RS - 1;
    *rho = __fmul_rn((float)(r - (hc.numrho - 1) / 2), hc.rho);
    *theta = __fmul_rn((float)n, hc.theta);
}

// first K lines in OpenCV order without a full sort: K rounds of block-wide minimum over the keys.
// grid = (2, n); writes lfd_result.top_equ / top_box and n_lines_*.
__global__ void __launch_bounds__(256)
k_hough_topk(const u64* __restrict__ keys, lfd_result* __restrict__ res, const FrameCtl* __restrict__ ctl,
             int pass, HoughCfg hc, size_t key_stride, int K)
{
    int which = blockIdx.x, f = blockIdx.y;
    if (!ctl[f].active[pass]) return;
    lfd_result* R = res + f;
    if (!ctl[f].hough[pass]) {
        if (threadIdx.x == 0) { if (which) R->n_lines_box[pass] = -1; else R->n_lines_equ[pass] = -1; }
        return;
    }
    const u64* Kk = keys + ((size_t)f * 2 + which) * key_stride;
    int np = ctl[f].npeaks[which];
    __shared__ u64 sm[256];
    __shared__ u64 prev_s;
    u64 prev = 0;
    bool have_prev = false;
    for (int k = 0; k < K; k++) {
        u64 best = ~0ull;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            u64 v = Kk[i];
            if ((!have_prev || v > prev) && v < best) best = v;
        }
        sm[threadIdx.x] = best;
        __syncthreads();
        for (int o = 128; o; o >>= 1) {
            if (threadIdx.x < o) { u64 b = sm[threadIdx.x + o]; if (b < sm[threadIdx.x]) sm[threadIdx.x] = b; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            prev_s = sm[0];
            float* dst = which ? R->top_box[pass][k] : R->top_equ[pass][k];
            if (sm[0] != ~0ull) key_to_line(sm[0], hc, &dst[0], &dst[1]);
            else { dst[0] = 0.f; dst[1] = 0.f; }
        }
        __syncthreads();
        prev = prev_s;
        have_prev = true;
        __syncthreads();
    }
    if (threadIdx.x == 0) { if (which) R->n_lines_box[pass] = np; else R->n_lines_equ[pass] = np; }
}

// full sort (taps / lfd_hough_lines): bitonic sort of the keys in global memory by one CTA,
// then (rho, theta) for every line.  key buffer must have room for next_pow2(npeaks).
__global__ void __launch_bounds__(1024)
k_hough_sort(u64* __restrict__ keys, float* __restrict__ lines, const FrameCtl* __restrict__ ctl, int pass,
             HoughCfg hc, size_t key_stride, size_t line_stride, int max_lines)
{
    int which = blockIdx.x, f = blockIdx.y;
    if (!ctl[f].active[pass] || !ctl[f].hough[pass]) return;
    u64* Kk = keys + ((size_t)f * 2 + which) * key_stride;
    int np = ctl[f].npeaks[which];
    int n2 = 1;
    while (n2 < np) n2 <<= 1;
    for (int i = np + threadIdx.x; i < n2; i += blockDim.x) Kk[i] = ~0ull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    u64 a = Kk[i], b = Kk[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { Kk[i] = b; Kk[ixj] = a; }
                }
            }
            __syncthreads();
        }
    float* L = lines + ((size_t)f * 2 + which) * line_stride;
    for (int i = threadIdx.x; i < np && i < max_lines; i += blockDim.x) key_to_line(Kk[i], hc, &L[2 * i], &L[2 * i + 1]);
}

// check_theta (processfield.py:36-150) + the pass verdict; one thread per frame.
// float64 arithmetic on float32-valued inputs like the NumPy original; a short second set leaves
// theta1[i] at 0 as well (the four assignments share one try block, :93-102).
__global__ void k_check_theta(lfd_result* __restrict__ res, FrameCtl* __restrict__ ctl, int pass, int n,
                              int navg, double dro, double thetaTresh, double lineSetTresh, int numangle,
                              unsigned long long* __restrict__ counters)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    lfd_result* R = res + f;
    if (!ctl[f].active[pass]) return;
    R->rect_detection[pass] = ctl[f].hough[pass];
    bool detected = false;
    if (ctl[f].hough[pass]) {
        int n1 = R->n_lines_equ[pass], n2 = R->n_lines_box[pass];
        if (n1 <= 0) atomicOr(&ctl[f].status, LFD_FRAME_NO_LINES_EQU);
        if (n2 <= 0) atomicOr(&ctl[f].status, LFD_FRAME_NO_LINES_BOX);
        if (n1 > 0 && n2 > 0) {
            double ro1[LFD_MAX_SET_LINES], ro2[LFD_MAX_SET_LINES], t1[LFD_MAX_SET_LINES], t2[LFD_MAX_SET_LINES];
            for (int i = 0; i < navg; i++) {
                ro1[i] = ro2[i] = t1[i] = t2[i] = 0.0;
                if (i < n1) {
                    ro1[i] = (double)R->top_equ[pass][i][0];
                    if (i < n2) {
                        ro2[i] = (double)R->top_box[pass][i][0];
                        t1[i] = (double)R->top_equ[pass][i][1];
                        t2[i] = (double)R->top_box[pass][i][1];
                    }
                }
            }
            double s1 = 0, s2 = 0, mx1 = t1[0], mn1 = t1[0], mx2 = t2[0], mn2 = t2[0], sd = 0;
            for (int i = 0; i < navg; i++) {
                s1 += ro1[i]; s2 += ro2[i];
                mx1 = fmax(mx1, t1[i]); mn1 = fmin(mn1, t1[i]);
                mx2 = fmax(mx2, t2[i]); mn2 = fmin(mn2, t2[i]);
                sd += fabs(t1[i] - t2[i]);
            }
            bool reject = false;
            if (fabs(s1 / navg - s2 / navg) > dro) reject = true;
            else if (fabs(mx1 - mn1) > thetaTresh) reject = true;
            else if (fabs(mx2 - mn2) > thetaTresh) reject = true;
            else if (sd / navg > lineSetTresh) reject = true;
            R->rejected[pass] = reject ? 1 : 0;
            detected = !reject;
        }
    }
    // the frame's verdict (which pass counts, rho/theta, status) is assembled by k_finalize: the two passes run
    // concurrently and only write their own fields
    ctl[f].detected = detected ? 1 : 0;
    // end-of-pass bookkeeping: contour counts for the RECTS tap, work counters of the batch
    FrameCtl* c = ctl + f;
    c->ncomp_saved[pass][0] = c->ncomp[0]; c->ncomp_saved[pass][1] = c->ncomp[1];
    atomicAdd(&counters[0], (unsigned long long)c->nnz[0]);
    atomicAdd(&counters[1], (unsigned long long)c->nnz[1]);
    atomicAdd(&counters[2], (unsigned long long)(c->nnz[0] + c->nnz[1]) * numangle);
    atomicAdd(&counters[3], (unsigned long long)c->nruns[0]);
    atomicAdd(&counters[4], (unsigned long long)c->nruns[1]);
    atomicAdd(&counters[5], (unsigned long long)(c->ncomp[0] + c->ncomp[1]));
    atomicAdd(&counters[6], (unsigned long long)c->npass);
    if (c->hough[pass]) atomicAdd(&counters[8], 1ull);
    atomicAdd(&counters[9 + pass], 1ull);
}


// Peak shared-memory atomic throughput of this GPU (the roofline the Hough vote kernel is measured against):
// every lane adds to its own bank-conflict-free word, `iters` times, on all SMs.
__global__ void __launch_bounds__(256)
k_smem_atomic_peak(int iters, unsigned* sink)
{
    __shared__ unsigned acc[256 * 8];
    for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    unsigned* p = acc + threadIdx.x;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) atomicAdd(p + 256 * k, 1u);
    }
    __syncthreads();
    unsigned v = 0;
    for (int k = 0; k < 8; k++) v += acc[threadIdx.x + 256 * k];
    if (v == 0xffffffffu) sink[0] = v;       // never true: keeps the loop alive
}
