// lfd_b200.cu - handle, device memory layout, pipeline orchestration and the C ABI of include/lfd_b200.h.
//
// One handle = one GPU, one stream, buffers for `max_batch` frames.  A batch goes through
//   star mask -> prep (blot+flip+clip+u8+hist, both passes in one read of the float frame)
//   per pass: LUT -> morphology -> Sobel/NMS -> run-CCL (hysteresis + outer contours) -> run-CCL of the
//   background (hole contours) -> minAreaRect/filter/boxPoints -> box fill -> Hough x2 -> check_theta
// with every kernel gridded over (work, frame); frames a pass does not apply to exit at the first line
// (bright detected -> no dim; no rectangle passed -> no Hough), so there is no host round trip inside a batch.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <fcntl.h>
#include <sys/file.h>
#include <time.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "common.cuh"
#include "geom.cuh"
#include "k_canny.cuh"
#include "k_ccl.cuh"
#include "k_hough.cuh"
#include "k_morph.cuh"
#include "k_mnms.cuh"
#include "k_prep.cuh"
#include "k_rects.cuh"
#include "host_ingest.cuh"

static std::string g_create_error;

// dynamic shared-memory opt-in (cudaFuncSetAttribute) is per function AND per device: high-water marks per device
#define LFD_MAX_DEVICES 64
#define LFD_MAX_SPLIT 4
struct DevLimits { size_t band_max = 48 * 1024, rects_max = 48 * 1024, anyk_max = 48 * 1024; };
static DevLimits g_dev_limits[LFD_MAX_DEVICES];

enum { T_PREP = 0, T_MORPH, T_CANNY, T_CCL_FG, T_CCL_BG, T_RECTS, T_HOUGH, T_CHECK, T_PER_PASS };
static const char* k_timing_names[] = {
    "setup(memset+star_mask)",
    "prep(blot+flip+clip+u8+hist)",
    "bright:lut+morph(+sobel+nms when fused)", "bright:sobel+nms", "bright:ccl_fg(hysteresis)", "bright:ccl_bg(holes)",
    "bright:rects+boxfill", "bright:hough", "bright:check_theta",
    "dim:lut+morph(+sobel+nms when fused)", "dim:sobel+nms", "dim:ccl_fg(hysteresis)", "dim:ccl_bg(holes)",
    "dim:rects+boxfill", "dim:hough", "dim:check_theta",
    "results_d2h"};
#define N_TIMINGS 17

// Brackets: a CUDA event on each side of the launches that carry the step (recorded on the launching stream, also
// inside the captured graph), so that bench.py can quote the average launch duration of the time-dominant kernel
// measured live in the timed region (lfd_get_kernel_times).
enum { KB_MORPH = 0, KB_NMS, KB_CCL_BAND_FG, KB_CCL_BAND_BG, KB_RECTS, KB_HOUGH_VOTE, KB_COUNT };
static const char* k_bracket_names[KB_COUNT] = {"k_morph_march", "k_nms_march", "k_ccl_band(fg)", "k_ccl_band(bg)",
                                                "k_rects_warp", "k_hough_vote"};

struct HoughBufs {
    HoughCfg hc;
    float* tabSin = nullptr;   // device
    float* tabCos = nullptr;
    size_t accum_stride = 0, key_stride = 0, line_stride = 0;
    int max_lines = 0;
    int* accum = nullptr;      // [B][2][accum_stride]
    u64* keys = nullptr;       // [B][2][key_stride]
    float* lines = nullptr;    // [B][2][line_stride]  (allocated on first FULL_LINES use)
    size_t smem = 0;
};

struct lfd_handle {
    int device = 0, B = 0, sm_count = 148;
    Dims d;
    lfd_config cfg;
    lfd_params params;
    bool have_params = false;
    cudaStream_t stream = nullptr;        // uploads, prep, bright pass, results
    // the batch is cut in `nsplit` parts; part k runs its bright pass on xs[k][0] and its dim pass on xs[k][1]
    // (xs[0][0] is `stream` itself); every extra stream joins `stream` through its own event
    cudaStream_t xs[LFD_MAX_SPLIT][2];
    cudaEvent_t xjoin[LFD_MAX_SPLIT][2];
    cudaEvent_t ev_fork = nullptr;
    int nsplit_cfg = 2;                   // env LFD_NSPLIT (1..LFD_MAX_SPLIT)
    // host->device copy gate (lfd_set_h2d_gate): an advisory file lock shared with the other ranks of the same host bridge
    // slot, held from the moment the batch copy is enqueued until it has completed (released by a stream host callback)
    int gate_fd = -1;
    int64_t gate_waits = 0, gate_timeouts = 0;
    std::string err;
    int64_t launches = 0;
    int last_n = 0, last_flags = 0;
    bool pending = false;

    // device
    float* in = nullptr;              // [B][N]
    u32* starmask = nullptr;          // [B][NW]
    int4* rects_d = nullptr;          // [B*max_star_rects]
    int* rect_off_d = nullptr;        // [B+1]
    u8* gray[2] = {nullptr, nullptr}; // [B][N]
    u32* hist = nullptr;              // [2][B][256]
    u8* lut = nullptr;                // [2][B][256]
    u8* morph[2] = {nullptr, nullptr};
    u32* nz[2] = {nullptr, nullptr};  // [B][NW]
    u32* cand[2] = {nullptr, nullptr};
    u32* strong[2] = {nullptr, nullptr};
    u32* edges[2] = {nullptr, nullptr};
    u32* box[2] = {nullptr, nullptr};
    u8* tap_u8 = nullptr;             // [N] scratch for expanded taps
    u8* eroded_tap = nullptr;         // [B][N] lazily
    u8* nms_tap[2] = {nullptr, nullptr};
    float* clipped = nullptr;         // [N] lazily (run_pass writeback)
    int* labels_tap = nullptr;        // [N] lazily
    FrameCtl* ctl = nullptr;          // [2][B]: one bookkeeping block per pass (the passes run concurrently)
    lfd_result* res_d = nullptr;      // [B]
    int64_t* counters_d = nullptr;    // [16]
    uint2* segs[2] = {nullptr, nullptr};     // per pass: [B][2][NW]
    CclBuf* ccl_d[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [pass][kind] -> [B]
    CompBuf* comp_d[2] = {nullptr, nullptr}; // per pass: [B]
    RectBuf* rbuf_d[2] = {nullptr, nullptr}; // per pass: [B]
    std::vector<CclBuf> ccl_h[2][2];
    std::vector<CompBuf> comp_h[2];
    std::vector<RectBuf> rbuf_h[2];
    std::vector<void*> allocs;
    HoughBufs hb[2];

    // host
    float* frames_h = nullptr;        // pinned [B][N]
    lfd_result* res_h = nullptr;      // pinned [B]
    int4* rects_h = nullptr;          // pinned
    int* rect_off_h = nullptr;        // pinned
    FrameCtl* ctl_h = nullptr;        // pinned [2][B]
    int64_t* counters_h = nullptr;    // pinned [16] (a pageable destination would make the D2H copy block the host)

    cudaEvent_t ev[N_TIMINGS + 1];
    bool ev_valid[N_TIMINGS + 1];
    float timings[N_TIMINGS];
    cudaEvent_t kb[LFD_MAX_SPLIT][2][KB_COUNT][2];    // [part][pass][kernel][begin / end], created on first use
    bool kb_used[LFD_MAX_SPLIT][2][KB_COUNT];         // recorded by the last run
    float kb_ms[2][KB_COUNT];                         // last run: per pass, summed over the parts
    int kb_launches[2][KB_COUNT];
    int cur_part = 0;                                 // part index of the run_pass_kernels call in progress
    // CUDA graph of the two passes + verdict + result copies (whole-frame path without taps): ~100 small launches
    // become one graph launch, which removes the per-launch gaps between the latency-bound kernels
    bool use_graphs = true;
    cudaGraphExec_t graph_exec = nullptr;
    int graph_n = 0, graph_flags = 0;
    int64_t graph_launches = 0;
    bool stage_timings_valid = true;
    // arbitrary structuring elements (lfd_set_kernels): [pass][0 erode / 1 dilate]; n == 0 -> the all-ones rectangle of lfd_pass_params
    AnyKernel* anyk_d = nullptr;      // [2][2]
    AnyKernel anyk_h[2][2];
    bool anyk_on[2] = {false, false};
    int prep_grid[3] = {0, 0, 0};     // resident CTAs of k_prep<mode> on this device (one full wave)
    // fused morphology + Sobel + NMS (k_mnms.cuh): 3-D tensor maps over gray[pass][frame][y][x], box 256 x FZ_RB x 1
    CUtensorMap tm_gray[2];
    bool fused_ok = false;            // frame shape allows the tensor maps (W % 16 == 0, W >= 256) and LFD_NO_FUSED is unset
    int fused_case[2] = {-1, -1};     // per pass: index into the instantiated kernel shapes, -1 = unfused kernels
    cudaEvent_t mark[4];              // caller-placed timestamps on the handle's stream (lfd_timer_mark)
    bool mark_valid[4];
    // developer aid (env LFD_KTIMING=1): one event after every launch, reported by source line
    bool ktiming = false;
    std::vector<cudaEvent_t> kev;
    std::vector<int> kline;
    int kn = 0;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            return LFD_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

#define LAUNCH_CHECK()                                                                             \
    do {                                                                                           \
        h->launches++;                                                                             \
        if (h->ktiming) ktime_mark(h, __LINE__);                                                   \
        cudaError_t e_ = cudaGetLastError();                                                       \
        if (e_ != cudaSuccess) {                                                                   \
            h->err = std::string("kernel launch: ") + cudaGetErrorString(e_) + " at line " + std::to_string(__LINE__); \
            return LFD_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

static void ktime_mark(lfd_handle* h, int line)
{
    if (h->kn >= (int)h->kev.size()) { cudaEvent_t e; cudaEventCreate(&e); h->kev.push_back(e); h->kline.push_back(0); }
    h->kline[h->kn] = line;
    cudaEventRecord(h->kev[h->kn], h->stream);
    h->kn++;
}

template <typename T>
static int dev_alloc(lfd_handle* h, T** p, size_t count)
{
    void* q = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    CK(cudaMalloc(&q, bytes));
    h->allocs.push_back(q);
    *p = (T*)q;
    return LFD_OK;
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ------------------------------------------------------------------------------------------------
// small bookkeeping kernels
// ------------------------------------------------------------------------------------------------
// ctl: [2][B] (one block per pass); block p only ever looks at active[p] / hough[p]
__global__ void k_ctl_init(FrameCtl* ctl, int B, lfd_result* res, int n, int act0, int act1)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    FrameCtl c;
    memset(&c, 0, sizeof(c));
    c.active[0] = act0; c.active[1] = act1;
    ctl[f] = c;
    ctl[B + f] = c;
    lfd_result r;
    memset(&r, 0, sizeof(r));
    r.pass = -1;
    for (int p = 0; p < 2; p++) { r.rect_detection[p] = -1; r.n_lines_equ[p] = -1; r.n_lines_box[p] = -1; }
    res[f] = r;
}

// Verdict of the frame from the two passes' private results (detecttrails.py:125-131): bright wins; the dim
// pass counts only where bright neither detected nor failed, and is reported as "not run" elsewhere.
// mode 0: both passes ; 1: bright only ; 2: dim only
__global__ void k_finalize(const FrameCtl* __restrict__ ctl, int B, lfd_result* __restrict__ res, int n, int mode,
                           unsigned long long* counters)
{
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const FrameCtl& c0 = ctl[f];
    const FrameCtl& c1 = ctl[B + f];
    lfd_result* R = res + f;
    const int errbits = LFD_FRAME_NO_LINES_EQU | LFD_FRAME_NO_LINES_BOX | LFD_FRAME_OVERFLOW;
    const bool ran0 = mode != 2, ran1 = mode != 1;
    const bool use1 = ran1 && !(ran0 && (c0.detected || (c0.status & errbits)));
    int status = ran0 ? c0.status : 0;
    if (use1) status |= c1.status;
    R->status = status;
    int pass = -1;
    if (ran0 && c0.detected) pass = 0;
    else if (use1 && c1.detected) pass = 1;
    R->detected = pass >= 0;
    R->pass = pass;
    if (pass >= 0) { R->rho = R->top_equ[pass][0][0]; R->theta = R->top_equ[pass][0][1]; }
    if (ran1 && !use1) {      // what a run that skipped the dim pass reports
        R->rect_detection[1] = -1; R->n_lines_equ[1] = -1; R->n_lines_box[1] = -1; R->rejected[1] = 0;
        for (int i = 0; i < LFD_MAX_SET_LINES; i++) {
            R->top_equ[1][i][0] = R->top_equ[1][i][1] = 0.f;
            R->top_box[1][i][0] = R->top_box[1][i][1] = 0.f;
        }
    }
    if (mode == 0 && use1) atomicAdd(&counters[7], 1ull);
}

__global__ void k_pack_mask(const u8* __restrict__ img, u32* __restrict__ mask, Dims d)
{
    // one warp per 32 pixels
    int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int i = gw; i < d.NW; i += nwarps) {
        int y = i / d.WW, w = i - y * d.WW;
        int x = (w << 5) + lane;
        bool nzp = x < d.W && img[(size_t)y * d.W + x] != 0;
        u32 b = __ballot_sync(FULLMASK, nzp);
        if (lane == 0) mask[i] = b;
    }
}

// ------------------------------------------------------------------------------------------------
// Hough configuration (host): OpenCV's numangle/numrho and float trig tables
// ------------------------------------------------------------------------------------------------
static void hough_dims_host(int H, int W, double rho_, double theta_, int* numangle, int* numrho)
{
    float rho = (float)rho_, theta = (float)theta_;
    int max_rho = W + H, min_rho = -max_rho;
    double min_theta = 0, max_theta = M_PI;
    int na = (int)floor((max_theta - min_theta) / theta) + 1;
    if (na > 1 && fabs(M_PI - (na - 1) * theta) < theta / 2) --na;
    *numangle = na;
    *numrho = (int)lrint(((max_rho - min_rho) + 1) / rho);
}

static int hough_setup(lfd_handle* h, HoughBufs* hb, int H, int W, double rho_, double theta_, int threshold,
                       int nframes, bool want_lines, int max_lines_cfg)
{
    if (!(rho_ > 0) || !(theta_ > 0)) { h->err = "HoughLines: rho and theta must be positive"; return LFD_E_ARG; }
    HoughCfg hc;
    hough_dims_host(H, W, rho_, theta_, &hc.numangle, &hc.numrho);
    if (hc.numangle < 1 || hc.numrho < 1 || (double)hc.numangle * hc.numrho > 2.0e8) {
        h->err = "HoughLines: accumulator size out of range";
        return LFD_E_ARG;
    }
    hc.RS = hc.numrho + 2;
    hc.rho = (float)rho_;
    hc.theta = (float)theta_;
    hc.threshold = threshold;
    const size_t smem_budget = 96 * 1024;
    int apc = (int)(smem_budget / ((size_t)hough_rss(hc.RS) * 4));
    if (apc > 32) apc = 32;
    if (apc < 1) { h->err = "HoughLines: rho too fine for the shared-memory accumulator (numrho too large)"; return LFD_E_UNSUPPORTED; }
    if (apc < 32) { int p = 1; while (p * 2 <= apc) p *= 2; apc = p; }   // power of two so sub-warps tile a warp
    if (apc > hc.numangle) apc = hc.numangle < 32 ? next_pow2(hc.numangle) : 32;
    hc.apc = apc;
    hc.ngroups = (hc.numangle + apc - 1) / apc;
    hb->smem = (size_t)apc * hough_rss(hc.RS) * 4;
    hb->hc = hc;
    // tables, exactly as OpenCV builds them (float accumulation of the angle)
    std::vector<float> ts(hc.numangle), tc(hc.numangle);
    float irho = 1 / hc.rho;
    float ang = 0.f;
    for (int n = 0; n < hc.numangle; ang += hc.theta, n++) {
        ts[n] = (float)(sin((double)ang) * irho);
        tc[n] = (float)(cos((double)ang) * irho);
    }
    size_t accum_stride = (size_t)(hc.numangle + 2) * hc.RS;
    size_t cells = (size_t)hc.numangle * hc.numrho;
    size_t key_stride = (size_t)next_pow2((int)cells);
    bool realloc_needed = (accum_stride != hb->accum_stride) || (key_stride != hb->key_stride) || !hb->accum;
    if (realloc_needed) {
        if (hb->accum) cudaFree(hb->accum);
        if (hb->keys) cudaFree(hb->keys);
        if (hb->lines) cudaFree(hb->lines);
        if (hb->tabSin) cudaFree(hb->tabSin);
        hb->accum = nullptr; hb->keys = nullptr; hb->lines = nullptr; hb->tabSin = nullptr;
        CK(cudaMalloc((void**)&hb->accum, (size_t)nframes * 2 * accum_stride * sizeof(int)));
        CK(cudaMalloc((void**)&hb->keys, (size_t)nframes * 2 * key_stride * sizeof(u64)));
        CK(cudaMalloc((void**)&hb->tabSin, (size_t)2 * hc.numangle * sizeof(float)));
        hb->tabCos = hb->tabSin + hc.numangle;
        hb->accum_stride = accum_stride;
        hb->key_stride = key_stride;
    }
    hb->max_lines = (max_lines_cfg > 0 && (size_t)max_lines_cfg < cells) ? max_lines_cfg : (int)cells;
    hb->line_stride = (size_t)2 * hb->max_lines;
    if (want_lines && !hb->lines) CK(cudaMalloc((void**)&hb->lines, (size_t)nframes * 2 * hb->line_stride * sizeof(float)));
    CK(cudaMemcpyAsync(hb->tabSin, ts.data(), hc.numangle * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(hb->tabCos, tc.data(), hc.numangle * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    // the limit is per function, not per launch: both passes (and lfd_hough_lines) share k_hough_vote
    CK(cudaFuncSetAttribute(k_hough_vote<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_budget));
    CK(cudaFuncSetAttribute(k_hough_vote<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_budget));
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------------
extern "C" int lfd_abi_version(void) { return LFD_ABI_VERSION; }

extern "C" const char* lfd_last_error(const lfd_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int lfd_destroy(lfd_handle* h)
{
    if (!h) return LFD_E_ARG;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int k = 0; k < LFD_MAX_SPLIT; k++)
        for (int p = 0; p < 2; p++)
            if (h->xs[k][p] && h->xs[k][p] != h->stream) cudaStreamSynchronize(h->xs[k][p]);
    if (h->gate_fd >= 0) { flock(h->gate_fd, LOCK_UN); close(h->gate_fd); h->gate_fd = -1; }
    for (void* p : h->allocs) cudaFree(p);
    for (int p = 0; p < 2; p++) {
        if (h->hb[p].accum) cudaFree(h->hb[p].accum);
        if (h->hb[p].keys) cudaFree(h->hb[p].keys);
        if (h->hb[p].lines) cudaFree(h->hb[p].lines);
        if (h->hb[p].tabSin) cudaFree(h->hb[p].tabSin);
    }
    if (h->frames_h) cudaFreeHost(h->frames_h);
    if (h->res_h) cudaFreeHost(h->res_h);
    if (h->rects_h) cudaFreeHost(h->rects_h);
    if (h->rect_off_h) cudaFreeHost(h->rect_off_h);
    if (h->ctl_h) cudaFreeHost(h->ctl_h);
    if (h->counters_h) cudaFreeHost(h->counters_h);
    for (int i = 0; i <= N_TIMINGS; i++) if (h->ev_valid[i]) cudaEventDestroy(h->ev[i]);
    for (int k = 0; k < LFD_MAX_SPLIT; k++)
        for (int p = 0; p < 2; p++)
            for (int b = 0; b < KB_COUNT; b++)
                for (int e = 0; e < 2; e++)
                    if (h->kb[k][p][b][e]) cudaEventDestroy(h->kb[k][p][b][e]);
    for (int i = 0; i < 4; i++) if (h->mark_valid[i]) cudaEventDestroy(h->mark[i]);
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    for (int k = 0; k < LFD_MAX_SPLIT; k++)
        for (int p = 0; p < 2; p++) {
            if (h->xjoin[k][p]) cudaEventDestroy(h->xjoin[k][p]);
            if (h->xs[k][p] && h->xs[k][p] != h->stream) cudaStreamDestroy(h->xs[k][p]);
        }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return LFD_OK;
}

static int create_impl(lfd_handle* h, int device, int max_batch, int H, int W, const lfd_config* cfg)
{
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { h->err = "no such CUDA device"; return LFD_E_CUDA; }
    CK(cudaSetDevice(device));
    h->device = device;
    { const char* kt = getenv("LFD_KTIMING"); h->ktiming = kt && kt[0] == '1'; }
    { const char* ng = getenv("LFD_NO_GRAPH"); h->use_graphs = !(ng && ng[0] == '1'); }
    if (max_batch < 1 || H < 8 || W < 8) { h->err = "bad batch or frame size"; return LFD_E_ARG; }
    if (W % 4 != 0) { h->err = "frame width must be a multiple of 4"; return LFD_E_UNSUPPORTED; }
    if (W > 4096 || H > 65535) { h->err = "frame too large (W <= 4096, H <= 65535)"; return LFD_E_UNSUPPORTED; }
    h->B = max_batch;
    h->d.H = H; h->d.W = W; h->d.WW = (W + 31) / 32; h->d.N = H * W; h->d.NW = H * h->d.WW;
    memset(&h->cfg, 0, sizeof(h->cfg));
    if (cfg) h->cfg = *cfg;
    if (h->cfg.max_runs <= 0) h->cfg.max_runs = 1 << 19;
    if (h->cfg.max_components <= 0) h->cfg.max_components = 1 << 16;
    if (h->cfg.max_star_rects <= 0) h->cfg.max_star_rects = 8192;
    // a frame cannot hold more runs than H * ceil(W/2)
    long long worst = (long long)H * ((W + 1) / 2);
    if (h->cfg.max_runs > worst) h->cfg.max_runs = (int)worst;
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (int k = 0; k < LFD_MAX_SPLIT; k++)
        for (int p = 0; p < 2; p++) {
            if (k == 0 && p == 0) { h->xs[0][0] = h->stream; continue; }
            CK(cudaStreamCreateWithFlags(&h->xs[k][p], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&h->xjoin[k][p], cudaEventDisableTiming));
        }
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    { const char* ns = getenv("LFD_NSPLIT"); if (ns && ns[0] >= '1' && ns[0] <= '0' + LFD_MAX_SPLIT) h->nsplit_cfg = ns[0] - '0'; }
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        h->sm_count = prop.multiProcessorCount;
        int per_sm[3] = {0, 0, 0};
        // k_prep keeps its TMA rings in dynamic shared memory (above the 48 KB default limit)
#define PREP_ATTR(M, E) CK(cudaFuncSetAttribute(k_prep<M, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM))
        PREP_ATTR(0, false); PREP_ATTR(0, true); PREP_ATTR(1, false); PREP_ATTR(1, true); PREP_ATTR(2, false); PREP_ATTR(2, true);
#undef PREP_ATTR
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], k_prep<0, false>, PR_WARPS * 32, PR_SMEM));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], k_prep<1, false>, PR_WARPS * 32, PR_SMEM));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[2], k_prep<2, false>, PR_WARPS * 32, PR_SMEM));
        for (int m = 0; m < 3; m++) h->prep_grid[m] = prop.multiProcessorCount * (per_sm[m] > 0 ? per_sm[m] : 1);
    }
    // the dynamic shared-memory limit is a property of the FUNCTION on one DEVICE, shared by every handle of the process
    // on that device: only ever raise it, and keep the high-water mark per device
    {
        DevLimits& L = g_dev_limits[device % LFD_MAX_DEVICES];
        if (ccl_band_smem(h->d.WW) > L.band_max) {
            CK(cudaFuncSetAttribute(k_ccl_band, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ccl_band_smem(h->d.WW)));
            L.band_max = ccl_band_smem(h->d.WW);
        }
        if (rects_smem(max_batch) > L.rects_max) {
            CK(cudaFuncSetAttribute(k_rects_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rects_smem(max_batch)));
            L.rects_max = rects_smem(max_batch);
        }
    }
    for (int i = 0; i <= N_TIMINGS; i++) { CK(cudaEventCreate(&h->ev[i])); h->ev_valid[i] = true; }
    for (int i = 0; i < 4; i++) { CK(cudaEventCreate(&h->mark[i])); h->mark_valid[i] = true; }
    for (int k = 0; k < LFD_MAX_SPLIT; k++)
        for (int p = 0; p < 2; p++)
            for (int b = 0; b < KB_COUNT; b++)
                for (int e = 0; e < 2; e++) CK(cudaEventCreate(&h->kb[k][p][b][e]));

    const int B = h->B;
    const size_t N = h->d.N, NW = h->d.NW;
    int rc;
#define DA(ptr, count) if ((rc = dev_alloc(h, &(ptr), (count))) != LFD_OK) return rc
    DA(h->in, (size_t)B * N);
    DA(h->starmask, (size_t)B * NW);
    DA(h->rects_d, (size_t)B * h->cfg.max_star_rects);
    DA(h->rect_off_d, (size_t)B + 1);
    DA(h->hist, (size_t)2 * B * 256);
    DA(h->lut, (size_t)2 * B * 256);
    for (int p = 0; p < 2; p++) {
        DA(h->gray[p], (size_t)B * N);
        DA(h->morph[p], (size_t)B * N);
        DA(h->nz[p], (size_t)B * NW);
        DA(h->cand[p], (size_t)B * NW);
        DA(h->strong[p], (size_t)B * NW);
        DA(h->edges[p], (size_t)B * NW);
        DA(h->box[p], (size_t)B * NW);
    }
    {
        // Fused morphology + Sobel + NMS (k_mnms.cuh) is opt-in (LFD_FUSED=1): measured on B200 it executes as many
        // instructions as the k_morph_march + k_nms_march pair and runs longer (profiles/r02c_*), so the pair stays the
        // production path.  It needs 16-byte row strides for the tensor map and a frame at least one box wide.
        const char* nf = getenv("LFD_FUSED");
        h->fused_ok = (nf && nf[0] == '1') && (W % 16) == 0 && W >= FZ_BOXW;
        h->fused_case[0] = h->fused_case[1] = -1;
    }
    DA(h->tap_u8, N);
    DA(h->ctl, (size_t)2 * B);
    DA(h->anyk_d, 4);
    memset(h->anyk_h, 0, sizeof(h->anyk_h));
    DA(h->res_d, (size_t)B);
    DA(h->counters_d, 16);
    DA(h->segs[0], (size_t)B * 2 * NW);
    DA(h->segs[1], (size_t)B * 2 * NW);
    const int MR = h->cfg.max_runs, MC = h->cfg.max_components;
    const int slotcap = 2 * MR + 4 * MC, hullcap = 2 * slotcap + 4 * MC;
    // one slab per field, sliced per frame; everything the two passes touch exists once per pass
    for (int p = 0; p < 2; p++)
    for (int k = 0; k < 2; k++) {
        h->ccl_h[p][k].resize(B);
        Run* runs; int *parent, *flag, *ymax, *compidx, *rowbase, *rowcnt; u16* wpre;
        DA(runs, (size_t)B * MR); DA(parent, (size_t)B * MR); DA(flag, (size_t)B * MR); DA(ymax, (size_t)B * MR);
        DA(compidx, (size_t)B * MR); DA(rowbase, (size_t)B * (H + 1)); DA(wpre, (size_t)B * NW); DA(rowcnt, (size_t)B * H);
        for (int f = 0; f < B; f++) {
            CclBuf& c = h->ccl_h[p][k][f];
            c.runs = runs + (size_t)f * MR; c.parent = parent + (size_t)f * MR; c.flag = flag + (size_t)f * MR;
            c.ymax = ymax + (size_t)f * MR; c.compidx = compidx + (size_t)f * MR;
            c.rowbase = rowbase + (size_t)f * (H + 1); c.wpre = wpre + (size_t)f * NW; c.rowcnt = rowcnt + (size_t)f * H;
        }
        DA(h->ccl_d[p][k], (size_t)B);
        CK(cudaMemcpy(h->ccl_d[p][k], h->ccl_h[p][k].data(), B * sizeof(CclBuf), cudaMemcpyHostToDevice));
    }
    for (int p = 0; p < 2; p++) {
        h->comp_h[p].resize(B);
        int *root, *y0, *hh, *slot, *hulloff, *rowmin, *rowmax;
        DA(root, (size_t)B * 2 * MC); DA(y0, (size_t)B * 2 * MC); DA(hh, (size_t)B * 2 * MC);
        DA(slot, (size_t)B * 2 * MC); DA(hulloff, (size_t)B * 2 * MC);
        DA(rowmin, (size_t)B * slotcap); DA(rowmax, (size_t)B * slotcap);
        for (int f = 0; f < B; f++) {
            CompBuf& c = h->comp_h[p][f];
            c.root = root + (size_t)f * 2 * MC; c.y0 = y0 + (size_t)f * 2 * MC; c.h = hh + (size_t)f * 2 * MC;
            c.slot = slot + (size_t)f * 2 * MC; c.hulloff = hulloff + (size_t)f * 2 * MC;
            c.rowmin = rowmin + (size_t)f * slotcap; c.rowmax = rowmax + (size_t)f * slotcap;
            c.slotcap = slotcap; c.hullcap = hullcap; c.maxcomp = MC;
        }
        DA(h->comp_d[p], (size_t)B);
        CK(cudaMemcpy(h->comp_d[p], h->comp_h[p].data(), B * sizeof(CompBuf), cudaMemcpyHostToDevice));
    }
    for (int p = 0; p < 2; p++) {
        h->rbuf_h[p].resize(B);
        lfdgeom::Pt* hulls; float* hullfs;
        DA(hulls, (size_t)B * hullcap); DA(hullfs, (size_t)B * 3 * hullcap);
        lfd_rect* rects; int* passing;
        DA(rects, (size_t)B * 2 * MC); DA(passing, (size_t)B * 2 * MC);
        for (int f = 0; f < B; f++) {
            RectBuf& r = h->rbuf_h[p][f];
            r.rects = rects + (size_t)f * 2 * MC; r.passing = passing + (size_t)f * 2 * MC;
            r.hull = hulls + (size_t)f * hullcap; r.hullf = hullfs + (size_t)f * 3 * hullcap;
        }
        DA(h->rbuf_d[p], (size_t)B);
        CK(cudaMemcpy(h->rbuf_d[p], h->rbuf_h[p].data(), B * sizeof(RectBuf), cudaMemcpyHostToDevice));
    }
#undef DA
    // Pinned staging of the frames.  LFD_STAGING=wc asks for write-combined pages: the loaders' stores then bypass the
    // caches (no read-for-ownership of lines that are only ever written by the CPU and read by the DMA engine); CPU READS of
    // such memory are very slow, so it is only for drivers that never read the staging back (profiles/lab/h2d_lab: the
    // copy itself runs at the same 55.5 GB/s from either kind).
    {
        const char* sk = getenv("LFD_STAGING");
        if (sk && sk[0] == 'w') CK(cudaHostAlloc((void**)&h->frames_h, (size_t)B * N * sizeof(float), cudaHostAllocWriteCombined));
        else CK(cudaMallocHost((void**)&h->frames_h, (size_t)B * N * sizeof(float)));
    }
    CK(cudaMallocHost((void**)&h->res_h, (size_t)B * sizeof(lfd_result)));
    CK(cudaMallocHost((void**)&h->rects_h, (size_t)B * h->cfg.max_star_rects * sizeof(int4)));
    CK(cudaMallocHost((void**)&h->rect_off_h, ((size_t)B + 1) * sizeof(int)));
    CK(cudaMallocHost((void**)&h->ctl_h, (size_t)2 * B * sizeof(FrameCtl)));
    CK(cudaMemset(h->counters_d, 0, 16 * sizeof(int64_t)));
    CK(cudaMallocHost((void**)&h->counters_h, 16 * sizeof(int64_t)));
    memset(h->counters_h, 0, 16 * sizeof(int64_t));
    memset(h->timings, 0, sizeof(h->timings));
    return LFD_OK;
}

extern "C" int lfd_create_ex(int device, int max_batch, int height, int width, const lfd_config* cfg, lfd_handle** out)
{
    if (!out) return LFD_E_ARG;
    *out = nullptr;
    lfd_handle* h = new lfd_handle();
    memset(h->ev_valid, 0, sizeof(h->ev_valid));
    memset(h->mark_valid, 0, sizeof(h->mark_valid));
    int rc = create_impl(h, device, max_batch, height, width, cfg);
    if (rc != LFD_OK) {
        g_create_error = h->err;
        lfd_destroy(h);
        return rc;
    }
    *out = h;
    return LFD_OK;
}

extern "C" int lfd_create(int device, int max_batch, int height, int width, lfd_handle** out)
{
    return lfd_create_ex(device, max_batch, height, width, nullptr, out);
}

extern "C" int lfd_host_frames(lfd_handle* h, float** out)
{
    if (!h || !out) return LFD_E_ARG;
    *out = h->frames_h;
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------------
static int check_pass_params(lfd_handle* h, const lfd_pass_params& p, bool dim)
{
    if (p.nlinesInSet < 1 || p.nlinesInSet > LFD_MAX_SET_LINES) { h->err = "nlinesInSet must be in 1..16"; return LFD_E_UNSUPPORTED; }
    if (p.contoursMode < 0 || p.contoursMode > 3) { h->err = "unknown contoursMode"; return LFD_E_ARG; }
    if (p.contoursMethod != 1 && p.contoursMethod != 2) { h->err = "contoursMethod CHAIN_APPROX_TC89_* is not implemented"; return LFD_E_UNSUPPORTED; }
    if (p.dilate_h < 1 || p.dilate_w < 1) { h->err = "dilateKernel missing"; return LFD_E_ARG; }
    int eh = dim ? p.erode_h : 0, ew = dim ? p.erode_w : 0;
    if ((eh > 0) != (ew > 0)) { h->err = "bad erodeKernel"; return LFD_E_ARG; }
    // combined halo must fit the shared-memory tile padding of k_morph
    int e_t = eh / 2, e_b = eh > 0 ? eh - 1 - eh / 2 : 0, e_l = ew / 2, e_r = ew > 0 ? ew - 1 - ew / 2 : 0;
    int d_t = p.dilate_h / 2, d_b = p.dilate_h - 1 - p.dilate_h / 2, d_l = p.dilate_w / 2, d_r = p.dilate_w - 1 - p.dilate_w / 2;
    if (e_t + d_t > MORPH_PADH || e_b + d_b > MORPH_PADH || e_l + d_l > 12 || e_r + d_r > 11) {
        h->err = "erode+dilate kernel reach exceeds this build's tile halo (rows <= 16, cols <= 12 left / 11 right)";
        return LFD_E_UNSUPPORTED;
    }
    return LFD_OK;
}

// Kernel shapes the fused kernel is instantiated for: {erode h, erode w, dilate h, dilate w}
static const int k_fused_shapes[][4] = {{0, 0, 4, 4},       // params_bright default (detecttrails.py:204)
                                         {3, 3, 9, 9},       // params_dim default (detecttrails.py:220-221)
                                         {3, 3, 15, 15},     // high-sensitivity dim (BASELINE.json config 4)
                                         {0, 0, 9, 9},
                                         {0, 0, 3, 3}};
#define FUSED_NCASES 5
#define FUSED_DISPATCH(CASE, MACRO)                                                                \
    switch (CASE) {                                                                                 \
    case 0: MACRO(0, 0, 4, 4) break;                                                                \
    case 1: MACRO(3, 3, 9, 9) break;                                                                \
    case 2: MACRO(3, 3, 15, 15) break;                                                              \
    case 3: MACRO(0, 0, 9, 9) break;                                                                \
    case 4: MACRO(0, 0, 3, 3) break;                                                                \
    default: break;                                                                                 \
    }

// Tensor map of gray[pass] ([B][H][W] uint8, box = 256 x U x 1) and the dynamic shared-memory opt-in of the kernel
// instantiation that matches this pass's structuring elements; fused_case[pass] = -1 when there is none.
static int fused_setup(lfd_handle* h, int pass, int eh, int ew, int dh, int dw)
{
    h->fused_case[pass] = -1;
    if (!h->fused_ok) return LFD_OK;
    int cs = -1;
    for (int i = 0; i < FUSED_NCASES; i++)
        if (k_fused_shapes[i][0] == eh && k_fused_shapes[i][1] == ew && k_fused_shapes[i][2] == dh && k_fused_shapes[i][3] == dw) cs = i;
    if (cs < 0) return LFD_OK;
    int U = 0, smem = 0;
    cudaError_t ae = cudaSuccess;
#define FZ_ATTR(EH_, EW_, DH_, DW_)                                                                                         \
    U = FzGeom<EH_, DH_>::U; smem = FzGeom<EH_, DH_>::SMEM_B;                                                                \
    ae = cudaFuncSetAttribute(k_morph_nms<EH_, EW_, DH_, DW_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   \
    if (ae == cudaSuccess) ae = cudaFuncSetAttribute(k_morph_nms<EH_, EW_, DH_, DW_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    FUSED_DISPATCH(cs, FZ_ATTR)
#undef FZ_ATTR
    if (ae != cudaSuccess) { h->err = std::string("cudaFuncSetAttribute(k_morph_nms): ") + cudaGetErrorString(ae); return LFD_E_CUDA; }
    if (h->d.H < U) return LFD_OK;
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn) { cudaGetLastError(); return LFD_OK; }
    const cuuint64_t gdim[3] = {(cuuint64_t)h->d.W, (cuuint64_t)h->d.H, (cuuint64_t)h->B};
    const cuuint64_t gstr[2] = {(cuuint64_t)h->d.W, (cuuint64_t)h->d.W * (cuuint64_t)h->d.H};      // bytes, dimensions 1 and 2
    const cuuint32_t box[3] = {FZ_BOXW, (cuuint32_t)U, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = ((EncodeFn)fn)(&h->tm_gray[pass], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, h->gray[pass], gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr == CUDA_SUCCESS) h->fused_case[pass] = cs;
    return LFD_OK;
}

extern "C" int lfd_set_params(lfd_handle* h, const lfd_params* p)
{
    if (!h || !p) return LFD_E_ARG;
    cudaSetDevice(h->device);
    int rc;
    if ((rc = check_pass_params(h, p->bright, false)) != LFD_OK) return rc;
    if ((rc = check_pass_params(h, p->dim, true)) != LFD_OK) return rc;
    CK(cudaStreamSynchronize(h->stream));
    for (int pass = 0; pass < 2; pass++) {
        const lfd_pass_params& pp = pass ? p->dim : p->bright;
        // theta = np.pi/180 and threshold = 1 are literals in the reference (processfield.py:370, :488)
        rc = hough_setup(h, &h->hb[pass], h->d.H, h->d.W, pp.houghMethod, M_PI / 180, 1, h->B, false, h->cfg.max_lines);
        if (rc != LFD_OK) return rc;
        rc = fused_setup(h, pass, pass ? pp.erode_h : 0, pass ? pp.erode_w : 0, pp.dilate_h, pp.dilate_w);
        if (rc != LFD_OK) return rc;
    }
    h->params = *p;
    h->have_params = true;
    h->anyk_on[0] = h->anyk_on[1] = false;       // back to all-ones rectangles until lfd_set_kernels says otherwise
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }   // kernel arguments changed
    return LFD_OK;
}

static int fill_anykernel(lfd_handle* h, AnyKernel* k, const uint8_t* mask, int kh, int kw, const char* what)
{
    memset(k, 0, sizeof(*k));
    if (!mask || kh == 0 || kw == 0) return LFD_OK;
    if (kh < 1 || kw < 1 || kh > ANYK_MAX || kw > ANYK_MAX) { h->err = std::string(what) + ": kernel side must be 1..31"; return LFD_E_UNSUPPORTED; }
    const int ay = kh / 2, ax = kw / 2;        // cv2's default anchor
    for (int i = 0; i < kh; i++)
        for (int j = 0; j < kw; j++)
            if (mask[i * kw + j]) {
                k->dy[k->n] = (signed char)(i - ay); k->dx[k->n] = (signed char)(j - ax);
                k->reach = std::max(k->reach, std::max(abs(i - ay), abs(j - ax)));
                k->n++;
            }
    if (k->n == 0) { h->err = std::string(what) + ": structuring element has no non-zero entry"; return LFD_E_ARG; }
    return LFD_OK;
}

// Arbitrary structuring elements for one pass (row-major uint8 masks, non-zero = member, anchor at the centre like
// cv2.erode / cv2.dilate called with the default anchor: processfield.py:354, :464, :471).  erode_mask may be NULL.
extern "C" int lfd_set_kernels(lfd_handle* h, int pass, const uint8_t* erode_mask, int eh, int ew,
                               const uint8_t* dilate_mask, int dh, int dw)
{
    if (!h || (pass != 0 && pass != 1) || !dilate_mask) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (!h->have_params) { h->err = "lfd_set_params has not been called"; return LFD_E_STATE; }
    CK(cudaStreamSynchronize(h->stream));
    int rc;
    if ((rc = fill_anykernel(h, &h->anyk_h[pass][0], pass == 1 ? erode_mask : nullptr, eh, ew, "erodeKernel")) != LFD_OK) return rc;
    if ((rc = fill_anykernel(h, &h->anyk_h[pass][1], dilate_mask, dh, dw, "dilateKernel")) != LFD_OK) return rc;
    CK(cudaMemcpy(h->anyk_d + pass * 2, &h->anyk_h[pass][0], 2 * sizeof(AnyKernel), cudaMemcpyHostToDevice));
    h->anyk_on[pass] = true;
    if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
    size_t need = anyk_smem(h->anyk_h[pass][0].reach, h->anyk_h[pass][1].reach);
    DevLimits& L = g_dev_limits[h->device % LFD_MAX_DEVICES];
    if (need > L.anyk_max) { CK(cudaFuncSetAttribute(k_morph_any, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need)); L.anyk_max = need; }
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// the pipeline
// ------------------------------------------------------------------------------------------------
// frames [f0, f0 + n) of the batch: every per-frame array is entered at frame f0, kernels index frames from 0
// `part`: 0 = the whole pass ; 1 = only the contour part (NMS .. box fill) on a morph plane + nz mask the caller put in place
static int run_pass_kernels(lfd_handle* h, int f0, int n, int pass, int flags, cudaStream_t s, bool stage_events, int part = 0)
{
#define STAGE_EVENT(i) do { if (stage_events) CK(cudaEventRecord(h->ev[i], s)); } while (0)
#define KB_MARK(id, e) do {                                                                          \
        if (!(flags & LFD_KERNEL_TIMES)) break;                                                        \
        cudaEvent_t& kev_ = h->kb[h->cur_part][pass][id][e];                                           \
        /* under stream capture a plain record is a dependency marker, not a node: ask for an external record node */ \
        CK(cudaEventRecordWithFlags(kev_, s, capturing ? cudaEventRecordExternal : cudaEventRecordDefault)); \
        h->kb_used[h->cur_part][pass][id] = true;                                                      \
    } while (0)
    const Dims d = h->d;
    bool capturing = false;
    { cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone; if (cudaStreamIsCapturing(s, &cst) == cudaSuccess) capturing = cst == cudaStreamCaptureStatusActive; }
    const lfd_pass_params& pp = pass ? h->params.dim : h->params.bright;
    FrameCtl* const C = h->ctl + (size_t)pass * h->B + f0;      // this pass's bookkeeping block
    const size_t fN = (size_t)f0 * d.N, fNW = (size_t)f0 * d.NW;
    u8* const v_gray = h->gray[pass] + fN;
    u8* const v_morph = h->morph[pass] + fN;
    u32* const v_nz = h->nz[pass] + fNW;
    u32* const v_cand = h->cand[pass] + fNW;
    u32* const v_strong = h->strong[pass] + fNW;
    u32* const v_edges = h->edges[pass] + fNW;
    u32* const v_box = h->box[pass] + fNW;
    CclBuf* const v_ccl0 = h->ccl_d[pass][0] + f0;
    CclBuf* const v_ccl1 = h->ccl_d[pass][1] + f0;
    CompBuf* const v_comp = h->comp_d[pass] + f0;
    RectBuf* const v_rbuf = h->rbuf_d[pass] + f0;
    uint2* const v_segs = h->segs[pass] + (size_t)f0 * 2 * d.NW;
    lfd_result* const v_res = h->res_d + f0;
    const bool taps = flags & LFD_KEEP_TAPS;
    const int tbase = 2 + pass * (T_PER_PASS - 1);   // event index preceding this pass's first stage
    HoughBufs& hb = h->hb[pass];
    dim3 rows((d.H + CCL_WARPS * CCL_RC_ROWS - 1) / (CCL_WARPS * CCL_RC_ROWS), n);
    const int nbands = (d.H + CCL_BAND - 1) / CCL_BAND;
    dim3 bands((n >= 16 && nbands > CCL_BAND_CTAS) ? CCL_BAND_CTAS : nbands, n), seams((nbands + CCL_WARPS - 1) / CCL_WARPS, n);

    // NMS tap (class per pixel) only with LFD_KEEP_TAPS
    u8* ntap = nullptr;
    if (taps) {
        if (!h->nms_tap[pass]) { int rc = dev_alloc(h, &h->nms_tap[pass], (size_t)h->B * d.N); if (rc) return rc; }
        ntap = h->nms_tap[pass] + fN;
    }
    bool fused = false;        // morphology + Sobel + NMS done by one k_morph_nms launch
    if (part == 0) {
    // LUT + morphology
    k_lut<<<dim3(n, 1), 256, 0, s>>>(h->hist + (size_t)f0 * 256, h->lut + (size_t)f0 * 256, C, h->B, d.N, pass); LAUNCH_CHECK();
    MorphCfg mc;
    mc.eh = pass ? pp.erode_h : 0; mc.ew = pass ? pp.erode_w : 0; mc.dh = pp.dilate_h; mc.dw = pp.dilate_w;
    u8* etap = nullptr;
    if (taps && pass == 1 && mc.eh > 0) {
        if (!h->eroded_tap) { int rc = dev_alloc(h, &h->eroded_tap, (size_t)h->B * d.N); if (rc) return rc; }
        etap = h->eroded_tap + fN;
    }
    KB_MARK(KB_MORPH, 0);
    {
        const u8* lutp = h->lut + ((size_t)pass * h->B + f0) * 256;
        const int nstrips = ((d.W >> 2) + MARCH_UW - 1) / MARCH_UW, nchunks = (d.H + MARCH_R - 1) / MARCH_R;
        const int nunits = nstrips * nchunks;
        dim3 gg((nunits + MARCH_WPC - 1) / MARCH_WPC, n);
        bool done = false;
        if (h->anyk_on[pass]) {
            const AnyKernel* ekd = (pass == 1 && h->anyk_h[pass][0].n > 0) ? h->anyk_d + pass * 2 : nullptr;
            const AnyKernel* dkd = h->anyk_d + pass * 2 + 1;
            const int re = ekd ? h->anyk_h[pass][0].reach : 0, rd = h->anyk_h[pass][1].reach;
            dim3 ag((d.W + ANYK_TW - 1) / ANYK_TW, (d.H + ANYK_TH - 1) / ANYK_TH, n);
            k_morph_any<<<ag, 256, anyk_smem(re, rd), s>>>(v_gray, lutp, v_morph, v_nz, etap, C, pass, d, ekd, dkd);
            done = true;
        }
        // the all-ones rectangles of configs 1-4: one fused launch, gray rows staged by TMA (k_mnms.cuh); the morph plane is
        // written only as a stage tap
        if (!done && h->fused_case[pass] >= 0) {
            const int fchunks = (d.H + FZ_R - 1) / FZ_R, funits = nstrips * fchunks;
            dim3 fg((funits + FZ_WARPS - 1) / FZ_WARPS, n);
#define FUSED_LAUNCH(EH_, EW_, DH_, DW_)                                                                              \
            if (taps) k_morph_nms<EH_, EW_, DH_, DW_, true><<<fg, FZ_WARPS * 32, FzGeom<EH_, DH_>::SMEM_B, s>>>(h->tm_gray[pass], f0, \
                                          lutp, v_morph, v_nz, etap, v_cand, v_strong, ntap, C, pass, d, nstrips, funits, 0, 255); \
            else k_morph_nms<EH_, EW_, DH_, DW_, false><<<fg, FZ_WARPS * 32, FzGeom<EH_, DH_>::SMEM_B, s>>>(h->tm_gray[pass], f0, \
                                          lutp, nullptr, v_nz, nullptr, v_cand, v_strong, nullptr, C, pass, d, nstrips, funits, 0, 255);
            FUSED_DISPATCH(h->fused_case[pass], FUSED_LAUNCH)
#undef FUSED_LAUNCH
            done = true; fused = true;
        }
#define MORPH_CASE(EH_, EW_, DH_, DW_)                                                                              \
        if (!done && (d.W % 8) == 0 && mc.eh == EH_ && mc.ew == EW_ && mc.dh == DH_ && mc.dw == DW_) {                \
            k_morph_march<EH_, EW_, DH_, DW_><<<gg, MARCH_WPC * 32, 0, s>>>(v_gray, lutp, v_morph, v_nz, etap, \
                                                               C, pass, d, nstrips, nunits);                    \
            done = true;                                                                                             \
        }
        MORPH_CASE(0, 0, 4, 4)
        MORPH_CASE(3, 3, 9, 9)
        MORPH_CASE(3, 3, 15, 15)
        MORPH_CASE(0, 0, 9, 9)
        MORPH_CASE(0, 0, 3, 3)
#undef MORPH_CASE
        if (!done) {                // any other all-ones rectangle: shared-memory tile kernel
            dim3 mg((d.W + MORPH_TW - 1) / MORPH_TW, (d.H + MORPH_TH - 1) / MORPH_TH, n);
            k_morph<<<mg, 256, 0, s>>>(v_gray, lutp, v_morph, v_nz, etap, C, pass, d, mc);
        }
        LAUNCH_CHECK();
    }
    KB_MARK(KB_MORPH, 1);
    STAGE_EVENT(tbase + 1);
    }
    // Sobel + NMS (done by the fused kernel when that is selected)
    if (!fused) {
    if ((d.W % 8) == 0) {        // k_nms_march stores only non-zero mask words
        CK(cudaMemsetAsync(v_cand, 0, (size_t)n * d.NW * sizeof(u32), s));
        CK(cudaMemsetAsync(v_strong, 0, (size_t)n * d.NW * sizeof(u32), s));
    }
    KB_MARK(KB_NMS, 0);
    if ((d.W % 8) == 0) {
        const int nstrips = ((d.W >> 2) + MARCH_UW - 1) / MARCH_UW, nchunks = (d.H + NMS_R - 1) / NMS_R;
        const int nunits = nstrips * nchunks;
        dim3 gg((nunits + NMS_WPC - 1) / NMS_WPC, n);
        if (ntap) k_nms_march<true><<<gg, NMS_WPC * 32, 0, s>>>(v_morph, v_nz, v_cand, v_strong, ntap, C, pass, d, nstrips, nunits, 0, 255);
        else k_nms_march<false><<<gg, NMS_WPC * 32, 0, s>>>(v_morph, v_nz, v_cand, v_strong, ntap, C, pass, d, nstrips, nunits, 0, 255);
    } else {
        dim3 cg((d.W + CANNY_TW - 1) / CANNY_TW, (d.H + CANNY_TH - 1) / CANNY_TH, n);
        k_canny_nms<<<cg, 256, 0, s>>>(v_morph, v_nz, v_cand, v_strong, ntap, C, pass, d, 0, 255);
    }
    LAUNCH_CHECK();
    KB_MARK(KB_NMS, 1);
    }
    STAGE_EVENT(tbase + 2);
    // foreground runs: hysteresis + outer contours
    k_ccl_rowcount<<<rows, CCL_WARPS * 32, 0, s>>>(v_cand, v_ccl0, C, pass, d, 0); LAUNCH_CHECK();
    k_ccl_rowscan<<<n, 1024, 0, s>>>(v_ccl0, C, pass, d, 0, h->cfg.max_runs); LAUNCH_CHECK();
    KB_MARK(KB_CCL_BAND_FG, 0);
    k_ccl_band<<<bands, 256, ccl_band_smem(d.WW), s>>>(v_cand, v_ccl0, C, pass, d, 0); LAUNCH_CHECK();
    KB_MARK(KB_CCL_BAND_FG, 1);
    k_ccl_merge<<<seams, CCL_WARPS * 32, 0, s>>>(v_cand, v_ccl0, C, pass, d, 0); LAUNCH_CHECK();
    const dim3 flat(CCL_FLAT_CTAS, n);
    CK(cudaMemsetAsync(v_edges, 0, (size_t)n * d.NW * sizeof(u32), s));
    k_ccl_stats_flat<<<flat, 256, 0, s>>>(v_strong, v_ccl0, C, pass, d, 0); LAUNCH_CHECK();
    k_ccl_alloc_flat<<<flat, 256, 0, s>>>(v_ccl0, v_edges, v_comp, C, pass, d, 0); LAUNCH_CHECK();
    k_ccl_extremes_flat<<<flat, 256, 0, s>>>(v_edges, v_ccl0, v_comp, C, pass, d, 0); LAUNCH_CHECK();
    STAGE_EVENT(tbase + 3);
    // background runs: hole contours
    k_ccl_rowcount<<<rows, CCL_WARPS * 32, 0, s>>>(v_edges, v_ccl1, C, pass, d, 1); LAUNCH_CHECK();
    k_ccl_rowscan<<<n, 1024, 0, s>>>(v_ccl1, C, pass, d, 1, h->cfg.max_runs); LAUNCH_CHECK();
    KB_MARK(KB_CCL_BAND_BG, 0);
    k_ccl_band<<<bands, 256, ccl_band_smem(d.WW), s>>>(v_edges, v_ccl1, C, pass, d, 1); LAUNCH_CHECK();
    KB_MARK(KB_CCL_BAND_BG, 1);
    k_ccl_merge<<<seams, CCL_WARPS * 32, 0, s>>>(v_edges, v_ccl1, C, pass, d, 1); LAUNCH_CHECK();
    k_ccl_stats_flat<<<flat, 256, 0, s>>>(nullptr, v_ccl1, C, pass, d, 1); LAUNCH_CHECK();
    k_ccl_alloc_flat<<<flat, 256, 0, s>>>(v_ccl1, nullptr, v_comp, C, pass, d, 1); LAUNCH_CHECK();
    k_ccl_extremes_flat<<<flat, 256, 0, s>>>(v_edges, v_ccl1, v_comp, C, pass, d, 1); LAUNCH_CHECK();
    STAGE_EVENT(tbase + 4);
    // rectangles + box image
    KB_MARK(KB_RECTS, 0);
    k_rects_warp<<<h->sm_count * 16, RECT_WARPS * 32, rects_smem(n), s>>>(v_comp, v_rbuf, v_ccl0, v_ccl1, C, pass, n, d, pp.minAreaRectMinLen, pp.lwTresh, (unsigned long long*)h->counters_d, v_edges, pp.contoursMode == 0 ? 1 : 0); LAUNCH_CHECK();
    KB_MARK(KB_RECTS, 1);
    CK(cudaMemsetAsync(v_box, 0, (size_t)n * d.NW * sizeof(u32), s));
    k_fill_boxes<<<dim3(16, n), 128, 0, s>>>(v_rbuf, v_box, C, pass, d); LAUNCH_CHECK();
    STAGE_EVENT(tbase + 5);
    if (part == 1) return LFD_OK;
    // Hough on the morphology output and on the box image
    int* const v_accum = hb.accum + (size_t)f0 * 2 * hb.accum_stride;
    u64* const v_keys = hb.keys + (size_t)f0 * 2 * hb.key_stride;
    CK(cudaMemsetAsync(v_accum, 0, (size_t)n * 2 * hb.accum_stride * sizeof(int), s));
    // a thread's iterations are a serial chain (load -> ballot -> slot atomic -> store): many CTAs, few iterations each
    k_hough_compact<<<dim3(128, n, 2), 256, 0, s>>>(v_nz, v_box, v_segs, C, pass, d, (size_t)d.NW); LAUNCH_CHECK();
    KB_MARK(KB_HOUGH_VOTE, 0);
    if (hough_magic_ok(d.H, d.W, hb.hc.rho))
        k_hough_vote<true><<<dim3(32, hb.hc.ngroups, 2 * n), HOUGH_THREADS, hb.smem, s>>>(v_segs, v_accum, hb.tabSin, hb.tabCos, C, pass,
                                                                                  hb.hc, (size_t)d.NW, hb.accum_stride, (unsigned long long*)h->counters_d);
    else
        k_hough_vote<false><<<dim3(32, hb.hc.ngroups, 2 * n), HOUGH_THREADS, hb.smem, s>>>(v_segs, v_accum, hb.tabSin, hb.tabCos, C, pass,
                                                                                   hb.hc, (size_t)d.NW, hb.accum_stride, (unsigned long long*)h->counters_d);
    LAUNCH_CHECK();
    KB_MARK(KB_HOUGH_VOTE, 1);
    int cells = hb.hc.numangle * hb.hc.numrho;
    int pblocks = (cells + 255) / 256; if (pblocks > 64) pblocks = 64;
    k_hough_peaks<<<dim3(pblocks, 2 * n), 256, 0, s>>>(v_accum, v_keys, C, pass, hb.hc, hb.accum_stride, hb.key_stride); LAUNCH_CHECK();
    k_hough_topk<<<dim3(2, n), 256, 0, s>>>(v_keys, v_res, C, pass, hb.hc, hb.key_stride, pp.nlinesInSet); LAUNCH_CHECK();
    if (flags & LFD_FULL_LINES) {
        if (!hb.lines) CK(cudaMalloc((void**)&hb.lines, (size_t)h->B * 2 * hb.line_stride * sizeof(float)));
        k_hough_sort<<<dim3(2, n), 1024, 0, s>>>(v_keys, hb.lines + (size_t)f0 * 2 * hb.line_stride, C, pass, hb.hc, hb.key_stride, hb.line_stride, hb.max_lines); LAUNCH_CHECK();
    }
    STAGE_EVENT(tbase + 6);
    k_check_theta<<<(n + 63) / 64, 64, 0, s>>>(v_res, C, pass, n, pp.nlinesInSet, pp.dro, pp.thetaTresh, pp.lineSetTresh,
                                              hb.hc.numangle, (unsigned long long*)h->counters_d); LAUNCH_CHECK();
    STAGE_EVENT(tbase + 7);
    return LFD_OK;
#undef STAGE_EVENT
#undef KB_MARK
}

// mode 0: whole-frame pipeline on h->in (un-flipped) ; mode 1/2: standalone bright/dim on frame slot 0
static int run_pipeline(lfd_handle* h, int n, int flags, int mode, bool want_clipped)
{
    const Dims d = h->d;
    cudaStream_t s = h->stream;
    if (!h->have_params) { h->err = "lfd_set_params has not been called"; return LFD_E_STATE; }
    CK(cudaMemsetAsync(h->counters_d, 0, 16 * sizeof(int64_t), s));
    CK(cudaEventRecord(h->ev[0], s));
    const bool replay = h->graph_exec && h->use_graphs && h->graph_n == n && h->graph_flags == flags && mode == 0 &&
                        !(flags & (LFD_SERIAL_PASSES | LFD_KEEP_TAPS | LFD_FULL_LINES)) && !h->ktiming;
    if (!replay) memset(h->kb_used, 0, sizeof(h->kb_used));       // a replayed graph records the brackets of its capture again
    if (h->ktiming) { h->kn = 0; ktime_mark(h, 0); }
    k_ctl_init<<<(n + 127) / 128, 128, 0, s>>>(h->ctl, h->B, h->res_d, n, mode != 2, mode != 1); LAUNCH_CHECK();
    CK(cudaMemsetAsync(h->hist, 0, (size_t)2 * h->B * 256 * sizeof(u32), s));
    if (mode == 0) {
        CK(cudaMemsetAsync(h->starmask, 0, (size_t)n * d.NW * sizeof(u32), s));
        const int total_rects = h->rect_off_h[n];
        if (total_rects > 0) { k_star_mask<<<(total_rects + 3) / 4, 128, 0, s>>>(h->rects_d, h->rect_off_d, n, total_rects, h->starmask, d); LAUNCH_CHECK(); }
    }
    float* clipped = nullptr;
    if (want_clipped) {
        if (!h->clipped) { int rc = dev_alloc(h, &h->clipped, (size_t)d.N); if (rc) return rc; }
        clipped = h->clipped;
    }
    const lfd_pass_params& pd = h->params.dim;
    CK(cudaEventRecord(h->ev[1], s));
    const int be = (flags & LFD_INPUT_BIGENDIAN) ? 1 : 0;
    u32* hist1 = h->hist + (size_t)h->B * 256;
    if (clipped) {
        int pblocks = (d.N / 4 + 255) / 256; if (pblocks > 1184) pblocks = 1184;
        k_prep_generic<<<dim3(pblocks, n), 256, 0, s>>>(h->in, h->starmask, h->gray[0], h->gray[1], h->hist, hist1, clipped, d, mode, be,
                                                      (float)pd.minFlux, (float)pd.addFlux); LAUNCH_CHECK();
    } else {
        // one wave of resident CTAs spread over the frames of the batch (a second, partial wave would double the time)
        const float mf = (float)pd.minFlux, af = (float)pd.addFlux;
        // at most 16 frames in flight at a time (fewer concurrent DRAM streams: measured best of 16 / 24 / 32 / all
        // with profiles/lab/prep_lab.cu); the CTAs of the single resident wave walk the rest of the batch
        const int gy = n < 16 ? n : 16;
        int rb = h->prep_grid[mode] / gy; if (rb < 1) rb = 1;
        const int chunks = d.H * ((d.W * 4 + PR_CB - 1) / PR_CB);
        if (rb > (chunks + PR_WARPS - 1) / PR_WARPS) rb = (chunks + PR_WARPS - 1) / PR_WARPS;
        dim3 pg(rb, gy);
#define PREP_LAUNCH(M, E) k_prep<M, E><<<pg, PR_WARPS * 32, PR_SMEM, s>>>(h->in, h->starmask, h->gray[0], h->gray[1], h->hist, hist1, d, n, mf, af)
        if (mode == 0) { if (be) PREP_LAUNCH(0, true); else PREP_LAUNCH(0, false); }
        else if (mode == 1) { if (be) PREP_LAUNCH(1, true); else PREP_LAUNCH(1, false); }
        else { if (be) PREP_LAUNCH(2, true); else PREP_LAUNCH(2, false); }
#undef PREP_LAUNCH
        LAUNCH_CHECK();
    }
    CK(cudaEventRecord(h->ev[2], s));
    int rc;
    // The bright and the dim pass of a batch are independent until the verdict (the reference runs dim only
    // when bright finds nothing, detecttrails.py:125-129; here dim runs for every frame on a second stream and
    // k_finalize drops it where bright detected).  Most kernels after the morphology are latency-bound, so the
    // two passes overlap almost completely.  LFD_SERIAL_PASSES keeps one stream (clean per-stage timings).
    const bool overlap = mode == 0 && !(flags & LFD_SERIAL_PASSES) && !h->ktiming;
    const bool graph = overlap && h->use_graphs && !(flags & (LFD_KEEP_TAPS | LFD_FULL_LINES));
    h->stage_timings_valid = !graph;
    // everything from the fork to the result copies; captured into a CUDA graph when `graph`
    // with overlap and a batch of >= 16 frames the batch is also cut in nsplit parts (default two halves): 2 x nsplit
    // independent chains (bright / dim x part) keep the SMs busy through the latency-bound CCL / geometry phases
    int nsplit = (overlap && n >= 16) ? h->nsplit_cfg : 1;
    if (nsplit > 1 && n / nsplit < 8) nsplit = n / 8 > 0 ? n / 8 : 1;
    auto passes = [&](bool stage_events) -> int {
        int rc2;
        if (overlap) {
            CK(cudaEventRecord(h->ev_fork, s));
            for (int k = 0; k < nsplit; k++)
                for (int p = 0; p < 2; p++)
                    if (k || p) CK(cudaStreamWaitEvent(h->xs[k][p], h->ev_fork, 0));
        }
        int fstart = 0;
        for (int k = 0; k < nsplit; k++) {
            const int cnt = (n - fstart + (nsplit - k) - 1) / (nsplit - k);     // first parts take the remainder
            cudaStream_t sb = overlap ? h->xs[k][0] : s, sd = overlap ? h->xs[k][1] : s;
            const bool ev = stage_events && k == 0;
            h->cur_part = k;
            if (mode != 2) { if ((rc2 = run_pass_kernels(h, fstart, cnt, 0, flags, sb, ev)) != LFD_OK) return rc2; }
            else if (k == 0) for (int i = 3; i <= T_PER_PASS + 1; i++) CK(cudaEventRecord(h->ev[i], s));
            if (mode != 1) { if ((rc2 = run_pass_kernels(h, fstart, cnt, 1, flags, sd, ev)) != LFD_OK) return rc2; }
            else if (k == 0) for (int i = T_PER_PASS + 2; i < N_TIMINGS; i++) CK(cudaEventRecord(h->ev[i], s));
            fstart += cnt;
        }
        if (overlap)
            for (int k = 0; k < nsplit; k++)
                for (int p = 0; p < 2; p++)
                    if (k || p) { CK(cudaEventRecord(h->xjoin[k][p], h->xs[k][p])); CK(cudaStreamWaitEvent(s, h->xjoin[k][p], 0)); }
        k_finalize<<<(n + 127) / 128, 128, 0, s>>>(h->ctl, h->B, h->res_d, n, mode, (unsigned long long*)h->counters_d); LAUNCH_CHECK();
        CK(cudaMemcpyAsync(h->res_h, h->res_d, (size_t)n * sizeof(lfd_result), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->ctl_h, h->ctl, (size_t)n * sizeof(FrameCtl), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->ctl_h + h->B, h->ctl + h->B, (size_t)n * sizeof(FrameCtl), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(h->counters_h, h->counters_d, 16 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        return LFD_OK;
    };
    if (graph) {
        if (!h->graph_exec || h->graph_n != n || h->graph_flags != flags) {
            if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
            const int64_t l0 = h->launches;
            CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            rc = passes(false);
            cudaGraph_t g = nullptr;
            cudaError_t ce = cudaStreamEndCapture(s, &g);
            if (rc != LFD_OK) { if (g) cudaGraphDestroy(g); cudaGetLastError(); return rc; }
            if (ce != cudaSuccess) { h->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return LFD_E_CUDA; }
            ce = cudaGraphInstantiate(&h->graph_exec, g, 0);
            cudaGraphDestroy(g);
            if (ce != cudaSuccess) { h->graph_exec = nullptr; h->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); return LFD_E_CUDA; }
            h->graph_launches = h->launches - l0;
            h->launches = l0;
            h->graph_n = n; h->graph_flags = flags;
        }
        CK(cudaGraphLaunch(h->graph_exec, s));
        h->launches += h->graph_launches;
    } else {
        if ((rc = passes(true)) != LFD_OK) return rc;
    }
    CK(cudaEventRecord(h->ev[N_TIMINGS], s));
    h->last_n = n; h->last_flags = flags; h->pending = true;
    return LFD_OK;
}

static void CUDART_CB gate_release_cb(void* p) { flock((int)(intptr_t)p, LOCK_UN); }

static int stage_rects(lfd_handle* h, int n, const int32_t* rects, const int32_t* rect_offsets)
{
    if (!rects || !rect_offsets) {
        for (int f = 0; f <= n; f++) h->rect_off_h[f] = 0;
    } else {
        if (rect_offsets[0] != 0) { h->err = "rect_offsets[0] must be 0"; return LFD_E_ARG; }
        for (int f = 0; f < n; f++) {
            int c = rect_offsets[f + 1] - rect_offsets[f];
            if (c < 0 || c > h->cfg.max_star_rects) { h->err = "too many star rectangles for a frame (lfd_config.max_star_rects)"; return LFD_E_CAPACITY; }
        }
        memcpy(h->rect_off_h, rect_offsets, ((size_t)n + 1) * sizeof(int));
        memcpy(h->rects_h, rects, (size_t)rect_offsets[n] * sizeof(int4));
        CK(cudaMemcpyAsync(h->rects_d, h->rects_h, (size_t)rect_offsets[n] * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaMemcpyAsync(h->rect_off_d, h->rect_off_h, ((size_t)n + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    return LFD_OK;
}

extern "C" int lfd_upload(lfd_handle* h, const float* frames, int n, const int32_t* rects, const int32_t* rect_offsets, int flags)
{
    if (!h || n < 1 || n > h->B) { if (h) h->err = "bad batch size"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "previous batch not collected (call lfd_wait)"; return LFD_E_STATE; }
    int rc = stage_rects(h, n, rects, rect_offsets);
    if (rc != LFD_OK) return rc;
    const float* src = frames ? frames : h->frames_h;
    // Copy gate: wait (bounded) for this rank's slot of its host bridge, copy, and let a host callback on the stream free
    // the slot when the copy has completed.  With four GPUs behind one ~117 GB/s bridge the four concurrent 781 MB copies
    // are served unfairly (21-37 GB/s per GPU, profiles/h2d_lab_g8.txt) and the slowest rank sets the pace of a job that
    // gives every rank the same work; two copies at a time run at ~55 GB/s each and every rank gets the same share.
    bool held = false;
    if (h->gate_fd >= 0) {
        struct timespec t0, t1;
        clock_gettime(CLOCK_MONOTONIC, &t0);
        for (;;) {
            if (flock(h->gate_fd, LOCK_EX | LOCK_NB) == 0) { held = true; break; }
            clock_gettime(CLOCK_MONOTONIC, &t1);
            if ((t1.tv_sec - t0.tv_sec) * 1000.0 + (t1.tv_nsec - t0.tv_nsec) / 1e6 > 250.0) { h->gate_timeouts++; break; }   // never deadlock on a peer
            usleep(40);
        }
        h->gate_waits++;
    }
    cudaError_t ce = cudaMemcpyAsync(h->in, src, (size_t)n * h->d.N * sizeof(float), cudaMemcpyHostToDevice, h->stream);
    if (held) {
        if (ce != cudaSuccess || cudaLaunchHostFunc(h->stream, gate_release_cb, (void*)(intptr_t)h->gate_fd) != cudaSuccess)
            flock(h->gate_fd, LOCK_UN);
    }
    if (ce != cudaSuccess) { h->err = std::string("cudaMemcpyAsync(H2D): ") + cudaGetErrorString(ce); return LFD_E_CUDA; }
    h->last_flags = flags;
    return LFD_OK;
}

// Share the host->device copy slot named by `lock_path` (an advisory lock file, created if missing) with the other
// processes that name the same file; NULL or "" removes the gate.  See lfd_upload.
extern "C" int lfd_set_h2d_gate(lfd_handle* h, const char* lock_path)
{
    if (!h) return LFD_E_ARG;
    if (h->gate_fd >= 0) { flock(h->gate_fd, LOCK_UN); close(h->gate_fd); h->gate_fd = -1; }
    if (!lock_path || !lock_path[0]) return LFD_OK;
    int fd = open(lock_path, O_RDWR | O_CREAT | O_CLOEXEC, 0666);
    if (fd < 0) { h->err = std::string("lfd_set_h2d_gate: cannot open ") + lock_path; return LFD_E_ARG; }
    h->gate_fd = fd;
    return LFD_OK;
}

extern "C" int lfd_submit(lfd_handle* h, const float* frames, int n, const int32_t* rects, const int32_t* rect_offsets, int flags)
{
    int rc = lfd_upload(h, frames, n, rects, rect_offsets, flags);
    if (rc != LFD_OK) return rc;
    return run_pipeline(h, n, flags, 0, false);
}

extern "C" int lfd_run_resident(lfd_handle* h, int n, int flags)
{
    if (!h || n < 1 || n > h->B) { if (h) h->err = "bad batch size"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "previous batch not collected (call lfd_wait)"; return LFD_E_STATE; }
    return run_pipeline(h, n, flags, 0, false);
}

extern "C" int lfd_wait(lfd_handle* h, lfd_result* out)
{
    if (!h) return LFD_E_ARG;
    cudaSetDevice(h->device);
    if (!h->pending) { h->err = "nothing submitted"; return LFD_E_STATE; }
    h->pending = false;
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < N_TIMINGS; i++) {
        float ms = 0.f;
        // in graph mode only the setup and k_prep brackets (recorded outside the graph) exist
        if (h->stage_timings_valid || i < 2)
            if (cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]) != cudaSuccess) { ms = 0.f; cudaGetLastError(); }
        h->timings[i] = ms;
    }
    memset(h->kb_ms, 0, sizeof(h->kb_ms));
    memset(h->kb_launches, 0, sizeof(h->kb_launches));
    for (int k = 0; k < LFD_MAX_SPLIT; k++)
        for (int p = 0; p < 2; p++)
            for (int b = 0; b < KB_COUNT; b++) {
                if (!h->kb_used[k][p][b]) continue;
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, h->kb[k][p][b][0], h->kb[k][p][b][1]) != cudaSuccess) { cudaGetLastError(); continue; }
                h->kb_ms[p][b] += ms;
                h->kb_launches[p][b]++;
            }
    if (out) memcpy(out, h->res_h, (size_t)h->last_n * sizeof(lfd_result));
    return LFD_OK;
}

// Bracketed kernels of the last collected batch: name, pass, summed duration over the batch parts and number of launches.
extern "C" int lfd_get_kernel_times(lfd_handle* h, int index, const char** name, int* pass, float* ms, int* launches)
{
    if (!h || index < 0 || index >= 2 * KB_COUNT) return LFD_E_ARG;
    const int p = index / KB_COUNT, b = index % KB_COUNT;
    if (name) *name = k_bracket_names[b];
    if (pass) *pass = p;
    if (ms) *ms = h->kb_ms[p][b];
    if (launches) *launches = h->kb_launches[p][b];
    return LFD_OK;
}

extern "C" int lfd_run_pass(lfd_handle* h, int pass, float* img, int flags, int writeback, lfd_result* out)
{
    if (!h || !img || (pass != 0 && pass != 1)) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "previous batch not collected (call lfd_wait)"; return LFD_E_STATE; }
    h->rect_off_h[0] = h->rect_off_h[1] = 0;
    CK(cudaMemcpyAsync(h->in, img, (size_t)h->d.N * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    int rc = run_pipeline(h, 1, flags, pass == 0 ? 1 : 2, writeback != 0);
    if (rc != LFD_OK) return rc;
    if (writeback) CK(cudaMemcpyAsync(img, h->clipped, (size_t)h->d.N * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    return lfd_wait(h, out);
}

// remove_stars' blot on a host image
__global__ void k_blot(float* img, const int4* rects, int nrects, Dims d)
{
    int ri = blockIdx.x;
    if (ri >= nrects) return;
    int4 r = rects[ri];
    int r0 = max(r.x, 0), r1 = min(r.y, d.H), c0 = max(r.z, 0), c1 = min(r.w, d.W);
    if (r0 >= r1 || c0 >= c1) return;
    int w = c1 - c0, total = (r1 - r0) * w;
    for (int i = threadIdx.x; i < total; i += blockDim.x) img[(size_t)(r0 + i / w) * d.W + c0 + i % w] = 0.0f;
}

extern "C" int lfd_blot(lfd_handle* h, float* img, const int32_t* rects, int nrects)
{
    if (!h || !img || nrects < 0 || (nrects > 0 && !rects)) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "previous batch not collected (call lfd_wait)"; return LFD_E_STATE; }
    if (nrects > h->B * h->cfg.max_star_rects) { h->err = "too many rectangles"; return LFD_E_CAPACITY; }
    if (nrects == 0) return LFD_OK;
    memcpy(h->rects_h, rects, (size_t)nrects * sizeof(int4));
    CK(cudaMemcpyAsync(h->rects_d, h->rects_h, (size_t)nrects * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->in, img, (size_t)h->d.N * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    k_blot<<<nrects, 128, 0, h->stream>>>(h->in, h->rects_d, nrects, h->d); LAUNCH_CHECK();
    CK(cudaMemcpyAsync(img, h->in, (size_t)h->d.N * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// taps
// ------------------------------------------------------------------------------------------------
extern "C" int lfd_get_stage_count(lfd_handle* h, int frame, int pass, int stage, int* count)
{
    if (!h || !count || frame < 0 || frame >= h->last_n || (pass != 0 && pass != 1)) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    const HoughBufs& hb = h->hb[pass];
    switch (stage) {
    case LFD_STAGE_RECTS: {
        const FrameCtl& ch = h->ctl_h[(size_t)pass * h->B + frame];
        *count = min(ch.ncomp_saved[pass][0], h->cfg.max_components) + min(ch.ncomp_saved[pass][1], h->cfg.max_components);
        return LFD_OK;
    }
    case LFD_STAGE_LINES_EQU: *count = h->res_h[frame].n_lines_equ[pass]; return LFD_OK;
    case LFD_STAGE_LINES_BOX: *count = h->res_h[frame].n_lines_box[pass]; return LFD_OK;
    case LFD_STAGE_ACCUM_EQU:
    case LFD_STAGE_ACCUM_BOX: *count = (int)hb.accum_stride; return LFD_OK;
    default: h->err = "stage has no count"; return LFD_E_ARG;
    }
}

extern "C" int lfd_get_stage(lfd_handle* h, int frame, int pass, int stage, void* host_out, size_t bytes)
{
    if (!h || !host_out || frame < 0 || frame >= h->last_n || (pass != 0 && pass != 1)) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "batch still pending (call lfd_wait)"; return LFD_E_STATE; }
    const Dims d = h->d;
    cudaStream_t s = h->stream;
    const size_t N = d.N;
    Dims d1 = d;
    const void* src = nullptr;
    size_t need = N;
    HoughBufs& hb = h->hb[pass];
    switch (stage) {
    case LFD_STAGE_MASK:
        k_expand_mask<<<dim3(256, 1), 256, 0, s>>>(h->starmask + (size_t)frame * d.NW, h->tap_u8, d1); LAUNCH_CHECK();
        src = h->tap_u8; break;
    case LFD_STAGE_GRAY: src = h->gray[pass] + (size_t)frame * N; break;
    case LFD_STAGE_EQU:
        k_apply_lut<<<dim3(256, 1), 256, 0, s>>>(h->gray[pass] + (size_t)frame * N, h->lut + ((size_t)pass * h->B + frame) * 256, h->tap_u8, d1); LAUNCH_CHECK();
        src = h->tap_u8; break;
    case LFD_STAGE_ERODED:
        if (!h->eroded_tap || pass != 1) { h->err = "eroded tap not recorded (dim pass with LFD_KEEP_TAPS)"; return LFD_E_STATE; }
        src = h->eroded_tap + (size_t)frame * N; break;
    case LFD_STAGE_MORPH:
        // the fused production kernel never writes the morphology plane; it exists when the taps were kept (or on the
        // unfused fallback path)
        if (h->fused_case[pass] >= 0 && !(h->last_flags & LFD_KEEP_TAPS) && !h->anyk_on[pass]) { h->err = "morph plane not recorded (LFD_KEEP_TAPS)"; return LFD_E_STATE; }
        src = h->morph[pass] + (size_t)frame * N; break;
    case LFD_STAGE_CANNY:
        k_expand_mask<<<dim3(256, 1), 256, 0, s>>>(h->edges[pass] + (size_t)frame * d.NW, h->tap_u8, d1); LAUNCH_CHECK();
        src = h->tap_u8; break;
    case LFD_STAGE_BOX:
        k_expand_mask<<<dim3(256, 1), 256, 0, s>>>(h->box[pass] + (size_t)frame * d.NW, h->tap_u8, d1); LAUNCH_CHECK();
        src = h->tap_u8; break;
    case LFD_STAGE_HIST: src = h->hist + ((size_t)pass * h->B + frame) * 256; need = 256 * sizeof(u32); break;
    case LFD_STAGE_LUT: src = h->lut + ((size_t)pass * h->B + frame) * 256; need = 256; break;
    case LFD_STAGE_NMS:
        if (!h->nms_tap[pass]) { h->err = "NMS tap not recorded (LFD_KEEP_TAPS)"; return LFD_E_STATE; }
        src = h->nms_tap[pass] + (size_t)frame * N; break;
    case LFD_STAGE_FG_LABELS:
    case LFD_STAGE_BG_LABELS: {
        if (!h->labels_tap) { int rc = dev_alloc(h, &h->labels_tap, N); if (rc) return rc; }
        int kind = stage == LFD_STAGE_BG_LABELS;
        k_ccl_labels<<<(d.H + CCL_WARPS - 1) / CCL_WARPS, CCL_WARPS * 32, 0, s>>>(h->ccl_d[pass][kind], h->labels_tap, h->ctl + (size_t)pass * h->B, pass, d, kind, frame); LAUNCH_CHECK();
        src = h->labels_tap; need = N * sizeof(int); break;
    }
    case LFD_STAGE_RECTS: {
        const FrameCtl& ch = h->ctl_h[(size_t)pass * h->B + frame];
        int n0 = min(ch.ncomp_saved[pass][0], h->cfg.max_components), n1 = min(ch.ncomp_saved[pass][1], h->cfg.max_components);
        need = (size_t)(n0 + n1) * sizeof(lfd_rect);
        if (bytes < need) { h->err = "buffer too small"; return LFD_E_ARG; }
        const RectBuf& rb = h->rbuf_h[pass][frame];
        if (n0) CK(cudaMemcpyAsync(host_out, rb.rects, (size_t)n0 * sizeof(lfd_rect), cudaMemcpyDeviceToHost, s));
        if (n1) CK(cudaMemcpyAsync((char*)host_out + (size_t)n0 * sizeof(lfd_rect), rb.rects + h->cfg.max_components, (size_t)n1 * sizeof(lfd_rect), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return LFD_OK;
    }
    case LFD_STAGE_ACCUM_EQU:
    case LFD_STAGE_ACCUM_BOX:
        src = hb.accum + ((size_t)frame * 2 + (stage == LFD_STAGE_ACCUM_BOX)) * hb.accum_stride;
        need = hb.accum_stride * sizeof(int); break;
    case LFD_STAGE_LINES_EQU:
    case LFD_STAGE_LINES_BOX: {
        if (!hb.lines || !(h->last_flags & LFD_FULL_LINES)) { h->err = "full line lists not recorded (LFD_FULL_LINES)"; return LFD_E_STATE; }
        int which = stage == LFD_STAGE_LINES_BOX;
        int nl = which ? h->res_h[frame].n_lines_box[pass] : h->res_h[frame].n_lines_equ[pass];
        if (nl < 0) nl = 0;
        if (nl > hb.max_lines) nl = hb.max_lines;
        src = hb.lines + ((size_t)frame * 2 + which) * hb.line_stride;
        need = (size_t)nl * 2 * sizeof(float);
        if (need == 0) return LFD_OK;
        break;
    }
    case LFD_STAGE_CLIPPED:
        if (!h->clipped) { h->err = "clipped image not recorded"; return LFD_E_STATE; }
        src = h->clipped; need = N * sizeof(float); break;
    default: h->err = "unknown stage"; return LFD_E_ARG;
    }
    if (bytes < need) { h->err = "buffer too small"; return LFD_E_ARG; }
    CK(cudaMemcpyAsync(host_out, src, need, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// standalone cv2.HoughLines
// ------------------------------------------------------------------------------------------------
struct HoughScratch {
    int H = 0, W = 0;
    u8* img = nullptr; u32* mask = nullptr; uint2* segs = nullptr; FrameCtl* ctl = nullptr; lfd_result* res = nullptr;
    HoughBufs hb;
};
static HoughScratch g_hs;   // one per process is enough for the microbenchmark entry point

extern "C" int lfd_hough_dims(int height, int width, double rho, double theta, int* numangle, int* numrho)
{
    if (!numangle || !numrho || !(rho > 0) || !(theta > 0)) return LFD_E_ARG;
    hough_dims_host(height, width, rho, theta, numangle, numrho);
    return LFD_OK;
}

extern "C" int lfd_hough_lines(lfd_handle* h, const uint8_t* img, int height, int width, double rho, double theta,
                               int threshold, float* lines, int max_lines, int* n_lines, int32_t* accum)
{
    if (!h || !img || height < 1 || width < 1 || height > 65535 || width > 65535 * 32) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "previous batch not collected (call lfd_wait)"; return LFD_E_STATE; }
    cudaStream_t s = h->stream;
    Dims d; d.H = height; d.W = width; d.WW = (width + 31) / 32; d.N = height * width; d.NW = height * d.WW;
    HoughScratch& hs = g_hs;
    if (hs.H != height || hs.W != width) {
        if (hs.img) { cudaFree(hs.img); cudaFree(hs.mask); cudaFree(hs.segs); cudaFree(hs.ctl); cudaFree(hs.res); }
        hs.img = nullptr;
        CK(cudaMalloc((void**)&hs.img, (size_t)d.N));
        CK(cudaMalloc((void**)&hs.mask, (size_t)d.NW * sizeof(u32)));
        CK(cudaMalloc((void**)&hs.segs, (size_t)2 * d.NW * sizeof(uint2)));
        CK(cudaMalloc((void**)&hs.ctl, 2 * sizeof(FrameCtl)));
        CK(cudaMalloc((void**)&hs.res, sizeof(lfd_result)));
        hs.H = height; hs.W = width;
    }
    int rc = hough_setup(h, &hs.hb, height, width, rho, theta, threshold, 1, true, 0);
    if (rc != LFD_OK) return rc;
    HoughBufs& hb = hs.hb;
    CK(cudaMemcpyAsync(hs.img, img, (size_t)d.N, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(h->counters_d, 0, 16 * sizeof(int64_t), s));
    k_ctl_init<<<1, 32, 0, s>>>(hs.ctl, 1, hs.res, 1, 1, 0); LAUNCH_CHECK();
    // mark Hough as enabled for pass 0
    FrameCtl c; memset(&c, 0, sizeof(c)); c.active[0] = 1; c.hough[0] = 1;
    CK(cudaMemcpyAsync(hs.ctl, &c, sizeof(c), cudaMemcpyHostToDevice, s));
    // device times of the four phases -> lfd_get_timings entries 0..3 (pack+compact, vote, peaks, sort)
    CK(cudaEventRecord(h->ev[0], s));
    k_pack_mask<<<592, 256, 0, s>>>(hs.img, hs.mask, d); LAUNCH_CHECK();
    CK(cudaMemsetAsync(hb.accum, 0, (size_t)2 * hb.accum_stride * sizeof(int), s));
    // which = 0 only: pass the same mask twice and ignore slot 1
    k_hough_compact<<<dim3(64, 1, 1), 256, 0, s>>>(hs.mask, hs.mask, hs.segs, hs.ctl, 0, d, (size_t)d.NW); LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev[1], s));
    if (hough_magic_ok(height, width, hb.hc.rho))
        k_hough_vote<true><<<dim3(32, hb.hc.ngroups, 1), HOUGH_THREADS, hb.smem, s>>>(hs.segs, hb.accum, hb.tabSin, hb.tabCos, hs.ctl, 0, hb.hc,
                                                                               (size_t)d.NW, hb.accum_stride, (unsigned long long*)h->counters_d);
    else
        k_hough_vote<false><<<dim3(32, hb.hc.ngroups, 1), HOUGH_THREADS, hb.smem, s>>>(hs.segs, hb.accum, hb.tabSin, hb.tabCos, hs.ctl, 0, hb.hc,
                                                                                (size_t)d.NW, hb.accum_stride, (unsigned long long*)h->counters_d);
    LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev[2], s));
    long long cells = (long long)hb.hc.numangle * hb.hc.numrho;
    int pblocks = (int)((cells + 255) / 256); if (pblocks > 1184) pblocks = 1184;
    k_hough_peaks<<<dim3(pblocks, 1), 256, 0, s>>>(hb.accum, hb.keys, hs.ctl, 0, hb.hc, hb.accum_stride, hb.key_stride); LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev[3], s));
    // the full sort is only needed when the caller wants the line list (n_lines and the accumulator do not depend on it)
    if (lines && max_lines > 0) { k_hough_sort<<<dim3(1, 1), 1024, 0, s>>>(hb.keys, hb.lines, hs.ctl, 0, hb.hc, hb.key_stride, hb.line_stride, hb.max_lines); LAUNCH_CHECK(); }
    CK(cudaEventRecord(h->ev[4], s));
    FrameCtl out;
    CK(cudaMemcpyAsync(&out, hs.ctl, sizeof(out), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(h->counters_h, h->counters_d, 16 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memset(h->timings, 0, sizeof(h->timings));
    for (int i = 0; i < 4; i++)
        if (cudaEventElapsedTime(&h->timings[i], h->ev[i], h->ev[i + 1]) != cudaSuccess) { h->timings[i] = 0.f; cudaGetLastError(); }
    h->counters_h[0] = out.nnz[0];
    h->counters_h[2] = (int64_t)out.nnz[0] * hb.hc.numangle;
    int np = out.npeaks[0];
    if (n_lines) *n_lines = np;
    int ncopy = np < max_lines ? np : max_lines;
    if (ncopy > hb.max_lines) ncopy = hb.max_lines;
    if (lines && ncopy > 0) CK(cudaMemcpy(lines, hb.lines, (size_t)ncopy * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    if (accum) CK(cudaMemcpy(accum, hb.accum, hb.accum_stride * sizeof(int), cudaMemcpyDeviceToHost));
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// standalone cv2.Canny (aperture 3, L1 gradient) on a host uint8 image of the handle's frame size
// ------------------------------------------------------------------------------------------------
extern "C" int lfd_canny(lfd_handle* h, const uint8_t* img, int low, int high, uint8_t* edges_out)
{
    if (!h || !img || !edges_out || low < 0 || high < low) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "previous batch not collected (call lfd_wait)"; return LFD_E_STATE; }
    const Dims d = h->d;
    cudaStream_t s = h->stream;
    const int pass = 0, n = 1;
    FrameCtl* const C = h->ctl;
    CK(cudaMemcpyAsync(h->morph[pass], img, (size_t)d.N, cudaMemcpyHostToDevice, s));
    k_ctl_init<<<1, 32, 0, s>>>(h->ctl, h->B, h->res_d, 1, 1, 0); LAUNCH_CHECK();
    // device times -> lfd_get_timings entries 0..1 (Sobel + NMS incl. the input mask, hysteresis by run CCL)
    CK(cudaEventRecord(h->ev[0], s));
    k_pack_mask<<<592, 256, 0, s>>>(h->morph[pass], h->nz[pass], d); LAUNCH_CHECK();
    if ((d.W % 8) == 0) {
        const int nstrips = ((d.W >> 2) + MARCH_UW - 1) / MARCH_UW, nchunks = (d.H + NMS_R - 1) / NMS_R;
        const int nunits = nstrips * nchunks;
        CK(cudaMemsetAsync(h->cand[pass], 0, (size_t)d.NW * sizeof(u32), s));
        CK(cudaMemsetAsync(h->strong[pass], 0, (size_t)d.NW * sizeof(u32), s));
        k_nms_march<false><<<dim3((nunits + NMS_WPC - 1) / NMS_WPC, n), NMS_WPC * 32, 0, s>>>(h->morph[pass], h->nz[pass], h->cand[pass], h->strong[pass], nullptr, C, pass, d,
                                                                  nstrips, nunits, low, high);
    } else {
        dim3 cg((d.W + CANNY_TW - 1) / CANNY_TW, (d.H + CANNY_TH - 1) / CANNY_TH, n);
        k_canny_nms<<<cg, 256, 0, s>>>(h->morph[pass], h->nz[pass], h->cand[pass], h->strong[pass], nullptr, C, pass, d, low, high);
    }
    LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev[1], s));
    dim3 rows((d.H + CCL_WARPS * CCL_RC_ROWS - 1) / (CCL_WARPS * CCL_RC_ROWS), n);
    const int nbands = (d.H + CCL_BAND - 1) / CCL_BAND;
    dim3 bands(nbands, n), seams((nbands + CCL_WARPS - 1) / CCL_WARPS, n);
    k_ccl_rowcount<<<rows, CCL_WARPS * 32, 0, s>>>(h->cand[pass], h->ccl_d[pass][0], C, pass, d, 0); LAUNCH_CHECK();
    k_ccl_rowscan<<<n, 1024, 0, s>>>(h->ccl_d[pass][0], C, pass, d, 0, h->cfg.max_runs); LAUNCH_CHECK();
    k_ccl_band<<<bands, 256, ccl_band_smem(d.WW), s>>>(h->cand[pass], h->ccl_d[pass][0], C, pass, d, 0); LAUNCH_CHECK();
    k_ccl_merge<<<seams, CCL_WARPS * 32, 0, s>>>(h->cand[pass], h->ccl_d[pass][0], C, pass, d, 0); LAUNCH_CHECK();
    CK(cudaMemsetAsync(h->edges[pass], 0, (size_t)d.NW * sizeof(u32), s));
    k_ccl_stats_flat<<<dim3(CCL_FLAT_CTAS, n), 256, 0, s>>>(h->strong[pass], h->ccl_d[pass][0], C, pass, d, 0); LAUNCH_CHECK();
    k_ccl_alloc_flat<<<dim3(CCL_FLAT_CTAS, n), 256, 0, s>>>(h->ccl_d[pass][0], h->edges[pass], h->comp_d[pass], C, pass, d, 0); LAUNCH_CHECK();
    CK(cudaEventRecord(h->ev[2], s));
    k_expand_mask<<<dim3(592, 1), 256, 0, s>>>(h->edges[pass], h->tap_u8, d); LAUNCH_CHECK();
    FrameCtl out;
    CK(cudaMemcpyAsync(&out, C, sizeof(out), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(edges_out, h->tap_u8, (size_t)d.N, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memset(h->timings, 0, sizeof(h->timings));
    for (int i = 0; i < 2; i++)
        if (cudaEventElapsedTime(&h->timings[i], h->ev[i], h->ev[i + 1]) != cudaSuccess) { h->timings[i] = 0.f; cudaGetLastError(); }
    if (out.status & LFD_FRAME_OVERFLOW) { h->err = "edge map has more runs than lfd_config.max_runs"; return LFD_E_CAPACITY; }
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// standalone fit_minAreaRect (processfield.py:201-263): Canny -> contours -> minAreaRect filter -> box image
// ------------------------------------------------------------------------------------------------
extern "C" int lfd_fit_min_area_rect(lfd_handle* h, const uint8_t* img, int contoursMode, int contoursMethod,
                                     double minAreaRectMinLen, double lwTresh, uint8_t* box_out, int* detection)
{
    if (!h || !img || !box_out || !detection) { if (h) h->err = "bad argument"; return LFD_E_ARG; }
    cudaSetDevice(h->device);
    if (h->pending) { h->err = "previous batch not collected (call lfd_wait)"; return LFD_E_STATE; }
    if (!h->have_params) { h->err = "lfd_set_params has not been called"; return LFD_E_STATE; }
    if (contoursMode < 0 || contoursMode > 3) { h->err = "unknown contoursMode"; return LFD_E_ARG; }
    if (contoursMethod != 1 && contoursMethod != 2) { h->err = "contoursMethod CHAIN_APPROX_TC89_* is not implemented"; return LFD_E_UNSUPPORTED; }
    const Dims d = h->d;
    cudaStream_t s = h->stream;
    const lfd_pass_params saved = h->params.bright;
    h->params.bright.contoursMode = contoursMode; h->params.bright.contoursMethod = contoursMethod;
    h->params.bright.minAreaRectMinLen = minAreaRectMinLen; h->params.bright.lwTresh = lwTresh;
    int rc = LFD_OK;
    do {
        cudaError_t ce;
        if ((ce = cudaMemcpyAsync(h->morph[0], img, (size_t)d.N, cudaMemcpyHostToDevice, s)) != cudaSuccess) { h->err = cudaGetErrorString(ce); rc = LFD_E_CUDA; break; }
        k_ctl_init<<<1, 32, 0, s>>>(h->ctl, h->B, h->res_d, 1, 1, 0);
        k_pack_mask<<<592, 256, 0, s>>>(h->morph[0], h->nz[0], d);
        h->launches += 2;
        h->cur_part = 0;
        if ((rc = run_pass_kernels(h, 0, 1, 0, 0, s, false, 1)) != LFD_OK) break;
        k_expand_mask<<<dim3(592, 1), 256, 0, s>>>(h->box[0], h->tap_u8, d);
        h->launches++;
        FrameCtl out;
        if ((ce = cudaMemcpyAsync(&out, h->ctl, sizeof(out), cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
            (ce = cudaMemcpyAsync(box_out, h->tap_u8, (size_t)d.N, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
            (ce = cudaStreamSynchronize(s)) != cudaSuccess) { h->err = cudaGetErrorString(ce); rc = LFD_E_CUDA; break; }
        if (out.status & LFD_FRAME_OVERFLOW) { h->err = "per-frame work list overflow (lfd_config.max_runs / max_components)"; rc = LFD_E_CAPACITY; break; }
        *detection = out.hough[0] ? 1 : 0;
    } while (0);
    h->params.bright = saved;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// micro-benchmark: peak shared-memory atomic rate (denominator for the Hough vote kernel)
// ------------------------------------------------------------------------------------------------
extern "C" int lfd_smem_atomic_peak(lfd_handle* h, double* gops)
{
    if (!h || !gops) return LFD_E_ARG;
    cudaSetDevice(h->device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    const int iters = 4096, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_smem_atomic_peak<<<blocks, 256, 0, h->stream>>>(64, (unsigned*)h->counters_d + 30); LAUNCH_CHECK();      // warm-up
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(a, h->stream));
        k_smem_atomic_peak<<<blocks, 256, 0, h->stream>>>(iters, (unsigned*)h->counters_d + 30); LAUNCH_CHECK();
        CK(cudaEventRecord(b, h->stream));
        CK(cudaEventSynchronize(b));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    *gops = (double)blocks * 256.0 * 8.0 * iters / (best * 1e6);
    return LFD_OK;
}

// ------------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------------
extern "C" int lfd_get_timings(lfd_handle* h, float* ms, int max_entries, int* n_entries)
{
    if (!h || !ms) return LFD_E_ARG;
    int n = N_TIMINGS < max_entries ? N_TIMINGS : max_entries;
    for (int i = 0; i < n; i++) ms[i] = h->timings[i];
    if (n_entries) *n_entries = n;
    return LFD_OK;
}

// caller-placed CUDA-event timestamps on the handle's stream (benchmark timing on the device)
extern "C" int lfd_timer_mark(lfd_handle* h, int slot)
{
    if (!h || slot < 0 || slot >= 4) return LFD_E_ARG;
    cudaSetDevice(h->device);
    CK(cudaEventRecord(h->mark[slot], h->stream));
    return LFD_OK;
}

extern "C" int lfd_timer_elapsed(lfd_handle* h, int slot_start, lfd_handle* h_end, int slot_end, float* ms)
{
    if (!h || !h_end || !ms || slot_start < 0 || slot_start >= 4 || slot_end < 0 || slot_end >= 4) return LFD_E_ARG;
    cudaSetDevice(h->device);
    CK(cudaEventSynchronize(h->mark[slot_start]));
    CK(cudaEventSynchronize(h_end->mark[slot_end]));
    CK(cudaEventElapsedTime(ms, h->mark[slot_start], h_end->mark[slot_end]));
    return LFD_OK;
}

// developer aid: per-launch device times of the last run when LFD_KTIMING=1 (source line of the launch, ms)
extern "C" int lfd_get_ktimings(lfd_handle* h, int* lines, float* ms, int max_entries, int* n_entries)
{
    if (!h || !lines || !ms || !n_entries) return LFD_E_ARG;
    int n = 0;
    for (int i = 1; i < h->kn && n < max_entries; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->kev[i - 1], h->kev[i]) != cudaSuccess) { cudaGetLastError(); t = 0.f; }
        lines[n] = h->kline[i]; ms[n] = t; n++;
    }
    *n_entries = n;
    return LFD_OK;
}

extern "C" const char* lfd_timing_name(int i) { return (i >= 0 && i < N_TIMINGS) ? k_timing_names[i] : ""; }

extern "C" int64_t lfd_kernel_launches(const lfd_handle* h) { return h ? h->launches : 0; }

extern "C" int lfd_get_counters(lfd_handle* h, int64_t* out, int n)
{
    if (!h || !out) return LFD_E_ARG;
    for (int i = 0; i < n && i < 16; i++) out[i] = h->counters_h[i];
    return LFD_OK;
}
