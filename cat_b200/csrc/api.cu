// C ABI of the library (include/ctc_crf_b200.h): den-graph lifetime, workspace carving, launch sequencing.
// Mirrors the roles of the reference's binding.cpp:51-117 + den_calculate.cu:288-481 host code +
// gpu_ctc/ctc_entrypoint.cu:29-109, without torch and without ever calling exit().
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ctc_crf_b200.h"
#include "common.cuh"

using namespace ccb;

int DEN_NUM_ARCS = 0;
int DEN_NUM_STATES = 0;

namespace {

constexpr int kMaxDevices = 64;
thread_local std::string g_err;
std::mutex g_mu;
DenPlan g_plan;                       // host copy (one den graph per process, like den_calculate.cu:263-273)
bool g_plan_valid = false;
DeviceGraph g_dev[kMaxDevices];
int g_refs[kMaxDevices] = {0};        // Init() calls not yet matched by Release(), per device: `ctx = CRFContext(new_graph)` runs the
                                      // old context's __del__ AFTER the new Init -- that Release must not free the new graph
std::atomic<long> g_launches{0};

// scratch owned by the library for the reference-signature entry points (Section 1)
struct LegacyScratch {
    void *aux = nullptr; size_t aux_bytes = 0;
    float *alpha = nullptr; size_t alpha_floats = 0;
    void *ctc = nullptr; size_t ctc_bytes = 0;
    int N = 0, T = 0;
    bool fwd_valid = false;            // compute_alpha ran and compute_beta_and_grad has not consumed it yet
    cudaStream_t stream = nullptr;     // stream the scratch was last allocated / used on (stream-ordered allocation)
    bool used = false;
};
LegacyScratch g_legacy[kMaxDevices];

// Side stream of the fused loss (OPT-IN, CCB_OVERLAP=1 at Init; measured and left off, profiles/r02_experiments.md section 14):
// the numerator's alpha || beta chains (a latency-bound ~1 ms on 2N small CTAs) run here next to the den forward pass, whose
// persistent CTAs leave room for one of them per SM (78 x 512 + 39 x 256 registers, 167 + 11 KB of shared memory).  On the
// B200 the two grids do become co-resident, but not profitably: either the cooperative den grid is admitted only after ~1 ms
// (step -0.4 ms), or both run at once and BOTH crawl until the numerator is through (numerator 1.5 -> 5.4 ms, den forward + 5 ms).
// Created at Init, destroyed by Release.
struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
    int order = 0;                     // CCB_OVERLAP_ORDER (tuning): 0 numerator launched before the den forward pass, 1 after it
    bool carve = true;                 // CCB_OVERLAP_CARVE=0 (tuning): leave the numerator kernel's L1 / shared-memory split at its default
    bool trace = false;                // CCB_OVERLAP_TRACE=1 (diagnostics): timed events around each part, printed to stderr (synchronises!)
    cudaEvent_t tr[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
SideStream g_side[kMaxDevices];

void FreeSide(SideStream &sd) {
    if (sd.fork) cudaEventDestroy(sd.fork);
    if (sd.join) cudaEventDestroy(sd.join);
    if (sd.s) cudaStreamDestroy(sd.s);
    for (cudaEvent_t e : sd.tr) if (e) cudaEventDestroy(e);
    sd = SideStream();
}

void MakeSide(SideStream &sd) {   // (current device); failure just leaves the single-stream path
    FreeSide(sd);
    const char *e = getenv("CCB_OVERLAP");
    if (!(e && e[0] == '1')) return;
    if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        FreeSide(sd);
        return;
    }
    sd.ok = true;
    { const char *o = getenv("CCB_OVERLAP_ORDER"); sd.order = o ? atoi(o) : 0; }
    { const char *c = getenv("CCB_OVERLAP_CARVE"); sd.carve = !(c && c[0] == '0'); }
    { const char *t = getenv("CCB_OVERLAP_TRACE"); sd.trace = t && t[0] == '1'; }
    if (sd.trace) for (cudaEvent_t &e : sd.tr) if (cudaEventCreate(&e) != cudaSuccess) { sd.trace = false; break; }
}

// profiling aid (ccb_debug_timeline)
unsigned long long *g_timeline = nullptr;
int g_tl_step0 = 0, g_tl_steps = 0;

int Fail(const std::string &m) { g_err = m; return 1; }
int FailCuda(const char *what, cudaError_t e) { g_err = std::string(what) + ": " + cudaGetErrorString(e); return (int)e ? (int)e : 1; }

#define CCB_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return FailCuda(#call, e__);             \
    } while (0)

template <typename T>
int Upload(T **dst, const std::vector<T> &src) {
    CCB_CUDA(cudaMalloc((void **)dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
    if (!src.empty()) CCB_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

void FreeDevice(DeviceGraph &d) {
    if (!d.loaded) return;
    cudaFree(d.state_label); cudaFree(d.state_pos); cudaFree(d.final_lin); cudaFree(d.start_arcs); cudaFree(d.hub_states);
    for (DevicePass *p : {&d.fwd, &d.bwd}) { cudaFree(p->arcs); cudaFree(p->chunk_state); cudaFree(p->chunk_arc); cudaFree(p->chunk_pair); cudaFree(p->cta_labels); cudaFree(p->w1); cudaFree(p->tq); cudaFree(p->own_c); }
    d = DeviceGraph();
}

int UploadPass(const PassPlan &h, DevicePass *d) {
    if (Upload(&d->arcs, h.arcs)) return 1;
    if (Upload(&d->chunk_state, h.chunk_state)) return 1;
    if (Upload(&d->chunk_arc, h.chunk_arc)) return 1;
    if (Upload(&d->chunk_pair, h.chunk_pair)) return 1;
    if (Upload(&d->cta_labels, h.cta_labels)) return 1;
    if (Upload(&d->w1, h.w1)) return 1;
    d->num_arcs = (int)h.arcs.size();
    d->max_tile_arcs = h.max_tile_arcs;
    d->max_tile_labels = h.max_tile_labels;
    d->max_tile_rows = h.max_tile_rows;
    return 0;
}

// The backward pass's quads as the streamed-arc TMA kernel reads them (DevicePass::tq): per quad {row coordinate of slot 0..3}
// {first-weight bits 0..3}{second-weight bits 0..3}; padding slots (both weights zero) name an out-of-bounds row, which the
// TMA unit zero-fills without a fetch.
int UploadTransposedQuads(const PassPlan &h, DevicePass *d) {
    const size_t nq = h.arcs.size() / kQuad;
    std::vector<uint4> tq(nq * 3);
    for (size_t i = 0; i < nq; ++i) {
        uint32_t c[kQuad], w0[kQuad], w1[kQuad];
        for (int j = 0; j < kQuad; ++j) {
            const Arc &a = h.arcs[i * kQuad + j];
            memcpy(&w0[j], &a.w, 4);
            memcpy(&w1[j], &h.w1[i * kQuad + j], 4);
            c[j] = ((w0[j] & 0x7fffffffu) == 0u && w1[j] == 0u) ? kOobRow : (uint32_t)a.peer;
        }
        tq[3 * i] = make_uint4(c[0], c[1], c[2], c[3]);
        tq[3 * i + 1] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        tq[3 * i + 2] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    }
    return Upload(&d->tq, tq);
}

int CurrentGraph(DeviceGraph **out) {
    int dev = 0;
    CCB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices || !g_dev[dev].loaded)
        return Fail("den graph not loaded on CUDA device " + std::to_string(dev) + " (call Init / CRFContext first)");
    *out = &g_dev[dev];
    return 0;
}

// warps per CTA of the persistent den kernels (512 threads, one CTA per SM); tuning builds: CCB_DEN_WARPS=8 for experiments
int DenWarps() {
#ifdef CCB_TUNING
    const char *e = getenv("CCB_DEN_WARPS");
    if (e && (atoi(e) == 8 || atoi(e) == 24)) return atoi(e);
#endif
    return 16;
}

int InitImpl(const char *fst_name, int n_gpus, const int *gpus) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!fst_name || n_gpus <= 0 || !gpus) return Fail("Init: bad arguments");
    HostFst fst;
    std::string err;
    if (!ReadFstFile(fst_name, &fst, &err)) return Fail(err);
    int prev = 0;
    CCB_CUDA(cudaGetDevice(&prev));
    int count = 0;
    CCB_CUDA(cudaGetDeviceCount(&count));
    for (int i = 0; i < n_gpus; ++i)
        if (gpus[i] < 0 || gpus[i] >= count || gpus[i] >= kMaxDevices) return Fail("Init: invalid GPU id " + std::to_string(gpus[i]));
    cudaDeviceProp prop;
    CCB_CUDA(cudaGetDeviceProperties(&prop, gpus[0]));
    if (!prop.cooperativeLaunch) return Fail("device does not support cooperative launch");
    if (!BuildDenPlan(fst, prop.multiProcessorCount, DenWarps(), &g_plan, &err)) return Fail(err);
    g_plan_valid = true;
    DEN_NUM_STATES = g_plan.num_states + g_plan.num_pairs;   // rows per alpha frame (what binding.cpp sizes scratch with)
    DEN_NUM_ARCS = (int)g_plan.fwd.arcs.size();
    int rc = 0;
    for (int i = 0; i < n_gpus && rc == 0; ++i) {
        DeviceGraph &d = g_dev[gpus[i]];
        CCB_CUDA(cudaSetDevice(gpus[i]));
        FreeDevice(d);
        cudaDeviceProp p2;
        CCB_CUDA(cudaGetDeviceProperties(&p2, gpus[i]));
        if (p2.multiProcessorCount != prop.multiProcessorCount) { rc = Fail("Init: heterogeneous GPUs are not supported"); break; }
        d.device = gpus[i];
        d.S = g_plan.num_states; d.P = g_plan.num_pairs; d.start = g_plan.start; d.num_labels = g_plan.num_labels;
        d.scale_exp = g_plan.scale_exp;
        d.n_ctas = g_plan.n_ctas; d.n_warps = g_plan.n_warps;
        d.max_smem_optin = (int)p2.sharedMemPerBlockOptin;
        rc = Upload(&d.state_label, g_plan.state_label) || Upload(&d.state_pos, g_plan.state_pos) ||
             Upload(&d.final_lin, g_plan.final_lin) || Upload(&d.start_arcs, g_plan.start_arcs) ||
             Upload(&d.hub_states, g_plan.hub_states) ||
             UploadPass(g_plan.fwd, &d.fwd) || UploadPass(g_plan.bwd, &d.bwd) ||
             Upload(&d.fwd.own_c, g_plan.own_fwd) || Upload(&d.bwd.own_c, g_plan.own_bwd);
        d.own_rows = g_plan.own_rows;
        { const char *e = getenv("CCB_ARCS_IN_GLOBAL"); d.tune_arcs_in_global = e && e[0] == '1'; }
        { const char *e = getenv("CCB_W1_IN_GLOBAL"); d.tune_w1_in_global = e && e[0] == '1'; }
        { const char *e = getenv("CCB_NO_TMA"); d.tune_no_tma = e && e[0] == '1'; }   // A/B: register gathers instead of TMA gather4
        { const char *e = getenv("CCB_RING_ROWS"); d.tune_ring_rows = e ? atoi(e) : 0; }   // A/B: rows per TMA ring stage (16 / 8)
        // streamed-arc tier: only graphs whose arc stream may not fit shared memory next to the TMA rings carry the copy
        if (rc == 0 && g_plan.hub_states.empty() && !d.tune_no_tma &&
            (d.tune_arcs_in_global || (size_t)d.bwd.max_tile_arcs * 12 > kStreamTierArcBytes))
            rc = UploadTransposedQuads(g_plan.bwd, &d.bwd);   // (the forward pass of such graphs keeps register gathers)
        if (d.tune_arcs_in_global || d.tune_w1_in_global)
            fprintf(stderr, "ctc_crf_b200: CCB_ARCS_IN_GLOBAL / CCB_W1_IN_GLOBAL set -- den arc tiles forced out of shared memory (test hook, slow)\n");
        {   // small-batch kernels: both arc streams in shared memory next to 4 KB (8 KB) of rings per warp, no hub rows
            const size_t budget = (size_t)d.max_smem_optin > 2048 ? (size_t)d.max_smem_optin - 1024 : 0;
            // (the backward pass may keep only offsets + first weights resident, 8 bytes per slot, and stream the second weights)
            const size_t ring = (size_t)d.n_warps * (4 * 16 * 64 + 4 * (8 + 64));
            const size_t fwd_need = (size_t)d.fwd.max_tile_arcs * sizeof(Arc) + (size_t)(16 + d.fwd.max_tile_rows) * 4 + ring + 512;
            const size_t bwd_need = (size_t)d.bwd.max_tile_arcs * 8 + ((size_t)(2 + d.bwd.max_tile_labels) * 16 + 2 * (size_t)d.bwd.max_tile_rows) * 4 + ring + 512;
            d.small_ok = !d.tune_no_tma && !d.tune_arcs_in_global && g_plan.hub_states.empty() &&
                         fwd_need <= budget && bwd_need <= budget && getenv("CCB_NO_SMALL") == nullptr;
        }
        d.n_start_arcs = (int)g_plan.start_arcs.size();
        d.n_hubs = (int)g_plan.hub_states.size();
        d.start_final = g_plan.final_lin[(size_t)g_plan.start];
        d.loaded = (rc == 0);
        if (rc == 0) { ++g_refs[gpus[i]]; MakeSide(g_side[gpus[i]]); }
    }
    cudaSetDevice(prev);
    return rc;
}

int ReleaseImpl(int n_gpus, const int *gpus) {
    std::lock_guard<std::mutex> lk(g_mu);
    int prev = 0;
    cudaGetDevice(&prev);
    for (int i = 0; i < n_gpus; ++i) {
        if (gpus[i] < 0 || gpus[i] >= kMaxDevices) continue;
        if (g_refs[gpus[i]] > 0 && --g_refs[gpus[i]] > 0) continue;   // a newer Init() still owns this device's graph
        if (cudaSetDevice(gpus[i]) != cudaSuccess) continue;
        cudaDeviceSynchronize();
        FreeDevice(g_dev[gpus[i]]);
        FreeSide(g_side[gpus[i]]);
        LegacyScratch &ls = g_legacy[gpus[i]];
        cudaFree(ls.aux); cudaFree(ls.alpha); cudaFree(ls.ctc);
        ls = LegacyScratch();
    }
    cudaSetDevice(prev);
    return 0;
}

DenParams BaseParams(const DeviceGraph &g, const void *y, int dtype, long sn, long st, int N, int T, int V,
                     const int *len, float *alpha, void *aux, const DenAuxLayout &L) {
    DenParams p;
    memset(&p, 0, sizeof(p));
    char *a = reinterpret_cast<char *>(aux);
    p.state_label = g.state_label; p.state_pos = g.state_pos; p.final_lin = g.final_lin;
    p.start_arcs = g.start_arcs; p.n_start_arcs = g.n_start_arcs;
    p.hub_states = g.hub_states; p.n_hubs = g.n_hubs;
    p.S = g.S; p.num_pairs = g.P; p.start = g.start; p.n_warps = g.n_warps;
    p.scale_exp = g.scale_exp;
    p.start_final = g.start_final;
    p.y = y; p.y_bf16 = (dtype == CCB_DTYPE_BF16); p.sn = sn; p.st = st;
    p.N = N; p.Npad = L.Npad; p.Tmax = T; p.V = V; p.len = len;
    p.alpha = alpha;
    p.bh = reinterpret_cast<float *>(a + L.bh);
    p.colsum_a = reinterpret_cast<float *>(a + L.colsum_a);
    p.colsum_b = reinterpret_cast<float *>(a + L.colsum_b);
    p.absum = reinterpret_cast<float *>(a + L.absum);
    p.zsum = reinterpret_cast<float *>(a + L.zsum);
    p.b0 = reinterpret_cast<float *>(a + L.b0);
    p.fmax = reinterpret_cast<float *>(a + L.fmax);
    p.timeline = g_timeline; p.tl_step0 = g_tl_step0; p.tl_steps = g_tl_steps;
#ifdef CCB_TUNING   // timing-experiment switches (skip row ends / barriers: WRONG RESULTS) exist in tuning builds only
    { const char *e = getenv("CCB_DEBUG"); p.debug = e ? atoi(e) : 0; }
#endif
    return p;
}

int CheckDen(const DeviceGraph &g, int dtype, int N, int T, int V) {
    if (N <= 0 || T <= 0 || V <= 0) return Fail("den: empty batch");
    if (dtype != CCB_DTYPE_F32 && dtype != CCB_DTYPE_BF16) return Fail("den: unsupported logits dtype");
    if (V < g.num_labels)
        return Fail("den graph uses label " + std::to_string(g.num_labels - 1) + " but logits have only " + std::to_string(V) + " classes");
    if ((size_t)(2 * (size_t)g.S + g.P) * (size_t)PadLanes(N, g.small_ok) * 4 >= ((size_t)1 << 32)) return Fail("den: states x batch too large for 32-bit row offsets; split the batch");
    if (PadLanes(N, g.small_ok) > g.n_warps * 32) return Fail("den: batch larger than " + std::to_string(g.n_warps * 32) + " utterances per call; split the batch");
    return 0;
}

// forward part: zero aux, frame max, alpha recursion; logz_a lands in aux
// raw: `y` holds unnormalised logits; the log-normalisers go to aux (lz, lnorm) and logZ comes out normalised
int DenForward(const DeviceGraph &g, const void *y, int dtype, long sn, long st, int N, int T, int V, const int *len,
               float *alpha, void *aux, cudaStream_t stream, bool raw = false) {
    const DenAuxLayout L = MakeDenAuxLayout(g.S, N, T, g.small_ok);
    char *a = reinterpret_cast<char *>(aux);
    CCB_CUDA(cudaMemsetAsync(a, 0, L.zero_bytes, stream));
    int rc = raw ? LaunchFrameLse(y, dtype == CCB_DTYPE_BF16, sn, st, N, T, V, len, reinterpret_cast<float *>(a + L.fmax),
                                  reinterpret_cast<float *>(a + L.lz), reinterpret_cast<double *>(a + L.lnorm), L.Npad, stream)
                 : LaunchFrameMax(y, dtype == CCB_DTYPE_BF16, sn, st, N, T, V, len, reinterpret_cast<float *>(a + L.fmax), L.Npad, stream);
    if (rc) return FailCuda("frame_max", (cudaError_t)rc);
    DenParams p = BaseParams(g, y, dtype, sn, st, N, T, V, len, alpha, aux, L);
    if (raw) p.lnorm = reinterpret_cast<const double *>(a + L.lnorm);
    p.barrier = reinterpret_cast<unsigned *>(a + L.barrier);
    p.logz = reinterpret_cast<float *>(a + L.logz_a);
    std::string err;
    rc = LaunchDenForward(g, p, stream, &err);
    if (rc) return Fail(err);
    return 0;
}

int DenBackward(const DeviceGraph &g, const void *y, int dtype, long sn, long st, int N, int T, int V, const int *len,
                float *alpha, void *aux, float *grad, long gsn, long gst, float grad_scale, cudaStream_t stream,
                bool raw = false) {
    const DenAuxLayout L = MakeDenAuxLayout(g.S, N, T, g.small_ok);
    char *a = reinterpret_cast<char *>(aux);
    DenParams p = BaseParams(g, y, dtype, sn, st, N, T, V, len, alpha, aux, L);
    if (raw) p.lnorm = reinterpret_cast<const double *>(a + L.lnorm);
    p.barrier = reinterpret_cast<unsigned *>(a + L.barrier + 128);
    p.logz = reinterpret_cast<float *>(a + L.logz_b);
    p.grad = grad; p.gsn = gsn; p.gst = gst;
    // the pass accumulates into colsum_b / absum / b0 and counts on its own barrier word: zero them here, so that a second
    // backward pass over the same forward pass (or a reused aux buffer) starts clean instead of relying on DenForward's memset
    CCB_CUDA(cudaMemsetAsync(a + L.colsum_b, 0, L.zsum - L.colsum_b, stream));
    CCB_CUDA(cudaMemsetAsync(a + L.b0, 0, (size_t)L.Npad * 4, stream));
    CCB_CUDA(cudaMemsetAsync(a + L.barrier + 128, 0, 128, stream));
    std::string err;
    int rc = LaunchDenBackward(g, p, stream, &err);
    if (rc) return Fail(err);
    rc = LaunchDenGradNormalize(grad, gsn, gst, p.absum, len, N, L.Npad, T, V, grad_scale, stream);
    if (rc) return FailCuda("den_grad_normalize", (cudaError_t)rc);
    return 0;
}

// Scratch of the reference-signature route, taken STREAM-ORDERED from the device's default memory pool
// (cudaMallocAsync / cudaFreeAsync on the caller's stream): growing it never synchronises the device.  It is kept
// between calls and returned by Release().  If the caller moves to another stream the old stream is drained first.
int GrowAsync(void **ptr, size_t *have, size_t need, cudaStream_t s) {
    if (*have >= need) return 0;
    if (*ptr) CCB_CUDA(cudaFreeAsync(*ptr, s));
    *ptr = nullptr; *have = 0;
    CCB_CUDA(cudaMallocAsync(ptr, need, s));
    *have = need;
    return 0;
}

int EnsureLegacy(int dev, const DeviceGraph &g, int N, int T, float *caller_alpha, size_t caller_floats, float **alpha_out,
                 cudaStream_t s) {
    LegacyScratch &ls = g_legacy[dev];
    if (ls.used && ls.stream != s) CCB_CUDA(cudaStreamSynchronize(ls.stream));
    ls.stream = s; ls.used = true;
    const DenAuxLayout L = MakeDenAuxLayout(g.S, N, T, g.small_ok);
    if (GrowAsync(&ls.aux, &ls.aux_bytes, L.total, s)) return 1;
    const size_t need = ccb_den_alpha_floats(N, T);
    if (caller_alpha && caller_floats >= need) { *alpha_out = caller_alpha; return 0; }
    size_t have_bytes = ls.alpha_floats * sizeof(float);
    void *ap = ls.alpha;
    if (GrowAsync(&ap, &have_bytes, need * sizeof(float), s)) { ls.alpha = nullptr; ls.alpha_floats = 0; return 1; }
    ls.alpha = reinterpret_cast<float *>(ap); ls.alpha_floats = have_bytes / sizeof(float);
    *alpha_out = ls.alpha;
    return 0;
}

}  // namespace

namespace ccb {
void CountLaunch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace ccb

extern "C" {

const char *ccb_last_error(void) { return g_err.c_str(); }
long ccb_launch_count(void) { return g_launches.load(); }
void ccb_debug_timeline(void *dev_buffer, int step0, int nsteps) {
    g_timeline = reinterpret_cast<unsigned long long *>(dev_buffer); g_tl_step0 = step0; g_tl_steps = nsteps;
}
int ccb_den_info(long *info) {
    if (!g_plan_valid || !info) return 1;
    info[0] = g_plan.file_states; info[1] = g_plan.file_arcs; info[2] = g_plan.num_states; info[3] = g_plan.num_pairs;
    info[4] = (long)g_plan.fwd.arcs.size(); info[5] = (long)g_plan.bwd.arcs.size();
    info[6] = g_plan.fwd.real_arcs; info[7] = g_plan.bwd.real_arcs;
    return 0;
}
int ccb_den_loaded(int device) { return device >= 0 && device < kMaxDevices && g_dev[device].loaded ? 1 : 0; }

void Init(const char *fst_name, int n_gpus, int *gpus) {
    g_err.clear();
    if (InitImpl(fst_name, n_gpus, gpus) != 0) fprintf(stderr, "ctc_crf_b200 Init failed: %s\n", g_err.c_str());
}

void Release(int n_gpus, int *gpus) { ReleaseImpl(n_gpus, gpus); }

// (S, P) of the graph on the CURRENT device when one is loaded there (devices may hold different graphs if Init was
// called per device with different files), else of the last plan built
static void CurrentSizes(int *S, int *P, bool *small_ok) {
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < kMaxDevices && g_dev[dev].loaded) {
        *S = g_dev[dev].S; *P = g_dev[dev].P; *small_ok = g_dev[dev].small_ok;
        return;
    }
    cudaGetLastError();
    *S = g_plan_valid ? g_plan.num_states : 0; *P = g_plan_valid ? g_plan.num_pairs : 0; *small_ok = false;
}

size_t ccb_den_alpha_floats(int N, int T) {
    int S, P;
    bool small_ok;
    CurrentSizes(&S, &P, &small_ok);
    // T+1 frames of S real rows; the virtual pair-sum rows of a frame are parked two frames ahead (den_kernels.cu), so two
    // more frames at the end when the plan has pairs
    return (size_t)(T + 1 + (P > 0 ? 2 : 0)) * (size_t)S * (size_t)PadLanes(N, small_ok);
}

size_t ccb_den_aux_bytes(int N, int T) {
    int S, P;
    bool small_ok;
    CurrentSizes(&S, &P, &small_ok);
    return S ? MakeDenAuxLayout(S, N, T, small_ok).total : 0;
}

size_t ccb_ctc_workspace_bytes(int N, int T, int max_label_len) {
    // per utterance: alpha and beta cells [T][2L+1] in double + the fp64 log-likelihood (ctc_kernels.cu)
    // ... and, behind them, N floats for the log-likelihoods of the fused loss (LossFwdImpl: `logp`)
    const size_t per_utt = 2 * (size_t)T * (2 * (size_t)max_label_len + 1) + 1;
    return ((size_t)N * per_utt + 32 + ((size_t)N + 1) / 2) * sizeof(double);
}

void compute_alpha(float *alpha, float *logits, const int batch_size, int T, const int alpha_size,
                   int logits_size, int *input_lengths, float *loglikelihood, void *stream) {
    g_err.clear();
    DeviceGraph *g;
    if (CurrentGraph(&g)) return;
    if (CheckDen(*g, CCB_DTYPE_F32, batch_size, T, logits_size)) return;
    float *al = nullptr;
    const size_t caller = (size_t)(T + 1) * (size_t)batch_size * (size_t)(alpha_size > 0 ? alpha_size : 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (EnsureLegacy(g->device, *g, batch_size, T, alpha, caller, &al, s)) return;
    LegacyScratch &ls = g_legacy[g->device];
    ls.N = batch_size; ls.T = T; ls.fwd_valid = false;
    if (DenForward(*g, logits, CCB_DTYPE_F32, (long)T * logits_size, logits_size, batch_size, T, logits_size,
                   input_lengths, al, ls.aux, s)) return;
    ls.fwd_valid = true;
    const DenAuxLayout L = MakeDenAuxLayout(g->S, batch_size, T, g->small_ok);
    cudaError_t e = cudaMemcpyAsync(loglikelihood, (char *)ls.aux + L.logz_a, sizeof(float) * batch_size, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) FailCuda("copy logZ", e);
}

void compute_beta_and_grad(float *beta, const float *const alpha, const float *const logits,
                           const float *const alpha_lld, float *grad_storage, float *grad_net,
                           const int batch_size, const int T, const int beta_size, const int logits_size,
                           const int *const input_lengths, float *loglikelihood, void *stream) {
    (void)beta; (void)alpha_lld; (void)grad_storage;
    g_err.clear();
    DeviceGraph *g;
    if (CurrentGraph(&g)) return;
    LegacyScratch &ls = g_legacy[g->device];
    // the backward pass continues from the forward pass's column sums / barrier words in the scratch: exactly one
    // compute_beta_and_grad per compute_alpha, same batch
    if (ls.N != batch_size || ls.T != T || !ls.aux || !ls.fwd_valid) { Fail("compute_beta_and_grad: call compute_alpha on the same batch first (once per backward call)"); return; }
    ls.fwd_valid = false;
    float *al = nullptr;
    const size_t caller = (size_t)(T + 1) * (size_t)batch_size * (size_t)(beta_size > 0 ? beta_size : 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (EnsureLegacy(g->device, *g, batch_size, T, const_cast<float *>(alpha), caller, &al, s)) return;
    if (DenBackward(*g, logits, CCB_DTYPE_F32, (long)T * logits_size, logits_size, batch_size, T, logits_size,
                    input_lengths, al, ls.aux, grad_net, (long)T * logits_size, logits_size, 1.f, s)) return;
    if (loglikelihood) {
        const DenAuxLayout L = MakeDenAuxLayout(g->S, batch_size, T, g->small_ok);
        cudaError_t e = cudaMemcpyAsync(loglikelihood, (char *)ls.aux + L.logz_b, sizeof(float) * batch_size, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) FailCuda("copy logZ(beta)", e);
    }
}

int ccb_den_forward_backward(const void *logits, int dtype, long sn, long st, int N, int T, int V,
                             const int *len_dev, float *alpha_ws, void *aux_ws, float *grad, long gsn, long gst,
                             float grad_scale, float *logz, float *logz_beta, void *stream) {
    g_err.clear();
    DeviceGraph *g;
    if (CurrentGraph(&g)) return 1;
    if (CheckDen(*g, dtype, N, T, V)) return 1;
    if (!alpha_ws || !aux_ws) return Fail("den: workspace missing");
    cudaStream_t s = (cudaStream_t)stream;
    if (DenForward(*g, logits, dtype, sn, st, N, T, V, len_dev, alpha_ws, aux_ws, s)) return 1;
    const DenAuxLayout L = MakeDenAuxLayout(g->S, N, T, g->small_ok);
    if (logz) CCB_CUDA(cudaMemcpyAsync(logz, (char *)aux_ws + L.logz_a, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
    if (grad) {
        if (DenBackward(*g, logits, dtype, sn, st, N, T, V, len_dev, alpha_ws, aux_ws, grad, gsn, gst, grad_scale, s)) return 1;
        if (logz_beta) CCB_CUDA(cudaMemcpyAsync(logz_beta, (char *)aux_ws + L.logz_b, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

#ifdef CCB_TUNING
/* tuning builds only: the backward pass alone over the spill of an earlier forward pass (overlap experiments) */
int ccb_debug_den_backward(const void *logits, int dtype, long sn, long st, int N, int T, int V, const int *len_dev,
                           float *alpha_ws, void *aux_ws, float *grad, long gsn, long gst, void *stream) {
    g_err.clear();
    DeviceGraph *g;
    if (CurrentGraph(&g)) return 1;
    return DenBackward(*g, logits, dtype, sn, st, N, T, V, len_dev, alpha_ws, aux_ws, grad, gsn, gst, 1.f, (cudaStream_t)stream);
}
#endif

int ccb_ctc_forward_backward(const void *logits, int dtype, long sn, long st, int N, int T, int V,
                             const int *labels_dev, const int *label_off_dev, const int *label_len_dev,
                             const int *len_dev, int max_label_len, int blank, void *workspace,
                             float *grad, long gsn, long gst, float grad_scale, float *logp, void *stream) {
    g_err.clear();
    if (N <= 0 || T <= 0 || V <= 0 || !logits || !workspace || !logp) return Fail("ctc: bad arguments");
    if (blank < 0 || blank >= V) return Fail("ctc: blank label out of range");
    std::string err;
    int rc = LaunchCtc(logits, dtype == CCB_DTYPE_BF16, sn, st, N, T, V, labels_dev, label_off_dev, label_len_dev,
                       len_dev, max_label_len, blank, reinterpret_cast<float *>(workspace), grad, gsn, gst, grad_scale,
                       logp, nullptr, (cudaStream_t)stream, &err);
    if (rc) return Fail(err);
    return 0;
}

/* CTC-only loss (WARP_CTC_LOSS, ctc_crf/__init__.py:25-56) on the (N,T,V) block in place: one numerator pass, the gradient
 * rows all written by the occupancy kernel (no zero fill), the scalar assembled on the device (no host round trip). */
int ccb_ctc_loss_fwd(const void *logits, int dtype, int N, int T, int V, int Tmax,
                     const int *labels_dev, const int *label_off_dev, const int *label_len_dev, const int *len_dev,
                     int max_label_len, int blank, float scale, void *ctc_ws, float *grad, float *loss, float *logp,
                     void *stream) {
    g_err.clear();
    if (N <= 0 || T <= 0 || V <= 0 || !logits || !ctc_ws || !grad || !loss || !logp) return Fail("ctc_loss_fwd: bad arguments");
    if (Tmax <= 0 || Tmax > T) return Fail("ctc_loss_fwd: Tmax must lie in [1, T]");
    if (dtype != CCB_DTYPE_F32 && dtype != CCB_DTYPE_BF16) return Fail("ctc_loss_fwd: unsupported logits dtype");
    if (blank < 0 || blank >= V) return Fail("ctc: blank label out of range");
    cudaStream_t s = (cudaStream_t)stream;
    const long sn = (long)T * V, st = V;
    std::string err;
    int rc = LaunchCtc(logits, dtype == CCB_DTYPE_BF16, sn, st, N, Tmax, V, labels_dev, label_off_dev, label_len_dev, len_dev,
                       max_label_len, blank, reinterpret_cast<float *>(ctc_ws), grad, sn, st, -scale, logp, nullptr, s, &err,
                       /*overwrite=*/1, /*Tfull=*/T);
    if (rc) return Fail(err);
    rc = LaunchSumScale(logp, N, -scale, loss, s);
    if (rc) return FailCuda("sum_scale", (cudaError_t)rc);
    return 0;
}

size_t ccb_ctc_align_workspace_bytes(int N, int T, int max_label_len) {
    return (size_t)N * (size_t)T * (2 * (size_t)max_label_len + 1) + 256;
}

int ccb_ctc_align(const void *logits, int dtype, long sn, long st, int N, int T, int V,
                  const int *labels_dev, const int *label_off_dev, const int *label_len_dev, const int *len_dev,
                  int max_label_len, int blank, void *workspace, int *align, float *score, void *stream) {
    g_err.clear();
    if (N <= 0 || T <= 0 || V <= 0 || !logits || !workspace || !align) return Fail("ctc_align: bad arguments");
    if (dtype != CCB_DTYPE_F32 && dtype != CCB_DTYPE_BF16) return Fail("ctc_align: unsupported logits dtype");
    if (blank < 0 || blank >= V) return Fail("ctc: blank label out of range");
    std::string err;
    int rc = LaunchCtcViterbi(logits, dtype == CCB_DTYPE_BF16, sn, st, N, T, V, labels_dev, label_off_dev, label_len_dev, len_dev,
                              max_label_len, blank, reinterpret_cast<unsigned char *>(workspace), align, score, (cudaStream_t)stream, &err);
    if (rc) return Fail(err);
    return 0;
}

static int LossFwdImpl(bool raw, const void *logits, int dtype, int N, int T, int V, int Tmax,
                       const int *labels_dev, const int *label_off_dev, const int *label_len_dev,
                       const int *len_dev, int max_label_len, float lamb, float scale,
                       float *alpha_ws, void *aux_ws, void *ctc_ws, float *grad, float *loss, float *parts,
                       void *stream) {
    g_err.clear();
    DeviceGraph *g;
    if (CurrentGraph(&g)) return 1;
    if (Tmax <= 0 || Tmax > T) return Fail("ctc_crf_loss_fwd: Tmax must lie in [1, T]");
    if (CheckDen(*g, dtype, N, Tmax, V)) return 1;
    if (!alpha_ws || !aux_ws || !ctc_ws || !grad || !loss) return Fail("ctc_crf_loss_fwd: missing buffer");
    cudaStream_t s = (cudaStream_t)stream;
    const long sn = (long)T * V, st = V;     // the (N,T,V) block is addressed in place; only Tmax frames are walked
    const DenAuxLayout L = MakeDenAuxLayout(g->S, N, Tmax, g->small_ok);
    float *logz = reinterpret_cast<float *>((char *)aux_ws + L.logz_a);
    // log p(l|x) per utterance: behind the numerator's cells in ITS workspace (not in the den aux block: the numerator may run
    // concurrently with the den passes)
    float *logp = reinterpret_cast<float *>(reinterpret_cast<double *>(ctc_ws) +
                                            (size_t)N * (2 * (size_t)Tmax * (2 * (size_t)max_label_len + 1) + 1) + 32);
    const double *lnorm = raw ? reinterpret_cast<const double *>((char *)aux_ws + L.lnorm) : nullptr;
    std::string err;
    int dev = 0;
    CCB_CUDA(cudaGetDevice(&dev));
    // Fork: the numerator's chains need only y and the labels.  (raw-logit entry: they also need the frame normalisers the den
    // prologue computes on the caller's stream -- kept in line there.)
    SideStream *sd = (!raw && dev < kMaxDevices && g_side[dev].ok) ? &g_side[dev] : nullptr;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (sd && (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone)) { cudaGetLastError(); sd = nullptr; }
    int rc = 0;
    auto side_numerator = [&]() -> int {
        if (sd->trace) cudaEventRecord(sd->tr[1], sd->s);
        int r = LaunchCtcAlphaBeta(logits, dtype == CCB_DTYPE_BF16, sn, st, N, Tmax, V, labels_dev, label_off_dev, label_len_dev, len_dev,
                                   max_label_len, 0, reinterpret_cast<float *>(ctc_ws), true, logp, nullptr, sd->s, &err, sd->carve);
        if (r) return Fail(err);
        if (sd->trace) cudaEventRecord(sd->tr[2], sd->s);
        CCB_CUDA(cudaEventRecord(sd->join, sd->s));
        return 0;
    };
    if (sd) {
        if (sd->trace) cudaEventRecord(sd->tr[0], s);
        CCB_CUDA(cudaEventRecord(sd->fork, s));
        CCB_CUDA(cudaStreamWaitEvent(sd->s, sd->fork, 0));
        if (sd->order == 0 && side_numerator()) return 1;
    }
    CCB_CUDA(cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)N * T * V, s));
    if (sd && sd->trace) cudaEventRecord(sd->tr[3], s);
    if (DenForward(*g, logits, dtype, sn, st, N, Tmax, V, len_dev, alpha_ws, aux_ws, s, raw)) rc = 1;
    if (sd && sd->trace) cudaEventRecord(sd->tr[4], s);
    if (sd && sd->order != 0 && side_numerator()) return 1;
    if (!rc && DenBackward(*g, logits, dtype, sn, st, N, Tmax, V, len_dev, alpha_ws, aux_ws, grad, sn, st, scale, s, raw)) rc = 1;
    if (sd && sd->trace) {
        cudaEventRecord(sd->tr[5], s);
        cudaDeviceSynchronize();
        float a0 = 0, a1 = 0, f0 = 0, f1 = 0, b1 = 0;
        cudaEventElapsedTime(&a0, sd->tr[0], sd->tr[1]); cudaEventElapsedTime(&a1, sd->tr[0], sd->tr[2]);
        cudaEventElapsedTime(&f0, sd->tr[0], sd->tr[3]); cudaEventElapsedTime(&f1, sd->tr[0], sd->tr[4]);
        cudaEventElapsedTime(&b1, sd->tr[0], sd->tr[5]);
        fprintf(stderr, "ccb overlap trace (ms from fork): numerator %.3f..%.3f  den forward %.3f..%.3f  den backward ..%.3f\n", a0, a1, f0, f1, b1);
    }
    if (sd) {   // join even on failure: the side stream must never outlive the call's buffers unordered
        const std::string keep = g_err;
        if (cudaStreamWaitEvent(s, sd->join, 0) != cudaSuccess && !rc) return FailCuda("cudaStreamWaitEvent(join)", cudaGetLastError());
        g_err = keep;
    }
    if (rc) return 1;
    if (!sd) {
        rc = LaunchCtcAlphaBeta(logits, dtype == CCB_DTYPE_BF16, sn, st, N, Tmax, V, labels_dev, label_off_dev, label_len_dev, len_dev,
                                max_label_len, 0, reinterpret_cast<float *>(ctc_ws), true, logp, lnorm, s, &err, false);
        if (rc) return Fail(err);
    }
    rc = LaunchCtcGamma(logits, dtype == CCB_DTYPE_BF16, sn, st, N, Tmax, V, labels_dev, label_off_dev, label_len_dev, len_dev,
                        max_label_len, 0, reinterpret_cast<float *>(ctc_ws), grad, sn, st, -(1.f + lamb) * scale, s, &err);
    if (rc) return Fail(err);
    if (raw) {   // chain through log_softmax: dL/dz = g - softmax(z) * sum_k g_k
        rc = LaunchLogitGrad(logits, dtype == CCB_DTYPE_BF16, sn, st, N, Tmax, V, len_dev,
                             reinterpret_cast<const float *>((char *)aux_ws + L.lz), L.Npad, grad, sn, st, s);
        if (rc) return FailCuda("logit_grad", (cudaError_t)rc);
    }
    rc = LaunchAssembleLoss(logz, logp, N, lamb, scale, loss, s);
    if (rc) return FailCuda("assemble_loss", (cudaError_t)rc);
    if (parts) {
        CCB_CUDA(cudaMemcpyAsync(parts, logz, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
        CCB_CUDA(cudaMemcpyAsync(parts + N, logp, sizeof(float) * N, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

int ccb_ctc_crf_loss_fwd(const void *logits, int dtype, int N, int T, int V, int Tmax,
                         const int *labels_dev, const int *label_off_dev, const int *label_len_dev,
                         const int *len_dev, int max_label_len, float lamb, float scale,
                         float *alpha_ws, void *aux_ws, void *ctc_ws, float *grad, float *loss, float *parts,
                         void *stream) {
    return LossFwdImpl(false, logits, dtype, N, T, V, Tmax, labels_dev, label_off_dev, label_len_dev, len_dev, max_label_len,
                       lamb, scale, alpha_ws, aux_ws, ctc_ws, grad, loss, parts, stream);
}

int ccb_ctc_crf_loss_logits_fwd(const void *logits, int dtype, int N, int T, int V, int Tmax,
                                const int *labels_dev, const int *label_off_dev, const int *label_len_dev,
                                const int *len_dev, int max_label_len, float lamb, float scale,
                                float *alpha_ws, void *aux_ws, void *ctc_ws, float *grad, float *loss, float *parts,
                                void *stream) {
    return LossFwdImpl(true, logits, dtype, N, T, V, Tmax, labels_dev, label_off_dev, label_len_dev, len_dev, max_label_len,
                       lamb, scale, alpha_ws, aux_ws, ctc_ws, grad, loss, parts, stream);
}

/* ---- gpu_ctc/ctc.h surface ------------------------------------------------------------------- */
const char *ctcGetStatusString(ctcStatus_t status) {
    switch (status) {
        case CTC_STATUS_SUCCESS: return "no error";
        case CTC_STATUS_MEMOPS_FAILED: return "cuda memcpy or memset failed";
        case CTC_STATUS_INVALID_VALUE: return "invalid value";
        case CTC_STATUS_EXECUTION_FAILED: return "execution failed";
        default: return "unknown error";
    }
}

// workspace: [meta ints: labels | label_off | label_len | len] then the alpha spill
ctcStatus_t get_workspace_size(const int *const label_lengths, const int *const input_lengths, int alphabet_size,
                               int minibatch, struct ctcOptions options, size_t *size_bytes) {
    (void)options;
    if (!label_lengths || !input_lengths || !size_bytes || alphabet_size <= 0 || minibatch <= 0) return CTC_STATUS_INVALID_VALUE;
    int maxL = 0, maxT = 0;
    long sumL = 0;
    for (int i = 0; i < minibatch; ++i) {
        if (label_lengths[i] < 0 || input_lengths[i] < 0) return CTC_STATUS_INVALID_VALUE;
        maxL = std::max(maxL, label_lengths[i]); maxT = std::max(maxT, input_lengths[i]); sumL += label_lengths[i];
    }
    size_t meta = ((size_t)(sumL + 3 * (size_t)minibatch + 1) * sizeof(int) + 255) & ~(size_t)255;
    *size_bytes = meta + ((size_t)minibatch * sizeof(float) + 255 & ~(size_t)255) + ccb_ctc_workspace_bytes(minibatch, maxT, maxL);
    return CTC_STATUS_SUCCESS;
}

ctcStatus_t compute_ctc_loss(const float *const activations, float *gradients, const int *const flat_labels,
                             const int *const label_lengths, const int *const input_lengths, int alphabet_size,
                             int minibatch, float *costs, void *workspace, struct ctcOptions options) {
    if (!activations || !flat_labels || !label_lengths || !input_lengths || !costs || !workspace || alphabet_size <= 0 || minibatch <= 0)
        return CTC_STATUS_INVALID_VALUE;
    int maxL = 0, maxT = 0;
    std::vector<int> meta;
    long sumL = 0;
    for (int i = 0; i < minibatch; ++i) { maxL = std::max(maxL, label_lengths[i]); maxT = std::max(maxT, input_lengths[i]); sumL += label_lengths[i]; }
    meta.reserve((size_t)sumL + 3 * (size_t)minibatch + 1);
    meta.insert(meta.end(), flat_labels, flat_labels + sumL);
    int off = 0;
    for (int i = 0; i < minibatch; ++i) { meta.push_back(off); off += label_lengths[i]; }
    meta.push_back(off);
    meta.insert(meta.end(), label_lengths, label_lengths + minibatch);
    meta.insert(meta.end(), input_lengths, input_lengths + minibatch);
    for (long i = 0; i < sumL; ++i) if (flat_labels[i] < 0 || flat_labels[i] >= alphabet_size) return CTC_STATUS_INVALID_VALUE;
    cudaStream_t s = (cudaStream_t)options.stream;
    char *ws = reinterpret_cast<char *>(workspace);
    const size_t meta_bytes = (meta.size() * sizeof(int) + 255) & ~(size_t)255;
    const size_t cost_bytes = ((size_t)minibatch * sizeof(float) + 255) & ~(size_t)255;
    if (cudaMemcpyAsync(ws, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, s) != cudaSuccess) return CTC_STATUS_MEMOPS_FAILED;
    // the host staging vector must outlive the (pageable => staged synchronously) copy: it does, pageable H2D
    // returns after the source has been consumed.
    const int *d_labels = reinterpret_cast<const int *>(ws);
    const int *d_off = d_labels + sumL;
    const int *d_llen = d_off + minibatch + 1;
    const int *d_len = d_llen + minibatch;
    float *d_costs = reinterpret_cast<float *>(ws + meta_bytes);
    float *alpha_ws = reinterpret_cast<float *>(ws + meta_bytes + cost_bytes);
    std::string err;
    // (T,N,V) layout: element (n,t,k) at t*(N*V) + n*V + k
    int rc = LaunchCtc(activations, 0, (long)alphabet_size, (long)minibatch * alphabet_size, minibatch, maxT, alphabet_size,
                       d_labels, d_off, d_llen, d_len, maxL, options.blank_label, alpha_ws, gradients,
                       (long)alphabet_size, (long)minibatch * alphabet_size, 1.f, d_costs, nullptr, s, &err);
    if (rc) { g_err = err; return CTC_STATUS_EXECUTION_FAILED; }
    if (cudaMemcpyAsync(costs, d_costs, sizeof(float) * minibatch, cudaMemcpyDeviceToHost, s) != cudaSuccess) return CTC_STATUS_MEMOPS_FAILED;
    if (cudaStreamSynchronize(s) != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;   // costs are host memory (gpu_ctc.h:365-368)
    return CTC_STATUS_SUCCESS;
}

/* ---- host-side plan inspection ------------------------------------------------------------------ */
void *ccb_plan_create(const char *fst_name, int n_ctas, int n_warps) {
    g_err.clear();
    HostFst fst;
    std::string err;
    if (!fst_name || !ReadFstFile(fst_name, &fst, &err)) { g_err = fst_name ? err : "null path"; return nullptr; }
    DenPlan *p = new DenPlan();
    if (!BuildDenPlan(fst, n_ctas, n_warps, p, &err)) { g_err = err; delete p; return nullptr; }
    return p;
}

void ccb_plan_destroy(void *plan) { delete reinterpret_cast<DenPlan *>(plan); }

int ccb_plan_info(void *plan, long *info) {
    if (!plan || !info) return 1;
    const DenPlan *p = reinterpret_cast<const DenPlan *>(plan);
    info[0] = p->file_states; info[1] = p->file_arcs; info[2] = p->num_states;
    info[3] = (long)p->fwd.arcs.size(); info[4] = (long)p->bwd.arcs.size(); info[5] = p->start;
    info[6] = p->num_labels; info[7] = p->n_ctas; info[8] = p->n_warps;
    info[9] = std::max(p->fwd.max_tile_arcs, p->bwd.max_tile_arcs);
    info[10] = p->num_pairs; info[11] = (long)p->start_arcs.size(); info[12] = (long)p->hub_states.size();
    return 0;
}

int ccb_plan_copy(void *plan, int which, void *dst, size_t dst_bytes) {
    if (!plan || !dst) return 1;
    const DenPlan *p = reinterpret_cast<const DenPlan *>(plan);
    const void *src = nullptr;
    size_t bytes = 0;
    switch (which) {
        case 0: src = p->state_label.data(); bytes = p->state_label.size() * 4; break;
        case 1: src = p->final_lin.data(); bytes = p->final_lin.size() * 4; break;
        case 2: src = p->orig_state.data(); bytes = p->orig_state.size() * 4; break;
        case 3: src = p->fwd.arcs.data(); bytes = p->fwd.arcs.size() * sizeof(Arc); break;
        case 4: src = p->fwd.chunk_state.data(); bytes = p->fwd.chunk_state.size() * 4; break;
        case 5: src = p->fwd.chunk_arc.data(); bytes = p->fwd.chunk_arc.size() * 4; break;
        case 6: src = p->bwd.arcs.data(); bytes = p->bwd.arcs.size() * sizeof(Arc); break;
        case 7: src = p->bwd.chunk_state.data(); bytes = p->bwd.chunk_state.size() * 4; break;
        case 8: src = p->bwd.chunk_arc.data(); bytes = p->bwd.chunk_arc.size() * 4; break;
        case 9: src = p->state_pos.data(); bytes = p->state_pos.size() * 4; break;
        case 10: src = p->fwd.chunk_pair.data(); bytes = p->fwd.chunk_pair.size() * 4; break;
        case 11: src = p->bwd.chunk_pair.data(); bytes = p->bwd.chunk_pair.size() * 4; break;
        case 12: src = p->start_arcs.data(); bytes = p->start_arcs.size() * sizeof(Arc); break;
        case 13: src = p->fwd.cta_labels.data(); bytes = p->fwd.cta_labels.size() * 4; break;
        case 14: src = p->bwd.cta_labels.data(); bytes = p->bwd.cta_labels.size() * 4; break;
        case 15: src = p->bwd.w1.data(); bytes = p->bwd.w1.size() * 4; break;
        case 16: src = p->hub_states.data(); bytes = p->hub_states.size() * 4; break;
        case 17: src = p->own_fwd.data(); bytes = (size_t)p->num_states * 2 * 4; break;   // (without the 4 floats of padding)
        case 18: src = p->own_bwd.data(); bytes = (size_t)p->num_states * 2 * 4; break;
        default: return 1;
    }
    if (bytes > dst_bytes) return 2;
    memcpy(dst, src, bytes);
    return 0;
}

}  // extern "C"
