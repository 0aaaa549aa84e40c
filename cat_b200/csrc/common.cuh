// Shared device helpers and the internal host-side launch interface (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "den_graph.h"

namespace ccb {

// ------------------------------------------------------------------------------------------------
// Device copy of the den plan (one per GPU listed in Init()).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kOobRow = 0x40000000u;   // a TMA row coordinate no gather table reaches (tables are limited to 2^30 rows)
// backward arc bytes per CTA tile (12 per slot) above which a graph also gets the streamed-arc copy of its quads: below it
// the stream fits shared memory next to the deepest rings at every batch width
constexpr size_t kStreamTierArcBytes = 72 * 1024;

struct DevicePass {
    Arc *arcs = nullptr;
    int *chunk_state = nullptr;
    int *chunk_arc = nullptr;
    int *chunk_pair = nullptr;
    int *cta_labels = nullptr;
    float *w1 = nullptr;      // backward pass: second weight per slot
    float *own_c = nullptr;   // [2S + 4] own-row coefficients of this pass (DenPlan::own_fwd / own_bwd)
    // Streamed-arc tier (backward pass of graphs whose arc stream does not fit shared memory): the pass's quads as the TMA
    // kernel stages them, 3 16-byte words per quad {row coordinate 0..3}{w0 0..3}{w1 0..3}; built at Init when the graph is
    // large enough to need it, else null
    uint4 *tq = nullptr;
    int num_arcs = 0;
    int max_tile_arcs = 0;
    int max_tile_labels = 0;
    int max_tile_rows = 0;
};

struct DeviceGraph {
    bool loaded = false;
    int device = -1;
    int S = 0, P = 0, start = 0, num_labels = 0;   // S states, P pairs (gather-table rows S..S+P-1 = virtual pair-sum rows)
    int scale_exp = 56;
    int n_ctas = 0, n_warps = 0;
    int *state_label = nullptr;
    int *state_pos = nullptr;
    float *final_lin = nullptr;
    DevicePass fwd, bwd;
    Arc *start_arcs = nullptr;   // out-arcs of the start state
    int n_start_arcs = 0;
    int *hub_states = nullptr;   // states whose forward row is accumulated from parts
    int n_hubs = 0;
    float start_final = 0.f;
    int max_smem_optin = 0;
    // test hooks, read ONCE at Init (never on the per-call path): force the large-graph tiers on a small graph
    bool tune_arcs_in_global = false, tune_w1_in_global = false, tune_no_tma = false;
    int tune_ring_rows = 0;    // > 0: force this many rows per TMA ring stage (A/B runs)
    // batches of <= 16 utterances run the small-batch TMA kernels (rows of 8 / 16 floats): needs both arc streams in
    // shared memory next to the rings, no hub rows, and a usable TMA descriptor -- decided once at Init
    bool small_ok = false;
    bool own_rows = false;     // DenPlan::own_rows: own-row coefficients instead of own-row arcs (DevicePass::own_c)
};

// Kernel parameter block shared by the two persistent den kernels (passed as a __grid_constant__ kernel parameter: the TMA
// descriptor inside it is read by the hardware straight from parameter memory and must sit on a 64-byte boundary).
struct alignas(64) DenParams {
    // TMA view of the pass's gather table (forward: the alpha spill, backward: the beta ping-pong) as a 2-D tensor
    // [rows][Npad] of fp32 with a box of one row x (32 * utterances per lane) columns; used by the gather4 kernels only
    CUtensorMap tmap;
    int use_tma;          // 1: rows are staged in shared memory by TMA gather4 (den_kernels.cu), 0: register gathers
    int ring_off;         // byte offset of the per-warp row rings in dynamic shared memory (128-byte aligned)
    int bar_off;          // byte offset of the rings' mbarriers
    // graph
    const Arc *arcs;
    const float *w1;      // backward pass only
    const int *chunk_state;
    const int *chunk_arc;
    const int *chunk_pair;
    const int *cta_labels;
    const int *state_label;
    const int *state_pos;
    const uint4 *tq;          // streamed-arc tier: the pass's transposed quads in global memory (DevicePass::tq)
    int aring_off, abar_off;  // ... byte offsets in dynamic shared memory of the per-warp arc rings and their mbarriers
    const float *final_lin;
    const Arc *start_arcs;
    int n_start_arcs;
    const int *hub_states;
    int n_hubs;
    int S, num_pairs, start, n_warps;
    int scale_exp;        // column sums are renormalised to ~2^scale_exp (DenPlan::scale_exp)
    float start_final;
    // problem
    const void *y;        // (N,T,V) log-probs, fp32 or bf16
    int y_bf16;
    long sn, st;          // element strides of y
    int N, Npad, Tmax, V;
    const int *len;       // [N] device
    // workspaces
    float *alpha;         // [(Tmax+3)][S][Npad] scaled-linear alpha spill; the pair-sum rows of frame t are parked in frame t+2
    float *bh;            // [2][S][Npad]          backward ping-pong (emission-weighted beta)
    float *colsum_a;      // [(Tmax+2)][Npad]
    float *colsum_b;      // [(Tmax+2)][Npad]
    float *absum;         // [(Tmax+2)][Npad]      sum_q alpha_t(q) beta_t(q)
    float *zsum;          // [Npad]
    float *b0;            // [Npad]
    const float *fmax;    // [Tmax][Npad]          per-frame max of y (emission shift)
    unsigned *barrier;    // grid barrier counter (zeroed before launch)
    float *logz;          // [N] out (forward: logZ from alpha; backward: logZ from beta)
    const float *own_c;   // [2S + 4] own-row coefficients of this pass (den_graph.h DenPlan::own_rows), used when own != 0
    int own;
    const double *lnorm;  // [N] or null: sum_t log-normaliser of raw logits, subtracted from logZ (raw-logit entry)
    // gradient (backward)
    float *grad;          // raw accumulation target, element (n,t,k) at n*gsn + t*gst + k
    long gsn, gst;
    int gacc_rows;        // rows of the shared-memory label accumulator (0 = direct global atomics)
    int tile_rows;        // max rows (states) owned by one CTA: size of the shared-memory row metadata
    // optional per-warp timeline (profiling aid, normally null)
    unsigned long long *timeline;
    int tl_step0, tl_steps;
    int debug;            // CCB_DEBUG bit 0: skip row-end work, bit 1: skip the grid barrier (TIMING EXPERIMENTS ONLY)
};

// workspace carving (all offsets in bytes, 256-aligned)
struct DenAuxLayout {
    int Npad = 0;
    size_t colsum_a = 0, colsum_b = 0, absum = 0, zsum = 0, b0 = 0, barrier = 0, zero_bytes = 0;
    size_t fmax = 0, bh = 0, logz_a = 0, logz_b = 0, lz = 0, lnorm = 0, total = 0;
};
// Lane padding of a batch: whole 32-lane groups carrying 1, 2 or 4 utterances per lane; batches of <= 16 utterances
// on a graph that supports the small-batch kernels (DeviceGraph::small_ok) use 8- or 16-float rows instead, with the
// lanes of a warp spread over (arc of the quad) x (utterance)  (den_kernels.cu, LPR template parameter).
inline int PadLanes(int N, bool small_ok) {
    // (8-float rows were measured slower than 16-float rows with half the lanes idle -- N=8, T=600: 11.1 + 14.2 ms against
    // 8.1 + 8.7 ms -- because the gather4 copies are bound by their count, not by their bytes: profiles/r02_experiments.md)
    if (small_ok && N <= 16) return 16;
    int g = (N + 31) / 32;
    if (g >= 3) g = (g + 3) / 4 * 4;   // lanes carry 1, 2 or 4 utterances each
    return g * 32;
}
DenAuxLayout MakeDenAuxLayout(int S, int N, int T, bool small_ok);
inline int LaneWidth(int Npad) { int g = Npad / 32; return g >= 4 ? 4 : (g < 1 ? 1 : g); }

// host launchers; return cudaError_t-compatible int (0 = ok) and fill *err
int LaunchFrameMax(const void *y, int y_bf16, long sn, long st, int N, int T, int V, const int *len, float *fmax,
                   int Npad, cudaStream_t stream);
int LaunchFrameLse(const void *y, int y_bf16, long sn, long st, int N, int T, int V, const int *len, float *fmax,
                   float *lz, double *lnorm, int Npad, cudaStream_t stream);
int LaunchLogitGrad(const void *z, int z_bf16, long sn, long st, int N, int T, int V, const int *len, const float *lz,
                    int Npad, float *grad, long gsn, long gst, cudaStream_t stream);
int LaunchDenForward(const DeviceGraph &g, DenParams &p, cudaStream_t stream, std::string *err);
int LaunchDenBackward(const DeviceGraph &g, DenParams &p, cudaStream_t stream, std::string *err);
// fills *tm with the 2-D view [rows][Npad] fp32 of `base`, box = 1 row x box_cols columns; false if the driver entry
// point is unavailable or refuses the shape (the callers then keep the register-gather kernels)
bool EncodeRowTensorMap(CUtensorMap *tm, const float *base, size_t rows, int Npad, int box_cols);
int LaunchDenGradNormalize(float *grad, long gsn, long gst, const float *absum, const int *len, int N, int Npad,
                           int T, int V, float scale, cudaStream_t stream);
int LaunchCtc(const void *y, int y_bf16, long sn, long st, int N, int T, int V, const int *labels,
              const int *label_off, const int *label_len, const int *len, int max_label_len, int blank,
              float *alpha_ws, float *grad, long gsn, long gst, float grad_scale, float *logp,
              const double *lnorm, cudaStream_t stream, std::string *err, int overwrite = 0, int Tfull = 0);
int LaunchCtcAlphaBeta(const void *y, int y_bf16, long sn, long st, int N, int T, int V, const int *labels,
                       const int *label_off, const int *label_len, const int *len, int max_label_len, int blank,
                       float *alpha_ws, bool want_beta, float *logp, const double *lnorm, cudaStream_t stream,
                       std::string *err, bool share_sm);
int LaunchCtcGamma(const void *y, int y_bf16, long sn, long st, int N, int T, int V, const int *labels,
                   const int *label_off, const int *label_len, const int *len, int max_label_len, int blank,
                   float *alpha_ws, float *grad, long gsn, long gst, float grad_scale, cudaStream_t stream,
                   std::string *err, int overwrite = 0, int Tfull = 0);
int LaunchSumScale(const float *logp, int N, float scale, float *loss, cudaStream_t stream);
int LaunchCtcViterbi(const void *y, int y_bf16, long sn, long st, int N, int T, int V, const int *labels,
                     const int *label_off, const int *label_len, const int *len, int max_label_len, int blank,
                     unsigned char *bp_ws, int *align, float *score, cudaStream_t stream, std::string *err);
int LaunchAssembleLoss(const float *logz, const float *logp, int N, float lamb, float scale, float *loss,
                       cudaStream_t stream);
void CountLaunch(int n = 1);

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float load_y(const void *y, int bf16, long idx) {
    return bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(y)[idx])
                : __ldg(reinterpret_cast<const float *>(y) + idx);
}

__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Grid-wide barrier for a co-resident (cooperatively launched) grid.  `target` is the value the
// monotonically increasing counter reaches once every CTA has arrived at this barrier.
// Arrive with a release reduction, poll with relaxed loads (an acquire load would invalidate L1 on every poll),
// then one acquire fence before the CTA is released.
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        red_release_add_u32(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// The same barrier in two halves: work placed between them (by any thread of the CTA) overlaps the barrier latency.
// Only writes made BEFORE the arrive are guaranteed visible to the other CTAs after their wait; writes made in between
// are ordered by the NEXT barrier's release (use them for data nobody reads in the next frame).
__device__ __forceinline__ void grid_barrier_arrive(unsigned *counter) {
    __syncthreads();
    if (threadIdx.x == 0) red_release_add_u32(counter, 1u);
}
__device__ __forceinline__ void grid_barrier_wait(unsigned *counter, unsigned target) {
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// log-semiring add in fp32 (numerator only), same formula as den_calculate.cu:28-35 / ctc_helper.h.  The operands
// are kept relative to a per-frame offset (|values| stay small), so the hardware ex2/lg2 approximations
// (abs. error ~1e-7 here) are as accurate as the libm versions were on the reference's un-normalised values.
__device__ __forceinline__ float log_add(float a, float b) {
    const float m = fmaxf(a, b);
    if (m == -INFINITY) return m;
    return m + __logf(1.f + __expf(-fabsf(a - b)));   // exp(-inf) = 0 covers a one-sided -inf
}
#endif

}  // namespace ccb
