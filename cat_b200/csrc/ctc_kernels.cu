// Numerator: per-utterance CTC forward-backward over the blank-expanded label lattice (log domain).
//
// Replaces src/ctc_crf/gpu_ctc (compute_alpha_kernel gpu_ctc_kernels.h:87-213,
// compute_betas_and_grad_kernel :218-458, host driver gpu_ctc.h:100-400).  Differences by design:
//   * one kernel does forward, backward and the gradient (the reference launches two and synchronises
//     the stream twice, gpu_ctc.h:272,368); nothing here touches the host;
//   * thread i owns lattice cells 2i (blank) and 2i+1 (label i): the 2-/3-way log-sum of a frame needs
//     one value from the neighbouring thread, exchanged through a shared-memory ping-pong, one
//     __syncthreads per frame; the label row y[n][t][:] is staged coalesced one frame ahead;
//   * occupancies are reduced per label with shared-memory float atomics in the LINEAR domain (blank cells
//     are first summed with warp shuffles) -- no sort, no segmented reduce, no moderngpu;
//   * any layout: element (n,t,k) at n*sn + t*st + k, so the (N,T,V) logits are read in place (the
//     reference needs a transposed (T,N,V) copy, ctc_crf/__init__.py:70) and the gradient is ACCUMULATED
//     with a caller scale into a buffer that may already hold the denominator part;
//   * no 2L+1 <= 1280 limit (gpu_ctc.h:294-311): cells beyond the thread count are looped.
// Semantics kept: inputs are log-softmax outputs (gpu_ctc/README.txt:1-2); costs = log p(l|x);
// grad[t][k] = exp(LSE_{s: l'_s=k}(alpha_t(s)+beta_t(s)) - y_t(k) - log p) written for labels occurring in the
// utterance and t < len only (:431-435); infeasible utterances (L+repeats > T, :108-109) get
// log p = -inf and no gradient (the reference leaves both unwritten).
#include "common.cuh"

namespace ccb {

namespace {

constexpr unsigned kFull = 0xffffffffu;

// Block-wide max of the values the warps published for the previous frame (one LDS + 5 shuffles).
__device__ __forceinline__ float block_max_from(const float *s_wmax, int nwarps) {
    const int lane = threadIdx.x & 31;
    float m = lane < nwarps ? s_wmax[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    return m == -INFINITY ? 0.f : m;
}

// Precision: cells are kept RELATIVE to a per-frame offset (the block max of the previous frame, accumulated in
// fp64), so fp32 log-add rounding stays ~1e-6 absolute instead of growing with |alpha| (the reference's plain fp32
// log domain loses ~1e-2 relative on the occupancies at T ~ 1000).  alpha_true_t(s) = a_rel_t(s) + C_t,
// beta_true_t(s) = b_rel_t(s) + D_t, occupancy = exp(a_rel + b_rel - y + (C_t + D_t - log p)).
// shared memory carve-up: lab[Lmax+1] ints | a[2][ScMax] | yrow[2][V] | gk[2][V] | wmax[2][32]
__global__ void ctc_fwd_bwd_kernel(const void *y, int y_bf16, long sn, long st, int T, int V,
                                   const int *labels, const int *label_off, const int *label_len, const int *len,
                                   int max_label_len, int blank, float *alpha_ws,
                                   float *grad, long gsn, long gst, float grad_scale, float *logp_out,
                                   const double *lnorm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = blockIdx.x;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int L = label_len[n], Tn = len[n];
    const int Sc = 2 * L + 1, L1 = L + 1;
    const int ScMax = 2 * max_label_len + 1;
    int *s_lab = reinterpret_cast<int *>(smem_raw);                 // [max_label_len + 1] label of cell 2i+1
    float *s_a = reinterpret_cast<float *>(s_lab + max_label_len + 1);   // [2][ScMax]
    float *s_y = s_a + 2 * ScMax;                                   // [2][V]
    float *s_g = s_y + 2 * V;                                       // [2][V]
    float *s_wmax = s_g + 2 * V;                                    // [2][32]
    const int *lab = labels + label_off[n];

    // feasibility (gpu_ctc_kernels.h:108-109)
    int rep = 0;
    for (int i = tid + 1; i < L; i += NT) rep += (lab[i] == lab[i - 1]);
    __shared__ int s_rep;
    if (tid == 0) s_rep = 0;
    __syncthreads();
    if (rep) atomicAdd(&s_rep, rep);
    __syncthreads();
    if (Tn <= 0 || L + s_rep > Tn) {
        if (tid == 0) logp_out[n] = -INFINITY;
        return;
    }
    for (int i = tid; i < L; i += NT) s_lab[i] = lab[i];
    for (int k = tid; k < 2 * V; k += NT) s_g[k] = 0.f;
    // per-utterance workspace: alpha_rel [T][ScMax] floats, then C_t [T] doubles
    const size_t per_utt = ((size_t)T * ScMax + 1) / 2 * 2 + 2 * (size_t)T;
    float *ws = alpha_ws + (size_t)n * per_utt;
    double *coff = reinterpret_cast<double *>(ws + ((size_t)T * ScMax + 1) / 2 * 2);
    const long ybase = n * sn;

    // ---- forward ---------------------------------------------------------------------------------
    for (int k = tid; k < V; k += NT) s_y[k] = load_y(y, y_bf16, ybase + k);
    __syncthreads();
    {
        float mx = -INFINITY;
        for (int s = tid; s < Sc; s += NT) {
            float v = -INFINITY;
            if (s == 0) v = s_y[blank];
            else if (s == 1) v = s_y[s_lab[0]];
            s_a[s] = v;
            ws[s] = v;
            mx = fmaxf(mx, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
        if (lane == 0) s_wmax[warp] = mx;
        if (tid == 0) coff[0] = 0.0;
    }
    double C = 0.0;
    // Row prefetch: the emission row of frame t+1 is loaded into registers while frame t is computed and parked in
    // shared memory afterwards, so no frame waits on a global-memory round trip (V <= kRowRegs * blockDim).
    constexpr int kRowRegs = 4;
    const bool row_in_regs = V <= kRowRegs * NT;
    float yreg[kRowRegs];
    if (row_in_regs && Tn > 1) {
#pragma unroll
        for (int j = 0; j < kRowRegs; ++j) { const int k = tid + j * NT; yreg[j] = k < V ? load_y(y, y_bf16, ybase + st + k) : 0.f; }
#pragma unroll
        for (int j = 0; j < kRowRegs; ++j) { const int k = tid + j * NT; if (k < V) s_y[V + k] = yreg[j]; }
    }
    for (int t = 1; t < Tn; ++t) {
        // One barrier per frame: it publishes row t (parked in slot t&1 during the previous frame), the previous
        // frame's cells and their per-warp maxima.
        float *yc = s_y + (t & 1) * V;
        const float *prev = s_a + ((t - 1) & 1) * ScMax;
        float *cur = s_a + (t & 1) * ScMax;
        if (!row_in_regs) for (int k = tid; k < V; k += NT) yc[k] = load_y(y, y_bf16, ybase + (long)t * st + k);
        else if (t + 1 < Tn) {
#pragma unroll
            for (int j = 0; j < kRowRegs; ++j) { const int k = tid + j * NT; yreg[j] = k < V ? load_y(y, y_bf16, ybase + (long)(t + 1) * st + k) : 0.f; }
        }
        __syncthreads();
        const float m = block_max_from(s_wmax + ((t - 1) & 1) * 32, nwarps);
        C += (double)m;
        if (tid == 0) coff[t] = C;
        float mx = -INFINITY;
        for (int i = tid; i < L1; i += NT) {
            // blank cell 2i
            const int sb = 2 * i;
            float vb = prev[sb];
            if (i > 0) vb = log_add(vb, prev[sb - 1]);
            vb += yc[blank] - m;
            cur[sb] = vb;
            ws[(size_t)t * ScMax + sb] = vb;
            mx = fmaxf(mx, vb);
            if (i < L) {   // label cell 2i+1
                const int sl = sb + 1;
                const int li = s_lab[i];
                float vl = log_add(prev[sl], prev[sb]);
                if (i > 0 && li != s_lab[i - 1]) vl = log_add(vl, prev[sl - 2]);
                vl += yc[li] - m;
                cur[sl] = vl;
                ws[(size_t)t * ScMax + sl] = vl;
                mx = fmaxf(mx, vl);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
        if (lane == 0) s_wmax[(t & 1) * 32 + warp] = mx;
        if (row_in_regs && t + 1 < Tn) {   // park row t+1 in the slot last read during frame t-1
            float *yn = s_y + ((t + 1) & 1) * V;
#pragma unroll
            for (int j = 0; j < kRowRegs; ++j) { const int k = tid + j * NT; if (k < V) yn[k] = yreg[j]; }
        }
    }
    __syncthreads();
    double logp_d;
    {
        const float *last = s_a + ((Tn - 1) & 1) * ScMax;
        float lp = last[Sc - 1];
        if (Sc > 1) lp = log_add(lp, last[Sc - 2]);
        logp_d = (double)lp + C;
    }
    // raw-logit entry: the reported log-likelihood is normalised; everything below stays in the raw domain
    if (tid == 0) logp_out[n] = (float)(logp_d - (lnorm ? lnorm[n] : 0.0));
    if (grad == nullptr || !(logp_d > -INFINITY)) return;
    __syncthreads();

    // ---- backward + occupancies ------------------------------------------------------------------
    // beta ping-pong reuses s_a (slot t&1 holds beta_rel_t) and s_wmax.
    double D = 0.0;
    // prefetch state: emission row and this thread's two alpha cells of the NEXT frame to be processed (t-1)
    const bool cells_in_regs = L1 <= NT;
    float a_sb = -INFINITY, a_sl = -INFINITY;   // alpha_rel of cells 2*tid, 2*tid+1 for the current frame
    {
        const int t = Tn - 1;
        float *yc = s_y + (t & 1) * V;
        for (int k = tid; k < V; k += NT) yc[k] = load_y(y, y_bf16, ybase + (long)t * st + k);
        if (cells_in_regs && tid < L1) {
            a_sb = ws[(size_t)t * ScMax + 2 * tid];
            if (tid < L) a_sl = ws[(size_t)t * ScMax + 2 * tid + 1];
        }
    }
    for (int t = Tn - 1; t >= 0; --t) {
        float *yc = s_y + (t & 1) * V;
        float *cur = s_a + (t & 1) * ScMax;
        const float *nxt = s_a + ((t + 1) & 1) * ScMax;
        float *gk = s_g + (t & 1) * V;
        // issue the loads for frame t-1 now; they are consumed after this frame's work
        float an_sb = -INFINITY, an_sl = -INFINITY;
        if (t > 0) {
            if (row_in_regs) {
#pragma unroll
                for (int j = 0; j < kRowRegs; ++j) { const int k = tid + j * NT; yreg[j] = k < V ? load_y(y, y_bf16, ybase + (long)(t - 1) * st + k) : 0.f; }
            }
            if (cells_in_regs && tid < L1) {
                an_sb = ws[(size_t)(t - 1) * ScMax + 2 * tid];
                if (tid < L) an_sl = ws[(size_t)(t - 1) * ScMax + 2 * tid + 1];
            }
        }
        const double Ct = coff[t];
        __syncthreads();   // the only barrier of the frame: publishes row t, beta_{t+1}, its maxima and gk(t+1)
        if (t + 1 < Tn) {  // flush the (now complete) occupancies of frame t+1 and clear their slot
            float *gp = s_g + ((t + 1) & 1) * V;
            float *grow = grad + n * gsn + (long)(t + 1) * gst;
            for (int k = tid; k < V; k += NT) {
                const float g = gp[k];
                if (g != 0.f) { atomicAdd(grow + k, grad_scale * g); gp[k] = 0.f; }
            }
        }
        float m = 0.f;
        if (t < Tn - 1) {
            m = block_max_from(s_wmax + ((t + 1) & 1) * 32, nwarps);
            D += (double)m;
        }
        const float K = (float)(Ct + D - logp_d);
        const float *al = ws + (size_t)t * ScMax;
        float mx = -INFINITY;
        for (int i0 = 0; i0 < L1; i0 += NT) {
            const int i = i0 + tid;
            float occ_blank = 0.f;
            if (i < L1) {
                const int sb = 2 * i;
                float vb;
                if (t == Tn - 1) {
                    vb = (sb == Sc - 1) ? yc[blank] : -INFINITY;
                } else {
                    vb = nxt[sb];
                    if (sb + 1 < Sc) vb = log_add(vb, nxt[sb + 1]);
                    vb += yc[blank] - m;
                }
                cur[sb] = vb;
                mx = fmaxf(mx, vb);
                const float ob = (cells_in_regs ? a_sb : al[sb]) + vb - yc[blank] + K;
                occ_blank = (ob == -INFINITY || ob != ob) ? 0.f : expf(ob);
                if (i < L) {
                    const int sl = sb + 1;
                    const int li = s_lab[i];
                    float vl;
                    if (t == Tn - 1) {
                        vl = (sl == Sc - 2) ? yc[li] : -INFINITY;
                    } else {
                        vl = log_add(nxt[sl], nxt[sl + 1]);
                        if (i + 1 < L && li != s_lab[i + 1]) vl = log_add(vl, nxt[sl + 2]);
                        vl += yc[li] - m;
                    }
                    cur[sl] = vl;
                    mx = fmaxf(mx, vl);
                    const float ol = (cells_in_regs ? a_sl : al[sl]) + vl - yc[li] + K;
                    if (ol != -INFINITY && ol == ol) atomicAdd(&gk[li], expf(ol));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) occ_blank += __shfl_xor_sync(kFull, occ_blank, o);
            if (lane == 0 && occ_blank != 0.f) atomicAdd(&gk[blank], occ_blank);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
        if (lane == 0) s_wmax[(t & 1) * 32 + warp] = mx;
        if (t > 0) {   // park row t-1 (slot last read during frame t+1) and rotate the alpha cells
            float *yn = s_y + ((t - 1) & 1) * V;
            if (row_in_regs) {
#pragma unroll
                for (int j = 0; j < kRowRegs; ++j) { const int k = tid + j * NT; if (k < V) yn[k] = yreg[j]; }
            } else {
                for (int k = tid; k < V; k += NT) yn[k] = load_y(y, y_bf16, ybase + (long)(t - 1) * st + k);
            }
            a_sb = an_sb; a_sl = an_sl;
        }
    }
    __syncthreads();
    {   // flush frame 0
        float *gp = s_g;
        float *grow = grad + n * gsn;
        for (int k = tid; k < V; k += NT) {
            const float g = gp[k];
            if (g != 0.f) atomicAdd(grow + k, grad_scale * g);
        }
    }
}

// loss[0] = scale * sum_n (logz[n] - (1+lamb) logp[n])      (ctc_crf/__init__.py:79-86)
__global__ void assemble_loss_kernel(const float *logz, const float *logp, int N, float lamb, float scale, float *loss) {
    double acc = 0.0;
    for (int n = threadIdx.x; n < N; n += 32) acc += (double)logz[n] - (1.0 + (double)lamb) * (double)logp[n];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
    if (threadIdx.x == 0) loss[0] = (float)(acc * (double)scale);
}

}  // namespace

int LaunchCtc(const void *y, int y_bf16, long sn, long st, int N, int T, int V, const int *labels,
              const int *label_off, const int *label_len, const int *len, int max_label_len, int blank,
              float *alpha_ws, float *grad, long gsn, long gst, float grad_scale, float *logp,
              const double *lnorm, cudaStream_t stream, std::string *err) {
    if (N == 0) return 0;
    const int ScMax = 2 * max_label_len + 1;
    const size_t smem = (size_t)(max_label_len + 1) * 4 + (size_t)2 * ScMax * 4 + (size_t)4 * V * 4 + 64 * 4;
    int threads = ((max_label_len + 1 + 31) / 32) * 32;
    threads = threads < 64 ? 64 : (threads > 1024 ? 1024 : threads);
    if (smem > 200 * 1024) { *err = "label sequence too long for the numerator kernel's shared memory"; return 1; }
    cudaError_t e = cudaFuncSetAttribute(ctc_fwd_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { *err = std::string("cudaFuncSetAttribute(ctc): ") + cudaGetErrorString(e); return (int)e; }
    ctc_fwd_bwd_kernel<<<N, threads, smem, stream>>>(y, y_bf16, sn, st, T, V, labels, label_off, label_len, len,
                                                     max_label_len, blank, alpha_ws, grad, gsn, gst, grad_scale, logp, lnorm);
    CountLaunch();
    e = cudaGetLastError();
    if (e != cudaSuccess) { *err = std::string("ctc launch: ") + cudaGetErrorString(e); return (int)e; }
    return 0;
}

int LaunchAssembleLoss(const float *logz, const float *logp, int N, float lamb, float scale, float *loss,
                       cudaStream_t stream) {
    assemble_loss_kernel<<<1, 32, 0, stream>>>(logz, logp, N, lamb, scale, loss);
    CountLaunch();
    return (int)cudaGetLastError();
}

}  // namespace ccb
