/*
 * ctc_crf_b200 -- C ABI of the B200-native CTC-CRF loss hot path (drop-in for thu-spmi/CAT src/ctc_crf).
 *
 * Plain C, no torch types: pointers, sizes, a cudaStream_t passed as void*.  All device pointers must
 * belong to the CUDA device that is current when the call is made (the reference has the same contract,
 * den_calculate.cu:436-438).  Citations are relative to /root/reference/src/ctc_crf.
 *
 * Section 1 keeps the exact names and signatures the reference's binding.cpp binds (binding.cpp:14-49 and
 * gpu_ctc/ctc.h:16-109), so the reference's own binding.cpp links against this library unchanged.
 * Section 2 is the fused entry the new autograd Function uses (SURVEY.md 8b "new fused entry").
 * Section 3 exposes the host-side den-graph plan for CPU-only tests.
 *
 * Error convention: section-1 den functions return void like the reference but never exit(); they record
 * an error retrievable with ccb_last_error().  Everything else returns 0 on success, non-zero on failure.
 */
#ifndef CTC_CRF_B200_H_
#define CTC_CRF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Section 1 -- reference-compatible entry points
 * ---------------------------------------------------------------------------------------------- */

/* replaces den_calculate.cu:263-273 globals read by binding.cpp:14-15,77-78 (scratch sizing).
 * DEN_NUM_STATES is the state count *after* the loader's in-label state split (== the file's state
 * count for every T-compose-LM graph). */
extern int DEN_NUM_ARCS;
extern int DEN_NUM_STATES;

/* replaces Init / Release (binding.cpp:22-24, den_calculate.cu:288-425): load an OpenFst binary
 * vector/standard den graph, build the kernel plan, upload it to each listed GPU. */
void Init(const char *fst_name, int n_gpus, int *gpus);
void Release(int n_gpus, int *gpus);

/* replaces compute_alpha (binding.cpp:26-34, den_calculate.cu:427-451).
 *   alpha      : scratch, >= ccb_den_alpha_floats(batch_size, T) floats (OPAQUE layout: [t][state][lane]).
 *                The need is (T+3) frames of DEN_NUM_STATES-P rows of PadLanes(N) floats (lanes padded to 32, 64 or a multiple
 *                of 128; 16 for batches of <= 16 utterances on small graphs): binding.cpp's (T+1)*N*DEN_NUM_STATES covers it
 *                for N = 32, 64, 128, 256 ... but NOT for every other N (e.g. 96 -> 128 lanes).  When the caller's buffer is
 *                too small the library takes its own from the device's default memory pool, stream-ordered
 *                (cudaMallocAsync on `stream`: no device synchronisation), keeps it between calls and returns it in Release().
 *                Size the buffer with ccb_den_alpha_floats() to avoid the second allocation.
 *   logits     : (N,T,V) fp32 log-probs, contiguous
 *   input_lengths : (N,) int32 DEVICE
 *   loglikelihood : (N,) fp32 DEVICE out, logZ_den per utterance */
void compute_alpha(float *alpha, float *logits, const int batch_size, int T, const int alpha_size,
                   int logits_size, int *input_lengths, float *loglikelihood, void *stream);

/* replaces compute_beta_and_grad (binding.cpp:36-48, den_calculate.cu:453-481).
 *   grad_net : (N,T,V) fp32, PRE-ZEROED by the caller (ctc_crf/__init__.py:66); rows t < len get the
 *              denominator occupancies, rows t >= len stay untouched.
 *   beta / grad_storage : scratch of the reference's sizes; unused here (kept for signature parity).
 *   loglikelihood : (N,) out, logZ_den recomputed from the backward pass (== alpha_lld up to rounding). */
void compute_beta_and_grad(float *beta, const float *const alpha, const float *const logits,
                           const float *const alpha_lld, float *grad_storage, float *grad_net,
                           const int batch_size, const int T, const int beta_size, const int logits_size,
                           const int *const input_lengths, float *loglikelihood, void *stream);

/* replaces gpu_ctc/ctc.h:16-109 */
typedef enum {
    CTC_STATUS_SUCCESS = 0,
    CTC_STATUS_MEMOPS_FAILED = 1,
    CTC_STATUS_INVALID_VALUE = 2,
    CTC_STATUS_EXECUTION_FAILED = 3,
    CTC_STATUS_UNKNOWN_ERROR = 4
} ctcStatus_t;

struct ctcOptions {
    void *stream;     /* CUstream */
    int blank_label;
};

const char *ctcGetStatusString(ctcStatus_t status);

/* activations: (T,N,V) fp32 log-probs on the device; gradients: (T,N,V) pre-zeroed device buffer or NULL;
 * flat_labels / label_lengths / input_lengths / costs: HOST pointers; costs[n] = log p(l_n | x_n). */
ctcStatus_t compute_ctc_loss(const float *const activations, float *gradients, const int *const flat_labels,
                             const int *const label_lengths, const int *const input_lengths, int alphabet_size,
                             int minibatch, float *costs, void *workspace, struct ctcOptions options);
ctcStatus_t get_workspace_size(const int *const label_lengths, const int *const input_lengths,
                               int alphabet_size, int minibatch, struct ctcOptions options, size_t *size_bytes);

/* ------------------------------------------------------------------------------------------------
 * Section 2 -- fused entry and helpers (additions)
 * ---------------------------------------------------------------------------------------------- */
#define CCB_DTYPE_F32 0
#define CCB_DTYPE_BF16 1

const char *ccb_last_error(void);            /* thread-local message of the last failure ("" if none) */
int ccb_den_loaded(int device);              /* 1 if Init() covered this device */
/* info[0..7] = file states, file arcs, states after the in-label split, pairs, forward / backward stream slots
 * (with padding), forward / backward arcs actually gathered per frame and utterance */
int ccb_den_info(long *info);

/* scratch sizes for N utterances x T frames on the current device's den graph */
size_t ccb_den_alpha_floats(int N, int T);   /* alpha spill, floats */
size_t ccb_den_aux_bytes(int N, int T);      /* everything else the den passes need */
size_t ccb_ctc_workspace_bytes(int N, int T, int max_label_len);

/* Denominator only: logZ (N,) fp32 device out; grad (strided, pre-zeroed, accumulated with
 * grad_scale); logz_beta may be NULL. */
int ccb_den_forward_backward(const void *logits, int dtype, long sn, long st, int N, int T, int V,
                             const int *len_dev, float *alpha_ws, void *aux_ws,
                             float *grad, long gsn, long gst, float grad_scale,
                             float *logz, float *logz_beta, void *stream);

/* Numerator only: meta_dev = device int32 block [labels(sumL) | label_off(N+1) | label_len(N) | len(N)];
 * grad accumulated with grad_scale (pass e.g. -(1+lamb)/N); logp (N,) fp32 device out. */
int ccb_ctc_forward_backward(const void *logits, int dtype, long sn, long st, int N, int T, int V,
                             const int *labels_dev, const int *label_off_dev, const int *label_len_dev,
                             const int *len_dev, int max_label_len, int blank, void *workspace,
                             float *grad, long gsn, long gst, float grad_scale, float *logp, void *stream);

/* Fused CTC-CRF loss (ctc_crf/__init__.py:58-90 in one call, no host synchronisation) of an (N,T,V) block:
 *   loss[0] = scale * sum_n (logZ_den[n] - (1+lamb) logp_ctc[n])
 *   grad    = scale * (gamma_den - (1+lamb) gamma_ctc)                     (N,T,V) fp32, zeroed here
 * scale = 1/batch for size_average (the caller may split a batch into several calls and sum the losses).
 * Tmax <= T is the number of frames actually walked (max input length of the block); workspaces are sized with
 * ccb_*_floats/bytes(N, Tmax).  parts (optional, may be NULL): 2N floats = [logZ_den | logp_ctc].
 * Everything is enqueued on `stream`.  (Opt-in, CCB_OVERLAP=1 at Init: the numerator's alpha || beta chains are forked onto a
 * library-owned side stream and joined back before the call's last kernels -- one such call per device at a time.) */
int ccb_ctc_crf_loss_fwd(const void *logits, int dtype, int N, int T, int V, int Tmax,
                         const int *labels_dev, const int *label_off_dev, const int *label_len_dev,
                         const int *len_dev, int max_label_len, float lamb, float scale,
                         float *alpha_ws, void *aux_ws, void *ctc_ws,
                         float *grad, float *loss, float *parts, void *stream);

/* Same, taking the RAW encoder outputs z (N,T,V) instead of log-probs (replaces the caller's
 * `logits.log_softmax(-1)` + its autograd backward, cat/ctc/train.py:173-174,184-190):
 *   loss as above evaluated on y = log_softmax(z);   grad = d loss / d z = g - softmax(z) * sum_k g_k.
 * No normalised copy of the logits is materialised. */
int ccb_ctc_crf_loss_logits_fwd(const void *logits, int dtype, int N, int T, int V, int Tmax,
                                const int *labels_dev, const int *label_off_dev, const int *label_len_dev,
                                const int *len_dev, int max_label_len, float lamb, float scale,
                                float *alpha_ws, void *aux_ws, void *ctc_ws,
                                float *grad, float *loss, float *parts, void *stream);

/* CTC-only loss (replaces the numerator-only path ctc_crf/__init__.py:25-56 = WARP_CTC_LOSS) on an (N,T,V) block of
 * log-probs, read in place (no (T,N,V) transpose copy, cf. __init__.py:31):
 *   logp[n] = log p(l_n | x_n)  (N,) device out;   loss[0] = -scale * sum_n logp[n];
 *   grad    = -scale * gamma_ctc, (N,T,V) fp32, every row WRITTEN here (zeros for t >= len and infeasible utterances), so
 *             the caller needs no zero fill (cf. __init__.py:32) and no host round trip for the costs (:35).
 * ctc_ws: ccb_ctc_workspace_bytes(N, Tmax, max_label_len). */
int ccb_ctc_loss_fwd(const void *logits, int dtype, int N, int T, int V, int Tmax,
                     const int *labels_dev, const int *label_off_dev, const int *label_len_dev, const int *len_dev,
                     int max_label_len, int blank, float scale, void *ctc_ws, float *grad, float *loss, float *logp,
                     void *stream);

/* Best-path (Viterbi) alignment over the numerator lattice -- the forced alignment that comes with the CTC numerator
 * (SURVEY.md 8f-4).  align: (N,T) int32 device out, the token emitted at every frame t < len (blank included), -1 beyond
 * the length and for infeasible utterances; score (optional, may be NULL): (N,) log-probability of the best path.
 * workspace: ccb_ctc_align_workspace_bytes(N, T, max_label_len) bytes. */
size_t ccb_ctc_align_workspace_bytes(int N, int T, int max_label_len);
int ccb_ctc_align(const void *logits, int dtype, long sn, long st, int N, int T, int V,
                  const int *labels_dev, const int *label_off_dev, const int *label_len_dev, const int *len_dev,
                  int max_label_len, int blank, void *workspace, int *align, float *score, void *stream);

/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long ccb_launch_count(void);

/* Profiling aid: when dev_buffer != NULL the den kernels record, for frames [step0, step0+nsteps) of each pass,
 * per warp chunk 4 x u64 {globaltimer at frame start, clock64 at frame start, clock64 at chunk done, clock64 after
 * the grid barrier} into dev_buffer[(frame-step0) * n_chunks + chunk][4].  Pass NULL to switch it off. */
void ccb_debug_timeline(void *dev_buffer, int step0, int nsteps);

/* ------------------------------------------------------------------------------------------------
 * Section 3 -- host-side den-graph plan (no GPU needed)
 * ---------------------------------------------------------------------------------------------- */
/* Parse + plan without touching CUDA.  n_ctas/n_warps describe the persistent grid the plan is cut for. */
void *ccb_plan_create(const char *fst_name, int n_ctas, int n_warps);
void ccb_plan_destroy(void *plan);
/* info[0..12] = S_file, A_file, S, A_fwd, A_bwd, start, num_labels, n_ctas, n_warps, max_tile_arcs, P (pairs), n_start_arcs,
 *               n_hubs (states whose forward row is split into parts) */
int ccb_plan_info(void *plan, long *info);
/* which: 0 state_label[S] i32, 1 final_lin[S] f32, 2 orig_state[S] i32,
 *        3 fwd_arcs[A_fwd] {u32 peer, f32 w (sign bits = segment events)}, 4 fwd_chunk_state[n_ctas*n_warps+1] i32,
 *        5 fwd_chunk_arc[n_ctas*n_warps+1] i32, 6 bwd_arcs[A_bwd], 7 bwd_chunk_state, 8 bwd_chunk_arc,
 *        9 state_pos[S] i32, 10 fwd_chunk_pair, 11 bwd_chunk_pair, 12 start_arcs[n_start_arcs],
 *        13 fwd_cta_labels[n_ctas*4] i32, 14 bwd_cta_labels, 15 bwd second weights f32[A_bwd], 16 hub_states[n_hubs] i32 */
int ccb_plan_copy(void *plan, int which, void *dst, size_t dst_bytes);

#ifdef __cplusplus
}
#endif
#endif /* CTC_CRF_B200_H_ */
