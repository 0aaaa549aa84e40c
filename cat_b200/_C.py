"""``ctc_crf._C`` -- host-side mirror of the reference's pybind module (src/ctc_crf/binding.cpp:51-126).

Same four entry points, same argument meaning and placement (``gpu_ctc`` takes CPU labels/lengths/costs and
(T,N,V) activations; ``gpu_den`` takes CUDA lengths and pre-zeroed gradients), implemented by calling the C ABI
of ``libctc_crf_b200.so`` through ctypes.  As in binding.cpp, scratch is taken from torch's caching allocator
and the work is enqueued on the current CUDA stream of the tensors' device.  Differences: a device guard is
installed (binding.cpp has none) and native errors raise ``RuntimeError`` instead of being
ignored (binding.cpp:105,111) or exiting the process (den_calculate.cu:16-25).

Addition: ``ctc_crf_loss_fwd`` -- the fused, synchronisation-free entry used by ``CTC_CRF_LOSS``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ctcOptions

_DTYPES = {torch.float32: 0, torch.bfloat16: 1}


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed: {_lib.last_error() or rc}")


def _raise_if_error(what: str) -> None:
    msg = _lib.last_error()
    if msg:
        raise RuntimeError(f"{what} failed: {msg}")


def _stream(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _iptr(t: torch.Tensor):
    return C.cast(t.data_ptr(), C.POINTER(C.c_int))


def init_env(fst_name: str, gpus: torch.Tensor) -> None:
    """binding.cpp:51-56"""
    assert gpus.dtype == torch.int32 and not gpus.is_cuda
    g = gpus.contiguous()
    _lib.lib().Init(fst_name.encode(), int(g.numel()), _iptr(g))
    _raise_if_error("init_env")


def release_env(gpus: torch.Tensor) -> None:
    """binding.cpp:58-63"""
    g = gpus.contiguous()
    _lib.lib().Release(int(g.numel()), _iptr(g))


def den_num_states() -> int:
    return C.c_int.in_dll(_lib.lib(), "DEN_NUM_STATES").value


def den_info() -> dict:
    """Sizes of the loaded den graph and of its kernel plan."""
    info = (C.c_long * 8)()
    if _lib.lib().ccb_den_info(info) != 0:
        raise RuntimeError("den graph not loaded")
    keys = ("file_states", "file_arcs", "states", "pairs", "fwd_slots", "bwd_slots", "fwd_arcs", "bwd_arcs")
    return dict(zip(keys, [int(x) for x in info]))


def den_num_arcs() -> int:
    return C.c_int.in_dll(_lib.lib(), "DEN_NUM_ARCS").value


def gpu_den(logits: torch.Tensor, grad_net: torch.Tensor, input_lengths: torch.Tensor,
            costs_alpha: torch.Tensor, costs_beta: torch.Tensor) -> None:
    """binding.cpp:65-84.  logits (N,T,V) fp32 CUDA contiguous; grad_net pre-zeroed, filled in place;
    input_lengths (N,) int32 CUDA; costs_* (N,) fp32 CUDA."""
    L = _lib.lib()
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous()
    assert grad_net.is_cuda and grad_net.dtype == torch.float32 and grad_net.is_contiguous()
    assert input_lengths.is_cuda and input_lengths.dtype == torch.int32
    N, T, V = logits.shape
    dev = logits.device
    with torch.cuda.device(dev):
        stream = _stream(dev)
        S = den_num_states()
        # binding.cpp:77-79 sizes; alpha is padded to whole 32-lane groups here
        # the library states its need through ccb_den_alpha_floats (lane padding, parked pair rows); alpha_size is the
        # per-(frame, utterance) row count binding.cpp would pass, rounded UP so that (T+1)*N*alpha_size covers the need
        per = (T + 1) * N
        alpha_states = max(S, -(-int(L.ccb_den_alpha_floats(N, T)) // per))
        alpha = torch.empty(per * alpha_states, dtype=torch.float32, device=dev)
        beta = torch.empty(1, dtype=torch.float32, device=dev)
        grad_storage = torch.empty(1, dtype=torch.float32, device=dev)
        L.compute_alpha(alpha.data_ptr(), logits.data_ptr(), N, T, alpha_states, V, input_lengths.data_ptr(),
                        costs_alpha.data_ptr(), stream)
        _raise_if_error("gpu_den/compute_alpha")
        L.compute_beta_and_grad(beta.data_ptr(), alpha.data_ptr(), logits.data_ptr(), costs_alpha.data_ptr(),
                                grad_storage.data_ptr(), grad_net.data_ptr(), N, T, alpha_states, V,
                                input_lengths.data_ptr(), costs_beta.data_ptr(), stream)
        _raise_if_error("gpu_den/compute_beta_and_grad")


def gpu_ctc(probs: torch.Tensor, grads: torch.Tensor, labels: torch.Tensor, label_sizes: torch.Tensor,
            sizes: torch.Tensor, minibatch_size: int, costs: torch.Tensor, blank_label: int) -> None:
    """binding.cpp:86-117.  probs/grads (T,N,V) fp32 CUDA (grads pre-zeroed); labels/label_sizes/sizes int32 CPU;
    costs (N,) fp32 CPU, receives log p(l|x) per utterance."""
    L = _lib.lib()
    assert probs.is_cuda and probs.dtype == torch.float32 and probs.is_contiguous()
    assert not labels.is_cuda and not label_sizes.is_cuda and not sizes.is_cuda and not costs.is_cuda
    assert labels.dtype == torch.int32 and label_sizes.dtype == torch.int32 and sizes.dtype == torch.int32
    assert costs.dtype == torch.float32 and costs.is_contiguous()
    labels, label_sizes, sizes = labels.contiguous(), label_sizes.contiguous(), sizes.contiguous()
    V = probs.size(2)
    dev = probs.device
    with torch.cuda.device(dev):
        opts = ctcOptions(stream=torch.cuda.current_stream(dev).cuda_stream, blank_label=int(blank_label))
        nbytes = C.c_size_t(0)
        st = L.get_workspace_size(_iptr(label_sizes), _iptr(sizes), V, minibatch_size, opts, C.byref(nbytes))
        if st != 0:
            raise RuntimeError(f"gpu_ctc/get_workspace_size: {L.ctcGetStatusString(st).decode()}")
        ws = torch.empty((nbytes.value + 3) // 4, dtype=torch.float32, device=dev)
        gptr = grads.data_ptr() if grads is not None and grads.numel() > 0 else None
        st = L.compute_ctc_loss(probs.data_ptr(), gptr, _iptr(labels), _iptr(label_sizes), _iptr(sizes), V,
                                minibatch_size, C.cast(costs.data_ptr(), C.POINTER(C.c_float)), ws.data_ptr(), opts)
        if st != 0:
            raise RuntimeError(f"gpu_ctc/compute_ctc_loss: {L.ctcGetStatusString(st).decode()} {_lib.last_error()}")


# ---------------------------------------------------------------------------------------------------
# fused entry
# ---------------------------------------------------------------------------------------------------
class _PinnedRing:
    """Small ring of pinned int32 staging buffers so the label/length upload is a true async H2D copy."""

    def __init__(self, slots: int = 4):
        self.bufs = [None] * slots
        self.events = [None] * slots
        self.i = 0

    def stage(self, parts, device) -> torch.Tensor:
        n = sum(int(p.numel()) for p in parts)
        i = self.i
        self.i = (i + 1) % len(self.bufs)
        if self.events[i] is not None:
            self.events[i].synchronize()
        if self.bufs[i] is None or self.bufs[i].numel() < n:
            self.bufs[i] = torch.empty(max(n, 1024), dtype=torch.int32, pin_memory=True)
        buf = self.bufs[i][:n]
        torch.cat(parts, out=buf)
        dev = buf.to(device, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(device))
        self.events[i] = ev
        return dev


_rings = {}


def upload_meta(labels: torch.Tensor, lx: torch.Tensor, ly: torch.Tensor, device) -> Tuple[torch.Tensor, int, int]:
    """Pack [labels | label_off(N+1) | ly | lx] into one pinned buffer and upload it asynchronously.
    Returns (device int32 tensor, sum(ly), max(ly))."""
    N = int(lx.numel())
    ly64 = ly.to(torch.int64)
    off = torch.zeros(N + 1, dtype=torch.int32)
    off[1:].copy_(torch.cumsum(ly64, 0))
    sum_l = int(off[-1])
    if sum_l != int(labels.numel()):
        raise RuntimeError(f"labels has {labels.numel()} entries but label lengths sum to {sum_l}")
    ring = _rings.setdefault(torch.device(device).index, _PinnedRing())
    meta = ring.stage([labels, off, ly, lx], device)
    return meta, sum_l, int(ly.max()) if N else 0


MAX_UTTS_PER_CALL = 64           # utterances per native call: the den kernels are fastest per utterance at two utterances
                                 # per lane (N=64: 35 us per frame; N=128 in one call: 79 us), and a slice of a
                                 # length-sorted batch walks only its own longest utterance (N=256, len~U{200..3000}:
                                 # 7.8 k frame steps in four slices instead of 12 k in one).  Hard limit of the kernels: 512.
_WS_FRACTION = 0.85              # of the currently free device memory a call may use for scratch


_budget_cache = {}    # device index -> [calls until the next refresh, budget in bytes]


def _scratch_budget(dev, need: int = 0) -> int:
    """What a torch allocation can get now: free device memory plus what the caching allocator holds but is not using.
    cudaMemGetInfo costs ~0.6 ms, so the answer is reused for a few calls unless the request does not fit it."""
    key = torch.device(dev).index
    ent = _budget_cache.get(key)
    if ent is not None and ent[0] > 0 and need <= ent[1]:
        ent[0] -= 1
        return ent[1]
    free, _ = torch.cuda.mem_get_info(dev)
    free += max(0, torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev))
    budget = int(free * _WS_FRACTION)
    _budget_cache[key] = [16, budget]
    return budget


def _plan_slices(L, lx, ly, N, V, dev):
    """Split the batch into contiguous slices whose scratch (alpha spill etc.) fits the free device memory.
    Each slice walks only max(lx[slice]) frames."""
    budget = _scratch_budget(dev)
    slices, n0 = [], 0
    while n0 < N:
        n1 = min(N, n0 + MAX_UTTS_PER_CALL)
        while True:
            tmax = max(1, int(lx[n0:n1].max()))
            maxl = int(ly[n0:n1].max())
            need = (int(L.ccb_den_alpha_floats(n1 - n0, tmax)) * 4 + int(L.ccb_den_aux_bytes(n1 - n0, tmax))
                    + int(L.ccb_ctc_workspace_bytes(n1 - n0, tmax, maxl)))
            if need > budget:
                budget = _scratch_budget(dev, need)     # a stale (cached) figure must not split a batch needlessly
            if need <= budget or n1 - n0 == 1:
                break
            n1 = n0 + max(1, (n1 - n0) // 2)
        if need > budget:
            raise RuntimeError(f"CTC-CRF scratch for a single utterance of {tmax} frames needs {need >> 20} MiB, "
                               f"only {budget >> 20} MiB free")
        slices.append((n0, n1, tmax, maxl))
        n0 = n1
    return slices


def ctc_crf_loss_fwd(logits: torch.Tensor, labels: torch.Tensor, lx: torch.Tensor, ly: torch.Tensor,
                     lamb: float, size_average: bool, want_parts: bool = False, from_logits: bool = False
                     ) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Fused CTC-CRF loss.  logits (N,T,V) fp32 or bf16 CUDA contiguous log-probs; labels/lx/ly int32 CPU.
    Returns (loss[1] fp32 CUDA, grad (N,T,V) fp32 CUDA, parts[2N] or None).  Never synchronises the host.
    Batches whose scratch does not fit the free memory (or > 512 utterances) are processed in slices.
    from_logits: `logits` are the RAW encoder outputs; log_softmax and its Jacobian are applied inside
    (grad = d loss / d raw logits), replacing cat/ctc/train.py:173-174."""
    L = _lib.lib()
    entry = L.ccb_ctc_crf_loss_logits_fwd if from_logits else L.ccb_ctc_crf_loss_fwd
    assert logits.is_cuda and logits.dim() == 3 and logits.is_contiguous()
    if logits.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported logits dtype {logits.dtype}")
    N, T, V = _check_batch_args(logits, labels, lx, ly)
    dev = logits.device
    esz = logits.element_size()
    scale = 1.0 / N if size_average else 1.0
    with torch.cuda.device(dev):
        meta, sum_l, _ = upload_meta(labels, lx, ly, dev)
        base = meta.data_ptr()
        p_labels, p_off = base, base + 4 * sum_l
        p_ly, p_lx = p_off + 4 * (N + 1), p_off + 4 * (2 * N + 1)
        grad = torch.empty((N, T, V), dtype=torch.float32, device=dev)
        parts = torch.empty(2 * N, dtype=torch.float32, device=dev) if want_parts else None
        stream = _stream(dev)

        def run_slices():
            slices = _plan_slices(L, lx, ly, N, V, dev)
            losses = torch.empty(len(slices), dtype=torch.float32, device=dev)
            for i, (n0, n1, tmax, maxl) in enumerate(slices):
                n = n1 - n0
                alpha_ws = torch.empty(int(L.ccb_den_alpha_floats(n, tmax)), dtype=torch.float32, device=dev)
                aux_ws = torch.empty(int(L.ccb_den_aux_bytes(n, tmax)), dtype=torch.uint8, device=dev)
                ctc_ws = torch.empty(int(L.ccb_ctc_workspace_bytes(n, tmax, maxl)), dtype=torch.uint8, device=dev)
                sub_parts = torch.empty(2 * n, dtype=torch.float32, device=dev) if want_parts else None
                rc = entry(logits.data_ptr() + n0 * T * V * esz, _DTYPES[logits.dtype], n, T, V, tmax,
                           p_labels, p_off + 4 * n0, p_ly + 4 * n0, p_lx + 4 * n0, maxl, float(lamb),
                           float(scale), alpha_ws.data_ptr(), aux_ws.data_ptr(), ctc_ws.data_ptr(),
                           grad.data_ptr() + n0 * T * V * 4, losses.data_ptr() + 4 * i,
                           sub_parts.data_ptr() if want_parts else None, stream)
                _check(rc, "ctc_crf_loss_fwd")
                if want_parts:
                    parts[n0:n1] = sub_parts[:n]
                    parts[N + n0:N + n1] = sub_parts[n:]
                del alpha_ws, aux_ws, ctc_ws      # stream-ordered reuse by the caching allocator for the next slice
            return slices, losses

        try:
            slices, losses = run_slices()
        except torch.cuda.OutOfMemoryError:       # the cached memory figure was stale: measure again, slice finer, redo
            _budget_cache.pop(torch.device(dev).index, None)
            torch.cuda.empty_cache()
            slices, losses = run_slices()
        loss = losses.sum().reshape(1) if len(slices) > 1 else losses
        meta.record_stream(torch.cuda.current_stream(dev))
    return loss, grad, parts


def _check_batch_args(logits, labels, lx, ly):
    assert logits.is_cuda and logits.dim() == 3 and logits.is_contiguous()
    if logits.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported logits dtype {logits.dtype}")
    N, T, V = logits.shape
    if N == 0:
        raise RuntimeError("empty batch")
    if int(lx.numel()) != N or int(ly.numel()) != N:
        raise RuntimeError("lx / ly must have one entry per utterance")
    if int(lx.max()) > T or int(lx.min()) < 0:
        raise RuntimeError("input lengths must lie in [0, T]")
    if int(ly.min()) < 0:
        raise RuntimeError("label lengths must be non-negative")
    if labels.numel() and (int(labels.min()) < 0 or int(labels.max()) >= V):
        raise RuntimeError("label out of range")
    return N, T, V


def ctc_loss_fwd(logits: torch.Tensor, labels: torch.Tensor, lx: torch.Tensor, ly: torch.Tensor,
                 size_average: bool, blank: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """CTC-only loss on the (N,T,V) log-probs in place (the native counterpart of ``_WARP_CTC_GPU``,
    ctc_crf/__init__.py:25-56): no (T,N,V) transpose copy, no zero fill, no host round trip for the costs.
    Returns (loss[1], grad (N,T,V) fp32 = d loss / d logits, logp[N]) on the device; never synchronises the host."""
    L = _lib.lib()
    N, T, V = _check_batch_args(logits, labels, lx, ly)
    dev = logits.device
    scale = 1.0 / N if size_average else 1.0
    with torch.cuda.device(dev):
        meta, sum_l, maxl = upload_meta(labels, lx, ly, dev)
        base = meta.data_ptr()
        p_labels, p_off = base, base + 4 * sum_l
        p_ly, p_lx = p_off + 4 * (N + 1), p_off + 4 * (2 * N + 1)
        tmax = max(1, int(lx.max()))
        grad = torch.empty((N, T, V), dtype=torch.float32, device=dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        logp = torch.empty(N, dtype=torch.float32, device=dev)
        ws = torch.empty(int(L.ccb_ctc_workspace_bytes(N, tmax, maxl)), dtype=torch.uint8, device=dev)
        rc = L.ccb_ctc_loss_fwd(logits.data_ptr(), _DTYPES[logits.dtype], N, T, V, tmax, p_labels, p_off, p_ly, p_lx, maxl,
                                int(blank), float(scale), ws.data_ptr(), grad.data_ptr(), loss.data_ptr(), logp.data_ptr(),
                                _stream(dev))
        _check(rc, "ctc_loss_fwd")
        meta.record_stream(torch.cuda.current_stream(dev))
    return loss, grad, logp


def ctc_align(logits: torch.Tensor, labels: torch.Tensor, lx: torch.Tensor, ly: torch.Tensor, blank: int = 0
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Best-path (Viterbi) forced alignment over the numerator lattice.  logits (N,T,V) fp32/bf16 CUDA
    log-probs; labels/lx/ly int32 CPU as for the loss.  Returns (align (N,T) int32 CUDA: the token of every frame
    t < lx[n], blank included, -1 beyond; score (N,) fp32 CUDA: log-probability of that path, -inf if infeasible)."""
    L = _lib.lib()
    N, T, V = _check_batch_args(logits, labels, lx, ly)
    dev = logits.device
    with torch.cuda.device(dev):
        meta, sum_l, maxl = upload_meta(labels, lx, ly, dev)
        base = meta.data_ptr()
        p_labels, p_off = base, base + 4 * sum_l
        p_ly, p_lx = p_off + 4 * (N + 1), p_off + 4 * (2 * N + 1)
        align = torch.empty((N, T), dtype=torch.int32, device=dev)
        score = torch.empty(N, dtype=torch.float32, device=dev)
        ws = torch.empty(int(L.ccb_ctc_align_workspace_bytes(N, T, maxl)), dtype=torch.uint8, device=dev)
        rc = L.ccb_ctc_align(logits.data_ptr(), _DTYPES[logits.dtype], T * V, V, N, T, V, p_labels, p_off, p_ly, p_lx, maxl,
                             int(blank), ws.data_ptr(), align.data_ptr(), score.data_ptr(), _stream(dev))
        _check(rc, "ctc_align")
        meta.record_stream(torch.cuda.current_stream(dev))
    return align, score


def launch_count() -> int:
    return int(_lib.lib().ccb_launch_count())
