"""Python surface of the CTC-CRF loss -- drop-in for the reference's ``ctc_crf`` package
(src/ctc_crf/ctc_crf/__init__.py:25-171): same class names, constructor arguments, argument placement
(logits on the GPU, labels / lx / ly int32 on the CPU), dtype asserts and error behaviour.

What changed underneath: the reference's forward makes ~11 passes over (N,T,V)
(transpose copy, two zero fills, combine, scale, backward multiply) and synchronises the host three times;
here ``_CTC_CRF.forward`` is ONE native call that writes the final gradient once and never synchronises.
"""
from __future__ import annotations

import os
from typing import List, Union

import torch
from torch.autograd import Function
from torch.nn import Module

from . import _C as core

__version__ = "0.1.0"


def _assert_no_grad(tensor):
    assert not tensor.requires_grad, "shouldn't require grads"


class _WARP_CTC_GPU(Function):
    """__init__.py:25-56: CTC loss on log-softmax inputs.  Same arguments, result and gradient as the reference's
    Function; underneath it is ONE native call on the (N,T,V) tensor in place -- no (T,N,V) transpose copy (:31), no
    zero-filled gradient (:32), no host round trip for the per-utterance costs (:33-35) -- see ``_C.ctc_loss_fwd``.
    (``_C.gpu_ctc`` keeps the reference's pybind signature for callers that use it directly.)"""

    @staticmethod
    def forward(ctx, logits, labels, input_lengths, label_lengths, size_average=True):
        logits = logits.contiguous()
        costs, grads, _ = core.ctc_loss_fwd(logits, labels, input_lengths, label_lengths, size_average)
        ctx.grads = grads
        return costs

    @staticmethod
    def backward(ctx, grad_output):
        return ctx.grads * grad_output.to(ctx.grads.device), None, None, None, None, None, None


class _CTC_CRF(Function):
    """__init__.py:58-94.  loss = sum_n [logZ_den(n) - (1+lamb) log p_ctc(n)] (/N),
    grad = [gamma_den - (1+lamb) gamma_ctc] (/N); backward = grads * grad_output."""

    @staticmethod
    def forward(ctx, logits, labels, input_lengths, label_lengths, lamb=0.1, size_average=True):
        logits = logits.contiguous()
        costs, grads, _ = core.ctc_crf_loss_fwd(logits, labels, input_lengths, label_lengths, lamb, size_average)
        ctx.grads = grads
        return costs

    @staticmethod
    def backward(ctx, grad_output):
        return ctx.grads * grad_output.to(ctx.grads.device), None, None, None, None, None, None


class _CTC_CRF_LOGITS(Function):
    """the same loss taken on the RAW encoder outputs.  Replaces, in one native call,
    ``logits = net_out.float().log_softmax(-1); loss = _CTC_CRF(logits, ...)`` and the autograd backward of the cast
    and of log_softmax (cat/ctc/train.py:173-174,184-190): no normalised fp32 copy is materialised and the gradient
    comes back already chained through the softmax Jacobian, in the dtype of the input."""

    @staticmethod
    def forward(ctx, logits, labels, input_lengths, label_lengths, lamb=0.1, size_average=True):
        logits = logits.contiguous()
        costs, grads, _ = core.ctc_crf_loss_fwd(logits, labels, input_lengths, label_lengths, lamb, size_average,
                                                from_logits=True)
        ctx.grads = grads
        ctx.in_dtype = logits.dtype
        return costs

    @staticmethod
    def backward(ctx, grad_output):
        g = ctx.grads * grad_output.to(ctx.grads.device)
        return g.to(ctx.in_dtype), None, None, None, None, None, None


class CTC_CRF_LOSS(Module):
    def __init__(self, lamb: float = 0.1, size_average: bool = True, from_logits: bool = False):
        """
        lamb (float): weight for auxiliary CTC loss, final loss = lamb * loss_ctc + loss_crf
        size_average (bool): whether to do average over batch size dimension.
        from_logits (bool): addition to the reference signature -- ``forward`` takes the raw encoder outputs and
            applies log_softmax (and its backward) inside the fused call.  Default False = reference behaviour.
        """
        super(CTC_CRF_LOSS, self).__init__()
        self.ctc_crf = _CTC_CRF_LOGITS.apply if from_logits else _CTC_CRF.apply
        self.lamb = lamb
        self.size_average = size_average
        self.from_logits = from_logits

    def forward(self, logits, labels, lx, ly) -> torch.FloatTensor:
        """
        logits (torch.FloatTensor): size (N, T, V), log-probabilities, on GPU device.
            (torch.bfloat16 is accepted too; accumulation is fp32 -- the reference asserts fp32.)
        labels (torch.IntTensor)  : size (sum {Ui}, ) flattened without paddings, on CPU.
        lx (torch.IntTensor) : size (N, ), on CPU
        ly (torch.IntTensor) : size (N, ), on CPU
        """
        assert len(labels.size()) == 1
        assert logits.dtype in (torch.float, torch.bfloat16), \
            f"expect logits to be torch.float (or torch.bfloat16) object, instead: {logits.dtype}"
        assert labels.dtype == torch.int, f"expect labels to be torch.int object, instead: {labels.dtype}"
        assert lx.dtype == torch.int, f"expect lx to be torch.int object, instead: {lx.dtype}"
        assert ly.dtype == torch.int, f"expect ly to be torch.int object, instead: {ly.dtype}"
        _assert_no_grad(labels)
        _assert_no_grad(lx)
        _assert_no_grad(ly)
        return self.ctc_crf(logits, labels, lx, ly, self.lamb, self.size_average)


class WARP_CTC_LOSS(Module):
    """__init__.py:128-144"""

    def __init__(self, size_average=True):
        super(WARP_CTC_LOSS, self).__init__()
        self.ctc = _WARP_CTC_GPU.apply
        self.size_average = size_average

    def forward(self, logits, labels, input_lengths, label_lengths):
        assert len(labels.size()) == 1
        _assert_no_grad(labels)
        _assert_no_grad(input_lengths)
        _assert_no_grad(label_lengths)
        return self.ctc(logits, labels, input_lengths, label_lengths, self.size_average)


def ctc_align(logits, labels, input_lengths, label_lengths):
    """Best-path forced alignment of every utterance to its label sequence over the CTC lattice -- the alignment that comes
    with the numerator (cat/ctc/train_jsa.py consumes frame-level paths).  Arguments as for WARP_CTC_LOSS.
    Returns (align (N,T) int32 on the GPU, token per frame incl. blank 0, -1 past each length; score (N,) fp32)."""
    assert len(labels.size()) == 1
    return core.ctc_align(logits.contiguous(), labels, input_lengths, label_lengths)


class CRFContext:
    def __init__(self, den_lm: str, gpus: Union[int, List[int]]) -> None:
        """
        den_lm (str): path to the denominator LM (OpenFst binary, see cat/utils/tool/prep_den_lm.sh)
        gpus   (int, List[int]): Specify which GPU to be used.
        """
        if not os.path.isfile(den_lm):
            raise RuntimeError(f"Denominator LM model location is invalid: {den_lm}.")
        if isinstance(gpus, int):
            gpus = [gpus]
        nprocs = torch.cuda.device_count()
        if not all([i >= 0 and i < nprocs for i in gpus]):
            raise RuntimeError(f"Available GPU={nprocs}, invalid GPU ids: {gpus}.")
        gpu_t = torch.IntTensor(gpus)
        core.init_env(den_lm, gpu_t)
        self._gpus = gpu_t

    def __del__(self):
        if hasattr(self, '_gpus'):
            try:
                core.release_env(self._gpus)
            except Exception:      # interpreter shutdown: module globals may already be gone
                pass
            del self._gpus
