"""Numpy emulation of the den kernels' arithmetic on the *product's* plan arrays (CPU-only tests).

Walks exactly what den_kernels.cu does -- scaled-linear alpha/beta with power-of-two per-frame scales,
hoisted emissions, state-product occupancies normalised by sum_q alpha*beta -- in float64 so that any
disagreement with the oracle is a plan/algorithm error, not rounding.
"""
import numpy as np


SCALE_EXP = 32


def _scale(s):
    if not (s > 0) or not np.isfinite(s):
        return 1.0, 0
    ex = int(np.floor(np.log2(s)))
    return 2.0 ** (SCALE_EXP - ex), SCALE_EXP - ex


def _rows(passview):
    row = passview.row_of_arc()
    peer = passview.arcs["peer"].astype(np.int64)
    w = passview.weights().astype(np.float64)
    keep = row < len(passview.row_ends())     # padding after the very last row belongs to no row
    return row[keep], peer[keep], w[keep]


def den_emulate(plan, y, lens):
    """Returns (logz_alpha, logz_beta, gamma[N,T,V]) for log-probs y (N,T,V)."""
    N, T, V = y.shape
    S = plan.num_states
    lab = plan.state_label.astype(np.int64)
    fin = plan.final_lin.astype(np.float64)
    frow, fpeer, fw = _rows(plan.fwd)
    brow, bpeer, bw = _rows(plan.bwd)
    logz_a = np.zeros(N)
    logz_b = np.zeros(N)
    gamma = np.zeros((N, T, V))
    for n in range(N):
        Tn = int(lens[n])
        yn = y[n].astype(np.float64)
        fmax = yn.max(-1)
        alpha = np.zeros((Tn + 1, S))
        alpha[0, plan.start] = 1.0
        colsum = 1.0
        runlog = 0.0
        for t in range(1, Tn + 1):
            r, sh = _scale(colsum)
            acc = np.bincount(frow, weights=fw * alpha[t - 1, fpeer], minlength=S)
            e = np.exp(yn[t - 1, lab] - fmax[t - 1])
            alpha[t] = acc * e * r
            runlog += fmax[t - 1] - sh * np.log(2.0)
            colsum = alpha[t].sum()
        with np.errstate(divide="ignore"):
            logz_a[n] = np.log((alpha[Tn] * fin).sum()) + runlog
        # backward
        bh_next = None
        colsum_b = 0.0
        runlog = 0.0
        for tau in range(Tn, 0, -1):
            if tau == Tn:
                b = fin.copy()
            else:
                rb, sh = _scale(colsum_b)
                b = rb * np.bincount(brow, weights=bw * bh_next[bpeer], minlength=S)
                runlog += fmax[tau] - sh * np.log(2.0)
            ab = alpha[tau] * b
            tot = ab.sum()
            if tot > 0:
                gamma[n, tau - 1] = np.bincount(lab, weights=ab, minlength=V)[:V] / tot
            e = np.exp(yn[tau - 1, lab] - fmax[tau - 1])
            bh_next = e * b
            colsum_b = bh_next.sum()
        if Tn > 0:
            rb, sh = _scale(colsum_b)
            runlog += fmax[0] - sh * np.log(2.0)
            m = brow == plan.start
            b0 = rb * (bw[m] * bh_next[bpeer[m]]).sum()
        else:
            b0 = fin[plan.start]
        with np.errstate(divide="ignore"):
            logz_b[n] = np.log(b0) + runlog
    return logz_a, logz_b, gamma
