"""Numpy emulation of the den kernels' arithmetic on the *product's* plan arrays (CPU-only tests).

Walks exactly what den_kernels.cu does -- scaled-linear alpha/beta with power-of-two per-frame scales,
hoisted emissions, state-product occupancies normalised by sum_q alpha*beta -- in float64 so that any
disagreement with the oracle is a plan/algorithm error, not rounding.
"""
import numpy as np


SCALE_EXP = 56      # DenPlan::scale_exp for probability-like graphs (den_graph.cc)


def _scale(s):
    if not (s > 0) or not np.isfinite(s):
        return 1.0, 0
    ex = int(np.floor(np.log2(s)))
    return 2.0 ** (SCALE_EXP - ex), SCALE_EXP - ex


def _decode_forward(plan):
    """Walk the forward stream like den_forward_kernel: returns (row, peer, w) triples over the gather table of
    S real rows + P virtual (pair-sum) rows, and for every pair its two member rows."""
    S = plan.num_states
    row, peer, w = [], [], []
    pairs = []
    q = 0
    arcs = plan.fwd.arcs
    hubs = set(int(h) for h in plan.hub_states)
    merged = plan.fwd_merged
    for a0, a1, ev, _chg in plan.fwd.segments():
        n = a1 - a0
        if ev == 3 and merged and plan.own_rows:   # both rows of a pair end; every slot belongs to the second member (q+1)
            assert plan.state_pos[q] == 0 and plan.state_pos[q + 1] == 1
            row.append(np.full(n, q + 1)); peer.append(arcs["peer"][a0:a1].astype(np.int64)); w.append(np.abs(arcs["w"][a0:a1]).astype(np.float64))
            pairs.append((q, q + 1))
            q += 2
            continue
        if ev == 3 and merged:   # both rows of a pair: slots 0..n-2 -> second member (q+1), last slot -> first member (q)
            assert plan.state_pos[q] == 0 and plan.state_pos[q + 1] == 1 and arcs["w"][a1 - 1] != 0
            row.append(np.full(n - 1, q + 1)); peer.append(arcs["peer"][a0:a1 - 1].astype(np.int64)); w.append(np.abs(arcs["w"][a0:a1 - 1]).astype(np.float64))
            row.append(np.full(1, q)); peer.append(arcs["peer"][a1 - 1:a1].astype(np.int64)); w.append(np.abs(arcs["w"][a1 - 1:a1]).astype(np.float64))
            pairs.append((q, q + 1))
            q += 2
            continue
        if ev == 3:     # a part of a high in-degree row: the last slot names the target row (weight 0)
            tgt = int(arcs["peer"][a1 - 1])
            assert tgt in hubs and arcs["w"][a1 - 1] == 0
            row.append(np.full(n - 1, tgt)); peer.append(arcs["peer"][a0:a1 - 1].astype(np.int64)); w.append(np.abs(arcs["w"][a0:a1 - 1]).astype(np.float64))
            if tgt == q:
                assert plan.state_pos[q] == 1
                q += 1
            continue
        assert q not in hubs
        row.append(np.full(n, q)); peer.append(arcs["peer"][a0:a1].astype(np.int64)); w.append(np.abs(arcs["w"][a0:a1]).astype(np.float64))
        if ev == 1:
            assert plan.state_pos[q] == 0
            pos0 = q
        else:
            assert plan.state_pos[q] == 1
            if ev == 2:
                assert pos0 == q - 1
                pairs.append((pos0, q))
        q += 1
    assert q == S and len(pairs) == plan.num_pairs
    return np.concatenate(row), np.concatenate(peer), np.concatenate(w), np.array(pairs, dtype=np.int64).reshape(-1, 2)


def _decode_backward(plan):
    """Walk the backward stream like den_backward_kernel: one segment per group, two weights per slot (w for the
    group's first row, w1 for its second row)."""
    S = plan.num_states
    row, peer, w = [], [], []
    q = 0
    arcs, w1 = plan.bwd.arcs, plan.bwd.w1
    for a0, a1, ev, _chg in plan.bwd.segments():
        pr = arcs["peer"][a0:a1].astype(np.int64)
        assert (pr < S).all()
        n = a1 - a0
        if ev == 2:     # pair: rows q (pos0) and q+1 (pos1)
            assert plan.state_pos[q] == 0 and plan.state_pos[q + 1] == 1
            row.append(np.full(n, q)); peer.append(pr); w.append(np.abs(arcs["w"][a0:a1]).astype(np.float64))
            row.append(np.full(n, q + 1)); peer.append(pr); w.append(w1[a0:a1].astype(np.float64))
            q += 2
        else:
            assert ev == 0 and plan.state_pos[q] == 1 and not w1[a0:a1].any()
            row.append(np.full(n, q)); peer.append(pr); w.append(np.abs(arcs["w"][a0:a1]).astype(np.float64))
            q += 1
    assert q == S
    assert (w1 >= 0).all()
    return np.concatenate(row), np.concatenate(peer), np.concatenate(w)


def den_emulate(plan, y, lens):
    """Returns (logz_alpha, logz_beta, gamma[N,T,V]) for log-probs y (N,T,V)."""
    N, T, V = y.shape
    S = plan.num_states
    lab = plan.state_label.astype(np.int64)
    fin = plan.final_lin.astype(np.float64)
    frow, fpeer, fw, pairs = _decode_forward(plan)
    brow, bpeer, bw = _decode_backward(plan)
    P = plan.num_pairs
    # own-row terms: row q receives c[q,0] * X(first row of its group) + c[q,1] * X(second row of its group, or q itself)
    pos = plan.state_pos.astype(np.int64)
    ids = np.arange(S)
    after_pos0 = np.concatenate([[False], pos[:-1] == 0])          # q is the second member of a pair
    g0 = np.where(pos == 0, ids, np.where(after_pos0, ids - 1, ids))
    g1 = np.minimum(np.where(pos == 0, ids + 1, ids), S - 1)
    cf, cb = plan.own_fwd.astype(np.float64), plan.own_bwd.astype(np.float64)
    unp = (pos == 1) & ~after_pos0                                   # unpaired rows: only the second coefficient may be set
    assert not cf[unp, 0].any() and not cb[unp, 0].any()
    sa = plan.start_arcs
    logz_a = np.zeros(N)
    logz_b = np.zeros(N)
    gamma = np.zeros((N, T, V))
    for n in range(N):
        Tn = int(lens[n])
        yn = y[n].astype(np.float64)
        fmax = yn.max(-1)
        alpha = np.zeros((Tn + 1, S + P))       # real rows, then the pair-sum rows the next frame gathers
        alpha[0, plan.start] = 1.0
        if P:
            alpha[0, S:] = alpha[0, pairs[:, 0]] + alpha[0, pairs[:, 1]]
        colsum = 1.0
        runlog = 0.0
        for t in range(1, Tn + 1):
            r, sh = _scale(colsum)
            acc = np.bincount(frow, weights=fw * alpha[t - 1, fpeer], minlength=S)
            acc = acc + cf[:, 0] * alpha[t - 1, g0] + cf[:, 1] * alpha[t - 1, g1]
            e = np.exp(yn[t - 1, lab] - fmax[t - 1])
            alpha[t, :S] = acc * e * r
            if P:
                alpha[t, S:] = alpha[t, pairs[:, 0]] + alpha[t, pairs[:, 1]]
            runlog += fmax[t - 1] - sh * np.log(2.0)
            colsum = alpha[t, :S].sum()
        with np.errstate(divide="ignore"):
            logz_a[n] = np.log((alpha[Tn, :S] * fin).sum()) + runlog
        # backward
        bh_next = None
        colsum_b = 0.0
        runlog = 0.0
        for tau in range(Tn, 0, -1):
            if tau == Tn:
                b = fin.copy()
            else:
                rb, sh = _scale(colsum_b)
                b = rb * (np.bincount(brow, weights=bw * bh_next[bpeer], minlength=S) + cb[:, 0] * bh_next[g0] + cb[:, 1] * bh_next[g1])
                runlog += fmax[tau] - sh * np.log(2.0)
            ab = alpha[tau, :S] * b
            tot = ab.sum()
            if tot > 0:
                gamma[n, tau - 1] = np.bincount(lab, weights=ab, minlength=V)[:V] / tot
            e = np.exp(yn[tau - 1, lab] - fmax[tau - 1])
            bh_next = e * b
            colsum_b = bh_next.sum()
        if Tn > 0:
            rb, sh = _scale(colsum_b)
            runlog += fmax[0] - sh * np.log(2.0)
            b0 = rb * (sa["w"].astype(np.float64) * bh_next[sa["peer"].astype(np.int64)]).sum()
        else:
            b0 = fin[plan.start]
        with np.errstate(divide="ignore"):
            logz_b[n] = np.log(b0) + runlog
    return logz_a, logz_b, gamma
