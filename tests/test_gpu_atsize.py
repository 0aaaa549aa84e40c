"""At-size GPU parity (BASELINE.json `configs` 1-5 and the headline shape) + adversarial inputs.

The CUDA path runs the FULL configuration; the fp64 oracle re-computes a subset of the utterances (first, last, and for
the variable-length batch the longest and the shortest) with the scaled-linear fp64 evaluation `oracle.den(fast=True)`,
which tests/test_oracle.py pins against the log-domain restatement of den_calculate.cu.  Tolerances are the north_star's:
loss 1e-4 relative, gradients (occupancies, in [0,1]) 1e-3 absolute -- both against the fp64 oracle.  The reference's own
fp32 log-domain CUDA code is itself several 1e-2 away from the oracle at T >= 800 (tests/test_gpu_parity.py::
test_full_size_properties prints the figure), so the tight comparison against it stays at T=120 there.

Synthetic inputs follow SURVEY.md 8(d): x = 3 randn, y = log_softmax(x); L_n = min(len/6, 400); T-compose-LM shaped
den graph with H LM states and d = 24 successors (H=20 000 -> 1.02 M arcs, H=100 000 -> 5.09 M arcs).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-4
GRAD_ATOL = 1e-3


@pytest.fixture(scope="module")
def graphs(tmp_path_factory):
    """Benchmark-sized den graphs, generated once per module and cached on disk."""
    from cat_b200 import fst
    d = tmp_path_factory.mktemp("atsize")
    cache = {}

    def get(H, V, few_finals=False):
        key = (H, V, few_finals)
        if key not in cache:
            g = fst.make_synthetic_den(H, 24, V, seed=7)
            if few_finals:      # a graph whose final states are a small minority: 6 of the 2H-1 states
                keep = np.flatnonzero(~np.isinf(g.final))[::max(1, (2 * H - 1) // 6)][:6]
                fin = np.full_like(g.final, np.inf)
                fin[keep] = g.final[keep]
                g.final = fin
            path = str(d / f"den_{H}_{V}_{int(few_finals)}.fst")
            fst.write_fst(path, g)
            cache[key] = (path, g)
        return cache[key]
    return get


def _ctx(path):
    import ctc_crf
    return ctc_crf.CRFContext(path, gpus=0)


def _run(y_t, labels, lens, ly, lamb, from_logits=False):
    """Full loss through the reference-facing module; returns (loss, grad tensor on the GPU, fp32)."""
    import ctc_crf
    logits = y_t.clone().requires_grad_(True)
    crit = ctc_crf.CTC_CRF_LOSS(lamb=lamb, from_logits=from_logits)
    loss = crit(logits, torch.tensor(labels, dtype=torch.int32), torch.tensor(lens, dtype=torch.int32),
                torch.tensor(ly, dtype=torch.int32))
    loss.backward()
    return float(loss.item()), logits.grad.float()


def _check_subset(g, y_sub, labels, lens, ly, lamb, sub, loss, grad, N, parts=None, grad_atol=GRAD_ATOL):
    """Oracle on the utterances `sub` (y_sub = the rows of the logits the kernel saw for them, in that order); compares
    their gradient rows (un-scaled: x N) and per-utterance likelihoods."""
    from oracle import oracle
    off = np.concatenate([[0], np.cumsum(ly)])
    sub_labels = np.concatenate([labels[off[i]:off[i + 1]] for i in sub]) if len(sub) else np.zeros(0, np.int32)
    oloss, ograd, oparts = oracle.ctc_crf(g, y_sub, sub_labels, lens[sub], ly[sub], lamb, size_average=False,
                                          nthreads=len(sub), fast=True)
    got = grad[torch.tensor(sub)].cpu().numpy() * N
    d = np.abs(got - ograd).max()
    assert np.isfinite(got).all()
    assert d < grad_atol, f"max |grad - oracle| = {d}"
    if parts is not None:
        p = parts.cpu().numpy()
        np.testing.assert_allclose(p[:N][sub], oparts["logz_alpha"], rtol=LOSS_RTOL)
        np.testing.assert_allclose(p[N:][sub], oparts["logp_ctc"], rtol=LOSS_RTOL)
    return d, oparts


def _parts(y_t, labels, lens, ly, lamb):
    import ctc_crf
    _, _, parts = ctc_crf._C.ctc_crf_loss_fwd(y_t.contiguous(), torch.tensor(labels, dtype=torch.int32),
                                              torch.tensor(lens, dtype=torch.int32), torch.tensor(ly, dtype=torch.int32),
                                              lamb, True, want_parts=True)
    return parts


def test_headline_n64_t1500_v218(graphs):
    """BASELINE.json metric shape: N=64, T=1500, V=218, 1.02 M-arc graph, fp32."""
    from oracle import oracle
    V, N, T, lamb = 218, 64, 1500, 0.01
    path, g = graphs(20000, V)
    ctx = _ctx(path)
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=1234)
    yt = torch.tensor(y, device="cuda")
    loss, grad = _run(yt, labels, lens, ly, lamb)
    assert np.isfinite(loss)
    rows = grad.sum(-1) * N + lamb          # sum_k (gamma_den - (1+lamb) gamma_ctc) = -lamb on every valid frame
    assert float(rows.abs().max()) < 2e-3
    d, _ = _check_subset(g, y[[0, 31, N - 1]], labels, lens, ly, lamb, [0, 31, N - 1], loss, grad, N, _parts(yt, labels, lens, ly, lamb))
    print("headline max |grad - oracle| (x N):", d)
    del ctx


def test_config3_bf16_v72(graphs):
    """Config 3: V=72, N=64, T=1500, bf16 logits / fp32 accumulation; the oracle is fed the bf16-rounded values."""
    from oracle import oracle
    V, N, T, lamb = 72, 64, 1500, 0.01
    path, g = graphs(20000, V)
    ctx = _ctx(path)
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=1234)
    yb = torch.tensor(y, device="cuda").bfloat16()
    y_used = yb.float().cpu().numpy()
    loss, grad = _run(yb, labels, lens, ly, lamb)
    assert np.isfinite(loss)
    d, _ = _check_subset(g, y_used[[0, N - 1]], labels, lens, ly, lamb, [0, N - 1], loss, grad, N, _parts(yb, labels, lens, ly, lamb),
                         grad_atol=GRAD_ATOL + 2.0 ** -8)   # the module returns the gradient in the input's dtype: bf16 rounding is up to 2^-8 relative (half an ulp)
    print("config 3 (bf16) max |grad - oracle| (x N):", d)
    # the fp32-held gradient of the same bf16 inputs (native entry) meets the fp32 tolerance
    import ctc_crf
    _, g32, _ = ctc_crf._C.ctc_crf_loss_fwd(yb.contiguous(), torch.tensor(labels, dtype=torch.int32), torch.tensor(lens, dtype=torch.int32),
                                            torch.tensor(ly, dtype=torch.int32), lamb, True)
    _check_subset(g, y_used[[0, N - 1]], labels, lens, ly, lamb, [0, N - 1], loss, g32, N)
    del ctx


def test_config4_5m_arc_graph(graphs):
    """Config 4's graph (H=100 000 -> 5.09 M arcs, S=199 999) at the per-GPU share of N=128 over 8 GPUs (16 utterances),
    natural tier selection (no env hook): the arc stream does not fit shared memory."""
    from oracle import oracle
    V, N, T, lamb = 218, 16, 400, 0.01
    path, g = graphs(100000, V)
    assert g.num_arcs > 5_000_000
    ctx = _ctx(path)
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=1234)
    yt = torch.tensor(y, device="cuda")
    loss, grad = _run(yt, labels, lens, ly, lamb)
    assert np.isfinite(loss)
    d, _ = _check_subset(g, y[[0, N - 1]], labels, lens, ly, lamb, [0, N - 1], loss, grad, N, _parts(yt, labels, lens, ly, lamb))
    print("config 4 (5M arcs) max |grad - oracle| (x N):", d)
    del ctx


def test_config5_varlen_n256(graphs):
    """Config 5: N=256, len ~ U{200..3000} sorted descending (cat/shared/data.py:397-412), V=218, 1.02 M-arc graph.
    The oracle checks the longest and the shortest utterance (and one in the middle)."""
    from oracle import oracle
    V, N, lamb = 218, 256, 0.01
    path, g = graphs(20000, V)
    ctx = _ctx(path)
    rng = np.random.default_rng(1234)
    lens = np.sort(rng.integers(200, 3001, size=N))[::-1].astype(np.int32)
    T = int(lens[0])
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=1234, lens=lens)
    yt = torch.tensor(y, device="cuda")
    del y
    loss, grad = _run(yt, labels, lens, ly, lamb)
    assert np.isfinite(loss)
    for n in (0, N // 2, N - 1):     # frames past each length stay zero
        assert not bool(grad[n, lens[n]:].any())
    sub = [0, N // 2, N - 1]
    y_sub = np.stack([yt[i].cpu().numpy() for i in sub])      # only the checked rows come back to the host
    d, _ = _check_subset(g, y_sub, labels, lens, ly, lamb, sub, loss, grad, N)
    print("config 5 (N=256 var-len) max |grad - oracle| (x N):", d)
    del ctx


def test_config1_warp_ctc_yesno():
    """Config 1: CTC-only loss (no den graph), N=4, T=100, V=5, L=8 through WARP_CTC_LOSS, against
    torch.nn.functional.ctc_loss in fp64 (loss and the gradient chained through log_softmax) and a central-difference
    check of the analytic gradient."""
    import ctc_crf
    N, T, V, L = 4, 100, 5, 8
    gen = torch.Generator().manual_seed(1234)
    z = (3 * torch.randn(N, T, V, generator=gen)).cuda()
    labels = torch.randint(1, V, (N * L,), generator=gen, dtype=torch.int32)
    lx = torch.tensor([100, 93, 71, 40], dtype=torch.int32)
    ly = torch.full((N,), L, dtype=torch.int32)
    crit = ctc_crf.WARP_CTC_LOSS()
    zz = z.clone().requires_grad_(True)
    loss = crit(zz.log_softmax(-1), labels, lx, ly)
    loss.backward()
    z64 = z.double().cpu().requires_grad_(True)
    ref = torch.nn.functional.ctc_loss(z64.log_softmax(-1).transpose(0, 1), labels.view(N, L).long(), lx.long(), ly.long(),
                                       blank=0, reduction="sum") / N
    ref.backward()
    assert abs(float(loss.item()) - float(ref.item())) <= LOSS_RTOL * max(1.0, abs(float(ref.item())))
    assert float((zz.grad.cpu().double() - z64.grad).abs().max()) < GRAD_ATOL
    for n in range(N):
        assert not bool(zz.grad[n, int(lx[n]):].any())
    # directional central difference of the op itself (fp32): d loss(y + eps d) / d eps at 0 == <grad_y, d>
    y = z.log_softmax(-1)
    yy = y.clone().requires_grad_(True)
    crit(yy, labels, lx, ly).backward()
    dirn = torch.randn(y.shape, generator=torch.Generator().manual_seed(5)).cuda()
    eps = 1e-2
    with torch.no_grad():
        fd = (float(crit(y + eps * dirn, labels, lx, ly).item()) - float(crit(y - eps * dirn, labels, lx, ly).item())) / (2 * eps)
    an = float((yy.grad * dirn).sum().item())
    assert abs(fd - an) < 2e-2 * max(1.0, abs(an)), (fd, an)


@pytest.mark.parametrize("kind", ["randn20", "one_hot_floor60"])
def test_peaky_logits(graphs, kind):
    """Trained-model-like, peaky posteriors: (a) x = 20 randn (gaps of ~60-120 nats between the best label and the rest),
    (b) half of the frames clamped to one-hot rows with -60 floors.  Scaled-linear arithmetic must neither underflow to
    logZ = -inf nor lose the tolerance."""
    from oracle import oracle
    V, N, T, lamb = 218, 8, 300, 0.01
    path, g = graphs(20000, V)
    ctx = _ctx(path)
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=77, lens=[300, 300, 280, 250, 200, 120, 60, 17],
                                             scale=20.0 if kind == "randn20" else 3.0)
    if kind == "one_hot_floor60":
        rng = np.random.default_rng(3)
        hot = rng.integers(0, V, size=(N, T))
        mask = rng.random((N, T)) < 0.5
        yh = np.full((N, T, V), -60.0, np.float32)
        np.put_along_axis(yh, hot[..., None], np.float32(np.log1p(-(V - 1) * np.exp(-60.0))), -1)
        y = np.where(mask[..., None], yh, y).astype(np.float32)
    yt = torch.tensor(y, device="cuda")
    loss, grad = _run(yt, labels, lens, ly, lamb)
    assert np.isfinite(loss), loss
    assert bool(torch.isfinite(grad).all())
    parts = _parts(yt, labels, lens, ly, lamb)
    sub = [0, 3, N - 1]
    d, oparts = _check_subset(g, y[sub], labels, lens, ly, lamb, sub, loss, grad, N, parts)
    print(f"peaky ({kind}) max |grad - oracle| (x N):", d, "logZ:", oparts["logz_alpha"])
    del ctx


def test_few_final_states(graphs):
    """A den graph where only 6 of 39 999 states are final, with moderately peaky logits: the mass that reaches a final
    state at t = len is a tiny fraction of the last column (about 2^-130 of it here).  The two shortest utterances cannot
    reach any final state at all: their logZ is -inf in the oracle too, and they must not poison the others."""
    from oracle import oracle
    import ctc_crf
    V, N, T, lamb = 218, 8, 200, 0.01
    path, g = graphs(20000, V, few_finals=True)
    assert int((~np.isinf(g.final)).sum()) == 6
    ctx = _ctx(path)
    y, labels, lens, ly = oracle.synth_batch(N, T, V, seed=78, lens=[200, 200, 150, 99, 64, 33, 21, 17], scale=8.0)
    yt = torch.tensor(y, device="cuda")
    loss, grad = _run(yt, labels, lens, ly, lamb)
    assert np.isfinite(loss), loss
    parts = _parts(yt, labels, lens, ly, lamb)
    d, oparts = _check_subset(g, y[[0, 4, 6, N - 1]], labels, lens, ly, lamb, [0, 4, 6, N - 1], loss, grad, N, parts)
    print("few finals max |grad - oracle| (x N):", d, "logZ:", oparts["logz_alpha"])
    # no path of 2 frames ends in a final state: logZ = -inf on both sides, occupancies all zero, nothing else disturbed
    lens2 = np.array([40, 2, 37], np.int32)
    y2, _, lens2, _ = oracle.synth_batch(3, 40, V, seed=79, lens=lens2, scale=8.0)
    la, lb, gd = oracle.den(g, y2, lens2, fast=True)
    assert np.isinf(la[1]) and la[1] < 0 and np.isfinite(la[[0, 2]]).all()
    lg = torch.tensor(y2, device="cuda")
    gden = torch.zeros_like(lg)
    ca, cb = torch.zeros(3, device="cuda"), torch.zeros(3, device="cuda")
    ctc_crf._C.gpu_den(lg, gden, torch.tensor(lens2).cuda(), ca, cb)
    ca, cb = ca.cpu().numpy(), cb.cpu().numpy()
    assert np.isinf(ca[1]) and ca[1] < 0 and np.isinf(cb[1]) and cb[1] < 0
    np.testing.assert_allclose(ca[[0, 2]], la[[0, 2]], rtol=LOSS_RTOL)
    np.testing.assert_allclose(cb[[0, 2]], lb[[0, 2]], rtol=LOSS_RTOL)
    gn = gden.cpu().numpy()
    assert np.isfinite(gn).all() and not gn[1].any()
    assert np.abs(gn[[0, 2]] - gd[[0, 2]]).max() < GRAD_ATOL
    del ctx
