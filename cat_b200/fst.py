"""Denominator-graph files: OpenFst binary ``vector``/``standard`` reader/writer (numpy)
and the synthetic T-compose-LM shaped generator used by tests and bench.py.

The product's loader is the C++ one in ``csrc/den_graph.cc`` (used by ``CRFContext``);
this module is the host-side tool for *producing* den graphs (tests, benchmarks) and an
independent second parser the tests cross-check the C++ reader against.

Reference semantics being mirrored (nothing is copied, OpenFst is not available here):
  * ``src/ctc_crf/gpu_den/fst_read.cc:40-60`` -- weight = -tropical, label = ilabel-1,
    start weight 0 on ``Start()``, end weight = -Final for non-Zero finals.
  * ``cat/utils/tool/build_ctc_topo.py:47-60`` -- CTC topology T: every arc *into* token
    state i carries ilabel i+1, arcs into the blank state carry ilabel 1.
File layout (little endian), decoded from ``src/ctc_crf/test/den_lm.fst``:
  int32 magic 0x7EB2FDD6 | str "vector" | str "standard" | int32 version(2) | int32 flags |
  uint64 properties | int64 start | int64 num_states | int64 num_arcs |
  per state: float32 final (inf = non-final), int64 narcs,
             narcs x {int32 ilabel, int32 olabel, float32 weight, int32 nextstate}
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

FST_MAGIC = 0x7EB2FDD6
# kExpanded | kMutable, and "not known" for everything else is acceptable to OpenFst readers
_DEFAULT_PROPERTIES = 0x0000000000000003


@dataclass
class DenGraph:
    """Arc list of a den graph in *file* convention (tropical weights, ilabel as stored)."""
    num_states: int
    start: int
    src: np.ndarray       # int32 [A]
    dst: np.ndarray       # int32 [A]
    ilabel: np.ndarray    # int32 [A]  (token id + 1)
    olabel: np.ndarray    # int32 [A]
    weight: np.ndarray    # float32 [A] tropical (-log p)
    final: np.ndarray     # float32 [S] tropical, +inf = non-final

    @property
    def num_arcs(self) -> int:
        return int(self.src.shape[0])

    # --- the view the loss uses (fst_read.cc:43-59) ---------------------------------
    def log_arcs(self):
        """(src, dst, label, logw) with label = ilabel-1 and logw = -weight."""
        return (self.src.astype(np.int32), self.dst.astype(np.int32),
                (self.ilabel - 1).astype(np.int32), (-self.weight).astype(np.float32))

    def start_weight(self) -> np.ndarray:
        w = np.full(self.num_states, -np.inf, dtype=np.float32)
        w[self.start] = 0.0
        return w

    def end_weight(self) -> np.ndarray:
        w = np.where(np.isinf(self.final), -np.inf, -self.final).astype(np.float32)
        return w


def _read_str(buf: bytes, off: int):
    (n,) = struct.unpack_from("<i", buf, off)
    off += 4
    return buf[off:off + n].decode("ascii"), off + n


def read_fst(path: str) -> DenGraph:
    buf = open(path, "rb").read()
    off = 0
    (magic,) = struct.unpack_from("<i", buf, off)
    off += 4
    if (magic & 0xFFFFFFFF) != FST_MAGIC:
        raise ValueError(f"{path}: not an OpenFst binary file (magic {magic & 0xFFFFFFFF:#x})")
    fst_type, off = _read_str(buf, off)
    arc_type, off = _read_str(buf, off)
    if fst_type != "vector" or arc_type != "standard":
        raise ValueError(f"{path}: unsupported fst/arc type {fst_type}/{arc_type}")
    version, flags = struct.unpack_from("<ii", buf, off)
    off += 8
    if flags & 0x3:
        raise ValueError(f"{path}: embedded symbol tables are not supported (flags={flags})")
    props, start, num_states, _num_arcs_hdr = struct.unpack_from("<Qqqq", buf, off)
    off += 32
    final = np.empty(num_states, dtype=np.float32)
    srcs, dsts, ils, ols, ws = [], [], [], [], []
    arc_dt = np.dtype([("il", "<i4"), ("ol", "<i4"), ("w", "<f4"), ("ns", "<i4")])
    for s in range(num_states):
        (f, narcs) = struct.unpack_from("<fq", buf, off)
        off += 12
        final[s] = f
        if narcs:
            a = np.frombuffer(buf, dtype=arc_dt, count=narcs, offset=off)
            off += 16 * narcs
            srcs.append(np.full(narcs, s, dtype=np.int32))
            dsts.append(a["ns"].astype(np.int32))
            ils.append(a["il"].astype(np.int32))
            ols.append(a["ol"].astype(np.int32))
            ws.append(a["w"].astype(np.float32))
    if off != len(buf):
        raise ValueError(f"{path}: trailing bytes ({len(buf) - off})")
    cat = (lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dtype=dt))
    return DenGraph(num_states=int(num_states), start=int(start),
                    src=cat(srcs, np.int32), dst=cat(dsts, np.int32),
                    ilabel=cat(ils, np.int32), olabel=cat(ols, np.int32),
                    weight=cat(ws, np.float32), final=final)


def write_fst(path: str, g: DenGraph, properties: int = _DEFAULT_PROPERTIES) -> None:
    order = np.argsort(g.src, kind="stable")
    src = g.src[order]
    counts = np.bincount(src, minlength=g.num_states).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)])
    arc_dt = np.dtype([("il", "<i4"), ("ol", "<i4"), ("w", "<f4"), ("ns", "<i4")])
    arcs = np.empty(g.num_arcs, dtype=arc_dt)
    arcs["il"] = g.ilabel[order]
    arcs["ol"] = g.olabel[order]
    arcs["w"] = g.weight[order]
    arcs["ns"] = g.dst[order]
    with open(path, "wb") as f:
        f.write(struct.pack("<I", FST_MAGIC))
        for s in (b"vector", b"standard"):
            f.write(struct.pack("<i", len(s)) + s)
        f.write(struct.pack("<ii", 2, 0))
        f.write(struct.pack("<Qqqq", properties, g.start, g.num_states, g.num_arcs))
        for s in range(g.num_states):
            f.write(struct.pack("<fq", float(g.final[s]), int(counts[s])))
            f.write(arcs[offs[s]:offs[s + 1]].tobytes())


def make_synthetic_den(H: int, d: int, V: int, seed: int = 7, p_final: float = 0.05) -> DenGraph:
    """T-compose-LM shaped den graph.

    LM: H history states, state 0 = start (no last label), every h>0 has a last label
    l(h) in [1,V-1]; each h has d distinct successors h'!=0 with Dirichlet(1) probabilities
    (+ p_final for h>0).  Den states: (h,B) for all h, (h,L) for h>0  => S = 2H-1.
    Arcs (weight 0 unless stated): (h,L)->(h,L) label l(h); (h,L)->(h,B) blank;
    (h,B)->(h,B) blank; per LM arc h->h' (weight w): (h,B)->(h',L) and, when
    l(h') != l(h), (h,L)->(h',L), both with label l(h').  Finals on both twins.
    """
    rng = np.random.default_rng(seed)
    assert H >= 2 and V >= 2 and 1 <= d <= H - 1
    last = np.zeros(H, dtype=np.int32)
    last[1:] = rng.integers(1, V, size=H - 1)
    # successors: d distinct h' in [1,H-1] for each h (vectorised rejection-free draw)
    succ = rng.integers(1, H, size=(H, d), dtype=np.int64)
    while True:   # re-draw the (rare) rows that contain a duplicate successor
        srt = np.sort(succ, axis=1)
        bad = np.nonzero((srt[:, 1:] == srt[:, :-1]).any(axis=1))[0]
        if bad.size == 0:
            break
        succ[bad] = rng.integers(1, H, size=(bad.size, d), dtype=np.int64)
    probs = rng.dirichlet(np.ones(d), size=H)
    pf = np.where(np.arange(H) > 0, p_final, 0.0)
    probs = probs * (1.0 - pf)[:, None]
    w_lm = (-np.log(np.maximum(probs, 1e-30))).astype(np.float32)

    B = lambda h: 2 * h - (h > 0)          # (0,B)=0, (h,B)=2h-1
    L = lambda h: 2 * h                    # (h,L)=2h  for h>0
    S = 2 * H - 1
    hs = np.arange(H)
    hpos = np.arange(1, H)
    src, dst, lab, w = [], [], [], []
    # self loops / blank arcs
    src += [L(hpos), L(hpos), B(hs)]
    dst += [L(hpos), B(hpos), B(hs)]
    lab += [last[hpos], np.zeros(H - 1, np.int32), np.zeros(H, np.int32)]
    w += [np.zeros(H - 1, np.float32), np.zeros(H - 1, np.float32), np.zeros(H, np.float32)]
    # LM arcs
    hh = np.repeat(hs, d)
    hn = succ.reshape(-1)
    ww = w_lm.reshape(-1)
    src.append(B(hh)); dst.append(L(hn)); lab.append(last[hn]); w.append(ww)
    m = (hh > 0) & (last[hn] != last[hh])
    src.append(L(hh[m])); dst.append(L(hn[m])); lab.append(last[hn[m]]); w.append(ww[m])

    final = np.full(S, np.inf, dtype=np.float32)
    fw = np.float32(-np.log(p_final))
    final[B(hpos)] = fw
    final[L(hpos)] = fw
    src = np.concatenate(src).astype(np.int32)
    dst = np.concatenate(dst).astype(np.int32)
    lab = np.concatenate(lab).astype(np.int32)
    return DenGraph(num_states=S, start=0, src=src, dst=dst, ilabel=lab + 1, olabel=lab.copy(),
                    weight=np.concatenate(w).astype(np.float32), final=final)


def make_random_den(S: int, A: int, V: int, seed: int = 0, n_final: int = 2) -> DenGraph:
    """Unstructured random graph: in-arcs of a state may carry *different* labels
    (exercises the loader's state-splitting path; not T-compose-LM shaped)."""
    rng = np.random.default_rng(seed)
    src = rng.integers(0, S, size=A).astype(np.int32)
    dst = rng.integers(0, S, size=A).astype(np.int32)
    # keep the graph connected enough: a chain 0->1->...->S-1
    src[:S - 1] = np.arange(S - 1)
    dst[:S - 1] = np.arange(1, S)
    lab = rng.integers(0, V, size=A).astype(np.int32)
    w = rng.uniform(0.1, 3.0, size=A).astype(np.float32)
    final = np.full(S, np.inf, dtype=np.float32)
    final[rng.choice(S, size=min(n_final, S), replace=False)] = 0.7
    final[S - 1] = 0.3
    return DenGraph(num_states=S, start=0, src=src, dst=dst, ilabel=lab + 1, olabel=lab.copy(),
                    weight=w, final=final)


# ---------------------------------------------------------------------------------------------------
# Den-graph construction without Kaldi/OpenFst: T o LM for an epsilon-free phone LM
# ---------------------------------------------------------------------------------------------------
@dataclass
class PhoneLm:
    """Epsilon-free weighted acceptor over phone ids 1..V-1 (0 is the CTC blank and never appears):
    what ``chain-est-phone-lm`` produces for ``cat/utils/tool/prep_den_lm.sh:40-45`` (no back-off arcs)."""
    num_states: int
    start: int
    src: np.ndarray      # int32 [A]
    dst: np.ndarray      # int32 [A]
    phone: np.ndarray    # int32 [A] in [1, V-1]
    weight: np.ndarray   # float32 [A] tropical (-log p)
    final: np.ndarray    # float32 [S] tropical, +inf = non-final


def compose_ctc_lm(lm: PhoneLm) -> DenGraph:
    """den graph = T o LM, with T the CTC topology of ``cat/utils/tool/build_ctc_topo.py:47-66`` (state 0 = blank,
    state i = token i; every arc INTO token state i reads label i and emits phone i, arcs into the blank state and
    token self loops emit nothing), i.e. the composition ``prep_den_lm.sh:46-51`` performs with OpenFst.

    States of the result: (h,B) "LM state h, last frame was blank or nothing yet" for every reachable h, and (h,i)
    "LM state h reached by phone i, last frame was i" for every phone i entering h.  Because T emits a phone exactly
    when a token state is entered, the composition needs no epsilon handling for an epsilon-free LM, and for a
    deterministic LM the result is already deterministic (what fstdeterminizestar would return up to state order)."""
    S, A = lm.num_states, int(lm.src.shape[0])
    # token states: one per distinct (destination, phone)
    key = lm.dst.astype(np.int64) * (int(lm.phone.max(initial=0)) + 1) + lm.phone
    uniq, tok_of_arc = np.unique(key, return_inverse=True)
    n_tok = int(uniq.shape[0])
    tok_h = (uniq // (int(lm.phone.max(initial=0)) + 1)).astype(np.int32)
    tok_p = (uniq % (int(lm.phone.max(initial=0)) + 1)).astype(np.int32)
    B = lambda h: np.asarray(h, np.int64)                   # (h,B) -> h
    TK = lambda k: S + np.asarray(k, np.int64)             # token state k -> S + k
    src, dst, lab, w = [], [], [], []
    hs = np.arange(S)
    ks = np.arange(n_tok)
    # blank self loop, token self loop, token -> blank
    src += [B(hs), TK(ks), TK(ks)]
    dst += [B(hs), TK(ks), B(tok_h)]
    lab += [np.zeros(S, np.int32), tok_p, np.zeros(n_tok, np.int32)]
    w += [np.zeros(S, np.float32), np.zeros(n_tok, np.float32), np.zeros(n_tok, np.float32)]
    # per LM arc h -p-> h': from (h,B), and from every token state (h,i) with i != p
    src.append(B(lm.src)); dst.append(TK(tok_of_arc)); lab.append(lm.phone); w.append(lm.weight)
    by_h = [[] for _ in range(S)]
    for k in range(n_tok):
        by_h[int(tok_h[k])].append(k)
    s2, d2, l2, w2 = [], [], [], []
    for a in range(A):
        for k in by_h[int(lm.src[a])]:
            if tok_p[k] != lm.phone[a]:
                s2.append(S + k); d2.append(S + int(tok_of_arc[a])); l2.append(int(lm.phone[a])); w2.append(float(lm.weight[a]))
    src.append(np.asarray(s2, np.int64)); dst.append(np.asarray(d2, np.int64))
    lab.append(np.asarray(l2, np.int32)); w.append(np.asarray(w2, np.float32))
    src = np.concatenate(src); dst = np.concatenate(dst)
    lab = np.concatenate(lab).astype(np.int32); w = np.concatenate(w).astype(np.float32)
    final = np.concatenate([lm.final, lm.final[tok_h]]).astype(np.float32)
    # keep what is reachable from (start,B), renumbered in discovery order
    n_all = S + n_tok
    order = np.argsort(src, kind="stable")
    starts = np.searchsorted(src[order], np.arange(n_all + 1))
    new_id = np.full(n_all, -1, np.int64)
    new_id[lm.start] = 0
    stack, n_new = [int(lm.start)], 1
    while stack:
        q = stack.pop()
        for a in order[starts[q]:starts[q + 1]]:
            t = int(dst[a])
            if new_id[t] < 0:
                new_id[t] = n_new; n_new += 1; stack.append(t)
    keep = new_id[src] >= 0
    inv = np.argsort(np.where(new_id >= 0, new_id, n_all), kind="stable")[:n_new]
    return DenGraph(num_states=n_new, start=0, src=new_id[src[keep]].astype(np.int32), dst=new_id[dst[keep]].astype(np.int32),
                    ilabel=(lab[keep] + 1).astype(np.int32), olabel=lab[keep].copy(), weight=w[keep], final=final[inv])


def lm_logprob(lm: PhoneLm, phones) -> float:
    """log p_LM(l) of one phone sequence under a deterministic LM (the per-utterance path weight CAT adds to the
    numerator, docs/toolkitworkflow.md:124-135); -inf if the LM does not accept it."""
    h, lp = int(lm.start), 0.0
    for p in phones:
        m = np.nonzero((lm.src == h) & (lm.phone == int(p)))[0]
        if m.size == 0:
            return float("-inf")
        assert m.size == 1, "lm_logprob needs a deterministic LM"
        lp -= float(lm.weight[m[0]]); h = int(lm.dst[m[0]])
    return lp - float(lm.final[h]) if np.isfinite(lm.final[h]) else float("-inf")


def make_random_lm(H: int, V: int, d: int, seed: int = 0, p_final: float = 0.2) -> PhoneLm:
    """Small deterministic phone LM for tests: every state has d out-arcs with distinct phones, random targets."""
    rng = np.random.default_rng(seed)
    assert 1 <= d <= V - 1
    src, dst, ph, w = [], [], [], []
    final = np.full(H, np.inf, np.float32)
    for h in range(H):
        fin = h > 0 and rng.random() < 0.7
        pr = rng.dirichlet(np.ones(d)) * (1.0 - (p_final if fin else 0.0))
        for p, q in zip(rng.choice(np.arange(1, V), size=d, replace=False), pr):
            src.append(h); dst.append(int(rng.integers(0, H))); ph.append(int(p)); w.append(-np.log(q))
        if fin:
            final[h] = -np.log(p_final)
    return PhoneLm(H, 0, np.asarray(src, np.int32), np.asarray(dst, np.int32), np.asarray(ph, np.int32),
                   np.asarray(w, np.float32), final)
