"""CPU: den-graph files and the host-side kernel plan (the product's C++ loader through the C ABI)."""
import numpy as np
import pytest

from cat_b200 import fst, plan, _lib
from oracle import oracle

import emulate


def test_fixture_parses(fixture_fst):
    g = fst.read_fst(fixture_fst)
    assert (g.num_states, g.num_arcs, g.start) == (9, 24, 0)
    assert sorted(np.nonzero(~np.isinf(g.final))[0].tolist()) == [4, 6]
    np.testing.assert_allclose(g.final[[4, 6]], 0.6931472, rtol=1e-6)
    P = plan.load_plan(fixture_fst, 4, 2)
    assert (P.file_states, P.file_arcs, P.num_states) == (9, 24, 9)      # T-compose-LM: no state split
    assert int((P.fwd.weights() > 0).sum()) == 24 and int((P.bwd.weights() > 0).sum()) == 24


def test_roundtrip_and_cxx_reader_agree(tmp_path):
    g = fst.make_synthetic_den(40, 5, 9, seed=1)
    p = str(tmp_path / "a.fst")
    fst.write_fst(p, g)
    g2 = fst.read_fst(p)
    for f in ("src", "dst", "ilabel", "weight", "final"):
        np.testing.assert_array_equal(getattr(g, f)[np.argsort(g.src, kind="stable")] if f not in ("final",) else g.final,
                                      getattr(g2, f))
    P = plan.load_plan(p, 7, 3)
    assert P.file_states == g.num_states and P.file_arcs == g.num_arcs
    # every file state appears, final weights carried over in the linear domain
    assert sorted(set(P.orig_state.tolist())) == list(range(g.num_states))
    np.testing.assert_allclose(P.final_lin, np.exp(g.end_weight()[P.orig_state]), rtol=1e-6)


@pytest.mark.parametrize("bad", ["missing", "garbage", "truncated", "wrongtype"])
def test_bad_files_raise_not_exit(tmp_path, fixture_fst, bad):
    p = tmp_path / "x.fst"
    raw = open(fixture_fst, "rb").read()
    if bad == "garbage":
        p.write_bytes(b"\x00" * 64)
    elif bad == "truncated":
        p.write_bytes(raw[:200])
    elif bad == "wrongtype":
        p.write_bytes(raw.replace(b"standard", b"log64xyz"))
    with pytest.raises(RuntimeError):
        plan.load_plan(str(p), 4, 2)
    assert _lib.last_error()
    if bad != "missing":
        with pytest.raises(ValueError):
            fst.read_fst(str(p))


@pytest.mark.parametrize("n_ctas,n_warps", [(1, 1), (4, 2), (148, 16), (148, 32)])
def test_plan_invariants(tmp_graphs, n_ctas, n_warps):
    for name in ("tlm_small", "random_split", "tlm_mid"):
        path, g, V = tmp_graphs[name]
        P = plan.load_plan(path, n_ctas, n_warps)
        S = P.num_states
        assert (np.diff(P.state_label) >= 0).all()                      # states sorted by label
        assert P.num_labels <= V
        for pv, real in ((P.fwd, None), (P.bwd, None)):
            ends = pv.row_ends()
            assert len(ends) == S                                       # one flagged quad per row
            assert len(pv.arcs) % plan.QUAD == 0
            assert pv.chunk_state[0] == 0 and pv.chunk_state[-1] == S
            assert (np.diff(pv.chunk_state) >= 0).all()
            assert (pv.chunk_arc % plan.CHUNK_ARC_PAD == 0).all() and pv.chunk_arc[-1] == len(pv.arcs)
            # a chunk's arc range holds exactly its rows
            for c in range(0, len(pv.chunk_state) - 1, max(1, (len(pv.chunk_state) - 1) // 97)):
                s0, s1 = pv.chunk_state[c], pv.chunk_state[c + 1]
                if s1 > s0:
                    assert (s0 == 0 or ends[s0 - 1] <= pv.chunk_arc[c]) and pv.chunk_arc[c] < ends[s0]
                    assert pv.chunk_arc[c + 1] - plan.CHUNK_ARC_PAD < ends[s1 - 1] <= pv.chunk_arc[c + 1]
                else:
                    assert pv.chunk_arc[c] == pv.chunk_arc[c + 1]
            assert (pv.arcs["peer"] < S).all()
            assert (pv.weights() >= 0).all()
            # the sign flag appears only on quad-final slots
            assert not np.signbit(pv.arcs["w"].reshape(-1, plan.QUAD)[:, :-1]).any()
        # arcs survive padding: the multiset of (row, peer, w>0) has the graph's size (x copies for split states)
        nz = int((P.fwd.weights() > 0).sum())
        if name == "random_split":
            assert S > g.num_states and nz >= g.num_arcs
        else:
            assert S == g.num_states and nz == g.num_arcs == int((P.bwd.weights() > 0).sum())


def test_plan_balance(tmp_path):
    g = fst.make_synthetic_den(4000, 24, 60, seed=7)
    p = str(tmp_path / "g.fst")
    fst.write_fst(p, g)
    P = plan.load_plan(p, 148, 16)
    for pv in (P.fwd, P.bwd):
        per_cta = np.diff(pv.chunk_arc[::16])
        assert per_cta.max() <= 1.25 * per_cta.mean() + 64
    assert P.max_tile_arcs * 8 < 200 * 1024


@pytest.mark.parametrize("name,lens", [("tlm_small", [30, 22, 9, 1]), ("random_split", [20, 13, 7, 2])])
def test_kernel_arithmetic_emulation_matches_oracle(tmp_graphs, name, lens):
    """The scaled-linear / hoisted-emission / state-product algorithm the kernels implement, emulated in numpy on
    the product's own plan arrays, equals the arc-based log-domain reference semantics."""
    path, g, V = tmp_graphs[name]
    P = plan.load_plan(path, 148, 16)
    y, _, lens, _ = oracle.synth_batch(len(lens), max(lens), V, seed=6, lens=lens)
    la, lb, gd = oracle.den(g, y, lens)
    ea, eb, eg = emulate.den_emulate(P, y, lens)
    np.testing.assert_allclose(ea, la, rtol=1e-7)
    np.testing.assert_allclose(eb, lb, rtol=1e-7)
    assert np.abs(eg - gd).max() < 1e-6


def test_fixture_emulation(fixture_fst, fixture_inputs):
    P = plan.load_plan(fixture_fst, 148, 16)
    ea, eb, eg = emulate.den_emulate(P, fixture_inputs["y"], fixture_inputs["lx"])
    assert abs(ea[0] + 6.25832785) < 1e-6 and abs(eb[0] + 6.25832785) < 1e-6
