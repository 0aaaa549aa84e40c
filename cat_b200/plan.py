"""Host-side view of the den-graph kernel plan (``ccb_plan_*``): no GPU needed.

Used by the CPU tests (a numpy emulation of the kernels' arithmetic runs on these exact arrays) and by
tooling that wants to inspect how a den graph is cut for the persistent grid.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib

QUAD = 4            # rows are padded to whole quads of arcs (den_graph.h kQuad)
CHUNK_ARC_PAD = 16  # chunk arc counts are padded to a multiple of this (kChunkArcPad)
ARC_DTYPE = np.dtype([("peer", "<u4"), ("w", "<f4")])


@dataclass
class PassView:
    arcs: np.ndarray          # ARC_DTYPE [A]
    chunk_state: np.ndarray   # int32 [n_chunks+1]
    chunk_arc: np.ndarray     # int32 [n_chunks+1]

    def row_ends(self) -> np.ndarray:
        """Arc index one past each row: rows end at quads whose 4th weight has its sign bit set."""
        w4 = self.arcs["w"][QUAD - 1::QUAD]
        return (np.nonzero(np.signbit(w4))[0] + 1) * QUAD

    def row_of_arc(self) -> np.ndarray:
        """Row id of every arc slot (chunk-tail padding quads attach to the following row, weight 0)."""
        ends = self.row_ends()
        return np.searchsorted(ends, np.arange(len(self.arcs)), side="right")

    def weights(self) -> np.ndarray:
        return np.abs(self.arcs["w"])


@dataclass
class PlanView:
    file_states: int
    file_arcs: int
    num_states: int
    start: int
    num_labels: int
    n_ctas: int
    n_warps: int
    max_tile_arcs: int
    state_label: np.ndarray
    final_lin: np.ndarray
    orig_state: np.ndarray
    fwd: PassView
    bwd: PassView


def load_plan(path: str, n_ctas: int = 148, n_warps: int = 16) -> PlanView:
    L = _lib.lib()
    h = L.ccb_plan_create(path.encode(), n_ctas, n_warps)
    if not h:
        raise RuntimeError(_lib.last_error())
    try:
        info = (C.c_long * 10)()
        assert L.ccb_plan_info(h, info) == 0
        S0, A0, S, Af, Ab, start, nl, nc, nw, mta = [int(x) for x in info]
        n_chunks = nc * nw

        def get(which, dtype, count):
            a = np.empty(count, dtype=dtype)
            rc = L.ccb_plan_copy(h, which, a.ctypes.data_as(C.c_void_p), a.nbytes)
            assert rc == 0, f"ccb_plan_copy({which}) -> {rc}"
            return a

        return PlanView(S0, A0, S, start, nl, nc, nw, mta,
                        get(0, np.int32, S), get(1, np.float32, S), get(2, np.int32, S),
                        PassView(get(3, ARC_DTYPE, Af), get(4, np.int32, n_chunks + 1), get(5, np.int32, n_chunks + 1)),
                        PassView(get(6, ARC_DTYPE, Ab), get(7, np.int32, n_chunks + 1), get(8, np.int32, n_chunks + 1)))
    finally:
        L.ccb_plan_destroy(h)
