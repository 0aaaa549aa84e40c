"""Regenerates tests/golden/den_lm_fixture.fst from the reference's test fixture.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The den graph of src/ctc_crf/test/den_lm.fst (9 states / 24 arcs) is parsed with cat_b200.fst.read_fst and
re-emitted with cat_b200.fst.write_fst (the reference file is data, not source; the copy lets the GPU box,
which has no /root/reference, run the fixture test)."""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cat_b200 import fst  # noqa: E402

SRC = "/root/reference/src/ctc_crf/test/den_lm.fst"
DST = os.path.join(ROOT, "tests", "golden", "den_lm_fixture.fst")

if __name__ == "__main__":
    g = fst.read_fst(SRC)
    raw = open(SRC, "rb").read()
    (props,) = struct.unpack_from("<Q", raw, 4 + 4 + 6 + 4 + 8 + 8)
    fst.write_fst(DST, g, properties=props)
    g2 = fst.read_fst(DST)
    assert g2.num_states == 9 and g2.num_arcs == 24 and g2.start == 0
    print("wrote", DST, os.path.getsize(DST), "bytes")
