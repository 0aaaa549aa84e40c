"""Experiment (tuning build only): do the forward pass of one batch and the backward pass of another overlap when the den
kernels run as two co-resident 8-warp cooperative grids (one CTA of each per SM) on two streams?  Bounds what a
forward/backward recursion run concurrently from both ends of the utterances could gain.
   CCB_NVCC_DEFS=-DCCB_TUNING CCB_LIB_NAME=libccb_tune.so python -m cat_b200.build
   CCB_LIB_NAME=libccb_tune.so CCB_DEN_WARPS=8 python tools/overlap_probe.py"""
import argparse, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctc_crf
from cat_b200 import _lib, fst

ap = argparse.ArgumentParser(); ap.add_argument("--T", type=int, default=600); ap.add_argument("--N", type=int, default=64)
a = ap.parse_args()
N, T, V = a.N, a.T, 218
path = "/tmp/ccb_tl_20000.fst"
if not os.path.exists(path): fst.write_fst(path, fst.make_synthetic_den(20000, 24, V, seed=7))
ctx = ctc_crf.CRFContext(path, gpus=0)
L = _lib.lib()
vp = C.c_void_p
L.ccb_debug_den_backward.argtypes = [vp, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_long, C.c_long, vp]
L.ccb_debug_den_backward.restype = C.c_int
y = torch.log_softmax(3 * torch.randn(N, T, V, device="cuda"), -1)
lens = torch.full((N,), T, dtype=torch.int32, device="cuda")
def ws():
    return (torch.empty(int(L.ccb_den_alpha_floats(N, T)), device="cuda"), torch.empty(int(L.ccb_den_aux_bytes(N, T)), dtype=torch.uint8, device="cuda"),
            torch.zeros(N, T, V, device="cuda"), torch.empty(N, device="cuda"))
w1, w2 = ws(), ws()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def fwd(w, s, with_bwd=False):
    rc = L.ccb_den_forward_backward(y.data_ptr(), 0, T * V, V, N, T, V, lens.data_ptr(), w[0].data_ptr(), w[1].data_ptr(),
                                    w[2].data_ptr() if with_bwd else None, T * V, V, 1.0, w[3].data_ptr(), None, s.cuda_stream)
    assert rc == 0, _lib.last_error()
def bwd(w, s):
    rc = L.ccb_debug_den_backward(y.data_ptr(), 0, T * V, V, N, T, V, lens.data_ptr(), w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(), T * V, V, s.cuda_stream)
    assert rc == 0, _lib.last_error()
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fwd(w2, s1, True); torch.cuda.synchronize()
t_f = timed(lambda: fwd(w1, s1))
t_b = timed(lambda: bwd(w2, s2))
def both():
    fwd(w1, s1); bwd(w2, s2)
t_fb = timed(both)
print(f"warps/CTA={os.environ.get('CCB_DEN_WARPS', '16')}  N={N} T={T}: forward alone {t_f:.2f} ms, backward alone {t_b:.2f} ms, "
      f"forward || backward on two streams {t_fb:.2f} ms  (sum {t_f + t_b:.2f})")
