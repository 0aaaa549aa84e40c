"""Where does the host time of one fused call go? (diagnostic)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctc_crf
from cat_b200 import _C, _lib, fst
N, T, V = 64, 1500, 218
path = "/tmp/ccb_tl_20000.fst"
if not os.path.exists(path): fst.write_fst(path, fst.make_synthetic_den(20000, 24, V, seed=7))
ctx = ctc_crf.CRFContext(path, gpus=0)
y = torch.log_softmax(3 * torch.randn(N, T, V, device="cuda"), -1)
lx = torch.full((N,), T, dtype=torch.int32); ly = torch.full((N,), 250, dtype=torch.int32)
labels = torch.randint(1, V, (int(ly.sum()),), dtype=torch.int32)
L = _lib.lib()
def t(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = (time.perf_counter() - t0) / n * 1e3; torch.cuda.synchronize(); return dt
for _ in range(2): _C.ctc_crf_loss_fwd(y, labels, lx, ly, 0.01, True)
print("full call host ms (GPU idle at start, 5 back-to-back):", round(t(lambda: _C.ctc_crf_loss_fwd(y, labels, lx, ly, 0.01, True)), 2))
print("mem_get_info ms:", round(t(lambda: torch.cuda.mem_get_info(0), 20), 3))
print("upload_meta ms:", round(t(lambda: _C.upload_meta(labels, lx, ly, y.device), 20), 3))
print("plan_slices ms:", round(t(lambda: _C._plan_slices(L, lx, ly, N, V, y.device), 20), 3))
a = int(L.ccb_den_alpha_floats(N, T))
print("torch.empty(alpha) ms:", round(t(lambda: torch.empty(a, dtype=torch.float32, device="cuda"), 20), 3))
# native pieces, GPU busy
alpha = torch.empty(a, device="cuda"); aux = torch.empty(int(L.ccb_den_aux_bytes(N, T)), dtype=torch.uint8, device="cuda")
lens = lx.cuda(); logz = torch.empty(N, device="cuda"); grad = torch.zeros(N, T, V, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def den(bwd):
    rc = L.ccb_den_forward_backward(y.data_ptr(), 0, T * V, V, N, T, V, lens.data_ptr(), alpha.data_ptr(), aux.data_ptr(), grad.data_ptr() if bwd else None, T * V, V, 1.0, logz.data_ptr(), None, st)
    assert rc == 0
for bwd in (False, True):
    torch.cuda.synchronize(); ts = []
    for i in range(4):
        t0 = time.perf_counter(); den(bwd); ts.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize(); print("den call (bwd=%s) host ms per consecutive call:" % bwd, [round(x, 2) for x in ts])
