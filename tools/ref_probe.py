"""Probe the reference CUDA builds on the GPU box: CCB_REF_VARIANT=unmodified|sm100fix python tools/ref_probe.py den|ctc"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle, ref_cuda
from cat_b200 import fst
which = sys.argv[1]
print("variant", ref_cuda.VARIANT, which)
g = fst.read_fst("tests/golden/golden_tlm_a.fst")
y, labels, lens, ly = oracle.synth_batch(4, 40, 12, seed=11, lens=[40, 31, 18, 6])
if which == "den":
    ctx = ref_cuda.RefContext("tests/golden/golden_tlm_a.fst", 0)
    logits = torch.tensor(y, device="cuda"); grad = torch.zeros_like(logits)
    ca = torch.zeros(4, device="cuda"); cb = torch.zeros(4, device="cuda")
    ref_cuda.gpu_den(ctx, logits, grad, torch.tensor(lens).cuda(), ca, cb)
    torch.cuda.synchronize()
    la, lb, gd = oracle.den(g, y, lens)
    print("ref den logz", ca.cpu().numpy(), "oracle", la, "beta", cb.cpu().numpy())
    print("ref den grad maxdiff", np.abs(grad.cpu().numpy() - gd).max())
else:
    act = torch.tensor(y, device="cuda").transpose(0, 1).contiguous(); grads = torch.zeros_like(act); costs = torch.zeros(4)
    st = ref_cuda.gpu_ctc(act, grads, torch.tensor(labels), torch.tensor(ly), torch.tensor(lens), 4, costs, 0)
    torch.cuda.synchronize()
    lp, gc = oracle.ctc(y, labels, ly, lens)
    print("ref ctc status", st, costs.numpy(), "oracle", lp, "grad maxdiff", np.abs(grads.transpose(0, 1).cpu().numpy() - gc).max())
