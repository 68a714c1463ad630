"""Quick A/B harness for K1 experiments: parity of the tensor-core kernel against the fp32 SIMT kernel on
ragged sizes, then several timing rounds (CUDA events, 1 M pairs).  Usage: python tools/quick.py [rounds]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = npl.NeuralPlda(bench.NC).to(dev)
sd = m.state_dict()
for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                  ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
    sd[name].copy_(kp[key])
n = 1_000_000
x1, x2, t = bench.synth_on_device(n, 1002, kp["mean"].to(dev), dev)
lib = _lib.lib()
pack = m.packed.get("nplda", m._params(), 512, 170, 170, mixed=("f8" in os.environ.get("IMPLS", "tc,f8")))
def run(impl, cnt):
    out = torch.empty(cnt, device=dev)
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), cnt, 512, 170, 170, _lib.ptr(pack), _lib.ptr(out), impl, _lib.stream_ptr()), "k1")
    return out
worst_all = 0.0
IMPLS = {"tc": npl.IMPL_TC, "f8": npl.IMPL_TC_F8, "bf16": npl.IMPL_TC_BF16}
which = os.environ.get("IMPLS", "tc,f8").split(",")
for cnt in (1, 63, 64, 65, 9471, 9472, 100_003, 1_000_000):
    ref = run(npl.IMPL_SIMT, cnt).double()
    bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
    for nm in which:
        got = run(IMPLS[nm], cnt).double()
        w = float(((got - ref).abs() / bound).max()); worst_all = max(worst_all, w)
        print(f"parity {nm} n={cnt}: worst/bound {w:.3f} {'OK' if w <= 1 else 'FAIL'}", flush=True)
if "f8" in which:       # range guard: scaled inputs must be recomputed by the bf16x3 pass (same answer as IMPL_TC)
    keep = (x1, x2)
    for sc in (1e-3, 0.05, 300.0):
        x1, x2 = keep[0][:200_000] * sc, keep[1][:200_000] * sc
        ref = run(npl.IMPL_SIMT, 200_000).double(); got = run(npl.IMPL_TC_F8, 200_000).double(); tc = run(npl.IMPL_TC_BF16, 200_000).double()
        bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
        w = float(((got - ref).abs() / bound).max()); worst_all = max(worst_all, w)
        print(f"guard x*{sc}: worst/bound {w:.3f}, identical to IMPL_TC: {bool((got == tc).all())}", flush=True)
    x1, x2 = keep
scores = torch.empty(n, device=dev)
import time
def k1(impl=npl.IMPL_TC):
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), impl, _lib.stream_ptr()), "k1")
res = {nm: [] for nm in which}
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    for nm in which:                      # interleaved, with an idle pause: burst-regime numbers
        time.sleep(1.0)
        for _ in range(2): k1(IMPLS[nm])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): k1(IMPLS[nm])
        e1.record(); torch.cuda.synchronize()
        res[nm].append(e0.elapsed_time(e1) / 10)
for nm in which:
    print(f"k1 {nm} burst ms:", " ".join(f"{v:.4f}" for v in res[nm]), flush=True)
for nm in which:                          # sustained: 300 launches back to back
    time.sleep(2.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300): k1(IMPLS[nm])
    e1.record(); torch.cuda.synchronize()
    print(f"k1 {nm} sustained (300 launches) ms: {e0.elapsed_time(e1) / 300:.4f}", flush=True)
print("PARITY", "OK" if worst_all <= 1 else "FAIL")
