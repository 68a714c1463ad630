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
pack = m.packed.get("nplda", m._params(), 512, 170, 170)
def run(impl, cnt):
    out = torch.empty(cnt, device=dev)
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), cnt, 512, 170, 170, _lib.ptr(pack), _lib.ptr(out), impl, _lib.stream_ptr()), "k1")
    return out
worst_all = 0.0
for cnt in (1, 63, 64, 65, 9471, 9472, 100_003, 1_000_000):
    ref = run(npl.IMPL_SIMT, cnt).double(); got = run(npl.IMPL_TC, cnt).double()
    bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
    w = float(((got - ref).abs() / bound).max()); worst_all = max(worst_all, w)
    print(f"parity n={cnt}: worst/bound {w:.3f} {'OK' if w <= 1 else 'FAIL'}", flush=True)
scores = torch.empty(n, device=dev)
def k1():
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), npl.IMPL_TC, _lib.stream_ptr()), "k1")
res = []
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 5):
    for _ in range(3): k1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): k1()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 20)
print("k1 ms per round:", " ".join(f"{v:.4f}" for v in res), f" min {min(res):.4f} median {sorted(res)[len(res)//2]:.4f}", flush=True)
print("PARITY", "OK" if worst_all <= 1 else "FAIL")
