"""Burst-regime ablation of K1: every measurement is preceded by an idle pause so that the power-cap
state of the previous one does not leak into it; each mask is measured twice in interleaved order."""
import os, sys, time, numpy as np, torch
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
scores = torch.empty(n, device=dev)
def k1():
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), npl.IMPL_TC, _lib.stream_ptr()), "k1")
def timeit(reps):
    for _ in range(2): k1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): k1()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 32, 36, 5, 7, 39, 63]
pause = float(os.environ.get("PAUSE", "1.5")); reps = int(os.environ.get("REPS", "10"))
out = {k: [] for k in masks}
for rnd in range(3):
    for mask in masks:
        os.environ["NPLDA_TC_DEBUG"] = str(mask)
        time.sleep(pause)
        out[mask].append(timeit(reps))
for mask in masks:
    print(f"dbg {mask:3d}: " + " ".join(f"{v:.4f}" for v in out[mask]) + " ms", flush=True)
os.environ["NPLDA_TC_DEBUG"] = "0"
time.sleep(pause)
print("sustained (dbg 0), 30 x 20 launches:", " ".join(f"{timeit(20):.3f}" for _ in range(30)))
