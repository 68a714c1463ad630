"""Times K1 (tensor-core kernel) under the NPLDA_TC_DEBUG ablation masks and prints one tile timeline."""
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
n = int(os.environ.get("N", "1000000"))
x1, x2, t = bench.synth_on_device(n, 1002, kp["mean"].to(dev), dev)
lib = _lib.lib()
pack = m.packed.get("nplda", m._params(), 512, 170, 170)
scores = torch.empty(n, device=dev)
def k1():
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), npl.IMPL_TC, _lib.stream_ptr()), "k1")
def timeit(reps=20):
    for _ in range(3): k1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): k1()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 32, 3, 5, 6, 7, 33, 36, 37, 39, 63, 0]
for mask in masks:
    os.environ["NPLDA_TC_DEBUG"] = str(mask)
    print(f"dbg {mask:3d}: {timeit():.4f} ms", flush=True)
for mask in [int(a) for a in os.environ.get("PROF_MASKS", "0,4,63").split(",") if a]:
    os.environ["NPLDA_TC_DEBUG"] = str(mask)
    os.environ["NPLDA_TC_PROF"] = "1"
    print(f"--- cycle accounting, dbg {mask}", flush=True)
    k1(); torch.cuda.synchronize()
    del os.environ["NPLDA_TC_PROF"]
os.environ["NPLDA_TC_DEBUG"] = "0"
