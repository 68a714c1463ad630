"""Burst-regime time of K1 (bf16x3, 1M pairs) under ONE debug mask (NPLDA_TC_DEBUG, read once per process by a DEBUG=1
build: 1 no x loads, 2 no weight copies after the first ring fill, 4 no MMAs, 8 no TMEM stores, 16 no LDS, 32 idle epilogue).
Usage: for m in 0 1 2 4 ...; do NPLDA_LIB=.../libnplda_dbg.so NPLDA_TC_DEBUG=$m python tools/ablate_k1.py; done"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = bench.load_kaldi_init(npl.NeuralPlda(bench.NC).to(dev), kp)
n = 1_000_000
x1, x2, t = bench.synth_on_device(n, 1002, kp["mean"].to(dev), dev)
lib = _lib.lib()
pack = m.packed.get("nplda", m._params(), 512, 170, 170)
scores = torch.empty(n, device=dev)
def k1():
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), npl.IMPL_TC_BF16, _lib.stream_ptr()), "k1")
def timeit(reps):
    for _ in range(2): k1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): k1()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = []
for _ in range(3):
    time.sleep(1.0)
    res.append(timeit(10))
time.sleep(1.0)
sus = timeit(300)
print(f"dbg {os.environ.get('NPLDA_TC_DEBUG', '0'):>3s}: burst " + " ".join(f"{v:.4f}" for v in res) + f" ms; sustained(300) {sus:.4f} ms", flush=True)
