"""Bottleneck ablations of K1p (TCP_DBG build: NPLDA_LIB=.../libnplda_tdbg.so, NPLDA_TCP_DEBUG=mask): burst timing of the
bf16x3 (impl 5) and mixed (impl 6) pair kernels on 1 M pairs.  Results are wrong by construction for masks != 0."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = npl.NeuralPlda(bench.NC).to(dev)
n = 1_000_000
x1, x2, t = bench.synth_on_device(n, 1002, kp["mean"].to(dev), dev)
lib = _lib.lib()
pack = m.packed.get("nplda", m._params(), 512, 170, 170)
scores = torch.empty(n, device=dev)
def k1(impl):
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), impl, _lib.stream_ptr()), "k1")
out = []
for impl in (5, 6):
    time.sleep(0.5)
    for _ in range(2): k1(impl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): k1(impl)
    e1.record(); torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / 10)
print(f"dbg {os.environ.get('NPLDA_TCP_DEBUG', '0'):>3}: pair bf16x3 {out[0]:.4f} ms   pair mixed {out[1]:.4f} ms", flush=True)
