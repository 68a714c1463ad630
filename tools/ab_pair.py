"""Same-process A/B of two libnplda builds on the pair kernel is not possible (one library per process); this runs the burst /
sustained timing of impl 5 (pair bf16x3) only, for interleaved process-level A/B from the shell: python tools/ab_pair.py [reps]"""
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
impl = int(os.environ.get("IMPL", "5"))
def k1():
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), impl, _lib.stream_ptr()), "k1")
res = []
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    time.sleep(1.0)
    for _ in range(2): k1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): k1()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 10)
time.sleep(1.0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(600): k1()
e1.record(); torch.cuda.synchronize()
print(f"{os.path.basename(os.environ.get('NPLDA_LIB', 'libnplda.so')):>18} impl {impl}: burst " + " ".join(f"{v:.4f}" for v in res) + f"  sustained {e0.elapsed_time(e1) / 600:.4f} ms  checksum {float(scores.double().sum()):.6f}", flush=True)
