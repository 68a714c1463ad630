import os, sys, torch
sys.path.insert(0, os.getcwd())
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
for _ in range(3):
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), npl.IMPL_TC_BF16, _lib.stream_ptr()), "k1")
torch.cuda.synchronize()
