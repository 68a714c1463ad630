"""Per-iteration forward / backward times of the frozen-LDA DPlda training step (fresh random logistic_regres per process)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
n = 1_000_000
x1, x2, t = bench.synth_on_device(n, 1005, kp["mean"].to(dev), dev)
class NCD(bench.NC):
    loss = "crossentropy"; beta = [99.0]
if len(sys.argv) > 1: torch.manual_seed(int(sys.argv[1]))
d = npl.DPlda(NCD).to(dev)
sd = d.state_dict()
sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
for p_ in (d.centering_and_LDA.weight, d.centering_and_LDA.bias): p_.requires_grad_(False)
fw, bw = [], []
l0 = _lib.launch_count()
for i in range(10):
    d.zero_grad(set_to_none=True)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record(); s = d(x1, x2); loss = d.loss(s, t); e[1].record(); loss.backward(); e[2].record()
    torch.cuda.synchronize()
    fw.append(e[0].elapsed_time(e[1])); bw.append(e[1].elapsed_time(e[2]))
print("launches per step", (_lib.launch_count() - l0) / 10, " |w| max", float(d.logistic_regres.weight.abs().max()))
print("fwd+loss ms:", " ".join(f"{v:.2f}" for v in fw))
print("bwd ms     :", " ".join(f"{v:.2f}" for v in bw))
