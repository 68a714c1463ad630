"""Kernel-level time breakdown (torch.profiler / CUPTI) of the 1M-pair training steps and the DPlda forward."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
class NCX(bench.NC):
    loss = "crossentropy"
class NCD(bench.NC):
    loss = "crossentropy"; beta = [99.0]
x1, x2, t = bench.synth_on_device(n, 1005, kp["mean"].to(dev), dev)
m = bench.load_kaldi_init(npl.NeuralPlda(NCX).to(dev), kp)
d = npl.DPlda(NCD).to(dev)
sd = d.state_dict()
sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
for p in (d.centering_and_LDA.weight, d.centering_and_LDA.bias):
    p.requires_grad_(False)
def nstep():
    m.zero_grad(set_to_none=True); m.loss(m(x1, x2), t).backward()
def dstep():
    d.zero_grad(set_to_none=True); d.loss(d(x1, x2), t).backward()
def dfwd():
    with torch.no_grad(): d(x1, x2)
def nfwd():
    with torch.no_grad(): m(x1, x2)
for name, fn in (("NeuralPlda train step", nstep), ("DPlda train step (LDA frozen)", dstep), ("DPlda forward", dfwd), ("NeuralPlda forward", nfwd)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"=== {name}, {n} pairs: {e0.elapsed_time(e1) / 5:.3f} ms per call", flush=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): fn()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
    tot = sum(r.device_time_total for r in rows)
    for r in rows[:14]:
        print(f"  {r.device_time_total / 3e3:8.3f} ms  {100 * r.device_time_total / tot:5.1f}%  x{r.count // 3:<3d} {r.key[:110]}")
