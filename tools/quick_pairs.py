"""Times the embed-once trial-list path on BASELINE.json configs[2] (10 M trials = 2500 x 4000 grid, 6500 vectors)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
from oracle import nplda_oracle as O
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = npl.NeuralPlda(bench.NC).to(dev)
sd = m.state_dict()
for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                  ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
    sd[name].copy_(kp[key])
table, i1, i2, lab = O.synth_grid(2500, 4000, 500, seed=1003, mean=kp["mean"])
t, a, b = table.to(dev), i1.to(dev), i2.to(dev)
n = a.numel()
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
s, _ = m.forward_indexed(t, a, b)
sub = torch.arange(0, n, 97)
ref = O.nplda_score(table[i1[sub]], table[i2[sub]], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]).double()
got = s[sub.to(dev)].cpu().double()
bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
print(f"parity on {sub.numel()} strided trials: worst/bound {float(((got - ref).abs() / bound).max()):.3f}")
ms = timeit(lambda: m.forward_indexed(t, a, b))
print(f"embed-once, rows cached: {ms:.3f} ms for {n} trials -> {n / ms / 1e6:.2f} G trials/s")
def cold():
    m.packed.rowtab_key = None
    m.forward_indexed(t, a, b)
ms = timeit(cold)
print(f"embed-once incl. table prepare (6500 rows): {ms:.3f} ms -> {n / ms / 1e6:.2f} G trials/s")
ms = timeit(lambda: m.forward_indexed(t, a[:1_000_000], b[:1_000_000], embed_once=False), 3)
print(f"per-trial fused gather kernel: {ms:.3f} ms per 1 M trials -> {1e6 / ms / 1e6:.3f} G trials/s")
ha, hb = i1.pin_memory(), i2.pin_memory()
hs = torch.empty(n, pin_memory=True)
def e2e():
    da, db = ha.to(dev, non_blocking=True), hb.to(dev, non_blocking=True)
    s, _ = m.forward_indexed(t, da, db)
    hs.copy_(s, non_blocking=True)
    torch.cuda.synchronize()
e2e(); t0 = time.perf_counter()
for _ in range(5): e2e()
dt = (time.perf_counter() - t0) / 5
print(f"host indices -> host scores: {dt * 1e3:.2f} ms -> {n / dt / 1e6:.1f} M trials/s (H2D 16 B, D2H 4 B per trial)")
