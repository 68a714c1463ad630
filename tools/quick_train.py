"""Throughput of the training step (forward + loss + backward into .grad) and of the DPlda forward, materialised pairs."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
def timeit(fn, reps=12):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    print(f"    per-iteration ms: min {ts[0]:.2f} median {ts[reps // 2]:.2f} max {ts[-1]:.2f}")
    return ts[reps // 2]
class NCX(bench.NC):
    loss = "crossentropy"
class NCD(bench.NC):
    loss = "crossentropy"; beta = [99.0]
for n in ([int(sys.argv[1])] if len(sys.argv) > 1 else [131072, 1_000_000]):
    x1, x2, t = bench.synth_on_device(n, 1005, kp["mean"].to(dev), dev)
    m = npl.NeuralPlda(NCX).to(dev)
    sd = m.state_dict()
    for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                      ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
        sd[name].copy_(kp[key])
    def step():
        m.zero_grad(set_to_none=True)
        loss = m.loss(m(x1, x2), t)
        loss.backward()
    ms = timeit(step)
    print(f"NeuralPlda train step (fwd + BCE + bwd), {n} pairs: {ms:.2f} ms -> {n / ms / 1e3:.1f} M pairs/s")
    d = npl.DPlda(NCD).to(dev)
    sd = d.state_dict()
    sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
    for p in (d.centering_and_LDA.weight, d.centering_and_LDA.bias):
        p.requires_grad_(False)                               # xvector_DPlda_pytorch.py:140-147 freezes the LDA
    with torch.no_grad():
        ms = timeit(lambda: d(x1, x2))
    print(f"DPlda forward, {n} pairs: {ms:.2f} ms -> {n / ms / 1e3:.1f} M pairs/s")
    def dstep():
        d.zero_grad(set_to_none=True)
        loss = d.loss(d(x1, x2), t)
        loss.backward()
    ms = timeit(dstep)
    print(f"DPlda train step (fwd + BCE + bwd, LDA frozen), {n} pairs: {ms:.2f} ms -> {n / ms / 1e3:.1f} M pairs/s")
