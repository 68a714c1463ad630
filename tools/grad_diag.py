"""Gradient accuracy against an fp64 ground truth: per parameter,
err(g) = max_i |g_i - g64_i| / max(|g64_i|, rms(g64)), for (a) the reference arithmetic in fp32 (oracle, CPU autograd)
and (b) this package on the GPU, at 2048 and 100k pairs, both losses, NeuralPlda and DPlda."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from oracle import nplda_oracle as O
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
names = ["W1", "b1", "W2", "b2", "P_sqrt", "Q"]

def err(g, g64):
    g, g64 = g.double().reshape(-1), g64.double().reshape(-1)
    return float(((g - g64).abs() / torch.maximum(g64.abs(), g64.pow(2).mean().sqrt())).max())

def oracle_grads(x1, x2, t, lossname, dtype, th, thx):
    ps = [kp[k].to(dtype).clone().requires_grad_(True) for k in names]
    thr = [torch.tensor(v, dtype=dtype, requires_grad=True) for v in th]
    tx = torch.tensor(thx, dtype=dtype, requires_grad=True)
    s = O.nplda_score(x1.to(dtype), x2.to(dtype), *ps)
    loss = O.softcdet(s, t.to(dtype), thr, bench.BETAS, 15.0) if lossname == "SoftCdet" else O.crossentropy(s, t.to(dtype), tx)
    loss.backward()
    return [p.grad for p in ps], loss.item()

for n in (2048, 100_000):
    x1, x2, t = O.synth_pairs(n, 200, seed=1001, mean=kp["mean"])
    for lossname in ("SoftCdet", "crossentropy"):
        th, thx = [0.31, 0.47], 0.25
        g64, l64 = oracle_grads(x1, x2, t, lossname, torch.float64, th, thx)
        g32, l32 = oracle_grads(x1, x2, t, lossname, torch.float32, th, thx)
        class C(bench.NC):
            loss = lossname
        m = npl.NeuralPlda(C).to(dev)
        sd = m.state_dict()
        for nm, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                        ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
            sd[nm].copy_(kp[key])
        with torch.no_grad():
            m.Th99.fill_(th[0]); m.Th199.fill_(th[1]); m.threshold_Xent.fill_(thx)
        for impl, iname in ((npl.IMPL_SIMT, "simt"), (npl.IMPL_AUTO, "auto")):
            m.impl = impl
            m.zero_grad(set_to_none=True)
            loss = m.loss(m(x1.to(dev), x2.to(dev)), t.to(dev))
            loss.backward()
            ours = [p.grad.cpu() for p in m._params()]
            print(f"n={n} {lossname} {iname}: loss rel err ours {abs(loss.item() - l64) / abs(l64):.2e} ref32 {abs(l32 - l64) / abs(l64):.2e}")
            for k, a, b, c in zip(names, ours, g32, g64):
                print(f"    {k:7s} err ours {err(a, c):.2e}   reference-fp32 {err(b, c):.2e}   |g|rms {float(c.pow(2).mean().sqrt()):.2e}")
