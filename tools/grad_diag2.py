"""Locates the gradient error stage: (a) dL/ds from the loss kernel, (b) parameter / input gradients of the score
backward for an EXACT upstream dL/ds, both against fp64 ground truth."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from oracle import nplda_oracle as O
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
names = ["W1", "b1", "W2", "b2", "P_sqrt", "Q"]
def err(g, g64):
    g, g64 = g.double().reshape(-1).cpu(), g64.double().reshape(-1)
    return float(((g - g64).abs() / torch.maximum(g64.abs(), g64.pow(2).mean().sqrt())).max())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x1, x2, t = O.synth_pairs(n, 200, seed=1001, mean=kp["mean"])
for lossname in ("SoftCdet", "crossentropy"):
    th, thx = [0.31, 0.47], 0.25
    ps = [kp[k].double().clone().requires_grad_(True) for k in names]
    X1, X2 = x1.double().requires_grad_(True), x2.double().requires_grad_(True)
    s64 = O.nplda_score(X1, X2, *ps)
    s64.retain_grad()
    loss = O.softcdet(s64, t.double(), [torch.tensor(v, dtype=torch.float64) for v in th], bench.BETAS, 15.0) if lossname == "SoftCdet" \
        else O.crossentropy(s64, t.double(), torch.tensor(thx, dtype=torch.float64))
    loss.backward()
    ds64 = s64.grad.clone()
    class C(bench.NC):
        loss = lossname
    m = bench.load_kaldi_init(npl.NeuralPlda(C).to(dev), kp)
    with torch.no_grad():
        m.Th99.fill_(th[0]); m.Th199.fill_(th[1]); m.threshold_Xent.fill_(thx)
    for impl, iname in ((npl.IMPL_SIMT, "simt"), (npl.IMPL_AUTO, "auto")):
        m.impl = impl
        with torch.no_grad():
            s = m(x1.to(dev), x2.to(dev))
        print(f"{lossname} {iname}: score err {err(s, s64.detach()):.2e}")
        sl = s.detach().clone().requires_grad_(True)
        m.loss(sl, t.to(dev)).backward()
        print(f"   dL/ds from the loss kernel at OUR scores      err {err(sl.grad, ds64):.2e}")
        sl2 = s64.detach().float().to(dev).requires_grad_(True)
        m.loss(sl2, t.to(dev)).backward()
        print(f"   dL/ds from the loss kernel at the fp64 scores err {err(sl2.grad, ds64):.2e}")
        m.zero_grad(set_to_none=True)
        a, b = x1.to(dev).requires_grad_(True), x2.to(dev).requires_grad_(True)
        m(a, b).backward(ds64.float().to(dev))
        for k, p, c in zip(names, m._params(), ps):
            print(f"   score backward with exact dL/ds: {k:7s} err {err(p.grad, c.grad):.2e}")
        print(f"   score backward with exact dL/ds: dx1     err {err(a.grad, X1.grad):.2e}   dx2 err {err(b.grad, X2.grad):.2e}")
        # fp32 torch autograd of the same graph on the GPU (library kernels) for scale
    p32 = [kp[k].to(dev).clone().requires_grad_(True) for k in names]
    torch.backends.cuda.matmul.allow_tf32 = False
    s32 = O.nplda_score(x1.to(dev), x2.to(dev), *p32)
    s32.backward(ds64.float().to(dev))
    for k, p, c in zip(names, p32, ps):
        print(f"   torch fp32 CUDA autograd with exact dL/ds: {k:7s} err {err(p.grad, c.grad):.2e}")
