"""Training trajectory: 30 Adam steps on the GPU module vs the oracle port on the CPU, same batches."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from oracle import nplda_oracle as O
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
for lossname in ("SoftCdet", "crossentropy"):
    class C(bench.NC):
        loss = lossname
    m = npl.NeuralPlda(C).to(dev)
    sd = m.state_dict()
    for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                      ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
        sd[name].copy_(kp[key])
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    W = {k: torch.nn.Parameter(kp[k].clone()) for k in ("W1", "b1", "W2", "b2", "P_sqrt", "Q")}
    th = [torch.nn.Parameter(torch.zeros(1)) for _ in C.beta]
    thx = torch.nn.Parameter(torch.zeros(1))
    # same parameter order as the module (P_sqrt, Q, Th*, threshold_Xent, W1, b1, W2, b2)
    copt = torch.optim.Adam([W["P_sqrt"], W["Q"]] + th + [thx, W["W1"], W["b1"], W["W2"], W["b2"]], lr=1e-4)
    for step in range(30):
        x1, x2, t = O.synth_pairs(512, 40, seed=100 + step, mean=kp["mean"])
        opt.zero_grad()
        out = m(x1.to(dev), x2.to(dev))
        loss = m.loss(out, t.to(dev))
        loss.backward(); opt.step()
        copt.zero_grad()
        s = O.nplda_score(x1, x2, W["W1"], W["b1"], W["W2"], W["b2"], W["P_sqrt"], W["Q"])
        closs = O.softcdet(s, t, torch.cat(th), C.beta, C.alpha) if lossname == "SoftCdet" else O.crossentropy(s, t, thx)
        closs.backward(); copt.step()
        gn = float(torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None]).norm())
        cgn = float(torch.cat([p.grad.flatten() for g in copt.param_groups for p in g["params"] if p.grad is not None]).norm())
        if step < 5 or step % 5 == 4:
            print(f"{lossname} step {step}: gpu loss {loss.item():.6f} cpu {closs.item():.6f} | grad norm gpu {gn:.5g} cpu {cgn:.5g} | targets {int(t.sum())}")
