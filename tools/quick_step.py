"""Latency of the reference's training loop body (xvector_NeuralPlda_pytorch.py:35-43) at its own batch sizes:
loader gather -> forward -> loss -> .item() -> backward -> Adam step, through the public API."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
from neuralplda_b200.sv_trials_loaders import load_xvec_trials_from_numbatch
from oracle import nplda_oracle as O
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
table, i1, i2, lab = O.synth_grid(200, 300, 10, seed=5, mean=kp["mean"])
mega = {"utt%05d" % i: table[i].numpy() for i in range(table.shape[0])}
num_to_id = {i: k for i, k in enumerate(mega)}
for lossname, fused in (("crossentropy", False), ("crossentropy", True), ("SoftCdet", False)):
    class C(bench.NC):
        loss = lossname
    m = npl.NeuralPlda(C).to(dev)
    sd = m.state_dict()
    for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                      ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
        sd[name].copy_(kp[key])
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=fused)
    for B in (128, 2048, 16384):
        perm = torch.randperm(i1.numel())[:B]
        d1, d2, tg = i1[perm], i2[perm], lab[perm]
        def step():
            opt.zero_grad()
            a, b, t = d1.to(dev), d2.to(dev), tg.to(dev)
            x1, x2 = load_xvec_trials_from_numbatch(mega, num_to_id, a, b, dev)
            out = m(x1, x2)
            loss = m.loss(out, t)
            v = loss.item()
            loss.backward()
            opt.step()
            return v
        for _ in range(5): step()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        t0 = time.perf_counter()
        n = 50
        for _ in range(n): v = step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"{lossname} fused_adam={fused} B={B}: {dt * 1e3:.3f} ms/step ({B / dt / 1e3:.1f} k pairs/s), {(_lib.launch_count() - l0) / n:.0f} libnplda launches/step, loss {v:.5f}")
# the loop body captured in a CUDA graph (neuralplda_b200.graphs)
from neuralplda_b200.graphs import GraphedTrainStep
class CX(bench.NC):
    loss = "crossentropy"
for B, fused in ((128, False), (2048, False), (128, True), (2048, True)):
    m = npl.NeuralPlda(CX).to(dev)
    sd = m.state_dict()
    for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                      ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
        sd[name].copy_(kp[key])
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, capturable=True, fused=True if fused else None)
    gstep = GraphedTrainStep(m, opt, mega, num_to_id, batch_size=B)
    perm = torch.randperm(i1.numel())[:B]
    d1, d2, tg = i1[perm], i2[perm], lab[perm]
    for _ in range(6): v = gstep(d1, d2, tg).item()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n): v = gstep(d1, d2, tg).item()              # .item() every step, like the reference's loop
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print(f"graphed step crossentropy fused_adam={fused} B={B}: {dt * 1e3:.3f} ms/step ({B / dt / 1e3:.1f} k pairs/s), loss {v:.5f}")
# the same loop on the CPU with the oracle port (reference arithmetic, torch autograd), B = 128
W = {k: torch.nn.Parameter(kp[k].clone()) for k in ("W1", "b1", "W2", "b2", "P_sqrt", "Q")}
thx = torch.nn.Parameter(torch.zeros(1))
opt = torch.optim.Adam(list(W.values()) + [thx], lr=1e-4)
for B in (128, 2048):
    perm = torch.randperm(i1.numel())[:B]
    d1, d2, tg = i1[perm], i2[perm], lab[perm]
    def cstep():
        opt.zero_grad()
        x1, x2 = O.gather_numbatch(mega, num_to_id, d1, d2)
        s = O.nplda_score(x1, x2, W["W1"], W["b1"], W["W2"], W["b2"], W["P_sqrt"], W["Q"])
        loss = O.crossentropy(s, tg, thx)
        v = loss.item(); loss.backward(); opt.step()
        return v
    for _ in range(3): cstep()
    t0 = time.perf_counter()
    for _ in range(20): cstep()
    dt = (time.perf_counter() - t0) / 20
    print(f"CPU oracle port, crossentropy B={B}: {dt * 1e3:.3f} ms/step ({B / dt / 1e3:.1f} k pairs/s), {torch.get_num_threads()} threads")
