"""cProfile of the B=128 training-loop body (host-side overheads)."""
import os, sys, time, cProfile, pstats, torch
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
class C(bench.NC):
    loss = "crossentropy"
m = npl.NeuralPlda(C).to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-4)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
perm = torch.randperm(i1.numel())[:B]
d1, d2, tg = i1[perm], i2[perm], lab[perm]
T = {}
def tick(name, t0):
    torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
def step(sync=False):
    t0 = time.perf_counter(); opt.zero_grad()
    if sync: tick("zero_grad", t0); t0 = time.perf_counter()
    a, b, t = d1.to(dev), d2.to(dev), tg.to(dev)
    if sync: tick("h2d", t0); t0 = time.perf_counter()
    x1, x2 = load_xvec_trials_from_numbatch(mega, num_to_id, a, b, dev)
    if sync: tick("gather", t0); t0 = time.perf_counter()
    out = m(x1, x2)
    if sync: tick("forward", t0); t0 = time.perf_counter()
    loss = m.loss(out, t)
    if sync: tick("loss", t0); t0 = time.perf_counter()
    v = loss.item()
    if sync: tick("item", t0); t0 = time.perf_counter()
    loss.backward()
    if sync: tick("backward", t0); t0 = time.perf_counter()
    opt.step()
    if sync: tick("adam", t0)
    return v
for _ in range(10): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
torch.cuda.synchronize()
print(f"B={B}: {(time.perf_counter() - t0) / 200 * 1e3:.3f} ms/step async")
for _ in range(200): step(True)
print("per-phase ms (synchronised after each phase):", {k: round(v / 200 * 1e3, 3) for k, v in T.items()})
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(20): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=60))
