"""Where the frozen-LDA DPlda training step stalls: host-side durations of forward / loss / backward (perf_counter, no syncs
inside the step), device time per step (events), and the caching allocator's segment counters."""
import os, sys, time, numpy as np, torch
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
d = npl.DPlda(NCD).to(dev)
sd = d.state_dict()
sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
for p_ in (d.centering_and_LDA.weight, d.centering_and_LDA.bias): p_.requires_grad_(False)
def stats():
    s = torch.cuda.memory_stats(dev)
    return s["segment.all.allocated"], s["segment.all.freed"], s["num_alloc_retries"], s["reserved_bytes.all.current"] >> 20
rows = []
ev = [torch.cuda.Event(enable_timing=True) for _ in range(17)]
torch.cuda.synchronize()
ev[0].record()
for i in range(16):
    d.zero_grad(set_to_none=True)
    h0 = time.perf_counter(); s = d(x1, x2); h1 = time.perf_counter(); loss = d.loss(s, t); h2 = time.perf_counter(); loss.backward(); h3 = time.perf_counter()
    del s, loss
    ev[i + 1].record()
    rows.append((h1 - h0, h2 - h1, h3 - h2, stats()))
torch.cuda.synchronize()
for i, r in enumerate(rows):
    print(f"step {i:2d}: device {ev[i].elapsed_time(ev[i + 1]):7.2f} ms | host fwd {r[0] * 1e3:7.2f} loss {r[1] * 1e3:6.2f} bwd {r[2] * 1e3:6.2f} ms | segments alloc/freed/retries/reservedMB {r[3]}")
