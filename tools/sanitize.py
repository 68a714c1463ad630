"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): the tcgen05 score kernel in
all its forms (plain, fp16x3 + guarded fallback, f8, EMIT, DPL, BWD), the CTA-pair kernel on materialised pairs (bf16x3 and mixed +
guarded fallback), SIMT, the CTA-pair kernel over a pre-split table,
trial lists (per-trial and sub-grid + gather), grids, losses, both backward paths, the DPlda gradient kernel, sort.
SANITIZE_SMALL=1 shrinks the batches (racecheck is ~100x slower than memcheck)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib, functional as F_
from oracle import nplda_oracle as O
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = npl.NeuralPlda(bench.NC).to(dev)
sd = m.state_dict()
for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                  ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
    sd[name].copy_(kp[key])
SMALL = os.environ.get("SANITIZE_SMALL") == "1"
x1, x2, t = O.synth_pairs((2000 if SMALL else 20000) + 37, 50, seed=3, mean=kp["mean"])
a, b, y = x1.to(dev), x2.to(dev), t.to(dev)
for impl in (npl.IMPL_TC, npl.IMPL_TC_F8, npl.IMPL_SIMT, npl.IMPL_TC_BF16, npl.IMPL_TC_PAIR, _lib.IMPL_TC_PAIR_F8):
    m.impl = impl
    with torch.no_grad():
        s = m(a, b); s2 = m(a * 1000, b * 1000)
    print("impl", impl, float(s.sum()), float(s2.sum()))
m.impl = npl.IMPL_AUTO
nb = 1000 if SMALL else 4096
loss = m.loss(m(a[:nb], b[:nb]), y[:nb]); loss.backward()          # default backward: BWD form of the score kernel + gemm
# tensor-core backward pieces (EMIT pass + tcgen05 weight gradients) and the DPlda tensor-core forward / backward
_lib.lib().nplda_debug_backward_paths(2, 2, 0)
n_tc = (1024 if SMALL else 8192) + 37
loss = m.loss(m(a[:n_tc], b[:n_tc]), y[:n_tc]); loss.backward()
class NCD(bench.NC):
    loss = "crossentropy"; beta = [99.0]
d = npl.DPlda(NCD).to(dev)
d.state_dict()["centering_and_LDA.weight"].copy_(kp["W1"]); d.state_dict()["centering_and_LDA.bias"].copy_(kp["b1"])
ld = d.loss(d(a[:n_tc], b[:n_tc]), y[:n_tc]); ld.backward()
print("tc backward", float(loss), float(ld), float(d.logistic_regres.weight.grad.abs().sum()))
# DPlda as the reference trains it: LDA frozen -> DPL kernel emitting the u rows + the gradient kernel of logistic_regres
for p_ in (d.centering_and_LDA.weight, d.centering_and_LDA.bias):
    p_.requires_grad_(False)
d.zero_grad(set_to_none=True)
ld = d.loss(d(a[:n_tc], b[:n_tc]), y[:n_tc]); ld.backward()
with torch.no_grad():
    sdp = d(a[:n_tc], b[:n_tc])
print("dplda frozen", float(ld), float(d.logistic_regres.weight.grad.abs().sum()), float(sdp.sum()))
_lib.lib().nplda_debug_backward_paths(0, 0, 0)
table, i1, i2, _ = O.synth_grid(40, 50, 7, seed=2, mean=kp["mean"])
s, f = m.forward_indexed(table.to(dev), i1.to(dev), i2.to(dev))
s, f = m.forward_indexed(table.to(dev), i1.to(dev), i2.to(dev), embed_once=False)
s, f = m.forward_indexed(table.to(dev), i1.to(dev), i2.to(dev), embed_once=False, use_split=True)      # CTA-pair kernel
F_.GRID_GATHER_MIN_TRIALS = 1
s, f = m.forward_indexed(table.to(dev), i1.to(dev), i2.to(dev), embed_once=True)                        # sub-grid + gather
from neuralplda_b200 import adaptive_score_normalization as asn
from neuralplda_b200.sv_trials_loaders import load_xvec_trials_from_numbatch
er, tr = torch.arange(40, device=dev), torch.arange(40, 90, device=dev)
S = m.forward_grid(table.to(dev), er, tr)[0]
S = m.forward_grid(table.to(dev), torch.arange(0, 90, device=dev), torch.arange(0, 90, device=dev).repeat(4))[0]   # 90 x 360, ragged tiles
st = asn.cohort_statistics(S, 100)
z = asn.normalize_scores(S[:, 0].contiguous(), torch.arange(90, device=dev), torch.arange(90, device=dev).flip(0), st)
mega = {"u%d" % i: table[i].numpy() for i in range(table.shape[0])}
n2i = dict(enumerate(mega))
for _ in range(2):
    g1, g2 = load_xvec_trials_from_numbatch(mega, n2i, i1[:300].to(dev), i2[:300].to(dev), dev)
print("grid", float(S.sum()), "norm", float(z.sum()), "gather", float(g1.sum()))
print("minc", m.minc(m(a[:nb], b[:nb]).detach(), y[:nb])[0].item())
torch.cuda.synchronize()
print("sanitize run ok")
