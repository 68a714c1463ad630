"""A/B harness for K1p (CTA-pair bf16x3 kernel, csrc/score_tcp.cu) against K1 MODE 0 (one-CTA bf16x3): bit identity and
parity against the fp32 SIMT kernel on ragged sizes, then burst / sustained timing on 1 M pairs.
Usage: python tools/quick_pair.py [rounds]   (NPLDA_LIB selects an experiment build)"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = npl.NeuralPlda(bench.NC).to(dev)
sd = m.state_dict()
for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                  ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
    sd[name].copy_(kp[key])
n = 1_000_000
x1, x2, t = bench.synth_on_device(n, 1002, kp["mean"].to(dev), dev)
lib = _lib.lib()
pack = m.packed.get("nplda", m._params(), 512, 170, 170, mixed=True)
def run(impl, cnt, a=None, b=None):
    a = x1 if a is None else a; b = x2 if b is None else b
    out = torch.full((cnt,), float("nan"), device=dev)
    _lib.check(lib.nplda_score_fwd(_lib.ptr(a), _lib.ptr(b), cnt, 512, 170, 170, _lib.ptr(pack), _lib.ptr(out), impl, _lib.stream_ptr()), "k1")
    torch.cuda.synchronize()
    return out
PAIR, BF16, PAIRF8 = 5, npl.IMPL_TC_BF16, 6
worst_all, ident_all = 0.0, True
for cnt in (1, 63, 64, 65, 127, 128, 129, 9471, 9472, 9473, 100_003, 1_000_000):
    ref = run(npl.IMPL_SIMT, cnt).double()
    bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
    got = run(PAIR, cnt); one = run(BF16, cnt)
    w = float(((got.double() - ref).abs() / bound).max()); worst_all = max(worst_all, w)
    ident = bool((got == one).all()); ident_all &= ident
    g8 = run(PAIRF8, cnt).double()
    w8 = float(((g8 - ref).abs() / bound).max()); worst_all = max(worst_all, w8)
    print(f"parity pair n={cnt}: worst/bound {w:.3f} {'OK' if w <= 1 else 'FAIL'}  identical to one-CTA bf16x3: {ident}"
          f"  max|diff| {float((got - one).abs().max()):.3e}   pair-f8 worst/bound {w8:.3f} {'OK' if w8 <= 1 else 'FAIL'}", flush=True)
for sc in (1e-3, 300.0):          # any fp32 range
    a, b = x1[:200_000] * sc, x2[:200_000] * sc
    ref = run(npl.IMPL_SIMT, 200_000, a, b).double(); got = run(PAIR, 200_000, a, b).double()
    bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
    w = float(((got - ref).abs() / bound).max()); worst_all = max(worst_all, w)
    g8 = run(PAIRF8, 200_000, a, b)
    w8 = float(((g8.double() - ref).abs() / bound).max()); worst_all = max(worst_all, w8)
    print(f"range x*{sc}: worst/bound {w:.3f}; pair-f8 (guard -> bf16x3 pass) worst/bound {w8:.3f}, identical to pair bf16x3: {bool((g8 == got.float()).all())}", flush=True)
chk = run(PAIRF8, 100_003)        # an in-range call after the guarded ones takes the mixed path again
print("pair-f8 after guard: differs from bf16x3 (mixed path taken):", bool((chk != run(PAIR, 100_003)).any()), flush=True)
scores = torch.empty(n, device=dev)
def k1(impl):
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), impl, _lib.stream_ptr()), "k1")
which = {"pair": PAIR, "pairf8": PAIRF8, "one": BF16}
res = {nm: [] for nm in which}
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    for nm, impl in which.items():        # interleaved, with an idle pause: burst-regime numbers
        time.sleep(1.0)
        for _ in range(2): k1(impl)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): k1(impl)
        e1.record(); torch.cuda.synchronize()
        res[nm].append(e0.elapsed_time(e1) / 10)
for nm in which:
    print(f"k1 {nm} burst ms:", " ".join(f"{v:.4f}" for v in res[nm]), flush=True)
for nm, impl in which.items():            # sustained: 600 launches back to back
    time.sleep(2.0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(600): k1(impl)
    e1.record(); torch.cuda.synchronize()
    print(f"k1 {nm} sustained (600 launches) ms: {e0.elapsed_time(e1) / 600:.4f}", flush=True)
print("PARITY", "OK" if worst_all <= 1 else "FAIL", "IDENTICAL" if ident_all else "NOT-IDENTICAL")
