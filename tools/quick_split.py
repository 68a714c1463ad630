"""Quick harness for the pre-split CTA-pair kernel (nplda_score_fwd_split): parity against the CPU oracle on ragged
sizes, then timings (CUDA events) on 1 M trials over tables of 6 500 and 100 000 utterances.
Usage: python tools/quick_split.py [rounds]"""
import os, sys, time, numpy as np, torch
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
args = (kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
g = torch.Generator().manual_seed(5)
U = 3000
spk = torch.randn(300, 512, generator=g)
table_h = kp["mean"] + spk[torch.randint(0, 300, (U,), generator=g)] + 0.7 * torch.randn(U, 512, generator=g)
table = table_h.to(dev)
worst_all = 0.0
for n in (() if os.environ.get("QUICK") else (1, 63, 64, 65, 127, 128, 129, 1000, 9472, 100_003)):
    i1 = torch.randint(0, U, (n,), generator=g); i2 = torch.randint(0, U, (n,), generator=g)
    ref = O.nplda_score(table_h[i1], table_h[i2], *args).double()
    got, flag = m.forward_indexed(table, i1.to(dev), i2.to(dev), embed_once=False, use_split=True)
    torch.cuda.synchronize()
    bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
    w = float(((got.cpu().double() - ref).abs() / bound).max()); worst_all = max(worst_all, w)
    print(f"parity split n={n}: worst/bound {w:.3f} flag {int(flag)} {'OK' if w <= 1 and int(flag) == 0 else 'FAIL'}", flush=True)
bad = torch.tensor([0, U, 5, -1] * 40); ok = torch.zeros(160, dtype=torch.int64)
_, flag = m.forward_indexed(table, bad.to(dev), ok.to(dev), embed_once=False, use_split=True)
print("bad-index flag raised:", int(flag) != 0, flush=True)

def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

n = 1_000_000
lib = _lib.lib()
for U2 in (6500, 100_000):
    tab = (kp["mean"].to(dev) + torch.randn(U2, 512, device=dev))
    a = torch.randint(0, U2, (n,), device=dev); b = torch.randint(0, U2, (n,), device=dev)
    split = F_.split_table(tab)
    pack = m.packed.get("nplda", m._params(), 512, 170, 170, pair=True)
    scores = torch.empty(n, device=dev); flag = torch.zeros(1, dtype=torch.int32, device=dev)
    def k():
        _lib.check(lib.nplda_score_fwd_split(_lib.ptr(split), U2, _lib.ptr(a), _lib.ptr(b), n, 512, 170, 170, _lib.ptr(pack),
                                             _lib.ptr(scores), _lib.ptr(flag), _lib.stream_ptr()), "split")
    res = []
    for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        time.sleep(1.0)
        res.append(timeit(k, 10))
    time.sleep(2.0)
    sus = timeit(k, 300)
    t_split = timeit(lambda: lib.nplda_table_split(_lib.ptr(tab), U2, 512, _lib.ptr(split), _lib.stream_ptr()), 20)
    print(f"table {U2} rows: split kernel burst ms {' '.join(f'{v:.4f}' for v in res)}  sustained(300) {sus:.4f} ms "
          f"-> {n / min(res) / 1e6:.1f} / {n / sus / 1e6:.1f} G pairs/s x1e-3; table split {t_split * 1e3:.1f} us", flush=True)
    x1, x2 = tab[a[:200_000]], tab[b[:200_000]]
    ref = m(x1, x2)      # materialised K1 on the same rows
    got = scores[:200_000]
    print("  vs K1 (materialised) max abs diff:", float((ref - got).abs().max()), flush=True)
print("PARITY", "OK" if worst_all <= 1 else "FAIL")
