"""Times the grid kernel (nplda_score_grid) on BASELINE.json configs[2] / [3] (2500 x 4000 and 5000 x 10000)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from oracle import nplda_oracle as O
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = npl.NeuralPlda(bench.NC).to(dev)
sd = m.state_dict()
for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                  ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
    sd[name].copy_(kp[key])
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for ne, nt, spk, seed in ((2500, 4000, 500, 1003), (5000, 10000, 700, 1004)):
    table, i1, i2, _ = O.synth_grid(ne, nt, spk, seed=seed, mean=kp["mean"])
    t = table.to(dev)
    er, tr = torch.arange(ne, device=dev), torch.arange(ne, ne + nt, device=dev)
    n = ne * nt
    s, _ = m.forward_grid(t, er, tr)
    sub = torch.arange(0, n, 997)
    ref = O.nplda_score(table[i1[sub]], table[i2[sub]], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]).double()
    got = s.flatten()[sub.to(dev)].cpu().double()
    bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
    print(f"{ne} x {nt}: parity on {sub.numel()} strided trials: worst/bound {float(((got - ref).abs() / bound).max()):.3f}")
    from neuralplda_b200 import _lib
    rowtab = m.packed.get_rowtab("nplda", t, m._params(), 512, 170, 170)
    sc, flag = torch.empty(ne, nt, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    for impl, nm in ((_lib.IMPL_AUTO, "tcgen05"), (_lib.IMPL_SIMT, "fp32 FFMA2")):
        k = lambda: _lib.check(_lib.lib().nplda_score_grid_impl(_lib.ptr(rowtab), t.shape[0], _lib.ptr(er), ne, _lib.ptr(tr), nt, _lib.ptr(sc), nt,
                                                                _lib.ptr(flag), impl, _lib.stream_ptr()), "grid")
        k(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()                    # 20 launches per graph replay: no host time between the kernels
        with torch.cuda.graph(gr):
            for _ in range(20): k()
        msk = timeit(gr.replay, 5) / 20
        print(f"  kernel alone ({nm}): {msk * 1e3:.1f} us -> {n / msk / 1e6:.1f} G trials/s, {4 * n / msk / 1e6:.0f} GB/s of scores")
    ms = timeit(lambda: m.forward_grid(t, er, tr))
    print(f"  grid, rows cached: {ms:.3f} ms -> {n / ms / 1e6:.1f} G trials/s, {2 * 176 * n / ms / 1e9:.1f} TFLOP/s fp32, "
          f"{4 * n / ms / 1e6:.0f} GB/s of scores")
    def cold():
        m.packed.rowtab_key = None
        m.forward_grid(t, er, tr)
    ms = timeit(cold)
    print(f"  grid incl. table prepare ({ne + nt} rows): {ms:.3f} ms -> {n / ms / 1e6:.1f} G trials/s")
