"""Per-role cycle buckets of K1p (TCP_PROF build, NPLDA_LIB=.../libnplda_prof.so): 1 M pairs, a few launches, the buckets
of the last one for both CTAs of cluster 0."""
import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
m = npl.NeuralPlda(bench.NC).to(dev)
x1, x2, t = bench.synth_on_device(n, 1002, kp["mean"].to(dev), dev)
lib = _lib.lib()
trace = torch.zeros(2 * 32 * 8, dtype=torch.int64).pin_memory()
lib.nplda_debug_set_tcp_trace.argtypes = [ctypes.c_void_p]
lib.nplda_debug_set_tcp_trace(trace.data_ptr())
pack = m.packed.get("nplda", m._params(), 512, 170, 170)
out = torch.empty(n, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for r in range(4):
    e0.record()
    _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(out), 5, _lib.stream_ptr()), "k1p")
    e1.record(); torch.cuda.synchronize()
print(f"last launch {e0.elapsed_time(e1):.4f} ms")
names = {"epi": ["wait d_full", "wait u_empty", "pass 1", "wait y_full", "pass 2", "tail+store"],
         "cv": ["wait x_full", "lds+convert", "wait a_empty", "st+arrive x", "st_wait+arrive a", "loop"],
         "mma": ["wait d_empty", "wait a_full", "wait b_full", "L1 issue+commit", "wait u_full", "L2 issue+commit", "other", "drain"],
         "bld": ["issue", "wait b_empty"], "xld": ["issue", "wait x_empty"]}
tr = trace.view(2, 32, 8).tolist()
for cta in range(2):
    for w, role in ((0, "epi"), (4, "epi"), (8, "cv"), (12, "cv"), (16, "cv"), (23, "cv"), (24, "mma"), (25, "bld"), (26, "xld")):
        b = tr[cta][w]; tot = sum(b)
        if tot == 0: continue
        print(f"[prof] CTA{cta} {role}{w:<2d} total {tot:9d} cyc: " + "  ".join(f"{nm} {100.0 * v / tot:.1f}%" for nm, v in zip(names[role], b)))
