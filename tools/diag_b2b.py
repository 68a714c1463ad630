import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import _lib
class NC:
    xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = 512, 170, 170
    alpha, device, beta, loss = 15.0, "cpu", [99.0, 199.0], "SoftCdet"
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = npl.NeuralPlda(NC).to(dev)
n = int(sys.argv[1]); reps = int(sys.argv[2]); mode = sys.argv[3]
x1 = torch.randn(n, 512, device=dev); x2 = torch.randn(n, 512, device=dev)
pack = m.packed.get("nplda", m._params(), 512, 170, 170)
scores = torch.empty(n, device=dev); other = torch.zeros(1000, device=dev)
lib = _lib.lib()
torch.cuda.synchronize()
try:
    for r in range(reps):
        _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, 512, 170, 170, _lib.ptr(pack), _lib.ptr(scores), 2, _lib.stream_ptr()), "fwd")
        if mode == "sync": torch.cuda.synchronize()
        if mode == "small": other.add_(1.0)
    torch.cuda.synchronize()
    print(n, reps, mode, "OK", float(scores.sum()))
except Exception as e:
    print(n, reps, mode, "FAILED", str(e)[:80])
