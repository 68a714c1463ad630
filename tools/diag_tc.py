import sys, os, numpy as np, torch, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from oracle import nplda_oracle as O
class NC:
    xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = 512, 170, 170
    alpha, device, beta, loss = 15.0, "cpu", [99.0, 199.0], "SoftCdet"
z = np.load("tests/golden/kaldi_init_params.npz"); kp = {k: torch.from_numpy(z[k].copy()) for k in z.files}
dev = torch.device("cuda:0")
m = npl.NeuralPlda(NC).to(dev)
sd = m.state_dict()
for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"), ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
    sd[name].copy_(kp[key])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
x1, x2, t = O.synth_pairs(n, 200, seed=5, mean=kp["mean"])
x1d, x2d = x1.to(dev), x2.to(dev)
with torch.no_grad():
    m.impl = npl.IMPL_SIMT; ref = m(x1d, x2d)
    m.impl = npl.IMPL_TC
    for rep in range(int(os.environ.get("REPS", "3"))):
        s = m(x1d, x2d)
        bad = torch.nonzero((s - ref).abs() > 1e-3 * ref.abs().clamp_min(0.5)).flatten().cpu().numpy()
        tiles = bad // 64; pl = bad % 64
        print(f"rep {rep}: n_bad {bad.size}; tiles with errors {len(set(tiles))} of {(n+63)//64}; by tile-round (tile//148): {dict(collections.Counter((tiles//148).tolist()))}; by pl//8: {dict(sorted(collections.Counter((pl//8).tolist()).items()))}")
        if bad.size: print("   first bad:", bad[:12].tolist(), "got", [round(float(s[i]),3) for i in bad[:6]], "ref", [round(float(ref[i]),3) for i in bad[:6]])
