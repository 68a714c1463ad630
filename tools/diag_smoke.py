import sys, os, numpy as np, torch
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
m.impl = int(os.environ.get("NPLDA_IMPL", "1"))
for (n, spk, seed) in ((4096, 64, 7), (4096, 200, 7), (4096, 64, 8), (10000, 200, 1001), (4000, 64, 7)):
    x1, x2, t = O.synth_pairs(n, spk, seed=seed, mean=kp["mean"])
    ref = O.nplda_score(x1, x2, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]).double()
    for mode in ("nograd", "grad"):
        if mode == "nograd":
            with torch.no_grad():
                s = m(x1.to(dev), x2.to(dev))
        else:
            s = m(x1.to(dev), x2.to(dev)).detach()
        s2 = m(x1.to(dev), x2.to(dev)).detach()
        err = (s.cpu().double() - ref).abs()
        bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
        r = err / bound
        w = torch.argsort(r, descending=True)[:6]
        print(n, spk, seed, mode, "worst", float(r.max()), "n_bad", int((r > 1).sum()), "det", bool((s == s2).all()),
              "idx", w.tolist(), "ref", [round(float(ref[i]), 4) for i in w], "got", [round(float(s[i]), 4) for i in w])
