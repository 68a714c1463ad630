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
x1, x2, t = O.synth_pairs(4096, 64, seed=7, mean=kp["mean"])
ref = O.nplda_score(x1, x2, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]).double()
bound = 1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt())
def chk(tag, s):
    r = (s.detach().cpu().double() - ref).abs() / bound
    print(tag, "worst", float(r.max()), "nbad", int((r > 1).sum()), "first bad", torch.nonzero(r > 1).flatten()[:8].tolist())
m.impl = npl.IMPL_SIMT
out = m(x1.to(dev), x2.to(dev))
chk("after fwd", out)
keep = out.detach().clone()
loss = m.loss(out, t.to(dev))
chk("after loss", out)
loss.backward()
torch.cuda.synchronize()
chk("after bwd", out)
print("out changed:", int((keep != out.detach()).sum()))
m.zero_grad()
chk("after zero_grad", out)
