"""One NeuralPlda and one DPlda training step at 1M pairs, no warm-up: the workload of the ncu captures of the backward kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
import bench
dev = torch.device("cuda:0")
kp = bench.kaldi_params()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
class NCX(bench.NC):
    loss = "crossentropy"
class NCD(bench.NC):
    loss = "crossentropy"; beta = [99.0]
x1, x2, t = bench.synth_on_device(n, 1005, kp["mean"].to(dev), dev)
m = bench.load_kaldi_init(npl.NeuralPlda(NCX).to(dev), kp)
m.loss(m(x1, x2), t).backward()
torch.cuda.synchronize()
if len(sys.argv) > 2:
    d = npl.DPlda(NCD).to(dev)
    sd = d.state_dict()
    sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
    for p in (d.centering_and_LDA.weight, d.centering_and_LDA.bias):
        p.requires_grad_(False)
    d.loss(d(x1, x2), t).backward()
    torch.cuda.synchronize()
print("done")
