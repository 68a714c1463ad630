"""Does the tcgen05 score kernel handle other input widths (layer-1 stage counts)?  TC vs fp32 SIMT on random parameters."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
dev = torch.device("cuda:0")
for d_in in (192, 64, 32, 1024):
    class C:
        xvector_dim = d_in; layer1_LDA_dim = 170; layer2_PLDA_spkfactor_dim = 170
        alpha = 15.0; beta = [99.0, 199.0]; device = dev; loss = "SoftCdet"
    torch.manual_seed(d_in)
    m = npl.NeuralPlda(C).to(dev)
    x1, x2 = torch.randn(5000 + 13, d_in, device=dev), torch.randn(5000 + 13, d_in, device=dev)
    with torch.no_grad():
        m.impl = npl.IMPL_SIMT; s0 = m(x1, x2)
        m.impl = npl.IMPL_TC; s1 = m(x1, x2)
    torch.cuda.synchronize()
    rms = s0.pow(2).mean().sqrt()
    print(f"d_in={d_in}: worst |tc - simt| / (1e-4 * max(|s|, rms)) = {float(((s1 - s0).abs() / (1e-4 * torch.maximum(s0.abs(), rms))).max()):.3f}")
