#!/usr/bin/env python3
"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors for this path,
so parity is pinned on outputs of the reference code itself: utils/models.py
(NeuralPlda, DPlda), utils/sv_trials_loaders.py and utils/scorefile_generator.py
are imported as they are, with empty stub modules for the two imports that are
absent in this image and unused on the path (matplotlib, kaldi_io), and with
subprocess.check_output answered from a pure-Python reader of the Kaldi model
files (the Kaldi binaries copy-matrix / copy-vector / ivector-copy-plda are
not installed) so that LoadPldaParamsFromKaldi (models.py:441-457) and
kaldiPlda2numpydict (kaldiPlda2numpydict.py:16-38) run their own parsing,
slicing and diagP/diagQ code.

Inputs are NOT stored: tests rebuild them from the seeds recorded here with
oracle.nplda_oracle.synth_pairs (CPU torch.Generator, deterministic for this
image's torch build); a checksum of every input tensor is stored instead.
"""
import os
import subprocess
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
for name in ("matplotlib", "matplotlib.pyplot", "kaldi_io"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, REF)

from oracle import nplda_oracle as O  # noqa: E402

# ---- answer the three Kaldi commands the reference shells out to -----------
_real_check_output = subprocess.check_output


def _fmt(v):
    return " ".join(repr(float(x)) for x in v)


def _fake_check_output(cmd, *a, **k):
    if cmd[0] == "copy-matrix":
        m = O.read_kaldi_matrix(cmd[2])
        body = "\n".join("  " + _fmt(r) for r in m)
        return (" [\n" + body + " ]\n").encode()
    if cmd[0] == "copy-vector":
        return (" [ " + _fmt(O.read_kaldi_vector(cmd[2])) + " ]\n").encode()
    if cmd[0] == "ivector-copy-plda":
        p = O.read_kaldi_plda(cmd[2])
        body = "\n".join("  " + _fmt(r) for r in p["diagonalizing_transform"])
        txt = "<Plda>  [ " + _fmt(p["plda_mean"]) + " ]\n [\n" + body + " ]\n [ " \
            + _fmt(p["Psi_across_covar_diag"]) + " ]\n</Plda> \n"
        return txt.encode()
    return _real_check_output(cmd, *a, **k)


subprocess.check_output = _fake_check_output

from utils.models import NeuralPlda, DPlda  # noqa: E402  (the reference)
from utils import sv_trials_loaders as ref_loaders  # noqa: E402
from utils import scorefile_generator as ref_scorefile  # noqa: E402


class NC:
    """Duck-typed NpldaConf with conf/sre_config.cfg values (15-21, 26-30)."""
    xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = 512, 170, 170
    alpha, device = 15.0, "cpu"
    beta = [99.0, 199.0]
    loss = "SoftCdet"


def checksum(t):
    return float(t.double().sum()), float(t.double().abs().sum())


def subsample_grad(g):
    if g is None:   # parameter not reached by this loss (autograd leaves .grad None)
        return {"sum": float("nan"), "norm": float("nan"), "sample": np.zeros(0, np.float32)}
    g = g.detach().reshape(-1)
    return {"sum": float(g.double().sum()), "norm": float(g.double().norm()),
            "sample": g[::7].numpy().copy()}


def main():
    km = os.path.join(REF, "Kaldi_Models")
    out = {}

    # ---------------- case K: Kaldi initialisation ---------------------------
    torch.manual_seed(1)
    model = NeuralPlda(NC)
    model.LoadPldaParamsFromKaldi(km + "/mean.vec", km + "/transform.mat", km + "/plda")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    np.savez_compressed(
        os.path.join(HERE, "kaldi_init_params.npz"),
        W1=sd["centering_and_LDA.weight"].numpy(), b1=sd["centering_and_LDA.bias"].numpy(),
        W2=sd["centering_and_wccn_plda.weight"].numpy(), b2=sd["centering_and_wccn_plda.bias"].numpy(),
        P_sqrt=sd["P_sqrt"].numpy(), Q=sd["Q"].numpy(),
        mean=O.read_kaldi_vector(km + "/mean.vec").astype(np.float32))
    mean = torch.from_numpy(O.read_kaldi_vector(km + "/mean.vec").astype(np.float32))
    out["param_names"] = np.asarray([n for n, _ in model.named_parameters()])

    # ---------------- case 1: cfg1, 10k pairs, NeuralPlda forward + losses ----
    x1, x2, t = O.synth_pairs(10000, 200, seed=1001, mean=mean)
    out["c1_seed"], out["c1_n"], out["c1_spk"] = 1001, 10000, 200
    out["c1_x1_sum"], out["c1_x2_sum"], out["c1_t_sum"] = checksum(x1), checksum(x2), checksum(t)
    model.eval()
    with torch.no_grad():
        s = model.forward(x1, x2)
        out["c1_scores"] = s.numpy().copy()
        out["c1_softcdet_th0"] = model.softcdet(s, t).item()
        out["c1_cdet_th0"] = model.cdet(s, t).item()
        out["c1_bce_th0"] = model.crossentropy(s, t).item()
        mc, mth = model.minc(s[:2000], t[:2000], update_thresholds=True)
        out["c1_minc2k"] = float(mc)
        out["c1_minc2k_th"] = np.asarray([float(mth[b]) for b in NC.beta])
        out["c1_softcdet_thminc"] = model.softcdet(s, t).item()
        out["c1_cdet_thminc"] = model.cdet(s, t).item()
        sd2 = model.state_dict()
        out["c1_th_state"] = np.asarray([float(sd2["Th99"]), float(sd2["Th199"])])

    # ---------------- case 2: training step, 2048 pairs, both losses ----------
    xb1, xb2, tb = x1[:2048], x2[:2048], t[:2048]
    for lossname in ("SoftCdet", "crossentropy"):
        model.zero_grad()
        model.lossfn = lossname
        with torch.no_grad():
            model.threshold_Xent.fill_(0.25)
        loss = model.loss(model(xb1, xb2), tb)
        loss.backward()
        out[f"c2_{lossname}_loss"] = loss.item()
        for n, p in model.named_parameters():
            g = subsample_grad(p.grad)
            for k, v in g.items():
                out[f"c2_{lossname}_grad_{n}_{k}"] = v
    with torch.no_grad():
        model.threshold_Xent.fill_(0.0)

    # ---------------- case 3: default-init NeuralPlda (torch.manual_seed(1)) --
    torch.manual_seed(1)
    m3 = NeuralPlda(NC)
    g0 = torch.Generator().manual_seed(0)
    z1, z2 = torch.randn(256, 512, generator=g0), torch.randn(256, 512, generator=g0)
    with torch.no_grad():
        out["c3_scores"] = m3(z1, z2).numpy().copy()
    np.savez_compressed(os.path.join(HERE, "default_init_params.npz"),
                        **{k: v.numpy() for k, v in m3.state_dict().items()})

    # ---------------- case 4: DPlda, literal 57970-d expansion ----------------
    class NCD(NC):
        beta = [99.0]
        loss = "crossentropy"
    torch.manual_seed(1)
    md = DPlda(NCD)
    md.LoadParamsFromKaldi(km + "/mean.vec", km + "/transform.mat")
    gd = torch.Generator().manual_seed(77)
    with torch.no_grad():
        # a non-degenerate logistic-regression weight (default init is ~4e-3 uniform)
        md.logistic_regres.weight.copy_((torch.rand(1, 57970, generator=gd) - 0.5) * 0.2)
        md.logistic_regres.bias.fill_(0.3)
    out["c4_seed_w"] = 77
    n4 = 768
    with torch.no_grad():
        sdp = md(x1[:n4], x2[:n4])
        out["c4_scores"] = sdp.numpy().copy()
        out["c4_bce"] = md.crossentropy(sdp, t[:n4]).item()
        out["c4_softcdet"] = md.softcdet(sdp, t[:n4]).item()
    md.zero_grad()
    n4b = 256
    loss = md.loss(md(x1[:n4b], x2[:n4b]), t[:n4b])
    loss.backward()
    out["c4_train_loss"] = loss.item()
    for n, p in md.named_parameters():
        g = subsample_grad(p.grad)
        for k, v in g.items():
            out[f"c4_grad_{n}_{k}"] = v
    out["c4_param_names"] = np.asarray([n for n, _ in md.named_parameters()])

    # ---------------- case 5: minc quirks on toy inputs -----------------------
    toy_s = torch.tensor([.1, .5, .9, -.2, .3, .7, -1.])
    toy_t = torch.tensor([1., 1., 1., 0., 0., 0., 0.])
    mc, mth = model.minc(toy_s, toy_t)
    out["c5_minc"] = float(mc)
    out["c5_th"] = np.asarray([float(mth[b]) for b in NC.beta])
    gq = torch.Generator().manual_seed(5)
    qs = torch.round(torch.randn(400, generator=gq) * 4) / 4      # many ties
    qt = (torch.rand(400, generator=gq) < 0.3).float()
    mc, mth = model.minc(qs, qt)
    out["c5b_minc"] = float(mc)
    out["c5b_th"] = np.asarray([float(mth[b]) for b in NC.beta])

    # ---------------- case 6: loaders + score files ---------------------------
    gi = torch.Generator().manual_seed(6)
    ids = [f"utt{i:03d}" for i in range(40)]
    mega = {u: torch.randn(512, generator=gi).numpy() + mean.numpy() for u in ids}
    num_to_id = dict(enumerate(list(mega)))
    d1 = torch.randint(0, 40, (33,), generator=gi)
    d2 = torch.randint(0, 40, (33,), generator=gi)
    X1, X2 = ref_loaders.load_xvec_trials_from_numbatch(mega, num_to_id, d1, d2, torch.device("cpu"))
    out["c6_d1"], out["c6_d2"] = d1.numpy(), d2.numpy()
    out["c6_x1_sum"], out["c6_x2_sum"] = checksum(X1), checksum(X2)
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    voices_trials = os.path.join(HERE, "c6_voices_trials.txt")
    sre_trials = os.path.join(HERE, "c6_sre_trials.tsv")
    with open(voices_trials, "w") as f:
        for a, b in zip(d1.tolist(), d2.tolist()):
            f.write(f"{ids[a]} wav/{ids[b]}.wav {'tgt' if a % 3 == 0 else 'imp'}\n")
    with open(sre_trials, "w") as f:
        f.write("modelid\tsegmentid\tside\n")
        for a, b in zip(d1.tolist(), d2.tolist()):
            f.write(f"{ids[a]}\t{ids[b]}.sph\ta\n")
    model.lossfn = "SoftCdet"
    ref_scorefile.generate_voices_scores(os.path.join(HERE, "c6_voices_scores.txt"), voices_trials,
                                         mega, model, torch.device("cpu"), batch_size=10)
    ref_scorefile.generate_sre_scores(os.path.join(HERE, "c6_sre_scores.tsv"), sre_trials,
                                      mega, model, torch.device("cpu"), batch_size=10)
    np.savez_compressed(os.path.join(HERE, "c6_mega.npz"), ids=np.asarray(ids),
                        vecs=np.stack([mega[u] for u in ids]).astype(np.float32))

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote goldens:", sorted(os.listdir(HERE)))
    print("first scores", out["c1_scores"][:4], "softcdet", out["c1_softcdet_th0"],
          "bce", out["c1_bce_th0"], "c3", out["c3_scores"][:4])


if __name__ == "__main__":
    main()
