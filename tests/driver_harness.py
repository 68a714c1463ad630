#!/usr/bin/env python3
"""Run the reference's UNCHANGED driver scripts end to end on a small synthetic corpus.

    python tests/driver_harness.py --impl {ours,reference} --driver {nplda,dplda} --workdir DIR [--loss L] [--format F]

Test infrastructure (used by tests/test_reference_drivers.py and by `bench.py`'s scoring-loop leg for the file
layout).  It builds, under DIR, exactly the tree the reference's scripts expect relative to their working directory
(`conf/voices_config.cfg` or `conf/voices_config_dplda.cfg`, `logs/`, `models/`, `scores/`, trial TSVs, the x-vector
pickle, `Kaldi_Models/`), chdirs there and then runs, from the verbatim copies in `baseline/_ref/`:

  * `xvector_NeuralPlda_pytorch.main_kaldiplda()` or `xvector_DPlda_pytorch.main_kaldiplda()` -- loaders, Kaldi init,
    threshold initialisation (`validate(update_thresholds=True)`), `train()` / `validate()` for every epoch,
    `SaveModel`, `nc.generate_scorefile` (xvector_NeuralPlda_pytorch.py:88-181);
  * `xvector_generate_scores.py` as a script (`runpy`): `pickle.load` of the saved model + score files
    (xvector_generate_scores.py:19-44; its hard-coded model name `models/NPLDA_13_1586347612.pt` is a copy of the last
    epoch's pickle).

`--impl reference` runs them against the reference's own utils/* on the CPU; the only shims are empty `matplotlib` /
`kaldi_io` modules and `subprocess.check_output` answered from the Kaldi files (the Kaldi binaries are not installed).
`--impl ours` calls `neuralplda_b200.dropin.install()` first, so `utils.models`, `utils.sv_trials_loaders` and
`utils.scorefile_generator` resolve to this package and everything runs on cuda:0.  No script is edited in either arm.
Results (per-batch training losses, validation metrics, thresholds, final parameters, score files) land in DIR/out.
"""
import argparse
import glob
import json
import os
import pickle
import runpy
import shutil
import subprocess
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")


def build_corpus(work, driver, loss, fmt, seed=11):
    """Synthetic speaker-structured x-vectors (SURVEY 8d generator), trial lists and the config file."""
    rng = np.random.RandomState(seed)
    nspk, per = 40, 10
    mean = np.loadtxt(os.path.join(REF, "Kaldi_Models", "mean.vec"), dtype=str)
    mean = np.asarray([float(v) for v in mean if v not in "[]"], dtype=np.float32)
    spk = rng.randn(nspk, 512).astype(np.float32)
    ids, spk_of, mega = [], {}, {}
    for s in range(nspk):
        for u in range(per):
            uid = f"spk{s:03d}-utt{u:02d}"
            ids.append(uid)
            spk_of[uid] = s
            mega[uid] = (mean + spk[s] + 0.7 * rng.randn(512)).astype(np.float32)
    os.makedirs(os.path.join(work, "xvectors"), exist_ok=True)
    with open(os.path.join(work, "xvectors", "mega.pkl"), "wb") as f:
        pickle.dump(mega, f)

    def trials(n, ptarget, ext2=""):
        rows = []
        for _ in range(n):
            a = ids[rng.randint(len(ids))]
            if rng.rand() < ptarget:
                b = ids[spk_of[a] * per + rng.randint(per)]
            else:
                b = ids[rng.randint(len(ids))]
            rows.append((a, b + ext2, "1" if spk_of[a] == spk_of[b] else "0"))
        return rows

    tk = os.path.join(work, "trials_and_keys")
    os.makedirs(tk, exist_ok=True)
    bs = 128 if driver == "nplda" else 64
    files = {"train_a": trials(20 * bs + 37, 0.3), "train_b": trials(6 * bs, 0.2),
             "valid_heldout": trials(1500, 0.2, ".wav"), "valid_other": trials(1100, 0.1, ".wav")}
    for name, rows in files.items():
        np.savetxt(os.path.join(tk, name + ".tsv"), np.asarray(rows), fmt="%s", delimiter="\t", comments="")
    test_rows = [(a + ".wav", "some/dir/" + b + ".wav", c) for a, b, c in trials(5 * bs * 4 + 333, 0.1)]
    test_file = os.path.join(tk, "test_trials.tsv")
    if fmt == "sre":
        np.savetxt(test_file, np.asarray(test_rows), fmt="%s", delimiter="\t", comments="",
                   header="modelid\tsegmentid\tside")
    else:
        np.savetxt(test_file, np.asarray(test_rows), fmt="%s", delimiter="\t", comments="")
    km = os.path.join(work, "Kaldi_Models")
    os.makedirs(km, exist_ok=True)
    for f in ("mean.vec", "transform.mat", "plda"):
        shutil.copyfile(os.path.join(REF, "Kaldi_Models", f), os.path.join(km, f))
    for d in ("conf", "logs", "models", "scores", "out"):
        os.makedirs(os.path.join(work, d), exist_ok=True)
    cfg = f"""[Paths]
training_data_trials_list = trials_and_keys/train_a.tsv,trials_and_keys/train_b.tsv
validation_trials_list = trials_and_keys/valid_heldout.tsv,trials_and_keys/valid_other.tsv
test_trials_list = trials_and_keys/test_trials.tsv
mega_xvector_scp = xvectors/mega.scp
mega_xvector_pkl = xvectors/mega.pkl
meanvec = Kaldi_Models/mean.vec
transformmat = Kaldi_Models/transform.mat
kaldiplda = Kaldi_Models/plda

[NPLDA]
xvector_dim = 512
layer1_LDA_dim = 170
layer2_PLDA_spkfactor_dim = 170
initialization = kaldi
device = {'cuda' if torch.cuda.is_available() else 'cpu'}
seed = 1
alpha = 15

[Training]
train_subsample_factors = 1.01,0.9
valid_subsample_factors = 1.01,1.01
loss = {loss}
cmiss = 1
cfa = 1
target_probs = {'0.01,0.005' if driver == 'nplda' else '0.01'}
batch_size = {bs}
n_epochs = 2
lr = 0.0001
heldout_set_for_th_init = valid_heldout
heldout_set_for_lr_decay = valid_other

[Scoring]
scorefile_format = {fmt}

[Logging]
log_interval = 7
"""
    name = "voices_config.cfg" if driver == "nplda" else "voices_config_dplda.cfg"
    with open(os.path.join(work, "conf", name), "w") as f:
        f.write(cfg)
    with open(os.path.join(work, "conf", "voices_config.cfg"), "w") as f:      # xvector_generate_scores.py:19
        f.write(cfg)


def _stub_modules():
    for name in ("matplotlib", "matplotlib.pyplot", "kaldi_io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]


def _fake_kaldi_binaries():
    """Reference arm only: models.py:441-448 / kaldiPlda2numpydict.py:17 shell out to copy-matrix, copy-vector and
    ivector-copy-plda, which this image does not have.  Answer them from the files (same shim as make_golden.py)."""
    sys.path.insert(0, ROOT)
    from oracle import nplda_oracle as O
    real = subprocess.check_output

    def fmt(v):
        return " ".join(repr(float(x)) for x in v)

    def fake(cmd, *a, **k):
        if cmd[0] == "copy-matrix":
            m = O.read_kaldi_matrix(cmd[2])
            return (" [\n" + "\n".join("  " + fmt(r) for r in m) + " ]\n").encode()
        if cmd[0] == "copy-vector":
            return (" [ " + fmt(O.read_kaldi_vector(cmd[2])) + " ]\n").encode()
        if cmd[0] == "ivector-copy-plda":
            p = O.read_kaldi_plda(cmd[2])
            body = "\n".join("  " + fmt(r) for r in p["diagonalizing_transform"])
            return ("<Plda>  [ " + fmt(p["plda_mean"]) + " ]\n [\n" + body + " ]\n [ "
                    + fmt(p["Psi_across_covar_diag"]) + " ]\n</Plda> \n").encode()
        return real(cmd, *a, **k)

    subprocess.check_output = fake


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", required=True, choices=["ours", "reference"])
    ap.add_argument("--driver", default="nplda", choices=["nplda", "dplda"])
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--loss", default="crossentropy")
    ap.add_argument("--format", default="voices", choices=["voices", "sre"])
    args = ap.parse_args()
    work = os.path.abspath(args.workdir)
    if not os.path.exists(os.path.join(REF, "MANIFEST.json")):
        raise SystemExit("baseline/_ref is missing: run python baseline/install_reference.py where /root/reference exists")
    build_corpus(work, args.driver, args.loss, args.format)
    os.chdir(work)
    if args.impl == "ours":
        sys.path.insert(0, ROOT)
        import neuralplda_b200.dropin as dropin
        dropin.install(REF)
    else:
        sys.path.insert(0, REF)
        _stub_modules()
        _fake_kaldi_binaries()
        torch.set_num_threads(min(8, os.cpu_count() or 1))
    import importlib
    drv = importlib.import_module("xvector_NeuralPlda_pytorch" if args.driver == "nplda" else "xvector_DPlda_pytorch")
    import utils.models as M
    cls = M.NeuralPlda if args.driver == "nplda" else M.DPlda
    record = {"impl": args.impl, "models_module": M.__name__, "train_losses": [], "validate": []}
    # observe without editing: wrap the two driver-level functions' collaborators on the CLASS the driver instantiates
    orig_loss, orig_minc = cls.loss, cls.minc

    def loss_spy(self, output, target):
        out = orig_loss(self, output, target)
        if self.training and out is not None:
            record["train_losses"].append(float(out.detach()))
        return out

    def minc_spy(self, output, target, update_thresholds=False, showplots=False):
        # what validate() itself computed just before calling minc (xvector_NeuralPlda_pytorch.py:67-69), with the
        # thresholds the model holds at that point; hard decisions also as raw counts (a score that EQUALS a threshold
        # -- minc picks thresholds among the scores -- may fall on either side when two implementations differ in the
        # last bits, so the test compares counts with a tolerance of a few trials instead of the normalised cost)
        th_now = [float(self.threshold[b].detach()) for b in self.beta]
        s, tg = output.detach().float().cpu(), target.detach().float().cpu()
        rec = {"n": int(output.numel()), "n_target": int(tg.sum()), "softcdet": float(self.softcdet(output, target)),
               "cdet": float(self.cdet(output, target)), "thresholds_before": th_now,
               "miss_counts": [int(((s < th) & (tg > 0.5)).sum()) for th in th_now],
               "fa_counts": [int(((s > th) & (tg < 0.5)).sum()) for th in th_now],
               "score_mean": float(s.double().mean()), "score_std": float(s.double().std())}
        mc, th = orig_minc(self, output, target, update_thresholds, showplots)
        rec.update({"minc": float(mc), "thresholds": {str(b): float(v) for b, v in th.items()}})
        record["validate"].append(rec)
        return mc, th

    cls.loss, cls.minc = loss_spy, minc_spy
    drv.main_kaldiplda()                                     # the reference's main, unchanged
    cls.loss, cls.minc = orig_loss, orig_minc
    saved = sorted(glob.glob("models/NPLDA_2_*.pt"))
    assert saved, "the driver saved no epoch-2 model"
    shutil.copyfile(saved[-1], "models/NPLDA_13_1586347612.pt")   # the name xvector_generate_scores.py:20-21,39 hard-codes
    runpy.run_path(os.path.join(REF, "xvector_generate_scores.py"), run_name="__main__")
    model = pickle.load(open(saved[-1], "rb"))
    record["model_class"] = type(model).__module__ + "." + type(model).__qualname__
    np.savez(os.path.join("out", "params.npz"), **{k: v.detach().cpu().numpy() for k, v in model.state_dict().items()})
    scores = {}
    for f in sorted(glob.glob("scores/*")):
        key = "_".join(os.path.basename(f).split("_")[:2]) + ("_rescored" if f.endswith("_scores.txt") else "")
        shutil.copyfile(f, os.path.join("out", key + ".txt"))
        scores[key] = f
    record["score_files"] = sorted(scores)
    with open(os.path.join("out", "record.json"), "w") as f:
        json.dump(record, f, indent=1)
    print("HARNESS OK", args.impl, args.driver, len(record["train_losses"]), "train batches,", len(record["validate"]), "validations")


if __name__ == "__main__":
    main()
