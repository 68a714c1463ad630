"""The reference's own driver scripts, UNCHANGED, on top of this package (north_star: "so xvector_NeuralPlda_pytorch.py
and xvector_generate_scores.py call it unchanged").

`baseline/_ref/` holds verbatim copies of the reference files (baseline/install_reference.py, sha256 manifest).
tests/driver_harness.py runs `main_kaldiplda()` of xvector_NeuralPlda_pytorch.py / xvector_DPlda_pytorch.py and then
xvector_generate_scores.py as scripts, once against the reference's own utils/* on the CPU and once with
`neuralplda_b200.dropin.install()` aliasing utils.models / utils.sv_trials_loaders / utils.scorefile_generator to this
package on cuda:0.  The two runs must agree: per-batch training losses, every validation metric (minC, thresholds,
softCdet, Cdet), the parameters after two epochs of Adam and every score file.
"""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(ROOT, "baseline", "_ref")
sys.path.insert(0, ROOT)

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "MANIFEST.json")),
                               reason="baseline/_ref not installed (python baseline/install_reference.py needs /root/reference)")


@needs_ref
def test_reference_copy_is_verbatim():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import install_reference
    assert install_reference.verify(), "a file under baseline/_ref differs from the manifest written at install time"
    if os.path.isdir("/root/reference"):                     # build container: compare with the source tree itself
        for rel in install_reference.FILES:
            assert open(os.path.join(REF, rel), "rb").read() == open(os.path.join("/root/reference", rel), "rb").read(), rel


_MAKE_REF_PICKLE = r"""
import sys, types, pickle, torch
for name in ("matplotlib", "matplotlib.pyplot", "kaldi_io"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, sys.argv[1])
from utils.models import NeuralPlda, DPlda
class NC:
    xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = 512, 170, 170
    alpha, device, beta, loss = 15.0, "cpu", [99.0, 199.0], "crossentropy"
torch.manual_seed(5)
m = NeuralPlda(NC)
m.threshold[99.0].data.fill_(0.25)
m.SaveModel(sys.argv[2] + "/ref_nplda.pt")
d = DPlda(NC)
d.SaveModel(sys.argv[2] + "/ref_dplda.pt")
torch.save({"n": m.state_dict(), "d": d.state_dict()}, sys.argv[2] + "/ref_sd.pt")
"""


@needs_ref
def test_dropin_aliases_and_reference_pickles(tmp_path):
    """utils.models resolves to this package after dropin.install(); a model pickled by the REFERENCE class (class path
    utils.models.NeuralPlda, models.py:459-461) unpickles into the drop-in class with identical parameters, parameter
    order and threshold aliasing (xvector_generate_scores.py:39); and pickles written here under the reference's class
    path load the same way."""
    subprocess.run([sys.executable, "-c", _MAKE_REF_PICKLE, REF, str(tmp_path)], check=True)
    import neuralplda_b200.dropin as dropin
    import neuralplda_b200.models as ours
    try:
        dropin.install(REF)
        import utils.models as um
        import utils.sv_trials_loaders as ul
        import utils.scorefile_generator as us
        assert um is ours and ul.__name__ == "neuralplda_b200.sv_trials_loaders" and us.__name__ == "neuralplda_b200.scorefile_generator"
        from utils.NpldaConf import NpldaConf                   # the REFERENCE's config class, importing our writers
        import utils.NpldaConf as unc
        assert unc.generate_voices_scores is us.generate_voices_scores
        assert os.path.samefile(unc.__file__, os.path.join(REF, "utils", "NpldaConf.py"))
        import importlib
        drv = importlib.import_module("xvector_NeuralPlda_pytorch")     # the driver, unchanged
        assert drv.NeuralPlda is ours.NeuralPlda
        assert drv.load_xvec_trials_from_numbatch is ul.load_xvec_trials_from_numbatch
        ref_sd = torch.load(str(tmp_path / "ref_sd.pt"))
        for fn, cls, key in (("ref_nplda.pt", ours.NeuralPlda, "n"), ("ref_dplda.pt", ours.DPlda, "d")):
            m = pickle.load(open(tmp_path / fn, "rb"))
            assert type(m) is cls
            sd = m.state_dict()
            assert list(sd) == list(ref_sd[key])
            for k in sd:
                assert torch.equal(sd[k], ref_sd[key][k]), k
            for beta in m.beta:
                assert m.threshold[beta] is m._parameters["Th{}".format(int(beta))]
            assert m.lossfn == "crossentropy" and m.beta == [99.0, 199.0]
        m = pickle.load(open(tmp_path / "ref_nplda.pt", "rb"))
        assert float(m.threshold[99.0]) == 0.25
        dropin.pickle_as_reference(m, str(tmp_path / "ours_as_ref.pt"))
        raw = open(tmp_path / "ours_as_ref.pt", "rb").read()
        assert b"utils.models" in raw and b"neuralplda_b200" not in raw
        m2 = pickle.load(open(tmp_path / "ours_as_ref.pt", "rb"))
        assert type(m2) is ours.NeuralPlda and torch.equal(m2.Q, m.Q)
    finally:
        dropin.uninstall()
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k.startswith("xvector_")]:
            sys.modules.pop(k, None)
        if REF in sys.path:
            sys.path.remove(REF)


def _keep(work, tag):
    """DRIVER_HARNESS_KEEP=dir keeps the harness records of both arms (for the round's profiles/)."""
    keep = os.environ.get("DRIVER_HARNESS_KEEP")
    if keep:
        import shutil
        dst = os.path.join(keep, tag)
        os.makedirs(dst, exist_ok=True)
        for f in ("record.json",):
            shutil.copyfile(os.path.join(work, "out", f), os.path.join(dst, f))


def _run(impl, driver, work, loss, fmt):
    cmd = [sys.executable, os.path.join(HERE, "driver_harness.py"), "--impl", impl, "--driver", driver, "--workdir", work,
           "--loss", loss, "--format", fmt]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "HARNESS OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
    rec = json.load(open(os.path.join(work, "out", "record.json")))
    _keep(work, f"{driver}_{loss}_{fmt}_{impl}")
    return rec, dict(np.load(os.path.join(work, "out", "params.npz")))


def _scores(path):
    rows = [ln.rstrip("\n").split("\t") for ln in open(path)]
    return rows


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("driver,loss,fmt", [("nplda", "crossentropy", "voices"), ("nplda", "SoftCdet", "sre"),
                                              ("dplda", "crossentropy", "voices")])
def test_unchanged_drivers_match_reference(tmp_path, driver, loss, fmt):
    ref, ref_p = _run("reference", driver, str(tmp_path / "ref"), loss, fmt)
    our, our_p = _run("ours", driver, str(tmp_path / "ours"), loss, fmt)
    assert ref["models_module"] == "utils.models" and our["models_module"] == "neuralplda_b200.models"
    assert our["model_class"].startswith("neuralplda_b200.models.")
    # same batches in the same order (same seeds, same RNG draw order in constructors and loaders)
    assert len(ref["train_losses"]) == len(our["train_losses"]) > 40
    a, b = np.asarray(ref["train_losses"]), np.asarray(our["train_losses"])
    assert np.allclose(a, b, rtol=2e-3, atol=1e-5), float(np.abs(a - b).max())
    assert len(ref["validate"]) == len(our["validate"]) == 7        # threshold init + 3 x 2 sets
    for va, vb in zip(ref["validate"], our["validate"]):
        assert va["n"] == vb["n"] and va["n_target"] == vb["n_target"]
        for key in ("minc", "softcdet"):
            assert abs(va[key] - vb[key]) <= 2e-3 * max(abs(va[key]), 1e-2) + 2e-3, (key, va, vb)
        for ca, cb in zip(va["miss_counts"] + va["fa_counts"], vb["miss_counts"] + vb["fa_counts"]):
            assert abs(ca - cb) <= 2, (va, vb)                       # hard decisions: scores that tie with a threshold
        for ta, tb in zip(va["thresholds_before"], vb["thresholds_before"]):
            assert abs(ta - tb) <= 5e-3, (va, vb)
        for beta in va["thresholds"]:
            assert abs(va["thresholds"][beta] - vb["thresholds"][beta]) <= 5e-3, (beta, va, vb)
    assert list(ref_p) == list(our_p)
    for k in ref_p:
        scale = max(float(np.abs(ref_p[k]).max()), 1e-3)
        assert float(np.abs(ref_p[k] - our_p[k]).max()) <= 5e-3 * scale, k
    assert ref["score_files"] == our["score_files"] and len(ref["score_files"]) == 3
    for key in ref["score_files"]:
        ra, rb = _scores(os.path.join(str(tmp_path / "ref"), "out", key + ".txt")), _scores(os.path.join(str(tmp_path / "ours"), "out", key + ".txt"))
        assert len(ra) == len(rb) > 1000
        first = 1 if fmt == "sre" else 0
        if first:
            assert ra[0] == rb[0]                                    # header row + LLR
        assert [r[:-1] for r in ra[first:]] == [r[:-1] for r in rb[first:]]      # id columns, byte for byte
        sa = np.asarray([float(r[-1]) for r in ra[first:]])
        sb = np.asarray([float(r[-1]) for r in rb[first:]])
        bound = 5e-3 * np.maximum(np.abs(sa), np.sqrt(np.mean(sa ** 2)))
        assert bool((np.abs(sa - sb) <= bound).all()), float((np.abs(sa - sb) / bound).max())
    # the model re-scored by xvector_generate_scores.py (pickle.load) is the epoch-2 model: identical scores
    pa = _scores(os.path.join(str(tmp_path / "ours"), "out", "kaldipldanet_epoch2.txt"))
    pb = _scores(os.path.join(str(tmp_path / "ours"), "out", "kaldipldanet_epoch13_rescored.txt"))
    assert pa == pb
