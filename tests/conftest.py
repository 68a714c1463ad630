import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class NC:
    """Duck-typed NpldaConf (reference utils/NpldaConf.py) with conf/sre_config.cfg values."""
    xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = 512, 170, 170
    alpha, device = 15.0, "cpu"
    beta = [99.0, 199.0]
    loss = "SoftCdet"


class NCD(NC):
    beta = [99.0]
    loss = "crossentropy"


@pytest.fixture(scope="session")
def ref_out():
    return np.load(os.path.join(GOLDEN, "reference_outputs.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def kaldi_params():
    z = np.load(os.path.join(GOLDEN, "kaldi_init_params.npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


@pytest.fixture(scope="session")
def cfg1(kaldi_params):
    """The 10k-pair synthetic workload of BASELINE.json configs[0], rebuilt from its seed."""
    from oracle import nplda_oracle as O
    return O.synth_pairs(10000, 200, seed=1001, mean=kaldi_params["mean"])


def parity_ok(s, s_ref, rel=1e-4):
    """|S - S_ref| <= rel * max(|S_ref|, rms(S_ref))   (SURVEY.md section 8d)."""
    s, s_ref = s.double().cpu(), s_ref.double().cpu()
    rms = s_ref.pow(2).mean().sqrt()
    bound = rel * torch.maximum(s_ref.abs(), rms)
    err = (s - s_ref).abs()
    return bool((err <= bound).all()), float((err / bound).max())
