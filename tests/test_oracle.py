"""The CPU oracle against the outputs of the unmodified reference (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from oracle import nplda_oracle as O
from conftest import GOLDEN, NC, parity_ok


def P(kp):
    return kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]


def test_inputs_rebuild_exactly(ref_out, cfg1):
    x1, x2, t = cfg1
    assert x1.double().sum().item() == pytest.approx(ref_out["c1_x1_sum"][0], rel=0, abs=0)
    assert x2.double().abs().sum().item() == pytest.approx(ref_out["c1_x2_sum"][1], rel=0, abs=0)
    assert t.sum().item() == ref_out["c1_t_sum"][0]


def test_kaldi_init_known_answers(kaldi_params):
    # diagP[0]=0.4911, diagQ[0]=-0.4740 from psi[0]=27.70
    assert kaldi_params["P_sqrt"][0].item() ** 2 == pytest.approx(0.4911, abs=2e-4)
    assert kaldi_params["Q"][0].item() == pytest.approx(-0.4740, abs=2e-4)
    assert kaldi_params["W1"].shape == (170, 512) and kaldi_params["W2"].shape == (170, 170)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Kaldi_Models"), reason="reference files only in the build container")
def test_kaldi_reader_matches_reference_init(kaldi_params):
    km = "/root/reference/Kaldi_Models"
    p = O.kaldi_init_params(km + "/mean.vec", km + "/transform.mat", km + "/plda")
    for k in ("W1", "b1", "W2", "b2", "P_sqrt", "Q"):
        np.testing.assert_array_equal(p[k], kaldi_params[k].numpy())


def test_nplda_scores(ref_out, kaldi_params, cfg1):
    x1, x2, _ = cfg1
    s = O.nplda_score(x1, x2, *P(kaldi_params))
    ref = torch.from_numpy(ref_out["c1_scores"])
    ok, worst = parity_ok(s, ref, rel=2e-6)
    assert ok, worst


def test_losses(ref_out, kaldi_params, cfg1):
    _, _, t = cfg1
    s = torch.from_numpy(ref_out["c1_scores"])
    th0 = [0.0, 0.0]
    assert O.softcdet(s, t, th0, NC.beta, NC.alpha).item() == pytest.approx(float(ref_out["c1_softcdet_th0"]), rel=1e-6)
    assert O.cdet(s, t, th0, NC.beta).item() == pytest.approx(float(ref_out["c1_cdet_th0"]), rel=1e-6)
    assert O.crossentropy(s, t, 0.0).item() == pytest.approx(float(ref_out["c1_bce_th0"]), rel=1e-6)
    th = [float(v) for v in ref_out["c1_th_state"]]
    assert O.softcdet(s, t, th, NC.beta, NC.alpha).item() == pytest.approx(float(ref_out["c1_softcdet_thminc"]), rel=1e-6)
    assert O.cdet(s, t, th, NC.beta).item() == pytest.approx(float(ref_out["c1_cdet_thminc"]), rel=1e-6)


def test_accumulators_reproduce_losses(ref_out, cfg1):
    _, _, t = cfg1
    s = torch.from_numpy(ref_out["c1_scores"])
    th = [float(v) for v in ref_out["c1_th_state"]]
    acc = O.loss_accumulators(s, t, th, NC.beta, NC.alpha, 0.0)
    K = 2
    nt, nn_, sb, n = acc[4 * K:].tolist()
    soft = sum(acc[4 * k].item() / nt + NC.beta[k] * acc[4 * k + 1].item() / nn_ for k in range(K)) / K
    hard = sum(acc[4 * k + 2].item() / nt + NC.beta[k] * acc[4 * k + 3].item() / nn_ for k in range(K)) / K
    assert soft == pytest.approx(float(ref_out["c1_softcdet_thminc"]), rel=1e-5)
    assert hard == pytest.approx(float(ref_out["c1_cdet_thminc"]), rel=1e-6)
    assert sb / n == pytest.approx(float(ref_out["c1_bce_th0"]), rel=1e-5)


def test_minc(ref_out, cfg1):
    _, _, t = cfg1
    s = torch.from_numpy(ref_out["c1_scores"])
    for fn in (O.minc, O.minc_loop):
        mc, th = fn(s[:2000], t[:2000], NC.beta)
        assert float(mc) == pytest.approx(float(ref_out["c1_minc2k"]), rel=1e-6)
        np.testing.assert_array_equal(np.asarray([float(th[b]) for b in NC.beta], dtype=np.float32),
                                      ref_out["c1_minc2k_th"].astype(np.float32))


def test_minc_quirks(ref_out):
    toy_s = torch.tensor([.1, .5, .9, -.2, .3, .7, -1.])
    toy_t = torch.tensor([1., 1., 1., 0., 0., 0., 0.])
    for fn in (O.minc, O.minc_loop):
        mc, th = fn(toy_s, toy_t, NC.beta)
        assert float(mc) == pytest.approx(float(ref_out["c5_minc"]), rel=1e-6)
        assert [float(th[b]) for b in NC.beta] == pytest.approx(list(ref_out["c5_th"]))
    g = torch.Generator().manual_seed(5)
    qs = torch.round(torch.randn(400, generator=g) * 4) / 4
    qt = (torch.rand(400, generator=g) < 0.3).float()
    for fn in (O.minc, O.minc_loop):
        mc, th = fn(qs, qt, NC.beta)
        assert float(mc) == pytest.approx(float(ref_out["c5b_minc"]), rel=1e-6)
        assert [float(th[b]) for b in NC.beta] == pytest.approx(list(ref_out["c5b_th"]))


def dplda_weights(ref_out):
    g = torch.Generator().manual_seed(int(ref_out["c4_seed_w"]))
    w = (torch.rand(1, 57970, generator=g) - 0.5) * 0.2
    return w, torch.tensor([0.3])


def test_dplda_scores(ref_out, kaldi_params, cfg1):
    x1, x2, t = cfg1
    w, c = dplda_weights(ref_out)
    ref = torch.from_numpy(ref_out["c4_scores"])
    n = ref.numel()
    for expanded in (True, False):
        s = O.dplda_score(x1[:n], x2[:n], kaldi_params["W1"], kaldi_params["b1"], w, c, expanded=expanded)
        ok, worst = parity_ok(s, ref, rel=5e-6)
        assert ok, (expanded, worst)
    assert O.crossentropy(ref, t[:n], 0.0).item() == pytest.approx(float(ref_out["c4_bce"]), rel=1e-6)
    assert O.softcdet(ref, t[:n], [0.0], [99.0], 15.0).item() == pytest.approx(float(ref_out["c4_softcdet"]), rel=1e-6)


def test_training_gradients_via_autograd_of_oracle(ref_out, kaldi_params, cfg1):
    """The oracle differentiated by autograd reproduces the reference's .grad (case 2)."""
    x1, x2, t = cfg1
    xb1, xb2, tb = x1[:2048], x2[:2048], t[:2048]
    names = ["P_sqrt", "Q", "Th99", "Th199", "threshold_Xent", "centering_and_LDA.weight",
             "centering_and_LDA.bias", "centering_and_wccn_plda.weight", "centering_and_wccn_plda.bias"]
    th = [float(v) for v in ref_out["c1_th_state"]]
    for lossname in ("SoftCdet", "crossentropy"):
        p = {k: kaldi_params[k].clone().requires_grad_(True) for k in ("W1", "b1", "W2", "b2", "P_sqrt", "Q")}
        ths = [torch.tensor([v], requires_grad=True) for v in th]
        thx = torch.tensor([0.25], requires_grad=True)
        s = O.nplda_score(xb1, xb2, p["W1"], p["b1"], p["W2"], p["b2"], p["P_sqrt"], p["Q"])
        loss = O.softcdet(s, tb, ths, NC.beta, NC.alpha) if lossname == "SoftCdet" else O.crossentropy(s, tb, thx)
        loss.backward()
        assert loss.item() == pytest.approx(float(ref_out[f"c2_{lossname}_loss"]), rel=1e-5)
        got = dict(zip(names, [p["P_sqrt"], p["Q"], ths[0], ths[1], thx, p["W1"], p["b1"], p["W2"], p["b2"]]))
        for n in names:
            sample = ref_out[f"c2_{lossname}_grad_{n}_sample"]
            if sample.size == 0:
                assert got[n].grad is None
                continue
            g = got[n].grad.reshape(-1)
            scale = float(ref_out[f"c2_{lossname}_grad_{n}_norm"]) / np.sqrt(g.numel()) + 1e-30
            np.testing.assert_allclose(g[::7].numpy(), sample, rtol=1e-3, atol=1e-3 * scale)


def test_gather_and_scorefile_formats(ref_out, kaldi_params):
    z = np.load(os.path.join(GOLDEN, "c6_mega.npz"))
    ids = [str(s) for s in z["ids"]]
    mega = {u: z["vecs"][i] for i, u in enumerate(ids)}
    num_to_id = dict(enumerate(ids))
    X1, X2 = O.gather_numbatch(mega, num_to_id, ref_out["c6_d1"], ref_out["c6_d2"])
    assert X1.double().sum().item() == pytest.approx(ref_out["c6_x1_sum"][0], rel=1e-12)
    assert X2.double().abs().sum().item() == pytest.approx(ref_out["c6_x2_sum"][1], rel=1e-12)
    # voices score file: enrol <TAB> test <TAB> str(np.float32 score), ids basename/splitext-stripped for lookup
    lines = open(os.path.join(GOLDEN, "c6_voices_scores.txt")).read().strip().split("\n")
    trials = [l.split() for l in open(os.path.join(GOLDEN, "c6_voices_trials.txt")).read().strip().split("\n")]
    a = torch.from_numpy(np.stack([mega[O.strip_id(tr[0])] for tr in trials]))
    b = torch.from_numpy(np.stack([mega[O.strip_id(tr[1])] for tr in trials]))
    kp = kaldi_params
    s = O.nplda_score(a, b, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    for line, tr, sc in zip(lines, trials, s.tolist()):
        e, tst, val = line.split("\t")
        assert (e, tst) == (tr[0], tr[1])
        assert float(val) == pytest.approx(sc, rel=2e-5, abs=2e-6)
    assert O.format_scores(np.float32([0.5]))[0] == "0.5"


def _read_norm_golden():
    raw = np.genfromtxt(os.path.join(GOLDEN, "c7_raw_scores.tsv"), dtype=str)[1:]
    coh = np.genfromtxt(os.path.join(GOLDEN, "c7_cohort_scores.tsv"), dtype=str, skip_header=1)
    c = len(np.unique(coh[:, 1]))
    ids = list(coh[:, 0].reshape(-1, c)[:, 0])
    mat = coh[:, -1].astype(float).reshape(-1, c)
    er = np.asarray([ids.index(e) for e in raw[:, 0]])
    tr = np.asarray([ids.index(t.replace(".sph", "")) for t in raw[:, 1]])
    want = np.stack([np.genfromtxt(os.path.join(GOLDEN, "c7_raw_scores.tsv_%s.tsv" % n), dtype=str)[:, -1].astype(float)
                     for n in ("znorm", "tnorm", "snorm", "asnorm1")])
    return raw[:, -1].astype(float), er, tr, mat, want


def test_score_normalisation_matches_reference_script():
    """oracle vs the four files the unmodified adaptive_score_normalization.py wrote
    (tests/golden/make_golden_norm.py); float64 throughout, so 1e-12 relative."""
    raw, er, tr, mat, want = _read_norm_golden()
    assert mat.shape[1] > 500                                   # the top-N slice is a strict subset
    got = O.score_norm(raw, er, tr, O.cohort_stats(mat, 500))
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
