"""Parity of the CUDA path (through the C ABI) against the oracle and the
reference's golden outputs.  Run on the B200 box: pytest -m gpu."""
import ctypes
import os
import pickle

import numpy as np
import pytest
import torch

import neuralplda_b200 as npl
from neuralplda_b200 import _lib, functional as F_
from oracle import nplda_oracle as O
from conftest import GOLDEN, NC, NCD, parity_ok

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_nplda(kp, impl=None, loss="SoftCdet"):
    class C(NC):
        pass
    C.loss = loss
    m = npl.NeuralPlda(C).to(DEV)
    sd = m.state_dict()
    sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
    sd["centering_and_wccn_plda.weight"].copy_(kp["W2"]); sd["centering_and_wccn_plda.bias"].copy_(kp["b2"])
    sd["P_sqrt"].copy_(kp["P_sqrt"]); sd["Q"].copy_(kp["Q"])
    if impl is not None:
        m.impl = impl
    return m


def dplda_weights(ref_out):
    g = torch.Generator().manual_seed(int(ref_out["c4_seed_w"]))
    return (torch.rand(1, 57970, generator=g) - 0.5) * 0.2, torch.tensor([0.3])


def make_dplda(kp, ref_out, impl=None):
    m = npl.DPlda(NCD).to(DEV)
    w, c = dplda_weights(ref_out)
    sd = m.state_dict()
    sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
    sd["logistic_regres.weight"].copy_(w); sd["logistic_regres.bias"].copy_(c)
    if impl is not None:
        m.impl = impl
    return m


_BWD = {"gemm": 0, "emit": 0, "du": 0}
_CODE = {None: 0, "simt": 1, "tc": 2, "0": 1, "1": 2}


def bwd_paths(**kw):
    """Force pieces of the backward (nplda_debug_backward_paths): gemm="simt"|"tc", emit="0"|"1", du="0"|"1"; None = auto."""
    for k, v in kw.items():
        _BWD[k] = _CODE[v]
    _lib.lib().nplda_debug_backward_paths(_BWD["gemm"], _BWD["emit"], _BWD["du"])


@pytest.fixture(autouse=True)
def _reset_bwd_paths():
    yield
    bwd_paths(gemm=None, emit=None, du=None)


IMPLS = [npl.IMPL_SIMT, npl.IMPL_AUTO]       # AUTO = tcgen05 kernel for the reference dims (512-170-170)


@pytest.mark.parametrize("impl", IMPLS)
def test_nplda_forward_golden_10k(ref_out, kaldi_params, cfg1, impl):
    """BASELINE.json configs[0]: 10k pairs vs the reference's CPU forward, 1e-4 relative."""
    x1, x2, _ = cfg1
    m = make_nplda(kaldi_params, impl)
    with torch.no_grad():
        s = m(x1.to(DEV), x2.to(DEV))
    ok, worst = parity_ok(s, torch.from_numpy(ref_out["c1_scores"]), rel=1e-4)
    assert ok, f"worst normalised error {worst} (bound 1.0)"


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 129, 1000, 4097])
def test_nplda_forward_ragged_sizes(kaldi_params, cfg1, impl, n):
    x1, x2, _ = cfg1
    m = make_nplda(kaldi_params, impl)
    kp = kaldi_params
    ref = O.nplda_score(x1[:n], x2[:n], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    with torch.no_grad():
        s = m(x1[:n].to(DEV), x2[:n].to(DEV))
    assert s.shape == (n,)
    # the SURVEY 8d criterion with the scale of the WHOLE 10k workload (the rms of one or two scores is not a scale):
    # |S - S_ref| <= 1e-4 * max(|S_ref|, rms(S_ref over configs[0]))
    full = O.nplda_score(x1, x2, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]).double()
    bound = 1e-4 * torch.maximum(ref.double().abs(), full.pow(2).mean().sqrt())
    err = (s.cpu().double() - ref.double()).abs()
    assert bool((err <= bound).all()), float((err / bound).max())


@pytest.mark.parametrize("n", [1, 63, 64, 65, 127, 128, 129, 4097, 9471, 9473, 10000])
def test_pair_kernel_ragged_sizes_identical_to_one_cta(kaldi_params, cfg1, n):
    """K1p (csrc/score_tcp.cu: two CTAs of a cluster share every M = 256 MMA, each converting its own 128 rows and holding
    half of the weight rows) on ragged batches -- one CTA of the pair without a single live row, partial tiles, an odd
    number of tiles: against the oracle (1e-4) and BIT-identical to the one-CTA bf16x3 kernel (same products, same
    accumulation order).  It is what IMPL_AUTO takes for materialised pairs from 9 472 pairs on."""
    x1, x2, _ = cfg1
    kp = kaldi_params
    m = make_nplda(kp, npl.IMPL_TC_PAIR)
    a, b = x1[:n].to(DEV), x2[:n].to(DEV)
    with torch.no_grad():
        s = m(a, b)
        m.impl = npl.IMPL_TC_BF16
        one = m(a, b)
        m.impl = npl.IMPL_AUTO
        auto = m(a, b)
    assert s.shape == (n,) and torch.equal(s, one) and torch.equal(s, auto)
    ref = O.nplda_score(x1[:n], x2[:n], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    full = O.nplda_score(x1, x2, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]).double()
    bound = 1e-4 * torch.maximum(ref.double().abs(), full.pow(2).mean().sqrt())
    err = (s.cpu().double() - ref.double()).abs()
    assert bool((err <= bound).all()), float((err / bound).max())


def test_pair_kernel_deterministic_any_range_and_dims(kaldi_params, cfg1):
    """Back-to-back launches of the pair kernel give the same bits (no races in the cross-CTA mbarrier protocol); inputs
    far outside fp16's range score like the fp32 kernel (bf16 halves keep fp32's exponent); NaN rows give NaN scores;
    a shape it does not take (fewer than two 32-wide stages) fails loudly when asked for explicitly."""
    x1, x2, _ = cfg1
    a, b = x1.to(DEV), x2.to(DEV)
    m = make_nplda(kaldi_params, npl.IMPL_TC_PAIR)
    with torch.no_grad():
        first = m(a, b)
        for _ in range(20):
            again = m(a, b)
        torch.cuda.synchronize()
        assert torch.equal(first, again)
        for scale in (1e-4, 3e4):
            s = m(a * scale, b * scale)
            m.impl = npl.IMPL_SIMT
            ref = m(a * scale, b * scale)
            m.impl = npl.IMPL_TC_PAIR
            ok, worst = parity_ok(s, ref.cpu(), rel=1e-4)
            assert ok, (scale, worst)
        bad = a.clone()
        bad[5, 17] = float("nan")
        s = m(bad, b)
        assert torch.isnan(s[5]) and torch.isfinite(s[:5]).all() and torch.isfinite(s[6:]).all()
    class C(NC):
        xvector_dim = 32
    m2 = npl.NeuralPlda(C).to(DEV)
    m2.impl = npl.IMPL_TC_PAIR
    with pytest.raises(RuntimeError):
        m2(torch.zeros(300, 32, device=DEV), torch.zeros(300, 32, device=DEV))


def test_pair_kernel_mixed_mode_parity_and_range_guard(ref_out, kaldi_params, cfg1):
    """IMPL_TC_PAIR_F8 (the pair kernel with layer 1 as fp16*fp16 + two e4m3*e4m3 products, four MMAs per K = 32 instead of
    six): 1e-4 parity on the golden 10k pairs; inputs outside the range the e4m3 terms cover (too small, too large, NaN)
    are recomputed on the device by the bf16x3 pair kernel behind it -- bit-identical to IMPL_TC_PAIR -- and a later
    in-range call takes the mixed path again (the guard slot was cleared)."""
    x1, x2, _ = cfg1
    a, b = x1.to(DEV), x2.to(DEV)
    m8 = make_nplda(kaldi_params, _lib.IMPL_TC_PAIR_F8)
    mp = make_nplda(kaldi_params, npl.IMPL_TC_PAIR)
    with torch.no_grad():
        s = m8(a, b)
        ok, worst = parity_ok(s, torch.from_numpy(ref_out["c1_scores"]), rel=1e-4)
        assert ok and worst <= 0.5, worst
        base = mp(a, b)
        assert not torch.equal(s, base)                       # the mixed path really ran
        for scale in (1e-3, 300.0):
            assert torch.equal(m8(a * scale, b * scale), mp(a * scale, b * scale))
        bad = a.clone()
        bad[7, 100] = float("nan")
        s_nan, p_nan = m8(bad, b), mp(bad, b)
        assert torch.isnan(s_nan[7]) and torch.equal(s_nan[:7], p_nan[:7]) and torch.equal(s_nan[8:], p_nan[8:])
        again = m8(a, b)
        assert torch.equal(again, s)


def test_tc_kernel_is_selected_and_deterministic(kaldi_params, cfg1):
    """IMPL_TC must run (no silent fallback) for the reference dims, agree with the SIMT kernel, and be
    bit-reproducible over back-to-back launches (the mbarrier pipelines have no data races)."""
    x1, x2, _ = cfg1
    a, b = x1.to(DEV), x2.to(DEV)
    m = make_nplda(kaldi_params, npl.IMPL_TC)
    with torch.no_grad():
        first = m(a, b)
        for _ in range(20):
            again = m(a, b)
        torch.cuda.synchronize()
        assert torch.equal(first, again)
        m.impl = npl.IMPL_SIMT
        simt = m(a, b)
    ok, worst = parity_ok(first, simt.cpu(), rel=1e-4)
    assert ok, worst
    class C(NC):
        xvector_dim = 96
    m2 = npl.NeuralPlda(C).to(DEV)
    m2.impl = npl.IMPL_TC
    with pytest.raises(RuntimeError):            # 96 % 32 == 0 is fine, but explicit TC on DPlda-only shapes must fail loudly
        class D(NC):
            xvector_dim = 100
        m3 = npl.NeuralPlda(D).to(DEV); m3.impl = npl.IMPL_TC
        m3(torch.zeros(4, 100, device=DEV), torch.zeros(4, 100, device=DEV))


def test_tc_f8_mode_parity_and_range_guard(ref_out, kaldi_params, cfg1):
    """IMPL_TC_F8 (layer 1 as fp16*fp16 + two e4m3*e4m3 products on one accumulator): 1e-4 parity on the golden
    10k config and on ragged sizes; inputs outside the range the e4m3 terms cover must be recomputed on the
    device by the bf16x3 kernel (bit-identical to IMPL_TC_BF16), and a later in-range call must again take the fp8 path
    (the guard slot is cleared)."""
    x1, x2, _ = cfg1
    a, b = x1.to(DEV), x2.to(DEV)
    m = make_nplda(kaldi_params, npl.IMPL_TC_F8)
    mt = make_nplda(kaldi_params, npl.IMPL_TC_BF16)
    with torch.no_grad():
        s = m(a, b)
        ok, worst = parity_ok(s, torch.from_numpy(ref_out["c1_scores"]), rel=1e-4)
        assert ok, worst
        assert not torch.equal(s, mt(a, b))                      # it really is the other arithmetic
        kp = kaldi_params
        for n in (1, 63, 65, 1000, 4097):
            ref = O.nplda_score(x1[:n], x2[:n], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
            np.testing.assert_allclose(m(a[:n], b[:n]).cpu().numpy(), ref.numpy(), rtol=2e-4, atol=2e-4)
        for scale in (1e-3, 300.0):                                # out of range -> guarded fallback pass
            got, want = m(a * scale, b * scale), mt(a * scale, b * scale)
            assert torch.equal(got, want), scale
            ref = O.nplda_score(x1 * scale, x2 * scale, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
            ok, worst = parity_ok(got, ref, rel=1e-4)
            assert ok, (scale, worst)
        one_bad = a.clone(); one_bad[777, 5] = 1.0e4                 # a single out-of-range element flags the call
        assert torch.equal(m(one_bad, b), mt(one_bad, b))
        again = m(a, b)                                            # guard cleared: fp8 path again, same bits as before
        assert torch.equal(again, s)


def test_tc_fp16x3_mode_parity_and_range_guard(ref_out, kaldi_params, cfg1):
    """IMPL_TC = "fp16x3" (both layers as split fp16, weights scaled into range at pack time): closer to the reference
    than bf16x3 on the golden 10k config (bound 1e-4; measured ~0.04 of it against 0.13), different bits from bf16x3;
    inputs outside fp16's range (|x| >= 2048 after the kernel's 2^4 scaling, inf) are recomputed on the device by the
    bf16x3 kernel -- bit-identical to IMPL_TC_BF16 -- and NaN inputs give NaN scores as in the reference; the guard slot
    is cleared afterwards.  IMPL_AUTO scores with the bf16x3 arithmetic (the fastest)."""
    x1, x2, _ = cfg1
    a, b = x1.to(DEV), x2.to(DEV)
    m16 = make_nplda(kaldi_params, npl.IMPL_TC)
    mb = make_nplda(kaldi_params, npl.IMPL_TC_BF16)
    ma = make_nplda(kaldi_params, npl.IMPL_AUTO)
    ref = torch.from_numpy(ref_out["c1_scores"])
    kp = kaldi_params
    with torch.no_grad():
        s16, sb = m16(a, b), mb(a, b)
        ok16, w16 = parity_ok(s16, ref, rel=1e-4)
        okb, wb = parity_ok(sb, ref, rel=1e-4)
        assert ok16 and okb and w16 < 0.6 * wb, (w16, wb)
        assert not torch.equal(s16, sb) and torch.equal(ma(a, b), sb)
        for n in (1, 63, 65, 1000, 4097):
            r = O.nplda_score(x1[:n], x2[:n], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
            np.testing.assert_allclose(m16(a[:n], b[:n]).cpu().numpy(), r.numpy(), rtol=1e-4, atol=1e-4)
        for scale in (1.0e3, 1.0e5):                               # |x| 2^4 beyond fp16 -> guarded bf16x3 pass
            got, want = m16(a * scale, b * scale), mb(a * scale, b * scale)
            assert torch.equal(got, want), scale
        small = m16(a * 1e-3, b * 1e-3)                            # small inputs stay on the fp16x3 path and stay accurate
        r = O.nplda_score(x1 * 1e-3, x2 * 1e-3, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
        ok, worst = parity_ok(small, r, rel=1e-4)
        assert ok, worst
        one_bad = a.clone(); one_bad[777, 5] = float("inf")
        assert torch.equal(m16(one_bad, b).nan_to_num(7.0), mb(one_bad, b).nan_to_num(7.0))
        nan_in = a.clone(); nan_in[12, 3] = float("nan")
        out = m16(nan_in, b)
        assert bool(torch.isnan(out[12])) and bool(torch.isfinite(out[:12]).all()) and bool(torch.isfinite(out[13:]).all())
        assert torch.equal(m16(a, b), s16)                         # guard cleared: the fp16x3 path again, same bits


def test_empty_input(kaldi_params):
    m = make_nplda(kaldi_params)
    with torch.no_grad():
        s = m(torch.zeros(0, 512, device=DEV), torch.zeros(0, 512, device=DEV))
    assert s.shape == (0,)


def test_cpu_tensors_raise(kaldi_params):
    m = make_nplda(kaldi_params)
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 512), torch.zeros(4, 512))
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 100, device=DEV), torch.zeros(4, 100, device=DEV))


@pytest.mark.parametrize("impl", IMPLS)
def test_default_init_and_random_params(ref_out, impl):
    z = np.load(os.path.join(GOLDEN, "default_init_params.npz"))
    m = npl.NeuralPlda(NC).to(DEV)
    m.impl = impl
    m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files})
    g0 = torch.Generator().manual_seed(0)
    z1, z2 = torch.randn(256, 512, generator=g0), torch.randn(256, 512, generator=g0)
    with torch.no_grad():
        s = m(z1.to(DEV), z2.to(DEV))
    ok, worst = parity_ok(s, torch.from_numpy(ref_out["c3_scores"]), rel=1e-4)
    assert ok, worst


@pytest.mark.parametrize("dims", [(512, 170, 170), (96, 40, 24), (257, 192, 192), (30, 7, 5), (64, 64, 100)])
def test_other_dims_simt(dims):
    d_in, d1, d2 = dims
    class C(NC):
        xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = d_in, d1, d2
    torch.manual_seed(3)
    m = npl.NeuralPlda(C)
    g = torch.Generator().manual_seed(9)
    x1, x2 = torch.randn(300, d_in, generator=g), torch.randn(300, d_in, generator=g)
    p = m.state_dict()
    ref = O.nplda_score(x1, x2, p["centering_and_LDA.weight"], p["centering_and_LDA.bias"],
                        p["centering_and_wccn_plda.weight"], p["centering_and_wccn_plda.bias"], p["P_sqrt"], p["Q"])
    m = m.to(DEV)
    with torch.no_grad():
        s = m(x1.to(DEV), x2.to(DEV))
    ok, worst = parity_ok(s, ref, rel=1e-4)
    assert ok, worst


def test_unsupported_dims_error():
    class C(NC):
        layer1_LDA_dim = 300
    m = npl.NeuralPlda(C).to(DEV)
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 512, device=DEV), torch.zeros(4, 512, device=DEV))


@pytest.mark.parametrize("impl", IMPLS + [npl.IMPL_TC])       # 768 pairs: AUTO stays on the fp32 kernel, TC = the fused kernel
def test_dplda_forward_golden(ref_out, kaldi_params, cfg1, impl):
    x1, x2, _ = cfg1
    m = make_dplda(kaldi_params, ref_out, impl)
    ref = torch.from_numpy(ref_out["c4_scores"])
    n = ref.numel()
    with torch.no_grad():
        s = m(x1[:n].to(DEV), x2[:n].to(DEV))
    ok, worst = parity_ok(s, ref, rel=1e-4)
    assert ok, worst


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 129, 1000, 4097])
def test_dplda_fused_forward_ragged_sizes(ref_out, kaldi_params, cfg1, n):
    """The one-pass tensor-core DPlda kernel (both square products per tile) on ragged sizes, forced with IMPL_TC, against
    the oracle's closed form; the criterion uses the score scale of the whole 10k workload."""
    x1, x2, _ = cfg1
    kp = kaldi_params
    m = make_dplda(kp, ref_out, npl.IMPL_TC)
    w, c = dplda_weights(ref_out)
    ref = O.dplda_score(x1[:n], x2[:n], kp["W1"], kp["b1"], w, c)
    full = O.dplda_score(x1[:2000], x2[:2000], kp["W1"], kp["b1"], w, c).double()
    with torch.no_grad():
        s = m(x1[:n].to(DEV), x2[:n].to(DEV))
    assert s.shape == (n,)
    bound = 1e-4 * torch.maximum(ref.double().abs(), full.pow(2).mean().sqrt())
    err = (s.cpu().double() - ref.double()).abs()
    assert bool((err <= bound).all()), float((err / bound).max())


def test_dplda_fused_forward_range_guard(ref_out, kaldi_params, cfg1):
    """Inputs outside fp16's range make the fused DPlda kernel's fp16x3 pass raise its guard; the bf16x3 pass behind it
    (same kernel, bf16 halves re-read from shared memory) recomputes the call: still within the bound of the oracle, and
    a later in-range call gives the same bits as before."""
    x1, x2, _ = cfg1
    kp = kaldi_params
    m = make_dplda(kp, ref_out, npl.IMPL_TC)
    w, c = dplda_weights(ref_out)
    n = 3000
    a, b = x1[:n].to(DEV), x2[:n].to(DEV)
    with torch.no_grad():
        s0 = m(a, b)
        big = m(a * 1000.0, b * 1000.0)
        s1 = m(a, b)
    ref = O.dplda_score(x1[:n] * 1000.0, x2[:n] * 1000.0, kp["W1"], kp["b1"], w, c)
    ok, worst = parity_ok(big, ref, rel=1e-4)
    assert ok, worst
    assert torch.equal(s0, s1)


@pytest.mark.parametrize("n", [64, 65, 256, 1000])
def test_dplda_frozen_lda_training_step(ref_out, kaldi_params, cfg1, n):
    """LDA frozen (xvector_DPlda_pytorch.py:140-147): the forward keeps only the normalised rows u and the backward is
    the gradient of logistic_regres alone (dplda_lr_bwd: Ps / Pd contractions).  Same loss and logistic_regres gradients
    as with a trainable LDA (golden .grad of the unmodified reference at n = 256), no gradient on the frozen parameters."""
    x1, x2, t = cfg1
    a, b, tt = x1[:n].to(DEV), x2[:n].to(DEV), t[:n].to(DEV)
    full = make_dplda(kaldi_params, ref_out)
    lf = full.loss(full(a, b), tt)
    lf.backward()
    m = make_dplda(kaldi_params, ref_out)
    for p in (m.centering_and_LDA.weight, m.centering_and_LDA.bias):
        p.requires_grad_(False)
    out = m(a, b)
    loss = m.loss(out, tt)
    loss.backward()
    assert m.centering_and_LDA.weight.grad is None and m.centering_and_LDA.bias.grad is None
    assert loss.item() == pytest.approx(lf.item(), rel=1e-5)
    for name in ("weight", "bias"):
        got, want = getattr(m.logistic_regres, name).grad, getattr(full.logistic_regres, name).grad
        scale = float(want.abs().max())
        assert float((got - want).abs().max()) <= 1e-4 * scale, name
    if n == 256:
        assert loss.item() == pytest.approx(float(ref_out["c4_train_loss"]), rel=1e-4)
        _check_grads(m, ref_out, "c4", ["logistic_regres.weight", "logistic_regres.bias"])


@pytest.mark.parametrize("impl", IMPLS)
def test_indexed_matches_materialised(kaldi_params, impl):
    table, i1, i2, _ = O.synth_grid(37, 53, 11, seed=1003, mean=kaldi_params["mean"])
    m = make_nplda(kaldi_params, impl)
    t = table.to(DEV)
    with torch.no_grad():
        s_idx, flag = m.forward_indexed(t, i1.to(DEV), i2.to(DEV), embed_once=False)   # gather fused into the score kernel
        s_mat = m(t[i1.to(DEV)], t[i2.to(DEV)])
    assert int(flag.item()) == 0
    if impl == npl.IMPL_SIMT:       # same kernel, same arithmetic: bit-identical
        np.testing.assert_array_equal(s_idx.cpu().numpy(), s_mat.cpu().numpy())
    else:                           # AUTO: materialised pairs take the tcgen05 kernel, the indexed gather the SIMT one
        ok, worst = parity_ok(s_idx, s_mat.cpu(), rel=1e-4)
        assert ok, worst
    kp = kaldi_params
    ref = O.nplda_score(table[i1], table[i2], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    ok, worst = parity_ok(s_idx, ref, rel=1e-4)
    assert ok, worst
    bad = i1.clone(); bad[5] = 10 ** 6
    with torch.no_grad():
        _, flag = m.forward_indexed(t, bad.to(DEV), i2.to(DEV), embed_once=False)
    assert int(flag.item()) != 0


@pytest.mark.parametrize("kind", ["nplda", "dplda"])
def test_embed_once_trial_list_scoring(ref_out, kaldi_params, kind):
    """every utterance of the table transformed once (nplda_table_prepare), trials scored as
    r[i] + r[j] + A[i].B[j] (nplda_score_pairs).  Against the oracle on the gathered pairs, both models; the
    cached row table must follow parameter updates and in-place table updates; bad indices are flagged."""
    kp = kaldi_params
    table, i1, i2, _ = O.synth_grid(61, 83, 17, seed=1004, mean=kp["mean"])
    if kind == "nplda":
        m = make_nplda(kp)
        oracle = lambda tb: O.nplda_score(tb[i1], tb[i2], *[m.state_dict()[k].cpu() for k in (
            "centering_and_LDA.weight", "centering_and_LDA.bias", "centering_and_wccn_plda.weight",
            "centering_and_wccn_plda.bias", "P_sqrt", "Q")])
    else:
        m = make_dplda(kp, ref_out)
        oracle = lambda tb: O.dplda_score(tb[i1], tb[i2], *[m.state_dict()[k].cpu() for k in (
            "centering_and_LDA.weight", "centering_and_LDA.bias", "logistic_regres.weight", "logistic_regres.bias")])
    t, a, b = table.to(DEV), i1.to(DEV), i2.to(DEV)
    s, flag = m.forward_indexed(t, a, b)                       # default: table has fewer rows than trials -> embed once
    assert int(flag.item()) == 0 and s.shape == (i1.numel(),)
    ok, worst = parity_ok(s, oracle(table), rel=1e-4)
    assert ok, worst
    s2, _ = m.forward_indexed(t, a[:100], b[:100])               # cached rows (validated on the device): same bits
    assert torch.equal(s2, s[:100])
    s_fused, _ = m.forward_indexed(t, a, b, embed_once=False)    # the per-trial kernel agrees
    ok, worst = parity_ok(s, s_fused.cpu(), rel=1e-4)
    assert ok, worst
    with torch.no_grad():                                        # parameter update -> rows rebuilt
        m.centering_and_LDA.bias.add_(0.05)
    s3, _ = m.forward_indexed(t, a, b)
    ok, worst = parity_ok(s3, oracle(table), rel=1e-4)
    assert ok, worst
    assert not torch.equal(s3, s)
    t.mul_(1.01)                                                 # in-place table update -> rows rebuilt
    s4, _ = m.forward_indexed(t, a, b)
    ok, worst = parity_ok(s4, oracle(t.cpu()), rel=1e-4)
    assert ok, worst
    bad = i2.clone(); bad[7] = -3
    _, flag = m.forward_indexed(t, a, bad.to(DEV))
    assert int(flag.item()) != 0
    s0, _ = m.forward_indexed(t, a[:0], b[:0])
    assert s0.shape == (0,)


def test_embedding_half_steps(ref_out, kaldi_params, cfg1):
    """extract_plda_embeddings / forward_from_plda_embeddings (models.py:366-376, 478-489)."""
    x1, x2, _ = cfg1
    n = 777
    kp = kaldi_params
    m = make_nplda(kp)
    with torch.no_grad():
        y1 = m.extract_plda_embeddings(x1[:n].to(DEV))
        y2 = m.extract_plda_embeddings(x2[:n].to(DEV))
        s = m.forward_from_plda_embeddings(y1, y2)
    ref_y = O.nplda_embed(x1[:n], kp["W1"], kp["b1"], kp["W2"], kp["b2"])
    assert y1.shape == (n, 170)
    np.testing.assert_allclose(y1.cpu().numpy(), ref_y.numpy(), rtol=1e-4, atol=1e-5)
    ok, worst = parity_ok(s, torch.from_numpy(ref_out["c1_scores"][:n]), rel=1e-4)
    assert ok, worst
    with pytest.raises(RuntimeError):
        m.extract_plda_embeddings(x1[:4].to(DEV))           # grad mode: not differentiable here
    d = make_dplda(kp, ref_out)
    nd = 300
    with torch.no_grad():
        u1 = d.extract_plda_embeddings(x1[:nd].to(DEV))
        u2 = d.extract_plda_embeddings(x2[:nd].to(DEV))
        sd = d.forward_from_plda_embeddings(u1, u2)
    np.testing.assert_allclose(u1.cpu().numpy(), O.dplda_embed(x1[:nd], kp["W1"], kp["b1"]).numpy(), rtol=1e-4, atol=1e-6)
    ok, worst = parity_ok(sd, torch.from_numpy(ref_out["c4_scores"][:nd]), rel=1e-4)
    assert ok, worst


def test_losses_match_reference(ref_out, kaldi_params, cfg1):
    _, _, t = cfg1
    s = torch.from_numpy(ref_out["c1_scores"]).to(DEV)
    t = t.to(DEV)
    m = make_nplda(kaldi_params)
    assert m.softcdet(s, t).item() == pytest.approx(float(ref_out["c1_softcdet_th0"]), rel=1e-4)
    assert m.cdet(s, t).item() == pytest.approx(float(ref_out["c1_cdet_th0"]), rel=1e-6)
    assert m.crossentropy(s, t).item() == pytest.approx(float(ref_out["c1_bce_th0"]), rel=1e-4)
    mc, th = m.minc(s[:2000], t[:2000], update_thresholds=True)
    assert float(mc) == pytest.approx(float(ref_out["c1_minc2k"]), rel=1e-6)
    np.testing.assert_array_equal(np.float32([float(th[b]) for b in NC.beta]), ref_out["c1_minc2k_th"].astype(np.float32))
    np.testing.assert_array_equal(np.float32([m.Th99.item(), m.Th199.item()]), ref_out["c1_th_state"].astype(np.float32))
    assert m.softcdet(s, t).item() == pytest.approx(float(ref_out["c1_softcdet_thminc"]), rel=1e-4)
    assert m.cdet(s, t).item() == pytest.approx(float(ref_out["c1_cdet_thminc"]), rel=1e-6)
    m.lossfn = "softCdet"          # the shipped voices_config.cfg spelling: loss() returns None (models.py:395-399)
    assert m.loss(s, t) is None


def test_accumulators_bitwise_layout(ref_out, cfg1):
    _, _, t = cfg1
    s = torch.from_numpy(ref_out["c1_scores"])
    th = torch.tensor([0.1, -0.3])
    acc = F_.loss_accumulators(s.to(DEV), t.to(DEV), th.to(DEV), 15.0, torch.tensor([0.25], device=DEV)).cpu()
    ref = O.loss_accumulators(s, t, th.tolist(), NC.beta, 15.0, 0.25)
    np.testing.assert_allclose(acc.numpy(), ref.numpy(), rtol=2e-6)
    assert acc[4 * 2 + 0].item() == t.sum().item() and acc[4 * 2 + 3].item() == 10000.0
    assert acc[2].item() == ((s < 0.1).float() * t).sum().item()          # hard counts are exact


def test_minc_quirks_gpu(ref_out, kaldi_params):
    m = make_nplda(kaldi_params)
    toy_s = torch.tensor([.1, .5, .9, -.2, .3, .7, -1.], device=DEV)
    toy_t = torch.tensor([1., 1., 1., 0., 0., 0., 0.], device=DEV)
    mc, th = m.minc(toy_s, toy_t)
    assert float(mc) == pytest.approx(float(ref_out["c5_minc"]), rel=1e-6)
    assert [float(th[b]) for b in NC.beta] == pytest.approx(list(ref_out["c5_th"]))
    g = torch.Generator().manual_seed(5)
    qs = torch.round(torch.randn(400, generator=g) * 4) / 4
    qt = (torch.rand(400, generator=g) < 0.3).float()
    mc, th = m.minc(qs.to(DEV), qt.to(DEV))
    assert float(mc) == pytest.approx(float(ref_out["c5b_minc"]), rel=1e-6)
    assert [float(th[b]) for b in NC.beta] == pytest.approx(list(ref_out["c5b_th"]))
    # large sweep against the O(N log N) oracle, exact
    g = torch.Generator().manual_seed(11)
    s = torch.randn(200000, generator=g)
    t = (torch.rand(200000, generator=g) < 0.05).float()
    s = s + 2.0 * t
    mc, th = m.minc(s.to(DEV), t.to(DEV))
    omc, oth = O.minc(s, t, NC.beta)
    assert float(mc) == pytest.approx(float(omc), rel=1e-6)
    assert [float(th[b]) for b in NC.beta] == [float(oth[b]) for b in NC.beta]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 1000, 2048, 2049, 4097, 100_003, 1 << 20, (1 << 20) + 5])
def test_sorted_populations_match_torch_sort(n):
    """csrc/sort.cu (label split + bitonic sort in its flip / disperse form, virtual +inf tail for lengths that are not
    a power of two) against `torch.sort(output[target > 0.5])` / `torch.sort(output[target < 0.5])` (models.py:407-408):
    bit-identical, labels of exactly 0.5 in neither population."""
    g = torch.Generator().manual_seed(n)
    s = torch.randn(n, generator=g)
    t = (torch.rand(n, generator=g) < 0.1).float()
    if n > 10:
        t[7] = 0.5
        s[3] = s[4]                                            # ties
    sd, td = s.to(DEV), t.to(DEV)
    tgt, non = F_.sorted_populations(sd, td)
    assert torch.equal(tgt.cpu(), torch.sort(s[t > 0.5])[0]) and torch.equal(non.cpu(), torch.sort(s[t < 0.5])[0])


def _check_grads(model, ref_out, prefix, names, rtol=1e-4):
    got = dict(model.named_parameters())
    for n in names:
        sample = ref_out[f"{prefix}_grad_{n}_sample"]
        p = got[n]
        if sample.size == 0:      # the reference's autograd leaves .grad None (Adam with weight decay then skips the parameter)
            assert p.grad is None, n
            continue
        assert p.grad is not None, n
        g = p.grad.reshape(-1).cpu()
        norm = float(ref_out[f"{prefix}_grad_{n}_norm"])
        assert float(g.double().norm()) == pytest.approx(norm, rel=rtol), n
        scale = norm / np.sqrt(g.numel()) + 1e-30
        np.testing.assert_allclose(g[::7].numpy(), sample, rtol=rtol, atol=rtol * scale, err_msg=n)


@pytest.mark.parametrize("lossname", ["SoftCdet", "crossentropy"])
def test_nplda_training_step_gradients(ref_out, kaldi_params, cfg1, lossname):
    """forward + loss + backward at the voices batch size against the reference's autograd."""
    x1, x2, t = cfg1
    m = make_nplda(kaldi_params, loss=lossname)
    with torch.no_grad():
        m.Th99.fill_(float(ref_out["c1_th_state"][0])); m.Th199.fill_(float(ref_out["c1_th_state"][1]))
        m.threshold_Xent.fill_(0.25)
    out = m(x1[:2048].to(DEV), x2[:2048].to(DEV))
    loss = m.loss(out, t[:2048].to(DEV))
    assert loss.item() == pytest.approx(float(ref_out[f"c2_{lossname}_loss"]), rel=1e-4)
    loss.backward()
    _check_grads(m, ref_out, f"c2_{lossname}", [str(n) for n in ref_out["param_names"]])


def test_dplda_training_step_gradients(ref_out, kaldi_params, cfg1):
    x1, x2, t = cfg1
    m = make_dplda(kaldi_params, ref_out)
    out = m(x1[:256].to(DEV), x2[:256].to(DEV))
    loss = m.loss(out, t[:256].to(DEV))
    assert loss.item() == pytest.approx(float(ref_out["c4_train_loss"]), rel=1e-4)
    loss.backward()
    _check_grads(m, ref_out, "c4", [str(n) for n in ref_out["c4_param_names"]])


def test_backward_linearity_large(kaldi_params, cfg1):
    """Size-independent property at a multi-chunk size: grads are linear in dS."""
    x1, x2, _ = cfg1
    reps = 56                                  # 560k pairs > one backward chunk (524288)
    X1 = x1.repeat(reps, 1).to(DEV); X2 = x2.repeat(reps, 1).to(DEV)
    m = make_nplda(kaldi_params)
    g = torch.Generator().manual_seed(4)
    w = torch.randn(10000, generator=g).to(DEV)
    m(X1, X2).mul(w.repeat(reps)).sum().backward()
    big = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    m(X1[:10000], X2[:10000]).mul(w).sum().backward()
    for n, p in m.named_parameters():
        if p.grad is None:
            continue
        scale = float(p.grad.abs().max()) * reps + 1e-30
        assert float((big[n] - reps * p.grad).abs().max()) <= 2e-4 * scale, n


def test_optimizer_step_repacks_weights(kaldi_params, cfg1):
    x1, x2, t = cfg1
    m = make_nplda(kaldi_params, loss="crossentropy")
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    a, b, tt = x1[:512].to(DEV), x2[:512].to(DEV), t[:512].to(DEV)
    s0 = m(a, b).detach().clone()
    loss = m.loss(m(a, b), tt); loss.backward(); opt.step()
    s1 = m(a, b).detach()
    assert float((s1 - s0).abs().max()) > 0
    p = m.state_dict()
    ref = O.nplda_score(a.cpu(), b.cpu(), *(p[k].cpu() for k in (
        "centering_and_LDA.weight", "centering_and_LDA.bias", "centering_and_wccn_plda.weight",
        "centering_and_wccn_plda.bias", "P_sqrt", "Q")))
    ok, worst = parity_ok(s1, ref, rel=1e-4)
    assert ok, worst


def test_pickle_roundtrip(tmp_path, kaldi_params, cfg1):
    x1, x2, _ = cfg1
    m = make_nplda(kaldi_params)
    a, b = x1[:100].to(DEV), x2[:100].to(DEV)
    with torch.no_grad():
        s0 = m(a, b)
    f = tmp_path / "m.pt"
    m.SaveModel(str(f))
    m2 = pickle.load(open(f, "rb"))
    assert m2.threshold[99.0] is m2.Th99
    with torch.no_grad():
        s1 = m2(a, b)
    np.testing.assert_array_equal(s0.cpu().numpy(), s1.cpu().numpy())


def test_host_entry_matches_device_entry(kaldi_params, cfg1):
    x1, x2, _ = cfg1
    n = 5000
    m = make_nplda(kaldi_params)
    with torch.no_grad():
        s_dev = m(x1[:n].to(DEV), x2[:n].to(DEV)).cpu()
    pack = m.packed.get("nplda", m._params(), 512, 170, 170)
    h1, h2 = x1[:n].contiguous().pin_memory(), x2[:n].contiguous().pin_memory()
    out = torch.empty(n).pin_memory()
    chunk = 1536
    nbytes = _lib.lib().nplda_host_scratch_bytes(chunk, 512)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    torch.cuda.synchronize()
    rc = _lib.lib().nplda_score_fwd_host(ctypes.c_void_p(h1.data_ptr()), ctypes.c_void_p(h2.data_ptr()), n, 512, 170,
                                         170, _lib.ptr(pack), ctypes.c_void_p(out.data_ptr()), chunk,
                                         _lib.ptr(scratch), nbytes, 0, _lib.IMPL_AUTO)
    assert rc == 0
    np.testing.assert_array_equal(out.numpy(), s_dev.numpy())


def test_scorefile_generation_matches_reference_files(tmp_path, kaldi_params):
    from neuralplda_b200 import scorefile_generator as sg
    z = np.load(os.path.join(GOLDEN, "c6_mega.npz"))
    mega = {str(u): z["vecs"][i] for i, u in enumerate(z["ids"])}
    m = make_nplda(kaldi_params)
    for fn, trials, golden in ((sg.generate_voices_scores, "c6_voices_trials.txt", "c6_voices_scores.txt"),
                               (sg.generate_sre_scores, "c6_sre_trials.tsv", "c6_sre_scores.tsv")):
        out = tmp_path / golden
        fn(str(out), os.path.join(GOLDEN, trials), mega, m, torch.device(DEV), batch_size=10)
        got = open(out).read().strip().split("\n")
        ref = open(os.path.join(GOLDEN, golden)).read().strip().split("\n")
        assert len(got) == len(ref)
        for lg, lr in zip(got, ref):
            cg, cr = lg.split("\t"), lr.split("\t")
            assert cg[:-1] == cr[:-1]
            if cr[-1] != "LLR":
                assert float(cg[-1]) == pytest.approx(float(cr[-1]), rel=1e-4, abs=1e-5)


def test_large_properties_1m(kaldi_params):
    """BASELINE.json configs[1] size: 1M pairs.  EVERY pair against the CPU oracle ("the CPU oracle checks
    all pairs for <= 1 M"), plus size-independent properties: pair symmetry S(x1,x2)=S(x2,x1) and chunking invariance."""
    mean = kaldi_params["mean"].to(DEV)
    g = torch.Generator(device=DEV).manual_seed(1002)
    n = 1_000_000
    spk = torch.randn(2000, 512, generator=g, device=DEV)
    s1 = torch.randint(0, 2000, (n,), generator=g, device=DEV)
    s2 = torch.where(torch.rand(n, generator=g, device=DEV) < 0.1, s1, torch.randint(0, 2000, (n,), generator=g, device=DEV))
    x1 = mean + spk[s1] + 0.7 * torch.randn(n, 512, generator=g, device=DEV)
    x2 = mean + spk[s2] + 0.7 * torch.randn(n, 512, generator=g, device=DEV)
    kp = kaldi_params
    ref = torch.cat([O.nplda_score(x1[c:c + 100_000].cpu(), x2[c:c + 100_000].cpu(), kp["W1"], kp["b1"], kp["W2"], kp["b2"],
                                   kp["P_sqrt"], kp["Q"]) for c in range(0, n, 100_000)])
    for impl in IMPLS:
        m = make_nplda(kp, impl)
        with torch.no_grad():
            s = m(x1, x2)
            s_sw = m(x2, x1)
            s_part = torch.cat([m(x1[a:a + 333_333], x2[a:a + 333_333]) for a in range(0, n, 333_333)])
        assert torch.isfinite(s).all()
        np.testing.assert_allclose(s.cpu().numpy(), s_sw.cpu().numpy(), rtol=1e-4, atol=1e-4)
        ok, worst = parity_ok(s_part, s.cpu(), rel=1e-4)
        assert ok, worst
        ok, worst = parity_ok(s, ref, rel=1e-4)                     # all 1,000,000 pairs
        assert ok, worst


@pytest.mark.parametrize("kind", ["nplda", "dplda"])
def test_dense_trial_list_subgrid_gather(ref_out, kaldi_params, kind, monkeypatch):
    """Dense trial lists go through a sub-grid product over the rows the list uses + a 4-byte gather per trial
    (nplda_trial_rows / nplda_score_grid / nplda_trial_grid_gather); sparse ones through one row gather per trial
    (nplda_score_pairs).  Same scores either way (both within the bound of the oracle), bad rows reported and scored 0."""
    kp = kaldi_params
    table, i1, i2, _ = O.synth_grid(300, 700, 60, seed=77, mean=kp["mean"])
    g = torch.Generator().manual_seed(5)
    keep = torch.rand(i1.numel(), generator=g) < 0.4                     # 40 % of a 300 x 700 grid, in list order
    a, b = i1[keep].clone(), i2[keep].clone()
    b[1234] = 10_000                                                     # a row outside the table
    m = make_nplda(kp) if kind == "nplda" else make_dplda(kp, ref_out)
    t = table.to(DEV)
    monkeypatch.setattr(F_, "GRID_GATHER_MIN_TRIALS", 1 << 40)
    s_pairs, f_pairs = m.forward_indexed(t, a.to(DEV), b.to(DEV), embed_once=True)
    monkeypatch.setattr(F_, "GRID_GATHER_MIN_TRIALS", 1)
    launches = _lib.lib().nplda_launch_count()
    s_grid, f_grid = m.forward_indexed(t, a.to(DEV), b.to(DEV), embed_once=True)
    assert int(f_pairs.item()) == 1 and int(f_grid.item()) == 1
    assert float(s_grid[1234]) == 0.0 and float(s_pairs[1234]) == 0.0
    ok = torch.ones(a.numel(), dtype=torch.bool); ok[1234] = False
    if kind == "nplda":
        ref = O.nplda_score(table[a[ok]], table[b[ok]], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    else:
        w, c = dplda_weights(ref_out)
        ref = O.dplda_score(table[a[ok]], table[b[ok]], kp["W1"], kp["b1"], w, c)
    for s in (s_pairs, s_grid):
        good, worst = parity_ok(s.cpu()[ok], ref, rel=1e-4)
        assert good, worst
    # a sparse list (one trial per enrol row) stays on the per-trial kernel: the sub-grid would have 300 cells per trial
    monkeypatch.setattr(F_, "GRID_GATHER_MIN_TRIALS", 1)
    sp_a, sp_b = torch.arange(300), 300 + torch.arange(300)
    s_sp, _ = m.forward_indexed(t, sp_a.to(DEV), sp_b.to(DEV), embed_once=True)
    np.testing.assert_allclose(s_sp.cpu().numpy(), m.forward_grid(t, sp_a.to(DEV), sp_b.to(DEV))[0].diagonal().cpu().numpy(),
                               rtol=1e-5, atol=1e-5)
    assert _lib.lib().nplda_launch_count() > launches


def test_config3_10m_grid_trial_list(kaldi_params):
    """BASELINE.json configs[2]: 10M trials = 2500 enrol x 4000 test grid over 6500 x-vectors, indexed layout.
    Oracle on a strided subsample; properties: the grid is symmetric under swapping the roles of the two index
    lists, and scoring the list in ragged chunks gives the same bits."""
    kp = kaldi_params
    table, i1, i2, lab = O.synth_grid(2500, 4000, 500, seed=1003, mean=kp["mean"])
    m = make_nplda(kp)
    t, a, b = table.to(DEV), i1.to(DEV), i2.to(DEV)
    s, flag = m.forward_indexed(t, a, b)
    assert int(flag.item()) == 0 and s.numel() == 10_000_000
    sub = torch.arange(0, s.numel(), 1009)
    ref = O.nplda_score(table[i1[sub]], table[i2[sub]], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    ok, worst = parity_ok(s[sub.to(DEV)], ref, rel=1e-4)
    assert ok, worst
    # S(i, j) == S(j, i) (models.py:373-375 is symmetric); the list is dense, so it is scored as a sub-grid on the tensor
    # cores and swapping the lists swaps the roles of the fp16 hi/lo operands: equal up to that rounding (a tenth of the bound)
    s_swapped, _ = m.forward_indexed(t, b, a)
    np.testing.assert_allclose(s_swapped.cpu().numpy(), s.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # ragged chunks: the one-trial chunk takes the per-trial fp32 kernel, the large ones their own sub-grids
    parts = [m.forward_indexed(t, a[lo:hi], b[lo:hi])[0] for lo, hi in ((0, 1), (1, 3_333_333), (3_333_333, 10_000_000))]
    np.testing.assert_allclose(torch.cat(parts).cpu().numpy(), s.cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert torch.equal(torch.cat(parts[1:]), s[1:])                 # the same grid cell gives the same bits in any sub-grid
    # targets score higher than non-targets on this generator: a degenerate kernel would not separate them
    labd = lab.to(DEV).bool()
    assert float(s[labd].mean()) > float(s[~labd].mean()) + 0.5


def test_config4_50m_sharded_accumulators(kaldi_params):
    """BASELINE.json configs[3]: 50M trials = 5000 x 10000 grid, sharded by contiguous trial ranges as the
    multi-GPU path shards them (dist.shard_range).  The raw fp64 loss accumulators of the shards must add up to
    the accumulators of the whole list (the all-reduce payload), and the loss finalised from the sum must equal
    the loss of the whole list -- while the mean of per-shard losses does not."""
    from neuralplda_b200 import dist as D
    kp = kaldi_params
    table, i1, i2, lab = O.synth_grid(5000, 10000, 700, seed=1004, mean=kp["mean"])
    m = make_nplda(kp)
    t, a, b, y = table.to(DEV), i1.to(DEV), i2.to(DEV), lab.to(DEV)
    n = a.numel()
    assert n == 50_000_000
    s, _ = m.forward_indexed(t, a, b)
    th = torch.cat([m.threshold[bt].detach() for bt in NC.beta])
    whole = F_.loss_accumulators(s, y, th, NC.alpha, m.threshold_Xent)
    world = 8
    shards = []
    for r in range(world):
        lo, hi = D.shard_range(n, world, r)
        sr, _ = m.forward_indexed(t, a[lo:hi], b[lo:hi])
        assert torch.equal(sr, s[lo:hi])
        shards.append(F_.loss_accumulators(sr, y[lo:hi], th, NC.alpha, m.threshold_Xent))
    summed = torch.stack(shards).sum(0)
    np.testing.assert_allclose(summed.cpu().numpy(), whole.cpu().numpy(), rtol=1e-12)
    assert float(summed[-1]) == n
    loss_whole = F_.finalize(whole, NC.beta)[0].item()
    loss_sum = F_.finalize(summed, NC.beta)[0].item()
    assert loss_sum == pytest.approx(loss_whole, rel=1e-6)
    per_shard = np.mean([F_.finalize(sh, NC.beta)[0].item() for sh in shards])
    assert abs(per_shard - loss_whole) > 1e-7 * abs(loss_whole)     # averaging per-rank losses is NOT the same thing
    sub = torch.arange(0, n, 50021)
    ref = O.nplda_score(table[i1[sub]], table[i2[sub]], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    ok, worst = parity_ok(s[sub.to(DEV)], ref, rel=1e-4)
    assert ok, worst


# ------------------------------------------------------------------------------
# f-4: cohort score normalisation (utils/adaptive_score_normalization.py)
# ------------------------------------------------------------------------------


@pytest.mark.parametrize("m,c,top_n", [(7, 510, 500), (3, 40, 500), (33, 4097, 500), (5, 1000, 1), (2, 3000, 2999)])
def test_cohort_stats_vs_oracle(m, c, top_n):
    """Row statistics incl. the N lowest scores (radix select == sort()[:N]); ties, negative zero, duplicates of
    the boundary value.  fp64 sums in a different order than numpy's: 1e-12 relative."""
    from neuralplda_b200 import adaptive_score_normalization as asn
    g = torch.Generator().manual_seed(100 + c)
    x = torch.randn(m, c, generator=g) * 0.3 - 0.8
    x[:, 5] = x[:, 17]
    x[0, : min(c, 300)] = x[0, 0]                                # a long run of equal values across the boundary
    x[-1, 1::2] = -0.0
    x[-1, 0::2] = 0.0
    got = asn.cohort_statistics(x.to(DEV), top_n).cpu().numpy()
    want = O.cohort_stats(x.numpy(), top_n)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)


def test_score_norm_golden_files(tmp_path):
    """The four normalised score files of the unmodified reference script (tests/golden/c7_*): values within
    1e-6 (raw scores travel as float32 here, the script parses the text as float64), same rows and header."""
    import shutil
    from neuralplda_b200 import adaptive_score_normalization as asn
    raw = str(tmp_path / "raw.tsv")
    shutil.copy(os.path.join(GOLDEN, "c7_raw_scores.tsv"), raw)
    out = asn.normalize_score_file(raw, os.path.join(GOLDEN, "c7_cohort_scores.tsv"), DEV)
    for k, name in enumerate(asn.NORMS):
        want = open(os.path.join(GOLDEN, "c7_raw_scores.tsv_%s.tsv" % name)).read().splitlines()
        got = open(raw + "_%s.tsv" % name).read().splitlines()
        assert len(got) == len(want) and got[0] == want[0]
        for lg, lw in zip(got[1:], want[1:]):
            fg, fw = lg.split("\t"), lw.split("\t")
            assert fg[:-1] == fw[:-1]
            assert abs(float(fg[-1]) - float(fw[-1])) <= 1e-6 * max(1.0, abs(float(fw[-1])))
        np.testing.assert_allclose(out[k], [float(l.split("\t")[-1]) for l in want[1:]], rtol=1e-6, atol=1e-6)


def test_score_norm_unknown_row_raises():
    from neuralplda_b200 import adaptive_score_normalization as asn
    stats = asn.cohort_statistics(torch.randn(4, 64, device=DEV), 10)
    raw = torch.randn(5, device=DEV)
    with pytest.raises(KeyError):
        asn.normalize_scores(raw, torch.tensor([0, 1, 2, 3, 4], device=DEV), torch.zeros(5, dtype=torch.int64, device=DEV), stats)


def test_cohort_grid_scoring_feeds_normalisation(kaldi_params):
    """End to end on the device: id x cohort grid through the embed-once kernel -> statistics -> normalised
    trial scores, against the oracle fed with the oracle's own scores (1e-3: score parity 1e-4 divided by std)."""
    from neuralplda_b200 import adaptive_score_normalization as asn
    kp = kaldi_params
    table, _, _, _ = O.synth_grid(30, 620, 25, seed=11, mean=kp["mean"])      # rows 0..29 ids, 30..649 cohort
    m = make_nplda(kp)
    ids, coh = torch.arange(30), torch.arange(30, 650)
    S = asn.score_cohort(m, table.to(DEV), ids.to(DEV), coh.to(DEV))
    with torch.no_grad():
        y = O.nplda_embed(table, kp["W1"], kp["b1"], kp["W2"], kp["b2"])
        Sref = O.nplda_score_from_embeddings(y[ids].repeat_interleave(620, 0), y[coh].repeat(30, 1), kp["P_sqrt"], kp["Q"]).view(30, 620)
    ok, worst = parity_ok(S.cpu().flatten(), Sref.flatten(), rel=1e-4)
    assert ok, worst
    e, t = torch.arange(0, 15).repeat(4), torch.arange(15, 30).repeat_interleave(4)
    raw, _ = m.forward_indexed(table.to(DEV), e.to(DEV), t.to(DEV))
    got = asn.normalize_scores(raw, e.to(DEV), t.to(DEV), asn.cohort_statistics(S, 500)).cpu().numpy()
    want = O.score_norm(raw.cpu().numpy(), e.numpy(), t.numpy(), O.cohort_stats(Sref.numpy(), 500))
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-3)


# ------------------------------------------------------------------------------
# Grid scoring (nplda_score_grid): enrol x test grids without an index pair per trial
# ------------------------------------------------------------------------------


@pytest.mark.parametrize("kind", ["nplda", "dplda"])
@pytest.mark.parametrize("ne,nt", [(1, 1), (61, 83), (128, 128), (129, 257), (300, 7)])
def test_grid_scoring_vs_oracle(ref_out, kaldi_params, kind, ne, nt):
    """Every enrol row against every test row, ragged tile tails and odd leading dimensions (scalar-store path),
    both models, against the oracle on the gathered pairs (1e-4) and against the trial-list kernel."""
    kp = kaldi_params
    table, i1, i2, _ = O.synth_grid(ne, nt, 17, seed=2000 + ne, mean=kp["mean"])
    if kind == "nplda":
        m = make_nplda(kp)
        ref = O.nplda_score(table[i1], table[i2], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    else:
        m = make_dplda(kp, ref_out)
        w, c = dplda_weights(ref_out)
        ref = O.dplda_score(table[i1], table[i2], kp["W1"], kp["b1"], w, c)
    t = table.to(DEV)
    er, tr = torch.arange(ne, device=DEV), torch.arange(ne, ne + nt, device=DEV)
    assert torch.equal(i1.view(ne, nt)[:, 0], er.cpu()) and torch.equal(i2.view(ne, nt)[0], tr.cpu())
    s, flag = m.forward_grid(t, er, tr)
    assert s.shape == (ne, nt) and int(flag.item()) == 0
    ok, worst = parity_ok(s.flatten(), ref, rel=1e-4)
    assert ok, worst
    s_list, _ = m.forward_indexed(t, i1.to(DEV), i2.to(DEV), embed_once=True)
    ok, worst = parity_ok(s.flatten(), s_list.cpu(), rel=2e-5)         # same rows, 170-term fp32 sums in another order
    assert ok, worst
    # permuted / repeated rows select the same scores
    perm_e, perm_t = torch.randperm(ne, device=DEV), torch.randint(0, nt, (nt + 3,), device=DEV)
    s2, _ = m.forward_grid(t, er[perm_e], tr[perm_t])
    assert torch.equal(s2, s[perm_e][:, perm_t])


def test_grid_scoring_bad_rows_and_empty(kaldi_params):
    kp = kaldi_params
    table, _, _, _ = O.synth_grid(20, 30, 5, seed=9, mean=kp["mean"])
    m = make_nplda(kp)
    t = table.to(DEV)
    er, tr = torch.arange(20, device=DEV), torch.arange(20, 50, device=DEV)
    good, _ = m.forward_grid(t, er, tr)
    er_bad = er.clone(); er_bad[3] = 50                                   # one past the table
    tr_bad = tr.clone(); tr_bad[7] = -1
    s, flag = m.forward_grid(t, er_bad, tr_bad)
    assert int(flag.item()) != 0
    assert torch.all(s[3] == 0) and torch.all(s[:, 7] == 0)
    keep_e, keep_t = [i for i in range(20) if i != 3], [j for j in range(30) if j != 7]
    assert torch.equal(s[keep_e][:, keep_t], good[keep_e][:, keep_t])
    s0, _ = m.forward_grid(t, er[:0], tr)
    assert s0.shape == (0, 30)
    with pytest.raises(RuntimeError):
        m.forward_grid(table, er, tr)                                      # CPU table: no fallback


def test_config3_10m_grid_kernel(kaldi_params):
    """BASELINE.json configs[2] through the grid kernel: 2500 x 4000 = 10M trials; oracle on a strided subsample,
    agreement with the trial-list kernel on every trial, symmetry S(i,j) = S(j,i) via the transposed grid."""
    kp = kaldi_params
    table, i1, i2, lab = O.synth_grid(2500, 4000, 500, seed=1003, mean=kp["mean"])
    m = make_nplda(kp)
    t = table.to(DEV)
    er, tr = torch.arange(2500, device=DEV), torch.arange(2500, 6500, device=DEV)
    s, flag = m.forward_grid(t, er, tr)
    assert int(flag.item()) == 0 and s.shape == (2500, 4000)
    sub = torch.arange(0, s.numel(), 1009)
    ref = O.nplda_score(table[i1[sub]], table[i2[sub]], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    ok, worst = parity_ok(s.flatten()[sub.to(DEV)], ref, rel=1e-4)
    assert ok, worst
    s_list, _ = m.forward_indexed(t, i1.to(DEV), i2.to(DEV))
    ok, worst = parity_ok(s.flatten(), s_list.cpu(), rel=2e-5)
    assert ok, worst
    st, _ = m.forward_grid(t, tr, er)              # roles of A' and B' swapped: equal up to the fp16x3 rounding (a tenth of the bound)
    np.testing.assert_allclose(st.t().cpu().numpy(), s.cpu().numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("batch", [512, 4608])
@pytest.mark.parametrize("lossname", ["SoftCdet", "crossentropy"])
def test_training_trajectory_matches_oracle(kaldi_params, lossname, batch):
    """The reference's training-loop body (xvector_NeuralPlda_pytorch.py:35-43) for 12 Adam steps on fresh batches
    (512 pairs: all-fp32 backward; 4608 pairs: activations saved by the tensor-core forward, tile phases around the
    tensor-core dL/du pass): forward -> loss -> backward -> optimizer.step through the drop-in module on the GPU vs
    the oracle port under torch autograd on the CPU, same parameter order.  Loss values within 1e-4 relative at
    every step, parameters within 1e-4 of the update scale at the end."""
    kp = kaldi_params
    m = make_nplda(kp, loss=lossname)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    W = {k: torch.nn.Parameter(kp[k].clone()) for k in ("W1", "b1", "W2", "b2", "P_sqrt", "Q")}
    th = [torch.nn.Parameter(torch.zeros(1)) for _ in NC.beta]
    thx = torch.nn.Parameter(torch.zeros(1))
    copt = torch.optim.Adam([W["P_sqrt"], W["Q"]] + th + [thx, W["W1"], W["b1"], W["W2"], W["b2"]], lr=1e-4)
    for step in range(12):
        x1, x2, t = O.synth_pairs(batch, 40, seed=300 + step, mean=kp["mean"])
        opt.zero_grad()
        loss = m.loss(m(x1.to(DEV), x2.to(DEV)), t.to(DEV))
        loss.backward()
        opt.step()
        copt.zero_grad()
        s = O.nplda_score(x1, x2, W["W1"], W["b1"], W["W2"], W["b2"], W["P_sqrt"], W["Q"])
        closs = O.softcdet(s, t, torch.cat(th), NC.beta, NC.alpha) if lossname == "SoftCdet" else O.crossentropy(s, t, thx)
        closs.backward()
        copt.step()
        assert abs(loss.item() - closs.item()) <= 1e-4 * max(abs(closs.item()), 1e-2), (step, loss.item(), closs.item())
    sd = m.state_dict()
    for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_wccn_plda.weight", "W2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
        moved = (W[key].detach() - kp[key]).abs().max().item()              # ~ 12 * lr
        diff = (sd[name].cpu() - W[key].detach()).abs().max().item()
        assert moved > 1e-4 and diff <= 0.05 * moved, (name, moved, diff)



@pytest.mark.parametrize("how", ["data_add", "data_copy", "fused_adam", "state_dict_data_copy"])
def test_parameter_changes_without_version_bump_are_seen(kaldi_params, how):
    """The reference writes parameters through `.data.copy_()` (models.py:449-457; minc :420) and users may train with
    fused optimisers: neither bumps tensor._version.  Every path (materialised forward, embed-once trial list, grid)
    must score with the CURRENT parameters afterwards; cached per-utterance rows are validated on the device."""
    kp = kaldi_params
    table, i1, i2, _ = O.synth_grid(40, 50, 9, seed=77, mean=kp["mean"])
    m = make_nplda(kp, loss="crossentropy")
    t, a, b = table.to(DEV), i1.to(DEV), i2.to(DEV)
    er, tr = torch.arange(40, device=DEV), torch.arange(40, 90, device=DEV)

    def all_paths():
        with torch.no_grad():
            return (m(t[a], t[b]), m.forward_indexed(t, a, b, embed_once=True)[0], m.forward_grid(t, er, tr)[0].flatten())

    def oracle():
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        return O.nplda_score(table[i1], table[i2], sd["centering_and_LDA.weight"], sd["centering_and_LDA.bias"],
                             sd["centering_and_wccn_plda.weight"], sd["centering_and_wccn_plda.bias"], sd["P_sqrt"], sd["Q"])

    before = all_paths()
    for s in before:
        ok, worst = parity_ok(s, oracle(), rel=1e-4)
        assert ok, worst
    p = m.centering_and_wccn_plda.bias
    v0 = p._version
    if how == "data_add":
        p.data.add_(0.05)
    elif how == "data_copy":
        p.data.copy_(p.detach() + 0.05)
    elif how == "state_dict_data_copy":
        m.state_dict()["centering_and_wccn_plda.bias"].data.copy_(p.detach() + 0.05)
    else:
        opt = torch.optim.Adam(m.parameters(), lr=1e-2, fused=True)
        x1, x2, lab = O.synth_pairs(512, 40, seed=5, mean=kp["mean"])
        m.loss(m(x1.to(DEV), x2.to(DEV)), lab.to(DEV)).backward()
        opt.step()
    if how != "fused_adam":
        assert p._version == v0                                   # the premise: nothing on the host saw the change
    after = all_paths()
    ref = oracle()
    for s0, s1 in zip(before, after):
        assert not torch.equal(s0, s1)
        ok, worst = parity_ok(s1, ref, rel=1e-4)
        assert ok, worst
    again = all_paths()                                           # unchanged parameters: cached rows reused, same bits
    for s1, s2 in zip(after, again):
        assert torch.equal(s1, s2)


def test_mixed_impl_without_and_with_mixed_image(kaldi_params, cfg1):
    """NPLDA_IMPL_TC_F8 selected on the module packs the mixed image; calling the C entry with impl F8 on a pack built
    WITHOUT it must fall back to the bf16x3 kernel on the device (hdr[2] == 0), never read a stale image."""
    x1, x2, _ = cfg1
    kp = kaldi_params
    m = make_nplda(kp, npl.IMPL_TC_F8)
    a, b = x1[:4096].to(DEV), x2[:4096].to(DEV)
    ref = O.nplda_score(x1[:4096], x2[:4096], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    with torch.no_grad():
        s = m(a, b)
    ok, worst = parity_ok(s, ref, rel=1e-4)
    assert ok, worst
    pack = m.packed.get("nplda", m._params(), 512, 170, 170, mixed=False)      # re-pack without the mixed image
    out = torch.empty(4096, device=DEV)
    _lib.check(_lib.lib().nplda_score_fwd(_lib.ptr(a), _lib.ptr(b), 4096, 512, 170, 170, _lib.ptr(pack), _lib.ptr(out),
                                          npl.IMPL_TC_F8, _lib.stream_ptr()), "nplda_score_fwd")
    m.impl = npl.IMPL_TC_BF16
    with torch.no_grad():
        s_tc = m(a, b)
    assert torch.equal(out, s_tc)                                # the fallback pass is the bf16x3 kernel
    # the same for the CTA-pair kernel's mixed mode: stale / absent mixed pair image -> every tile flagged -> bf16x3 pair pass
    _lib.check(_lib.lib().nplda_score_fwd(_lib.ptr(a), _lib.ptr(b), 4096, 512, 170, 170, _lib.ptr(pack), _lib.ptr(out),
                                          _lib.IMPL_TC_PAIR_F8, _lib.stream_ptr()), "nplda_score_fwd")
    assert torch.equal(out, s_tc)


def test_loader_gather_bit_exact(ref_out):
    """load_xvec_trials_from_numbatch / _from_idbatch (sv_trials_loaders.py:418-437) on the golden dict: the device
    gather returns exactly the rows the reference's Python gather returns (checksums of the unmodified reference in
    reference_outputs.npz; element-wise against the oracle's gather), for index tensors on the CPU and on the GPU."""
    from neuralplda_b200 import sv_trials_loaders as L
    z = np.load(os.path.join(GOLDEN, "c6_mega.npz"))
    ids = [str(s) for s in z["ids"]]
    mega = {u: z["vecs"][i] for i, u in enumerate(ids)}
    num_to_id = dict(enumerate(ids))
    d1, d2 = torch.from_numpy(ref_out["c6_d1"]), torch.from_numpy(ref_out["c6_d2"])
    R1, R2 = O.gather_numbatch(mega, num_to_id, d1, d2)
    for a, b in ((d1, d2), (d1.to(DEV), d2.to(DEV)), (d1.to(DEV), d2.to(DEV))):     # 2nd CUDA call: sync-free fast path
        X1, X2 = L.load_xvec_trials_from_numbatch(mega, num_to_id, a, b, torch.device(DEV))
        assert X1.is_cuda and X1.dtype == torch.float32 and X1.is_contiguous()
        assert torch.equal(X1.cpu(), R1) and torch.equal(X2.cpu(), R2)
        assert X1.double().sum().item() == pytest.approx(ref_out["c6_x1_sum"][0], rel=1e-12)
    trials = np.asarray([["/some/dir/" + ids[int(i)] + ".wav", ids[int(j)] + ".sph"] for i, j in zip(d1[:7], d2[:7])])
    Y1, Y2 = L.load_xvec_trials_from_idbatch(mega, trials, torch.device(DEV))
    assert torch.equal(Y1.cpu(), R1[:7]) and torch.equal(Y2.cpu(), R2[:7])
    # unknown rows: CPU indices raise at once (the reference's dict lookup raises KeyError); GPU indices are not read
    # back -- the gather kernel fills the row with NaN and raises a pinned flag that the next loader call (or
    # check_pending_errors(), e.g. after the last batch of a loop) reports
    with pytest.raises(KeyError):
        L.load_xvec_trials_from_numbatch(mega, num_to_id, torch.tensor([len(ids)]), torch.tensor([0]), torch.device(DEV))
    bad = torch.tensor([0, len(ids) + 3], device=DEV)
    B1, B2 = L.load_xvec_trials_from_numbatch(mega, num_to_id, bad, bad, torch.device(DEV))
    torch.cuda.synchronize()
    assert torch.equal(B1[0].cpu(), torch.from_numpy(z["vecs"][0]).float()) and bool(torch.isnan(B1[1]).all())
    with pytest.raises(KeyError):
        L.load_xvec_trials_from_numbatch(mega, num_to_id, d1.to(DEV), d2.to(DEV), torch.device(DEV))
    L.load_xvec_trials_from_numbatch(mega, num_to_id, bad, bad, torch.device(DEV))      # "last batch of an epoch"
    with pytest.raises(KeyError):
        L.check_pending_errors()
    L.check_pending_errors()                                                            # reported once
    X1, _ = L.load_xvec_trials_from_numbatch(mega, num_to_id, d1.to(DEV), d2.to(DEV), torch.device(DEV))   # flag cleared
    assert torch.equal(X1.cpu(), R1)


@pytest.mark.parametrize("kind", ["nplda", "dplda"])
def test_tensor_core_weight_gradients(ref_out, kaldi_params, cfg1, kind):
    """The dW contractions (dW1 = DA^T X, dW2 = DY^T U; DPlda dWw / dWb) on the tcgen05 bf16x3 kernel
    (csrc/gemm_tc.cu, taken for batches >= 8192 rows) against the fp32 SIMT contraction on the same batch: every
    parameter gradient within 1e-4 of its largest entry; ragged row counts exercise the zero-filled tails."""
    x1, x2, t = cfg1
    n = 9000 + 37
    a, b, y = x1[:n].to(DEV), x2[:n].to(DEV), t[:n].to(DEV)
    grads = {}
    for mode in ("simt", "tc"):
        bwd_paths(gemm=mode)
        try:
            m = make_nplda(kaldi_params, loss="crossentropy") if kind == "nplda" else make_dplda(kaldi_params, ref_out)
            m.loss(m(a, b), y).backward()
            torch.cuda.synchronize()
            grads[mode] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        finally:
            bwd_paths(gemm=None)
    assert set(grads["tc"]) == set(grads["simt"]) and len(grads["tc"]) >= 4
    for k, gs in grads["simt"].items():
        scale = float(gs.abs().max()) + 1e-30
        assert float((grads["tc"][k] - gs).abs().max()) <= 1e-4 * scale, k


@pytest.mark.parametrize("lossname", ["SoftCdet", "crossentropy"])
def test_tensor_core_weight_gradients_vs_reference_autograd(ref_out, kaldi_params, cfg1, lossname, monkeypatch):
    """Same golden check as test_nplda_training_step_gradients (the unmodified reference's .grad on 2048 pairs), with
    the tensor-core contraction forced for this small batch."""
    bwd_paths(gemm="tc")
    test_nplda_training_step_gradients(ref_out, kaldi_params, cfg1, lossname)


def test_config5_dplda_training_step_shard(ref_out, kaldi_params):
    """BASELINE.json configs[4]: DPlda, 10M trials over 8 GPUs = 1.25M pairs per rank, one forward + BCE + backward
    with the LDA frozen (xvector_DPlda_pytorch.py:140-147).  One rank's shard at full size: the loss and the
    gradients of the whole shard equal the pair-count-weighted sums over two unequal parts (what the all-reduce of
    raw sums relies on), frozen parameters get no gradient, and the forward agrees with the oracle on a subsample."""
    kp = kaldi_params
    n = 1_250_000
    g = torch.Generator(device=DEV).manual_seed(1005)
    mean = kp["mean"].to(DEV)
    spk = torch.randn(2000, 512, generator=g, device=DEV)
    s1 = torch.randint(0, 2000, (n,), generator=g, device=DEV)
    same = torch.rand(n, generator=g, device=DEV) < 0.1
    s2 = torch.where(same, s1, torch.randint(0, 2000, (n,), generator=g, device=DEV))
    x1 = mean + spk[s1] + 0.7 * torch.randn(n, 512, generator=g, device=DEV)
    x2 = mean + spk[s2] + 0.7 * torch.randn(n, 512, generator=g, device=DEV)
    t = (s1 == s2).float()
    m = make_dplda(kp, ref_out)
    for p in (m.centering_and_LDA.weight, m.centering_and_LDA.bias):
        p.requires_grad_(False)

    def step(lo, hi):
        m.zero_grad(set_to_none=True)
        out = m(x1[lo:hi], x2[lo:hi])
        loss = m.loss(out, t[lo:hi])
        loss.backward()
        return loss.item(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}, out.detach()

    L, G, out = step(0, n)
    assert "centering_and_LDA.weight" not in G and "logistic_regres.weight" in G
    cut = 400_003
    L1, G1, _ = step(0, cut)
    L2, G2, _ = step(cut, n)
    assert L == pytest.approx((cut * L1 + (n - cut) * L2) / n, rel=1e-5)
    for k in G:
        want = (cut * G1[k] + (n - cut) * G2[k]) / n
        scale = float(want.abs().max()) + 1e-30
        assert float((G[k] - want).abs().max()) <= 2e-4 * scale, k
    idx = torch.arange(0, n, 1223, device=DEV)
    w, c = dplda_weights(ref_out)
    ref = O.dplda_score(x1[idx].cpu(), x2[idx].cpu(), kp["W1"], kp["b1"], w, c)
    ok, worst = parity_ok(out[idx], ref, rel=1e-4)
    assert ok, worst


@pytest.mark.parametrize("kind", ["nplda", "dplda"])
def test_backward_with_emitted_activations(ref_out, kaldi_params, cfg1, kind):
    """Backward fed by the tensor cores -- a = W1 x + b1 and y (DPlda: R u and Pm u) either saved by the training
    forward (default for batches >= 4096 pairs) or emitted again by the tcgen05 kernel inside the backward
    (`packed.save_activations = False`) -- against the all-fp32 backward that recomputes them in the tile kernel:
    all parameter gradients and the input gradients within 1e-4 of their largest entry, scores identical in kind,
    on a ragged batch spanning a partial tile."""
    x1, x2, t = cfg1
    n = 9000 + 37
    y = t[:n].to(DEV)
    grads = {}
    for variant, save, emit in (("fp32", False, "0"), ("emit_in_backward", False, "1"), ("saved_by_forward", True, None)):
        if emit is not None:
            bwd_paths(emit=emit)
        bwd_paths(gemm="simt")
        try:
            m = make_nplda(kaldi_params, loss="SoftCdet") if kind == "nplda" else make_dplda(kaldi_params, ref_out)
            m.packed.save_activations = save
            a, b = x1[:n].to(DEV).requires_grad_(True), x2[:n].to(DEV).requires_grad_(True)
            out = m(a, b)
            m.loss(out, y).backward()
            torch.cuda.synchronize()
            grads[variant] = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
            grads[variant]["x1"], grads[variant]["x2"], grads[variant]["scores"] = a.grad.clone(), b.grad.clone(), out.detach().clone()
        finally:
            bwd_paths(emit=None)
            bwd_paths(gemm=None)
    for variant in ("emit_in_backward", "saved_by_forward"):
        for k, g0 in grads["fp32"].items():
            scale = float(g0.abs().max()) + 1e-30
            assert float((grads[variant][k] - g0).abs().max()) <= 1e-4 * scale, (variant, k)


def test_saved_activations_survive_retain_graph(kaldi_params, cfg1):
    """The backward only READS the activations the training forward saved: a second backward through the same graph
    (retain_graph=True) gives the same gradients."""
    x1, x2, t = cfg1
    n = 5000
    m = make_nplda(kaldi_params, loss="crossentropy")
    loss = m.loss(m(x1[:n].to(DEV), x2[:n].to(DEV)), t[:n].to(DEV))
    loss.backward(retain_graph=True)
    g1 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    loss.backward()
    for k, p in m.named_parameters():
        if p.grad is not None:
            scale = float(g1[k].abs().max()) + 1e-30
            assert float((p.grad - g1[k]).abs().max()) <= 1e-6 * scale, k       # atomics reorder the fp32 sums


@pytest.mark.parametrize("lossname", ["SoftCdet", "crossentropy"])
def test_emitted_activations_vs_reference_autograd(ref_out, kaldi_params, cfg1, lossname, monkeypatch):
    """The golden .grad of the unmodified reference (2048 pairs) with both tensor-core pieces of the backward forced."""
    bwd_paths(emit="1")
    bwd_paths(gemm="tc")
    test_nplda_training_step_gradients(ref_out, kaldi_params, cfg1, lossname)


@pytest.mark.parametrize("d_in", [32, 64, 192, 1024])
def test_tc_kernel_other_input_widths(d_in):
    """The tcgen05 kernel takes any input width that is a multiple of 32 (1 .. 32 layer-1 stages per tile; the backward
    runs it with 192-wide rows for dL/du = dL/dy . W2): against the oracle on random parameters, ragged batch."""
    class C(NC):
        xvector_dim = d_in
    torch.manual_seed(d_in)
    m = npl.NeuralPlda(C).to(DEV)
    m.impl = npl.IMPL_TC
    g = torch.Generator().manual_seed(d_in + 1)
    x1, x2 = torch.randn(3000 + 13, d_in, generator=g), torch.randn(3000 + 13, d_in, generator=g)
    with torch.no_grad():
        s = m(x1.to(DEV), x2.to(DEV))
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = O.nplda_score(x1, x2, sd["centering_and_LDA.weight"], sd["centering_and_LDA.bias"],
                        sd["centering_and_wccn_plda.weight"], sd["centering_and_wccn_plda.bias"], sd["P_sqrt"], sd["Q"])
    ok, worst = parity_ok(s, ref, rel=1e-4)
    assert ok, worst


@pytest.mark.parametrize("lossname", ["SoftCdet", "crossentropy"])
def test_saved_activations_vs_reference_autograd(ref_out, kaldi_params, cfg1, lossname, monkeypatch):
    """The golden .grad of the unmodified reference (2048 pairs) through the default large-batch training path:
    activations saved by the tensor-core forward, tensor-core dL/du pass and weight gradients."""
    monkeypatch.setattr(F_, "SAVE_ACTIVATIONS_MIN_PAIRS", 1)
    bwd_paths(gemm="tc")
    launches0 = _lib.launch_count()
    test_nplda_training_step_gradients(ref_out, kaldi_params, cfg1, lossname)
    assert _lib.launch_count() > launches0


def test_dplda_saved_activations_vs_reference_autograd(ref_out, kaldi_params, cfg1, monkeypatch):
    monkeypatch.setattr(F_, "SAVE_ACTIVATIONS_MIN_PAIRS", 1)
    bwd_paths(gemm="tc")
    test_dplda_training_step_gradients(ref_out, kaldi_params, cfg1)


@pytest.mark.parametrize("lossname", ["crossentropy", "SoftCdet"])
def test_graphed_train_step_matches_eager(kaldi_params, lossname):
    """neuralplda_b200.graphs.GraphedTrainStep (gather + forward + loss + backward + Adam captured in one CUDA graph)
    against the same loop run eagerly: same losses step by step, same parameters after 10 steps (fp32 atomics reorder
    sums: 1e-5 of the update), including a short last batch that falls back to the eager body."""
    from neuralplda_b200.graphs import GraphedTrainStep
    from neuralplda_b200.sv_trials_loaders import load_xvec_trials_from_numbatch
    kp = kaldi_params
    table, i1, i2, lab = O.synth_grid(60, 90, 6, seed=21, mean=kp["mean"])
    mega = {"u%04d" % k: table[k].numpy() for k in range(table.shape[0])}
    n2i = dict(enumerate(mega))
    B = 128
    g = torch.Generator().manual_seed(3)
    batches = [torch.randperm(i1.numel(), generator=g)[:B] for _ in range(10)] + [torch.randperm(i1.numel(), generator=g)[:50]]
    dev = torch.device(DEV)

    m_e = make_nplda(kp, loss=lossname)
    opt_e = torch.optim.Adam(m_e.parameters(), lr=1e-3, capturable=True)
    eager = []
    for b in batches:
        opt_e.zero_grad()
        x1, x2 = load_xvec_trials_from_numbatch(mega, n2i, i1[b].to(dev), i2[b].to(dev), dev)
        loss = m_e.loss(m_e(x1, x2), lab[b].to(dev))
        loss.backward()
        opt_e.step()
        eager.append(loss.item())

    m_g = make_nplda(kp, loss=lossname)
    opt_g = torch.optim.Adam(m_g.parameters(), lr=1e-3, capturable=True)
    step = GraphedTrainStep(m_g, opt_g, mega, n2i, batch_size=B)
    graphed = [float(step(i1[b], i2[b], lab[b])) for b in batches]
    assert step.graph is not None
    np.testing.assert_allclose(graphed, eager, rtol=2e-4, atol=1e-6)
    scale = len(batches) * 1e-3                                   # Adam moves an entry by ~lr per step
    for (k, pe), (_, pg) in zip(m_e.named_parameters(), m_g.named_parameters()):
        assert float((pe.detach() - pg.detach()).abs().max()) <= 2e-3 * scale, k
    with pytest.raises(RuntimeError):
        GraphedTrainStep(m_g, torch.optim.Adam(m_g.parameters(), lr=1e-3), mega, n2i, batch_size=B)


@pytest.mark.parametrize("kind", ["nplda", "dplda"])
@pytest.mark.parametrize("n", [1, 63, 64, 65, 130, 1000])
def test_backward_ragged_small_batches(ref_out, kaldi_params, cfg1, kind, n):
    """The default training path at the reference's batch sizes and around the 64-pair tile boundary (below it the
    all-fp32 tile kernel does everything; from it on the tensor-core forward saves the activations) against the forced
    all-fp32 path: gradients within 1e-4 of their largest entry, loss within 1e-5."""
    x1, x2, t = cfg1
    a, b, y = x1[:n].to(DEV), x2[:n].to(DEV), t[:n].clone().to(DEV)
    y[0] = 1.0                                                    # at least one target, or softCdet is 0/0 as in the reference
    res = {}
    for variant in ("default", "fp32"):
        if variant == "fp32":
            bwd_paths(emit="0"); bwd_paths(gemm="simt")
        try:
            m = make_nplda(kaldi_params, loss="crossentropy") if kind == "nplda" else make_dplda(kaldi_params, ref_out)
            m.packed.save_activations = variant == "default"
            loss = m.loss(m(a, b), y)
            loss.backward()
            torch.cuda.synchronize()
            res[variant] = (loss.item(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
        finally:
            bwd_paths(emit=None); bwd_paths(gemm=None)
    assert res["default"][0] == pytest.approx(res["fp32"][0], rel=1e-5, abs=1e-7)
    for k, g0 in res["fp32"][1].items():
        scale = float(g0.abs().max()) + 1e-30
        assert float((res["default"][1][k] - g0).abs().max()) <= 1e-4 * scale, (k, n)


def test_sharded_training_two_ranks(tmp_path):
    """BASELINE.json configs[4] in miniature (tools/dist_train_check.py): two ranks, each forward + BCE + backward on its
    contiguous trial range with the raw loss accumulators all-reduced (model.process_group) and the parameter gradients
    summed (dist.allreduce_gradients), against the whole batch in one process: same loss, gradients within 1e-4 -- for
    NeuralPlda and DPlda.  (The BCE backward once divided by the rank's trial count instead of the global one.)
    Runs under torchrun with the gloo backend so that both ranks can share this box's GPU."""
    import socket, subprocess, sys
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, DIST_BACKEND="gloo")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(root, "tools", "dist_train_check.py")],
                       cwd=root, env=env, capture_output=True, text=True, timeout=420)
    lines = [l for l in r.stdout.splitlines() if l.startswith(("nplda:", "dplda:"))]
    assert r.returncode == 0 and len(lines) == 2 and all(l.endswith("OK") for l in lines), (r.stdout[-2000:], r.stderr[-2000:])
