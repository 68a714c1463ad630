"""K1x (csrc/score_tcx.cu): trial scoring from a PRE-SPLIT x-vector table on CTA-pair tcgen05 MMAs, against the CPU
oracle (models.py:378-382 on the gathered rows, sv_trials_loaders.py:418-426) and against the materialised-pair kernel
K1; table-identity caching (a second table of the same shape must never be scored with the first
table's rows)."""
import gc

import numpy as np
import pytest
import torch

import neuralplda_b200 as npl
from neuralplda_b200 import _lib, functional as F_
from oracle import nplda_oracle as O
from conftest import NC, parity_ok
from test_gpu_parity import make_nplda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _table(kp, rows, seed, nspk=64):
    g = torch.Generator().manual_seed(seed)
    spk = torch.randn(nspk, 512, generator=g)
    return kp["mean"] + spk[torch.randint(0, nspk, (rows,), generator=g)] + 0.7 * torch.randn(rows, 512, generator=g)


def _oracle(kp, table, i1, i2):
    return O.nplda_score(table[i1], table[i2], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 127, 128, 129, 255, 257, 1000, 9473, 60_001])
def test_split_kernel_ragged_sizes(kaldi_params, n):
    kp = kaldi_params
    table = _table(kp, 2999, 21)
    g = torch.Generator().manual_seed(n)
    i1, i2 = torch.randint(0, 2999, (n,), generator=g), torch.randint(0, 2999, (n,), generator=g)
    m = make_nplda(kp)
    t = table.to(DEV)
    s, flag = m.forward_indexed(t, i1.to(DEV), i2.to(DEV), embed_once=False, use_split=True)
    assert s.shape == (n,) and int(flag) == 0
    ref = _oracle(kp, table, i1, i2)
    np.testing.assert_allclose(s.cpu().numpy(), ref.numpy(), rtol=2e-4, atol=2e-4)
    if n >= 64:
        ok, worst = parity_ok(s, ref, rel=1e-4)
        assert ok, worst
    # the same fp16x3 arithmetic as the materialised-pair kernel under IMPL_TC up to the power-of-two input scales
    # (which move the subnormal threshold of the lo halves): within a tenth of the parity bound of each other
    with torch.no_grad():
        m.impl = npl.IMPL_TC
        k1 = m(t[i1.to(DEV)], t[i2.to(DEV)])
    scale = torch.maximum(ref.abs(), ref.pow(2).mean().sqrt() if n >= 64 else torch.tensor(0.5))
    assert float(((k1.cpu() - s.cpu()).abs() / scale).max()) <= 1e-5


def test_split_kernel_default_route_and_sass_path(kaldi_params):
    """forward_indexed takes the split kernel by default when the table is large against the batch (training-sized
    batches over the whole corpus) -- no fp32 SIMT fallback for indexed NeuralPlda trials."""
    kp = kaldi_params
    table = _table(kp, 5000, 3)
    t = table.to(DEV)
    m = make_nplda(kp)
    g = torch.Generator().manual_seed(9)
    i1, i2 = torch.randint(0, 5000, (2048,), generator=g), torch.randint(0, 5000, (2048,), generator=g)
    before = len(F_._split_cache)
    s, _ = m.forward_indexed(t, i1.to(DEV), i2.to(DEV))             # 5000 rows > 2 * 2048 trials: not embed-once
    assert len(F_._split_cache) == before + 1 and id(t) in F_._split_cache
    s2, _ = m.forward_indexed(t, i1.to(DEV), i2.to(DEV), embed_once=False, use_split=True)
    assert torch.equal(s, s2)
    ok, worst = parity_ok(s, _oracle(kp, table, i1, i2), rel=1e-4)
    assert ok, worst
    m.impl = npl.IMPL_SIMT                                          # explicit fp32 request is honoured
    s3, _ = m.forward_indexed(t, i1.to(DEV), i2.to(DEV), embed_once=False)
    assert not torch.equal(s3, s)
    ok, worst = parity_ok(s3, s.cpu(), rel=1e-4)
    assert ok, worst


@pytest.mark.parametrize("dims", [(64, 170, 170), (256, 150, 120), (1024, 176, 176), (512, 8, 33)])
def test_split_kernel_other_shapes(dims):
    d_in, d1, d2 = dims

    class C(NC):
        xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = d_in, d1, d2
    torch.manual_seed(7)
    m = npl.NeuralPlda(C).to(DEV)
    g = torch.Generator().manual_seed(d_in + d1)
    table = torch.randn(777, d_in, generator=g)
    i1, i2 = torch.randint(0, 777, (3001,), generator=g), torch.randint(0, 777, (3001,), generator=g)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ref = O.nplda_score(table[i1], table[i2], sd["centering_and_LDA.weight"], sd["centering_and_LDA.bias"],
                        sd["centering_and_wccn_plda.weight"], sd["centering_and_wccn_plda.bias"], sd["P_sqrt"], sd["Q"])
    s, flag = m.forward_indexed(table.to(DEV), i1.to(DEV), i2.to(DEV), embed_once=False, use_split=True)
    assert int(flag) == 0
    ok, worst = parity_ok(s, ref, rel=1e-4)
    assert ok, worst


def test_split_kernel_bad_indices_never_fault(kaldi_params):
    kp = kaldi_params
    t = _table(kp, 300, 5).to(DEV)
    m = make_nplda(kp)
    i1 = torch.tensor([0, 300, 5, -1, 2 ** 40, 7] * 50, device=DEV)
    i2 = torch.zeros_like(i1)
    s, flag = m.forward_indexed(t, i1, i2, embed_once=False, use_split=True)
    torch.cuda.synchronize()
    assert int(flag) != 0
    good = torch.tensor([0, 5, 7], device=DEV)
    s_ok, flag_ok = m.forward_indexed(t, good.repeat(50), torch.zeros(150, dtype=torch.int64, device=DEV), embed_once=False, use_split=True)
    assert int(flag_ok) == 0
    assert torch.equal(s_ok[:3], s[[0, 2, 5]])


def test_second_table_of_same_shape_is_never_scored_with_stale_rows(kaldi_params):
    """the per-table caches (row table of the embed-once path, split image of the tensor-core path)
    were keyed on data_ptr; a freed table whose address is recycled by another table of the same shape was scored with
    the old rows.  Tables are now identified by the tensor object (held while cached); temporaries are never cached."""
    kp = kaldi_params
    m = make_nplda(kp)
    g = torch.Generator().manual_seed(1)
    i1, i2 = torch.randint(0, 400, (1500,), generator=g), torch.randint(0, 400, (1500,), generator=g)
    a, b = i1.to(DEV), i2.to(DEV)
    for use_split in (False, True):
        kw = dict(embed_once=False, use_split=True) if use_split else {}
        tab_a = _table(kp, 400, 100)
        ta = tab_a.to(DEV)
        ptr_a = ta.data_ptr()
        sa, _ = m.forward_indexed(ta, a, b, **kw)
        ok, worst = parity_ok(sa, _oracle(kp, tab_a, i1, i2), rel=1e-4)
        assert ok, worst
        del ta
        gc.collect()
        tab_b = _table(kp, 400, 200)
        tb = tab_b.to(DEV)                                   # the caching allocator hands the same block back
        sb, _ = m.forward_indexed(tb, a, b, **kw)
        ok, worst = parity_ok(sb, _oracle(kp, tab_b, i1, i2), rel=1e-4)
        assert ok, (worst, "same address" if tb.data_ptr() == ptr_a else "different address")
        # non-fp32 and non-contiguous tables go through a temporary copy every call: correct every time
        for conv in (lambda t: t.double(), lambda t: torch.cat([t, t], 1)[:, :512]):
            for tab in (tab_a, tab_b):
                s, _ = m.forward_indexed(conv(tab.to(DEV)), a, b, **kw)
                ok, worst = parity_ok(s, _oracle(kp, tab, i1, i2), rel=1e-4)
                assert ok, worst
        # in-place update through tensor ops (version counter) and parameter updates are both followed
        tb.mul_(1.02)
        s, _ = m.forward_indexed(tb, a, b, **kw)
        ok, worst = parity_ok(s, _oracle(kp, tb.cpu(), i1, i2), rel=1e-4)
        assert ok, worst
        with torch.no_grad():
            m.Q.mul_(1.1)
        s, _ = m.forward_indexed(tb, a, b, **kw)
        kq = dict(kp); kq["Q"] = m.Q.detach().cpu()
        ok, worst = parity_ok(s, _oracle(kq, tb.cpu(), i1, i2), rel=1e-4)
        assert ok, worst
        with torch.no_grad():
            m.Q.copy_(kp["Q"])


def test_split_kernel_1m_trials_full_oracle_check(kaldi_params):
    """BASELINE configs[1] size through the indexed layout: 1 M trials over 20 000 utterances, EVERY trial against the
    CPU oracle (the oracle embeds each utterance once: same arithmetic per row as models.py:366-370, then
    forward_from_plda_embeddings on the gathered embeddings, models.py:372-376)."""
    kp = kaldi_params
    table = _table(kp, 20_000, 77, nspk=500)
    g = torch.Generator().manual_seed(2)
    n = 1_000_000
    i1, i2 = torch.randint(0, 20_000, (n,), generator=g), torch.randint(0, 20_000, (n,), generator=g)
    y = O.nplda_embed(table, kp["W1"], kp["b1"], kp["W2"], kp["b2"])
    ref = torch.cat([O.nplda_score_from_embeddings(y[i1[c:c + 100_000]], y[i2[c:c + 100_000]], kp["P_sqrt"], kp["Q"])
                     for c in range(0, n, 100_000)])
    m = make_nplda(kp)
    s, flag = m.forward_indexed(table.to(DEV), i1.to(DEV), i2.to(DEV), embed_once=False, use_split=True)
    assert int(flag) == 0
    ok, worst = parity_ok(s, ref, rel=1e-4)
    assert ok, worst
    for _ in range(3):                                          # deterministic across launches (no races in the pipelines)
        s2, _ = m.forward_indexed(table.to(DEV), i1.to(DEV), i2.to(DEV), embed_once=False, use_split=True)
        assert torch.equal(s, s2)


@pytest.mark.parametrize("kind", ["nplda", "dplda"])
@pytest.mark.parametrize("shape", [(1, 1), (37, 53), (128, 128), (129, 257), (300, 1001), (1000, 131)])
def test_grid_tensor_core_kernel_vs_fp32_and_oracle(ref_out, kaldi_params, kind, shape):
    """K5-TC (csrc/grid_tc.cu): enrol x test grids on tcgen05 with r[i] + r[j] folded into the contraction, against the
    fp32 FFMA2 grid kernel and the CPU oracle on the gathered pairs; ragged tiles, repeated and permuted rows, a
    strided output is not needed (ld = T), bad indices flagged."""
    from test_gpu_parity import make_dplda
    kp = kaldi_params
    E, T = shape
    table = _table(kp, 700, 31 + E)
    g = torch.Generator().manual_seed(E * 7 + T)
    er, tr = torch.randint(0, 700, (E,), generator=g), torch.randint(0, 700, (T,), generator=g)
    if kind == "nplda":
        m = make_nplda(kp)
        oracle = lambda a, b: O.nplda_score(a, b, kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    else:
        m = make_dplda(kp, ref_out)
        sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
        oracle = lambda a, b: O.dplda_score(a, b, sd["centering_and_LDA.weight"], sd["centering_and_LDA.bias"],
                                            sd["logistic_regres.weight"], sd["logistic_regres.bias"])
    t = table.to(DEV)
    s_tc, flag = m.forward_grid(t, er.to(DEV), tr.to(DEV))
    assert s_tc.shape == (E, T) and int(flag) == 0
    m.impl = npl.IMPL_SIMT
    s_f32, _ = m.forward_grid(t, er.to(DEV), tr.to(DEV))
    m.impl = npl.IMPL_AUTO
    i1, i2 = er.repeat_interleave(T), tr.repeat(E)
    ref = oracle(table[i1], table[i2]).reshape(E, T)
    scale = torch.maximum(ref.abs(), ref.pow(2).mean().sqrt() if E * T >= 64 else torch.tensor(0.5)).double()
    for name, s in (("tc", s_tc), ("fp32", s_f32)):
        worst = float(((s.cpu().double() - ref.double()).abs() / (1e-4 * scale)).max())
        assert worst <= 1.0, (name, worst)
    assert float(((s_tc - s_f32).abs().cpu().double() / scale).max()) <= 2e-5
    bad = er.clone(); bad[0] = 700
    _, flag = m.forward_grid(t, bad.to(DEV), tr.to(DEV))
    assert int(flag) != 0
    s_again, flag = m.forward_grid(t, er.to(DEV), tr.to(DEV))      # cached rows + cached grid operands: same bits
    assert int(flag) == 0 and torch.equal(s_again, s_tc)
    with torch.no_grad():                                          # a parameter update rebuilds rows AND grid operands
        p = m.Q if kind == "nplda" else m.logistic_regres.bias
        p.add_(0.01)
    s_new, _ = m.forward_grid(t, er.to(DEV), tr.to(DEV))
    assert not torch.equal(s_new, s_tc)
    m.impl = npl.IMPL_SIMT
    s_new32, _ = m.forward_grid(t, er.to(DEV), tr.to(DEV))
    assert float(((s_new - s_new32).abs().cpu().double() / scale).max()) <= 2e-5
