"""Host-side mirror of the reference interface: module construction, parameter names, pickling,
loss dispatch, loaders, Kaldi readers, trial-list sharding and the multi-rank accumulator reduce
(gloo, world_size 2).  No GPU needed."""
import os
import pickle

import numpy as np
import pytest
import torch

import neuralplda_b200 as npl
from neuralplda_b200 import kaldi_io_lite, sv_trials_loaders as L, dist as D
from oracle import nplda_oracle as O
from conftest import GOLDEN, NC, NCD


def test_parameter_names_order_and_shapes(ref_out):
    m = npl.NeuralPlda(NC)
    assert [n for n, _ in m.named_parameters()] == [str(s) for s in ref_out["param_names"]]
    assert sum(p.numel() for p in m.parameters()) == 116623
    d = npl.DPlda(NCD)
    assert [n for n, _ in d.named_parameters()] == [str(s) for s in ref_out["c4_param_names"]]
    assert d.logistic_regres.weight.shape == (1, 57970)
    assert m.threshold[99.0] is m.Th99 and m.threshold[199.0] is m.Th199
    class C(NC):
        beta = [9.899999999999999]
    assert "Th9" in dict(npl.NeuralPlda(C).named_parameters())        # Th{int(beta)} naming (models.py:358)


def test_state_dict_loads_reference_checkpoint_keys():
    z = np.load(os.path.join(GOLDEN, "default_init_params.npz"))       # saved from the reference module
    m = npl.NeuralPlda(NC)
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(z[k]) for k in z.files}, strict=True)
    assert not missing and not unexpected


def test_pickle_roundtrip_keeps_threshold_aliasing_and_drops_handles(tmp_path):
    m = npl.NeuralPlda(NC)
    _ = m.packed                                   # create the transient workspace object
    m.impl = npl.IMPL_SIMT
    f = tmp_path / "m.pt"
    m.SaveModel(str(f))
    m2 = pickle.load(open(f, "rb"))
    assert "_packed" not in m2.__dict__ and "_impl" not in m2.__dict__
    assert m2.threshold[99.0] is m2.Th99
    assert m2.impl == npl.IMPL_AUTO
    m2.Th99.data.fill_(0.5)
    assert float(m2.threshold[99.0]) == 0.5


def test_cpu_tensors_raise_no_fallback():
    m = npl.NeuralPlda(NC)
    with pytest.raises(RuntimeError, match="GPU only"):
        m(torch.zeros(3, 512), torch.zeros(3, 512))
    with pytest.raises(RuntimeError, match="GPU only"):
        m.softcdet(torch.zeros(3), torch.zeros(3))
    with pytest.raises(RuntimeError):
        L.load_xvec_trials_from_numbatch({}, {}, torch.zeros(1), torch.zeros(1), torch.device("cpu"))


def test_loss_dispatch_is_case_sensitive():
    m = npl.NeuralPlda(NC)
    m.lossfn = "softCdet"                          # conf/voices_config.cfg:25 spelling
    assert m.loss(torch.zeros(2), torch.zeros(2)) is None


def test_kaldi_readers_text_and_binary(tmp_path, kaldi_params):
    v = tmp_path / "v.txt"
    v.write_text(" [ 1.5 -2 3e-1 ]\n")
    np.testing.assert_allclose(kaldi_io_lite.read_vector(str(v)), [1.5, -2, 0.3])
    mfile = tmp_path / "m.bin"
    mat = np.arange(6, dtype=np.float32).reshape(2, 3)
    mfile.write_bytes(b"\0BFM \x04" + np.int32(2).tobytes() + b"\x04" + np.int32(3).tobytes() + mat.tobytes())
    np.testing.assert_array_equal(kaldi_io_lite.read_matrix(str(mfile)), mat)
    if os.path.isdir("/root/reference/Kaldi_Models"):
        km = "/root/reference/Kaldi_Models"
        m = npl.NeuralPlda(NC)
        m.LoadPldaParamsFromKaldi(km + "/mean.vec", km + "/transform.mat", km + "/plda")
        sd = m.state_dict()
        np.testing.assert_array_equal(sd["centering_and_LDA.weight"].numpy(), kaldi_params["W1"].numpy())
        np.testing.assert_array_equal(sd["centering_and_LDA.bias"].numpy(), kaldi_params["b1"].numpy())
        np.testing.assert_array_equal(sd["centering_and_wccn_plda.bias"].numpy(), kaldi_params["b2"].numpy())
        np.testing.assert_array_equal(sd["P_sqrt"].numpy(), kaldi_params["P_sqrt"].numpy())
        np.testing.assert_array_equal(sd["Q"].numpy(), kaldi_params["Q"].numpy())
        d = npl.DPlda(NCD)
        d.LoadParamsFromKaldi(km + "/mean.vec", km + "/transform.mat")
        np.testing.assert_array_equal(d.state_dict()["centering_and_LDA.weight"].numpy(), kaldi_params["W1"].numpy())


def test_trial_loaders_batch_layout(tmp_path):
    ids = [f"u{i}" for i in range(10)]
    id_to_num = {u: i for i, u in enumerate(ids)}
    f = tmp_path / "keys.tsv"
    f.write_text("u1 u2 1\nu3 u4.wav 0\nu5 missing 1\nu7 u8 0\n")
    np.random.seed(0)
    loader = L.combine_trials_and_get_loader([str(f)], id_to_num, batch_size=2)
    batches = list(loader)
    d1, d2, t = batches[0]
    assert d1.dtype == torch.int64 and d2.dtype == torch.int64 and t.dtype == torch.float32
    assert sum(len(b[0]) for b in batches) == 2            # 'u4.wav' and 'missing' rows dropped (no ext stripping here)
    loaders = L.get_trials_loaders_dict([str(f)], id_to_num, batch_size=8)
    assert list(loaders) == ["keys"]
    (d1, d2, t), = list(loaders["keys"])
    assert len(d1) == 3                                     # column 2 loses its extension here (sv_trials_loaders.py:403)


def test_shard_ranges_cover_the_trial_list():
    for n, w in ((10, 3), (50_000_000, 8), (7, 8), (0, 2)):
        r = [D.shard_range(n, w, k) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _skewed_trials():
    """Scores + labels whose target rate differs a lot between the two halves of the list."""
    g = torch.Generator().manual_seed(3)
    s = torch.randn(1001, generator=g)
    rate = torch.where(torch.arange(1001) < 500, torch.tensor(0.4), torch.tensor(0.03))
    t = (torch.rand(1001, generator=g) < rate).float()
    return s + 0.5 * t, t


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, t = _skewed_trials()
    a, b = D.shard_range(1001, world, rank)
    th, betas = [0.1, -0.2], [99.0, 199.0]
    acc = O.loss_accumulators(s[a:b], t[a:b], th, betas, 15.0, 0.0)       # this rank's raw sums (oracle stands in for K2)
    acc = D.allreduce_accumulators(acc)                                   # the reduction the GPU path uses
    soft, bce, hard = D.losses_from_accumulators(acc, betas)
    q.put((rank, soft, bce, hard, float(O.softcdet(s[a:b], t[a:b], th, betas, 15.0))))
    dist.destroy_process_group()


def test_sharded_accumulators_reduce_to_the_global_loss_gloo():
    """N > 1 host logic: raw sums are all-reduced BEFORE normalising; averaging the
    per-rank losses would be wrong because per-rank label counts differ."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    s, t = _skewed_trials()
    full = float(O.softcdet(s, t, [0.1, -0.2], [99.0, 199.0], 15.0))
    for rank, soft, bce, hard, local in res:
        assert soft == pytest.approx(full, rel=1e-6)
        assert bce == pytest.approx(float(O.crossentropy(s, t, 0.0)), rel=1e-6)
        assert hard == pytest.approx(float(O.cdet(s, t, [0.1, -0.2], [99.0, 199.0])), rel=1e-6)
    assert abs(np.mean([r[4] for r in res]) - full) > 1e-4 * full      # the naive mean of rank losses differs


def _gloo_grad_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    m = npl.NeuralPlda(NC)                                   # parameters only: no kernel runs on the CPU
    m.Th99.requires_grad_(False)                             # a parameter without a gradient must be skipped
    params = [p for p in m.parameters() if p.requires_grad]
    # gradients as the backward returns them: views of ONE flat buffer (functional._zero_grads)
    flat = torch.arange(sum(p.numel() for p in params), dtype=torch.float32) * (rank + 1)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view(p.shape)
        off += p.numel()
    D.allreduce_gradients(m)
    want = torch.arange(off, dtype=torch.float32) * sum(r + 1 for r in range(world))
    got = torch.cat([p.grad.reshape(-1) for p in params])
    q.put((rank, bool(torch.equal(got, want)), m.Th99.grad is None))
    dist.destroy_process_group()


def test_gradient_allreduce_gloo():
    """Training on trial-list shards: parameter gradients are summed across ranks, DDP-style; gradients
    that are views of one flat buffer and parameters without gradient are handled."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert res == [(0, True, True), (1, True, True)]
