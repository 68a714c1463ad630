"""CPU oracle for the NeuralPlda / DPlda pairwise trial-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``neuralplda_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline -- never as the thing shipped.

It is an independent restatement (torch-CPU / numpy, explicit arithmetic) of
the algorithm in the reference ``/root/reference/utils/models.py`` and of the
data formats either side of it.  Each function cites the reference lines it
follows.  The arithmetic of the reference lives in PyTorch library calls
(nn.Linear, F.normalize, F.binary_cross_entropy, torch.sort); no version is
pinned by the reference (README.md:12-15), so "the reference" here means the
reference code on this image's torch 2.11 CPU.

Parity pinning: the reference ships NO tests or golden vectors for this path
(SURVEY.md section 8c).  The oracle is pinned instead against outputs of the
reference itself, imported unmodified in the build container by
``tests/golden/make_golden.py``; those outputs are committed under
``tests/golden/`` and ``tests/test_oracle.py`` checks this file against them.
"""
from __future__ import annotations

import os
import struct

import numpy as np
import torch

# ----------------------------------------------------------------------------
# Scores
# ----------------------------------------------------------------------------


def length_norm(a: torch.Tensor) -> torch.Tensor:
    """F.normalize(a) with its defaults p=2, dim=1, eps=1e-12 (models.py:368).

    Unit norm, denominator clamped from below: a / max(||a||_2, 1e-12).
    """
    nrm = a.pow(2).sum(dim=1, keepdim=True).sqrt()
    return a / nrm.clamp_min(1e-12)


def nplda_embed(x, W1, b1, W2, b2):
    """NeuralPlda.extract_plda_embeddings (models.py:366-370).

    a = x W1^T + b1 ; u = a / max(||a||, 1e-12) ; y = u W2^T + b2.
    """
    a = x @ W1.t() + b1
    u = length_norm(a)
    return u @ W2.t() + b2


def nplda_score_from_embeddings(y1, y2, P_sqrt, Q):
    """NeuralPlda.forward_from_plda_embeddings (models.py:372-376).

    P = P_sqrt * P_sqrt (diagonal), cross term carries a factor 2, no constant.
    """
    P = P_sqrt * P_sqrt
    return (y1 * Q * y1).sum(1) + (y2 * Q * y2).sum(1) + 2.0 * (y1 * P * y2).sum(1)


def nplda_score(x1, x2, W1, b1, W2, b2, P_sqrt, Q):
    """NeuralPlda.forward (models.py:378-382)."""
    return nplda_score_from_embeddings(
        nplda_embed(x1, W1, b1, W2, b2), nplda_embed(x2, W1, b1, W2, b2), P_sqrt, Q)


def dplda_embed(x, W1, b1):
    """DPlda.extract_plda_embeddings (models.py:478-481): affine + length norm."""
    return length_norm(x @ W1.t() + b1)


def dplda_score_expanded(u1, u2, w_lr, c_lr):
    """DPlda.forward_from_plda_embeddings, literal form (models.py:483-489).

    Builds the 2*L*L+L pair expansion [vec(u1u2^T+u2u1^T), vec(u1u1^T+u2u2^T),
    u1+u2] and applies Linear(.,1).  232 KB per trial at L=170: small N only.
    """
    n = u1.shape[0]
    between = (u1[:, :, None] * u2[:, None, :] + u2[:, :, None] * u1[:, None, :]).reshape(n, -1)
    within = (u1[:, :, None] * u1[:, None, :] + u2[:, :, None] * u2[:, None, :]).reshape(n, -1)
    feats = torch.cat((between, within, u1 + u2), dim=1)
    return feats @ w_lr.reshape(-1) + c_lr.reshape(())


def dplda_split_weight(w_lr, L):
    """Row-major L x L reshapes of logistic_regres.weight (models.py:467, 484-487)."""
    w = w_lr.reshape(-1)
    return w[: L * L].reshape(L, L), w[L * L: 2 * L * L].reshape(L, L), w[2 * L * L:]


def dplda_score_closed(u1, u2, w_lr, c_lr):
    """Closed form of the same score:

    S = u1^T (Wb+Wb^T) u2 + u1^T Ww u1 + u2^T Ww u2 + ws.(u1+u2) + c
    """
    L = u1.shape[1]
    Wb, Ww, ws = dplda_split_weight(w_lr, L)
    Pm = Wb + Wb.t()
    return ((u1 @ Pm) * u2).sum(1) + ((u1 @ Ww) * u1).sum(1) + ((u2 @ Ww) * u2).sum(1) \
        + (u1 + u2) @ ws + c_lr.reshape(())


def dplda_score(x1, x2, W1, b1, w_lr, c_lr, expanded=False):
    """DPlda.forward (models.py:491-495)."""
    u1, u2 = dplda_embed(x1, W1, b1), dplda_embed(x2, W1, b1)
    return (dplda_score_expanded if expanded else dplda_score_closed)(u1, u2, w_lr, c_lr)


# ----------------------------------------------------------------------------
# Losses and detection costs
# ----------------------------------------------------------------------------


def softcdet(s, t, thresholds, betas, alpha):
    """NeuralPlda.softcdet / DPlda.softcdet (models.py:384-388, 497-501).

    mean_k [ sum_i t_i sig(alpha (th_k - s_i)) / sum t
             + beta_k sum_i (1-t_i) sig(alpha (s_i - th_k)) / sum (1-t) ]
    Normalised by the GLOBAL label sums of the tensors passed in.
    """
    nt, nn_ = t.sum(), (1 - t).sum()
    terms = []
    for th, beta in zip(thresholds, betas):
        pmiss = (torch.sigmoid(alpha * (th - s)) * t).sum() / nt
        pfa = (torch.sigmoid(alpha * (s - th)) * (1 - t)).sum() / nn_
        terms.append(pmiss + beta * pfa)
    return sum(terms) / len(terms)


def crossentropy(s, t, threshold_xent=0.0):
    """NeuralPlda.crossentropy (models.py:390-393); DPlda's (503-506) passes 0.

    binary_cross_entropy(sigmoid(z), t), mean reduction, log clamped at -100
    (PyTorch semantics), z = s - threshold_Xent.
    """
    p = torch.sigmoid(s - threshold_xent)
    logp = torch.log(p).clamp_min(-100.0)
    log1mp = torch.log(1 - p).clamp_min(-100.0)
    return -(t * logp + (1 - t) * log1mp).mean()


def cdet(s, t, thresholds, betas):
    """NeuralPlda.cdet (models.py:401-404): hard decisions, strict < and >."""
    nt, nn_ = t.sum(), (1 - t).sum()
    terms = []
    for th, beta in zip(thresholds, betas):
        pmiss = ((s < th).float() * t).sum() / nt
        pfa = ((s > th).float() * (1 - t)).sum() / nn_
        terms.append(pmiss + beta * pfa)
    return sum(terms) / len(terms)


def minc_loop(s, t, betas):
    """NeuralPlda.minc, literal O(N_t * N) restatement (models.py:406-421, arr2val 23-27).

    For every target score s_j (ascending): the reference takes the LAST INDEX
    of the sorted targets strictly below s_j (= count-1), or 1.0 if none; and
    the last index of the descending-sorted non-targets >= s_j (= count-1), or
    1.0 if none.  Small inputs only.
    """
    st = np.sort(s[t > 0.5].numpy())
    sn = np.sort(s[t < 0.5].numpy())[::-1]
    pm, pf = [], []
    for v in st:
        below = np.nonzero(st < v)[0]
        pm.append(float(below[-1]) if below.size else 1.0)
        above = np.nonzero(sn >= v)[0]
        pf.append(float(above[-1]) if above.size else 1.0)
    pmiss = torch.tensor(pm).float() / t.sum()
    pfa = torch.tensor(pf).float() / (1 - t).sum()
    mins, ths = [], {}
    for beta in betas:
        c = pmiss + beta * pfa
        v, idx = torch.min(c, 0)
        mins.append(v)
        ths[beta] = torch.tensor(st[int(idx)])
    return sum(mins) / len(mins), ths


def minc(s, t, betas):
    """Same result as ``minc_loop`` in O(N log N) with searchsorted (numpy)."""
    sv = s.numpy()
    st = np.sort(sv[t.numpy() > 0.5])
    sn = np.sort(sv[t.numpy() < 0.5])
    cnt_below = np.searchsorted(st, st, side="left")              # targets strictly < s_j
    cnt_ge = sn.size - np.searchsorted(sn, st, side="left")       # non-targets >= s_j
    pm = np.where(cnt_below > 0, cnt_below - 1, 1).astype(np.float32)
    pf = np.where(cnt_ge > 0, cnt_ge - 1, 1).astype(np.float32)
    pmiss = torch.from_numpy(pm) / t.sum()
    pfa = torch.from_numpy(pf) / (1 - t).sum()
    mins, ths = [], {}
    for beta in betas:
        c = pmiss + beta * pfa
        v, idx = torch.min(c, 0)
        mins.append(v)
        ths[beta] = torch.tensor(st[int(idx)])
    return sum(mins) / len(mins), ths


# ----------------------------------------------------------------------------
# Loss accumulators (the payload that is all-reduced across GPUs)
# ----------------------------------------------------------------------------


def loss_accumulators(s, t, thresholds, betas, alpha, threshold_xent=0.0):
    """fp64 raw sums in the layout of include/nplda.h:

    [A_k (soft miss sum), B_k (soft fa sum), miss_cnt_k, fa_cnt_k] for k<K,
    then N_t, N_n, sum_bce, N.
    """
    s64, t64 = s.double(), t.double()
    acc = []
    for th in thresholds:
        th = float(th)
        acc.append((torch.sigmoid(alpha * (th - s64)) * t64).sum())
        acc.append((torch.sigmoid(alpha * (s64 - th)) * (1 - t64)).sum())
        acc.append(((s64 < th).double() * t64).sum())
        acc.append(((s64 > th).double() * (1 - t64)).sum())
    z = s64 - float(threshold_xent)
    # fp32 sigmoid saturates to exactly 0/1 where the reference clamps the log
    # at -100; mirror that by evaluating p in fp32 like the reference does.
    p = torch.sigmoid(z.float())
    bce = -(t * torch.log(p).clamp_min(-100.0) + (1 - t) * torch.log(1 - p).clamp_min(-100.0))
    acc += [t64.sum(), (1 - t64).sum(), bce.double().sum(), torch.tensor(float(s.numel()), dtype=torch.float64)]
    return torch.stack([a.double() for a in acc])


# ----------------------------------------------------------------------------
# Kaldi model files -> initial parameters
# ----------------------------------------------------------------------------


def _read_token(buf, pos):
    end = buf.index(b" ", pos)
    return buf[pos:end].decode(), end + 1


def _read_int32(buf, pos):
    assert buf[pos] == 4, "expected \\x04 size marker"
    return struct.unpack_from("<i", buf, pos + 1)[0], pos + 5


def _read_bin_vector(buf, pos):
    tok, pos = _read_token(buf, pos)
    dt = {"FV": np.float32, "DV": np.float64}[tok]
    n, pos = _read_int32(buf, pos)
    v = np.frombuffer(buf, dtype=dt, count=n, offset=pos).astype(np.float64)
    return v, pos + n * np.dtype(dt).itemsize


def _read_bin_matrix(buf, pos):
    tok, pos = _read_token(buf, pos)
    dt = {"FM": np.float32, "DM": np.float64}[tok]
    r, pos = _read_int32(buf, pos)
    c, pos = _read_int32(buf, pos)
    m = np.frombuffer(buf, dtype=dt, count=r * c, offset=pos).astype(np.float64).reshape(r, c)
    return m, pos + r * c * np.dtype(dt).itemsize


def read_kaldi_vector(path):
    """Kaldi vector, text (' [ v v v ]') or binary ('\\0B' + FV/DV).  What
    ``copy-vector --binary=false`` feeds models.py:446-448."""
    buf = open(path, "rb").read()
    if buf[:2] == b"\0B":
        return _read_bin_vector(buf, 2)[0]
    txt = buf.decode().replace("[", " ").replace("]", " ")
    return np.asarray(txt.split(), dtype=np.float64)


def read_kaldi_matrix(path):
    """Kaldi matrix, binary FM/DM or text.  What ``copy-matrix --binary=false``
    feeds models.py:443-445."""
    buf = open(path, "rb").read()
    if buf[:2] == b"\0B":
        return _read_bin_matrix(buf, 2)[0]
    rows = [r.split() for r in buf.decode().replace("[", " ").replace("]", " ").strip().split("\n")]
    return np.asarray([r for r in rows if r], dtype=np.float64)


def read_kaldi_plda(path):
    """Binary Kaldi <Plda> object: mean (vector), transform (matrix), psi (vector);
    derives diagP/diagQ exactly as kaldiPlda2numpydict.py:34-38."""
    buf = open(path, "rb").read()
    assert buf[:2] == b"\0B", "text-mode plda not supported by the oracle reader"
    tok, pos = _read_token(buf, 2)
    assert tok == "<Plda>"
    mean, pos = _read_bin_vector(buf, pos)
    transform, pos = _read_bin_matrix(buf, pos)
    psi, pos = _read_bin_vector(buf, pos)
    ac = psi
    tot = 1.0 + psi
    return {
        "plda_mean": mean,
        "diagonalizing_transform": transform,
        "Psi_across_covar_diag": psi,
        "diagP": ac / (tot * (tot - ac * ac / tot)),
        "diagQ": (1.0 / tot) - 1.0 / (tot - ac * ac / tot),
    }


def kaldi_init_params(mean_vec_file, transform_mat_file, plda_file):
    """NeuralPlda.LoadPldaParamsFromKaldi (models.py:441-457) without the Kaldi
    binaries: W1 = T[:, :-1]; b1 = T[:, -1] - W1.mean; W2 = plda transform;
    b2 = -W2.plda_mean; P_sqrt = sqrt(diagP); Q = diagQ.  float64 -> float32."""
    T = read_kaldi_matrix(transform_mat_file)
    mean = read_kaldi_vector(mean_vec_file)
    plda = read_kaldi_plda(plda_file)
    W1 = T[:, :-1]
    W2 = plda["diagonalizing_transform"]
    return {
        "W1": W1.astype(np.float32),
        "b1": (T[:, -1] - W1.dot(mean)).astype(np.float32),
        "W2": W2.astype(np.float32),
        "b2": (-W2.dot(plda["plda_mean"])).astype(np.float32),
        "P_sqrt": np.sqrt(plda["diagP"]).astype(np.float32),
        "Q": plda["diagQ"].astype(np.float32),
        "mean": mean.astype(np.float32),
    }


# ----------------------------------------------------------------------------
# Synthetic workloads.  Deterministic CPU generators so
# the same tensors can be rebuilt on the GPU box without shipping them.
# ----------------------------------------------------------------------------


def synth_pairs(n_pairs, n_speakers, seed, mean=None, p_target=0.1, dim=512, noise=0.7):
    """x = mu + c_spk + noise*eps, speaker-structured; labels 1 for same speaker."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.zeros(dim) if mean is None else torch.as_tensor(mean, dtype=torch.float32)
    centers = torch.randn(n_speakers, dim, generator=g)
    spk1 = torch.randint(0, n_speakers, (n_pairs,), generator=g)
    is_tgt = torch.rand(n_pairs, generator=g) < p_target
    shift = torch.randint(1, n_speakers, (n_pairs,), generator=g)
    spk2 = torch.where(is_tgt, spk1, (spk1 + shift) % n_speakers)
    x1 = mu + centers[spk1] + noise * torch.randn(n_pairs, dim, generator=g)
    x2 = mu + centers[spk2] + noise * torch.randn(n_pairs, dim, generator=g)
    return x1.contiguous(), x2.contiguous(), is_tgt.float()


def synth_grid(n_enrol, n_test, n_speakers, seed, mean=None, dim=512, noise=0.7):
    """Enrol x test grid: returns the unique-vector table, and per-trial indices
    (row-major, enrol-major) with label = same speaker."""
    g = torch.Generator().manual_seed(seed)
    mu = torch.zeros(dim) if mean is None else torch.as_tensor(mean, dtype=torch.float32)
    centers = torch.randn(n_speakers, dim, generator=g)
    spk = torch.randint(0, n_speakers, (n_enrol + n_test,), generator=g)
    table = mu + centers[spk] + noise * torch.randn(n_enrol + n_test, dim, generator=g)
    i1 = torch.arange(n_enrol).repeat_interleave(n_test)
    i2 = n_enrol + torch.arange(n_test).repeat(n_enrol)
    labels = (spk[i1] == spk[i2]).float()
    return table.contiguous(), i1, i2, labels


# ----------------------------------------------------------------------------
# Data formats either side of the path
# ----------------------------------------------------------------------------


def strip_id(s):
    """basename + splitext, as load_xvec_trials_from_idbatch applies to both ids
    (sv_trials_loaders.py:432)."""
    return os.path.splitext(os.path.basename(s))[0]


def gather_numbatch(mega_dict, num_to_id, d1, d2):
    """load_xvec_trials_from_numbatch (sv_trials_loaders.py:418-426)."""
    a = np.asarray([mega_dict[num_to_id[int(i)]] for i in d1])
    b = np.asarray([mega_dict[num_to_id[int(i)]] for i in d2])
    return torch.from_numpy(a).float(), torch.from_numpy(b).float()


def format_scores(scores_f32):
    """Score text as written by scorefile_generator.py:37,54: str(np.float32)."""
    return np.asarray(scores_f32, dtype=np.float32).astype(str)


# ----------------------------------------------------------------------------
# Cohort score normalisation
# ----------------------------------------------------------------------------


def cohort_stats(cohort_scores, top_n=500):
    """utils/adaptive_score_normalization.py:32-36.  cohort_scores [m, c] (one row per enrol / test id) ->
    [m, 4] float64: mean, std over the row; mean, std over the first top_n entries of the ASCENDING sort
    (the top_n lowest scores -- the script sorts ascending and slices [:ASnorm_topN]).  np.std is the
    population standard deviation."""
    s = np.sort(np.asarray(cohort_scores, dtype=np.float64), axis=1)
    return np.stack([np.mean(s, axis=1), np.std(s, axis=1),
                     np.mean(s[:, :top_n], axis=1), np.std(s[:, :top_n], axis=1)], axis=1)


def score_norm(raw, enrol_row, test_row, stats):
    """utils/adaptive_score_normalization.py:61-66 for every trial; rows index `stats` (the script's dicts).
    Returns [4, n] float64: znorm, tnorm, snorm, asnorm1."""
    raw = np.asarray(raw, dtype=np.float64)
    e, t = stats[np.asarray(enrol_row)], stats[np.asarray(test_row)]
    z = (raw - e[:, 0]) / e[:, 1]
    tn = (raw - t[:, 0]) / t[:, 1]
    sn = (z + tn) / 2
    asn = ((raw - e[:, 2]) / e[:, 3] + (raw - t[:, 2]) / t[:, 3]) / 2
    return np.stack([z, tn, sn, asn])
