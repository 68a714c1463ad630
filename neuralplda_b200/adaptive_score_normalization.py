"""Cohort score normalisation (Z / T / S / AS-norm) on the GPU -- the arithmetic of the reference's
utils/adaptive_score_normalization.py (a top-level script with hard-coded paths) as functions.

    :32-36  per id: mean / std of its cohort scores, and mean / std over the first ASnorm_topN entries of the
            ascending sort (the N lowest scores -- kept exactly as the script computes it)
    :61-66  znorm = (s - mean[e]) / std[e], tnorm = (s - mean[t]) / std[t], snorm = (z + t) / 2,
            asnorm1 = ((s - mean_top[e]) / std_top[e] + (s - mean_top[t]) / std_top[t]) / 2   (float64)

Device kernels: csrc/norm.cu (`nplda_cohort_stats`, `nplda_score_norm`); the id x cohort score matrix comes from
the grid kernel over rows embedded once (`model.forward_grid`).  GPU only, like the rest of the package.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import check, lib, on_device, ptr, require_cuda, stream_ptr
from .textio import TrialFile

ASnorm_topN = 500                                     # adaptive_score_normalization.py:12
NORMS = ("znorm", "tnorm", "snorm", "asnorm1")        # output order of normalize_scores / file suffixes (:75-78)


def cohort_statistics(cohort_scores, top_n=ASnorm_topN):
    """cohort_scores [m, c] float32 on the GPU (row = one enrol / test id against the c cohort utterances)
    -> [m, 4] float64: mean, std, mean_top, std_top (adaptive_score_normalization.py:32-36)."""
    require_cuda(cohort_scores)
    if cohort_scores.dim() != 2 or cohort_scores.shape[1] == 0:
        raise RuntimeError("cohort_scores must be [ids, cohort] with a non-empty cohort")
    if int(top_n) <= 0:
        raise RuntimeError("top_n must be positive")
    x = cohort_scores.to(torch.float32).contiguous()
    stats = torch.empty(x.shape[0], 4, dtype=torch.float64, device=x.device)
    with on_device(x.device):
        check(lib().nplda_cohort_stats(ptr(x), x.shape[0], x.shape[1], int(top_n), ptr(stats), stream_ptr()),
              "nplda_cohort_stats")
    return stats


def normalize_scores(raw_scores, enrol_rows, test_rows, stats):
    """raw_scores [n] float32, enrol_rows / test_rows [n] int64 rows of `stats` -> [4, n] float64 in the order
    of NORMS (adaptive_score_normalization.py:61-66).  A row outside the statistics raises KeyError, as the
    script's dict lookups do."""
    require_cuda(raw_scores, enrol_rows, test_rows, stats)
    raw = raw_scores.to(torch.float32).contiguous()
    er = enrol_rows.to(torch.int64).contiguous()
    tr = test_rows.to(torch.int64).contiguous()
    if not (raw.dim() == er.dim() == tr.dim() == 1 and raw.shape == er.shape == tr.shape):
        raise RuntimeError("raw_scores, enrol_rows and test_rows must be 1-D and the same length")
    if stats.dim() != 2 or stats.shape[1] != 4 or stats.dtype != torch.float64:
        raise RuntimeError("stats must be the [ids, 4] float64 tensor of cohort_statistics")
    stats = stats.contiguous()
    n = raw.numel()
    out = torch.empty(4, n, dtype=torch.float64, device=raw.device)
    flag = torch.zeros(1, dtype=torch.int32, device=raw.device)
    with on_device(raw.device):
        check(lib().nplda_score_norm(ptr(raw), ptr(er), ptr(tr), n, ptr(stats), stats.shape[0], ptr(out), ptr(flag),
                                     stream_ptr()), "nplda_score_norm")
    if n and int(flag.item()):
        raise KeyError("a trial refers to an id without cohort statistics")
    return out


def score_cohort(model, table, id_rows, cohort_rows):
    """The [ids, cohort] score matrix the script reads from its cohort score file: every id row of `table`
    against every cohort row, one grid product over rows that are transformed once (model.forward_grid)."""
    scores, flag = model.forward_grid(table, id_rows, cohort_rows)
    if scores.numel() and int(flag.item()):
        raise KeyError("an id or cohort row lies outside the x-vector table")
    return scores


def normalize_score_file(raw_score_filename, cohort_score_filename, device="cuda", top_n=ASnorm_topN):
    """The reference script end to end on two score files (adaptive_score_normalization.py:20-78): reads the raw
    trial scores (header row; enrol, test, ..., score) and the cohort scores (header row; id, cohort utterance,
    ..., score; rows grouped per id, every id against the same number of cohort utterances), normalises on the
    GPU and writes `<raw>_{znorm,tnorm,snorm,asnorm1}.tsv` in np.savetxt's format ('# ' header, tab separated).
    Returns the [4, n] float64 scores (host).  Scores travel to the GPU as float32 (what the score kernels
    emit); the script parses the text as float64, so values agree to ~1e-7 relative, not to the last digit."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("neuralplda_b200 normalises on the GPU only; got device '%s'" % device)
    with TrialFile(cohort_score_filename) as cf:
        ids_col = [cf.field(r, 0) for r in range(1, cf.rows)]                       # skip_header=1 (:28)
        num_unlabelled = len({cf.field(r, 1) for r in range(1, cf.rows)})            # :29
        coh, ok = cf.col_float(cf.cols - 1, first_row=1)
        if not ok.all():
            raise ValueError("could not convert a cohort score to float")
        if num_unlabelled == 0 or coh.size % num_unlabelled:
            raise ValueError("cannot reshape the cohort scores into rows of %d" % num_unlabelled)   # :33
        enrolls_of_cohort = ids_col[::num_unlabelled]                                # :38
    row_of = {}
    for k, name in enumerate(enrolls_of_cohort):                                     # dict(zip(...)): the last wins (:41-45)
        row_of[name] = k
    with TrialFile(raw_score_filename) as rf:
        header = rf.row(0)
        raw, ok = rf.col_float(rf.cols - 1, first_row=1)
        if not ok.all():
            raise ValueError("could not convert a raw score to float")
        body = [rf.row(r) for r in range(1, rf.rows)]
    er = np.asarray([row_of[b[0]] for b in body], dtype=np.int64)                    # KeyError like the script's dicts
    tr = np.asarray([row_of[b[1].replace('.sph', '')] for b in body], dtype=np.int64)  # :26
    stats = cohort_statistics(torch.from_numpy(coh.reshape(-1, num_unlabelled)).to(device), top_n)
    out = normalize_scores(torch.from_numpy(raw).to(device), torch.from_numpy(er).to(device),
                           torch.from_numpy(tr).to(device), stats).cpu().numpy()
    for k, name in enumerate(NORMS):                                                 # :75-78
        txt = out[k].astype(str)
        with open(raw_score_filename + "_%s.tsv" % name, "w") as f:
            f.write("# " + "\t".join(header) + "\n")
            for b, s in zip(body, txt):
                f.write("\t".join(b[:-1] + [s]) + "\n")
    return out
