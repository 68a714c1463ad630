"""Score-file writers: same signatures and file formats as the reference's
utils/scorefile_generator.py:22-56, but the model stays on the GPU (the
reference moves it to the CPU and back, :25/:39/:44/:56), the x-vector dict is
uploaded once, and the pair gather is fused into the score kernel.

Formats kept: voices = first two columns + score, tab separated, no header
(:45,55); sre = header row + all original columns + 'LLR' (:27-28,38); scores
are written as str(np.float32).  The reference's crash when len(trials) is a
multiple of batch_size (an empty last slice reaches nn.Linear) is not
reproduced: empty slices are skipped.
"""
from __future__ import annotations

import numpy as np
import torch

from .sv_trials_loaders import get_table, strip_id
from .textio import TrialFile, MODE_BASENAME_SPLITEXT


def _score_rows(r1, r2, tab, model, device, batch_size):
    was_training = model.training
    model = model.to(device).eval()
    out = []
    r1, r2 = torch.from_numpy(r1), torch.from_numpy(r2)
    for i in range(0, len(r1), batch_size):
        s, flag = model.forward_indexed(tab.table, r1[i:i + batch_size].to(device), r2[i:i + batch_size].to(device))
        out.append(s)
    scores = torch.cat(out).cpu().numpy() if out else np.zeros(0, np.float32)
    model.train(was_training)
    return scores.astype(np.float32)


def _generate(score_filename, trials_file, mega_dict, model, device, batch_size, sre):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("neuralplda_b200 scores on the GPU only; got device '%s'" % device)
    tab = get_table(mega_dict, device)
    with TrialFile(trials_file) as tf:
        first = 1 if sre else 0                      # sre files carry a header row (scorefile_generator.py:27-28)
        if tf.cols < 2:
            raise IndexError("trial files need at least the enrol and test columns")
        r1 = tf.map_ids(0, tab.ids, mode=MODE_BASENAME_SPLITEXT, first_row=first)     # sv_trials_loaders.py:432
        r2 = tf.map_ids(1, tab.ids, mode=MODE_BASENAME_SPLITEXT, first_row=first)
        bad = np.nonzero((r1 < 0) | (r2 < 0))[0]
        if bad.size:                                  # the reference's dict lookup raises KeyError
            row = int(bad[0]) + first
            raise KeyError(strip_id(tf.field(row, 0 if r1[bad[0]] < 0 else 1)))
        scores = _score_rows(r1, r2, tab, model, device, batch_size)
        if sre:
            header = "\t".join(tf.row(0)) + "\tLLR" if tf.rows else "\tLLR"
            tf.write_scores(score_filename, scores, tf.cols, header_line=header, first_row=first)
        else:
            tf.write_scores(score_filename, scores, 2)


def generate_sre_scores(score_filename, trials_file, mega_dict, model, device, batch_size=102400):
    """scorefile_generator.py:22-39: header row kept, all columns + "LLR"."""
    _generate(score_filename, trials_file, mega_dict, model, device, batch_size, sre=True)


def generate_voices_scores(score_filename, trials_file, mega_dict, model, device, batch_size=102400):
    """scorefile_generator.py:41-56: first two columns + score, no header."""
    _generate(score_filename, trials_file, mega_dict, model, device, batch_size, sre=False)
