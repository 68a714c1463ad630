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


def _score_trials(trials, mega_dict, model, device, batch_size):
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("neuralplda_b200 scores on the GPU only; got device '%s'" % device)
    tab = get_table(mega_dict, device)
    r1 = tab.rows_for_ids([strip_id(d) for d in trials[:, 0]])
    r2 = tab.rows_for_ids([strip_id(d) for d in trials[:, 1]])
    was_training = model.training
    model = model.to(device).eval()
    out = []
    for i in range(0, len(trials), batch_size):
        s, flag = model.forward_indexed(tab.table, r1[i:i + batch_size].to(device), r2[i:i + batch_size].to(device))
        out.append(s)
    scores = torch.cat(out).cpu().numpy() if out else np.zeros(0, np.float32)
    model.train(was_training)
    return scores.astype(np.float32).astype(str)


def generate_sre_scores(score_filename, trials_file, mega_dict, model, device, batch_size=102400):
    trials = np.genfromtxt(trials_file, dtype='str')
    header = '\t'.join(trials[0]) + '\tLLR'
    trials = trials[1:]
    scores = _score_trials(trials, mega_dict, model, device, batch_size)
    np.savetxt(score_filename, np.c_[trials, scores], header=header, fmt='%s', delimiter='\t', comments='')


def generate_voices_scores(score_filename, trials_file, mega_dict, model, device, batch_size=102400):
    trials = np.genfromtxt(trials_file, dtype='str')[:, :2]
    scores = _score_trials(trials, mega_dict, model, device, batch_size)
    np.savetxt(score_filename, np.c_[trials, scores], fmt='%s', delimiter='\t', comments='')
