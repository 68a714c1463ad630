"""Drop-in NeuralPlda / DPlda modules backed by libnplda.so (sm_100a kernels).

Mirrors the public surface of the reference's utils/models.py classes
(NeuralPlda 348-461, DPlda 463-569): same constructor (an `nc` config object),
same parameter names and registration order, same method names, argument
meaning and return types, so xvector_NeuralPlda_pytorch.py,
xvector_DPlda_pytorch.py, scorefile_generator.py and xvector_generate_scores.py
call it unchanged.  Every arithmetic step of forward / losses / backward / minc
runs in hand-written CUDA kernels through the C ABI in include/nplda.h; there
is no CPU or PyTorch-op fallback (CPU tensors raise).
"""
from __future__ import annotations

import pickle

import numpy as np
import torch
import torch.nn as nn

from . import _lib, functional as F_
from . import kaldi_io_lite
from .lazy import dense, lazy_pair

_TRANSIENT = ("_packed", "_impl", "_group", "_alpha_cache")


class _PldaBase(nn.Module):
    """State and methods shared by NeuralPlda and DPlda."""

    def _init_common(self, nc):
        # models.py:355-363 / 468-475 -- names, order and initial values kept
        self.threshold = {}
        for beta in nc.beta:
            self.threshold[beta] = nn.Parameter(0 * torch.rand(1, requires_grad=True))
            self.register_parameter("Th{}".format(int(beta)), self.threshold[beta])

    def _init_tail(self, nc):
        self.alpha = torch.tensor(nc.alpha).to(nc.device)
        self.beta = nc.beta
        self.dropout = nn.Dropout(p=0.5)
        self.lossfn = nc.loss

    # ---- transient (unpicklable / per-process) state -------------------------
    @property
    def packed(self):
        p = self.__dict__.get("_packed")
        if p is None:
            p = F_.PackedWeights()
            self.__dict__["_packed"] = p
        return p

    @property
    def impl(self):
        """Kernel selection: _lib.IMPL_AUTO (default), IMPL_SIMT or IMPL_TC."""
        return self.__dict__.get("_impl", _lib.IMPL_AUTO)

    @impl.setter
    def impl(self, v):
        self.__dict__["_impl"] = int(v)

    @property
    def process_group(self):
        """When set (True = default group), losses all-reduce their raw
        accumulators across ranks before normalising (trial-list sharding)."""
        return self.__dict__.get("_group")

    @process_group.setter
    def process_group(self, g):
        self.__dict__["_group"] = g

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in _TRANSIENT:
            state.pop(k, None)
        return state

    # The reference builds self.threshold before register_parameter, so the dict
    # holds the very Parameter objects in _parameters (models.py:355-358).  After
    # .to(device) nn.Module keeps Parameter identity (it swaps .data), so the
    # aliasing survives; after unpickling it is restored here.
    def __setstate__(self, state):
        super().__setstate__(state)
        if isinstance(self.__dict__.get("threshold"), dict):
            for beta in list(self.threshold):
                name = "Th{}".format(int(beta))
                if name in self._parameters:
                    self.threshold[beta] = self._parameters[name]

    def _apply(self, fn, *a, **k):
        super()._apply(fn, *a, **k)
        # the reference leaves alpha on nc.device; following the parameters is
        # required for .to(device) to be usable here
        try:
            dev = next(self.parameters()).device
            if isinstance(self.alpha, torch.Tensor) and self.alpha.device != dev:
                self.alpha = self.alpha.to(dev)
        except StopIteration:
            pass
        return self

    # ---- losses (models.py:384-404; 497-517) -----------------------------------
    def _alpha_value(self):
        """float(self.alpha) without a device->host read per loss call: alpha is a tensor that follows the module
        to the GPU (models.py:360); its value is cached against the tensor's identity and version."""
        a = self.alpha
        if not isinstance(a, torch.Tensor):
            return float(a)
        key = (id(a), a._version)
        c = self.__dict__.get("_alpha_cache")
        if c is None or c[0] != key:
            c = (key, float(a))
            self.__dict__["_alpha_cache"] = c
        return c[1]

    def _thresholds(self):
        if len(self.beta) == 0:
            return None
        return torch.cat([self.threshold[b] for b in self.beta])

    def _th_xent(self):
        return None

    def softcdet(self, output, target):
        return F_.LossFn.apply(output, target, self._thresholds(), self._th_xent(), self.beta,
                               self._alpha_value(), _lib.LOSS_SOFTCDET, self.process_group)

    def crossentropy(self, output, target):
        return F_.LossFn.apply(output, target, self._thresholds(), self._th_xent(), self.beta,
                               self._alpha_value(), _lib.LOSS_CROSSENTROPY, self.process_group)

    def loss(self, output, target):
        # case-sensitive dispatch, anything else returns None (models.py:395-399)
        if self.lossfn == 'SoftCdet':
            return self.softcdet(output, target)
        elif self.lossfn == 'crossentropy':
            return self.crossentropy(output, target)

    def cdet(self, output, target):
        with torch.no_grad():
            acc = F_.loss_accumulators(output, target, self._thresholds(), self._alpha_value(), self._th_xent(),
                                       self.process_group)
            return F_.finalize(acc, self.beta)[2].clone()

    def minc(self, output, target, update_thresholds=False, showplots=False):
        """Threshold sweep over every target score (models.py:406-436) on the GPU:
        sort both populations, then one sweep kernel (binary searches + first-argmin)."""
        if showplots:
            raise NotImplementedError("minc(showplots=True) is broken in the reference (removed matplotlib "
                                      "kwarg, models.py:427-432) and is not provided")
        _lib.require_cuda(output, target)
        with torch.no_grad():
            output = output.detach().float()
            target = target.detach().float()
            scores_target, scores_nontarget = F_.sorted_populations(output, target)
            if scores_target.numel() == 0:
                raise RuntimeError("minc: no target trials (the reference fails in torch.min on an empty tensor)")
            sums = torch.stack((target.sum(), (1 - target).sum())).tolist()          # synchronises ...
            from .sv_trials_loaders import check_pending_errors
            check_pending_errors()                                                    # ... so this is free here
            out_min, out_arg = F_.minc_sweep(scores_target, scores_nontarget, sums[0], sums[1], self.beta)
            minc_threshold = {}
            for k, beta in enumerate(self.beta):
                minc_threshold[beta] = scores_target[out_arg[k]]
                if update_thresholds:
                    self.state_dict()["Th{}".format(int(beta))].data.copy_(minc_threshold[beta])
            minc_avg = out_min.sum() / len(self.beta)
        return minc_avg, minc_threshold

    def _forward_lazy(self, kind, x1, x2):
        """Scores for a pair of lazily gathered loader outputs (lazy.py) straight from (table, row indices): the
        materialised [B, D] pair the reference's loop passes between its loader and forward is never written.  Returns
        None when that shortcut does not apply (gradients wanted, or the rows were already materialised)."""
        lp = lazy_pair(x1, x2)
        if lp is None or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            return None
        tab, r1, r2 = lp
        from .sv_trials_loaders import bad_flag_of
        scores, _ = F_.score_indexed(kind, tab.table, r1, r2, self._params(), self._dims(), self.packed, self.impl,
                                     flag_ptr=bad_flag_of(tab).ptr)
        return scores

    def SaveModel(self, filename):
        with open(filename, 'wb') as f:
            pickle.dump(self, f)


class NeuralPlda(_PldaBase):
    def __init__(self, nc):
        super(NeuralPlda, self).__init__()
        self.centering_and_LDA = nn.Linear(nc.xvector_dim, nc.layer1_LDA_dim)  # Centering, wccn
        self.centering_and_wccn_plda = nn.Linear(nc.layer1_LDA_dim, nc.layer2_PLDA_spkfactor_dim)
        self.P_sqrt = nn.Parameter(torch.rand(nc.layer2_PLDA_spkfactor_dim, requires_grad=True))
        self.Q = nn.Parameter(torch.rand(nc.layer2_PLDA_spkfactor_dim, requires_grad=True))
        self._init_common(nc)
        self.threshold_Xent = nn.Parameter(0 * torch.rand(1, requires_grad=True))
        self._init_tail(nc)

    def _th_xent(self):
        return self.threshold_Xent

    def _dims(self):
        return (self.centering_and_LDA.in_features, self.centering_and_LDA.out_features,
                self.centering_and_wccn_plda.out_features)

    def _params(self):
        return (self.centering_and_LDA.weight, self.centering_and_LDA.bias,
                self.centering_and_wccn_plda.weight, self.centering_and_wccn_plda.bias, self.P_sqrt, self.Q)

    def forward(self, x1, x2):
        s = self._forward_lazy("nplda", x1, x2)
        if s is not None:
            return s
        return F_.NpldaScoreFn.apply(dense(x1), dense(x2), *self._params(), self.packed, self.impl, torch.is_grad_enabled())

    def forward_indexed(self, table, idx1, idx2, embed_once=None, use_split=None):
        """Scores for trials given as row indices into a device-resident x-vector table (no gradient).
        Every table row is transformed once and cached (F_.score_indexed); `embed_once=False` recomputes both sides
        per trial -- on the tensor cores from the pre-split image of the table (`use_split`, the default for shapes
        the CTA-pair kernel takes) or with the fp32 kernel that fuses the gather."""
        with torch.no_grad():
            scores, flag = F_.score_indexed("nplda", table, idx1, idx2, self._params(), self._dims(),
                                            self.packed, self.impl, embed_once, use_split)
        return scores, flag

    def forward_grid(self, table, enrol_rows, test_rows):
        """[E, T] scores of every enrol row against every test row of a device-resident x-vector table
        (enrol-major, the order of an enrol x test trial list); no gradient.  Returns (scores, bad_index_flag)."""
        with torch.no_grad():
            return F_.score_grid("nplda", table, enrol_rows, test_rows, self._params(), self._dims(), self.packed, self.impl)

    # The two half-steps of forward are part of the reference's public surface (models.py:366-376).
    # No reference call site uses them separately (forward is the hot path and never materialises
    # embeddings); they are served by the fp32 kernels and are not differentiable here.
    def extract_plda_embeddings(self, x):
        return F_.embed("nplda", x, self._params(), self._dims(), self.packed)

    def forward_from_plda_embeddings(self, x1, x2):
        return F_.score_from_embeddings("nplda", dense(x1), dense(x2), self._params(), self._dims(), self.packed)

    def LoadPldaParamsFromKaldi(self, mean_vec_file, transform_mat_file, PldaFile):
        """models.py:441-457 without the Kaldi binaries (files parsed directly)."""
        plda = kaldi_io_lite.read_plda(PldaFile)
        transform_mat = kaldi_io_lite.read_matrix(transform_mat_file)
        mean_vec = kaldi_io_lite.read_vector(mean_vec_file)
        mdsd = self.state_dict()
        mdsd['centering_and_LDA.weight'].data.copy_(torch.from_numpy(transform_mat[:, :-1]).float())
        mdsd['centering_and_LDA.bias'].data.copy_(
            torch.from_numpy(transform_mat[:, -1] - transform_mat[:, :-1].dot(mean_vec)).float())
        mdsd['centering_and_wccn_plda.weight'].data.copy_(torch.from_numpy(plda['diagonalizing_transform']).float())
        mdsd['centering_and_wccn_plda.bias'].data.copy_(
            torch.from_numpy(-plda['diagonalizing_transform'].dot(plda['plda_mean'])).float())
        mdsd['P_sqrt'].data.copy_(torch.from_numpy(np.sqrt(plda['diagP'])).float())
        mdsd['Q'].data.copy_(torch.from_numpy(plda['diagQ']).float())


class DPlda(_PldaBase):
    def __init__(self, nc):
        super(DPlda, self).__init__()
        self.centering_and_LDA = nn.Linear(nc.xvector_dim, nc.layer1_LDA_dim)  # Centering, wccn
        self.logistic_regres = nn.Linear(nc.layer1_LDA_dim * nc.layer1_LDA_dim * 2 + nc.layer1_LDA_dim, 1)
        self._init_common(nc)
        self._init_tail(nc)

    def _dims(self):
        d1 = self.centering_and_LDA.out_features
        return (self.centering_and_LDA.in_features, d1, d1)

    def _params(self):
        return (self.centering_and_LDA.weight, self.centering_and_LDA.bias,
                self.logistic_regres.weight, self.logistic_regres.bias)

    def forward(self, x1, x2):
        s = self._forward_lazy("dplda", x1, x2)
        if s is not None:
            return s
        return F_.DpldaScoreFn.apply(dense(x1), dense(x2), *self._params(), self.packed, self.impl, torch.is_grad_enabled())

    def forward_indexed(self, table, idx1, idx2, embed_once=None):
        with torch.no_grad():
            scores, flag = F_.score_indexed("dplda", table, idx1, idx2, self._params(), self._dims(),
                                            self.packed, self.impl, embed_once)
        return scores, flag

    def forward_grid(self, table, enrol_rows, test_rows):
        """[E, T] scores of every enrol row against every test row of a device-resident x-vector table
        (enrol-major, the order of an enrol x test trial list); no gradient.  Returns (scores, bad_index_flag)."""
        with torch.no_grad():
            return F_.score_grid("dplda", table, enrol_rows, test_rows, self._params(), self._dims(), self.packed, self.impl)

    def extract_plda_embeddings(self, x):
        return F_.embed("dplda", x, self._params(), self._dims(), self.packed)

    def forward_from_plda_embeddings(self, x1, x2):
        return F_.score_from_embeddings("dplda", dense(x1), dense(x2), self._params(), self._dims(), self.packed)

    def LoadParamsFromKaldi(self, mean_vec_file, transform_mat_file):
        """models.py:551-563 without the Kaldi binaries."""
        transform_mat = kaldi_io_lite.read_matrix(transform_mat_file)
        mean_vec = kaldi_io_lite.read_vector(mean_vec_file)
        mdsd = self.state_dict()
        mdsd['centering_and_LDA.weight'].data.copy_(torch.from_numpy(transform_mat[:, :-1]).float())
        mdsd['centering_and_LDA.bias'].data.copy_(
            torch.from_numpy(transform_mat[:, -1] - transform_mat[:, :-1].dot(mean_vec)).float())
