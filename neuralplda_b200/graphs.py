"""CUDA-graph replay of the training-loop body for small batches.

The reference trains with 128-pair batches (conf/sre_config.cfg:30) and its loop body
(xvector_NeuralPlda_pytorch.py:35-43: gather, forward, loss, backward, optimizer.step) is then ~35 small kernel launches
bound by launch and Python latencies, not by the kernels.  `GraphedTrainStep` captures that body once -- the batch
gather from the device-resident x-vector table, the module's forward, `model.loss`, the backward and the optimiser step
-- and replays it per batch: one graph launch instead of ~35, no Python between the kernels.

    step = GraphedTrainStep(model, optimizer, mega_xvec_dict, num_to_id_dict, batch_size=nc.batch_size)
    for data1, data2, target in train_loader:                       # the reference's loop, lines 36-43 replaced
        loss = step(data1, data2, target)                           # 0-d CUDA tensor; .item() only when logging

It is opt-in and changes nothing else: same kernels, same arithmetic, parameters and optimiser state are updated in
place.  Requirements of CUDA-graph capture: a fixed batch size (a short last batch runs eagerly through the same
code), an optimiser whose step is capturable (`torch.optim.Adam(..., capturable=True)`, optionally `fused=True`), and no
host reads inside the step (the loss comes back as a device tensor).
"""
from __future__ import annotations

import torch

from .sv_trials_loaders import _device_rows, _gather, _rows_from_nums, get_table


class GraphedTrainStep:
    def __init__(self, model, optimizer, mega_dict, num_to_id_dict, batch_size, device=None, warmup=3):
        self.model, self.optimizer = model, optimizer
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedTrainStep needs the module on a CUDA device")
        self.device = dev
        d = optimizer.defaults
        if not d.get("capturable"):
            raise RuntimeError("GraphedTrainStep needs an optimiser whose step can be captured in a CUDA graph: "
                               "torch.optim.Adam(..., capturable=True), optionally with fused=True")
        self.mega_dict, self.num_to_id = mega_dict, num_to_id_dict
        self.tab = get_table(mega_dict, dev)
        self.batch_size = int(batch_size)
        B = self.batch_size
        self.i12 = torch.zeros(2 * B, dtype=torch.int64, device=dev)       # static inputs of the graph
        self.i1, self.i2 = self.i12[:B], self.i12[B:]
        self.t = torch.zeros(B, dtype=torch.float32, device=dev)
        # batches that arrive on the host (DataLoader output) are staged in pinned memory: two asynchronous copies
        self.h_i12 = torch.zeros(2 * B, dtype=torch.int64).pin_memory()
        self.h_t = torch.zeros(B, dtype=torch.float32).pin_memory()
        self._staged = torch.cuda.Event()
        self.loss = None
        self.graph = None
        self._warmup = int(warmup)
        self._seen = 0

    def flush(self):
        """Synchronise and report a trial of ANY earlier batch (the last one included) that referred to a row outside
        the x-vector table -- the reference raises KeyError from its dict lookup at once."""
        from .sv_trials_loaders import check_pending_errors
        check_pending_errors()

    def _body(self, i1, i2, t):
        self.optimizer.zero_grad(set_to_none=False)                 # static .grad buffers: the graph accumulates into them
        x1, x2 = _gather(self.tab, i1, i2, lazy=False)
        loss = self.model.loss(self.model(x1, x2), t)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def _capture(self):
        # every tensor the body allocates comes from the graph's private pool; the inputs are the static buffers
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._body(self.i1, self.i2, self.t)

    def __call__(self, data1, data2, target):
        dev = self.device
        B = self.batch_size
        if not (data1.is_cuda or data2.is_cuda or target.is_cuda) and data1.numel() == B:
            r1 = _rows_from_nums(self.tab, self.num_to_id, data1)   # host int64 rows, range-checked (KeyError)
            r2 = _rows_from_nums(self.tab, self.num_to_id, data2)
            self._staged.synchronize()                              # the previous batch has left the staging buffers
            self.h_i12[:B].copy_(r1); self.h_i12[B:].copy_(r2); self.h_t.copy_(target)
            self.i12.copy_(self.h_i12, non_blocking=True)
            self.t.copy_(self.h_t, non_blocking=True)
            self._staged.record(torch.cuda.current_stream(dev))
        else:
            r1 = _device_rows(self.tab, self.num_to_id, data1.to(dev) if not data1.is_cuda else data1, dev)
            r2 = _device_rows(self.tab, self.num_to_id, data2.to(dev) if not data2.is_cuda else data2, dev)
            target = target.to(dev, torch.float32)
            if r1.numel() != B:                                     # short last batch: same code, eagerly
                return self._body(r1, r2, target)
            self.i1.copy_(r1); self.i2.copy_(r2); self.t.copy_(target)
        if self.graph is None:
            if self._seen < self._warmup:                           # eager warm-up steps (real training steps) on a side stream
                self._seen += 1
                s = torch.cuda.Stream(device=dev)
                s.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(s):
                    loss = self._body(self.i1, self.i2, self.t)
                torch.cuda.current_stream(dev).wait_stream(s)
                return loss
            self._capture()                                         # capturing does not run the step ...
        self.graph.replay()                                         # ... replaying does
        return self.loss
