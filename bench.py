#!/usr/bin/env python3
"""Benchmark of the pairwise trial-scoring hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic trial pairs:
the fused score kernel K1 over all pairs, the loss-accumulator kernel K2 over
the scores, the all-reduce of the 4K+4 fp64 accumulators (N > 1) and the
finalisation of softCdet / BCE / Cdet.  Workload at every N: BASELINE.json
configs[1], 1M synthetic 512-d trial pairs PER GPU (weak scaling: the trial
list is sharded by contiguous ranges, per-GPU work fixed), NeuralPlda with
sre_config.cfg dims (512 -> 170 -> 170, K = 2 betas), Kaldi-init parameters.

Prints ONE JSON line (rank 0).  `value` = pairs/s with inputs resident in HBM;
`e2e` = the same through the public module API from pinned HOST buffers (H2D of
the pairs and labels and the D2H of the loss inside the timed region);
`roofline` = K1's algorithmic bytes (4100 B/pair) / its CUDA-event duration
against the measured HBM copy bandwidth; `cpu_baseline` = the CPU oracle port
of the reference forward timed on this box's host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAIRS_PER_GPU = 1_000_000
D_IN, D1, D2 = 512, 170, 170
BETAS = [99.0, 199.0]
ALPHA = 15.0
BYTES_PER_PAIR = 2 * D_IN * 4 + 4          # 4100 B algorithmic per pair
REF_CHUNK = 102_400                         # scorefile_generator.py:22 default scoring batch


class NC:
    xvector_dim, layer1_LDA_dim, layer2_PLDA_spkfactor_dim = D_IN, D1, D2
    alpha, device, beta, loss = ALPHA, "cpu", BETAS, "SoftCdet"


def kaldi_params():
    z = np.load(os.path.join(ROOT, "tests", "golden", "kaldi_init_params.npz"))
    return {k: torch.from_numpy(z[k].copy()) for k in z.files}


def measured_traffic():
    """DRAM bytes per K1 launch from the committed `ncu --set full` capture of this workload (profiles/)."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        t = json.load(open(p))
        if t.get("pairs") == PAIRS_PER_GPU:
            return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def synth_on_device(n, seed, mean, dev):
    """SURVEY 8d generator (x = mu + c_spk + 0.7 eps, 10% targets), built on the GPU."""
    g = torch.Generator(device=dev).manual_seed(seed)
    nspk = 2000
    spk = torch.randn(nspk, D_IN, generator=g, device=dev)
    s1 = torch.randint(0, nspk, (n,), generator=g, device=dev)
    tgt = torch.rand(n, generator=g, device=dev) < 0.1
    s2 = torch.where(tgt, s1, (s1 + torch.randint(1, nspk, (n,), generator=g, device=dev)) % nspk)
    x1 = torch.empty(n, D_IN, device=dev)
    x2 = torch.empty(n, D_IN, device=dev)
    step = 131072
    for a in range(0, n, step):                      # bounded temporaries
        b = min(n, a + step)
        x1[a:b] = mean + spk[s1[a:b]] + 0.7 * torch.randn(b - a, D_IN, generator=g, device=dev)
        x2[a:b] = mean + spk[s2[a:b]] + 0.7 * torch.randn(b - a, D_IN, generator=g, device=dev)
    return x1, x2, tgt.float()


def cpu_forward_timing(kp, n_chunks, repeats, threads=None, min_seconds=0.0):
    """The reference's CPU forward restated by the oracle, in chunks of 102,400
    pairs as scorefile_generator.py scores, under no_grad, all host threads.
    Repeats passes of `n_chunks` chunks until `repeats` passes are done AND at least
    `min_seconds` of CPU work has been timed; returns the best pass."""
    from oracle import nplda_oracle as O
    if threads:
        torch.set_num_threads(threads)
    x1, x2, _ = O.synth_pairs(REF_CHUNK, 2000, seed=1002, mean=kp["mean"])
    args = (kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"])
    with torch.no_grad():
        O.nplda_score(x1[:4096], x2[:4096], *args)   # warm the thread pool
        best, total, passes = None, 0.0, 0
        while passes < repeats or total < min_seconds:
            t0 = time.perf_counter()
            for _ in range(n_chunks):
                s = O.nplda_score(x1, x2, *args)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            total += dt
            passes += 1
            if total > 30.0:
                break
    cpu_forward_timing.last = {"passes": passes, "total_s": total}
    return n_chunks * REF_CHUNK / best, best, float(s[0])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kp = kaldi_params()
    threads = os.cpu_count()
    n_chunks = 10                                    # one step = the configs[1] workload: 1,024,000 pairs in scoring-loop chunks
    torch.set_num_threads(threads)
    for _ in range(args.warmup):
        cpu_forward_timing(kp, 1, 1, threads)
    t0 = time.perf_counter()
    rates = [cpu_forward_timing(kp, n_chunks, 1, threads)[0] for _ in range(args.steps)]
    wall = time.perf_counter() - t0
    value = float(np.median(rates))
    sample = f"{n_chunks} x {REF_CHUNK}-pair chunks per step (same synthetic chunk re-scored), oracle port of NeuralPlda.forward, no_grad"
    out = {
        "impl": "reference", "metric": "trial-pairs scored/sec (512-d xvec)", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * n_chunks * REF_CHUNK / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: 1M synthetic 512-d trial pairs per GPU, NeuralPlda 512-170-170 forward",
                   "sample_pairs_per_step": n_chunks * REF_CHUNK},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample,
                         "torch_threads": torch.get_num_threads(), "wall_s": wall},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_ours(args):
    # libraries (NCCL's version banner) write to the C-level stdout: keep fd 1 for the ONE JSON line
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _run_ours(args, saved_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)


def _emit(saved_stdout, obj):
    sys.stdout.flush()
    os.write(saved_stdout, (json.dumps(obj) + "\n").encode())


def _run_ours(args, saved_stdout):
    import torch.distributed as dist
    import neuralplda_b200 as npl
    from neuralplda_b200 import _lib, functional as F_

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the scoring path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    group = True if world > 1 else None

    kp = kaldi_params()
    model = npl.NeuralPlda(NC).to(dev)
    sd = model.state_dict()
    for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"),
                      ("centering_and_wccn_plda.weight", "W2"), ("centering_and_wccn_plda.bias", "b2"),
                      ("P_sqrt", "P_sqrt"), ("Q", "Q")):
        sd[name].copy_(kp[key])
    model.impl = {"auto": npl.IMPL_AUTO, "simt": npl.IMPL_SIMT, "tc": npl.IMPL_TC, "f8": npl.IMPL_TC_F8}[args.kernel]
    model.process_group = group
    model.eval()

    n = args.pairs
    x1, x2, t = synth_on_device(n, 1002 + rank, kp["mean"].to(dev), dev)
    lib = _lib.lib()
    pack = model.packed.get("nplda", model._params(), D_IN, D1, D2)
    scores = torch.empty(n, device=dev)
    thresholds = torch.cat([model.threshold[b].detach() for b in BETAS])
    thx = model.threshold_Xent.detach()
    acc = torch.zeros(4 * len(BETAS) + 4, dtype=torch.float64, device=dev)
    out = torch.empty(4, device=dev)
    betas_arr = _lib.betas_array(BETAS)
    stream = torch.cuda.current_stream()

    def k1():
        _lib.check(lib.nplda_score_fwd(_lib.ptr(x1), _lib.ptr(x2), n, D_IN, D1, D2, _lib.ptr(pack), _lib.ptr(scores),
                                       model.impl, _lib.stream_ptr()), "nplda_score_fwd")

    def step(ev=None):
        if ev:
            ev[0].record(stream)
        k1()
        if ev:
            ev[1].record(stream)
        acc.zero_()
        _lib.check(lib.nplda_loss_accum(_lib.ptr(scores), _lib.ptr(t), n, _lib.ptr(thresholds), len(BETAS), ALPHA,
                                        _lib.ptr(thx), _lib.ptr(acc), _lib.stream_ptr()), "nplda_loss_accum")
        if world > 1:
            dist.all_reduce(acc)
        _lib.check(lib.nplda_loss_finalize(_lib.ptr(acc), betas_arr, len(BETAS), _lib.ptr(out), _lib.stream_ptr()),
                   "nplda_loss_finalize")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    k1_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(k1_events[i])
    e1.record(stream)
    barrier()
    launches = _lib.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in k1_events]))
    sampler.stop_flag = True
    sampler.join(timeout=1.0)
    loss_vals = out.tolist()

    # ---- end to end through the module API from pinned host buffers ----------------
    e2e_steps = max(3, min(args.steps, 10))
    if args.skip_e2e:                                # profiling runs (ncu) only: never a bench line
        if rank == 0:
            _emit(saved_stdout, {"profiling_only": True, "k1_ms": k1_ms, "ms_per_step": total_ms / args.steps})
        if world > 1:
            dist.destroy_process_group()
        return
    h1 = torch.empty(n, D_IN, pin_memory=True); h1.copy_(x1)
    h2 = torch.empty(n, D_IN, pin_memory=True); h2.copy_(x2)
    ht = torch.empty(n, pin_memory=True); ht.copy_(t)
    d1_, d2_, dt_ = torch.empty_like(x1), torch.empty_like(x2), torch.empty_like(t)

    def e2e_step():
        d1_.copy_(h1, non_blocking=True); d2_.copy_(h2, non_blocking=True); dt_.copy_(ht, non_blocking=True)
        with torch.no_grad():
            s = model(d1_, d2_)                      # NeuralPlda.forward, the reference-facing call
            return model.loss(s, dt_).item()         # softCdet incl. all-reduce; D2H of the scalar

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_loss = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0

    tm = torch.tensor([total_ms, k1_ms, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, k1_ms, e2e_s = tm.tolist()

    if rank == 0:
        peak, peak_src = measured_peak()
        value = world * n * args.steps / (total_ms * 1e-3)
        achieved = n * BYTES_PER_PAIR / (k1_ms * 1e-3) / 1e9
        cpu_rate, cpu_s, _ = (cpu_forward_timing(kp, 10, 3, os.cpu_count(), min_seconds=12.0)
                              if not args.no_cpu_baseline else (None, 0, 0))
        tl = trial_list_leg(model, kp, dev) if (not args.skip_trial_list and world == 1) else None
        tr = train_leg(kp, dev, x1, x2, t) if (not args.skip_trial_list and world == 1) else None
        res = {
            "metric": "trial-pairs scored/sec (512-d xvec)", "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: 1M synthetic 512-d trial pairs per GPU (materialised [N,512] x2 fp32), "
                                   "NeuralPlda 512-170-170 Kaldi-init, score + softCdet/BCE/Cdet accumulators (K=2)",
                       "pairs_per_gpu": n, "kernel": args.kernel, "sharding": "contiguous trial-list ranges, "
                       "one all-reduce of 12 fp64 accumulators per step" if world > 1 else "single GPU",
                       "l2": "inputs (4.1 GB per step) larger than L2 (126 MB); no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(), "kernel": "K1 fused score kernel", "k1_ms": k1_ms,
                         "bytes_per_pair": BYTES_PER_PAIR, "peak_source": peak_src},
            "e2e": {"value": world * n * e2e_steps / e2e_s, "unit": "pairs/s",
                    "h2d_bytes_per_step": n * (2 * D_IN * 4 + 4), "d2h_bytes_per_step": 4, "steps": e2e_steps,
                    "api": "NeuralPlda.forward + .loss(...).item() on pinned host inputs"},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "loss": {"softcdet": loss_vals[0], "bce": loss_vals[1], "cdet": loss_vals[2], "e2e_softcdet": e2e_loss},
        }
        if cpu_rate is not None:
            res["cpu_baseline"] = {"value": cpu_rate, "unit": "pairs/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"passes of 10 x {REF_CHUNK}-pair chunks (1,024,000 pairs, same synthetic chunk "
                                             f"re-scored) for {cpu_forward_timing.last['total_s']:.1f} s of CPU work "
                                             f"({cpu_forward_timing.last['passes']} passes), best pass {cpu_s:.2f} s; oracle "
                                             f"port of NeuralPlda.forward under no_grad",
                                   "torch_threads": torch.get_num_threads()}
        if tl is not None:
            res["trial_list"] = tl
        if tr is not None:
            res["train_step"] = tr
        _emit(saved_stdout, res)
    if world > 1:
        dist.destroy_process_group()


def train_leg(kp, dev, x1, x2, t):
    """Extra, separately-labelled measurement: one training step (forward + BCE loss + backward into .grad) through the
    module API on the resident 1M-pair batch, for NeuralPlda and for DPlda with the LDA frozen
    (BASELINE.json configs[4]; its per-GPU share at 8 GPUs is 1.25M pairs).  CUDA events, 5 steps after one warm-up."""
    import neuralplda_b200 as npl

    class NCX(NC):
        loss = "crossentropy"

    class NCDP(NC):
        loss = "crossentropy"
        beta = [99.0]

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    n = x1.shape[0]
    out = {"pairs": n, "loss": "crossentropy", "api": "model(x1, x2) -> model.loss(...) -> .backward()"}
    m = npl.NeuralPlda(NCX).to(dev)
    sd = m.state_dict()
    for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"),
                      ("centering_and_wccn_plda.weight", "W2"), ("centering_and_wccn_plda.bias", "b2"),
                      ("P_sqrt", "P_sqrt"), ("Q", "Q")):
        sd[name].copy_(kp[key])

    def nstep():
        m.zero_grad(set_to_none=True)
        m.loss(m(x1, x2), t).backward()

    ms = timeit(nstep)
    out["nplda"] = {"ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "pairs/s"}
    d = npl.DPlda(NCDP).to(dev)
    sd = d.state_dict()
    sd["centering_and_LDA.weight"].copy_(kp["W1"]); sd["centering_and_LDA.bias"].copy_(kp["b1"])
    for p_ in (d.centering_and_LDA.weight, d.centering_and_LDA.bias):
        p_.requires_grad_(False)                               # xvector_DPlda_pytorch.py:140-147

    def dstep():
        d.zero_grad(set_to_none=True)
        d.loss(d(x1, x2), t).backward()

    ms = timeit(dstep)
    out["dplda_lda_frozen"] = {"ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "pairs/s"}
    with torch.no_grad():
        ms = timeit(lambda: d(x1, x2))
    out["dplda_forward"] = {"ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "pairs/s"}
    return out


def trial_list_leg(model, kp, dev):
    """Extra, separately-labelled measurement (never mixed into `value` / `roofline`): BASELINE.json
    configs[2] in the INDEXED layout the reference's scoring loop actually has (a table of unique
    x-vectors + a trial list, scorefile_generator.py:29-36): 10 M trials = 2500 enrol x 4000 test over 6500
    utterances, every utterance transformed once (nplda_table_prepare) and each trial scored as
    r[i] + r[j] + A[i].B[j] (nplda_score_pairs)."""
    from oracle import nplda_oracle as O
    table, i1, i2, _ = O.synth_grid(2500, 4000, 500, seed=1003, mean=kp["mean"])
    t, a, b = table.to(dev), i1.to(dev), i2.to(dev)
    n = a.numel()
    stream = torch.cuda.current_stream()

    def full():
        model.packed.rowtab_key = None               # table prepare inside every step
        return model.forward_indexed(t, a, b)[0]

    s = full()
    sub = torch.arange(0, n, 997)
    ref = O.nplda_score(table[i1[sub]], table[i2[sub]], kp["W1"], kp["b1"], kp["W2"], kp["b2"], kp["P_sqrt"], kp["Q"]).double()
    got = s[sub.to(dev)].cpu().double()
    worst = float(((got - ref).abs() / (1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt()))).max())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps):
        full()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ha, hb, hs = i1.pin_memory(), i2.pin_memory(), torch.empty(n, pin_memory=True)

    def e2e():
        model.packed.rowtab_key = None
        s_ = model.forward_indexed(t, ha.to(dev, non_blocking=True), hb.to(dev, non_blocking=True))[0]
        hs.copy_(s_, non_blocking=True)
        torch.cuda.synchronize()

    e2e()
    t0 = time.perf_counter()
    for _ in range(5):
        e2e()
    dt = (time.perf_counter() - t0) / 5
    # the same 10M trials as an enrol x test GRID (nplda_score_grid): no index pair per trial, one fp32 product
    er, tr = torch.arange(2500, device=dev), torch.arange(2500, 6500, device=dev)

    def grid_full():
        model.packed.rowtab_key = None
        return model.forward_grid(t, er, tr)[0]

    sg = grid_full()
    gotg = sg.flatten()[sub.to(dev)].cpu().double()
    worst_g = float(((gotg - ref).abs() / (1e-4 * torch.maximum(ref.abs(), ref.pow(2).mean().sqrt()))).max())
    gms = {}
    for name, fn in (("with_table_prepare", grid_full), ("rows_cached", lambda: model.forward_grid(t, er, tr)[0])):
        fn()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(20):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        gms[name] = e0.elapsed_time(e1) / 20
    hg = torch.empty(2500, 4000, pin_memory=True)

    def grid_e2e():
        model.packed.rowtab_key = None
        hg.copy_(model.forward_grid(t, er, tr)[0], non_blocking=True)
        torch.cuda.synchronize()

    grid_e2e()
    t0 = time.perf_counter()
    for _ in range(5):
        grid_e2e()
    gdt = (time.perf_counter() - t0) / 5
    grid = {"workload": "configs[2] as an enrol x test grid: 2500 x 4000 over 6500 x-vectors, table prepare + one "
                        "[2500,176] x [176,4000] fp32 grid product (nplda_score_grid) every step",
            "value": n / (gms["with_table_prepare"] * 1e-3), "unit": "trials/s", "ms_per_step": gms["with_table_prepare"],
            "ms_rows_cached": gms["rows_cached"], "value_rows_cached": n / (gms["rows_cached"] * 1e-3),
            "fp32_tflops_rows_cached": 2 * 176 * n / (gms["rows_cached"] * 1e-3) / 1e12,
            "e2e": {"value": n / gdt, "unit": "trials/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * n,
                    "api": "NeuralPlda.forward_grid, scores copied back to pinned host memory"},
            "bytes_per_trial": {"hbm_score": 4}, "parity_worst_over_bound_strided_sample": worst_g}
    return {"grid": grid, "workload": "configs[2] indexed: 10M trials = 2500 x 4000 grid over 6500 x-vectors (table resident in HBM), "
                        "table prepare + pair scoring every step",
            "value": n / (ms * 1e-3), "unit": "trials/s", "ms_per_step": ms,
            "e2e": {"value": n / dt, "unit": "trials/s", "h2d_bytes_per_step": 16 * n, "d2h_bytes_per_step": 4 * n,
                    "api": "NeuralPlda.forward_indexed on pinned host index tensors, scores copied back to pinned host memory"},
            "bytes_per_trial": {"hbm_indices_and_score": 20, "l2_row_gather": 1408},
            "parity_worst_over_bound_strided_sample": worst}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tc", "f8"])
    ap.add_argument("--skip-trial-list", action="store_true", help="skip the extra indexed trial-list measurement")
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs only; prints no bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
