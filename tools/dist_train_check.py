"""2-GPU check of sharded training (BASELINE.json configs[4] in miniature): each rank runs forward + BCE + backward on
its contiguous trial range with the loss accumulators all-reduced (model.process_group), then the parameter gradients
are summed (dist.allreduce_gradients); rank 0 compares with the whole batch on one GPU.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_train_check.py"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import dist as D
from oracle import nplda_oracle as O
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
local = local % torch.cuda.device_count()              # DIST_BACKEND=gloo: several ranks may share one GPU (single-GPU boxes)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("DIST_BACKEND", "nccl") == "gloo":
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=dev)
kp = bench.kaldi_params()
x1, x2, t = O.synth_pairs(20000 + 13, 300, seed=77, mean=kp["mean"])
n = x1.shape[0]
class NCD(bench.NC):
    loss = "crossentropy"; beta = [99.0]
ok = True
for kind in ("nplda", "dplda"):
    def build():
        torch.manual_seed(5)
        if kind == "nplda":
            class C(bench.NC):
                loss = "crossentropy"
            m = npl.NeuralPlda(C).to(dev)
            sd = m.state_dict()
            for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                              ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
                sd[name].copy_(kp[key])
        else:
            m = npl.DPlda(NCD).to(dev)
            m.state_dict()["centering_and_LDA.weight"].copy_(kp["W1"]); m.state_dict()["centering_and_LDA.bias"].copy_(kp["b1"])
            for p in (m.centering_and_LDA.weight, m.centering_and_LDA.bias):
                p.requires_grad_(False)
        return m
    m = build()
    m.process_group = True
    lo, hi = D.shard_range(n, world, rank)
    loss = m.loss(m(x1[lo:hi].to(dev), x2[lo:hi].to(dev)), t[lo:hi].to(dev))
    loss.backward()
    D.allreduce_gradients(m)
    torch.cuda.synchronize()
    if rank == 0:
        w = build()
        lw = w.loss(w(x1.to(dev), x2.to(dev)), t.to(dev))
        lw.backward()
        worst = 0.0
        for (k, p), (_, q) in zip(m.named_parameters(), w.named_parameters()):
            if q.grad is None:
                assert p.grad is None
                continue
            rel = float((p.grad - q.grad).abs().max()) / (float(q.grad.abs().max()) + 1e-30)
            if os.environ.get("DIST_VERBOSE"):
                print(f"   {kind} {k}: |sharded| {float(p.grad.norm()):.6g} |whole| {float(q.grad.norm()):.6g} rel {rel:.2e}")
            worst = max(worst, rel)
        good = abs(loss.item() - lw.item()) <= 1e-6 * abs(lw.item()) + 1e-9 and worst <= 1e-4
        ok = ok and good
        print(f"{kind}: sharded loss {loss.item():.7f} whole {lw.item():.7f}; worst gradient difference / max |grad| = {worst:.2e} -> {'OK' if good else 'MISMATCH'}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
