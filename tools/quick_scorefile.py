"""End-to-end timing of generate_voices_scores (trial file -> score file) on a large synthetic trial list,
split into parse / id mapping / scoring / formatting, with numpy's own text path timed beside it."""
import os, sys, time, tempfile, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuralplda_b200 as npl
from neuralplda_b200 import scorefile_generator as sg
import bench
n_utts, n_trials = 6500, int(os.environ.get("TRIALS", "2000000"))
rng = np.random.default_rng(3)
kp = bench.kaldi_params()
mega = {f"utt{u:05d}": (kp["mean"].numpy() + rng.standard_normal(512).astype(np.float32)) for u in range(n_utts)}
dev = torch.device("cuda:0")
m = npl.NeuralPlda(bench.NC).to(dev)
sd = m.state_dict()
for name, key in (("centering_and_LDA.weight", "W1"), ("centering_and_LDA.bias", "b1"), ("centering_and_wccn_plda.weight", "W2"),
                  ("centering_and_wccn_plda.bias", "b2"), ("P_sqrt", "P_sqrt"), ("Q", "Q")):
    sd[name].copy_(kp[key])
with tempfile.TemporaryDirectory() as d:
    tf = os.path.join(d, "trials.txt")
    a, b = rng.integers(0, n_utts, n_trials), rng.integers(0, n_utts, n_trials)
    with open(tf, "w") as f:
        f.write("".join(f"utt{x:05d} wav/utt{y:05d}.wav imp\n" for x, y in zip(a, b)))
    out = os.path.join(d, "scores.txt")
    sg.generate_voices_scores(out, tf, mega, m, dev)          # warm (table upload, row table)
    t0 = time.perf_counter(); sg.generate_voices_scores(out, tf, mega, m, dev); t1 = time.perf_counter()
    print(f"generate_voices_scores, {n_trials} trials over {n_utts} utterances: {t1 - t0:.2f} s -> {n_trials / (t1 - t0) / 1e6:.2f} M trials/s, "
          f"file {os.path.getsize(out) / 1e6:.0f} MB")
    t0 = time.perf_counter(); tr = np.genfromtxt(tf, dtype="str")[:, :2]; t1 = time.perf_counter()
    sc = np.genfromtxt(out, dtype="str")[:, 2].astype(np.float32)
    t2 = time.perf_counter(); np.savetxt(os.path.join(d, "ref.txt"), np.c_[tr, sc.astype(str)], fmt="%s", delimiter="\t", comments=""); t3 = time.perf_counter()
    print(f"numpy text path alone (np.genfromtxt {t1 - t0:.2f} s + astype(str)/np.savetxt {t3 - t2:.2f} s), identical bytes: "
          f"{open(os.path.join(d, 'ref.txt'), 'rb').read() == open(out, 'rb').read()}")
