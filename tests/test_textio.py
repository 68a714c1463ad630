"""Host-side text I/O: the native trial-list reader and score-file writer against numpy's own
behaviour (np.genfromtxt(dtype='str'), ndarray.astype(str), np.savetxt) and the files the unmodified reference
wrote (tests/golden/c6_*).  CPU only: these entry points do no device work."""
import os

import numpy as np
import pytest

from neuralplda_b200 import textio as T
from neuralplda_b200 import sv_trials_loaders as L
from conftest import GOLDEN


def test_float32_formatting_matches_numpy():
    rng = np.random.default_rng(5)
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 0.1, 1e-4, 9.9999e-5, 1e-5, 1.5e-5, 123456.79, 1e15, 1e16, 1.2345679e16, 3.4e38,
            1e-38, 1e-45, 16777216.0, 0.967479, -0.85290444, np.inf, -np.inf, np.nan, 999999.94, 1e7, 12345678.0]
    arr = np.concatenate([np.float32(vals),
                          rng.standard_normal(200000).astype(np.float32),
                          (rng.standard_normal(100000) * 10.0 ** rng.integers(-12, 20, 100000)).astype(np.float32),
                          rng.integers(0, 2 ** 32, 100000, dtype=np.uint64).astype(np.uint32).view(np.float32)])
    want = arr.astype(str)
    for v, w in zip(arr[:3000], want[:3000]):
        assert T.format_f32(v) == w, (v, T.format_f32(v), w)
    # the rest through the writer (one native pass), compared line by line
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "t.txt")
        with open(src, "w") as f:
            f.write("\n".join("a b" for _ in range(len(arr))) + "\n")
        with T.TrialFile(src) as tf:
            tf.write_scores(os.path.join(d, "o.txt"), arr, 0)
        got = open(os.path.join(d, "o.txt")).read().split("\n")[:-1]
    assert len(got) == len(want)
    bad = [(g, w) for g, w in zip(got, want) if g != w]
    assert not bad, bad[:5]


def test_reader_matches_genfromtxt(tmp_path):
    f = tmp_path / "trials.txt"
    f.write_text("# a comment line\n"
                 "spk1-a  wav/u1.wav\ttgt   # trailing comment\n"
                 "\n"
                 "spk2 dir.v2/u2.tar.gz imp\r\n"
                 "   spk3\tu3 tgt\n"
                 "spk4 .hidden imp")                      # no final newline
    ref = np.genfromtxt(str(f), dtype="str")
    with T.TrialFile(f) as tf:
        assert (tf.rows, tf.cols) == ref.shape
        for r in range(tf.rows):
            assert tf.row(r) == list(ref[r])
        ids = ["u1", "u2.tar", "u3", ".hidden", "spk1-a", "dir.v2/u2.tar"]
        # mode 2 = os.path.splitext(os.path.basename(x))[0], mode 1 = os.path.splitext(x)[0]
        want2 = [ids.index(os.path.splitext(os.path.basename(x))[0]) if os.path.splitext(os.path.basename(x))[0] in ids else -1 for x in ref[:, 1]]
        want1 = [ids.index(os.path.splitext(x)[0]) if os.path.splitext(x)[0] in ids else -1 for x in ref[:, 1]]
        assert list(tf.map_ids(1, ids, mode=T.MODE_BASENAME_SPLITEXT)) == want2
        assert list(tf.map_ids(1, ids, mode=T.MODE_SPLITEXT)) == want1
        assert list(tf.map_ids(0, ids, values=[10, 11, 12, 13, 14, 15])) == [14, -1, -1, -1]
        assert list(tf.map_ids(0, ids, values=[10, 11, 12, 13, 14, 15], first_row=2)) == [-1, -1]
    g = tmp_path / "ragged.txt"
    g.write_text("a b c\nd e\n")
    with pytest.raises(ValueError):
        np.genfromtxt(str(g), dtype="str")
    with pytest.raises(ValueError):
        T.TrialFile(g)
    with pytest.raises(OSError):
        T.TrialFile(tmp_path / "missing.txt")
    e = tmp_path / "empty.txt"
    e.write_text("")
    with T.TrialFile(e) as tf:
        assert tf.rows == 0


def test_labels_parse_like_python_float(tmp_path):
    f = tmp_path / "k.txt"
    labels = ["1", "0", "1.0", "-2.5e-3", "nan", "inf", "tgt", "0x10", "1e", "", "+3"]
    f.write_text("\n".join(f"a b {l}" if l else "a b" for l in labels if l) + "\n")
    labels = [l for l in labels if l]
    with T.TrialFile(f) as tf:
        out, ok = tf.col_float(2)
    for l, v, k in zip(labels, out, ok):
        try:
            want = float(l)
        except ValueError:
            assert not k, l
            continue
        assert k, l
        assert (np.isnan(v) and np.isnan(want)) or np.float32(want) == v


def test_loader_reader_drops_unknown_rows_like_the_reference(tmp_path):
    f = tmp_path / "keys.tsv"
    f.write_text("u1 u2.wav 1\nu1 u3.wav 0\nzz u2.wav 1\nu2 u1.wav notanumber\nu3 u3 0\n")
    id_to_num = {"u1": 7, "u2": 5, "u3": 9, "u2.wav": 1, "u3.wav": 2, "u1.wav": 3}
    ds = L._read_trials(str(f), id_to_num, strip_ext_col2=False)
    assert [tuple(int(v) for v in t[:2]) + (float(t[2]),) for t in ds] == [(7, 1, 1.0), (7, 2, 0.0), (9, 9, 0.0)]
    ds = L._read_trials(str(f), id_to_num, strip_ext_col2=True)        # column 2 loses its extension (:403)
    assert [tuple(int(v) for v in t[:2]) + (float(t[2]),) for t in ds] == [(7, 5, 1.0), (7, 9, 0.0), (9, 9, 0.0)]
    assert ds.tensors[0].dtype.is_floating_point is False and str(ds.tensors[2].dtype) == "torch.float32"
    two = tmp_path / "two.txt"
    two.write_text("u1 u2\nu2 u3\n")
    assert len(L._read_trials(str(two), id_to_num, False)) == 0         # tr[2] raises in the reference: all rows dropped


@pytest.mark.parametrize("trials,golden,sre", [("c6_voices_trials.txt", "c6_voices_scores.txt", False),
                                               ("c6_sre_trials.tsv", "c6_sre_scores.tsv", True)])
def test_writer_reproduces_reference_score_files_byte_for_byte(tmp_path, trials, golden, sre):
    """The golden score files were written by the unmodified reference (tests/golden/make_golden.py).  Feeding the
    scores parsed back from them through the native writer must reproduce the files exactly."""
    ref_bytes = open(os.path.join(GOLDEN, golden), "rb").read()
    ref_tab = np.genfromtxt(os.path.join(GOLDEN, golden), dtype="str")
    scores = ref_tab[1 if sre else 0:, -1].astype(np.float32)
    out = tmp_path / golden
    with T.TrialFile(os.path.join(GOLDEN, trials)) as tf:
        if sre:
            tf.write_scores(out, scores, tf.cols, header_line="\t".join(tf.row(0)) + "\tLLR", first_row=1)
        else:
            tf.write_scores(out, scores, 2)
    assert open(out, "rb").read() == ref_bytes
