#!/usr/bin/env python3
"""Install the UNMODIFIED reference into git-ignored ``baseline/_ref/`` so that the GPU box can run it.

    python baseline/install_reference.py            (also run by __graft_entry__.build() when /root/reference exists)

The reference is plain Python with no setup.py / pyproject.toml, so `pip install --target baseline/_ref
/root/reference` has nothing to build ("neither 'setup.py' nor 'pyproject.toml' found"); the files of the hot path
and their callers are copied verbatim instead.  Nothing is edited: a
sha256 manifest of every copied file is written next to them and `verify()` re-checks it, which is what
`bench.py --impl reference` and the driver tests call before importing anything from there.

`baseline/_ref/` is listed in .gitignore (reference sources never enter this repository's history) and NOT in
.gpurunignore (it travels to the GPU box like the built .so files).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"

FILES = [
    "utils/models.py",
    "utils/sv_trials_loaders.py",
    "utils/scorefile_generator.py",
    "utils/NpldaConf.py",
    "utils/Kaldi2NumpyUtils/kaldiPlda2numpydict.py",
    "utils/adaptive_score_normalization.py",
    "xvector_NeuralPlda_pytorch.py",
    "xvector_DPlda_pytorch.py",
    "xvector_generate_scores.py",
    "conf/sre_config.cfg",
    "conf/voices_config.cfg",
    "conf/voices_config_dplda.cfg",
    "Kaldi_Models/mean.vec",
    "Kaldi_Models/transform.mat",
    "Kaldi_Models/plda",
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def install(src=SRC_DEFAULT, dest=DEST):
    if not os.path.isdir(src):
        raise FileNotFoundError(src)
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = _sha(d)
        assert manifest[rel] == _sha(s)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dest=DEST):
    """True when every file of the manifest is present and unmodified."""
    mf = os.path.join(dest, "MANIFEST.json")
    if not os.path.exists(mf):
        return False
    files = json.load(open(mf))["files"]
    return all(os.path.exists(os.path.join(dest, rel)) and _sha(os.path.join(dest, rel)) == sha
               for rel, sha in files.items()) and set(files) == set(FILES)


if __name__ == "__main__":
    m = install(sys.argv[1] if len(sys.argv) > 1 else SRC_DEFAULT)
    print(f"installed {len(m)} reference files into {DEST}; verify: {verify()}")
