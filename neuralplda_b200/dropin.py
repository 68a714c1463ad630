"""Put this package under the reference's UNCHANGED scripts.

The reference's callers import three modules by name (`/root/reference`):

    xvector_NeuralPlda_pytorch.py:23,28   from utils.sv_trials_loaders import ...; from utils.models import NeuralPlda
    xvector_DPlda_pytorch.py:23,28        the same with DPlda
    utils/NpldaConf.py:10                 from utils.scorefile_generator import generate_voices_scores, generate_sre_scores
    xvector_generate_scores.py:39         pickle.load of an object whose class path is utils.models.NeuralPlda

`install()` registers this package's modules under those names in `sys.modules`, so that the scripts -- and pickles
written by either implementation -- resolve `utils.models`, `utils.sv_trials_loaders` and `utils.scorefile_generator`
to the B200 path while everything else of the reference (`utils.NpldaConf`, the drivers' `train` / `validate` loops)
runs as it is.  Call it before importing the reference's scripts:

    import neuralplda_b200.dropin as dropin
    dropin.install(reference_root)          # reference_root: directory holding utils/ and the xvector_*.py scripts
    import xvector_NeuralPlda_pytorch as drv
    drv.train(nc, model, device, ...)

`uninstall()` restores what was there.  `pickle_as_reference(model, path)` writes a pickle whose class path is
`utils.models.<Class>` (what `SaveModel` of the reference produces, models.py:459-461), loadable by
xvector_generate_scores.py:39 in either world.
"""
from __future__ import annotations

import importlib
import pickle
import sys
import types

_ALIASES = ("models", "sv_trials_loaders", "scorefile_generator")
_saved = {}
_STUBS = ("matplotlib", "matplotlib.pyplot", "kaldi_io")


def install(reference_root=None, stub_missing=True):
    """Alias utils.{models,sv_trials_loaders,scorefile_generator} to this package.  `reference_root` (optional) is put
    on sys.path so that `utils.NpldaConf` and the driver scripts import from the reference tree.  `stub_missing`
    registers empty modules for matplotlib / kaldi_io when they are not installed: the reference imports them at
    module level (models.py:20, sv_trials_loaders.py:18) and never uses them on this path."""
    from . import models, sv_trials_loaders, scorefile_generator
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    if stub_missing:
        for name in _STUBS:
            if name in sys.modules:
                continue
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
        if "matplotlib.pyplot" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    mods = {"models": models, "sv_trials_loaders": sv_trials_loaders, "scorefile_generator": scorefile_generator}
    pkg = sys.modules.get("utils")
    if pkg is None:
        try:
            pkg = importlib.import_module("utils")          # the reference's namespace package, when on sys.path
        except Exception:
            pkg = types.ModuleType("utils")
            pkg.__path__ = []
            sys.modules["utils"] = pkg
            _saved.setdefault("utils", None)
    for short, mod in mods.items():
        full = "utils." + short
        if full not in _saved:
            _saved[full] = sys.modules.get(full)
        sys.modules[full] = mod
        setattr(pkg, short, mod)
    return mods


def uninstall():
    for full, old in list(_saved.items()):
        if old is None:
            sys.modules.pop(full, None)
        else:
            sys.modules[full] = old
        pkg = sys.modules.get("utils")
        short = full.partition(".")[2]
        if pkg is not None and short and hasattr(pkg, short):
            try:
                if old is None:
                    delattr(pkg, short)
                else:
                    setattr(pkg, short, old)
            except Exception:
                pass
        _saved.pop(full)


def pickle_as_reference(model, filename):
    """SaveModel (models.py:459-461) with the reference's class path: the pickle names utils.models.<Class>, as files
    written by the reference do, so xvector_generate_scores.py:39 loads it with or without this package aliased."""
    cls = type(model)
    mod, qual = cls.__module__, cls.__qualname__
    try:
        cls.__module__ = "utils.models"
        # pickle checks that utils.models.<qualname> IS the class: needs the alias in place while dumping
        had = "utils.models" in _saved
        install()
        with open(filename, "wb") as f:
            pickle.dump(model, f)
        if not had:
            uninstall()
    finally:
        cls.__module__ = mod
        cls.__qualname__ = qual
