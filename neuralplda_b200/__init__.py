"""neuralplda_b200 -- B200-native pairwise trial scoring (NeuralPlda / DPlda hot path).

    from neuralplda_b200 import NeuralPlda, DPlda      # drop-in for utils.models

See DESIGN.md for the kernels and INTEGRATION.md for wiring it under the
reference's driver scripts.
"""
from .models import NeuralPlda, DPlda  # noqa: F401
from . import _lib  # noqa: F401

IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_TC_F8, IMPL_TC_BF16, IMPL_TC_PAIR = (_lib.IMPL_AUTO, _lib.IMPL_SIMT, _lib.IMPL_TC,
                                                                          _lib.IMPL_TC_F8, _lib.IMPL_TC_BF16, _lib.IMPL_TC_PAIR)

__all__ = ["NeuralPlda", "DPlda", "IMPL_AUTO", "IMPL_SIMT", "IMPL_TC", "IMPL_TC_F8", "IMPL_TC_BF16", "IMPL_TC_PAIR"]
