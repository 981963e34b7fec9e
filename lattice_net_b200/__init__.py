"""lattice_net_b200 -- B200-native backend for the permutohedral-lattice hot path of LatticeNet.

`Lattice` / `HashTable` mirror the reference's `latticenet` pybind module, `lattice_funcs` /
`lattice_modules` / `lattice_wrapper` mirror `latticenet_py.lattice.*`; all device work is done by
hand-written sm_100a kernels behind the C ABI of include/lattice_b200.h.
"""
from .params import EvalParams, ModelParams, TrainParams   # noqa: F401
from .lattice import HashTable, Lattice, set_conv_precision   # noqa: F401

__all__ = ["Lattice", "HashTable", "TrainParams", "ModelParams", "EvalParams", "set_conv_precision"]
