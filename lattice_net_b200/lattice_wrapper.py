"""`LatticeWrapper`: lets a `Lattice` handle travel through `torch.autograd.Function.apply`, whose
outputs must be tensors (reference: /root/reference/latticenet_py/lattice/lattice_wrapper.py:12-17)."""
import torch


class LatticeWrapper(torch.Tensor):
    @staticmethod
    def wrap(lattice):
        carrier = LatticeWrapper()
        carrier.lattice = lattice
        return carrier
