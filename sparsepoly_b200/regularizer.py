"""Host-callable proximal operators with the reference's class surface (sparsepoly/regularizer/*.py): the
psgd-side protocol `init_cache_psgd(degree, n_features, n_components)` + `prox(P, strength, degree)` on a
numpy array P [n_features, n_components], computed on the device by sp_prox (csrc/psgd.cu: l1.py:50-51,
l21.py:43-48, squaredl12.py:66-78, squaredl21.py:63-74, utils.py:26-70).  The coordinate-wise protocols
(prox_cd / prox_bcd and their caches) live inside the sweep kernels and have no host-side object."""
import numpy as np
import torch

from . import solvers
from .dataset import _device


class _Prox:
    _name = None

    def __init__(self, transpose=None):
        self.transpose = transpose

    def init_cache_psgd(self, degree, n_features, n_components):
        pass

    def prox(self, P, strength, degree=2):
        """In place: P <- prox_{strength * Omega}(P)."""
        dev = _device()
        Pd = torch.from_numpy(np.ascontiguousarray(P, dtype=np.float64)).to(dev)
        solvers.prox(Pd, self._name, float(strength), solvers.prox_work(Pd.shape[0], Pd.shape[1], dev))
        P[...] = Pd.cpu().numpy()


class L1(_Prox):
    _name = "l1"


class L21(_Prox):
    _name = "l21"


class SquaredL12(_Prox):
    _name = "squaredl12"


class SquaredL21(_Prox):
    _name = "squaredl21"


REGULARIZATION = {"l1": L1, "l21": L21, "squaredl12": SquaredL12, "squaredl21": SquaredL21}
