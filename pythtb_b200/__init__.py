"""pythtb_b200 — a B200-native (sm_100a) k-mesh engine behind the PythTB API.

``from pythtb_b200 import *`` gives ``tb_model``, ``wf_array`` and ``w90`` with
the call signatures of PythTB 1.8.0; Bloch-Hamiltonian assembly, batched
Hermitian diagonalisation and the Berry-phase / Wilson-loop / Berry-flux
overlap products run in hand-written CUDA kernels behind a C ABI
(``include/tbk.h``).  Model building and file parsing are host code and work
without a GPU; every numerical call requires the built library and a CUDA
device (no CPU fallback).
"""
from .model import tb_model
from .wfarray import wf_array
from .w90 import w90

__version__ = "0.2.0"
__all__ = ["tb_model", "wf_array", "w90"]


def get_backend():
    """Always 'b200' here; ``shim/pythtb`` (``import pythtb`` with PYTHTB_BACKEND=b200|reference, or
    ``pythtb.set_backend``) is the switch for unmodified scripts."""
    return "b200"
