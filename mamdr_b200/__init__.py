"""mamdr_b200 -- B200-native (sm_100a) implementation of MAMDR's meta-training hot path.

Host side mirrors the reference's wrapper protocol (``MAML`` / ``DomainNegotiation`` / ``MAMDR``
over a ``BaseModel``); the arithmetic runs in ``lib/libmamdr_b200.so`` (C-ABI: ``include/mamdr_b200.h``).
"""
__version__ = "0.1.0"

# The reference's user-facing class names (run.py:6-15: `from model_zoo import MAML, Reptile, DomainNegotiation`, `from
# model_zoo.mamdr import MAMDR`, `from model_zoo.DeepCTR import DeepCTR`, ..., `from utils import MultiDomainDataset`), resolved
# lazily so that importing the package does not import torch / load the CUDA library.
_EXPORTS = {"MAML": "maml", "DomainNegotiation": "domain_negotiation", "MAMDR": "mamdr", "Reptile": "reptile",
            "MLDG": "mldg", "PCGrad": "pcgrad", "BaseModel": "base_model", "DeepCTR": "deepctr", "Star": "star", "DeepMTLCTR": "deep_mtl_ctr",
            "MultiDomainDataset": "dataset"}
__all__ = sorted(_EXPORTS)


def __getattr__(name):
    if name in _EXPORTS:
        import importlib
        return getattr(importlib.import_module("." + _EXPORTS[name], __name__), name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
