"""mamdr_b200 -- B200-native (sm_100a) implementation of MAMDR's meta-training hot path.

Host side mirrors the reference's wrapper protocol (``MAML`` / ``DomainNegotiation`` / ``MAMDR``
over a ``BaseModel``); the arithmetic runs in ``lib/libmamdr_b200.so`` (C-ABI: ``include/mamdr_b200.h``).
"""
__version__ = "0.1.0"
