"""B200-native ingest + label-aggregation hot path of Elmer-Carvalho/Image-Classification-System.

Import as ``ics_b200`` (see the shim at the repository root).  Layout:

  csrc/ + libb2ingest.so   hand-written sm_100a kernels behind the C ABI of include/b2ingest.h
  _lib.py                  ctypes binding (fails loudly when the library is missing)
  hostapi.py               host-pointer layer: ctypes + NumPy only (what the reference service needs)
  ingest.py, labels.py     batched host entry points (hash+dedupe+thumbnails; tally+kappa)
  services/, crud/, api/   mirrors of the reference's own functions on the path
  engine.py, pipeline.py   device-pointer layer (PyTorch tensors as buffers) and the streaming wrapper
  dist.py                  sharding + the two integer collectives (NCCL)
  store.py                 the persistence seam (storage engine is out of scope)

Importing the package does not import PyTorch: the device-pointer layer (``engine``, ``pipeline``, ``dist``,
``get_plan``, ``ResizePlan``) is loaded on first access.
"""
import importlib as _importlib

from ._lib import B2Error, LIB_PATH  # noqa: F401
from . import hostapi, ingest, labels, store  # noqa: F401
from .hostapi import hash_batch, thumbnails  # noqa: F401
from .ingest import hash_and_dedupe, ingest_batch  # noqa: F401
from .labels import label_tally, fleiss_kappa, fleiss_kappa_general, fleiss_kappa_from_hist, TallyResult  # noqa: F401

__version__ = "0.1.0"

_LAZY_MODULES = ("engine", "pipeline", "dist")
_LAZY_NAMES = {"get_plan": "engine", "ResizePlan": "engine"}


def __getattr__(name):
    if name in _LAZY_MODULES:
        return _importlib.import_module(f"{__name__}.{name}")
    if name in _LAZY_NAMES:
        return getattr(_importlib.import_module(f"{__name__}.{_LAZY_NAMES[name]}"), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
