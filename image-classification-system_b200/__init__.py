"""B200-native ingest + label-aggregation hot path of Elmer-Carvalho/Image-Classification-System.

Import as ``ics_b200`` (see the shim at the repository root).  Layout:

  csrc/ + libb2ingest.so   hand-written sm_100a kernels behind the C ABI of include/b2ingest.h
  _lib.py                  ctypes binding (fails loudly when the library is missing)
  engine.py                device-level entry points (tensors in, tensors out, no sync)
  ingest.py, labels.py     batched host entry points (hash+dedupe+thumbnails; tally+kappa)
  services/, crud/, api/   mirrors of the reference's own functions on the path
  dist.py                  sharding + the two integer collectives (NCCL)
  store.py                 the persistence seam (storage engine is out of scope)
"""
from ._lib import B2Error, LIB_PATH  # noqa: F401
from . import engine, ingest, labels, store, dist  # noqa: F401
from .engine import hash_batch, thumbnails, get_plan, ResizePlan  # noqa: F401
from .ingest import hash_and_dedupe, ingest_batch  # noqa: F401
from .labels import label_tally, fleiss_kappa, fleiss_kappa_general, TallyResult  # noqa: F401

__version__ = "0.1.0"
