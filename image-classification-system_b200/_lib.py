"""ctypes binding of libb2ingest.so (C ABI declared in include/b2ingest.h).

The library is the product; there is NO CPU fallback.  If the shared object is missing the
import of this module raises, and every compute entry point raises ``B2Error`` when there is
no Blackwell GPU.  ctypes releases the GIL for the duration of each call, so the service's
threads can call concurrently.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2_LIB_PATH") or os.path.join(HERE, "libb2ingest.so")   # B2_LIB_PATH: A/B builds of the library

B2_OK = 0
B2_ERR_BAD_ARG = -1
B2_ERR_CUDA = -2
B2_ERR_NOT_SORTED = -3
B2_ERR_NO_DEVICE = -4
B2_ERR_WORKSPACE = -5
B2_ERR_NCCL = -6

B2_TALLY_SORTED = 1
B2_PARTIALS_EXTRA = 7
B2_AGREE_BINS = 1024
B2_COMM_ID_BYTES = 128


class B2Error(RuntimeError):
    """A libb2ingest entry point returned a negative status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libb2ingest error {code}: {message}")
        self.code = code


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        # a source-only checkout: compile the library in place if the CUDA toolchain is here (nvcc cross-compiles
        # sm_100a without a GPU); otherwise fail loudly — there is no CPU fallback for this path
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_b2_build", os.path.join(HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        except Exception as e:  # noqa: BLE001 - whatever went wrong, the message below is what the user needs
            raise ImportError(
                f"{LIB_PATH} not found and could not be built ({e}): build it with "
                "`python image-classification-system_b200/build.py` (or __graft_entry__.build()).  "
                "There is no CPU fallback for this path.") from e
    return C.CDLL(LIB_PATH)


lib = _load()

_u8p, _u32p, _u64p = C.c_void_p, C.c_void_p, C.c_void_p   # device pointers travel as integers
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/b2ingest.h one to one
SIGNATURES = {
    "b2_version": (C.c_int, []),
    "b2_last_error": (C.c_char_p, []),
    "b2_init": (C.c_int, [C.c_int]),
    "b2_device_sm_count": (C.c_int, [C.c_int, C.POINTER(C.c_int)]),
    "b2_shutdown": (C.c_int, []),
    "b2_sha256_batch": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp]),
    "b2_digest_hex": (C.c_int, [_vp, C.c_uint32, _vp, _vp]),
    "b2_dedupe_workspace_bytes": (C.c_uint64, [C.c_uint32]),
    "b2_dedupe": (C.c_int, [_vp, _vp, _vp, C.c_uint32, _vp, C.c_uint64, _vp, _vp, _vp, _vp, _vp, C.c_uint64, _vp]),
    "b2_lookup_sorted": (C.c_int, [_vp, C.c_uint32, _vp, C.c_uint64, _vp, _vp]),
    "b2_resize_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "b2_resize_plan_destroy": (C.c_int, [_vp]),
    "b2_resize_plan_taps": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int), _vp, _vp, C.c_uint64]),
    "b2_resize_normalize_batch": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, _vp, _vp,
                                            C.POINTER(C.c_float), C.POINTER(C.c_float), _vp]),
    "b2_label_tally": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 _vp, _vp, _vp, _vp]),
    "b2_label_tally_status": (C.c_int, [_vp, C.c_uint32, C.c_uint64]),
    "b2_fleiss_workspace_bytes": (C.c_uint64, [C.c_uint32]),
    "b2_fleiss_partials": (C.c_int, [_vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp, _vp, C.c_uint64, _vp]),
    "b2_distinct_images_per_annotator": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, _vp, _vp]),
    "b2_encode_label_rows": (C.c_int, [_vp, _vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_uint32, _vp, _vp, _vp, _vp, _vp]),
    "b2_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_uint64]),
    "b2_host_free": (C.c_int, [_vp]),
    "b2_ingest_ring_create": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "b2_ingest_ring_destroy": (C.c_int, [_vp]),
    "b2_ingest_ring_submit": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_uint32, _vp, C.c_uint64,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_uint64)]),
    "b2_ingest_ring_wait": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "b2_ingest_ring_poll": (C.c_int, [_vp, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]),
    "b2_ingest_ring_stats": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_uint64)]),
    "b2_sha256_host": (C.c_int, [C.c_int, _vp, _vp, C.c_uint32, _vp, _vp]),
    "b2_dedupe_host": (C.c_int, [C.c_int, _vp, _vp, C.c_uint32, _vp, C.c_uint64, _vp, _vp, _vp, _vp]),
    "b2_thumbnails_host": (C.c_int, [C.c_int, _vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp,
                                     C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "b2_label_tally_host": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                      _vp, _vp, _vp]),
    "b2_distinct_images_host": (C.c_int, [C.c_int, _vp, _vp, _vp, C.c_uint64, C.c_uint32, _vp]),
    "b2_comm_unique_id": (C.c_int, [_vp]),
    "b2_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_int, _vp, C.POINTER(_vp)]),
    "b2_comm_destroy": (C.c_int, [_vp]),
    "b2_comm_info": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2_allgather_digests": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp]),
    "b2_allreduce_i64": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "b2_comm_enable_peer_reduce": (C.c_int, [_vp, C.c_uint32]),
    "b2_comm_peer_status": (C.c_int, [_vp, C.POINTER(C.c_int)]),
    "b2_peer_allreduce_i64": (C.c_int, [_vp, _vp, C.c_uint32, _vp]),
    "b2_label_tally_reduce": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                        _vp, _vp, _vp]),
    "b2_dedupe_global_workspace_bytes": (C.c_uint64, [C.c_uint32, C.c_uint32]),
    "b2_dedupe_global": (C.c_int, [_vp, _vp, _vp, _vp, C.c_uint32, C.c_uint32, _vp, C.c_uint64, _vp, _vp, _vp, _vp,
                                   _vp, C.c_uint64, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.b2_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != B2_OK:
        raise B2Error(rc, last_error())
