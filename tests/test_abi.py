"""The C-ABI library loads on a CPU-only box and exports every symbol include/b2ingest.h
declares; without a GPU every compute entry point fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2ingest.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("b2_sha256_batch", "b2_dedupe", "b2_resize_normalize_batch", "b2_label_tally",
              "b2_fleiss_partials", "b2_last_error", "b2_digest_hex", "b2_lookup_sorted"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    import ics_b200
    lib = ctypes.CDLL(ics_b200.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/b2ingest.h but not exported"


def test_binding_covers_header():
    from ics_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    assert _lib.lib.b2_version() == 2


def test_no_cpu_fallback_without_gpu():
    import torch
    import ics_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ics_b200.B2Error) as e:
        ics_b200.hash_batch([b"abc"])
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(ics_b200.B2Error):
        ics_b200.label_tally([0], [0], [1], 1, 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "image-classification-system_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
