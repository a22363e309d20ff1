"""Shared fixtures.  ``-m "not gpu"`` runs on the CPU-only build container (oracle vs golden
vectors, host logic, C-ABI symbol check, gloo world_size-2); ``-m gpu`` are the parity tests
proper: the CUDA path, called through the C ABI, against the oracle.  Nothing here reads
/root/reference."""
import base64
import json
import os
import sys
from datetime import datetime

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_ingest():
    return load_json("reference_ingest.json")


@pytest.fixture(scope="session")
def ref_labels():
    return load_json("reference_labels.json")


@pytest.fixture(scope="session")
def sha_kat():
    return load_json("sha256_kat.json")


@pytest.fixture(scope="session")
def fleiss71():
    return load_json("fleiss_1971.json")


@pytest.fixture(scope="session")
def pillow_cases():
    z = np.load(os.path.join(GOLDEN, "pillow_resize.npz"))
    n = sum(1 for k in z.files if k.startswith("in_"))
    return [(z[f"in_{i}"], z[f"out_{i}"]) for i in range(n)]


class Resp:
    def __init__(self, content):
        self.content = content


class FakeNextCloud:
    """Stand-in for NextCloudClient (network is out of scope): serves the golden files and
    raises for the paths the golden scenario marks as failing."""

    def __init__(self, files, failures):
        self.files, self.failures = files, failures

    def get_file(self, path):
        kind = self.failures.get(path)
        if kind == "connection":
            raise ConnectionError("stub connection error")
        if kind == "timeout":
            raise TimeoutError("stub timeout")
        if kind == "other":
            raise ValueError("stub failure")
        return Resp(self.files[path])

    def list_folder(self, folder_path, depth=0):
        name = folder_path.rsplit("/", 1)[-1]
        return [{"is_collection": True, "file_id": "fid-" + (name or "root"), "name": name, "path": folder_path}]

    def filter_images(self, items):
        return items


def ingest_scenario(ref_ingest):
    files = {p: base64.b64decode(d) for p, d in ref_ingest["files"].items()}
    infos = []
    for i in ref_ingest["infos"]:
        d = dict(i)
        d["last_modified"] = datetime.fromisoformat(d["last_modified"]) if d["last_modified"] else None
        infos.append(d)
    return files, infos, FakeNextCloud(files, ref_ingest["failures"])


def dump_rows(rows):
    """Same projection of table ``imagens`` as the golden generator's dump_imagens()."""
    out = {}
    for h, r in rows.items():
        md = r.get("metadados") or {}
        out[h] = {
            "nome_img": r["nome_img"], "caminho_img": r["caminho_img"],
            "existe_no_nextcloud": r["existe_no_nextcloud"], "id_cnj": str(r["id_cnj"]),
            "image_meta": md.get("image"), "nextcloud_meta": md.get("nextcloud"),
            "sync_method": (md.get("sync") or {}).get("sync_method"),
            "first_seen": r["data_proc"] == r["data_sinc"],
        }
    return out
