"""CPU oracle for the ingest + label-aggregation hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, what the reference
(Elmer-Carvalho/Image-Classification-System, mounted at /root/reference during the
build) computes on the one hot path this repository accelerates.  It is the *checker*:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under
``image-classification-system_b200/`` (the product) imports it, and the product has no
CPU fallback.

Parity status ("pinned" = anchored on vectors that do not come from this repository):

* ``sha256_hex`` — the reference's own arithmetic: ``hashlib.sha256(data).hexdigest()``
  (``app/services/webdav_sync.py:59``, ``app/services/activity_api_sync.py:798``,
  ``app/api/routes/images.py:62``).  Pinned on the FIPS 180-4 known answers in
  ``tests/golden/sha256_kat.json``.  ``sha256_restated`` (pure Python, FIPS 180-4
  section 6.2) documents the algorithm the CUDA kernel implements and is itself checked
  against ``hashlib``.
* dedupe / stats, upload lookup, per-user grouping, distinct count, history grouping,
  progress-counter rule — restated from the cited reference lines and pinned on
  ``tests/golden/reference_*.json``, which were produced by RUNNING THE REFERENCE'S OWN
  FUNCTIONS in the build container against stub DB sessions
  (``tests/golden/make_reference_golden.py``).
* thumbnail / preview — **parity unpinned by the reference**: the reference has no
  resize.  The oracle is Pillow (the reference's pinned image library,
  ``requirements.txt:9``) called directly, plus ``resample_restated`` (NumPy
  restatement of Pillow's 8-bit two-pass resampler) which is checked against Pillow.
* label tally / Fleiss kappa — **parity unpinned by the reference**: the reference has
  no cross-annotator tally and no kappa.  The oracle is NumPy ``bincount`` + the textbook
  formula, pinned on the worked example of Fleiss (1971) in
  ``tests/golden/fleiss_1971.json`` (kappa = 0.20993...).

The reference ships no tests, golden vectors or fixtures of its own.
"""
from .hashing import sha256_hex, sha256_digest, sha256_restated  # noqa: F401
from .ingest import (  # noqa: F401
    image_metadata,
    validate_image,
    dedupe_batch,
    process_image_batch,
    buscar_por_hash,
)
from .resize import (  # noqa: F401
    thumbnail_u8,
    preview_f32,
    resample_restated,
    precompute_coeffs,
    scatter_table,
    kernel_model,
)
from .labels import (  # noqa: F401
    label_tally,
    fleiss_partials,
    fleiss_kappa,
    fleiss_kappa_general,
    agreement_hist,
    group_by_image,
    distinct_image_count,
    history_grouping,
    classification_delta,
    encode_label_rows,
)
from .synth import (  # noqa: F401
    synth_image,
    synth_images,
    synth_label_rows,
    synth_duplicate_map,
)
