#!/bin/bash
# A/B of the pair-column resize kernel against the band kernel over the config shapes (same box, same call).
python -m pytest tests/test_resize_gpu.py tests/test_pipeline_gpu.py -q -m gpu -x 2>&1 | tail -3
echo "== pairs"; python tools/resize_shapes.py 2>&1 | tail -7 | cut -c1-75
echo "== band kernel only"; B2_RESIZE_NO_PAIRS=1 python tools/resize_shapes.py 2>&1 | tail -7 | cut -c1-75
for kb in 46 52 60; do echo "1080p pair smem ${kb} KB"; B2_RESIZE_PAIR_SMEM_KB=$kb python tools/profile_kernels.py resize; done
echo "3 stages 54 KB"; B2_RESIZE_STAGES=3 B2_RESIZE_PAIR_SMEM_KB=54 python tools/profile_kernels.py resize
