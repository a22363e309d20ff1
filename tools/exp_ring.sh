#!/bin/bash
# Experiments on the ingest ring's concurrency knobs (config 3 stream): hardware connections, hash streams, chunk / group size.
run() { echo "=== $*"; env "$@" python bench.py --only c3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)['configs']['c3']
        print({k: d[k] for k in ('e2e_h2d_gbs_per_gpu','steady_h2d_gbs_per_gpu_min','ring_stalls','listings')}, d['parity']['ok'])
    elif 'rror' in l: print(l.strip())
"; }
run A=1
run CUDA_DEVICE_MAX_CONNECTIONS=8
run CUDA_DEVICE_MAX_CONNECTIONS=1
run B2_RING_HASH_GROUP_MB=1
run B2_RING_HASH_GROUP_MB=1 CUDA_DEVICE_MAX_CONNECTIONS=8
run B2_RING_HASH_GROUP_MB=8192
python bench.py --only c5,e2e,c1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', d['configs']['c5']['e2e_h2d_gbs_per_gpu'], d['configs']['c5']['parity']['ok'], 'e2e', d['e2e']['value'], d['e2e']['h2d_gbs'], d['e2e']['parity'], 'c1', d['configs']['c1']['e2e_images_per_s'])"
