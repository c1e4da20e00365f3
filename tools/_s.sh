timeout 300 python -m pytest tests/test_gpu_dense.py -m gpu -q -x 2>&1 | tail -1
HN_DENSE_PROF=1 timeout 100 python tools/dense_case.py 18944 2>&1 | grep "dense prof" | tail -1 | cut -c1-10,60-280
timeout 200 python bench.py --no-cpu --no-delta --no-f32 --steps 2 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['stage_ms_per_step'])"
