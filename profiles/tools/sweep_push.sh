#!/bin/bash
# repeatability of the overlap configuration (bulk K1 beside K3): value (M audio-s/s), us/step, kernel us
for i in 1 2 3; do
  timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('overlap', round(d['value']/1e6,2), round(d['ms_per_step']*1e3,1), {k:round(v*1e3,1) for k,v in d['kernel_ms_per_step'].items()}, round(d['e2e']['value']/1e6,3))"
done
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-overlap 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sequential', round(d['value']/1e6,2), round(d['ms_per_step']*1e3,1), round(d['e2e']['value']/1e6,3))"
