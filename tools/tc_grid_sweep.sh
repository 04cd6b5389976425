#!/bin/bash
# whole-step time against the tiles per CTA of the per-tap TMA kernels when the passes share the SMs (ATVS_TC_MINTILES)
for m in 1 2 3 4 6 8 12; do
  echo -n "ATVS_TC_MINTILES=$m  "
  ATVS_TC_MINTILES=$m python bench.py --no-extras --no-cpu-baseline --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
done
