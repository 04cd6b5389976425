#!/bin/bash
# whole-step A/B of environment knobs:  bash tools/step_ab.sh "A=1" "B=2 C=3" ...   (first line = defaults)
run() { env "$@" python bench.py --no-extras --no-cpu-baseline --steps 30 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])"; }
echo -n "defaults  "; run X=1
for v in "$@"; do echo -n "$v  "; run $v; done
echo -n "defaults  "; run X=1
