mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],3), round(d['roofline']['ms_per_launch']*1e3,1))" ; }
run base X=1
run passes4 ATVS_PASSES=4
run passes6 ATVS_PASSES=6
python tools/deconv_probe.py
