run() { tag=$1; shift; env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],3), round(d['roofline']['ms_per_launch']*1e3,1))" ; }
run base X=1
run base X=1
run dring_all ATVS_DECONV_RING_MINVOX=16384
run dring_all ATVS_DECONV_RING_MINVOX=16384
run pg2 ATVS_RING_PG=2 ATVS_RING_R=12
