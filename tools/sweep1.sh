run() { tag=$1; shift; env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['ms_per_step'],3), round(d['roofline']['ms_per_launch']*1e3,1))" ; }
run base X=1
run base2 X=1
run r16_pf8 ATVS_RING_R=16 ATVS_RING_PF=8
run r12_pf6 ATVS_RING_R=12 ATVS_RING_PF=6
run r8_pf7 ATVS_RING_R=8 ATVS_RING_PF=7
run r6_pf2 ATVS_RING_R=6 ATVS_RING_PF=2
run minb1 ATVS_RING_MINB=1
