python -m pytest tests/test_gpu_parity.py -m "gpu and not slow" -q -x -k "ring or conv3d or pipeline or cfg2_sized or properties" 2>&1 | tail -2
echo "ring 8->8 mt4"; python tools/conv_probe.py 8 8 1 0 128 128 160 10
for zs in 16 22 32 43 64; do echo "ring 8->8 mt4 zs=$zs"; ATVS_RING_ZS=$zs python tools/conv_probe.py 8 8 1 0 128 128 160 10; done
echo "ring 8->8 mt2"; ATVS_RING_MT=2 python tools/conv_probe.py 8 8 1 0 128 128 160 10
echo "ring 8->8 mt4 minb1"; ATVS_RING_MINB=1 python tools/conv_probe.py 8 8 1 0 128 128 160 10
for v in "X=1" "ATVS_RING_MT=2" "ATVS_RING_ZS_8_8=22" "ATVS_RING_ZS_8_8=32" "ATVS_RING_ZS_8_8=64"; do env $v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v step', round(d['ms_per_step'],3), round(d['roofline']['ms_per_launch']*1e3,1))"; done
