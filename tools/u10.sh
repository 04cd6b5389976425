timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "conv or ring or reasoning or split or attention or cfg2" 2>&1 | tail -3
python tools/steady_probe.py 8,8,1,0 32,8,1,0 2>&1 | grep '"dbg": 0\|"dbg": 4\|"dbg": 3'
echo "early release (dbg 16)"
python - <<'P'
import os, sys, json
sys.path.insert(0, '.')
sys.argv = ['x', '8,8,1,0', '32,8,1,0', '16,16,1,0']
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw
import importlib.util
def timed(fn, iters=3):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters): fn()
    torch.cuda.current_stream().wait_stream(st)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
for cin, cout, D in ((8, 8, 1024), (32, 8, 1024), (16, 16, 1024), (8, 8, 128), (32, 8, 128)):
    x = torch.randn(1, D, 128, 160, cin, device='cuda').to(torch.float16)
    w = torch.randn(3, 3, 3, cin, cout, device='cuda') * 0.05
    for env in ({'ATVS_RING_DEBUG': '0'}, {'ATVS_RING_DEBUG': '16'}, {'ATVS_RING_DEBUG': '0', 'ATVS_RING_MINB': '1'}):
        for k in ('ATVS_RING_DEBUG', 'ATVS_RING_MINB'): os.environ.pop(k, None)
        os.environ.update(env)
        stats = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
        us = timed(lambda: conv3d_raw(x, 'u10_%d_%d' % (cin, cout), w, cout, 1, False, True, stats_buf=stats, raw_dtype=torch.float16))
        print(json.dumps(dict(cin=cin, cout=cout, D=D, env=env, us=round(us, 1), per_cfg2=round(us * 128 / D, 1))), flush=True)
    del x
P
bash tools/tc_grid_sweep.sh
