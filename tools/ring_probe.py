"""time one stride-1 conv layer (ring kernel) with the debug knobs of conv_ring.cu"""
import sys, os, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    import atvsnet_b200 as A
    from atvsnet_b200.network import conv3d_raw
    cin, cout = int(sys.argv[2]), int(sys.argv[3])
    D, H, W = 128, 128, 160
    x = torch.randn(1, D, H, W, cin, device='cuda').to(torch.bfloat16)
    w = torch.randn(3, 3, 3, cin, cout, device='cuda') * 0.05
    for _ in range(3):
        conv3d_raw(x, 'probe', w, cout, 1, False, True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        conv3d_raw(x, 'probe', w, cout, 1, False, True)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps(dict(cin=cin, cout=cout, dbg=os.environ.get('ATVS_RING_DEBUG', '0'), us=e0.elapsed_time(e1) * 100)))
else:
    for cin, cout in ((8, 8),):
        for dbg in (0, 7, 7 + 8, 7 + 16, 7 + 32, 7 + 8 + 16 + 32, 8, 16):
            env = dict(os.environ, ATVS_RING_DEBUG=str(dbg))
            r = subprocess.run([sys.executable, __file__, 'child', str(cin), str(cout)], env=env, capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
