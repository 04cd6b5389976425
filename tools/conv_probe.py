"""time one bf16 conv layer (C-ABI path): python tools/conv_probe.py cin cout [stride transposed D H W iters]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw
a = [int(v) for v in sys.argv[1:]]
cin, cout = a[0], a[1]
stride = a[2] if len(a) > 2 else 1
tr = bool(a[3]) if len(a) > 3 else False
D, H, W = (a[4], a[5], a[6]) if len(a) > 6 else (128, 128, 160)
iters = a[7] if len(a) > 7 else 10
x = torch.randn(1, D, H, W, cin, device='cuda').to(torch.bfloat16)
w = torch.randn(3, 3, 3, cout, cin, device='cuda') * 0.05 if tr else torch.randn(3, 3, 3, cin, cout, device='cuda') * 0.05
for _ in range(3):
    conv3d_raw(x, 'probe', w, cout, stride, tr, True)
torch.cuda.synchronize()
# CUDA-graph replay: the Python launch path (~50 us per call) must not be what is timed
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(st):
    with torch.cuda.graph(g, stream=st):
        for _ in range(iters):
            conv3d_raw(x, 'probe', w, cout, stride, tr, True)
torch.cuda.current_stream().wait_stream(st)
g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record(); torch.cuda.synchronize()
print(json.dumps(dict(cin=cin, cout=cout, stride=stride, transposed=tr, shape=[D, H, W], dbg=os.environ.get('ATVS_RING_DEBUG', '0'),
                      us=e0.elapsed_time(e1) * 1e3 / iters)))
