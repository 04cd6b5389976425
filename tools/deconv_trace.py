"""one launch of the deconv plane-ring kernel on the cfg2 shape with the trace library: per-role timelines of CTA 0"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
from atvsnet_b200.network import conv3d_raw
cin, cout, shape = 16, 8, (1, 64, 64, 80)
x = torch.randn(shape + (cin,), device='cuda').half()
w = torch.randn(3, 3, 3, cout, cin, device='cuda') * 0.05
stats = torch.zeros(128, dtype=torch.float64, device='cuda')
os.environ['ATVS_RING_TRACE_QUIET'] = '1'
for i in range(3):
    conv3d_raw(x, 'probe', w, cout, 2, True, True, stats, raw_dtype=torch.float16)
torch.cuda.synchronize()
print("=====LAST", flush=True)
