"""one FEM forward at 512x640 between cudaProfilerStart/Stop (ncu --profile-from-start off)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import atvsnet_b200 as A
A.variables.load_weights(A.variables.synthetic_fem_weights())
img = torch.rand(1, 512, 640, 3, device='cuda') * 255
A.fem.ResNetDS2SPP(img)
torch.cuda.synchronize()
torch.cuda.profiler.start()
A.fem.ResNetDS2SPP(img)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
