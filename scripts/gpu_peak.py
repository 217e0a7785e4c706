import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayes_drt_b200 import capi
print('device', torch.cuda.get_device_name(0))
for i in range(3):
    print('FP64 peak TFLOP/s (dfma, dmma):', capi.peak_fp64())
