import sys, json
sys.path.insert(0, '/root/repo')
import torch, bench
from wps_b200 import capi
dev = torch.device("cuda", 0)
h = capi.Handle(0, capi.PTR_DEVICE)
print(json.dumps(bench.aux_kernels_arm(h, torch, dev)))
