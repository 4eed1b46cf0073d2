import sys, os, time
sys.path.insert(0, os.getcwd())
from fc_planner_b200 import _lib
import torch
torch.cuda.init(); torch.zeros(1, device='cuda')
lib = _lib.load()
for mb in (64, 256, 600, 3100):
    t = time.perf_counter(); p = lib.fcvm_host_alloc(mb << 20); dt = time.perf_counter() - t
    t = time.perf_counter(); lib.fcvm_host_free(p); df = time.perf_counter() - t
    print(mb, 'MB pinned alloc', round(dt * 1e3, 1), 'ms free', round(df * 1e3, 1), 'ms')
x = torch.empty(600 << 20, dtype=torch.uint8, device='cuda')
import numpy as np
h = np.empty(600 << 20, np.uint8); h[:] = 0
ht = torch.from_numpy(h)
torch.cuda.synchronize(); t = time.perf_counter(); ht.copy_(x); torch.cuda.synchronize(); print('600 MB D2H pageable', round((time.perf_counter() - t) * 1e3, 1), 'ms')
hp = torch.empty(600 << 20, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize(); t = time.perf_counter(); hp.copy_(x); torch.cuda.synchronize(); print('600 MB D2H pinned', round((time.perf_counter() - t) * 1e3, 1), 'ms')
