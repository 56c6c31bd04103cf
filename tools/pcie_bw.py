import torch, time, numpy as np, ctypes, sys, os
sys.path.insert(0, os.getcwd())
from src.x265_b200 import abi
x = torch.empty(64<<20, dtype=torch.uint8).pin_memory()
d = torch.empty(64<<20, dtype=torch.uint8, device='cuda')
for name, fn in (('H2D', lambda: d.copy_(x, non_blocking=True)), ('D2H', lambda: x.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(10): fn()
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    print(name, 'pinned %.1f GB/s' % (10*64/1024/dt))
a = np.zeros(64<<20, dtype=np.uint8)
r = abi.lib_cu().x265cu_host_register(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes))
print('host_register rc', r)
xa = torch.from_numpy(a)
print('is_pinned', xa.is_pinned())
d.copy_(xa, non_blocking=True); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(10): d.copy_(xa, non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('registered numpy H2D %.1f GB/s' % (10*64/1024/dt))
# small chunk copies 2MB
t=time.perf_counter()
for i in range(32): d[i*(2<<20):(i+1)*(2<<20)].copy_(x[i*(2<<20):(i+1)*(2<<20)], non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('2MB chunks H2D %.1f GB/s' % (64/1024/dt))
