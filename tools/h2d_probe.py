import sys, time, ctypes, torch
sys.path.insert(0, '/root/repo')
from viterbi_spl_b200 import _lib
B,T,S=1024,3000,361
host=torch.empty((B,T,S),dtype=torch.float32).pin_memory(); host.normal_()
dev=torch.empty((B,T,S),dtype=torch.float32,device='cuda')
torch.cuda.synchronize()
for k in range(3):
    t0=time.perf_counter(); dev.copy_(host,non_blocking=True); torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print('contiguous copy %.1f ms %.1f GB/s'%(dt*1e3, host.numel()*4/dt/1e9))
L=_lib.load(); st=torch.cuda.current_stream()
for slab in (94, 188, 375, 3000):
    t0=time.perf_counter()
    for a in range(0,T,slab):
        L.vit_upload_frames_f32(ctypes.c_void_p(dev.data_ptr()), ctypes.c_void_p(host.data_ptr()), B,T,S,a,min(T,a+slab), ctypes.c_void_p(st.cuda_stream))
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    print('slab %d: %.1f ms %.1f GB/s'%(slab, dt*1e3, host.numel()*4/dt/1e9))
p=torch.empty((B,T),dtype=torch.int64,device='cuda'); hp=torch.empty((B,T),dtype=torch.int64).pin_memory()
t0=time.perf_counter(); hp.copy_(p,non_blocking=True); torch.cuda.synchronize(); print('d2h paths %.2f ms'%((time.perf_counter()-t0)*1e3))
