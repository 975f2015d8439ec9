import sys, json, statistics, torch
sys.path.insert(0, '.')
from protstruc_b200 import _cabi
lib=_cabi.load(); s=torch.cuda.current_stream().cuda_stream
B,L,A=16,512,15
g=torch.Generator(device='cuda').manual_seed(0)
xyz=10*torch.randn(B,L,A,3,device='cuda',generator=g); mask=torch.rand(B,L,A,device='cuda',generator=g)<0.5
dist=torch.empty(B,L,L,A,A,device='cuda'); dm=torch.empty(B,L,L,A,A,dtype=torch.bool,device='cuda')
nbytes=B*L*L*A*A*5
def t(fn,it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for wpt_bit in (0,):
  for lock in (0,1):
    for slots in (2,3,4,5,6):
        for so in (1,0):
            v=(lock<<11)|(so<<10)|(wpt_bit<<9)|(slots<<4)
            ms=t(lambda: _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(),mask.data_ptr(),0,dist.data_ptr(),dm.data_ptr(),B,L,A,v,s),'k'))
            print(f"wpt={1 if wpt_bit else 2} lockstep={lock} slots={slots} stores_only={so}: {ms:.4f} ms {nbytes/ms/1e6:7.0f} GB/s")
# dist-only stores-only
for slots in (2,3,4,6):
    v=(1<<10)|(slots<<4)
    ms=t(lambda: _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(),None,0,dist.data_ptr(),None,B,L,A,v,s),'k'))
    print(f"dist-only stores_only slots={slots}: {ms:.4f} ms {B*L*L*A*A*4/ms/1e6:7.0f} GB/s")
ms=t(lambda: dist.zero_()); print(f"torch fill dist: {ms:.4f} ms {B*L*L*A*A*4/ms/1e6:7.0f} GB/s")
ms=t(lambda: dm.zero_()); print(f"torch fill mask: {ms:.4f} ms {B*L*L*A*A/ms/1e6:7.0f} GB/s")
for bps in (4, 8, 16, 32):
    ms=t(lambda: _cabi.check(lib.ps_debug_fill_pattern(dist.data_ptr(), dist.numel(), bps, s),'f'))
    print(f"plain STG.128 non-uniform fill, {bps} blocks/SM: {ms:.4f} ms {dist.numel()*4/ms/1e6:7.0f} GB/s")
src=torch.randn(B,L,L,A,A,device='cuda')
ms=t(lambda: dist.copy_(src)); print(f"torch copy (read+write): {ms:.4f} ms  write side {dist.numel()*4/ms/1e6:7.0f} GB/s, read+write {2*dist.numel()*4/ms/1e6:7.0f} GB/s")
