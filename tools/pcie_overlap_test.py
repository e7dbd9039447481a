import torch, time
torch.cuda.set_device(0)
MB=1<<20
h_src=[torch.empty(3*MB, dtype=torch.uint8).pin_memory() for _ in range(4)]
h_dst=[torch.empty(12*MB, dtype=torch.uint8).pin_memory() for _ in range(4)]
d_src=[torch.empty(3*MB, dtype=torch.uint8, device='cuda') for _ in range(4)]
d_dst=[torch.empty(12*MB, dtype=torch.uint8, device='cuda') for _ in range(4)]
streams=[torch.cuda.Stream() for _ in range(4)]
def run(mode, n=48):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(n):
        k=i%4; s=streams[k]
        with torch.cuda.stream(s):
            if mode in ('both','h2d'): d_src[k].copy_(h_src[k], non_blocking=True)
            if mode=='both': d_dst[k][:1024].add_(1)
            if mode in ('both','d2h'): h_dst[k].copy_(d_dst[k], non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t0
    b={'both':15,'h2d':3,'d2h':12}[mode]*MB*n
    print(mode, f'{dt/n*1e6:.1f} us/frame  {b/dt/1e9:.1f} GB/s')
for m in ('h2d','d2h','both','both'): run(m)
