import torch, time
dev="cuda"
x=torch.randn(8,12,64,160,320,device=dev)
conv=torch.nn.Conv3d(12,1,3,padding=1,bias=False).to(dev)
def t(fn,n=5):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
with torch.no_grad():
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32=tf32
        for bm in (False, True):
            torch.backends.cudnn.benchmark=bm
            print("tf32",tf32,"benchmark",bm, round(t(lambda: conv(x)),3),"ms")
    xc = x.to(memory_format=torch.channels_last_3d)
    torch.backends.cudnn.allow_tf32=True; torch.backends.cudnn.benchmark=True
    print("channels_last tf32", round(t(lambda: conv(xc)),3))
