import torch, sys
sys.path.insert(0, ".")
import hairfastgan_b200.op as OP
dev="cuda"
kern = torch.tensor([1., 3., 3., 1.], device=dev); k2 = kern[None,:]*kern[:,None]; k2 = k2/k2.sum()
flush = torch.empty(256<<20, dtype=torch.uint8, device=dev)
def avg(fn,n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); tot=0
    for _ in range(n):
        flush.fill_(1); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); tot+=e0.elapsed_time(e1)
    return tot/n
xu = torch.randn(1,256,1025,1025,device=dev)
ms = avg(lambda: OP.upfirdn2d(xu, k2*4, pad=(1,1))); print("up1 k4: %.3f ms %.0f GB/s" % (ms, (xu.numel()+256*1024*1024)*4/ms/1e6))
xd = torch.randn(1,256,1024,1024,device=dev)
ms = avg(lambda: OP.upfirdn2d(xd, k2, down=2, pad=(1,1))); print("down2 k4: %.3f ms %.0f GB/s" % (ms, (xd.numel()*1.25)*4/ms/1e6))
xs = torch.randn(48,3,512,512,device=dev)
ms = avg(lambda: OP.upfirdn2d(xs, k2*4, up=2, pad=(2,1))); print("up2 k4: %.3f ms %.0f GB/s" % (ms, xs.numel()*5*4/ms/1e6))
