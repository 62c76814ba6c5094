import sys, time; sys.path.insert(0,'.')
import torch
from qcgraft import nets
torch.backends.cudnn.benchmark=True
dev='cuda'
def bench(m, x, n=10):
    with torch.no_grad():
        for _ in range(3): m(x)
        torch.cuda.synchronize()
        g=torch.cuda.CUDAGraph()
        with torch.cuda.graph(g): y=m(x)
        g.replay(); torch.cuda.synchronize()
        s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): g.replay()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e)/n
for dt in (torch.bfloat16, torch.float16):
    for fold in (True, "fused"):
        m = nets.build_resnet50(5,0); m = nets.fold_batchnorm(m)
        m = m.to(dev, dt).to(memory_format=torch.channels_last)
        if fold == "fused":
            ref = nets.fold_batchnorm(nets.build_resnet50(5,0)).to(dev, dt).to(memory_format=torch.channels_last)
            try:
                m = nets.fuse_resnet_epilogues(m)
                xx = torch.randn(8,3,224,224,device=dev,dtype=dt).contiguous(memory_format=torch.channels_last)
                with torch.no_grad(): print("fused vs ref maxdiff", (m(xx).float()-ref(xx).float()).abs().max().item())
            except Exception as e:
                print("fused failed:", repr(e)[:300]); continue
        for bs in (256,):
            x = torch.randn(bs,3,224,224,device=dev,dtype=dt).contiguous(memory_format=torch.channels_last)
            t = bench(m,x)
            print(dt, 'fold',fold, 'bs',bs, f'{t:.2f} ms', f'{bs/t*1000:.0f} img/s', f'{bs*4.1/t:.0f} TFLOP/s')
y = nets.build_yolo('s',5,0).to(dev, torch.bfloat16).to(memory_format=torch.channels_last)
for bs in (64,):
    x = torch.randn(bs,3,640,640,device=dev,dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t = bench(y,x); print('yolo s bs',bs,f'{t:.2f} ms')
y = y.half()
x = torch.randn(64,3,640,640,device=dev,dtype=torch.float16).contiguous(memory_format=torch.channels_last)
print('yolo fp16', bench(y,x))
