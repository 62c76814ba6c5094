"""Probe: ResNet-50 bf16 channels_last, batch 256 -- conv1 input with 3 / 4 / 8 channels (zero-padded weights)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qcgraft import nets
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
for C in (3, 4, 8):
    m = nets.fold_batchnorm(nets.build_resnet50(5, 0).eval())
    conv1 = m.conv1
    if C != 3:
        new = torch.nn.Conv2d(C, conv1.out_channels, 7, 2, 3, bias=conv1.bias is not None)
        with torch.no_grad():
            new.weight.zero_()
            new.weight[:, :3] = conv1.weight
            if conv1.bias is not None:
                new.bias.copy_(conv1.bias)
        m.conv1 = new
    m = m.to(dev, torch.bfloat16).to(memory_format=torch.channels_last)
    m = nets.fuse_resnet_epilogues(m)
    x = torch.randn(256, C, 224, 224, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = m(x)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            g.replay()
        e.record(); torch.cuda.synchronize()
        print(f"C={C}: {s.elapsed_time(e) / 10:.3f} ms per 256")
