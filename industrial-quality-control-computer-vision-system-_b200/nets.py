"""Plain-torch YOLOv8 (n/s/m) and ResNet-50 graphs with random-init weights.

These conv backbones are OUTSIDE the graft target (BASELINE.json north_star): torch/cuDNN runs
them; they exist so that end-to-end images/sec is measured with the real amount of network
work between the hand-written kernels.  ultralytics is not installed, so the YOLOv8 graph is
rebuilt from its published yolov8.yaml; Conv = Conv2d+BN+SiLU is built
directly in its inference-fused form (Conv2d with bias + SiLU), which is what ultralytics runs
after ``model.fuse()``.  The detector returns the three raw Detect-head tensors
[B, 4*reg_max+nc, Hi, Wi] -- decode/NMS is K2's job.

ResNet-50 + head follows models/resent_model.py:68-83 (torchvision resnet50, fc replaced by
Dropout-Linear(2048,512)-ReLU-Dropout-Linear(512,nc)); weights=None because there is no network.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torchvision

YOLO_SCALES = {"n": (0.33, 0.25, 1024), "s": (0.33, 0.50, 1024), "m": (0.67, 0.75, 768)}


def _div8(c: float) -> int:
    return int(math.ceil(c / 8) * 8)


class Conv(nn.Module):
    def __init__(self, c1, c2, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=True)
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        return self.act(self.conv(x))


class Bottleneck(nn.Module):
    def __init__(self, c, shortcut=True):
        super().__init__()
        self.cv1 = Conv(c, c, 3)
        self.cv2 = Conv(c, c, 3)
        self.add = shortcut

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1, 2 * self.c, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        for m in self.m:
            y.append(m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1)
        self.cv2 = Conv(c_ * 4, c2, 1)
        self.m = nn.MaxPool2d(k, 1, k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(3):
            y.append(self.m(y[-1]))
        return self.cv2(torch.cat(y, 1))


class DetectHead(nn.Module):
    def __init__(self, nc, ch, reg_max=16, strides=(8, 16, 32)):
        super().__init__()
        self.nc, self.reg_max, self.strides = nc, reg_max, strides
        c2 = max(16, ch[0] // 4, reg_max * 4)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        for a, b, s in zip(self.cv2, self.cv3, strides):      # ultralytics Detect.bias_init
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[:nc] = math.log(5 / nc / (640 / s) ** 2)

    def forward(self, feats):
        return [torch.cat((self.cv2[i](f), self.cv3[i](f)), 1) for i, f in enumerate(feats)]


class YOLOv8(nn.Module):
    """yolov8.yaml with (depth, width, max_channels) scaling; forward -> [P3, P4, P5] head tensors."""

    def __init__(self, scale: str = "s", nc: int = 5):
        super().__init__()
        d, w, mc = YOLO_SCALES[scale]
        ch = lambda c: _div8(min(c, mc) * w)
        n = lambda k: max(round(k * d), 1)
        c64, c128, c256, c512, c1024 = ch(64), ch(128), ch(256), ch(512), ch(1024)
        self.b0 = Conv(3, c64, 3, 2)
        self.b1 = Conv(c64, c128, 3, 2)
        self.b2 = C2f(c128, c128, n(3), True)
        self.b3 = Conv(c128, c256, 3, 2)
        self.b4 = C2f(c256, c256, n(6), True)
        self.b5 = Conv(c256, c512, 3, 2)
        self.b6 = C2f(c512, c512, n(6), True)
        self.b7 = Conv(c512, c1024, 3, 2)
        self.b8 = C2f(c1024, c1024, n(3), True)
        self.b9 = SPPF(c1024, c1024, 5)
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.h12 = C2f(c1024 + c512, c512, n(3))
        self.h15 = C2f(c512 + c256, c256, n(3))
        self.h16 = Conv(c256, c256, 3, 2)
        self.h18 = C2f(c256 + c512, c512, n(3))
        self.h19 = Conv(c512, c512, 3, 2)
        self.h21 = C2f(c512 + c1024, c1024, n(3))
        self.detect = DetectHead(nc, (c256, c512, c1024))
        self.nc = nc
        self.scale = scale

    def forward(self, x):
        x = self.b2(self.b1(self.b0(x)))
        p3 = self.b4(self.b3(x))
        p4 = self.b6(self.b5(p3))
        p5 = self.b9(self.b8(self.b7(p4)))
        f12 = self.h12(torch.cat((self.up(p5), p4), 1))
        f15 = self.h15(torch.cat((self.up(f12), p3), 1))
        f18 = self.h18(torch.cat((self.h16(f15), f12), 1))
        f21 = self.h21(torch.cat((self.h19(f18), p5), 1))
        return self.detect([f15, f18, f21])


def build_yolo(scale: str = "s", nc: int = 5, seed: int = 0) -> YOLOv8:
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = YOLOv8(scale, nc).eval()
    torch.random.set_rng_state(g)
    return m


def build_resnet50(nc: int = 5, seed: int = 0) -> nn.Module:
    """models/resent_model.py:68-83 (weights=None: no network for the ImageNet checkpoint)."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed + 1)
    m = torchvision.models.resnet50(weights=None)
    nf = m.fc.in_features
    m.fc = nn.Sequential(nn.Dropout(0.5), nn.Linear(nf, 512), nn.ReLU(inplace=True), nn.Dropout(0.3),
                         nn.Linear(512, nc))
    torch.random.set_rng_state(g)
    return m.eval()


# ---- checkpoints ---------------------------------------------------------------------------------------------
# ultralytics' DetectionModel is nn.Sequential "model.{i}"; the same layers by name here (yolov8.yaml order)
_ULTRA_INDEX = {"b0": 0, "b1": 1, "b2": 2, "b3": 3, "b4": 4, "b5": 5, "b6": 6, "b7": 7, "b8": 8, "b9": 9,
                "h12": 12, "h15": 15, "h16": 16, "h18": 18, "h19": 19, "h21": 21, "detect": 22}
_SCALE_BY_STEM = {16: "n", 32: "s", 48: "m"}


def _checkpoint_state_dict(path: str) -> dict:
    """torch file -> flat state dict.  Accepts a plain state_dict, {'state_dict': ...} (models/resent_model.py:93-99),
    {'model_state_dict': ...} (scripts/download_models.py:153-157), {'model': state_dict}, and -- only where the
    ultralytics package is importable, its pickled {'model': DetectionModel} checkpoints (models/yolo_model.py:52)."""
    try:
        ck = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        ck = torch.load(path, map_location="cpu", weights_only=False)      # pickled modules need their package
    if isinstance(ck, dict):
        for k in ("state_dict", "model_state_dict", "model", "ema"):
            v = ck.get(k)
            if isinstance(v, dict) and v and all(torch.is_tensor(t) for t in v.values()):
                return v
            if isinstance(v, nn.Module):
                return v.float().state_dict()
        if ck and all(torch.is_tensor(t) for t in ck.values()):
            return ck
    if isinstance(ck, nn.Module):
        return ck.float().state_dict()
    raise ValueError(f"{path}: no state dict found in checkpoint")


def load_resnet_checkpoint(path: str, nc: int = 5) -> nn.Module:
    """models/resent_model.py:85-107: resnet50 + head, state dict loaded strictly."""
    m = torchvision.models.resnet50(weights=None)
    nf = m.fc.in_features
    m.fc = nn.Sequential(nn.Dropout(0.5), nn.Linear(nf, 512), nn.ReLU(inplace=True), nn.Dropout(0.3), nn.Linear(512, nc))
    sd = {k[7:] if k.startswith("module.") else k: v for k, v in _checkpoint_state_dict(path).items()}
    m.load_state_dict(sd, strict=True)
    return m.eval()


def ultralytics_state_dict(model: "YOLOv8") -> dict:
    """This module's weights under ultralytics' key names (fused form: model.{i}...conv.weight / conv.bias)."""
    out = {}
    for k, v in model.state_dict().items():
        head, rest = k.split(".", 1)
        out[f"model.{_ULTRA_INDEX[head]}.{rest}"] = v.detach().clone()
    return out


def load_yolo_checkpoint(path: str):
    """A YOLOv8 detection checkpoint in ultralytics' key layout -> (YOLOv8 module, scale, nc).  Unfused Conv+BN pairs
    ('...conv.weight' + '...bn.*', eps 1e-3 as ultralytics builds them) are folded into the biased convs this module
    runs -- what ``model.fuse()`` does before inference; scale and class count are read off the tensor shapes."""
    sd = _checkpoint_state_dict(path)
    sd = {(k[6:] if k.startswith("model.model.") else k): v.float() for k, v in sd.items()}
    stem = sd["model.0.conv.weight"].shape[0]
    if stem not in _SCALE_BY_STEM:
        raise ValueError(f"{path}: unsupported YOLOv8 width (stem channels {stem}); n / s / m are built")
    scale = _SCALE_BY_STEM[stem]
    nc = sd["model.22.cv3.0.2.weight"].shape[0]
    m = YOLOv8(scale, nc).eval()
    new = {}
    for k, ref in m.state_dict().items():
        head, rest = k.split(".", 1)
        src = f"model.{_ULTRA_INDEX[head]}.{rest}"
        if src in sd:
            t = sd[src]
        elif rest.endswith("conv.bias") and src[:-9] + "bn.weight" in sd:
            t = None                                                   # produced with the weight below
        else:
            raise KeyError(f"{path}: missing {src}")
        if rest.endswith("conv.weight") and src[:-11] + "bn.weight" in sd:
            pre = src[:-11] + "bn."
            g, b, mu, var = sd[pre + "weight"], sd[pre + "bias"], sd[pre + "running_mean"], sd[pre + "running_var"]
            sc = g / torch.sqrt(var + 1e-3)
            new[k] = t * sc.view(-1, 1, 1, 1)
            new[k[:-6] + "bias"] = b - mu * sc
        elif t is not None:
            new[k] = t
        if k in new and new[k].shape != ref.shape:
            raise ValueError(f"{path}: {src} has shape {tuple(new[k].shape)}, expected {tuple(ref.shape)}")
    m.load_state_dict(new, strict=True)
    return m, scale, nc


def fold_batchnorm(m: nn.Module) -> nn.Module:
    """Inference-time Conv+BN folding (same function, fewer elementwise passes over the activations):
    every (convN, bnN) / downsample (conv, bn) pair of a torchvision ResNet becomes one biased conv."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    m.eval()
    m.conv1 = fuse_conv_bn_eval(m.conv1, m.bn1)
    m.bn1 = nn.Identity()
    for layer in (m.layer1, m.layer2, m.layer3, m.layer4):
        for blk in layer:
            for i in (1, 2, 3):
                conv, bn = getattr(blk, f"conv{i}", None), getattr(blk, f"bn{i}", None)
                if conv is not None and isinstance(bn, nn.BatchNorm2d):
                    setattr(blk, f"conv{i}", fuse_conv_bn_eval(conv, bn))
                    setattr(blk, f"bn{i}", nn.Identity())
            if blk.downsample is not None and isinstance(blk.downsample[1], nn.BatchNorm2d):
                blk.downsample = nn.Sequential(fuse_conv_bn_eval(blk.downsample[0], blk.downsample[1]))
    return m


class _FusedBottleneck(nn.Module):
    """torchvision Bottleneck (BN already folded) with cuDNN's fused conv+bias+ReLU / conv+bias+add+ReLU
    epilogues instead of separate elementwise kernels (same arithmetic, fewer passes over the activations)."""

    def __init__(self, blk):
        super().__init__()
        self.c1, self.c2, self.c3 = blk.conv1, blk.conv2, blk.conv3
        self.down = blk.downsample[0] if blk.downsample is not None else None

    @staticmethod
    def _cr(x, c):
        return torch.cudnn_convolution_relu(x, c.weight, c.bias, c.stride, c.padding, c.dilation, c.groups)

    def forward(self, x):
        idt = x
        if self.down is not None:
            d = self.down
            idt = torch.nn.functional.conv2d(x, d.weight, d.bias, d.stride, d.padding, d.dilation, d.groups)
        out = self._cr(self._cr(x, self.c1), self.c2)
        c = self.c3
        return torch.cudnn_convolution_add_relu(out, c.weight, idt, 1.0, c.bias, c.stride, c.padding, c.dilation,
                                                c.groups)


def fuse_resnet_epilogues(m: nn.Module) -> nn.Module:
    """Call after fold_batchnorm() and after moving the model to CUDA."""
    for name in ("layer1", "layer2", "layer3", "layer4"):
        layer = getattr(m, name)
        setattr(m, name, nn.Sequential(*[_FusedBottleneck(b) for b in layer]))
    return m


def yolo_gflops(scale: str, H: int, W: int) -> float:
    """Published YOLOv8 forward cost at 640^2 (8.7/28.6/78.9 GFLOP), scaled by area."""
    base = {"n": 8.7, "s": 28.6, "m": 78.9}[scale]
    return base * (H * W) / (640 * 640)


class GraphedNet:
    """CUDA-graph replay of a conv net at fixed input buffers (one graph per (shape, buffer), captured lazily after a
    cuDNN-autotune warm-up; all graphs of one net share a memory pool -- they never run concurrently).  The producer
    kernel writes straight into the input: either ``input_buffer(shape)`` + ``run(shape)``, or any caller-owned view
    (a slice of an arena) through ``run_on(view)``.  Outputs are valid until the next replay of the same graph."""

    def __init__(self, module: nn.Module, device, dtype, channels_last: bool = True):
        self.module, self.device, self.dtype, self.cl = module, device, dtype, channels_last
        self.slots = {}
        self.pool = None

    def input_buffer(self, shape) -> torch.Tensor:
        shape = tuple(int(v) for v in shape)
        if shape not in self.slots:
            self.slots[shape] = {"in": torch.zeros(shape, dtype=self.dtype, device=self.device), "graph": None}
        return self.slots[shape]["in"]

    def _fwd(self, x):
        if self.cl:
            x = x.contiguous(memory_format=torch.channels_last)
        return self.module(x)

    @torch.no_grad()
    def _replay(self, slot):
        if slot["graph"] is None:
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(3):                     # cuDNN autotune + allocator warm-up
                    self._fwd(slot["in"])
            cur.wait_stream(side)
            torch.cuda.synchronize(self.device)
            if self.pool is None:
                self.pool = torch.cuda.graph_pool_handle()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=self.pool):
                slot["out"] = self._fwd(slot["in"])
            slot["graph"] = g
        slot["graph"].replay()
        return slot["out"]

    def run(self, shape):
        return self._replay(self.slots[tuple(int(v) for v in shape)])

    def run_on(self, view: torch.Tensor):
        """Replay with ``view`` (contiguous, fixed address) as the static input."""
        key = (tuple(view.shape), view.data_ptr())
        slot = self.slots.get(key)
        if slot is None:
            assert view.is_contiguous() and view.dtype == self.dtype
            slot = self.slots[key] = {"in": view, "graph": None}
        return self._replay(slot)
