import numpy as np
import pytest
import torch


@pytest.fixture(scope="module")
def qc():
    from qcgraft import ffi
    return ffi.Context.current(0)


def bf16_bits(t: torch.Tensor) -> torch.Tensor:
    return t.detach().cpu().contiguous().view(torch.int16)


def bf16_ulp_diff(a: torch.Tensor, b: torch.Tensor) -> int:
    """max distance in bf16 ulps (monotone integer mapping of the bit patterns)."""
    def key(t):
        i = t.detach().cpu().contiguous().view(torch.int16).to(torch.int32) & 0xFFFF
        return torch.where(i >= 0x8000, 0x8000 - i, i)
    return int((key(a) - key(b)).abs().max()) if a.numel() else 0
