"""Synthetic workloads of the shapes BASELINE.json names, generated on the device.

Used by bench.py, tools/ and the full-size GPU tests.  No oracle imports here.
"""
from dataclasses import dataclass

import torch


@dataclass
class TnWorkload:
    sims: torch.Tensor   # float32 [n_pairs * lq * lr]
    off: torch.Tensor    # int64 [n_pairs]
    lq: torch.Tensor     # int32 [n_pairs]
    lr: torch.Tensor     # int32 [n_pairs]
    n_pairs: int
    max_lq: int
    max_lr: int


def tn_pairs_device(n_pairs: int, lq: int, lr: int, seed: int, device, dim: int = 64,
                    bias: float = 0.5, jitter: float = 0.1, chunk: int = 1000) -> TnWorkload:
    """configs[3]: `n_pairs` frame-similarity matrices (lq x lr, float32), sims = Q.R^T + bias from
    L2-normalised descriptors with 0-2 planted diagonal copies of 20-80 frames each."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    sims = torch.empty((n_pairs, lq, lr), dtype=torch.float32, device=device)
    longest = min(lq, lr)
    for start in range(0, n_pairs, chunk):
        n = min(chunk, n_pairs - start)
        a = torch.randn((n, lq, dim), generator=gen, device=device)
        b = torch.randn((n, lr, dim), generator=gen, device=device)
        rows = torch.arange(lq, device=device)[None, :]
        for copy in range(2):
            use = torch.rand((n,), generator=gen, device=device) < 0.5
            length = torch.randint(min(20, longest), min(80, longest) + 1, (n,), generator=gen, device=device)
            qs = (torch.rand((n,), generator=gen, device=device) * (lq - length + 1).float()).long()
            rs = (torch.rand((n,), generator=gen, device=device) * (lr - length + 1).float()).long()
            planted = use[:, None] & (rows >= qs[:, None]) & (rows < (qs + length)[:, None])   # [n, lq]
            src_row = (rs[:, None] + rows - qs[:, None]).clamp(0, lr - 1)
            src = torch.gather(b, 1, src_row[:, :, None].expand(-1, -1, dim))
            src = src + jitter * torch.randn(src.shape, generator=gen, device=device)
            a = torch.where(planted[:, :, None], src, a)
        a = torch.nn.functional.normalize(a, dim=2)
        b = torch.nn.functional.normalize(b, dim=2)
        torch.baddbmm(torch.full((1, 1, 1), bias, device=device), a, b.transpose(1, 2), out=sims[start:start + n])
    off = torch.arange(n_pairs, device=device, dtype=torch.int64) * (lq * lr)
    lqs = torch.full((n_pairs,), lq, dtype=torch.int32, device=device)
    lrs = torch.full((n_pairs,), lr, dtype=torch.int32, device=device)
    return TnWorkload(sims.reshape(-1), off, lqs, lrs, n_pairs, lq, lr)
