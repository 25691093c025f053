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


# ---------------------------------------------------------------------------------------------------------------------
# configs[3] as the matching track sees it: candidate pairs over frame descriptors (numpy, per-video seeds, so the GPU
# arm, the CPU arm of bench.py and the tests can each generate exactly the videos they need and get identical values)
# ---------------------------------------------------------------------------------------------------------------------
import numpy as np  # noqa: E402


class C4Workload:
    """`n_pairs` candidate pairs: query video q has `per_query` candidates drawn from a pool of `n_refs` reference
    videos; every video has `frames` L2-normalised `dim`-d float32 descriptors.  Every second pair carries a planted
    copy: 20-80 consecutive reference frames (+ jitter, re-normalised) written into the query video."""

    def __init__(self, n_pairs=8000, per_query=5, n_refs=1600, frames=300, dim=512, seed=4, jitter=0.1):
        self.n_pairs, self.per_query, self.n_refs, self.frames, self.dim = n_pairs, per_query, n_refs, frames, dim
        self.seed, self.jitter = seed, jitter
        self.n_queries = (n_pairs + per_query - 1) // per_query
        rng = np.random.default_rng([seed, 0])
        self.pair_query = np.arange(n_pairs) // per_query
        self.pair_ref = rng.integers(0, n_refs, size=n_pairs)
        self.planted = (np.arange(n_pairs) % 2) == 0
        self.plant_len = rng.integers(min(20, frames), min(80, frames) + 1, size=n_pairs)
        self.plant_q = (rng.random(n_pairs) * (frames - self.plant_len + 1)).astype(np.int64)
        self.plant_r = (rng.random(n_pairs) * (frames - self.plant_len + 1)).astype(np.int64)

    @staticmethod
    def _unit(x):
        return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)

    def ref_video(self, r: int, out=None) -> np.ndarray:
        x = np.random.default_rng([self.seed, 1, int(r)]).standard_normal((self.frames, self.dim), dtype=np.float32)
        x = self._unit(x)
        if out is not None:
            out[...] = x
            return out
        return x

    def query_video(self, q: int, refs=None, out=None) -> np.ndarray:
        """`refs`: optional {ref id: array} cache of already generated reference videos."""
        rng = np.random.default_rng([self.seed, 2, int(q)])
        x = self._unit(rng.standard_normal((self.frames, self.dim), dtype=np.float32))
        for p in range(q * self.per_query, min(self.n_pairs, (q + 1) * self.per_query)):
            if not self.planted[p]:
                continue
            r = int(self.pair_ref[p])
            rv = refs[r] if refs is not None and r in refs else self.ref_video(r)
            n, a, b = int(self.plant_len[p]), int(self.plant_q[p]), int(self.plant_r[p])
            noise = rng.standard_normal((n, self.dim), dtype=np.float32) * np.float32(self.jitter / np.sqrt(self.dim))
            x[a:a + n] = self._unit(rv[b:b + n] + noise)
        if out is not None:
            out[...] = x
            return out
        return x

    def videos_for(self, pair_lo: int, pair_hi: int, q_base=None, r_base=None):
        """Descriptors of every video the pairs [pair_lo, pair_hi) touch.  Returns (query ids, ref ids, Q, R):
        Q / R are [n_videos * frames, dim] float32 arrays (written into q_base / r_base when given -- e.g. pinned
        buffers), video i of the returned id lists occupying rows [i * frames, (i + 1) * frames)."""
        q_ids = sorted(set(self.pair_query[pair_lo:pair_hi].tolist()))
        r_ids = sorted(set(self.pair_ref[pair_lo:pair_hi].tolist()))
        f = self.frames
        R = r_base if r_base is not None else np.empty((len(r_ids) * f, self.dim), dtype=np.float32)
        Q = q_base if q_base is not None else np.empty((len(q_ids) * f, self.dim), dtype=np.float32)
        cache = {}
        for i, r in enumerate(r_ids):
            cache[r] = self.ref_video(r, out=R[i * f:(i + 1) * f])
        for i, q in enumerate(q_ids):
            self.query_video(q, refs=cache, out=Q[i * f:(i + 1) * f])
        return q_ids, r_ids, Q, R
