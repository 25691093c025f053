"""Seeded synthetic inputs shared by the oracle, the tests and bench.py (TEST/BENCH DATA ONLY).

All descriptors are N(0,1) float32 rounded to bf16-representable values, so every
product q_k * r_k is exact in float32 on both the CPU and the tensor-core paths
(SURVEY.md section 8d).
"""
import numpy as np


def to_bf16_grid(x: np.ndarray) -> np.ndarray:
    """Round float32 to the nearest bf16-representable float32 (ties to even)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(x.shape)


def c1_videos(seed: int = 1, n_q: int = 2, n_r: int = 2, n_noise: int = 3, frames: int = 32,
              dim: int = 512):
    """BASELINE.json configs[0]: 2 videos x 32 frames x 512-d with one planted copy."""
    rng = np.random.default_rng(seed)
    q = to_bf16_grid(rng.normal(size=(n_q, frames, dim)).astype(np.float32))
    r = to_bf16_grid(rng.normal(size=(n_r, frames, dim)).astype(np.float32))
    noise = to_bf16_grid(rng.normal(size=(n_noise, frames, dim)).astype(np.float32))
    q[0, 8:24] = r[1, 4:20]  # planted copy Q0[8:24] == R1[4:20]
    ts = np.stack([np.arange(frames, dtype=np.float64), np.arange(frames, dtype=np.float64) + 1], axis=1)
    return q, r, noise, ts


def unit_rows(x: np.ndarray) -> np.ndarray:
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


def sim_matrix(rng: np.random.Generator, lq: int, lr: int, dim: int = 64, max_copies: int = 2,
               bias: float = 0.5, jitter: float = 0.1, quant: float = 0.0) -> np.ndarray:
    """One query-vs-ref frame-similarity matrix with 0..max_copies planted diagonal copies."""
    a = rng.normal(size=(lq, dim)).astype(np.float32)
    b = rng.normal(size=(lr, dim)).astype(np.float32)
    longest = min(lq, lr)
    if longest >= 12:
        for _ in range(int(rng.integers(0, max_copies + 1))):
            n = int(rng.integers(6, max(7, min(80, longest // 2 + 2))))
            n = min(n, longest)
            qs = int(rng.integers(0, lq - n + 1))
            rs = int(rng.integers(0, lr - n + 1))
            a[qs:qs + n] = b[rs:rs + n] + rng.normal(scale=jitter, size=(n, dim)).astype(np.float32)
    s = unit_rows(a) @ unit_rows(b).T + np.float32(bias)
    if quant:
        s = np.round(s * quant) / quant
    return np.ascontiguousarray(s, dtype=np.float32)
