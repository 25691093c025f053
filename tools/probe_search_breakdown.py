"""Where the time of the C3 global top-K search goes (dev tool): wraps the phases with synchronising timers."""
import collections
import sys
import time

import torch

sys.path.insert(0, ".")
from vsc2022_b200 import gemm  # noqa: E402
from vsc2022_b200.index import FlatIndex, VideoIndex  # noqa: E402

nqv, nrv, frames, d = 1250, 6250, 32, 512
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(3)
q = torch.randn((nqv * frames, d), generator=g, device=dev).bfloat16().float()
r = torch.randn((nrv * frames, d), generator=g, device=dev).bfloat16().float()
for v in range(0, nqv, 20):
    rv = (v * 7919) % nrv
    q[v * frames + 8:v * frames + 24] = r[rv * frames + 4:rv * frames + 20]
K = 1200 * nqv
index = VideoIndex(d)
index.index.add_device(r)
index.global_topk_device(q, K)
torch.cuda.synchronize()

acc = collections.defaultdict(float)
cnt = collections.Counter()


def timed(name, fn):
    def wrapper(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc[name] += time.perf_counter() - t0; cnt[name] += 1
        return out
    return wrapper


FlatIndex._kth_best = staticmethod(timed("kth_best", FlatIndex._kth_best))
FlatIndex._refilter = staticmethod(timed("refilter", FlatIndex._refilter))
_emit = gemm.gemm_emit
per_call = []


def emit_logged(oa, ob, hits, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = _emit(oa, ob, hits, *a, **k)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    rows = k.get("rows")
    per_call.append((rows.stop - rows.start if rows is not None else oa.rows, dt * 1e3, hits.read_counters()))
    return out


gemm.gemm_emit = timed("gemm_emit", emit_logged)
gemm.prepare_pair = timed("prepare_pair", gemm.prepare_pair)
gemm.HitBuffer.read_counters = timed("read_counters", gemm.HitBuffer.read_counters)
index.index.range_search_max_results = timed("range_search_total", index.index.range_search_max_results)
torch.cuda.synchronize(); t0 = time.perf_counter()
index.global_topk_device(q, K)
torch.cuda.synchronize(); total = time.perf_counter() - t0
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print(f"{k:22s} {v * 1e3:8.2f} ms  x{cnt[k]}")
for rows, ms, (stored, counted) in per_call:
    print(f"   emit rows={rows:6d} {ms:7.3f} ms  stored(after)={stored:9d} counted={counted:9d}  "
          f"{2.0 * rows * r.shape[0] * d / ms / 1e9:7.0f} TFLOP/s")
print(f"{'whole search':22s} {total * 1e3:8.2f} ms (with the synchronising timers; final ordering = whole - range_search_total)")
