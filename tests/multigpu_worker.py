"""Worker of tests/test_multigpu_gpu.py: run under `python -m torch.distributed.run --nproc-per-node 2`."""
import faulthandler
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(120, exit=True)
rank, ws = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))

from oracle import synth  # noqa: E402
from vsc2022_b200 import distributed as D, vta  # noqa: E402
from vsc2022_b200.candidates import CandidateGeneration, MaxScoreAggregation  # noqa: E402
from vsc2022_b200.index import VideoFeature  # noqa: E402

rng = np.random.default_rng(3)
grid = lambda n: (rng.integers(-16, 17, size=(n, 64)) / 16.0).astype(np.float32)  # noqa: E731
q = [VideoFeature(video_id=i, timestamps=np.arange(24) * 1.0, feature=grid(24)) for i in range(60)]
r = [VideoFeature(video_id=1000 + i, timestamps=np.arange(24) * 1.0, feature=grid(24)) for i in range(150)]
cg = CandidateGeneration(r, MaxScoreAggregation())
K = 1200 * len(q) // 40
flat = lambda cs: [(c.query_id, c.ref_id, float(c.score)) for c in cs]  # noqa: E731
sharded = flat(cg.query(q, K, group=dist.group.WORLD))     # query rows sharded, radius agreed across ranks
single = flat(cg.query(q, K))                              # every rank alone
assert sharded == single and len(single) > 10, "sharded search differs from single-GPU search"

# Gaussian float32 descriptors: the sharded device-side schedule with filtered batches (single-product candidates + exact
# re-score, every rank its slice of every batch, all-reduced counts / histograms) against the same search on one GPU
from vsc2022_b200.index import VideoIndex  # noqa: E402
grng = np.random.default_rng(8)
unit = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)  # noqa: E731
gq, gr = unit(grng.normal(size=(4800, 64))), unit(grng.normal(size=(6000, 64)))
gi = VideoIndex(64)
gi.index.filter_from_rows = 64
gi.index.add(gr)
gq_dev = torch.from_numpy(gq).cuda()
Kg = 150_000
a = [t.cpu().numpy() for t in gi.global_topk_device(gq_dev, Kg, group=dist.group.WORLD)]
b = [t.cpu().numpy() for t in gi.global_topk_device(gq_dev, Kg)]
sa, sb = set(zip(a[0].tolist(), a[1].tolist())), set(zip(b[0].tolist(), b[1].tolist()))
kth = float(b[2][-1])
exact = gq.astype(np.float64) @ gr.astype(np.float64).T
assert len(sa) == Kg and len(sa ^ sb) <= 8 and all(abs(exact[p] - kth) <= 4e-6 for p in sa ^ sb), len(sa ^ sb)
assert max(abs(float(s) - exact[i, j]) for i, j, s in zip(a[0][::97], a[1][::97], a[2][::97])) <= 4e-6

srng = np.random.default_rng(5)
sims = [synth.sim_matrix(srng, 64, 64) for _ in range(37)]
model = vta.build_vta_model("TN", tn_max_step=5, min_length=4)
lo, hi = D.shard_bounds(len(sims), rank, ws)
everything = D.gather_lists(model.forward_sim([(str(i), sims[i]) for i in range(lo, hi)]))
alone = model.forward_sim([(str(i), s) for i, s in enumerate(sims)])
assert everything == alone, "pair-sharded TN differs from single-GPU TN"
# stage A: videos sharded i % world, descriptors all-gathered (NCCL) == every video inferred on one GPU
from vsc2022_b200 import inference_impl  # noqa: E402
from vsc2022_b200.sscd import SSCDResNet50, TorchReference  # noqa: E402
ref = TorchReference(seed=1, device=f"cuda:{rank}")
sscd = SSCDResNet50(ref.trunk, ref.head)
vrng = np.random.default_rng(9)
videos = [(f"R{i:06d}", np.arange(n) * 1.0, vrng.integers(0, 256, size=(n, 64, 64, 3), dtype=np.uint8))
          for i, n in enumerate([5, 9, 3, 12, 7])]
dev = torch.device("cuda", rank)
mine = inference_impl.infer_videos([v for _, v in inference_impl.select_videos(videos, rank, ws)], sscd, batch_size=16, device=dev)
gathered = D.all_gather_video_features(mine, len(videos), device=dev)
alone_feats = inference_impl.infer_videos(videos, sscd, batch_size=16, device=dev)
assert all(a.video_id == b.video_id and np.array_equal(a.feature, b.feature) for a, b in zip(gathered, alone_feats)), \
    "all-gathered descriptors differ from single-GPU inference"
print(f"MULTIGPU_OK rank={rank} candidates={len(single)} boxes={sum(len(b) for _, b in alone)}", flush=True)
dist.destroy_process_group()
