"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: sharding, gathers, radius agreement."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vsc2022_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        out = {}
        items = list(range(11))
        mine = D.shard(items)
        out["shard"] = mine
        out["gathered"] = D.gather_lists([(rank, x) for x in mine])
        rng = np.random.default_rng(7)
        scores = torch.from_numpy(rng.normal(size=1000).astype(np.float32))
        lo, hi = D.shard_bounds(1000, rank, ws)
        local = scores[lo:hi][: 300 + 100 * rank]          # ranks hold different numbers of survivors
        out["var"] = D.all_gather_variable(local).numpy()
        out["radius_ip"] = D.agree_radius(local, 50, True)
        out["radius_l2"] = D.agree_radius(local, 50, False)
        out["count"] = D.global_count(local.numel(), "cpu")
        out["tmax"] = D.max_over_ranks(1.0 + rank, "cpu")
        out["union"] = torch.cat([scores[a:b][: 300 + 100 * r] for r, (a, b) in
                                  enumerate(D.shard_bounds(1000, r, ws) for r in range(ws))]).numpy()
        # stage A sharding (video i on rank i % ws) followed by the all-gather of the descriptors
        from vsc2022_b200 import inference_impl
        from vsc2022_b200.index import VideoFeature
        vrng = np.random.default_rng(11)
        videos = [VideoFeature(video_id=f"R{i:06d}", timestamps=np.arange(n) * 1.0,
                               feature=vrng.normal(size=(n, 8)).astype(np.float32)) for i, n in enumerate([3, 5, 1, 4, 2])]
        mine_v = [v for _, v in inference_impl.select_videos(videos, rank, ws)]
        everyone = D.all_gather_video_features(mine_v, len(videos), device="cpu")
        out["videos_ok"] = all(a.video_id == b.video_id and np.array_equal(a.feature, b.feature) and
                               np.array_equal(a.timestamps, b.timestamps) for a, b in zip(everyone, videos))
        results[rank] = out
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ws, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(ws, port, results), nprocs=ws, join=True)
        r0, r1 = results[0], results[1]
    assert r0["shard"] == [0, 1, 2, 3, 4, 5] and r1["shard"] == [6, 7, 8, 9, 10]
    assert r0["gathered"] == r1["gathered"] == [(0, x) for x in range(6)] + [(1, x) for x in range(6, 11)]
    union = r0["union"]
    assert np.array_equal(r0["var"], union) and np.array_equal(r1["var"], union)
    assert r0["radius_ip"] == r1["radius_ip"] == float(np.sort(union)[::-1][49])
    assert r0["radius_l2"] == r1["radius_l2"] == float(np.sort(union)[49])
    assert r0["count"] == r1["count"] == 300 + 400
    assert r0["tmax"] == r1["tmax"] == 2.0
    assert r0["videos_ok"] and r1["videos_ok"]


def test_shard_bounds_cover_everything_in_order():
    for n in (0, 1, 7, 8, 8000):
        for ws in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert D.world() == (0, 1)
    assert D.gather_lists([1, 2]) == [1, 2]
