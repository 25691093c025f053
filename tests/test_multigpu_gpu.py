"""Two-GPU tests (skipped on a single-GPU box): query-sharded global top-K agrees with the single-GPU result,
pair-sharded TN alignment gathers to the same boxes."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, results):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        from oracle import synth
        from vsc2022_b200 import distributed as D, vta
        from vsc2022_b200.candidates import CandidateGeneration, MaxScoreAggregation
        from vsc2022_b200.index import VideoFeature
        rng = np.random.default_rng(3)
        grid = lambda n: (rng.integers(-16, 17, size=(n, 64)) / 16.0).astype(np.float32)
        q = [VideoFeature(video_id=i, timestamps=np.arange(24) * 1.0, feature=grid(24)) for i in range(60)]
        r = [VideoFeature(video_id=1000 + i, timestamps=np.arange(24) * 1.0, feature=grid(24)) for i in range(150)]
        cg = CandidateGeneration(r, MaxScoreAggregation())
        K = 1200 * len(q) // 40
        sharded = cg.query(q, K, group=dist.group.WORLD)
        single = cg.query(q, K) if rank == 0 else None
        # stage C: shard the pairs, gather the boxes
        srng = np.random.default_rng(5)
        sims = [synth.sim_matrix(srng, 64, 64) for _ in range(37)]
        model = vta.build_vta_model("TN", tn_max_step=5, min_length=4)
        lo, hi = D.shard_bounds(len(sims), rank, ws)
        mine = model.forward_sim([(str(i), sims[i]) for i in range(lo, hi)])
        everything = D.gather_lists(mine)
        alone = model.forward_sim([(str(i), s) for i, s in enumerate(sims)]) if rank == 0 else None
        results[rank] = {"sharded": [(c.query_id, c.ref_id, float(c.score)) for c in sharded],
                         "single": None if single is None else [(c.query_id, c.ref_id, float(c.score)) for c in single],
                         "tn": everything, "tn_alone": alone}
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_two_gpu_sharding_matches_single_gpu():
    import torch.multiprocessing as mp
    ws, port = 2, _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(ws, port, results), nprocs=ws, join=True)
        r0, r1 = dict(results[0]), dict(results[1])
    assert r0["sharded"] == r1["sharded"] == r0["single"]
    assert len(r0["single"]) > 10
    assert r0["tn"] == r1["tn"] == r0["tn_alone"]
