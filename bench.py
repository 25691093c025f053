#!/usr/bin/env python3
"""Headline benchmark: query-ref pairs localized / second (BASELINE.json metric).

Workload (config.workload = "c4_tn_localization"): BASELINE.json configs[3] -- per GPU 8000
candidate pairs, each a 300x300 float32 frame-similarity matrix (sims = Q.R^T + 0.5 from
L2-normalised descriptors with 0-2 planted diagonal copies), temporal-network alignment with the
vsc2022 parameters (tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3).
One step = one pass of the hot path over the whole batch.

  value   device time only: similarity matrices already resident in HBM (2.88 GB per GPU, larger
          than the 126 MB L2, so no flush is needed between steps); CUDA events on the launch stream.
  e2e     the same batch through the public host-buffer API (vcsl.vta-style TN.forward_packed):
          pinned host matrices -> H2D -> kernels -> boxes D2H, all inside the timed region.
  roofline  HBM: algorithmic bytes = 4*Lq*Lr per pair over the device time of the
          whole TN pipeline call; per-kernel times are listed under roofline.stages_ms.
  cpu_baseline  the oracle's port of the reference CPU path (VCSL TN on networkx, one process per
          host core) on a bounded sample of the same workload -- a reported baseline, not a target.

`--impl reference` runs only that CPU arm (bounded sample per step) and prints the same JSON shape.
Multi-GPU (`torchrun ... bench.py --gpus N`): pairs are independent, every rank aligns its own 8000
pairs (weak scaling), no data-path collective; time = max over ranks.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

PAIRS_PER_GPU = 8000
LQ = LR = 300
TN_CFG = dict(tn_max_step=5, tn_top_k=5, max_path=10, min_sim=0.2, min_length=4, max_iou=0.3)
METRIC = "query-ref pairs localized/sec"
UNIT = "pairs/s"


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _tn_job(args):
    from oracle import tn_networkx
    return tn_networkx.tn(args, **TN_CFG)


def cpu_reference_run(n_pairs, seed):
    """The reference CPU path (oracle port: VCSL TN on networkx) over all host cores."""
    import multiprocessing as mp
    import numpy as np
    from oracle import synth
    rng = np.random.default_rng(seed)
    sims = [synth.sim_matrix(rng, LQ, LR, dim=64, max_copies=2, bias=0.5) for _ in range(n_pairs)]
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        boxes = pool.map(_tn_job, sims, chunksize=max(1, n_pairs // (cores * 4)))
    dt = time.perf_counter() - t0
    return n_pairs / dt, cores, dt, sum(len(b) for b in boxes)


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 64
    rates = []
    for step in range(args.warmup + args.steps):
        rate, cores, dt, _ = cpu_reference_run(sample, seed=100 + step)
        if step >= args.warmup:
            rates.append(rate)
    value = statistics.mean(rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c4_tn_localization", "pairs_per_step": sample, "lq": LQ, "lr": LR, **TN_CFG},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} pairs of {LQ}x{LR} per step; VCSL TN restated on networkx "
                                   f"(real VCSL/FAISS not installable here), multiprocessing.Pool({cores})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def sscd_layerwise_bound(peaks, hw=288):
    """Frames/s bound of the ResNet-50 trunk when every convolution runs at max(tensor time, HBM time): bf16 NHWC
    activations read and written once per layer (implicit GEMM, no cross-layer fusion).  34 of the 53 convolutions
    are HBM-bound at 288x288, so this -- not the 13.5 GFLOP/frame tensor bound -- is the ceiling of a per-layer design."""
    t_peak, h_peak = peaks["bf16_tflops"] * 1e12, peaks["hbm_gbs"] * 1e9
    total = [0.0]

    def conv(cin, cout, k, s, hin, residual=False):
        hout = (hin + 2 * (k // 2) - k) // s + 1
        flops = 2.0 * cin * cout * k * k * hout * hout
        byts = 2.0 * (cin * hin * hin + cout * hout * hout * (2 if residual else 1))
        total[0] += max(flops / t_peak, byts / h_peak)
        return hout

    h = conv(3, 64, 7, 2, hw)
    total[0] += 2.0 * 64 * (h * h + ((h + 1) // 2) ** 2) / h_peak   # max pool
    h = (h + 2 - 3) // 2 + 1
    cin = 64
    for mid, blocks, stride in ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)):
        for b in range(blocks):
            s = stride if b == 0 else 1
            if b == 0:
                conv(cin, mid * 4, 1, s, h)
            conv(cin, mid, 1, 1, h)
            h2 = conv(mid, mid, 3, s, h)
            conv(mid, mid * 4, 1, 1, h2, residual=True)
            h, cin = h2, mid * 4
    return 1.0 / total[0]


def stage_numbers(dev, peaks):
    """Secondary measurements of the other two stages of the path (N=1 only; device-timed, synthetic data)."""
    import torch
    from vsc2022_b200 import gemm
    from vsc2022_b200.index import VideoIndex
    out = {}

    def timed(fn, n):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    # ---- stage B, BASELINE.json configs[2]: 40k query x 200k ref 512-d descriptors, global top-K (K = 1200/query video)
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    nqv, nrv, frames, d = 1250, 6250, 32, 512
    q = torch.randn((nqv * frames, d), generator=g, device=dev).bfloat16().float()   # bf16-representable descriptors
    r = torch.randn((nrv * frames, d), generator=g, device=dev).bfloat16().float()
    for v in range(0, nqv, 20):                                                     # 5 % planted 16-frame copies
        rv = (v * 7919) % nrv
        q[v * frames + 8:v * frames + 24] = r[rv * frames + 4:rv * frames + 20]
    oa, ob = gemm.prepare_pair(q, r)
    flops = 2.0 * q.shape[0] * r.shape[0] * d
    ms = timed(lambda: gemm.gemm_rowmax(oa, ob), 5)
    out["descriptor_gemm_rowmax"] = {
        "shape": [q.shape[0], r.shape[0], d], "ms": ms, "tflops": flops / ms / 1e9,
        "roofline": {"bound": "tensor", "achieved": flops / ms / 1e9, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": flops / ms / 1e9 / peaks["bf16_tflops"]}}
    index = VideoIndex(d)
    index.index.add_device(r)
    K = 1200 * nqv
    ms = timed(lambda: index.global_topk_device(q, K), 3)
    out["descriptor_search_global_topk"] = {
        "workload": "c3: 40k query x 200k ref x 512-d, K=1.5M, FAISS radius schedule (13 batches) + final ordering",
        "ms": ms, "descriptors_per_s": (q.shape[0] + r.shape[0]) / ms * 1e3,
        "gemm_equivalent_tflops": flops / ms / 1e9}
    del q, r, oa, ob, index
    torch.cuda.empty_cache()

    # ---- stage A, BASELINE.json configs[1]: SSCD ResNet-50 on synthetic 288x288 frames (random weights: the
    # checkpoint is a download), bf16 tensor-core GEMMs
    from vsc2022_b200.sscd import SSCDResNet50, TorchReference
    ref = TorchReference(seed=0, device=dev)
    model = SSCDResNet50(ref.trunk, ref.head, device=dev)
    n = 2048
    frames_u8 = torch.randint(0, 256, (n, 288, 288, 3), generator=g, device=dev, dtype=torch.uint8)
    model.forward(frames_u8[:256], batch=128)
    torch.cuda.synchronize(dev)
    passes = []
    for _ in range(3):      # median of three passes: a single pass right after the allocator was emptied is noisy
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.forward(frames_u8, batch=128)
        e1.record()
        torch.cuda.synchronize(dev)
        passes.append(e0.elapsed_time(e1))
    ms = statistics.median(passes)
    tf = n * 13.513e9 / ms / 1e9
    bound = sscd_layerwise_bound(peaks)
    out["sscd_resnet50_inference"] = {
        "workload": "c2 slice: 2048 synthetic 288x288 uint8 frames, batch 128, bf16 (full config: 10k frames)",
        "ms": ms, "passes_ms": passes, "frames_per_s": n / ms * 1e3,
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": tf / peaks["bf16_tflops"], "algorithmic_flops_per_frame": 13.513e9},
        "layerwise_bound": {"frames_per_s": bound, "frac": n / ms * 1e3 / bound,
                            "note": "sum over the 53 convolutions of max(tensor time, HBM time of its bf16 "
                                    "activations); 34 of them are HBM-bound at 288x288"}}
    del frames_u8
    torch.cuda.empty_cache()

    # ---- BASELINE.json configs[4] at single-GPU test size: frames -> SSCD -> score-norm search -> TN localization
    # through the reference-shaped host API (numpy VideoFeatures between the stages, like the reference's .npz files)
    from vsc2022_b200 import inference_impl, sscd_baseline
    from vsc2022_b200.score_normalization import score_normalize
    import numpy as np
    nq, nr, nn, fr = 50, 400, 50, 40
    ts = np.stack([np.arange(fr) * 1.0, np.arange(fr) * 1.0 + 1.0], axis=1)
    make = lambda prefix, count: [(f"{prefix}{i:06d}", ts, torch.randint(0, 256, (fr, 288, 288, 3), generator=g, device=dev,
                                                                       dtype=torch.uint8)) for i in range(count)]
    refs_v, queries_v, noise_v = make("R", nr), make("Q", nq), make("N", nn)
    for i in range(0, nq, 2):                      # every second query carries a 20-frame copy of a reference
        queries_v[i][2][8:28] = refs_v[(i * 37) % nr][2][4:24]
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    feats = [inference_impl.infer_videos(v, model, batch_size=128, device=dev) for v in (queries_v, refs_v, noise_v)]
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    sn_q, sn_r = score_normalize(feats[0], feats[1], feats[2], beta=1.2)
    cands = sscd_baseline.search(sn_q, sn_r)
    torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    matches = sscd_baseline.localize_and_verify(sn_q, sn_r, cands, score_normalization=True)
    torch.cuda.synchronize(dev)
    t3 = time.perf_counter()
    planted = {(f"Q{i:06d}", f"R{(i * 37) % nr:06d}") for i in range(0, nq, 2)}
    found = {(m.query_id, m.ref_id) for m in matches}
    n_frames = (nq + nr + nn) * fr
    out["pipeline_frames_to_matches"] = {
        "workload": f"c5 at 1-GPU test size: {nq} query + {nr} ref + {nn} noise videos x {fr} frames of 288x288 -> SSCD -> "
                    "score-norm + global top-K candidates -> TN localization (host API, numpy features between stages)",
        "seconds": {"descriptors": t1 - t0, "score_norm_and_search": t2 - t1, "localize": t3 - t2, "total": t3 - t0},
        "frames_per_s_descriptors": n_frames / (t1 - t0), "frames_per_s_total": n_frames / (t3 - t0),
        "candidates": len(cands), "pairs_localized": min(len(cands), 5 * nq), "matches": len(matches),
        "planted_pairs_found": len(planted & found), "planted_pairs": len(planted)}
    return out


def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    from vsc2022_b200 import _lib, vta, workloads

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    n = PAIRS_PER_GPU
    w = workloads.tn_pairs_device(n, LQ, LR, seed=4 + rank, device=dev)
    model = vta.build_vta_model("TN", concurrency=16, **{k: v for k, v in TN_CFG.items()})
    stream = torch.cuda.current_stream(dev)

    def device_step():
        return model.align_device(w.sims, w.off, w.lq, w.lr, n, LQ, LR, want_maxsim=False)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: device-resident inputs
    for _ in range(max(args.warmup, 3)):
        res = device_step()
    lib.vsc_tn_set_profiling(1)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        res = device_step()
    ev1.record(stream)
    barrier()
    sampler.stop_flag = True
    launches = _lib.launch_count() - launches0
    ms = ev0.elapsed_time(ev1) / args.steps
    import ctypes
    stage = (ctypes.c_float * 4)()
    lib.vsc_tn_last_stage_ms(stage)
    lib.vsc_tn_set_profiling(0)
    _, n_boxes, _, status = res.to_host()

    # ---------------- e2e: pinned host buffers through the host-facing API
    try:
        host_sims = torch.empty(w.sims.numel() + 4, dtype=torch.float32, pin_memory=True)
        pinned = True
    except RuntimeError:   # N ranks pin N x 2.88 GB; a host that refuses still gets an (honest, slower) e2e number
        host_sims = torch.empty(w.sims.numel() + 4, dtype=torch.float32)
        pinned = False
    host_sims[:w.sims.numel()].copy_(w.sims)
    off_h, lq_h, lr_h = w.off.cpu().numpy(), w.lq.cpu().numpy(), w.lr.cpu().numpy()
    for _ in range(2):
        boxes_h = model.forward_packed(host_sims, off_h, lq_h, lr_h)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        boxes_h = model.forward_packed(host_sims, off_h, lq_h, lr_h)
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = w.sims.numel() * 4 + off_h.nbytes + lq_h.nbytes + lr_h.nbytes
    d2h = n * (TN_CFG["max_path"] + 1) * 4 * 4 + n * 4

    # ---------------- max over ranks
    times = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = times.tolist()

    if rank == 0:
        peaks, peak_src = measured_peaks()
        algo_bytes = n * 4 * LQ * LR
        achieved = algo_bytes / (ms_max * 1e-3) / 1e9
        stages = None
        if world == 1 and not args.no_stages:
            del host_sims
            stages = stage_numbers(dev, peaks)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, dt, _ = cpu_reference_run(96, seed=4)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"96 pairs of {LQ}x{LR} ({dt:.1f} s): VCSL TN restated on networkx "
                             f"(real VCSL not installable here), multiprocessing.Pool({cores})"}
        line = {
            "metric": METRIC, "value": world * n / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "c4_tn_localization", "pairs_per_gpu": n, "lq": LQ, "lr": LR, **TN_CFG,
                       "l2": "inputs (2.88 GB/GPU) larger than L2; no flush needed",
                       "parallelism": f"pairs sharded over {world} GPU(s), no data-path collective"},
            "clocks": sampler.summary(),
            "e2e": {"value": world * n / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_max,
                    "api": "vsc2022_b200.vta.TN.forward_packed (%s host similarity matrices in, boxes out)"
                           % ("pinned" if pinned else "pageable")},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "peak_source": peak_src,
                         "kernel": "TN pipeline of one vcsl_tn_batch call (tn_topk + tn_edges + tn_dp)",
                         "algorithmic_bytes_per_launch": algo_bytes,
                         "traffic": TRAFFIC_BYTES_PER_CALL,
                         "stages_ms": {"tn_topk_kernel": stage[0], "tn_edges_kernel": stage[1],
                                       "tn_dp_kernel": stage[2]},
                         "tn_topk_frac": algo_bytes / (stage[0] * 1e-3) / 1e9 / peaks["hbm_gbs"] if stage[0] > 0 else None},
            "cpu_baseline": cpu,
            "stages": stages,
            "result_check": {"boxes_per_pair": float(np.mean(n_boxes)),
                             "pairs_on_fast_pipeline": int((status == 0).sum()),
                             "pairs_on_general_kernel": int((status == 2).sum()),
                             "pairs_on_exact_order_kernel": int((status == 1).sum())},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum over the three kernels of one call, from the ncu capture
# committed under profiles/ (profiles/r01_tn_launches_final.csv, per-launch table r01_tn_launches_final.txt).
TRAFFIC_BYTES_PER_CALL = 3.88e9  # tn_topk 3.014+0.213, tn_edges 0.152+0.058, tn_dp 0.326+0.116 GB (profiles/r01_tn_summary.md)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stages", action="store_true", help="skip the secondary stage A / stage B measurements")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            sys.exit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
