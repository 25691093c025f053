#!/usr/bin/env python3
"""`python -m vsc.baseline.inference` mirror (vsc/baseline/inference.py:28-175, inference_impl.py:72-207): descriptors of a
directory of videos with the SSCD model on B200.

    python -m vsc2022_b200.inference --torchscript_path sscd_disc_mixup.no_l2_norm.torchscript.pt --accelerator cuda \
        --dataset_path videos/ --output_file out/descriptors.npz [--processes 8] [--transforms RESIZE_320_CENTER] \
        [--batch_size 32] [--fps 1] [--store_fp16] [--video_extensions mp4] [--ffmpeg_path ffmpeg] [--scratch_path tmp]

Same flags, defaults, sharding (video i -> rank i % world), per-rank scratch files merged into `--output_file`, and log
lines as the reference.  What runs differently: frames are decoded by the same ffmpeg command line but stay uint8; resize /
crop (`--transforms`), ToTensor + Normalize and the ResNet-50 -> GeM -> Linear model run on the GPU (csrc/resize.cu,
sscd_ops.cu, gemm_tc.cu); frames of consecutive videos are packed into full batches of `--batch_size`.  The TorchScript file
is only read for its weights (`sscd.load_torchscript`).  There is no CPU path: `--accelerator cpu` (the reference's default)
is refused, and `--baseline dns / dino` are outside the vsc2022 hot path.
"""
import argparse
import enum
import glob
import logging
import multiprocessing
import os
import tempfile

import numpy as np

from .preprocess import InferenceTransforms


class Accelerator(enum.Enum):
    CPU = enum.auto()
    CUDA = enum.auto()


class VideoReaderType(enum.Enum):
    FFMPEG = enum.auto()


class Baseline(enum.Enum):
    SSCD = enum.auto()
    DNS = enum.auto()
    DINO = enum.auto()


parser = argparse.ArgumentParser()
inference_parser = parser.add_argument_group("Inference")
inference_parser.add_argument("--baseline", default="sscd", choices=[x.name.lower() for x in Baseline])
inference_parser.add_argument("--torchscript_path", default=None)
inference_parser.add_argument("--batch_size", type=int, default=32)
inference_parser.add_argument("--distributed_rank", type=int, default=0)
inference_parser.add_argument("--distributed_size", type=int, default=1)
inference_parser.add_argument("--processes", type=int, default=1)
inference_parser.add_argument("--transforms", choices=[x.name for x in InferenceTransforms], default="RESIZE_320_CENTER")
inference_parser.add_argument("--accelerator", choices=[x.name.lower() for x in Accelerator], default="cpu")
inference_parser.add_argument("--output_file", required=True)
inference_parser.add_argument("--scratch_path", required=False)
inference_parser.add_argument("--store_fp16", action="store_true")
dataset_parser = parser.add_argument_group("Dataset")
dataset_parser.add_argument("--dataset_path", required=True)
dataset_parser.add_argument("--fps", default=1, type=float)
dataset_parser.add_argument("--video_extensions", default="mp4")
dataset_parser.add_argument("--video_reader", choices=[x.name for x in VideoReaderType], default="FFMPEG")
dataset_parser.add_argument("--ffmpeg_path", default="ffmpeg")

logging.basicConfig(format="%(asctime)s %(levelname)-8s %(message)s", level=logging.INFO, datefmt="%Y-%m-%d %H:%M:%S")
logger = logging.getLogger("inference.py")
logger.setLevel(logging.INFO)


def list_videos(path: str, extensions):
    """VideoDataset.__init__ (inference_impl.py:95-102): sorted files of the wanted extensions."""
    if len(extensions) == 1:
        filenames = glob.glob(os.path.join(path, f"*.{extensions[0]}"))
    else:
        filenames = [fn for fn in glob.glob(os.path.join(path, "*.*")) if fn.rsplit(".", 1)[-1] in extensions]
    videos = sorted(filenames)
    if not videos:
        raise Exception("No videos found!")
    return videos


def decode_video(video: str, fps: float, video_reader: VideoReaderType, ffmpeg_path: str):
    """(name, timestamps [n, 2], frames uint8 [n, H, W, 3]) of one video (VideoDataset.read_frames, :126-143)."""
    from .video_reader import FFMpegVideoReader
    name = os.path.basename(video).split(".")[0]
    if video_reader != VideoReaderType.FFMPEG:
        raise ValueError(f"VideoReaderType: {video_reader} not supported")
    reader = FFMpegVideoReader(video_path=video, required_fps=fps, ffmpeg_path=ffmpeg_path)
    stamps, frames = [], []
    for start, end, frame in reader.frames():
        stamps.append((start, end))
        frames.append(frame)
    ts = np.array(stamps, dtype=np.float64).reshape(-1, 2)
    return name, ts, (np.stack(frames) if frames else np.zeros((0, 1, 1, 3), np.uint8))


def get_device(args, rank, world_size):
    """inference_impl.py:146-166, CUDA only."""
    import torch
    if Accelerator[args.accelerator.upper()] != Accelerator.CUDA:
        from ._lib import EngineError
        raise EngineError("vsc2022_b200.inference runs on CUDA only (there is no CPU path): pass --accelerator cuda")
    assert torch.cuda.is_available()
    num_devices = torch.cuda.device_count()
    if args.processes > num_devices:
        raise Exception(f"Asked for {args.processes} processes and cuda, but only {num_devices} devices found")
    device_num = rank if (args.processes > 1 or world_size <= num_devices) else 0
    torch.cuda.set_device(device_num)
    return torch.device("cuda", device_num)


def worker_process(args, rank, world_size, output_filename):
    from . import inference_impl, sscd
    from .preprocess import build_transforms
    from .storage import store_features
    logger.info(f"Starting worker {rank} of {world_size}.")
    if Baseline[args.baseline.upper()] != Baseline.SSCD:
        raise NotImplementedError(f"--baseline {args.baseline}: only the SSCD baseline is on the vsc2022 hot path")
    device = get_device(args, rank, world_size)
    logger.info("Loading model")
    model = sscd.load_torchscript(args.torchscript_path, device=device)
    logger.info("Setting up dataset")
    transform = build_transforms(InferenceTransforms[args.transforms], device=device)
    videos = list_videos(args.dataset_path, args.video_extensions.split(","))
    mine = inference_impl.select_videos(videos, rank, world_size)
    reader = VideoReaderType[args.video_reader.upper()]
    vfs = []
    # videos of one geometry are packed into full batches; a change of geometry flushes what has been decoded so far
    pending, geometry = [], None

    def flush():
        if pending:
            vfs.extend(inference_impl.infer_videos(pending, model, batch_size=args.batch_size, store_fp16=args.store_fp16,
                                                   device=device, transform=transform))
            pending.clear()
    for _, video in mine:
        name, ts, frames = decode_video(video, args.fps, reader, args.ffmpeg_path)
        if len(frames) == 0:
            continue                      # the reference yields nothing for a video without frames
        if geometry is not None and frames.shape[1:] != geometry:
            flush()
        geometry = frames.shape[1:]
        pending.append((name, ts, frames))
        if sum(len(f) for _, _, f in pending) >= 8 * args.batch_size:
            flush()
    flush()
    logger.info(f"Storing worker {rank} outputs")
    store_features(output_filename, vfs)
    logger.info(f"Wrote worker {rank} features for {len(vfs)} videos to {output_filename}")


def distributed_worker_process(args, rank, world_size, backend, output_filename):
    from torch import distributed
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = "19529"
    distributed.init_process_group(backend, rank=rank, world_size=world_size)
    worker_process(args, rank, world_size, output_filename)


def main(args):
    success = False
    if args.processes > 1 and args.distributed_size > 1:
        raise Exception("Set either --processes (single-machine distributed) or both --distributed_size and "
                        "--distributed_rank (arbitrary distributed)")
    with tempfile.TemporaryDirectory() as tmp_path:
        os.makedirs(os.path.dirname(args.output_file), exist_ok=True)
        if args.scratch_path:
            os.makedirs(args.scratch_path, exist_ok=True)
        else:
            args.scratch_path = tmp_path
        if args.processes > 1:
            processes = []
            logger.info(f"Spawning {args.processes} processes")
            backend = "nccl" if Accelerator[args.accelerator.upper()] == Accelerator.CUDA else "gloo"
            ctx = multiprocessing.get_context("spawn")
            worker_files = []
            try:
                for rank in range(args.processes):
                    worker_file = os.path.join(args.scratch_path, f"{rank}.npz")
                    worker_files.append(worker_file)
                    p = ctx.Process(target=distributed_worker_process, args=(args, rank, args.processes, backend, worker_file))
                    processes.append(p)
                    p.start()
                worker_success = []
                for p in processes:
                    p.join()
                    worker_success.append(p.exitcode == os.EX_OK)
                success = all(worker_success)
            finally:
                for p in processes:
                    p.kill()
            if success:
                from .inference_impl import merge_feature_files
                num_files = merge_feature_files(worker_files, args.output_file)
                logger.info(f"Features for {num_files} videos saved to {args.output_file}")
        else:
            worker_process(args, args.distributed_rank, args.distributed_size, args.output_file)
            success = True
    if success:
        logger.info("Inference succeeded.")
    else:
        logger.error("Inference FAILED!")


if __name__ == "__main__":
    main(parser.parse_args())
