"""`vsc.baseline.inference_impl` mirror for the descriptor stage (inference_impl.py:103-109, 210-253).

What is mirrored: `run_inference(dataloader, model, device, store_fp16)` (single-video batches of
`{"name", "timestamp", "input"}` -> one `VideoFeature` per video), the `i % world_size == rank` video sharding of
`VideoDataset` and `merge_feature_files`.  What is not: ffmpeg decoding and the PIL resize --
frames arrive decoded.  `ToTensor` + `Normalize` (inference_impl.py:39-69) run on the GPU when frames are uint8 NHWC.

`infer_videos` is the B200-shaped entry point: the reference feeds the model at most one video (<= 32 frames) per
forward; here frames of consecutive videos are packed into full device batches (the forward is batch-invariant,
tests/test_sscd_gpu.py::test_batching_is_transparent) and regrouped per video afterwards, with a single device->host
copy per packed batch instead of one synchronous `.cpu()` per video chunk.
"""
from typing import Callable, Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from .index import VideoFeature
from .storage import load_features, store_features


def frame_timestamps(n_frames: int, original_fps=None) -> np.ndarray:
    """The [start, end] timestamps the reference attaches to decoded frame i: ((i) / original_fps, (i + 1) / original_fps)
    with original_fps = max(1, reader.fps) if reader.fps else 1 (video_reader/video_reader.py:18) -- and since
    FFMpegVideoReader.fps is always None (ffmpeg_video_reader.py:28-30), every frame extracted at `--fps` gets the
    interval [i, i + 1] whatever the sampling rate.  Reproduced as it is: downstream timestamps (Match rows, metrics)
    depend on it."""
    fps = max(1, original_fps) if original_fps else 1
    i = np.arange(n_frames, dtype=np.float64)
    return np.stack([i / fps, (i + 1) / fps], axis=1)


def select_videos(videos: Sequence, rank: int = 0, world_size: int = 1) -> List[Tuple[int, object]]:
    """`VideoDataset.selected_videos` (inference_impl.py:103-109): video i belongs to rank i % world_size."""
    assert rank < world_size
    return [(i, v) for i, v in enumerate(videos) if i % world_size == rank]


def _to_numpy(features, store_fp16: bool) -> np.ndarray:
    import torch
    if isinstance(features, torch.Tensor):
        features = features.detach().cpu()
        if store_fp16:
            features = features.half()
        return features.numpy()
    features = np.asarray(features)
    return features.astype(np.float16) if store_fp16 else features


def run_inference(dataloader: Iterable[dict], model: Callable, device=None, store_fp16: bool = False) -> Iterator[VideoFeature]:
    """Same contract as the reference generator: consecutive batches with the same `name` form one video."""
    import torch
    name = None
    embeddings: List[np.ndarray] = []
    timestamps: List[np.ndarray] = []
    with torch.no_grad():
        for batch in dataloader:
            names = batch["name"]
            if isinstance(names, str):
                names = [names]
            assert names[0] == names[-1]  # single-video batches
            if name is not None and name != names[0]:
                yield VideoFeature(video_id=name, timestamps=np.concatenate(timestamps, axis=0),
                                   feature=np.concatenate(embeddings, axis=0))
                embeddings, timestamps = [], []
            name = names[0]
            img = batch["input"]
            if device is not None and isinstance(img, torch.Tensor):
                img = img.to(device)
            embeddings.append(_to_numpy(model(img), store_fp16))
            timestamps.append(np.asarray(batch["timestamp"]))
    if name is not None:
        yield VideoFeature(video_id=name, timestamps=np.concatenate(timestamps, axis=0),
                           feature=np.concatenate(embeddings, axis=0))


def infer_videos(videos: Sequence[Tuple[object, np.ndarray, object]], model: Callable, batch_size: int = 128,
                 store_fp16: bool = False, device=None, on_device: bool = False, transform=None) -> List[VideoFeature]:
    """videos: (video_id, timestamps [n] or [n, 2], frames) with frames uint8 [n, H, W, 3] or normalised float32
    [n, 3, H, W] (numpy or torch).  Frames of all videos must share one geometry (after `transform`).  Returns one
    VideoFeature per video, in order, identical to feeding every video on its own.  on_device=True (extension): the
    descriptors stay on the GPU -- every VideoFeature holds a row view of one device matrix, ready for score_normalize /
    CandidateGeneration.  transform: a preprocess.GpuTransform (build_transforms, inference_impl.py:39-69) applied to the
    decoded uint8 frames of every video on the device before they are packed into batches."""
    import torch
    if transform is not None:
        model = _TransformedModel(model, transform)
    if on_device:
        return _infer_videos_device(videos, model, batch_size, store_fp16, device)
    counts = [int(len(f)) for _, _, f in videos]
    out: List[Optional[np.ndarray]] = [None] * len(videos)
    pending: List[Tuple[int, int, int]] = []      # (video, first frame, frames) of the batch being packed
    chunks: List[List[np.ndarray]] = [[] for _ in videos]

    def flush():
        if not pending:
            return
        parts = []
        for v, s, n in pending:
            fr = videos[v][2][s:s + n]
            parts.append(fr if isinstance(fr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(fr)))
        frames = torch.cat(parts)
        if device is not None:
            frames = frames.to(device, non_blocking=True)
        feats = _to_numpy(model(frames), store_fp16)
        at = 0
        for v, _, n in pending:
            chunks[v].append(feats[at:at + n])
            at += n
        pending.clear()

    room = batch_size
    for v, n in enumerate(counts):
        s = 0
        while s < n:
            take = min(room, n - s)
            pending.append((v, s, take))
            s += take
            room -= take
            if room == 0:
                flush()
                room = batch_size
    flush()
    dim = next((c[0].shape[1] for c in chunks if c), 0)
    for v, (vid, ts, _) in enumerate(videos):
        feat = np.concatenate(chunks[v], axis=0) if chunks[v] else np.zeros((0, dim), np.float16 if store_fp16 else np.float32)
        out[v] = VideoFeature(video_id=vid, timestamps=np.asarray(ts), feature=feat)
    return out  # type: ignore[return-value]


class _TransformedModel:
    """model(transform(frames)).  Packed batches hold frames of ONE source geometry (the caller's contract), so the
    whole batch is resized in one launch pair."""

    def __init__(self, model, transform):
        self.model, self.transform = model, transform

    def __call__(self, frames):
        return self.model(self.transform(frames))


def _infer_videos_device(videos, model, batch_size, store_fp16, device):
    import torch
    counts = [int(len(f)) for _, _, f in videos]
    total = sum(counts)
    out_mat = None
    flat = [(v, s) for v, n in enumerate(counts) for s in range(0, n, 1)]   # frame -> (video, frame): packed batches
    at = 0
    while at < total:
        take = min(batch_size, total - at)
        parts, v, s = [], flat[at][0], flat[at][1]
        left = take
        while left:
            n = min(left, counts[v] - s)
            fr = videos[v][2][s:s + n]
            parts.append(fr if isinstance(fr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(fr)))
            left -= n
            v, s = v + 1, 0
        frames = torch.cat(parts) if len(parts) > 1 else parts[0]
        if device is not None:
            frames = frames.to(device, non_blocking=True)
        feats = model(frames)
        if out_mat is None:
            out_mat = torch.empty((total, feats.shape[1]), dtype=torch.float16 if store_fp16 else feats.dtype, device=feats.device)
        out_mat[at:at + take] = feats
        at += take
    if out_mat is None:
        out_mat = torch.zeros((0, 0), dtype=torch.float32, device=device)
    out, at = [], 0
    for (vid, ts, _), n in zip(videos, counts):
        out.append(VideoFeature(video_id=vid, timestamps=np.asarray(ts), feature=out_mat[at:at + n]))
        at += n
    return out


def merge_feature_files(filenames: List[str], output_filename: str) -> int:
    """inference_impl.py:242-247: concatenate the per-rank feature files."""
    features: List[VideoFeature] = []
    for fn in filenames:
        features.extend(load_features(fn))
    store_features(output_filename, features)
    return len(features)
