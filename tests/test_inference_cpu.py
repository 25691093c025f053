"""Host logic of the stage-A mirror (vsc2022_b200/inference_impl.py) with a stand-in model: no GPU needed."""
import numpy as np

from vsc2022_b200 import inference_impl
from vsc2022_b200.storage import load_features, store_features


def fake_model(frames):
    """A per-frame function (batch-invariant like the real forward): 4 'descriptor' dims from pixel statistics."""
    import torch
    x = torch.as_tensor(frames).float().reshape(len(frames), -1)
    return torch.stack([x.mean(1), x.std(1), x.min(1).values, x.max(1).values], dim=1)


def make_videos(rng, lengths):
    vids = []
    for i, n in enumerate(lengths):
        frames = rng.integers(0, 256, size=(n, 8, 8, 3), dtype=np.uint8)
        ts = np.stack([np.arange(n, dtype=np.float64), np.arange(n, dtype=np.float64) + 1.0], axis=1)
        vids.append((f"Q{i:05d}", ts, frames))
    return vids


def loader(videos, batch_size):
    """What the reference's VideoDataset + DataLoader yield: single-video batches of at most batch_size frames."""
    for name, ts, frames in videos:
        for s in range(0, len(frames), batch_size):
            yield {"name": [name] * len(frames[s:s + batch_size]), "timestamp": ts[s:s + batch_size],
                   "input": frames[s:s + batch_size]}


def test_video_sharding_matches_reference_rule():
    vids = [f"v{i}" for i in range(11)]
    shards = [inference_impl.select_videos(vids, r, 4) for r in range(4)]
    assert [[i for i, _ in s] for s in shards] == [[0, 4, 8], [1, 5, 9], [2, 6, 10], [3, 7]]
    assert sorted(v for s in shards for _, v in s) == sorted(vids)


def test_run_inference_groups_per_video_and_packed_path_is_identical():
    rng = np.random.default_rng(0)
    videos = make_videos(rng, [5, 70, 1, 33, 32, 64])
    per_video = list(inference_impl.run_inference(loader(videos, 32), fake_model, None, store_fp16=False))
    assert [v.video_id for v in per_video] == [name for name, _, _ in videos]
    packed = inference_impl.infer_videos(videos, fake_model, batch_size=48)
    for a, b, (_, ts, frames) in zip(per_video, packed, videos):
        assert a.video_id == b.video_id and len(a) == len(frames)
        assert np.array_equal(a.timestamps, ts) and np.array_equal(b.timestamps, ts)
        assert a.feature.dtype == np.float32 and np.array_equal(a.feature, b.feature)


def test_store_fp16_and_merge(tmp_path):
    rng = np.random.default_rng(1)
    videos = make_videos(rng, [9, 17, 4])
    feats = inference_impl.infer_videos(videos, fake_model, batch_size=16, store_fp16=True)
    assert all(f.feature.dtype == np.float16 for f in feats)
    files = []
    for rank in range(2):   # one file per rank, like inference.py's workers
        mine = [feats[i] for i, _ in inference_impl.select_videos(feats, rank, 2)]
        fn = str(tmp_path / f"rank{rank}.npz")
        store_features(fn, mine)
        files.append(fn)
    merged = str(tmp_path / "merged.npz")
    assert inference_impl.merge_feature_files(files, merged) == 3
    back = {v.video_id: v for v in load_features(merged)}
    for f in feats:
        assert np.array_equal(back[f.video_id].feature, f.feature)
        assert np.array_equal(back[f.video_id].timestamps, f.timestamps)


def test_frame_timestamps_quirk():
    """ffmpeg_video_reader.py:54 with original_fps == 1 (video_reader.py:18, FFMpegVideoReader.fps is None)."""
    from vsc2022_b200.inference_impl import frame_timestamps
    ts = frame_timestamps(4)
    assert ts.tolist() == [[0.0, 1.0], [1.0, 2.0], [2.0, 3.0], [3.0, 4.0]]
    assert frame_timestamps(2, original_fps=0.5).tolist() == [[0.0, 1.0], [1.0, 2.0]]      # max(1, fps)
    assert frame_timestamps(2, original_fps=4).tolist() == [[0.0, 0.25], [0.25, 0.5]]
    assert frame_timestamps(0).shape == (0, 2)
