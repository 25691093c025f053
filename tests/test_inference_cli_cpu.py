"""`python -m vsc2022_b200.inference` (mirror of vsc/baseline/inference.py + inference_impl.py:72-207): flag surface, video
listing, the ffmpeg reader (against a stand-in binary) and the refusal to run without CUDA.  CPU only."""
import os
import stat
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture
def ffmpeg(tmp_path):
    """An executable named `ffmpeg` that runs tests/helpers/fake_ffmpeg.py with this interpreter."""
    exe = tmp_path / "ffmpeg"
    exe.write_text(f"#!/bin/sh\nexec {sys.executable} {os.path.join(HERE, 'helpers', 'fake_ffmpeg.py')} \"$@\"\n")
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    return str(exe)


def _video(path, n, h=24, w=40, seed=0):
    frames = np.random.default_rng(seed).integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    with open(path, "wb") as f:
        np.save(f, frames)
    return frames


def test_flags_and_defaults_match_the_reference():
    """vsc/baseline/inference.py:50-81."""
    from vsc2022_b200 import inference
    args = inference.parser.parse_args(["--output_file", "o.npz", "--dataset_path", "d"])
    assert vars(args) == dict(
        baseline="sscd", torchscript_path=None, batch_size=32, distributed_rank=0, distributed_size=1, processes=1,
        transforms="RESIZE_320_CENTER", accelerator="cpu", output_file="o.npz", scratch_path=None, store_fp16=False,
        dataset_path="d", fps=1, video_extensions="mp4", video_reader="FFMPEG", ffmpeg_path="ffmpeg")
    with pytest.raises(SystemExit):
        inference.parser.parse_args(["--dataset_path", "d"])                       # --output_file is required
    with pytest.raises(SystemExit):
        inference.parser.parse_args(["--output_file", "o", "--dataset_path", "d", "--transforms", "RESIZE_999"])
    with pytest.raises(SystemExit):
        inference.parser.parse_args(["--output_file", "o", "--dataset_path", "d", "--baseline", "resnet"])


def test_list_videos(tmp_path):
    from vsc2022_b200 import inference
    for name in ("b.mp4", "a.mp4", "c.mkv", "notes.txt"):
        (tmp_path / name).write_bytes(b"x")
    assert [os.path.basename(v) for v in inference.list_videos(str(tmp_path), ["mp4"])] == ["a.mp4", "b.mp4"]
    assert [os.path.basename(v) for v in inference.list_videos(str(tmp_path), ["mp4", "mkv"])] == ["a.mp4", "b.mp4", "c.mkv"]
    with pytest.raises(Exception, match="No videos found"):
        inference.list_videos(str(tmp_path), ["avi"])


def test_ffmpeg_reader_frames_and_timestamps(tmp_path, ffmpeg):
    """ffmpeg_video_reader.py:28-54: frames numbered from 0, interval [i, i + 1] per frame whatever --fps is."""
    from vsc2022_b200 import inference
    from vsc2022_b200.video_reader import FFMpegVideoReader
    frames = _video(tmp_path / "Q100001.mp4", 5)
    got = list(FFMpegVideoReader(str(tmp_path / "Q100001.mp4"), required_fps=1, ffmpeg_path=ffmpeg).frames())
    assert [(a, b) for a, b, _ in got] == [(0.0, 1.0), (1.0, 2.0), (2.0, 3.0), (3.0, 4.0), (4.0, 5.0)]
    assert all(np.array_equal(f, frames[i]) for i, (_, _, f) in enumerate(got))
    name, ts, dec = inference.decode_video(str(tmp_path / "Q100001.mp4"), 2.0, inference.VideoReaderType.FFMPEG, ffmpeg)
    assert name == "Q100001" and dec.shape == (10, 24, 40, 3) and dec.dtype == np.uint8
    assert ts.tolist() == [[float(i), float(i + 1)] for i in range(10)]            # original_fps == 1: the reference's quirk
    assert np.array_equal(dec[3], frames[1])


def test_cpu_accelerator_is_refused(tmp_path, ffmpeg):
    from vsc2022_b200 import inference
    from vsc2022_b200._lib import EngineError
    _video(tmp_path / "a.mp4", 2)
    args = inference.parser.parse_args(["--output_file", str(tmp_path / "out" / "f.npz"), "--dataset_path", str(tmp_path),
                                        "--ffmpeg_path", ffmpeg, "--torchscript_path", "none.pt"])
    with pytest.raises(EngineError, match="CUDA only"):
        inference.main(args)
    args.baseline, args.accelerator = "dino", "cuda"
    with pytest.raises(NotImplementedError):
        inference.main(args)
    args.baseline, args.processes, args.distributed_size = "sscd", 2, 2
    with pytest.raises(Exception, match="Set either --processes"):
        inference.main(args)
