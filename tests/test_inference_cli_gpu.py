"""`python -m vsc2022_b200.inference` end to end on the GPU: a traced TorchScript model in the adapted SSCD layout
(adapt_sscd_model.py:56-70) + a directory of "videos" decoded through a stand-in ffmpeg -> descriptors.npz, against the
reference's own stack computed in the test: PIL frames -> torchvision Compose (Resize, CenterCrop, ToTensor, Normalize) ->
the same TorchScript model in fp32 PyTorch (inference_impl.py:39-69, 173, 210-239)."""
import collections
import os
import stat
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _traced_sscd(path, seed=3):
    import torch
    import torchvision

    class GlobalGeMPool2d(torch.nn.Module):
        def forward(self, x):
            return x.clamp(min=1e-6).pow(3.0).mean(dim=(2, 3)).pow(1.0 / 3.0)

    torch.manual_seed(seed)
    resnet = torchvision.models.resnet50(weights=None)
    for mod in resnet.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.normal_(0, 0.1)
    model = torch.nn.Sequential(collections.OrderedDict([
        ("backbone", torch.nn.Sequential(*list(resnet.children())[:-2])), ("pool", GlobalGeMPool2d()),
        ("project", torch.nn.Linear(2048, 512))])).eval()
    torch.jit.save(torch.jit.trace(model, torch.randn(2, 3, 64, 64)), path)
    return model


def test_cli_end_to_end(tmp_path):
    import torch
    from PIL import Image
    from torchvision import transforms
    from vsc2022_b200 import inference
    from vsc2022_b200.storage import load_features
    exe = tmp_path / "ffmpeg"
    exe.write_text(f"#!/bin/sh\nexec {sys.executable} {os.path.join(HERE, 'helpers', 'fake_ffmpeg.py')} \"$@\"\n")
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    model = _traced_sscd(str(tmp_path / "sscd.pt"))
    rng = np.random.default_rng(4)
    data = tmp_path / "videos"
    data.mkdir()
    videos = {}
    for name, n in (("R200002", 7), ("R200001", 3), ("R200003", 5)):
        frames = rng.integers(0, 256, size=(n, 90, 160, 3), dtype=np.uint8)
        with open(data / f"{name}.mp4", "wb") as f:
            np.save(f, frames)
        videos[name] = frames
    common = ["--torchscript_path", str(tmp_path / "sscd.pt"), "--accelerator", "cuda", "--dataset_path", str(data),
              "--ffmpeg_path", str(exe), "--batch_size", "4", "--transforms", "RESIZE_320_CENTER"]
    out = tmp_path / "out" / "descriptors.npz"
    inference.main(inference.parser.parse_args(common + ["--output_file", str(out)]))
    feats = load_features(str(out))
    assert [f.video_id for f in feats] == ["R200001", "R200002", "R200003"]         # sorted file order
    compose = transforms.Compose([transforms.Resize(320), transforms.CenterCrop(320), transforms.ToTensor(),
                                  transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    for f in feats:
        frames = videos[f.video_id]
        assert f.timestamps.tolist() == [[float(i), float(i + 1)] for i in range(len(frames))]
        with torch.no_grad():
            want = model(torch.stack([compose(Image.fromarray(x)) for x in frames])).numpy()
        got = f.feature
        assert got.dtype == np.float32 and got.shape == want.shape
        cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
        rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
        assert cos.min() >= 0.999 and rel.max() <= 3e-2, (f.video_id, cos.min(), rel.max())
    # --distributed_rank / --distributed_size: video i -> rank i % 2; merged per-rank files == the single run; --store_fp16
    parts = []
    for rank in range(2):
        part = tmp_path / "scratch" / f"{rank}.npz"
        inference.main(inference.parser.parse_args(common + ["--output_file", str(part), "--distributed_rank", str(rank),
                                                             "--distributed_size", "2", "--store_fp16"]))
        parts.append(str(part))
    assert [f.video_id for f in load_features(parts[0])] == ["R200001", "R200003"]
    from vsc2022_b200.inference_impl import merge_feature_files
    merged = tmp_path / "merged.npz"
    assert merge_feature_files(parts, str(merged)) == 3
    by_id = {f.video_id: f for f in load_features(str(merged))}
    for f in feats:
        assert by_id[f.video_id].feature.dtype == np.float16
        assert np.array_equal(by_id[f.video_id].feature, f.feature.astype(np.float16))
