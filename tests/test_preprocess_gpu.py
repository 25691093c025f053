"""GPU frame resize (csrc/resize.cu through vsc2022_b200.preprocess) against the reference's own transform stack:
torchvision.transforms on PIL images (vsc/baseline/inference_impl.py:39-69).  Bar: bit-exact uint8 pixels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GEOMETRIES = [(360, 640), (640, 360), (288, 288), (300, 288), (1080, 1920), (97, 131), (720, 406), (32, 500)]


def _tv(t):
    from torchvision import transforms
    from vsc2022_b200.preprocess import InferenceTransforms as T
    return {T.RESIZE_288: transforms.Resize(288),
            T.RESIZE_320_CENTER: transforms.Compose([transforms.Resize(320), transforms.CenterCrop(320)]),
            T.RESIZE_224_SQUARE: transforms.Resize((224, 224))}[t]


@pytest.mark.parametrize("h,w", GEOMETRIES)
def test_resize_equals_pil_bit_for_bit(h, w):
    from PIL import Image
    from vsc2022_b200.preprocess import InferenceTransforms as T, build_transforms
    rng = np.random.default_rng(h * 31 + w)
    frames = rng.integers(0, 256, size=(3, h, w, 3), dtype=np.uint8)
    frames[1, : h // 2] = (frames[1, : h // 2] // 128) * 255            # saturated edges
    frames[2] = np.linspace(0, 255, w, dtype=np.uint8)[None, :, None]   # smooth ramp: rounding of near-exact values
    for t in T:
        got = build_transforms(t)(frames).cpu().numpy()
        for i in range(3):
            want = np.asarray(_tv(t)(Image.fromarray(frames[i])))
            assert got[i].shape == want.shape, (t, h, w)
            assert np.array_equal(got[i], want), (t, h, w, i, int(np.abs(got[i].astype(int) - want).max()))


def test_transform_then_model_matches_reference_stack():
    """Decoded frames -> GPU transform -> SSCD on tensor cores  vs  PIL frames -> the reference's Compose (Resize, ToTensor,
    Normalize) -> the fp32 PyTorch model: same tolerance as the model test (cosine >= 0.999, relative L2 <= 3e-2)."""
    import torch
    from PIL import Image
    from torchvision import transforms
    from vsc2022_b200 import inference_impl
    from vsc2022_b200.preprocess import InferenceTransforms as T, build_transforms
    from vsc2022_b200.sscd import SSCDResNet50, TorchReference
    ref = TorchReference(seed=2)
    model = SSCDResNet50(ref.trunk, ref.head)
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, size=(5, 180, 320, 3), dtype=np.uint8)
    compose = transforms.Compose([transforms.Resize(288), transforms.ToTensor(),
                                  transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    want = ref(torch.stack([compose(Image.fromarray(f)) for f in frames]).cuda()).float().cpu().numpy()
    feats = inference_impl.infer_videos([("v", np.arange(5) * 1.0, frames)], model, batch_size=4,
                                        transform=build_transforms(T.RESIZE_288))
    got = feats[0].feature
    assert got.shape == want.shape == (5, 512)
    cos = (got * want).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(want, axis=1))
    rel = np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)
    assert cos.min() >= 0.999 and rel.max() <= 3e-2, (cos.min(), rel.max())


def test_full_hd_batch_and_bad_input():
    import torch
    from vsc2022_b200.preprocess import InferenceTransforms as T, build_transforms
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    frames = torch.randint(0, 256, (16, 1080, 1920, 3), generator=g, device="cuda", dtype=torch.uint8)
    out = build_transforms(T.RESIZE_320_CENTER)(frames)
    assert out.shape == (16, 320, 320, 3) and out.is_cuda
    assert build_transforms(T.RESIZE_288)(frames[:0]).shape == (0, 288, 512, 3)
    with pytest.raises(ValueError):
        build_transforms(T.RESIZE_288)(torch.zeros((2, 8, 8, 4), dtype=torch.uint8))
