"""Host logic of the GPU frame resize (vsc2022_b200/preprocess.py) against torchvision / PIL on the CPU:
output geometry of the three reference transforms, and the coefficient tables through the numpy restatement of
Pillow's resample (oracle/pil_resize.py) -- bit-exact against PIL's own Image.resize."""
import numpy as np
import pytest

GEOMETRIES = [(360, 640), (640, 360), (288, 288), (300, 288), (1080, 1920), (97, 131), (240, 426), (720, 406), (32, 500)]


def test_geometry_matches_torchvision():
    from PIL import Image
    from torchvision import transforms
    from vsc2022_b200.preprocess import InferenceTransforms as T, resized_geometry
    tv = {T.RESIZE_288: transforms.Resize(288),
          T.RESIZE_320_CENTER: transforms.Compose([transforms.Resize(320), transforms.CenterCrop(320)]),
          T.RESIZE_224_SQUARE: transforms.Resize((224, 224))}
    for h, w in GEOMETRIES + [(321, 480), (481, 320), (1000, 333)]:
        img = Image.fromarray(np.zeros((h, w, 3), np.uint8))
        for t, fn in tv.items():
            rh, rw, top, left, oh, ow = resized_geometry(t, h, w)
            got = fn(img)
            assert (got.height, got.width) == (oh, ow), (t, h, w)
            assert 0 <= top and top + oh <= rh and 0 <= left and left + ow <= rw


@pytest.mark.parametrize("h,w", GEOMETRIES)
def test_numpy_restatement_equals_pil(h, w):
    from PIL import Image
    from oracle import pil_resize
    from vsc2022_b200.preprocess import InferenceTransforms as T, resized_geometry
    rng = np.random.default_rng(h * 10007 + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 3] = (img[: h // 3] // 128) * 255          # hard edges: the clip to [0, 255] and the rounding both matter
    for t in T:
        rh, rw, top, left, oh, ow = resized_geometry(t, h, w)
        want = np.asarray(Image.fromarray(img).resize((rw, rh), Image.BILINEAR))
        got = pil_resize.resize_bilinear(img, rh, rw)
        assert np.array_equal(got, want), (t, h, w, int(np.abs(got.astype(int) - want).max()))


def test_enum_and_builder_mirror_the_reference_names():
    from vsc2022_b200.preprocess import InferenceTransforms, build_transforms
    assert [t.name for t in InferenceTransforms] == ["RESIZE_288", "RESIZE_320_CENTER", "RESIZE_224_SQUARE"]
    assert build_transforms("RESIZE_320_CENTER").transform is InferenceTransforms.RESIZE_320_CENTER
