"""Freeze a stage-A fixture: descriptors of the plain PyTorch fp32 model (ResNet-50 trunk -> GeM(p=3) -> Linear 2048->512,
the architecture of adapt_sscd_model.py:56-70) computed ON THE CPU for seeded weights and seeded uint8 frames.

TEST INFRASTRUCTURE.  The reference pins nothing for this stage (the SSCD checkpoint is a download), so this fixture
pins the restatement itself: the GPU test compares the tcgen05 forward with these CPU float32 numbers, independently
of cuDNN on the GPU box.  Weights are not stored: `TorchReference(seed)` draws them on the CPU generator, identically
on every machine with this torch version (recorded in the fixture).

    python -m oracle.make_golden_sscd
"""
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    from vsc2022_b200.sscd import TorchReference, normalize_pixels
    seed = 5
    ref = TorchReference(seed=seed, device="cpu")
    rng = np.random.default_rng(seed)
    out = {"seed": np.int64(seed), "torch_version": np.array(torch.__version__)}
    for tag, (n, h, w) in {"a": (3, 64, 64), "b": (2, 75, 101)}.items():
        frames = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        with torch.no_grad():
            desc = ref(normalize_pixels(torch.from_numpy(frames))).numpy()
        out[f"frames_{tag}"], out[f"desc_{tag}"] = frames, desc.astype(np.float32)
    # fingerprint of the drawn weights, so a torch that initialises differently fails loudly instead of subtly
    out["weight_probe"] = np.array([float(ref.trunk.conv1.weight.detach().double().sum()), float(ref.head.weight.detach().double().sum()),
                                    float(ref.trunk.layer4[2].bn3.running_mean.double().sum())])
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, "sscd_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
