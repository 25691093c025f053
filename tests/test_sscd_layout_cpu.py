"""CPU check of the stem's weight-panel contract (vsc_conv_stem, include/vsc_b200.h): the 7x7 / stride-2 filter laid
out as a 4x4 / stride-1 filter over the 2x2 space-to-depth image, K index = ky2*64 + kx2*16 + (dy*2+dx)*3 + c.
The space-to-depth image and the 64-element windows are restated in torch here; the GPU kernels are tested against
the full model in tests/test_sscd_gpu.py."""
import torch
import torch.nn.functional as F

from vsc2022_b200.sscd import _Conv


def test_stem_panel_reproduces_conv_bn():
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
    bn = torch.nn.BatchNorm2d(64)
    bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.1)
    bn.eval()
    n, h, w = 2, 22, 30
    x = torch.randn(n, 3, h, w)
    with torch.no_grad():
        want = bn(conv(x))
    ho, wo = want.shape[2:]
    yd, xd = ho + 3, wo + 3

    stem = _Conv(torch, conv, bn, "cpu", stem_s2d=True)
    panel = stem.weight.float()                       # [64][256], bf16-rounded folded weights
    assert panel.shape == (64, 256)
    # padding channels and the taps of filter row / column 7 carry zero weights
    taps = panel.reshape(64, 4, 4, 16)
    assert torch.all(taps[..., 12:] == 0) and torch.all(taps[:, 3, :, 6:12] == 0) and torch.all(taps[:, :, 3, [3, 4, 5, 9, 10, 11]] == 0)

    # S[n][Y][X][(dy*2+dx)*3 + c] = x[n, c, 2Y+dy-3, 2X+dx-3] (zero outside, zero in channels 12..15), +4 cells of slack
    s2d = torch.zeros(n * yd * xd + 4, 16)
    xp = F.pad(x, (3, 2 * xd - w - 3, 3, 2 * yd - h - 3))             # padded so that (2Y+dy, 2X+dx) indexes it directly
    cells = xp.reshape(n, 3, yd, 2, xd, 2).permute(0, 2, 4, 3, 5, 1).reshape(n * yd * xd, 12)
    s2d[:n * yd * xd, :12] = cells
    flat = s2d.reshape(-1)
    # GEMM row m = (img*yd + oy)*xd + ox; k-block ky2 = the 64 contiguous elements starting at cell m + ky2*xd
    m = torch.arange(n * yd * xd)
    cols = torch.arange(64)[None, :]
    a = torch.cat([flat[((m[:, None] + ky2 * xd) * 16 + cols).clamp(max=flat.numel() - 1)] for ky2 in range(4)], dim=1)
    # (rows whose windows run past the image -- oy >= ho -- read clamped garbage; they are cut off below, like the
    # GPU kernel's padded output grid)
    out = (a @ panel.T + stem.bias).reshape(n, yd, xd, 64)[:, :ho, :wo].permute(0, 3, 1, 2)
    # only the bf16 rounding of the folded weights separates the two
    torch.testing.assert_close(out, want, rtol=0, atol=0.05)
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w_fold = (conv.weight * scale[:, None, None, None]).detach().to(torch.bfloat16).float()
    with torch.no_grad():
        exact = F.conv2d(x, w_fold, stride=2, padding=3) + (bn.bias - bn.running_mean * scale)[None, :, None, None]
    torch.testing.assert_close(out, exact, rtol=0, atol=2e-5)
