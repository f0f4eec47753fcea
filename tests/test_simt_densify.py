"""Device-side densification (gflow_b200/csrc/densify.cu) through the SIMT shim against the numpy / torch
restatement of trainer.py:878-939 in oracle/fit_ref.py."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

from oracle import fit_ref as FR

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "simt"))
import emu  # noqa: E402

p = emu.p


def _prepare(err, mask, thr, W, H):
    L = emu.load()
    ws = torch.zeros(L.gfb_densify_workspace_bytes(W, H), dtype=torch.uint8)
    emu.ok(L.gfb_densify_prepare(p(err), p(mask), W, H, thr, p(ws), None), "prepare")
    stats = ws[:32].view(torch.int32)
    P = W * H
    off_w = 256
    weights = ws[off_w:off_w + 4 * P].view(torch.float32).reshape(H, W)
    return ws, int(stats[1]), float(ws[12:16].view(torch.float32)[0]), weights


@pytest.mark.parametrize("use_mask", [False, True])
def test_weights_and_mask_count(use_mask):
    W, H = 70, 45
    g = torch.Generator().manual_seed(0)
    err = torch.rand(H, W, generator=g) ** 4 * 0.01
    err[torch.rand(H, W, generator=g) < 0.2] = 0.0
    mask = (torch.rand(H, W, generator=g) > 0.6).to(torch.uint8) if use_mask else None
    ws, count, total, weights = _prepare(err, mask, 1e-3, W, H)
    w_o, m_o, ratio = FR.densify_weights(err.numpy(), 1e-3, None if mask is None else mask.numpy())
    assert count == int(m_o.sum())
    assert np.array_equal(weights.numpy(), w_o.astype(np.float32))
    assert abs(total - float(w_o.sum())) <= 1e-5 * float(w_o.sum())
    assert FR.densify_count(5000, ratio, 0.1) == int(5000 * (count / (W * H)) * 0.1)


def test_rgb_error_map():
    W, H = 33, 21
    g = torch.Generator().manual_seed(1)
    rendered, gt = torch.rand(3, H, W, generator=g), torch.rand(H, W, 3, generator=g)
    mask = (torch.rand(H, W, generator=g) > 0.3)
    L = emu.load()
    for m in (None, mask):
        out = torch.full((H, W), float("nan"))
        m8 = None if m is None else m.to(torch.uint8)  # keep the temporary alive across the call
        emu.ok(L.gfb_rgb_error_map(p(rendered), p(gt), p(m8), W, H, p(out), None), "map")
        r, t = (rendered, gt) if m is None else (rendered * m[None], gt * m[..., None])
        ref = ((r.permute(1, 2, 0) - t) ** 2).mean(dim=2)  # trainer.py:457
        assert torch.allclose(out, ref, rtol=1e-6, atol=1e-9)


def test_sampling_is_proportional_and_attributes_match_the_reference():
    W, H, num_points = 64, 40, 3000
    g = torch.Generator().manual_seed(2)
    err = torch.zeros(H, W)
    err[5:15, 10:30] = 0.004     # 200 px, weight ~ 4 units each
    err[25:35, 40:60] = 0.001    # 200 px, weight ~ 1 unit each -> about to be cut by the threshold (0.001 + minpos > 1e-3)
    err[20, 5] = 0.05            # one hot pixel
    gt_image = torch.rand(H, W, 3, generator=g)
    gt_image[6, 11] = torch.tensor([0.0, 1.0, 0.5])  # logit(0) / logit(1) as the reference computes them in float32
    gt_depth = 1.0 + 3.0 * torch.rand(H, W, 1, generator=g)
    intr = torch.tensor([40.0, 35.0, 31.5, 19.0])
    from gflow_b200.synthetic import make_camera

    _, extr = make_camera(W, H, g)
    ws, count_mask, total, weights = _prepare(err, None, 1e-3, W, H)
    w_o, m_o, _ = FR.densify_weights(err.numpy(), 1e-3, None)
    assert count_mask == int(m_o.sum()) == 401
    L = emu.load()
    n = 40000
    new = {k: torch.full((n, w), float("nan")) for k, w in (("xyz", 3), ("scale", 3), ("rotate", 4), ("opacity", 1), ("rgb", 3))}
    sampled = torch.full((n,), -7, dtype=torch.int32)
    emu.ok(L.gfb_densify_sample(p(ws), p(gt_image), p(gt_depth), p(intr), p(extr), W, H, n, num_points, 1234, p(new["xyz"]),
                                p(new["scale"]), p(new["rotate"]), p(new["opacity"]), p(new["rgb"]), p(sampled), None), "sample")
    s = sampled.long()
    assert int(s.min()) >= 0 and int(s.max()) < W * H
    wflat = torch.from_numpy(w_o.reshape(-1))
    assert bool((wflat[s] > 0).all()), "only pixels with a positive weight are ever drawn"
    # empirical frequencies against p = w / sum(w), region by region (4 sigma of the binomial)
    prob = wflat / wflat.sum()
    regions = {"hot": [20 * W + 5], "strong": [y * W + x for y in range(5, 15) for x in range(10, 30)],
               "weak": [y * W + x for y in range(25, 35) for x in range(40, 60)]}
    for name, pix in regions.items():
        pr = float(prob[pix].sum())
        got = float(torch.isin(s, torch.tensor(pix)).float().mean())
        assert abs(got - pr) <= 4 * (pr * (1 - pr) / n) ** 0.5 + 1e-4, (name, got, pr)
    # per-pixel uniformity inside a flat region: chi-square of 200 equally likely cells
    cnt = torch.bincount(s, minlength=W * H)[regions["strong"]].float()
    chi2 = float(((cnt - cnt.mean()) ** 2 / cnt.mean()).sum())
    assert chi2 < 300, chi2  # 199 degrees of freedom: mean 199, sd 20
    # a different seed gives a different draw, the same seed the same
    again = torch.zeros(n, dtype=torch.int32)
    emu.ok(L.gfb_densify_sample(p(ws), p(gt_image), p(gt_depth), p(intr), p(extr), W, H, n, num_points, 1234, p(new["xyz"]),
                                p(new["scale"]), p(new["rotate"]), p(new["opacity"]), p(new["rgb"]), p(again), None), "sample")
    assert torch.equal(again, sampled)
    # everything derived from the drawn pixels equals the reference's arithmetic on the same pixels
    ref = FR.densify_attributes(sampled, gt_image, gt_depth.squeeze(-1), intr, extr, num_points, W)
    assert torch.allclose(new["xyz"], ref["xyz"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(new["scale"], ref["scale"], rtol=1e-6)
    assert torch.equal(new["rotate"], ref["rotate"])
    assert torch.allclose(new["opacity"], ref["opacity"], rtol=1e-6)
    fin = torch.isfinite(ref["rgb"])
    assert torch.equal(torch.isfinite(new["rgb"]), fin) and torch.allclose(new["rgb"][fin], ref["rgb"][fin], rtol=1e-5, atol=1e-6)
    assert torch.equal(new["rgb"][~fin], ref["rgb"][~fin])  # +inf where the target is exactly 1 (reference behaviour)
    other = torch.zeros(n, dtype=torch.int32)
    emu.ok(L.gfb_densify_sample(p(ws), p(gt_image), p(gt_depth), p(intr), p(extr), W, H, n, num_points, 99, p(new["xyz"]),
                                p(new["scale"]), p(new["rotate"]), p(new["opacity"]), p(new["rgb"]), p(other), None), "sample")
    assert not torch.equal(other, sampled)


def test_degenerate_maps():
    W, H = 40, 30
    L = emu.load()
    # all-zero error and no mask: nothing qualifies
    ws, count, total, _ = _prepare(torch.zeros(H, W), None, 1e-3, W, H)
    assert count == 0 and total == 0.0
    # uniform error with an explicit mask (the occlusion densification of trainer.py:562-564)
    mask = torch.zeros(H, W, dtype=torch.uint8)
    mask[3:9, 4:20] = 1
    ws, count, total, weights = _prepare(torch.ones(H, W), mask, 0.0, W, H)
    assert count == 96 and abs(total - 2.0 * 96) < 1e-3  # ones + min positive (1)
    assert L.gfb_densify_sample(p(ws), None, None, None, None, W, H, 0, 100, 0, None, None, None, None, None, None, None) == 0
    assert L.gfb_densify_prepare(None, None, W, H, 0.0, p(ws), None) == -1
    assert L.gfb_densify_workspace_bytes(0, 5) == 0
