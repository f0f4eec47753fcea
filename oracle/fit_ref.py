"""CPU restatement (PyTorch autograd) of the per-frame optimisation iteration GFlow runs.

TEST INFRASTRUCTURE ONLY (same rules as oracle/splat_ref.py: only tests/, smoke() and the
cpu_baseline / --impl reference legs of bench.py may import it).  PARITY UNPINNED for the rasteriser
ops underneath (see oracle/splat_ref.py); everything restated HERE is plain PyTorch in the reference
and is followed line by line:

  activations            /root/reference/gflow/trainer.py:62-69
  pose -> extrinsics     /root/reference/gflow/trainer.py:115-121 (roma.RigidUnitQuat(q_xyzw, t).normalize()
                         .to_homogeneous()[:3]; roma is absent here, the quaternion -> matrix formula is
                         the standard one roma documents)
  render rgb + depth map /root/reference/gflow/utils/render.py:21-74 (one projection, one sort, two blends)
  loss_rgb               /root/reference/gflow/trainer.py:452-464: mean squared error over (H,W,3) plus
                         1 - SSIM, both multiplied by lambda_rgb
  SSIM                   /root/reference/gflow/utils/pytorch_ssim.py:7-37 (11x11 Gaussian window, sigma 1.5,
                         zero padding 5, C1 = 0.01^2, C2 = 0.03^2, mean over all elements)
  loss_depth             /root/reference/gflow/trainer.py:476-488: ((a D + b) - D_gt)^2 / ((a D + b) + D_gt), mean
  loss_var / loss_scale  /root/reference/gflow/trainer.py:490-503
  loss_still / loss_flow /root/reference/gflow/trainer.py:504-530
  gradient masks         /root/reference/gflow/trainer.py:535-551
  Adam + LinearLR        /root/reference/gflow/trainer.py:123-153,383-384,554-555 (torch.optim.Adam defaults,
                         LinearLR 1.0 -> 0.1 over `iterations`, stepped after the optimiser)

The loop is the one gflow_b200/fit.py drives through the CUDA operators and the one
gflow_b200/csrc/fit.cu runs natively; tests compare both with this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import splat_ref as R

ATTRS = ("xyz", "scale", "rotate", "opacity", "rgb")


def activate(name: str, x: torch.Tensor) -> torch.Tensor:
    if name == "scale":
        return torch.abs(x)
    if name == "rotate":
        return F.normalize(x)
    if name == "opacity":
        return torch.sigmoid(x * 10.0)
    if name == "rgb":
        return torch.sigmoid(x)
    return x


def pose_to_extr(pose: torch.Tensor) -> torch.Tensor:
    q = pose[:4] / pose[:4].norm()
    x, y, z, w = q[0], q[1], q[2], q[3]
    Rm = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]),
    ])
    return torch.cat([Rm, pose[4:7].reshape(3, 1)], dim=1)


def ssim_window(dtype=torch.float32) -> torch.Tensor:
    """pytorch_ssim.py:7-15: normalised 1-D Gaussian (11 taps, sigma 1.5), outer product."""
    g = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
    g = g / g.sum()
    return (g[:, None] @ g[None, :]).to(dtype)


def ssim(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """pytorch_ssim.py:17-37 for (1,C,H,W) inputs, size_average=True."""
    C = img1.shape[1]
    win = ssim_window(img1.dtype).expand(C, 1, 11, 11).contiguous()
    mu1 = F.conv2d(img1, win, padding=5, groups=C)
    mu2 = F.conv2d(img2, win, padding=5, groups=C)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=5, groups=C) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=5, groups=C) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=5, groups=C) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))
    return m.mean()


@dataclass
class FitRefConfig:
    iterations: int = 300
    lr: float = 4e-3
    lr_camera: float = 0.0
    lambda_rgb: float = 1.0
    use_ssim: bool = False          # GFlow: loss_rgb = mse + (1 - ssim); gflow_b200.fit default: mse only
    lambda_depth: float = 0.1
    lambda_var: float = 0.0
    lambda_scale: float = 0.0
    lambda_still: float = 0.0
    lambda_flow: float = 0.0
    camera_only: bool = False
    freeze_rgb: bool = False         # frames >= 1: rgb gradient zeroed (trainer.py:537-540)
    background: float = 0.0
    depth_den_min: float = 1e-6      # gflow_b200.fit clamps the depth-loss denominator (uncovered synthetic pixels)


def render(raw: Dict[str, torch.Tensor], pose, intr, W, H, bg, want_depth=True):
    xyz, scale, rot = raw["xyz"], activate("scale", raw["scale"]), activate("rotate", raw["rotate"])
    op, rgb = activate("opacity", raw["opacity"]), activate("rgb", raw["rgb"])
    extr = pose_to_extr(pose)
    uv, depth = R.project_point(xyz, intr, extr, W, H)
    vis = depth != 0
    cov = R.compute_cov3d(scale, rot, vis)
    conic, radius, tiles = R.ewa_project(xyz, cov, intr, extr, uv, W, H, vis)
    ids, rng = R.sort_gaussian(uv, depth, W, H, radius, tiles)
    img = R.alpha_blending(uv, conic, op, rgb, ids, rng, bg, W, H)
    dmap = R.alpha_blending(uv, conic, op, depth, ids, rng, bg, W, H) if want_depth else None
    return img, dmap, uv, depth


def flow_and_mask(last_uv, W, H, still_mask=None, camera_only=False):
    """trainer.py:511-517: Gaussians whose previous-frame centre lies inside the image, restricted to the
    still (camera-only stage) or moving (full stage) set."""
    m = (last_uv[:, 0] > 0) & (last_uv[:, 0] < W - 1) & (last_uv[:, 1] > 0) & (last_uv[:, 1] < H - 1)
    if still_mask is not None:
        n = still_mask.shape[0]
        m[:n] = (still_mask if camera_only else ~still_mask) & m[:n]
    return m.detach()


def iteration_loss(raw, pose, depth_ab, intr, gt_image, gt_depth, pixel_mask, W, H, cfg: FitRefConfig, prev=None,
                   still_mask=None):
    """One forward of the iteration; returns (loss, dict of parts).

    prev (optional, frames >= 1): dict(last_xyz (n,3), last_still_mask (n,), last_uv (n,2), gt_flow (H,W,2),
    and_mask (n,)) -- the state trainer.py:619-625 saves after the previous frame, and flow_and_mask()."""
    use_depth = gt_depth is not None and cfg.lambda_depth > 0
    img, dmap, uv, depth = render(raw, pose, intr, W, H, cfg.background, want_depth=use_depth)
    parts = {}
    gt = gt_image
    if pixel_mask is not None:  # trainer.py:453-455: both images are masked before the loss
        img = img * pixel_mask.to(img.dtype)[None]
        gt = gt * pixel_mask.to(gt.dtype)[..., None]
    mse = torch.mean((img.permute(1, 2, 0) - gt) ** 2)
    loss_rgb = mse
    parts["mse"] = mse.detach()
    if cfg.use_ssim:
        s = ssim(img[None], gt.permute(2, 0, 1)[None])
        loss_rgb = loss_rgb + (1 - s)
        parts["ssim"] = s.detach()
    loss = cfg.lambda_rgb * loss_rgb
    if use_depth:
        d = depth_ab[0] * dmap.permute(1, 2, 0) + depth_ab[1]
        den = d + gt_depth
        ld = (d - gt_depth) ** 2 / (den.clamp_min(cfg.depth_den_min) if cfg.depth_den_min > 0 else den)
        if pixel_mask is not None:
            ld = ld * pixel_mask.to(ld.dtype)[..., None]
        ld = ld.mean()
        parts["depth"] = ld.detach()
        loss = loss + cfg.lambda_depth * ld
    if cfg.lambda_var:
        lv = torch.mean(torch.std(activate("scale", raw["scale"]), dim=1))
        parts["var"] = lv.detach()
        loss = loss + cfg.lambda_var * lv
    if cfg.lambda_scale:
        within = (uv[:, 0] > 0) & (uv[:, 0] < W - 1) & (uv[:, 1] > 0) & (uv[:, 1] < H - 1)
        if still_mask is not None:  # trainer.py:467-471 narrows the index in place; 495-501 reads it via self.within_index
            n = still_mask.shape[0]
            within[:n] = (still_mask if cfg.camera_only else ~still_mask) & within[:n]
        ls = torch.norm(activate("scale", raw["scale"])[within], dim=1) * (1.0 / depth[within]).squeeze(-1)
        ls = ls.mean()
        parts["scale"] = ls.detach()
        loss = loss + cfg.lambda_scale * ls
    if cfg.lambda_still and prev is not None and "last_still_mask" in prev:  # trainer.py:504-508
        m = prev["last_still_mask"]
        n = m.shape[0]
        lst = torch.norm(raw["xyz"][:n][m] - prev["last_xyz"][:n][m], dim=1).mean()
        parts["still"] = lst.detach()
        loss = loss + cfg.lambda_still * lst
    if cfg.lambda_flow and prev is not None and "gt_flow" in prev:  # trainer.py:510-530
        am, last_uv = prev["and_mask"], prev["last_uv"]
        pred = uv[: last_uv.shape[0]][am] - last_uv[am]
        gtf = prev["gt_flow"][last_uv[am][:, 1].long(), last_uv[am][:, 0].long()]
        lfl = F.mse_loss(pred, gtf)
        parts["flow"] = lfl.detach()
        loss = loss + cfg.lambda_flow * lfl
    parts["total"] = loss.detach()
    return loss, parts


def moving_footprint(raw, pose, intr, W, H, bg, tentative_still):
    """trainer.py:427-451: the Gaussians that are NOT tentatively still, rendered (detached) under the current
    pose; a pixel belongs to the moving footprint when its grey value is > 0."""
    n = tentative_still.shape[0]
    sub = {k: raw[k].detach()[:n][~tentative_still] for k in ATTRS}
    with torch.no_grad():
        img, _, _, _ = render(sub, pose.detach(), intr, W, H, bg, want_depth=False)
    return (0.299 * img[0] + 0.587 * img[1] + 0.114 * img[2]) > 0.0


def fit_loop(raw0: Dict[str, torch.Tensor], pose0, intr, gt_image, gt_depth, W, H, cfg: FitRefConfig,
             pixel_mask: Optional[torch.Tensor] = None, still_mask: Optional[torch.Tensor] = None, n_iters=None,
             record=None, prev=None, tentative_still: Optional[torch.Tensor] = None):
    """Runs `n_iters` (default cfg.iterations) iterations; returns (raw, pose, depth_ab, history).

    history[i] = dict(parts of iteration i, grads = raw gradients BEFORE masking, params after the step).
    `record(i, dict)` may be passed to stream the same information instead of keeping it.
    """
    raw = {k: raw0[k].detach().clone().requires_grad_(True) for k in ATTRS}
    pose = pose0.detach().clone().requires_grad_(True)
    depth_ab = torch.tensor([1.0, 0.0], dtype=pose.dtype).requires_grad_(True)
    opt = torch.optim.Adam([{"params": list(raw.values()), "lr": cfg.lr}, {"params": [pose], "lr": cfg.lr_camera},
                            {"params": [depth_ab], "lr": cfg.lr}])
    sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.1, total_iters=cfg.iterations)
    history = []
    for it in range(cfg.iterations if n_iters is None else n_iters):
        if cfg.camera_only and tentative_still is not None:  # move_mask = move_gs_mask | move_mask, cumulative
            keep = torch.ones(H, W, dtype=torch.bool) if pixel_mask is None else pixel_mask.bool()
            pixel_mask = keep & ~moving_footprint(raw, pose, intr, W, H, cfg.background, tentative_still)
        loss, parts = iteration_loss(raw, pose, depth_ab, intr, gt_image, gt_depth, pixel_mask, W, H, cfg, prev, still_mask)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        grads = {k: (v.grad.detach().clone() if v.grad is not None else torch.zeros_like(v)) for k, v in raw.items()}
        grads["pose"] = pose.grad.detach().clone() if pose.grad is not None else torch.zeros_like(pose)
        grads["depth_ab"] = depth_ab.grad.detach().clone() if depth_ab.grad is not None else None
        if cfg.freeze_rgb and raw["rgb"].grad is not None:
            raw["rgb"].grad.zero_()
        if still_mask is not None and raw["xyz"].grad is not None:
            raw["xyz"].grad[: still_mask.shape[0]][still_mask] = 0.0
        if cfg.camera_only:
            for p in raw.values():
                if p.grad is not None:
                    p.grad.zero_()
        opt.step()
        sched.step()
        rec = dict(parts, grads=grads, pixel_mask=None if pixel_mask is None else pixel_mask.clone(), params={k: v.detach().clone() for k, v in raw.items()},
                   pose=pose.detach().clone(), depth_ab=depth_ab.detach().clone())
        if record is not None:
            record(it, rec)
        else:
            history.append(rec)
    return {k: v.detach() for k, v in raw.items()}, pose.detach(), depth_ab.detach(), history


# --------------------------------------------------------------------------- densification
def densify_weights(error_map, error_threshold=1e-3, mask=None):
    """trainer.py:880-895 on numpy: returns (weights (H,W), mask (H,W) bool, mask_ratio)."""
    import numpy as np

    e = np.asarray(error_map, dtype=np.float32)
    pos = e[e > 0]
    e = e + (np.nanmin(pos) if pos.size else np.float32(0))
    if mask is None:
        m = (e > error_threshold).squeeze()
    else:
        m = np.asarray(mask).squeeze()
    m = m > 0
    w = e * m
    return w, m, float(np.sum(m)) / m.size


def densify_count(num_points, mask_ratio, percent):
    """trainer.py:900."""
    return int(num_points * mask_ratio * percent)


def densify_attributes(sampled_pixels, gt_image, gt_depth, intr, extr, num_points, W):
    """trainer.py:904-933 for an already drawn set of flat pixel indices: new raw attributes."""
    idx = torch.as_tensor(sampled_pixels, dtype=torch.long)
    ys, xs = idx // W, idx % W
    xys = torch.stack([xs, ys], dim=1).float()
    depths = gt_depth[ys, xs].reshape(-1, 1).float()
    scales = torch.ones(idx.numel()) * (1.0 / num_points) * (depths / depths.min()).squeeze(-1)
    rgbs = gt_image[ys, xs]
    # geometry.py:104-116
    rel = torch.cat((depths * (xys - intr[2:]) / intr[0], depths), dim=-1)
    extr_h = torch.cat((extr, torch.tensor([[0.0, 0.0, 0.0, 1.0]])), dim=0)
    c2w = torch.linalg.inv(extr_h)
    xyz = rel @ c2w[:3, :3].T + c2w[:3, 3]
    new_scale = torch.abs(scales.unsqueeze(1).repeat(1, 3))
    rgbs = torch.clamp(rgbs.contiguous(), min=1e-15, max=1 - 1e-15)
    new_rgb = torch.logit(rgbs)
    new_rot = torch.tensor([1.0, 0.0, 0.0, 0.0]).repeat(idx.numel(), 1)
    new_op = torch.logit(0.99 * torch.ones(idx.numel(), 1)) / 10.0
    return dict(xyz=xyz, scale=new_scale, rotate=new_rot, opacity=new_op, rgb=new_rgb)
