"""Per-frame optimisation loop on top of the msplat operator surface (BASELINE configs 3 and 4).

A compact restatement of the inner loop GFlow runs per frame
(/root/reference/gflow/trainer.py:383-558, driven by /root/reference/gflow/fit_video.py:288-315): raw
parameters -> activations -> render rgb + depth map from ONE sort -> photometric + scale/shift
invariant depth loss -> Adam with a LinearLR 1.0 -> 0.1 schedule.  It exists so the frame-sharded
mode of SURVEY.md 8e has something to shard and so BASELINE config 3 (300-iteration Adam loop) can
be timed on the GPU box, where /root/reference is not available; GFlow's own trainer runs unmodified
on the drop-in `msplat` module.  Deliberately NOT rebuilt here: SSIM, flow / still /
variance terms, densification, masks, logging, checkpoints (SURVEY.md 8f "next" rows).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import frames as _frames
from . import ops

ATTRS = ("xyz", "scale", "rotate", "opacity", "rgb")


def activate(name: str, x: torch.Tensor) -> torch.Tensor:
    """GFlow's activations (/root/reference/gflow/trainer.py:62-69)."""
    if name == "scale":
        return torch.abs(x)
    if name == "rotate":
        return torch.nn.functional.normalize(x)
    if name == "opacity":
        return torch.sigmoid(x * 10.0)
    if name == "rgb":
        return torch.sigmoid(x)
    return x


def inverse_activate(name: str, x: torch.Tensor) -> torch.Tensor:
    """Inverse activations (/root/reference/gflow/trainer.py:72-77)."""
    if name == "opacity":
        return torch.logit(x) / 10.0
    if name == "rgb":
        return torch.logit(x)
    if name == "rotate":
        return torch.nn.functional.normalize(x)
    if name == "scale":
        return torch.abs(x)
    return x


def pose_to_extr(pose: torch.Tensor) -> torch.Tensor:
    """(qx, qy, qz, qw, tx, ty, tz) -> world->camera [R|t] (3,4).

    Same convention as roma.RigidUnitQuat(Q, T).normalize().to_homogeneous()[:3] used by
    /root/reference/gflow/trainer.py:115-121 (xyzw quaternion, identity pose = (0,0,0,1,0,0,0)).
    """
    q = pose[:4] / pose[:4].norm()
    x, y, z, w = q[0], q[1], q[2], q[3]
    R = torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]),
    ])
    return torch.cat([R, pose[4:7].reshape(3, 1)], dim=1)


def extr_to_pose(extr: torch.Tensor) -> torch.Tensor:
    """Inverse of pose_to_extr for a proper rotation (Shepperd's method, xyzw order)."""
    R, t = extr[:, :3], extr[:, 3]
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    w = torch.sqrt(torch.clamp(1.0 + tr, min=1e-12)) / 2.0
    x = (R[2, 1] - R[1, 2]) / (4.0 * w)
    y = (R[0, 2] - R[2, 0]) / (4.0 * w)
    z = (R[1, 0] - R[0, 1]) / (4.0 * w)
    return torch.cat([torch.stack([x, y, z, w]), t])


@dataclass
class FitConfig:
    """Defaults follow /root/reference/scripts/fit_video.sh:16-41 (per-frame stage after frame 0)."""
    iterations: int = 300
    lr: float = 4e-3
    lr_camera: float = 0.0
    lambda_rgb: float = 1.0
    lambda_depth: float = 0.1
    camera_only: bool = False
    background: float = 0.0
    fused: bool = False  # True: render rgb through msplat.rasterization (single feature map only)


@dataclass
class FitResult:
    losses: List[float] = field(default_factory=list)
    image: Optional[torch.Tensor] = None   # (3,H,W) final render
    pose: Optional[torch.Tensor] = None    # (7,)
    uv: Optional[torch.Tensor] = None      # (N,2)


class FrameFitter:
    """Holds the raw Gaussian attributes + pose of one frame and optimises them against a target."""

    def __init__(self, state: Dict[str, torch.Tensor], intr: torch.Tensor, pose: torch.Tensor, W: int, H: int):
        dev = state["xyz"].device
        self.attrs = {k: torch.nn.Parameter(state[k].detach().clone().to(dev)) for k in ATTRS}
        self.intr = intr.detach().clone().to(dev)
        self.pose = torch.nn.Parameter(pose.detach().clone().to(dev))
        self.depth_a = torch.nn.Parameter(torch.ones(1, device=dev))
        self.depth_b = torch.nn.Parameter(torch.zeros(1, device=dev))
        self.W, self.H = int(W), int(H)

    def get_attribute(self, name: str) -> torch.Tensor:
        return activate(name, self.attrs[name])

    def get_extr(self) -> torch.Tensor:
        return pose_to_extr(self.pose)

    def state(self) -> Dict[str, torch.Tensor]:
        return {k: v.detach() for k, v in self.attrs.items()}

    def render(self, bg: float = 0.0, want_depth: bool = True):
        """rgb (3,H,W) and depth map (1,H,W) from one projection + one sort, the way
        /root/reference/gflow/utils/render.py:21-74 chains the operators."""
        xyz, scale, rotate = self.get_attribute("xyz"), self.get_attribute("scale"), self.get_attribute("rotate")
        opacity, rgb = self.get_attribute("opacity"), self.get_attribute("rgb")
        extr = self.get_extr()
        uv, depth = ops.project_point(xyz, self.intr, extr, self.W, self.H)
        visible = depth != 0
        cov3d = ops.compute_cov3d(scale, rotate, visible)
        conic, radius, tiles = ops.ewa_project(xyz, cov3d, self.intr, extr, uv, self.W, self.H, visible)
        ids, tile_range = ops.sort_gaussian(uv, depth, self.W, self.H, radius, tiles)
        img = ops.alpha_blending(uv, conic, opacity, rgb, ids, tile_range, bg, self.W, self.H)
        dmap = ops.alpha_blending(uv, conic, opacity, depth, ids, tile_range, bg, self.W, self.H) if want_depth else None
        return img, dmap, uv

    def train(self, gt_image: torch.Tensor, gt_depth: Optional[torch.Tensor], cfg: FitConfig) -> FitResult:
        """gt_image (H,W,3) in [0,1]; gt_depth (H,W,1) or None.  Returns per-iteration losses and the final render."""
        groups = [{"params": list(self.attrs.values()), "lr": cfg.lr},
                  {"params": [self.pose], "lr": cfg.lr_camera},
                  {"params": [self.depth_a, self.depth_b], "lr": cfg.lr}]
        opt = torch.optim.Adam(groups)
        sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.1, total_iters=cfg.iterations)
        use_depth = gt_depth is not None and cfg.lambda_depth > 0
        res = FitResult()
        loss_hist = []
        for _ in range(cfg.iterations):
            if cfg.fused and not use_depth:
                img = ops.rasterization(self.get_attribute("xyz"), self.get_attribute("scale"),
                                        self.get_attribute("rotate"), self.get_attribute("opacity"),
                                        self.get_attribute("rgb"), self.intr, self.get_extr(), self.W, self.H,
                                        cfg.background)
                dmap = None
            else:
                img, dmap, _ = self.render(cfg.background, want_depth=use_depth)
            loss = cfg.lambda_rgb * torch.mean((img.permute(1, 2, 0) - gt_image) ** 2)
            if use_depth:
                d = self.depth_a * dmap.permute(1, 2, 0) + self.depth_b
                # (a D + b - D_gt)^2 / (a D + b + D_gt), /root/reference/gflow/trainer.py:476-488; the clamp only
                # protects pixels where both depths are 0 (uncovered synthetic targets)
                loss = loss + cfg.lambda_depth * torch.mean((d - gt_depth) ** 2 / (d + gt_depth).clamp_min(1e-6))
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if cfg.camera_only:  # /root/reference/gflow/trainer.py:548-551
                for p in self.attrs.values():
                    if p.grad is not None:
                        p.grad.zero_()
            opt.step()
            sched.step()
            loss_hist.append(loss.detach())
        res.losses = [float(v) for v in torch.stack(loss_hist).cpu()] if loss_hist else []
        with torch.no_grad():
            res.image, _, res.uv = self.render(cfg.background, want_depth=False)
            res.pose = self.pose.detach().clone()
        return res


def fit_sequence_sharded(state0: Optional[Dict[str, torch.Tensor]], intr: torch.Tensor, frame_targets, W: int, H: int,
                         cfg: FitConfig, device: torch.device):
    """Frame-sharded sequence fit: every rank fits its own contiguous chunk of frames,
    each frame starting from the broadcast frame-0 state.

    `frame_targets(i)` returns (gt_image, gt_depth_or_None, pose0 (7,)) for frame i and `len(frame_targets)`
    is the number of frames; rank 0 passes `state0` (activated attributes are NOT expected: raw
    parameters as in the checkpoint), other ranks pass None.  Collectives: one broadcast before the
    loop, one all_gather after it.  Returns (local results keyed by frame index, gathered final frames
    on rank 0).
    """
    import torch.distributed as dist

    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    state = _frames.broadcast_state(state0, src=0, device=device) if world > 1 else state0
    mine = _frames.shard_frames(len(frame_targets), world, rank)
    results = {}
    last_img, last_pose = None, None
    for i in mine:
        gt_image, gt_depth, pose0 = frame_targets(i)
        fitter = FrameFitter(state, intr.to(device), pose0.to(device), W, H)
        results[i] = fitter.train(gt_image.to(device), None if gt_depth is None else gt_depth.to(device), cfg)
        last_img, last_pose = results[i].image, pose_to_extr(results[i].pose)
    gathered = None
    if world > 1:
        if last_img is None:  # a rank without frames still takes part in the collective
            last_img = torch.zeros(3, H, W, device=device)
            last_pose = torch.zeros(3, 4, device=device)
        gathered = _frames.gather_frames(last_img, last_pose, dst=0)
    return results, gathered
