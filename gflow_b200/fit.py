"""Per-frame optimisation loop on top of the msplat operator surface (BASELINE configs 3 and 4).

A compact restatement of the inner loop GFlow runs per frame
(/root/reference/gflow/trainer.py:383-558, driven by /root/reference/gflow/fit_video.py:288-315): raw
parameters -> activations -> render rgb + depth map from ONE sort -> photometric + scale/shift
invariant depth loss -> Adam with a LinearLR 1.0 -> 0.1 schedule.  It exists so the frame-sharded
mode of SURVEY.md 8e has something to shard and so BASELINE config 3 (300-iteration Adam loop) can
be timed on the GPU box, where /root/reference is not available; GFlow's own trainer runs unmodified
on the drop-in `msplat` module.

Two execution modes with the same semantics:
  * operator path (default): the msplat operators one by one + torch autograd + torch.optim.Adam, the way
    trainer.py drives them (~150 launches and 2-3 ms of host work per iteration);
  * native path (FitConfig.native): the whole iteration as seven kernels with no host synchronisation
    (csrc/fit.cu through gfb_fit_init / gfb_fit_iterate): activations + geometry + binning, one 4-channel
    blend for rgb and the depth map, fused losses, blend backward, geometry backward with the Adam update
    in the same kernel.
Terms covered (both modes): mse [+ 1 - SSIM], depth, scale-variance and scale/depth regularisers, the still-position
and flow-consistency terms, the camera-only / frozen-rgb / still-xyz gradient masks, the pixel mask and the moving
subset's footprint (trainer.py:427-551), error-driven and occlusion densification (trainer.py:560-571, 878-951).
Frame-state files and the wire format live in gflow_b200/checkpoint.py.  Not rebuilt: the concave-hull
segmentation, logging and video dumps after a stage.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import ctypes
import math

import torch

from . import capi
from . import densify as _densify
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
    """Inverse of pose_to_extr for a proper rotation: Shepperd's method with all four branches (xyzw order), valid for
    every rotation including the 180-degree ones (axis flips such as an OpenGL <-> OpenCV extrinsic), like the
    roma.rotmat_to_unitquat the reference uses (/root/reference/gflow/trainer.py:177).  The branch with the largest
    pivot is taken, so no component is ever divided by a value near zero.  Returned with w >= 0."""
    R, t = extr[:, :3], extr[:, 3]
    r = [[float(R[i, j]) for j in range(3)] for i in range(3)]
    tr = r[0][0] + r[1][1] + r[2][2]
    pivots = [tr, r[0][0], r[1][1], r[2][2]]
    k = max(range(4), key=lambda i: pivots[i])
    if k == 0:
        s4 = 2.0 * (1.0 + tr) ** 0.5            # 4 w
        w, x, y, z = 0.25 * s4, (r[2][1] - r[1][2]) / s4, (r[0][2] - r[2][0]) / s4, (r[1][0] - r[0][1]) / s4
    elif k == 1:
        s4 = 2.0 * (1.0 + r[0][0] - r[1][1] - r[2][2]) ** 0.5  # 4 x
        w, x, y, z = (r[2][1] - r[1][2]) / s4, 0.25 * s4, (r[0][1] + r[1][0]) / s4, (r[0][2] + r[2][0]) / s4
    elif k == 2:
        s4 = 2.0 * (1.0 + r[1][1] - r[0][0] - r[2][2]) ** 0.5  # 4 y
        w, x, y, z = (r[0][2] - r[2][0]) / s4, (r[0][1] + r[1][0]) / s4, 0.25 * s4, (r[1][2] + r[2][1]) / s4
    else:
        s4 = 2.0 * (1.0 + r[2][2] - r[0][0] - r[1][1]) ** 0.5  # 4 z
        w, x, y, z = (r[1][0] - r[0][1]) / s4, (r[0][2] + r[2][0]) / s4, (r[1][2] + r[2][1]) / s4, 0.25 * s4
    if w < 0.0:
        w, x, y, z = -w, -x, -y, -z
    q = torch.tensor([x, y, z, w], dtype=extr.dtype, device=extr.device)
    return torch.cat([q, t])


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
    native: bool = False  # True: the whole iteration runs natively (csrc/fit.cu), no autograd / torch.optim
    use_ssim: bool = False  # loss_rgb = mse + (1 - SSIM) as in trainer.py:459-462 (False: mse only)
    lambda_var: float = 0.0  # trainer.py:490-492
    lambda_scale: float = 0.0  # trainer.py:495-501
    lambda_still: float = 0.0  # trainer.py:504-508 (needs `prev`)
    lambda_flow: float = 0.0  # trainer.py:510-530 (needs `prev`)
    freeze_rgb: bool = False  # frames >= 1: rgb gradient zeroed (trainer.py:537-540)
    check_every: int = 50  # native path: iterations enqueued between two looks at the intersection count
    depth_den_min: float = 1e-6  # lower clamp of the depth-loss denominator (0 = the reference's unclamped quotient)
    # error-driven densification, trainer.py:566-571 ((iteration + 1) % interval == 0, at most `times` times)
    densify_interval: int = 0
    densify_times: int = 0
    densify_err_thre: float = 1e-3
    densify_err_percent: float = 0.1
    densify_occ_percent: float = 0.1  # trainer.py:562-564: after iteration 0 of a later frame, inside the occlusion mask
    num_points: Optional[int] = None  # the reference's self.num_points (configured count); default: initial N
    densify_seed: int = 0


_SSIM_WINDOW = None


def ssim(img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """SSIM of two (1,C,H,W) images exactly as /root/reference/gflow/utils/pytorch_ssim.py:7-37 computes it
    (11x11 Gaussian window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2, mean over all elements)."""
    global _SSIM_WINDOW
    C = img1.shape[1]
    if _SSIM_WINDOW is None:
        g = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], dtype=torch.float32)
        g = g / g.sum()
        _SSIM_WINDOW = g[:, None] @ g[None, :]
    win = _SSIM_WINDOW.to(img1.device, img1.dtype).expand(C, 1, 11, 11).contiguous()
    conv = lambda x: torch.nn.functional.conv2d(x, win, padding=5, groups=C)  # noqa: E731
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1, s2, s12 = conv(img1 * img1) - mu1_sq, conv(img2 * img2) - mu2_sq, conv(img1 * img2) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


@dataclass
class PrevFrame:
    """What GFlow carries from the previous frame into the still / flow terms (trainer.py:619-625) plus this
    frame's flow prior: last_xyz (n,3), last_still_mask (n,) bool, last_uv (n,2), gt_flow (H,W,2)."""
    last_xyz: Optional[torch.Tensor] = None
    last_still_mask: Optional[torch.Tensor] = None
    last_uv: Optional[torch.Tensor] = None
    gt_flow: Optional[torch.Tensor] = None

    def flow_and_mask(self, W: int, H: int, still_mask: Optional[torch.Tensor], camera_only: bool) -> torch.Tensor:
        """trainer.py:511-517."""
        uv = self.last_uv
        m = (uv[:, 0] > 0) & (uv[:, 0] < W - 1) & (uv[:, 1] > 0) & (uv[:, 1] < H - 1)
        if still_mask is not None:
            n = still_mask.shape[0]
            m[:n] = (still_mask if camera_only else ~still_mask) & m[:n]
        return m

    def flow_target(self, and_mask: torch.Tensor) -> torch.Tensor:
        """(n,2): where each selected Gaussian's centre should land, last_uv + gt_flow[last_uv] (rows outside
        and_mask are unused)."""
        uv = self.last_uv
        x = uv[:, 0].long().clamp(0, self.gt_flow.shape[1] - 1)
        y = uv[:, 1].long().clamp(0, self.gt_flow.shape[0] - 1)
        return (uv + self.gt_flow[y, x]).contiguous()


def warp_moving_by_flow(xyz: torch.Tensor, prev: "PrevFrame", gt_depth: torch.Tensor, intr: torch.Tensor,
                        extr: torch.Tensor, W: int, H: int) -> torch.Tensor:
    """The pre-update step of a full (not camera-only) stage on frames >= 1, trainer.py:348-376: every MOVING
    Gaussian of the previous frame whose centre was inside the image is moved to where the flow prior sends its
    centre, back-projected with the new depth prior (geometry.py:104-116, focal = fx on both axes).  Returns the
    new raw xyz; runs once per frame, plain tensor ops on whatever device the inputs live on."""
    m = prev.last_still_mask
    n = m.shape[0]
    uv_move = prev.last_uv[:n][~m]
    inside = (uv_move[:, 0] > 0) & (uv_move[:, 0] < W - 1) & (uv_move[:, 1] > 0) & (uv_move[:, 1] < H - 1)
    uv_in = uv_move[inside]
    uv_new = uv_in + prev.gt_flow[uv_in[:, 1].long(), uv_in[:, 0].long()]
    yc = uv_new[:, 1].long().clamp(0, H - 1)
    xc = uv_new[:, 0].long().clamp(0, W - 1)
    depth = gt_depth[yc, xc].reshape(-1, 1)
    cam = torch.cat((depth * (uv_new - intr[2:]) / intr[0], depth), dim=-1)
    R, t = extr[:, :3], extr[:, 3]
    world = (cam - t) @ R  # R^T (p_cam - t) for the rigid world->camera [R|t]
    out = xyz.detach().clone()
    idx = torch.nonzero(~m, as_tuple=False).squeeze(-1)[inside]
    out[idx] = world.to(out.dtype)
    return out


@dataclass
class FitResult:
    losses: List[float] = field(default_factory=list)
    image: Optional[torch.Tensor] = None   # (3,H,W) final render
    pose: Optional[torch.Tensor] = None    # (7,)
    uv: Optional[torch.Tensor] = None      # (N,2) of the final parameters
    last_uv: Optional[torch.Tensor] = None     # (N,2) of the LAST iteration's forward (what trainer.py:619-622 keeps)
    last_depth: Optional[torch.Tensor] = None  # (N,1) likewise


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

    def render(self, bg: float = 0.0, want_depth: bool = True, with_depth: bool = False):
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
        if with_depth:
            return img, dmap, uv, depth
        return img, dmap, uv

    def train(self, gt_image: torch.Tensor, gt_depth: Optional[torch.Tensor], cfg: FitConfig,
              pixel_mask: Optional[torch.Tensor] = None, still_mask: Optional[torch.Tensor] = None,
              prev: Optional[PrevFrame] = None, tentative_still: Optional[torch.Tensor] = None,
              occlusion_mask: Optional[torch.Tensor] = None) -> FitResult:
        """gt_image (H,W,3) in [0,1]; gt_depth (H,W,1) or None; pixel_mask (H,W) bool (True = pixel counts,
        trainer.py:452-455); still_mask (n,) bool (True = xyz frozen, trainer.py:542-546); prev: previous-frame
        state for the still / flow terms; tentative_still (n,) bool: in a camera-only stage the Gaussians that are
        NOT tentatively still are re-rendered every iteration and their footprint leaves the losses, cumulatively
        (trainer.py:427-451); occlusion_mask (H,W): on frames >= 1 (prev given), new Gaussians are drawn uniformly
        inside it right after iteration 0 (trainer.py:562-564).  Returns per-iteration losses and the final render."""
        occ = None
        if occlusion_mask is not None and prev is not None and not cfg.camera_only and bool((occlusion_mask > 0).any()):
            om = occlusion_mask > 0
            if om.dim() == 3:  # (H,W,1) from the reference's image reader, or an RGB mask image
                om = om.any(dim=-1)
            occ = om.reshape(self.H, self.W)
        if cfg.native:
            return self._train_native(gt_image, gt_depth, cfg, pixel_mask, still_mask, prev, tentative_still, occ)
        dynamic = cfg.camera_only and tentative_still is not None
        if dynamic:
            keep = torch.ones(self.H, self.W, dtype=torch.bool, device=gt_image.device) if pixel_mask is None else pixel_mask.bool()
        use_still = bool(cfg.lambda_still) and prev is not None and prev.last_still_mask is not None
        use_flow = bool(cfg.lambda_flow) and prev is not None and prev.gt_flow is not None
        if use_flow:
            and_mask = prev.flow_and_mask(self.W, self.H, still_mask, cfg.camera_only)
            flow_target = prev.flow_target(and_mask)
        groups = [{"params": list(self.attrs.values()), "lr": cfg.lr},
                  {"params": [self.pose], "lr": cfg.lr_camera},
                  {"params": [self.depth_a, self.depth_b], "lr": cfg.lr}]
        opt = torch.optim.Adam(groups)
        sched = torch.optim.lr_scheduler.LinearLR(opt, start_factor=1.0, end_factor=0.1, total_iters=cfg.iterations)
        use_depth = gt_depth is not None and cfg.lambda_depth > 0
        res = FitResult()
        loss_hist = []
        pm = None if pixel_mask is None else pixel_mask.to(gt_image.dtype)
        gt = gt_image if pm is None else gt_image * pm[..., None]
        self._num_points0 = cfg.num_points or int(self.attrs["xyz"].shape[0])
        for it in range(cfg.iterations):
            uv = depth = None
            if dynamic:
                keep = keep & ~self._moving_footprint(tentative_still, cfg.background)
                pm = keep.to(gt_image.dtype)
                gt = gt_image * pm[..., None]
            if cfg.fused and not use_depth and not cfg.lambda_scale and not use_flow:
                img = ops.rasterization(self.get_attribute("xyz"), self.get_attribute("scale"),
                                        self.get_attribute("rotate"), self.get_attribute("opacity"),
                                        self.get_attribute("rgb"), self.intr, self.get_extr(), self.W, self.H,
                                        cfg.background)
                dmap = None
            else:
                img, dmap, uv, depth = self.render(cfg.background, want_depth=use_depth, with_depth=True)
            if pm is not None:
                img = img * pm[None]
            loss_rgb = torch.mean((img.permute(1, 2, 0) - gt) ** 2)
            if cfg.use_ssim:
                loss_rgb = loss_rgb + (1.0 - ssim(img[None], gt.permute(2, 0, 1)[None]))
            loss = cfg.lambda_rgb * loss_rgb
            if use_depth:
                d = self.depth_a * dmap.permute(1, 2, 0) + self.depth_b
                # (a D + b - D_gt)^2 / (a D + b + D_gt), /root/reference/gflow/trainer.py:476-488; the clamp only
                # protects pixels where both depths are 0 (uncovered synthetic targets)
                den = d + gt_depth
                ld = (d - gt_depth) ** 2 / (den.clamp_min(cfg.depth_den_min) if cfg.depth_den_min > 0 else den)
                if pm is not None:
                    ld = ld * pm[..., None]
                loss = loss + cfg.lambda_depth * torch.mean(ld)
            if cfg.lambda_var:
                loss = loss + cfg.lambda_var * torch.mean(torch.std(self.get_attribute("scale"), dim=1))
            if cfg.lambda_scale:
                within = (uv[:, 0] > 0) & (uv[:, 0] < self.W - 1) & (uv[:, 1] > 0) & (uv[:, 1] < self.H - 1)
                if still_mask is not None:  # trainer.py:467-471,495-501 (in-place narrowing seen through an alias)
                    n = still_mask.shape[0]
                    within[:n] = (still_mask if cfg.camera_only else ~still_mask) & within[:n]
                ls = torch.norm(self.get_attribute("scale")[within], dim=1) * (1.0 / depth[within]).squeeze(-1)
                loss = loss + cfg.lambda_scale * ls.mean()
            if use_still:
                m = prev.last_still_mask
                n = m.shape[0]
                loss = loss + cfg.lambda_still * torch.norm(self.get_attribute("xyz")[:n][m] - prev.last_xyz[:n][m],
                                                            dim=1).mean()
            if use_flow:
                n = flow_target.shape[0]
                loss = loss + cfg.lambda_flow * torch.mean((uv[:n][and_mask] - flow_target[and_mask]) ** 2)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if cfg.freeze_rgb and self.attrs["rgb"].grad is not None:  # trainer.py:537-540
                self.attrs["rgb"].grad.zero_()
            if still_mask is not None and self.attrs["xyz"].grad is not None:  # trainer.py:542-546
                self.attrs["xyz"].grad[: still_mask.shape[0]][still_mask] = 0.0
            if cfg.camera_only:  # /root/reference/gflow/trainer.py:548-551
                for p in self.attrs.values():
                    if p.grad is not None:
                        p.grad.zero_()
            opt.step()
            sched.step()
            loss_hist.append(loss.detach())
            if it == 0 and occ is not None:
                if self._densify_operator_path(img.detach(), gt, gt_image, gt_depth, cfg, it, occ):
                    opt = torch.optim.Adam(list(self.attrs.values()), lr=cfg.lr)
            if self._densify_due(cfg, it):
                # trainer.py:566-571 + 941-951: append Gaussians drawn from the error map, then a NEW Adam over
                # the attributes only (initial lr, no scheduler; pose / depth_a / depth_b drop out)
                if self._densify_operator_path(img.detach(), gt, gt_image, gt_depth, cfg, it):
                    opt = torch.optim.Adam(list(self.attrs.values()), lr=cfg.lr)
        res.losses = [float(v) for v in torch.stack(loss_hist).cpu()] if loss_hist else []
        if cfg.iterations > 0 and uv is not None:
            res.last_uv, res.last_depth = uv.detach(), depth.detach()
        with torch.no_grad():
            res.image, _, res.uv = self.render(cfg.background, want_depth=False)
            res.pose = self.pose.detach().clone()
        return res

    def _moving_footprint(self, tentative_still: torch.Tensor, bg: float) -> torch.Tensor:
        """(H,W) bool: pixels the not-tentatively-still Gaussians touch under the current pose (trainer.py:427-451)."""
        n = tentative_still.shape[0]
        sel = ~tentative_still
        with torch.no_grad():
            a = {k: self.get_attribute(k).detach()[:n][sel] for k in ATTRS}
            img = ops.rasterization(a["xyz"], a["scale"], a["rotate"], a["opacity"], a["rgb"], self.intr, self.get_extr().detach(),
                                    self.W, self.H, bg)
        return (0.299 * img[0] + 0.587 * img[1] + 0.114 * img[2]) > 0.0

    @staticmethod
    def _densify_due(cfg: FitConfig, it: int) -> bool:
        k = cfg.densify_interval
        return bool(k) and not cfg.camera_only and (it + 1) % k == 0 and (it + 1) // k <= cfg.densify_times

    def _densify_operator_path(self, img, gt_masked, gt_image, gt_depth, cfg: FitConfig, it: int, occ=None) -> bool:
        if gt_depth is None:
            raise RuntimeError("gflow_b200: densification back-projects with the depth prior; gt_depth is required")
        dens = _densify.Densifier(self.W, self.H, self.attrs["xyz"].device)
        if occ is not None:  # uniform draw inside the occlusion mask
            err, thre, pct = torch.ones(self.H, self.W, device=img.device), 0.0, cfg.densify_occ_percent
        else:
            err, thre, pct = dens.rgb_error_map(img.contiguous(), gt_masked.contiguous()), cfg.densify_err_thre, cfg.densify_err_percent
        new = dens.sample(err, gt_image, gt_depth, self.intr, self.get_extr().detach(), self._num_points0, thre, pct, mask=occ,
                          seed=cfg.densify_seed + it)
        if new is None:
            return False
        for k in ATTRS:
            self.attrs[k] = torch.nn.Parameter(torch.cat([self.attrs[k].data, new[k]], dim=0))
        return True

    def _train_native(self, gt_image, gt_depth, cfg: FitConfig, pixel_mask, still_mask, prev=None,
                      tentative_still=None, occ=None) -> FitResult:
        if cfg.iterations <= 0:  # nothing to optimise: the reference's loop body simply does not run
            res = FitResult()
            with torch.no_grad():
                res.image, _, res.uv = self.render(cfg.background, want_depth=False)
                res.pose = self.pose.detach().clone()
            return res
        loop = NativeFitLoop(self, gt_image, gt_depth, cfg, pixel_mask=pixel_mask, still_mask=still_mask, prev=prev,
                             tentative_still=tentative_still)
        if occ is not None and cfg.iterations > 0:
            loop.run(1)
            loop.densify(0.0, cfg.densify_occ_percent, mask=occ, uniform_error=True, seed=cfg.densify_seed)
        if cfg.densify_interval and not cfg.camera_only:
            k, done = cfg.densify_interval, loop.done
            while done < cfg.iterations:
                nxt = min(cfg.iterations, (done // k + 1) * k)
                loop.run(nxt - done)
                done = nxt
                if self._densify_due(cfg, done - 1):
                    loop.densify(cfg.densify_err_thre, cfg.densify_err_percent, seed=cfg.densify_seed + done - 1)
        else:
            loop.run(cfg.iterations - loop.done)
        res = FitResult()
        res.losses = [float(v) for v in loop.loss_history()[:, 0].cpu()]
        if loop.done > 0:
            res.last_uv, res.last_depth = loop.last_uv(), loop.last_depth()
        with torch.no_grad():
            res.image, _, res.uv = self.render(cfg.background, want_depth=False)
            res.pose = self.pose.detach().clone()
        return res


class NativeFitLoop:
    """Host driver of the native iteration (csrc/fit.cu): owns the workspace, enqueues iterations in chunks
    and only looks at the device between chunks (intersection count against the workspace capacity).

    The raw attributes / pose / depth_a,b of `fitter` are updated IN PLACE by the kernels.  There is no
    CPU fallback: everything must live on one CUDA device."""

    def __init__(self, fitter: "FrameFitter", gt_image: torch.Tensor, gt_depth: Optional[torch.Tensor], cfg: FitConfig,
                 pixel_mask: Optional[torch.Tensor] = None, still_mask: Optional[torch.Tensor] = None,
                 capacity: Optional[int] = None, debug: bool = False, prev: Optional[PrevFrame] = None,
                 tentative_still: Optional[torch.Tensor] = None, sub_capacity: Optional[int] = None, stream=None):
        dev = fitter.attrs["xyz"].device
        self._require_device(dev)
        self.lib = self._library()
        self.stream = stream  # torch.cuda.Stream this loop enqueues on (None: the current stream at each call)
        self._pending = None
        self.fitter, self.cfg, self.dev = fitter, cfg, dev
        self.N, self.W, self.H = int(fitter.attrs["xyz"].shape[0]), fitter.W, fitter.H
        if self.N <= 0:
            raise RuntimeError("gflow_b200: the native fit loop needs at least one Gaussian")
        for k, w in _frames.STATE_KEYS:
            t = fitter.attrs[k].data
            if tuple(t.shape) != (self.N, w) or t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError(f"gflow_b200: attribute {k} must be a contiguous float32 ({self.N}, {w}) tensor")
        f32 = dict(device=dev, dtype=torch.float32)
        self.gt_image = gt_image.detach().to(**f32).contiguous()
        if tuple(self.gt_image.shape) != (self.H, self.W, 3):
            raise RuntimeError(f"gflow_b200: gt_image must be ({self.H}, {self.W}, 3), got {tuple(gt_image.shape)}")
        self.use_depth = gt_depth is not None and cfg.lambda_depth > 0
        self.gt_depth = gt_depth.detach().to(**f32).contiguous() if self.use_depth else None
        if self.use_depth and self.gt_depth.numel() != self.W * self.H:
            raise RuntimeError("gflow_b200: gt_depth must have H*W elements")
        self.pixel_mask = None if pixel_mask is None else pixel_mask.to(dev).to(torch.uint8).contiguous()
        self.still_mask = None if still_mask is None else still_mask.to(dev).to(torch.uint8).contiguous()
        if self.still_mask is not None and self.still_mask.numel() > self.N:
            raise RuntimeError("gflow_b200: still_mask is longer than the number of Gaussians")
        # camera-only stage: compact raw copy of the moving subset (attributes are frozen in this stage), the
        # cumulative pixel mask it carves, and a workspace for its per-iteration render
        self.sub = self.dyn_mask = self.sub_ws = None
        self.sub_N = self.sub_capacity = 0
        if cfg.camera_only and tentative_still is not None:
            n = tentative_still.shape[0]
            sel = ~tentative_still.to(dev).bool()
            if int(sel.sum()) > 0:
                self.sub = {k: fitter.attrs[k].data[:n][sel].contiguous() for k in ATTRS}
                self.sub_N = int(self.sub["xyz"].shape[0])
                self.sub_capacity = int(sub_capacity) if sub_capacity is not None else 8 * self.sub_N + 16384
                self.dyn_mask = (torch.ones(self.H, self.W, dtype=torch.uint8, device=dev) if self.pixel_mask is None
                                 else self.pixel_mask.clone())
            elif cfg.background > 0:
                # an empty moving set renders the bare background, whose grey value is > 0: the reference then
                # drops EVERY pixel from the losses (trainer.py:446-451)
                self.pixel_mask = torch.zeros(self.H, self.W, dtype=torch.uint8, device=dev)
        # loss_scale runs over the still (camera-only) / moving (full stage) Gaussians once a still mask exists
        self.scale_sel = None
        if cfg.lambda_scale and self.still_mask is not None:
            sel = torch.ones(self.N, dtype=torch.uint8, device=dev)
            n = self.still_mask.numel()
            sel[:n] = self.still_mask if cfg.camera_only else (1 - self.still_mask)
            self.scale_sel = sel
        # still / flow terms: everything that is constant over the call is folded on the host once
        self.still_ref = self.still_sel = self.flow_target = self.flow_sel = None
        self.still_count = self.flow_count = 0
        if cfg.lambda_still and prev is not None and prev.last_still_mask is not None:
            self.still_ref = prev.last_xyz.detach().to(**f32).contiguous()
            self.still_sel = prev.last_still_mask.to(dev).to(torch.uint8).contiguous()
            self.still_count = int(self.still_sel.sum())
        if cfg.lambda_flow and prev is not None and prev.gt_flow is not None:
            pv = PrevFrame(last_uv=prev.last_uv.detach().to(**f32), gt_flow=prev.gt_flow.detach().to(**f32))
            and_mask = pv.flow_and_mask(self.W, self.H, None if still_mask is None else still_mask.to(dev).bool(),
                                        cfg.camera_only)
            self.flow_target = pv.flow_target(and_mask)
            self.flow_sel = and_mask.to(torch.uint8).contiguous()
            self.flow_count = int(self.flow_sel.sum())
        # depth_a / depth_b live in one 2-float tensor on the device side; copied back after every run()
        self.depth_ab = torch.cat([fitter.depth_a.data.reshape(1), fitter.depth_b.data.reshape(1)]).contiguous()
        self.dbg_grads = torch.zeros(self.N, 14, **f32) if debug else None
        self.dbg_act = torch.zeros(self.N, 14, **f32) if debug else None
        self.iters = int(cfg.iterations)
        self.done = 0
        self.num_points0 = int(cfg.num_points or self.N)
        self.adam_t0, self.constant_lr, self.freeze_camera = 0, 0, 0  # change when a densification re-creates Adam
        self._densifier = None
        self.capacity = int(capacity) if capacity is not None else 6 * self.N + 65536
        self.ws = None
        with self._device_guard(), self._stream_guard():
            self._alloc(self.capacity)  # allocated under the loop's stream: the caching allocator keys blocks by stream
            capi.check(self.lib.gfb_fit_init(ctypes.addressof(self.problem), self.ws.data_ptr(), self.capacity, self.iters,
                                             self._stream()), "fit init")

    # ------------------------------------------------------------------ plumbing
    # (separate methods so the CPU test-suite can drive this host logic against the emulated kernel library;
    #  the product class itself only accepts CUDA tensors and only loads libgflow_b200.so)
    def _require_device(self, dev) -> None:
        if dev.type != "cuda":
            raise RuntimeError("gflow_b200: the native fit loop needs CUDA tensors (no CPU fallback exists)")

    def _library(self):
        return capi.load()

    def _stream(self) -> int:
        return self.stream.cuda_stream if self.stream is not None else ops._stream()

    def _device_guard(self):
        return ops._on_device(self.dev)

    def _stream_guard(self):
        """torch-side copies (snapshots, roll-backs, status reads) must be ordered on the loop's stream too."""
        import contextlib

        return torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def _alloc(self, capacity: int) -> None:
        lay = capi.FitLayout()
        capi.check(self.lib.gfb_fit_get_layout(self.N, self.W, self.H, capacity, self.iters, ctypes.addressof(lay)),
                   "fit layout")
        old, old_lay = self.ws, getattr(self, "lay", None)
        self.ws = torch.empty(lay.total, dtype=torch.uint8, device=self.dev)
        if old is not None:  # everything before `uv` (status, losses, camera, Adam state) is capacity independent
            self.ws[: lay.uv].copy_(old[: old_lay.uv])
        self.lay, self.capacity = lay, capacity
        f, c, pr = self.fitter, self.cfg, capi.FitProblem()
        for k in ATTRS:
            setattr(pr, k, f.attrs[k].data.data_ptr())
        pr.pose, pr.depth_ab, pr.intr = f.pose.data.data_ptr(), self.depth_ab.data_ptr(), f.intr.data_ptr()
        pr.gt_image, pr.gt_depth = self.gt_image.data_ptr(), ops._ptr(self.gt_depth)
        pr.pixel_mask, pr.still_mask = ops._ptr(self.pixel_mask), ops._ptr(self.still_mask)
        pr.dbg_grads, pr.dbg_act = ops._ptr(self.dbg_grads), ops._ptr(self.dbg_act)
        pr.scale_sel = ops._ptr(self.scale_sel)
        pr.sub_N, pr.sub_capacity = self.sub_N, self.sub_capacity
        if self.sub_N > 0:
            need = self.lib.gfb_fit_sub_workspace_bytes(self.sub_N, self.W, self.H, self.sub_capacity)
            if self.sub_ws is None or self.sub_ws.numel() != need:
                self.sub_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
            pr.sub_xyz, pr.sub_scale, pr.sub_rotate = (self.sub[k].data_ptr() for k in ("xyz", "scale", "rotate"))
            pr.sub_opacity, pr.sub_rgb = self.sub["opacity"].data_ptr(), self.sub["rgb"].data_ptr()
            pr.dyn_mask, pr.sub_workspace = self.dyn_mask.data_ptr(), self.sub_ws.data_ptr()
        pr.still_ref, pr.still_sel = ops._ptr(self.still_ref), ops._ptr(self.still_sel)
        pr.flow_target, pr.flow_sel = ops._ptr(self.flow_target), ops._ptr(self.flow_sel)
        pr.N, pr.W, pr.H = self.N, self.W, self.H
        pr.n_still = 0 if self.still_mask is None else self.still_mask.numel()
        pr.n_still_ref = 0 if self.still_sel is None else min(self.N, self.still_sel.numel())
        pr.n_flow = 0 if self.flow_sel is None else min(self.N, self.flow_sel.numel())
        pr.still_count, pr.flow_count = self.still_count, self.flow_count
        pr.lambda_still, pr.lambda_flow = float(c.lambda_still), float(c.lambda_flow)
        pr.total_iters = self.iters
        pr.camera_only, pr.freeze_rgb, pr.use_ssim = int(c.camera_only), int(c.freeze_rgb), int(c.use_ssim)
        pr.adam_t0, pr.constant_lr, pr.freeze_camera = self.adam_t0, self.constant_lr, self.freeze_camera
        pr.bg, pr.nearest, pr.extent = float(c.background), 0.2, 1.3
        pr.lr, pr.lr_camera = float(c.lr), float(c.lr_camera)
        pr.lambda_rgb, pr.lambda_depth = float(c.lambda_rgb), float(c.lambda_depth if self.use_depth else 0.0)
        pr.lambda_var, pr.lambda_scale = float(c.lambda_var), float(c.lambda_scale)
        pr.beta1, pr.beta2, pr.eps = 0.9, 0.999, 1e-8
        pr.depth_den_min = float(c.depth_den_min)
        self.problem = pr

    def _view(self, off: int, count: int, dtype=torch.float32) -> torch.Tensor:
        return self.ws[off: off + 4 * count].view(dtype)

    def _snapshot(self):
        f = self.fitter
        return ([f.attrs[k].data.clone() for k in ATTRS], f.pose.data.clone(), self.depth_ab.clone(),
                self.ws[: self.lay.uv].clone(), self.done, None if self.dyn_mask is None else self.dyn_mask.clone())

    def _restore(self, snap) -> None:
        attrs, pose, ab, head, done, dyn = snap
        if dyn is not None:
            self.dyn_mask.copy_(dyn)
        for k, t in zip(ATTRS, attrs):
            self.fitter.attrs[k].data.copy_(t)
        self.fitter.pose.data.copy_(pose)
        self.depth_ab.copy_(ab)
        self.ws[: self.lay.uv].copy_(head)
        self.done = done

    # ------------------------------------------------------------------ public
    def enqueue(self, n_iters: Optional[int] = None) -> int:
        """Enqueues the next chunk (at most cfg.check_every iterations) on this loop's stream and returns without
        waiting; check() must follow before the next enqueue.  Returns the number of iterations enqueued."""
        if self._pending is not None:
            raise RuntimeError("gflow_b200: check() the previous chunk before enqueueing the next one")
        n = min(max(1, int(self.cfg.check_every)), self.iters - self.done)
        if n_iters is not None:
            n = min(n, int(n_iters))
        if n <= 0:
            return 0
        with self._device_guard(), self._stream_guard():
            snap = self._snapshot()
            capi.check(self.lib.gfb_fit_iterate(ctypes.addressof(self.problem), self.ws.data_ptr(), self.capacity,
                                                self.iters, self.done, n, self._stream()), "fit iterate")
        self.done += n
        self._pending = snap
        return n

    def check(self) -> bool:
        """Waits for the enqueued chunk (one 16-byte read of the status block) and compares the largest
        intersection counts seen with the workspace capacities.  True: the chunk stands.  False: its tiles were
        truncated -- parameters, pose, Adam state and pixel mask are back at the chunk start, the workspace has
        grown, and the chunk has to be enqueued again."""
        if self._pending is None:
            return True
        snap, self._pending = self._pending, None
        with self._device_guard(), self._stream_guard():
            st = self._view(self.lay.status, 16, torch.int32)[:4].tolist()  # synchronises this stream
            k_max, sub_k_max = st[2], st[3]
            if k_max <= self.capacity and sub_k_max <= self.sub_capacity:
                return True
            self._restore(snap)
            if sub_k_max > self.sub_capacity:
                self.sub_capacity = int(1.5 * sub_k_max) + 16384
            self._alloc(int(1.5 * k_max) + 65536 if k_max > self.capacity else self.capacity)
        return False

    def finish(self) -> None:
        """Copies depth_a / depth_b back into the fitter's parameters (they live in one 2-float device tensor)."""
        with self._device_guard(), self._stream_guard():
            self.fitter.depth_a.data.copy_(self.depth_ab[0:1])
            self.fitter.depth_b.data.copy_(self.depth_ab[1:2])

    def run(self, n_iters: int) -> None:
        """Runs `n_iters` iterations in chunks of cfg.check_every.  After each chunk the largest intersection
        count seen is compared with the workspace capacity; a chunk that overflowed is rolled back and redone with
        a larger workspace (its tiles were truncated)."""
        end = min(self.iters, self.done + int(n_iters))
        while self.done < end:
            self.enqueue(end - self.done)
            self.check()  # False: rolled back, self.done is at the chunk start again and the loop re-enqueues it
        self.finish()

    def _make_densifier(self):
        return _densify.Densifier(self.W, self.H, self.dev)

    def densify(self, error_threshold: float = 1e-3, percent: float = 0.1, mask: Optional[torch.Tensor] = None,
                uniform_error: bool = False, seed: int = 0) -> int:
        """densify_by_pixels (trainer.py:878-951) between two iterations, entirely on the device: error map of
        the last iteration's render (or ones, with a mask: the occlusion densification of trainer.py:562-564) ->
        weighted draw -> new Gaussians appended -> optimiser re-created the way the reference does it (Adam state
        zeroed, constant lr, pose / depth_a / depth_b no longer updated).  Returns the number added."""
        if self.gt_depth is None:
            raise RuntimeError("gflow_b200: densification back-projects with the depth prior; gt_depth is required")
        if self.done == 0 and not uniform_error:
            raise RuntimeError("gflow_b200: no iteration has run yet, there is no error map to densify from")
        if self._densifier is None:
            self._densifier = self._make_densifier()
        d, f = self._densifier, self.fitter
        if self._pending is not None:
            raise RuntimeError("gflow_b200: check() the enqueued chunk before densifying")
        with self._device_guard(), self._stream_guard():
            if uniform_error:
                err = torch.ones(self.H, self.W, dtype=torch.float32, device=self.dev)
            else:
                err = d.rgb_error_map(self.rendered(), self.gt_image, self.pixel_mask)
            extr = self.camera()[:12].reshape(3, 4)  # camera after the last update = get_extr() at this point
            new = d.sample(err, self.gt_image, self.gt_depth, f.intr, extr, self.num_points0, error_threshold, percent, mask=mask,
                           seed=seed)
            if new is None:
                return 0
            count = int(new["xyz"].shape[0])
            self._last_densify_pixels = new["pixels"]
            for k in ATTRS:
                f.attrs[k] = torch.nn.Parameter(torch.cat([f.attrs[k].data, new[k]], dim=0))
            head = self.ws[: self.lay.adam_m].clone()  # status | loss sums | camera | loss history: N independent
            # projected centres / depths of the last forward survive the re-allocation: a stage that ENDS with a
            # densification hands them to the caller (still mask, last_uv, move_seg bookkeeping), and the fresh workspace
            # would otherwise return uninitialised rows there
            n_old = self.N
            old_uv, old_depth = (self.last_uv(), self.last_depth()) if self.done > 0 else (None, None)
            self.N += count
            if self.scale_sel is not None:
                self.scale_sel = torch.cat([self.scale_sel, torch.ones(count, dtype=torch.uint8, device=self.dev)])
            if self.dbg_grads is not None:  # per-Gaussian debug outputs grow with N
                self.dbg_grads = torch.zeros(self.N, 14, dtype=torch.float32, device=self.dev)
                self.dbg_act = torch.zeros(self.N, 14, dtype=torch.float32, device=self.dev)
            self.adam_t0, self.constant_lr, self.freeze_camera = self.done, 1, 1
            self.ws = None
            self._alloc(self.capacity)
            capi.check(self.lib.gfb_fit_init(ctypes.addressof(self.problem), self.ws.data_ptr(), self.capacity, self.iters,
                                             self._stream()), "fit init after densify")
            self.ws[: self.lay.adam_m].copy_(head)
            uv_new = self._view(self.lay.uv, 2 * self.N).reshape(self.N, 2)
            d_new = self._view(self.lay.depth, self.N)
            if old_uv is not None:
                uv_new[:n_old] = old_uv
                d_new[:n_old] = old_depth.reshape(-1)
            else:
                uv_new[:n_old] = 0.0
                d_new[:n_old] = 0.0
            # a new Gaussian is the back-projection of its pixel at the prior depth under the current camera
            # (trainer.py:904-933), so it projects onto that pixel at that depth
            # (geometry.py:104-116 back-projects with fx on both axes, so v lands on cy + (row - cy) fy / fx)
            px = new["pixels"].long()
            fx, fy, cy = f.intr[0], f.intr[1], f.intr[3]
            row = torch.div(px, self.W, rounding_mode="floor").to(torch.float32)
            uv_new[n_old:, 0] = (px % self.W).to(torch.float32)
            uv_new[n_old:, 1] = cy + (row - cy) * (fy / fx)
            d_new[n_old:] = self.gt_depth.reshape(-1)[px]
        return count

    def last_uv(self) -> torch.Tensor:
        """(N,2) projected centres of the last iteration's forward pass."""
        return self._view(self.lay.uv, 2 * self.N).reshape(self.N, 2).clone()

    def last_depth(self) -> torch.Tensor:
        return self._view(self.lay.depth, self.N).reshape(self.N, 1).clone()

    def pixel_keep_mask(self) -> Optional[torch.Tensor]:
        """(H,W) bool: pixels that still take part in the losses (static mask, or the one the moving subset carved)."""
        m = self.dyn_mask if self.dyn_mask is not None else self.pixel_mask
        return None if m is None else m.bool()

    def loss_history(self) -> torch.Tensor:
        """(iterations done, 8): total, mse, ssim, depth, var, scale, still, flow per iteration."""
        return self._view(self.lay.loss_hist, self.iters * 8).reshape(self.iters, 8)[: self.done]

    def status(self) -> torch.Tensor:
        return self._view(self.lay.status, 16, torch.int32)

    def d_pose(self) -> torch.Tensor:
        return self._view(self.lay.status + 32, 7)

    def camera(self) -> torch.Tensor:
        return self._view(self.lay.cam, 16)

    def rendered(self) -> torch.Tensor:
        """(C,H,W) output of the last iteration's blend: rgb, plus the depth map as channel 3 with a depth term."""
        C = 4 if self.use_depth else 3
        return self._view(self.lay.out, C * self.H * self.W).reshape(C, self.H, self.W)


_SIDE_STREAMS: Dict[tuple, list] = {}


def fit_frames_concurrently(fitters, targets, cfg: FitConfig, streams=None, loop_cls=None):
    """Runs the native loop of several independent frames of ONE GPU side by side, each on its own CUDA stream.

    A 60k-Gaussian / 480p iteration under-fills a B200 (1620 tiles against 148 SMs x several resident CTAs, and a
    tail of small per-Gaussian kernels and one-warp kernels), so kernels of different frames are left to overlap:
    chunks are enqueued round-robin without waiting, then checked in the same order (one status read each).
    `fitters[i]` is trained against `targets[i] = (gt_image, gt_depth_or_None)`.  Returns the loops (loss
    histories, final state).  No densification in this mode."""
    loop_cls = loop_cls or NativeFitLoop
    if streams is None:
        # the same streams for every group of frames: the caching allocator keeps one pool per stream, and a fresh stream
        # per call would send every workspace allocation of every group to cudaMalloc
        dev_ = fitters[0].attrs["xyz"].device
        pool = _SIDE_STREAMS.setdefault((dev_.type, dev_.index), [])
        while len(pool) < len(fitters):
            pool.append(torch.cuda.Stream(device=dev_))
        streams = pool[:len(fitters)]
        cur = torch.cuda.current_stream(fitters[0].attrs["xyz"].device)
        for s_ in streams:  # the fitters' tensors were produced on the current stream
            s_.wait_stream(cur)
    loops = [loop_cls(f, gi, gd, cfg, stream=s_) for f, (gi, gd), s_ in zip(fitters, targets, streams)]
    while any(lp.done < lp.iters for lp in loops):
        for lp in loops:
            lp.enqueue()
        for lp in loops:
            lp.check()  # False = rolled back and grown; the next round re-enqueues that chunk
    for lp in loops:
        lp.finish()
    if streams and streams[0] is not None:
        cur = torch.cuda.current_stream(fitters[0].attrs["xyz"].device)
        for s_ in streams:
            cur.wait_stream(s_)
    return loops


def fit_sequence_sharded(state0: Optional[Dict[str, torch.Tensor]], intr: torch.Tensor, frame_targets, W: int, H: int,
                         cfg: FitConfig, device: torch.device, concurrent_frames: int = 1):
    """Frame-sharded sequence fit: every rank fits its own contiguous chunk of frames,
    each frame starting from the broadcast frame-0 state.

    `frame_targets(i)` returns (gt_image, gt_depth_or_None, pose0 (7,)) for frame i and `len(frame_targets)`
    is the number of frames; rank 0 passes `state0` (activated attributes are NOT expected: raw
    parameters as in the checkpoint), other ranks pass None.  Collectives: one broadcast before the
    loop, one all_gather after it.  Returns (local results keyed by frame index, gathered final frames
    on rank 0).  With cfg.native and concurrent_frames > 1 a rank works on that many of its frames at a time,
    one CUDA stream each (fit_frames_concurrently).
    """
    import torch.distributed as dist

    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    state = _frames.broadcast_state(state0, src=0, device=device) if world > 1 else state0
    mine = _frames.shard_frames(len(frame_targets), world, rank)
    results = {}
    last_img, last_pose = None, None
    mine = list(mine)
    group = max(1, int(concurrent_frames)) if (cfg.native and not cfg.densify_interval) else 1
    for g0 in range(0, len(mine), group):
        idx = mine[g0:g0 + group]
        if group == 1:
            gt_image, gt_depth, pose0 = frame_targets(idx[0])
            fitter = FrameFitter(state, intr.to(device), pose0.to(device), W, H)
            results[idx[0]] = fitter.train(gt_image.to(device), None if gt_depth is None else gt_depth.to(device), cfg)
        else:
            tg = [frame_targets(i) for i in idx]
            fitters = [FrameFitter(state, intr.to(device), t[2].to(device), W, H) for t in tg]
            loops = fit_frames_concurrently(fitters, [(t[0].to(device), None if t[1] is None else t[1].to(device)) for t in tg], cfg)
            for i, f, lp in zip(idx, fitters, loops):
                r = FitResult(losses=[float(v) for v in lp.loss_history()[:, 0].cpu()])
                with torch.no_grad():
                    r.image, _, r.uv = f.render(cfg.background, want_depth=False)
                    r.pose = f.pose.detach().clone()
                results[i] = r
        last_img, last_pose = results[idx[-1]].image, pose_to_extr(results[idx[-1]].pose)
    gathered = None
    if world > 1:
        if last_img is None:  # a rank without frames still takes part in the collective
            last_img = torch.zeros(3, H, W, device=device)
            last_pose = torch.zeros(3, 4, device=device)
        gathered = _frames.gather_frames(last_img, last_pose, dst=0)
    return results, gathered
