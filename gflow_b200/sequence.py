"""Sequential per-frame fit of a video -- the frame loop of /root/reference/gflow/fit_video.py:104-349 with
the per-frame optimisation stages handed to gflow_b200.fit (operator path or native kernels).

What fit_video.py does per sequence, and where it lives here:

  frame 0     trainer.train(iterations_first, lr, lr_camera, densify every `densify_interval`)   fit_first()
  frame i>0   set_gt_image / depth / flow (+ load_camera(extr))                                   fit_next()
              [camera_first] trainer.train(iterations_camera, lr_camera_after, camera_only=True)
              trainer.train(iterations_after, lr_after, lr_camera=0, mask=occ_mask)
  inside train(): pre-update flow warp of the moving Gaussians (trainer.py:348-376),              fit.warp_moving_by_flow
                  the optimisation loop (trainer.py:387-571),                                     fit.FrameFitter.train
                  post-update bookkeeping (trainer.py:587-625): still / tentative masks from the  _post_update()
                  move mask at the projected centres, last_uv / last_xyz / last_still_mask

Not rebuilt: prior readers, concave-hull segmentation, trajectory renders, PNG / MP4
dumps -- callers pass tensors and get tensors.  Hyper-parameter defaults are scripts/fit_video.sh:16-41.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import fit as _fit

ATTRS = _fit.ATTRS


@dataclass
class SequenceConfig:
    """scripts/fit_video.sh:16-41 (the shipped launch line)."""
    num_points: int = 50000
    lr: float = 4e-3
    lr_camera: float = 0.0
    iterations_first: int = 500
    lr_after: float = 4e-3
    iterations_after: int = 300
    camera_first: bool = True
    lr_camera_after: float = 1e-3
    iterations_camera: int = 150
    densify_interval: int = 150
    densify_times: int = 2
    densify_occ_percent: float = 0.5
    densify_interval_after: int = 100
    densify_times_after: int = 2
    densify_err_thre: float = 1e-2
    densify_err_percent: float = 1.0
    lambda_rgb: float = 1.0
    lambda_depth: float = 0.1
    lambda_var: float = 50.0
    lambda_still: float = 0.0
    lambda_flow: float = 0.01
    lambda_scale: float = 0.0
    background: float = 0.0
    use_ssim: bool = True
    native: bool = True
    check_every: int = 50
    seed: int = 0


@dataclass
class FrameOutput:
    losses: Dict[str, List[float]] = field(default_factory=dict)   # per stage
    image: Optional[torch.Tensor] = None
    pose: Optional[torch.Tensor] = None
    num_points: int = 0
    chunk_head: bool = False   # fit_video_sharded: first frame of a rank's chunk (its tracks / masks start afresh)


class SequenceFitter:
    """Holds what SimpleGaussian keeps between frames: raw attributes, pose, still / tentative masks and the
    last_* tensors (trainer.py:587-625)."""

    def __init__(self, raw0: Dict[str, torch.Tensor], intr: torch.Tensor, pose0: torch.Tensor, W: int, H: int,
                 cfg: SequenceConfig):
        self.cfg, self.W, self.H = cfg, int(W), int(H)
        self.attrs = {k: raw0[k].detach().clone() for k in ATTRS}
        self.intr = intr.detach().clone()
        self.pose = pose0.detach().clone()
        self.still_mask = self.still_mask_tentative = self.last_still_mask = None
        self.last_uv = self.last_xyz = self.last_depth = None
        self.frame = 0

    # ------------------------------------------------------------------ one train() call of the reference
    def _stage(self, name: str, gt_image, gt_depth, fc: _fit.FitConfig, move_mask, gt_flow=None, occ_mask=None) -> List[float]:
        later = self.last_xyz is not None
        prev = None
        if later:
            prev = _fit.PrevFrame(last_xyz=self.last_xyz, last_still_mask=self.last_still_mask, last_uv=self.last_uv,
                                  gt_flow=gt_flow)
            if not fc.camera_only and self.still_mask is not None:  # pre-update processing, trainer.py:348-376
                self.attrs["xyz"] = _fit.warp_moving_by_flow(self.attrs["xyz"], prev, gt_depth, self.intr,
                                                             _fit.pose_to_extr(self.pose), self.W, self.H)
        fitter = _fit.FrameFitter(self.attrs, self.intr, self.pose, self.W, self.H)
        fc.freeze_rgb = later  # trainer.py:537-540
        res = fitter.train(gt_image, gt_depth, fc, pixel_mask=(~move_mask) if fc.camera_only else None,
                           still_mask=self.still_mask, prev=prev,
                           tentative_still=self.still_mask_tentative if fc.camera_only else None, occlusion_mask=occ_mask)
        self.attrs = {k: fitter.attrs[k].data for k in ATTRS}
        self.pose = fitter.pose.data
        self._last_result = res
        if not fc.camera_only:
            self._post_update(fitter, res, move_mask)
        return res.losses

    def _post_update(self, fitter, res, move_mask) -> None:
        """trainer.py:587-625.  `uv` is the one of the LAST iteration's forward pass (before its parameter update),
        exactly the variable the reference still holds when the loop ends."""
        uv = res.last_uv if res.last_uv is not None else res.uv
        W, H = self.W, self.H
        within = (uv[:, 0] > 0) & (uv[:, 0] < W - 1) & (uv[:, 1] > 0) & (uv[:, 1] < H - 1)
        labels = ~move_mask[uv[within][:, 1].long(), uv[within][:, 0].long()]
        N = uv.shape[0]
        still = torch.ones(N, dtype=torch.bool, device=uv.device)
        still[within] = labels
        self.still_mask_tentative = still.clone()
        if self.last_still_mask is not None:
            n = self.last_still_mask.shape[0]
            still[:n] = self.last_still_mask
        self.still_mask = still
        self.last_still_mask = still.clone()
        self.last_uv = uv.detach().clone()
        self.last_depth = None if res.last_depth is None else res.last_depth.detach().clone()
        self.last_xyz = self.attrs["xyz"].detach().clone()

    def _fit_config(self, **kw) -> _fit.FitConfig:
        c = self.cfg
        base = dict(lambda_rgb=c.lambda_rgb, lambda_depth=c.lambda_depth, background=c.background, use_ssim=c.use_ssim,
                    native=c.native, check_every=c.check_every, densify_err_thre=c.densify_err_thre,
                    densify_err_percent=c.densify_err_percent, densify_occ_percent=c.densify_occ_percent,
                    num_points=c.num_points, densify_seed=c.seed + 1000 * self.frame)
        base.update(kw)
        return _fit.FitConfig(**base)

    # ------------------------------------------------------------------ public
    def fit_first(self, gt_image, gt_depth, move_mask) -> FrameOutput:
        """fit_video.py:119-142."""
        c = self.cfg
        fc = self._fit_config(iterations=c.iterations_first, lr=c.lr, lr_camera=c.lr_camera, lambda_var=c.lambda_var,
                              lambda_scale=c.lambda_scale, densify_interval=c.densify_interval, densify_times=c.densify_times)
        out = FrameOutput(losses={"first": self._stage("first", gt_image, gt_depth, fc, move_mask)})
        return self._finish(out)

    def fit_chunk_head(self, gt_image, gt_depth, move_mask, have_camera_prior: bool) -> FrameOutput:
        """First frame of a later chunk in fit_video_sharded: the broadcast Gaussians describe frame 0, the chunk's
        first frame was taken from another viewpoint.  Without a camera prior the pose is first moved there by a
        camera-only stage (frame 0's recipe keeps lr_camera at 0 in the shipped configuration and would leave the
        chunk's camera at pose0), then the frame is fitted with the first-frame recipe."""
        c = self.cfg
        cam_losses = None
        if not have_camera_prior and c.iterations_camera > 0:
            move = torch.zeros(self.H, self.W, dtype=torch.bool, device=gt_image.device) if move_mask is None else move_mask
            fc = self._fit_config(iterations=c.iterations_camera, lr=1e-2, lr_camera=c.lr_camera_after, lambda_var=0.0,
                                  lambda_still=0.0, lambda_flow=0.0, camera_only=True)
            cam_losses = self._stage("camera", gt_image, gt_depth, fc, move)
        out = self.fit_first(gt_image, gt_depth, move_mask)
        out.chunk_head = True
        if cam_losses is not None:
            out.losses["camera"] = cam_losses
        return out

    def fit_next(self, gt_image, gt_depth, gt_flow, move_mask, occ_mask=None, extr: Optional[torch.Tensor] = None) -> FrameOutput:
        """fit_video.py:242-315.  `extr` (3,4): load_camera(extr=...) when camera priors are used (load_extr)."""
        c = self.cfg
        self.frame += 1
        if extr is not None:
            self.pose = _fit.extr_to_pose(extr.detach().float().cpu()).to(self.pose.device)
        out = FrameOutput()
        if c.camera_first:
            fc = self._fit_config(iterations=c.iterations_camera, lr=1e-2, lr_camera=c.lr_camera_after, lambda_var=0.0,
                                  lambda_still=0.0, lambda_flow=c.lambda_flow, camera_only=True)
            out.losses["camera"] = self._stage("camera", gt_image, gt_depth, fc, move_mask, gt_flow)
        if c.iterations_after > 0:
            fc = self._fit_config(iterations=c.iterations_after, lr=c.lr_after, lr_camera=0.0, lambda_var=c.lambda_var,
                                  lambda_still=c.lambda_still, lambda_scale=c.lambda_scale, lambda_flow=c.lambda_flow,
                                  densify_interval=c.densify_interval_after, densify_times=c.densify_times_after)
            out.losses["all"] = self._stage("all", gt_image, gt_depth, fc, move_mask, gt_flow, occ_mask)
        return self._finish(out)

    def _finish(self, out: FrameOutput) -> FrameOutput:
        out.image, out.pose = self._last_result.image, self.pose.detach().clone()
        out.num_points = int(self.attrs["xyz"].shape[0])
        return out

    def state(self):
        """checkpoint.FrameState of the current frame (trainer.py:252-272)."""
        from . import checkpoint as _ck

        return _ck.FrameState(attributes={k: v.detach() for k, v in self.attrs.items()}, intr=self.intr,
                              extr=_fit.pose_to_extr(self.pose), width=self.W, height=self.H, still_mask=self.still_mask,
                              last_uv=self.last_uv)


def fit_video_sharded(state0: Optional[Dict[str, torch.Tensor]], intr: torch.Tensor, pose0: torch.Tensor, num_frames: int,
                      frame_inputs, W: int, H: int, cfg: SequenceConfig, device):
    """The frame loop of fit_video.py sharded by frame across the ranks of one box: one NCCL broadcast
    of the frame-0 Gaussian state at sequence start, no collective inside the loop, one gather of the per-frame
    outputs (final render + pose) at the end.

    Rank r owns the contiguous chunk frames.shard_frames(num_frames, world, r) and runs it SEQUENTIALLY with the
    reference's state carry-over inside the chunk: the following frames get camera-only + full stages, flow warp and
    still-mask bookkeeping; the chunk's FIRST frame starts from the broadcast frame-0 state -- its camera comes from the
    `extr` prior when there is one, otherwise from a camera-only stage (SequenceFitter.fit_chunk_head) -- and is then
    fitted with the first-frame recipe.

    THROUGHPUT MODE, not the reference's semantics: the reference is strictly sequential, chunks here do not see each
    other.  Gaussian identities, still masks, last_uv and therefore tracks / trajectories are continuous only INSIDE
    a chunk, and poses of different chunks are only comparable when every chunk head got an `extr` prior or converged
    in its camera stage.  Outputs of chunk heads carry `chunk_head=True` so a consumer can cut tracks there.

    frame_inputs(i) -> dict(image (H,W,3), depth (H,W,1), move_mask (H,W) bool[, flow (H,W,2), occ_mask (H,W,1), extr (3,4)]).
    Returns (outputs of the local frames keyed by frame index, list of (image, extr) per rank on rank 0 or None).
    """
    import torch.distributed as dist

    from . import frames as _frames

    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    state = _frames.broadcast_state(state0, src=0, device=device) if world > 1 else {k: v.to(device) for k, v in state0.items()}
    mine = list(_frames.shard_frames(num_frames, world, rank))
    outputs: Dict[int, FrameOutput] = {}
    seq = None
    for j, i in enumerate(mine):
        fi = frame_inputs(i)
        to = lambda t: None if t is None else t.to(device)  # noqa: E731
        if j == 0:
            seq = SequenceFitter(state, intr.to(device), pose0.to(device), W, H, cfg)
            if fi.get("extr") is not None:
                seq.pose = _fit.extr_to_pose(fi["extr"].detach().float().cpu()).to(device)
            if i == 0:
                outputs[i] = seq.fit_first(to(fi["image"]), to(fi["depth"]), to(fi["move_mask"]))
            else:
                outputs[i] = seq.fit_chunk_head(to(fi["image"]), to(fi["depth"]), to(fi["move_mask"]),
                                                have_camera_prior=fi.get("extr") is not None)
        else:
            outputs[i] = seq.fit_next(to(fi["image"]), to(fi["depth"]), to(fi["flow"]), to(fi["move_mask"]),
                                      occ_mask=to(fi.get("occ_mask")), extr=fi.get("extr"))
    gathered = None
    if world > 1:
        last = outputs[mine[-1]] if mine else None
        img = last.image if last is not None and last.image is not None else torch.zeros(3, H, W, device=device)
        extr = _fit.pose_to_extr(last.pose) if last is not None else torch.zeros(3, 4, device=device)
        gathered = _frames.gather_frames(img, extr, dst=0)
    return outputs, gathered
