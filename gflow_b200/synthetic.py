"""Seeded synthetic Gaussians + camera shared by tests and bench.py.

Everything is generated on the CPU with a ``torch.Generator`` so the same scene is
reproduced bit for bit here, on the GPU box and by the oracle.
Camera defaults follow /root/reference/gflow/trainer.py:37-41 (fov 90 deg:
fx = W/2, fy = H/2, principal point at the image centre).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

# named BASELINE.json configs: (N, W, H)
CONFIGS = {
    "cfg1": (1_000, 256, 256),
    "cfg2": (60_000, 854, 480),
    "cfg5": (200_000, 1280, 720),
}


@dataclass
class Scene:
    xyz: torch.Tensor       # (N,3) world positions
    scale: torch.Tensor     # (N,3) activated (positive) scales
    rotate: torch.Tensor    # (N,4) unit quaternions, w first
    opacity: torch.Tensor   # (N,1) in (0,1)
    rgb: torch.Tensor       # (N,3) in (0,1)
    intr: torch.Tensor      # (4,)  fx fy cx cy
    extr: torch.Tensor      # (3,4) world->camera
    W: int
    H: int
    bg: float = 0.0

    def to(self, device):
        kw = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()}
        return Scene(**kw)

    def tensors(self):
        return self.xyz, self.scale, self.rotate, self.opacity, self.rgb


def _rodrigues(w: torch.Tensor) -> torch.Tensor:
    th = float(w.norm())
    if th < 1e-12:
        return torch.eye(3)
    k = w / th
    K = torch.tensor([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])
    return torch.eye(3) + math.sin(th) * K + (1.0 - math.cos(th)) * (K @ K)


def make_camera(W: int, H: int, gen: torch.Generator, fx=None, fy=None):
    """fov-90 intrinsics and a small random rigid world->camera motion."""
    fx = 0.5 * W if fx is None else fx
    fy = 0.5 * H if fy is None else fy
    intr = torch.tensor([fx, fy, 0.5 * W, 0.5 * H], dtype=torch.float32)
    aa = (torch.rand(3, generator=gen) - 0.5) * 0.1
    t = (torch.rand(3, generator=gen) - 0.5) * 0.2
    extr = torch.cat([_rodrigues(aa), t.reshape(3, 1)], dim=1).to(torch.float32)
    return intr, extr


def make_scene(N: int, W: int, H: int, seed: int = 0, profile: str = "synthetic", outside_frac: float = 0.05,
               bg: float = 0.0) -> Scene:
    """pixels ~ U[0,W)x[0,H), z ~ U[1,5], sigma_px log-uniform.

    profile "synthetic": sigma_px in [0.5,4];  "gflow": sigma_px in [0.4,1.2]
    (matches the init of /root/reference/gflow/trainer.py:223-225).
    ``outside_frac`` of the points are pushed outside the frustum / behind the
    camera to exercise culling.
    """
    gen = torch.Generator().manual_seed(seed)
    intr, extr = make_camera(W, H, gen)
    fx, fy, cx, cy = (float(v) for v in intr)
    u = torch.rand(N, generator=gen) * W
    v = torch.rand(N, generator=gen) * H
    z = 1.0 + 4.0 * torch.rand(N, generator=gen)
    n_out = int(N * outside_frac)
    if n_out > 0:
        kind = torch.randint(0, 3, (n_out,), generator=gen)
        idx = torch.randperm(N, generator=gen)[:n_out]
        u[idx] = torch.where(kind == 0, u[idx] + 3.0 * W, u[idx])
        v[idx] = torch.where(kind == 1, v[idx] - 3.0 * H, v[idx])
        z[idx] = torch.where(kind == 2, -z[idx], z[idx])
    # back-project like /root/reference/gflow/utils/geometry.py:104-113 (pix2world)
    xc = (u - cx) / fx * z
    yc = (v - cy) / fy * z
    pc = torch.stack([xc, yc, z], dim=1)
    R, t = extr[:, :3], extr[:, 3]
    xyz = (pc - t) @ R  # R^T (pc - t)
    lo, hi = (0.5, 4.0) if profile == "synthetic" else (0.4, 1.2)
    sig = torch.exp(math.log(lo) + (math.log(hi) - math.log(lo)) * torch.rand(N, 3, generator=gen))
    scale = sig * (z.abs() / fx).unsqueeze(1)
    q = torch.randn(N, 4, generator=gen)
    rotate = q / q.norm(dim=1, keepdim=True)
    opacity = torch.sigmoid(10.0 * (torch.rand(N, 1, generator=gen) - 0.5) * 0.6)
    rgb = torch.rand(N, 3, generator=gen)
    return Scene(xyz.to(torch.float32).contiguous(), scale.to(torch.float32).contiguous(),
                 rotate.to(torch.float32).contiguous(), opacity.to(torch.float32).contiguous(),
                 rgb.to(torch.float32).contiguous(), intr, extr.contiguous(), W, H, bg)


def make_grad_image(C: int, W: int, H: int, seed: int = 1) -> torch.Tensor:
    """Fixed random dL/d(out) of shape (C,H,W): loss = sum(out * G)."""
    gen = torch.Generator().manual_seed(seed)
    return torch.randn(C, H, W, generator=gen).to(torch.float32)
