"""TEST-ONLY stand-ins for Python packages GFlow imports that are absent from this image:
roma, imageio, matplotlib, shapely, concave_hull.  They let /root/reference/gflow/trainer.py be imported and
run UNMODIFIED on the CPU by tests/test_reference_trainer.py; each implements exactly the calls the reference
makes.  Nothing under gflow_b200/ imports this package."""
from __future__ import annotations

import sys
import types

import numpy as np
import torch


# ----------------------------------------------------------------------------- roma
class RigidUnitQuat:
    """roma.RigidUnitQuat(linear=q_xyzw, translation=t): .normalize().to_homogeneous() -> (4,4)
    (used by /root/reference/gflow/trainer.py:115-121,196-201)."""

    def __init__(self, linear, translation):
        self.linear, self.translation = linear, translation

    def normalize(self):
        return RigidUnitQuat(self.linear / torch.linalg.norm(self.linear, dim=-1, keepdim=True), self.translation)

    def to_homogeneous(self):
        x, y, z, w = self.linear.unbind(-1)
        R = torch.stack([
            torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], dim=-1),
            torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], dim=-1),
            torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1),
        ], dim=-2)
        top = torch.cat([R, self.translation.unsqueeze(-1)], dim=-1)
        bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=top.dtype, device=top.device)
        return torch.cat([top, bottom], dim=-2)


def rotmat_to_unitquat(R):
    """xyzw unit quaternion of a rotation matrix (trainer.py:177), via scipy (scalar-last convention)."""
    from scipy.spatial.transform import Rotation

    q = Rotation.from_matrix(R.detach().cpu().double().numpy()).as_quat()
    return torch.as_tensor(q, dtype=R.dtype, device=R.device)


# ----------------------------------------------------------------------------- matplotlib.cm
def _get_cmap(name):
    def cmap(idx):
        v = np.asarray(idx, dtype=np.float64) / 255.0
        return np.stack([v, 1.0 - v, 0.5 + 0.5 * np.sin(6.0 * v), np.ones_like(v)], axis=-1)  # any smooth RGBA map

    return cmap


def install():
    """Registers the stand-ins in sys.modules (only for names that cannot be imported for real).
    Returns the list of names added so the caller can remove them again."""
    added = []

    def add(name, **attrs):
        try:
            __import__(name)
            return
        except ImportError:
            pass
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        added.append(name)
        return m

    add("roma", RigidUnitQuat=RigidUnitQuat, rotmat_to_unitquat=rotmat_to_unitquat)
    noop = lambda *a, **k: None  # noqa: E731
    add("imageio", imwrite=noop, mimwrite=noop, imread=noop, get_writer=noop, mimsave=noop, imsave=noop)
    mpl = add("matplotlib")
    cm = add("matplotlib.cm", get_cmap=_get_cmap)
    plt = add("matplotlib.pyplot")
    if mpl is not None:
        mpl.cm, mpl.pyplot = cm, plt
    geom = add("shapely.geometry", Polygon=object)
    sh = add("shapely", LineString=object, MultiLineString=object, MultiPolygon=object)
    if sh is not None:
        sh.geometry = geom
    add("concave_hull", concave_hull=lambda pts, *a, **k: pts)
    return added
