"""Per-frame Adam loop (BASELINE config 3) and its helpers."""
import math

import pytest
import torch

from gflow_b200 import fit
from gflow_b200.synthetic import make_scene


def test_pose_extr_roundtrip_and_identity():
    ident = torch.tensor([0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0])
    E = fit.pose_to_extr(ident)
    assert torch.allclose(E, torch.cat([torch.eye(3), torch.zeros(3, 1)], dim=1))
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        q = torch.randn(4, generator=g)
        q = q / q.norm()
        if q[3] < 0:
            q = -q
        pose = torch.cat([q, torch.randn(3, generator=g)])
        E = fit.pose_to_extr(pose)
        assert torch.allclose(E[:, :3] @ E[:, :3].T, torch.eye(3), atol=1e-5)
        assert abs(float(torch.det(E[:, :3])) - 1.0) < 1e-5
        assert torch.allclose(fit.extr_to_pose(E), pose, atol=1e-5)


def test_activations_have_inverses():
    g = torch.Generator().manual_seed(1)
    x = torch.rand(10, 3, generator=g) * 0.9 + 0.05
    assert torch.allclose(fit.activate("rgb", fit.inverse_activate("rgb", x)), x, atol=1e-6)
    o = torch.rand(10, 1, generator=g) * 0.9 + 0.05
    assert torch.allclose(fit.activate("opacity", fit.inverse_activate("opacity", o)), o, atol=1e-6)
    q = torch.randn(10, 4, generator=g)
    assert torch.allclose(fit.activate("rotate", q).norm(dim=1), torch.ones(10), atol=1e-6)
    assert torch.all(fit.activate("scale", -x) == x)


def _raw_state(sc):
    return {"xyz": sc.xyz, "scale": sc.scale, "rotate": sc.rotate,
            "opacity": fit.inverse_activate("opacity", sc.opacity.clamp(0.02, 0.98)),
            "rgb": fit.inverse_activate("rgb", sc.rgb.clamp(0.02, 0.98))}


@pytest.mark.gpu
def test_fit_recovers_colour_and_pose_gradient_flows():
    dev = torch.device("cuda:0")
    sc = make_scene(8000, 320, 200, seed=5, profile="synthetic")
    pose = fit.extr_to_pose(sc.extr)
    raw = {k: v.to(dev) for k, v in _raw_state(sc).items()}
    target = fit.FrameFitter(raw, sc.intr.to(dev), pose.to(dev), sc.W, sc.H)
    with torch.no_grad():
        gt_img, gt_depth, _ = target.render(0.0, want_depth=True)
    gt_image = gt_img.permute(1, 2, 0).contiguous()
    gt_d = gt_depth.permute(1, 2, 0).contiguous()
    # start from grey colours: the loop must pull the loss down
    start = dict(raw)
    start["rgb"] = torch.zeros_like(raw["rgb"])
    for fused in (False, True):
        f = fit.FrameFitter(start, sc.intr.to(dev), pose.to(dev), sc.W, sc.H)
        cfg = fit.FitConfig(iterations=60, lr=1e-2, lambda_depth=0.0 if fused else 0.1, fused=fused)
        res = f.train(gt_image, None if fused else gt_d, cfg)
        assert len(res.losses) == 60 and all(math.isfinite(v) for v in res.losses)
        assert res.losses[-1] < 0.8 * res.losses[0], (fused, res.losses[0], res.losses[-1])
        assert res.image.shape == (3, sc.H, sc.W)
    # camera-only stage: attributes frozen, pose moves (trainer.py:548-551)
    shifted = pose.clone()
    shifted[4] += 0.02
    f = fit.FrameFitter(raw, sc.intr.to(dev), shifted.to(dev), sc.W, sc.H)
    before = {k: v.detach().clone() for k, v in f.attrs.items()}
    res = f.train(gt_image, gt_d, fit.FitConfig(iterations=25, lr=4e-3, lr_camera=1e-3, camera_only=True))
    assert all(torch.equal(before[k], f.attrs[k].detach()) for k in before)
    assert abs(float(res.pose[4] - pose[4].to(dev))) < 0.02  # moved back towards the true pose
    assert res.losses[-1] < res.losses[0]


def test_accelerate_import_hook_patches_the_trainer_module(tmp_path, monkeypatch):
    """GFLOW_B200_NATIVE_TRAIN=1: the drop-in msplat module installs a post-import hook that swaps
    SimpleGaussian.train for the native adapter as soon as GFlow's `trainer` module has been executed."""
    import importlib
    import sys

    from gflow_b200 import accelerate

    (tmp_path / "fake_gflow_trainer.py").write_text(
        "class SimpleGaussian:\n    def train(self, iterations=1):\n        return 'reference loop'\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    before = list(sys.meta_path)
    try:
        accelerate.install_import_hook("fake_gflow_trainer")
        mod = importlib.import_module("fake_gflow_trainer")
        assert mod.SimpleGaussian.train is accelerate.native_train
        assert mod.SimpleGaussian.train_reference(mod.SimpleGaussian()) == "reference loop"
        import json  # unrelated imports are untouched

        assert json.loads("1") == 1
    finally:
        sys.meta_path[:] = before
        sys.modules.pop("fake_gflow_trainer", None)


def test_zero_edit_switch_works_in_the_real_import_order(tmp_path):
    """The order the reference's entry point really imports in (fit_video.py:3-5, trainer.py:7): `import utils`, then
    `from trainer import SimpleGaussian`, whose module imports `msplat` at its top -- i.e. the drop-in (and its hook)
    arrives while `trainer` is half executed -- then `from utils.traj_visualizer import ...`.  Round 1's hook never
    fired in that order (the switch was a silent no-op).  Run in a process of its own with GFLOW_B200_NATIVE_TRAIN=1;
    the CUDA-free stand-in for gflow_b200.ops keeps this a pure import-machinery test."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    (tmp_path / "utils").mkdir()
    (tmp_path / "utils" / "__init__.py").write_text("")
    (tmp_path / "utils" / "traj_visualizer.py").write_text("class TrajVisualizer:\n    pass\n")
    (tmp_path / "utils" / "render.py").write_text("import msplat\n")
    (tmp_path / "trainer.py").write_text(
        "import math\nimport msplat\nimport utils\nimport utils.render as render\n"
        "class SimpleGaussian:\n    def train(self, iterations=1):\n        return 'reference loop'\n")
    (tmp_path / "fit_video.py").write_text(
        "import utils\nfrom trainer import SimpleGaussian\n"
        "after_trainer = SimpleGaussian.train.__name__\n"
        "from utils.traj_visualizer import TrajVisualizer\n"
        "print('AFTER_TRAINER', after_trainer)\nprint('AFTER_NEXT_IMPORT', SimpleGaussian.train.__name__)\n"
        "import msplat\nprint('OPS', msplat.project_point.__name__)\n")
    (tmp_path / "late.py").write_text(  # no import at all follows trainer: the first operator call patches
        "from trainer import SimpleGaussian\nimport sys\nmsplat = sys.modules['msplat']\n"
        "print('BEFORE_CALL', SimpleGaussian.train.__name__)\n"
        "try:\n    msplat.project_point()\nexcept TypeError:\n    pass\n"
        "print('AFTER_CALL', SimpleGaussian.train.__name__, msplat.project_point.__name__)\n")
    env = dict(os.environ, GFLOW_B200_NATIVE_TRAIN="1", GFLOW_B200_NO_BUILD="1",
               PYTHONPATH=os.pathsep.join([str(tmp_path), os.path.join(root, "gflow_b200", "dropin"), root]))
    res = subprocess.run([sys.executable, str(tmp_path / "fit_video.py")], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "AFTER_NEXT_IMPORT native_train" in res.stdout, res.stdout
    assert "OPS project_point" in res.stdout
    res = subprocess.run([sys.executable, str(tmp_path / "late.py")], capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "AFTER_CALL native_train project_point" in res.stdout, res.stdout


def test_extr_to_pose_round_trips_every_rotation_including_180_degrees():
    """Shepperd's method with all four branches: round 1 implemented the w-dominant branch only and returned the
    identity for 180-degree rotations (axis flips such as an OpenGL <-> OpenCV extrinsic)."""
    import math

    g = torch.Generator().manual_seed(0)
    quats = [torch.tensor(q) for q in ([1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0.70710678, 0.70710678, 0, 0],
                                       [0, 0.70710678, 0.70710678, 0], [0, 0, 0, 1.0], [0.5, 0.5, 0.5, 0.5])]
    for _ in range(500):
        q = torch.randn(4, generator=g)
        quats.append(q / q.norm())
    for _ in range(200):  # within 1e-4 rad of 180 degrees about a random axis
        ax = torch.randn(3, generator=g)
        ax = ax / ax.norm()
        ang = math.pi - 1e-4 * float(torch.rand(1, generator=g))
        quats.append(torch.cat([ax * math.sin(ang / 2), torch.tensor([math.cos(ang / 2)])]))
    for q in quats:
        E = fit.pose_to_extr(torch.cat([q, torch.tensor([0.1, -0.2, 0.3])]))
        pose = fit.extr_to_pose(E)
        assert float(pose[3]) >= 0.0 and abs(float(pose[:4].norm()) - 1.0) < 1e-5
        assert float((fit.pose_to_extr(pose) - E).abs().max()) < 2e-6, q
    flip = torch.tensor([[1.0, 0, 0, 0.5], [0, -1.0, 0, 0.0], [0, 0, -1.0, 2.0]])  # OpenGL <-> OpenCV axis flip
    assert float((fit.pose_to_extr(fit.extr_to_pose(flip)) - flip).abs().max()) < 1e-6
