"""Golden vectors produced by the UNMODIFIED reference training loop (run from the repo root, where /root/reference
is mounted):

    python tests/golden/make_trainer_golden.py

/root/reference/gflow/trainer.py is imported as it is (tests/ref_harness.py: msplat computed by the CPU oracle, absent
packages shimmed) and SimpleGaussian.train is run for the three kinds of stage GFlow has -- first frame, later-frame
camera-only, later-frame full.  For every iteration the raw state the reference holds when it renders (attributes,
pose, depth_a / depth_b) and the losses it then posts are recorded, plus everything the stage was given (targets,
masks, previous-frame state).  The file travels to the GPU box, where /root/reference does not exist: there the native
loop (csrc/fit.cu) is compared with what the reference itself produced.
PARITY of the rasteriser underneath stays unpinned (the reference's msplat is not available; the oracle computes it)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_harness  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trainer_stages.npz")
W, H, N = 48, 32, 200
ATTRS = ("xyz", "scale", "rotate", "opacity", "rgb")


def record_stage(ref_trainer, t, store, name, **train_kw):
    """Runs t.train(**train_kw) and stores, per iteration, the state at render time and the posted losses."""
    states = []
    orig = ref_trainer.render.render_multiple

    def spy(input_group, keys):
        if len(keys) > 1:  # the full render of the iteration (the moving-subset render asks for ["rgb"] only)
            states.append({**{k: t._attributes[k].detach().clone() for k in ATTRS}, "pose": t.pose.detach().clone(),
                           "ab": torch.cat([t.depth_a.detach(), t.depth_b.detach()])})
        return orig(input_group, keys)

    ref_trainer.render.render_multiple = spy
    ref_harness.Bar.posted = []
    try:
        t.train(**train_kw)
    finally:
        ref_trainer.render.render_multiple = orig
    iters = train_kw["iterations"]
    states = states[:iters]
    posted = ref_harness.Bar.posted
    for k in ATTRS + ("pose", "ab"):
        store[f"{name}/state/{k}"] = torch.stack([s[k] for s in states]).numpy()
    for k in ATTRS:  # .clone(): later stages update these tensors in place
        store[f"{name}/final/{k}"] = t._attributes[k].detach().clone().numpy()
    store[f"{name}/final/pose"] = t.pose.detach().clone().numpy()
    store[f"{name}/final/ab"] = torch.cat([t.depth_a.detach(), t.depth_b.detach()]).numpy()
    for key in ("total", "rgb", "depth", "var", "scale", "still", "flow"):
        store[f"{name}/posted/{key}"] = np.array([float(p.get(key, "nan")) for p in posted], dtype=np.float64)


def main():
    store = {}
    with ref_harness.reference_trainer("/tmp/gflow_b200_golden") as (ref_trainer, calls):
        t, img0, depth0 = ref_harness.new_trainer(ref_trainer, "/tmp/gflow_b200_golden", W, H, N)
        g = torch.Generator().manual_seed(5)
        mm0 = torch.zeros(H, W, dtype=torch.bool)
        mm0[10:20, 5:25] = True
        mm1 = torch.zeros(H, W, dtype=torch.bool)
        mm1[9:19, 7:27] = True
        img1, depth1 = torch.roll(img0, shifts=1, dims=1).contiguous(), (depth0 * 1.02).contiguous()
        gt_flow = torch.zeros(H, W, 2)
        gt_flow[..., 0] = 1.0 + 0.2 * torch.rand(H, W, generator=g)
        gt_flow[..., 1] = 0.3 * torch.randn(H, W, generator=g)
        store.update(W=W, H=H, N=N, intr=t.intr.clone().numpy(), img0=img0.numpy(), depth0=depth0.numpy(), img1=img1.numpy(),
                     depth1=depth1.numpy(), gt_flow=gt_flow.numpy(), move_mask0=mm0.numpy(), move_mask1=mm1.numpy())
        lam = dict(lambda_rgb=1.0, lambda_depth=0.1, lambda_var=0.2, lambda_scale=0.05)
        store["first/hyper"] = np.array([4, 4e-3, 1e-3, 1.0, 0.1, 0.2, 0.05, 0.0, 0.0])  # iters lr lr_cam rgb depth var scale still flow
        record_stage(ref_trainer, t, store, "first", iterations=4, lr=4e-3, lr_camera=1e-3, move_mask=mm0, densify_interval=500,
                     densify_times=0, **lam)

        def prev_state(tag):
            store[f"{tag}/still_mask"] = t.still_mask.clone().numpy()
            store[f"{tag}/tentative"] = t.still_mask_tentative.clone().numpy()
            store[f"{tag}/last_still_mask"] = t.last_still_mask.clone().numpy()
            store[f"{tag}/last_uv"] = t.last_uv.clone().numpy()
            store[f"{tag}/last_xyz"] = t.last_xyz.clone().numpy()

        t.set_gt_image(img1)
        t.set_gt_depth(depth1)
        t.set_gt_flow(gt_flow)
        prev_state("camera")
        store["camera/hyper"] = np.array([3, 1e-2, 2e-3, 1.0, 0.1, 0.0, 0.0, 0.0, 0.01])
        record_stage(ref_trainer, t, store, "camera", iterations=3, lr_camera=2e-3, camera_only=True, move_mask=mm1, lambda_rgb=1.0,
                     lambda_depth=0.1, lambda_var=0.0, lambda_still=0.0, lambda_flow=0.01, densify_interval=500, densify_times=0)
        prev_state("all")
        store["all/hyper"] = np.array([4, 2e-3, 0.0, 1.0, 0.1, 0.2, 0.05, 0.3, 0.01])
        record_stage(ref_trainer, t, store, "all", iterations=4, lr=2e-3, lr_camera=0.0, mask=torch.zeros(H, W, 1), move_mask=mm1,
                     lambda_still=0.3, lambda_flow=0.01, densify_interval=500, densify_times=0, **lam)
    np.savez_compressed(OUT, **store)
    print(f"wrote {OUT}: {os.path.getsize(OUT) / 1024:.0f} KiB, {len(store)} arrays")


if __name__ == "__main__":
    if not ref_harness.available():
        raise SystemExit("needs /root/reference (not available on the GPU box)")
    main()
