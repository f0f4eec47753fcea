"""Frame-state files in GFlow's checkpoint layout (gflow_b200/checkpoint.py) and their wire form.

Where /root/reference is mounted, the reference's OWN save_checkpoint / load_checkpoint method bodies
(/root/reference/gflow/trainer.py:252-288) are compiled from where they lie (ast, nothing copied) and run
against our files and vice versa; trainer.py as a whole cannot be imported here (msplat, roma, imageio are
absent)."""
import ast
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gflow_b200 import checkpoint as ck
from gflow_b200 import fit, frames

TRAINER_PY = "/root/reference/gflow/trainer.py"


def _state(N=23, seed=0, with_extras=True):
    g = torch.Generator().manual_seed(seed)
    attrs = {k: torch.randn(N, w, generator=g) for k, w in frames.STATE_KEYS}
    q = torch.randn(4, generator=g)
    pose = torch.cat([q / q.norm(), torch.randn(3, generator=g)])
    if pose[3] < 0:
        pose[:4] = -pose[:4]
    return ck.FrameState(attributes=attrs, intr=torch.tensor([32.0, 24.0, 32.0, 24.0]), extr=fit.pose_to_extr(pose), width=64,
                         height=48, still_mask=(torch.rand(N, generator=g) > 0.5) if with_extras else None,
                         move_seg=(np.arange(48 * 64).reshape(48, 64) % 255).astype(np.uint8) if with_extras else None,
                         last_uv=torch.rand(N, 2, generator=g) * 60 if with_extras else None), pose


def _same(a: ck.FrameState, b: ck.FrameState, wire=False):
    assert all(torch.equal(a.attributes[k], b.attributes[k]) for k in ck.ATTRS)
    assert torch.equal(a.intr, b.intr) and torch.equal(a.extr, b.extr) and (a.width, a.height) == (b.width, b.height)
    for x, y in ((a.still_mask, b.still_mask), (a.last_uv, b.last_uv)):
        assert (x is None) == (y is None) and (x is None or torch.equal(x, y))
    if not wire:
        assert (a.move_seg is None) == (b.move_seg is None) and (a.move_seg is None or np.array_equal(a.move_seg, b.move_seg))


@pytest.mark.parametrize("with_extras", [True, False])
def test_file_roundtrip_and_pose(tmp_path, with_extras):
    st, pose = _state(with_extras=with_extras)
    path = ck.save_frame(str(tmp_path / "ckpt" / "0003.tar"), st)
    back = ck.load_frame(path)
    _same(st, back)
    assert torch.allclose(back.pose(), pose, atol=1e-5)  # load_camera(extr=...) semantics: extr -> xyzw quaternion + t
    raw = torch.load(path, weights_only=False)
    assert sorted(raw) == ["attributes", "extr", "height", "intr", "last_uv", "move_seg", "still_mask", "width"]
    assert all(isinstance(raw["attributes"][k], torch.nn.Parameter) for k in ck.ATTRS)


def test_bad_inputs(tmp_path):
    st, _ = _state()
    st.attributes["rgb"] = st.attributes["rgb"][:-1]
    with pytest.raises(ValueError):
        ck.save_frame(str(tmp_path / "x.tar"), st)
    torch.save({"something": 1}, str(tmp_path / "y.tar"))
    with pytest.raises(ValueError, match="not a GFlow checkpoint"):
        ck.load_frame(str(tmp_path / "y.tar"))
    with pytest.raises(ValueError):
        ck.from_wire(torch.zeros(40))


@pytest.mark.parametrize("with_extras", [True, False])
def test_wire_roundtrip(with_extras):
    st, _ = _state(N=101, seed=3, with_extras=with_extras)
    buf = ck.to_wire(st)
    assert buf.dtype == torch.float32 and buf.numel() == ck.wire_numel(101, with_extras, with_extras)
    _same(st, ck.from_wire(buf), wire=True)
    with pytest.raises(ValueError):
        ck.from_wire(buf[:-1])


# ----------------------------------------------------------------------------- against the reference's own methods
def _reference_methods():
    tree = ast.parse(open(TRAINER_PY).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SimpleGaussian")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("save_checkpoint", "load_checkpoint")]
    assert len(fns) == 2
    mod = ast.Module(body=fns, type_ignores=[])
    torch_proxy = types.SimpleNamespace(save=torch.save, load=lambda p: torch.load(p, weights_only=False),
                                        cuda=types.SimpleNamespace(empty_cache=lambda: None))
    ns = {"torch": torch_proxy, "os": os}
    exec(compile(mod, TRAINER_PY, "exec"), ns)
    return ns["save_checkpoint"], ns["load_checkpoint"]


@pytest.mark.skipif(not os.path.exists(TRAINER_PY), reason="reference sources not mounted")
def test_reference_reads_our_file_and_we_read_the_reference_file(tmp_path):
    ref_save, ref_load = _reference_methods()
    st, pose = _state(N=17, seed=5)
    # ours -> reference load_checkpoint
    path = ck.save_frame(str(tmp_path / "a.tar"), st)
    seen = {}
    obj = types.SimpleNamespace(load_camera=lambda extr=None, show=True: seen.update(extr=extr))
    ref_load(obj, path, show=False)
    assert all(torch.equal(obj._attributes[k], st.attributes[k]) for k in ck.ATTRS)
    assert torch.equal(obj.intr, st.intr) and torch.equal(seen["extr"], st.extr)
    assert torch.equal(obj.still_mask, st.still_mask) and torch.equal(obj.last_uv, st.last_uv)
    assert np.array_equal(obj.move_seg, st.move_seg)
    # reference save_checkpoint -> ours
    obj2 = types.SimpleNamespace(_attributes={k: torch.nn.Parameter(v.clone()) for k, v in st.attributes.items()}, intr=st.intr,
                                 get_extr=lambda: st.extr, still_mask=st.still_mask, move_seg=st.move_seg, last_uv=st.last_uv,
                                 W=st.width, H=st.height, dir=str(tmp_path))
    ref_save(obj2, ckpt_name="0007")
    back = ck.load_frame(os.path.join(str(tmp_path), "ckpt", "0007.tar"))
    _same(st, back)
    assert torch.allclose(back.pose(), pose, atol=1e-5)


# ----------------------------------------------------------------------------- one packed broadcast (gloo, world 2)
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st, _ = _state(N=57, seed=9)
        got = frames.broadcast_frame_state(st if rank == 0 else None, src=0, device=torch.device("cpu"))
        _same(st, got, wire=True)
        q.put((rank, True))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_broadcast_frame_state_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)], res
