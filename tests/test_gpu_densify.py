"""GPU parity of the device-side densification (gflow_b200/csrc/densify.cu through gflow_b200.densify.Densifier)
against the numpy / torch restatement of /root/reference/gflow/trainer.py:878-939 in oracle/fit_ref.py -- the same
checks tests/test_simt_densify.py makes under CPU emulation, on the real kernels: the masked weights and the mask
count equal the oracle's, the draw follows the oracle's inverse-CDF distribution over the same error map, and every
attribute derived from the drawn pixels equals the reference's arithmetic on those pixels."""
import numpy as np
import pytest
import torch

from gflow_b200.synthetic import make_camera
from oracle import fit_ref as FR

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _weights_view(dens, W, H):
    P = W * H
    stats = dens.ws[:32].view(torch.int32).cpu()
    total = float(dens.ws[12:16].view(torch.float32).cpu()[0])
    weights = dens.ws[256:256 + 4 * P].view(torch.float32).reshape(H, W).cpu()
    return int(stats[1]), total, weights


@pytest.mark.parametrize("use_mask", [False, True])
def test_weights_and_mask_count_equal_the_oracle(use_mask):
    from gflow_b200 import capi, ops
    from gflow_b200.densify import Densifier

    W, H = 854, 480
    g = torch.Generator().manual_seed(0)
    err = torch.rand(H, W, generator=g) ** 4 * 0.01
    err[torch.rand(H, W, generator=g) < 0.2] = 0.0
    mask = (torch.rand(H, W, generator=g) > 0.6) if use_mask else None
    dens = Densifier(W, H, DEV)
    m8 = None if mask is None else mask.to(torch.uint8).to(DEV)
    e = err.to(DEV)
    capi.check(dens.lib.gfb_densify_prepare(e.data_ptr(), ops._ptr(m8), W, H, 1e-3, dens.ws.data_ptr(), dens._stream()), "prepare")
    torch.cuda.synchronize()
    count, total, weights = _weights_view(dens, W, H)
    w_o, m_o, ratio = FR.densify_weights(err.numpy(), 1e-3, None if mask is None else mask.numpy())
    assert count == int(m_o.sum())
    assert np.array_equal(weights.numpy(), w_o.astype(np.float32)), "per-pixel weights are bit-identical to the oracle's"
    assert abs(total - float(w_o.sum())) <= 1e-5 * float(w_o.sum())


def test_draw_follows_the_oracle_distribution_and_attributes_match():
    from gflow_b200.densify import Densifier

    W, H, num_points = 320, 200, 60000
    g = torch.Generator().manual_seed(2)
    err = torch.zeros(H, W)
    err[25:75, 50:150] = 0.004      # 5 000 px
    err[125:175, 200:300] = 0.0011  # 5 000 px, just above the threshold
    err[100, 25] = 0.05             # one hot pixel
    err += (torch.rand(H, W, generator=g) < 0.01).float() * 0.002  # scattered single pixels
    gt_image = torch.rand(H, W, 3, generator=g)
    gt_depth = 1.0 + 3.0 * torch.rand(H, W, 1, generator=g)
    intr = torch.tensor([160.0, 150.0, 159.5, 99.5])
    _, extr = make_camera(W, H, g)
    dens = Densifier(W, H, DEV)
    new = dens.sample(err.to(DEV), gt_image.to(DEV), gt_depth.to(DEV), intr.to(DEV), extr.to(DEV), num_points, 1e-3, 1.0, seed=1234)
    w_o, m_o, ratio = FR.densify_weights(err.numpy(), 1e-3, None)
    n = FR.densify_count(num_points, ratio, 1.0)
    assert new is not None and new["xyz"].shape[0] == n and n > 5000, (n,)
    s = new["pixels"].cpu().long()
    wflat = torch.from_numpy(w_o.reshape(-1))
    assert int(s.min()) >= 0 and int(s.max()) < W * H and bool((wflat[s] > 0).all()), "only weighted pixels are drawn"
    # region frequencies against the oracle's p = w / sum(w) (what np.random.choice(p=...) samples from), 4 sigma
    prob = wflat / wflat.sum()
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    regions = {"strong": (yy >= 25) & (yy < 75) & (xx >= 50) & (xx < 150), "weak": (yy >= 125) & (yy < 175) & (xx >= 200) & (xx < 300),
               "hot": (yy == 100) & (xx == 25)}
    regions["scattered"] = ~(regions["strong"] | regions["weak"] | regions["hot"])
    for name, m in regions.items():
        pr = float(prob[m.reshape(-1)].sum())
        got = float(m.reshape(-1)[s].float().mean())
        assert abs(got - pr) <= 4 * (pr * (1 - pr) / n) ** 0.5 + 1e-4, (name, got, pr)
    # inverse-CDF check: the empirical CDF over the pixel order equals the oracle's cumulative weights (KS statistic)
    cdf = torch.cumsum(prob.double(), 0)
    emp = torch.cumsum(torch.bincount(s, minlength=W * H).double() / n, 0)
    assert float((emp - cdf).abs().max()) <= 1.63 / n ** 0.5, "Kolmogorov-Smirnov at the 1 % level"
    # same seed -> same draw, other seed -> another draw
    again = dens.sample(err.to(DEV), gt_image.to(DEV), gt_depth.to(DEV), intr.to(DEV), extr.to(DEV), num_points, 1e-3, 1.0, seed=1234)
    other = dens.sample(err.to(DEV), gt_image.to(DEV), gt_depth.to(DEV), intr.to(DEV), extr.to(DEV), num_points, 1e-3, 1.0, seed=99)
    assert torch.equal(again["pixels"], new["pixels"]) and not torch.equal(other["pixels"], new["pixels"])
    # attributes of the drawn pixels = the reference's arithmetic (trainer.py:905-939) on those pixels
    ref = FR.densify_attributes(new["pixels"].cpu(), gt_image, gt_depth.squeeze(-1), intr, extr, num_points, W)
    assert torch.allclose(new["xyz"].cpu(), ref["xyz"], rtol=1e-5, atol=1e-5)
    assert torch.allclose(new["scale"].cpu(), ref["scale"], rtol=1e-6)
    assert torch.equal(new["rotate"].cpu(), ref["rotate"])
    assert torch.allclose(new["opacity"].cpu(), ref["opacity"], rtol=1e-6)
    fin = torch.isfinite(ref["rgb"])
    assert torch.equal(torch.isfinite(new["rgb"].cpu()), fin) and torch.allclose(new["rgb"].cpu()[fin], ref["rgb"][fin], rtol=1e-5, atol=1e-6)
