"""CPU: host-side logic of the drop-in modules that needs no GPU — loss modules, grid, checkpoint io, init, weight packing,
tile pickers (the numerics of the kernels themselves are covered by the -m gpu tests)."""
import io
import contextlib
import os
import types

import numpy as np
import pytest
import torch

from helpers import gen_opt, tocg_opt
from hrviton_b200 import ops


def test_make_grid_matches_reference_formula():
    import networks
    g4 = networks.make_grid(2, 5, 7, types.SimpleNamespace(cuda=False))
    g3 = networks.make_grid(2, 5, 7)  # stale 3-argument form (train_condition.py:241)
    assert g4.shape == (2, 5, 7, 2)
    assert torch.equal(g4[0, 0, :, 0], torch.linspace(-1, 1, 7)) and torch.equal(g4[0, :, 0, 1], torch.linspace(-1, 1, 5))
    assert torch.equal(g3.cpu(), g4)


def test_hinge_ganloss_matches_reference_semantics():
    import network_generator
    crit = network_generator.GANLoss("hinge")
    p = [[torch.tensor([[0.5, -2.0]]), torch.tensor([[1.5, 0.2]])], [torch.tensor([[3.0]])]]  # 2 scales, last entry = logits
    # discriminator, real: -mean(min(x-1,0)) per scale, averaged over scales (network_generator.py:369-398)
    want = (-(torch.clamp(p[0][-1] - 1, max=0)).mean() + -(torch.clamp(p[1][-1] - 1, max=0)).mean()) / 2
    assert torch.allclose(crit(p, True, for_discriminator=True), want)
    want_f = (-(torch.clamp(-p[0][-1] - 1, max=0)).mean() + -(torch.clamp(-p[1][-1] - 1, max=0)).mean()) / 2
    assert torch.allclose(crit(p, False, for_discriminator=True), want_f)
    assert torch.allclose(crit(p, True, for_discriminator=False), (-(p[0][-1].mean()) - p[1][-1].mean()) / 2)
    with pytest.raises(ValueError):
        network_generator.GANLoss("bogus")


def test_lsgan_ganloss():
    import networks
    crit = networks.GANLoss(use_lsgan=True)
    pred = [[torch.tensor([[0.5, 2.0]])], [torch.tensor([[1.0]])]]
    assert torch.allclose(crit(pred, True), ((pred[0][-1] - 1) ** 2).mean() + ((pred[1][-1] - 1) ** 2).mean())
    assert torch.allclose(crit(pred, False), (pred[0][-1] ** 2).mean() + (pred[1][-1] ** 2).mean())


def test_checkpoint_roundtrip(tmp_path):
    import networks
    opt = tocg_opt(False)
    with contextlib.redirect_stdout(io.StringIO()):
        a = networks.ConditionGenerator(opt, 4, 16, 13, ngf=8, norm_layer=torch.nn.BatchNorm2d)
        b = networks.ConditionGenerator(opt, 4, 16, 13, ngf=8, norm_layer=torch.nn.BatchNorm2d)
    path = os.path.join(tmp_path, "sub", "tocg.pth")
    networks.save_checkpoint(a, path, opt)  # creates the directory like the reference (networks.py:411-417)
    networks.load_checkpoint(b, path, opt)
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k


def test_init_weights_and_print_network(capsys):
    import network_generator
    d = network_generator.MultiscaleDiscriminator(gen_opt(256, 256))
    d.print_network()
    assert "MultiscaleDiscriminator" in capsys.readouterr().out
    d.init_weights("xavier", 0.02)
    w = d.discriminator_0.model0[0].weight
    assert float(w.std()) < 0.02  # xavier_normal with gain 0.02
    with pytest.raises(NotImplementedError):
        d.init_weights("bogus")


def test_tile_pickers():
    assert ops.pick_bk(128) == 64 and ops.pick_bk(80) == 32 and ops.pick_bk(7) == 16 and ops.pick_bk(1040) == 64
    for n in (16, 160, 256, 544, 1056, 2080, 2048):
        bn = ops.pick_bn(n)
        assert bn % 16 == 0 and 16 <= bn <= 256
        assert ops.round_up(n, bn) - n < 0.08 * n + 16  # padding stays small


def test_s2d_weight_equivalence_on_cpu():
    """The stride-2 -> space-to-depth rewrite is exact: conv(x, w, stride 2) == conv(s2d(x), s2d_weight(w), stride 1)."""
    import torch.nn.functional as F
    for k, pad, h, w in [(4, 2, 11, 8), (3, 1, 10, 8), (3, 1, 9, 7), (4, 2, 12, 9)]:
        x = torch.randn(2, 5, h, w)
        wt = torch.randn(6, 5, k, k)
        ref = F.conv2d(x, wt, stride=2, padding=pad)
        c8 = 8
        xp = F.pad(x, (0, w % 2, 0, h % 2, 0, c8 - 5))
        n, c, hh, ww = xp.shape
        s = xp.reshape(n, c, hh // 2, 2, ww // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(n, 4 * c, hh // 2, ww // 2)  # channel (py*2+px)*c8+ci
        got = F.conv2d(s, ops.s2d_weight(wt, pad), stride=1, padding=1)[:, :, :ref.shape[2], :ref.shape[3]]
        assert torch.allclose(got, ref, atol=1e-4), (k, pad, h, w)
