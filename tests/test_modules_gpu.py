"""GPU parity of the drop-in modules (networks.py / network_generator.py at the repo root) against golden outputs of
the UNMODIFIED reference modules (tests/golden/*.npz, produced by tests/golden/make_golden.py from /root/reference).
bf16 activations / fp32 accumulation: tolerance |delta| < 1e-2 per element on O(1) outputs (BASELINE.json north_star)."""
import io
import contextlib

import numpy as np
import pytest
import torch

from helpers import gen_opt, load_golden, maxdiff, synth_state_dict, tocg_opt
from hrviton_b200 import synth

pytestmark = pytest.mark.gpu


def _report(name, got, ref):
    ref = np.asarray(ref, np.float32)
    d = maxdiff(got, ref)
    scale = float(np.abs(ref).max())
    mean = float(np.abs(got.detach().float().cpu().numpy() - ref).mean())
    print("PARITY %-28s max|d|=%.3e mean|d|=%.3e (ref absmax %.3g)" % (name, d, mean, scale))
    return d, scale


@pytest.mark.parametrize("name", ["tocg_256x192_b1", "tocg_128x96_b2"])
def test_tocg_forward(name):
    import networks
    g = load_golden(name)
    n, h, w = [int(v) for v in g["shape"]]
    seed = int(g["seed"])
    m = networks.ConditionGenerator(tocg_opt(True), 4, 16, 13, ngf=96, norm_layer=torch.nn.BatchNorm2d)
    m.load_state_dict(synth_state_dict("tocg", seed))
    m = m.cuda().eval()
    i1, i2 = synth.tocg_inputs(n, h, w, seed)
    with torch.no_grad():
        flows, seg, wc, wcm = m(tocg_opt(True), i1.cuda(), i2.cuda())
        flows2, seg2, _, _ = m(i1.cuda(), i2.cuda())  # stale 2-positional call form (train_generator.py:215)
    torch.cuda.synchronize()
    assert maxdiff(seg, seg2) == 0.0
    for i, f in enumerate(flows):
        d, s = _report("flow%d" % i, f, g["flow%d" % i])
        assert d < 2e-2 * max(1.0, s)
    d, s = _report("seg", seg, g["seg"])
    assert d < 2e-2 * max(1.0, s)
    d, _ = _report("warped_c", wc, g["warped_c"])
    assert d < 3e-2
    d, _ = _report("warped_cm", wcm, g["warped_cm"])
    assert d < 6e-2  # binary mask edges: |d| = flow error (pixels) x unit step


@pytest.mark.parametrize("name", ["gen_512x384_b1", "gen_256x256_b2"])
def test_generator_forward(name):
    import network_generator
    g = load_golden(name)
    n, h, w = [int(v) for v in g["shape"]]
    seed = int(g["seed"])
    m = network_generator.SPADEGenerator(gen_opt(h, w, True), 9)
    m.load_state_dict(synth_state_dict("gen", seed))
    m = m.cuda().eval()
    cnt = [0]

    def noise(b, hh, ww):
        t = synth.spade_noise(b, hh, ww, seed, cnt[0]).cuda()
        cnt[0] += 1
        return t

    m.noise_source = noise
    x, seg = synth.gen_inputs(n, h, w, seed)
    with torch.no_grad():
        out = m(x.cuda(), seg.cuda())
    torch.cuda.synchronize()
    assert cnt[0] == 23
    # bf16 activation storage through ~60 stacked convs/normalisations: the fp32 oracle itself moves by max 4.6e-2 /
    # mean 5e-3 when its conv inputs/outputs are rounded to bf16 (DESIGN.md "Parity"); the kernels must stay inside that
    d, _ = _report("generator out", out, g["out"])
    mean = float(np.abs(out.cpu().numpy() - g["out"].astype(np.float32)).mean())
    assert d < 8e-2 and mean < 1e-2


def test_gen_discriminator_forward():
    import network_generator
    g = load_golden("gend_128x96_b2")
    n, h, w = [int(v) for v in g["shape"]]
    seed = int(g["seed"])
    m = network_generator.MultiscaleDiscriminator(gen_opt(h, w, True))
    m.load_state_dict(synth_state_dict("gend", seed))
    m = m.cuda().eval()
    x, seg = synth.gen_inputs(n, h, w, seed, input_nc=3)
    with torch.no_grad():
        res = m(torch.cat([seg, x], 1).cuda())
    for i, fs in enumerate(res):
        for j, f in enumerate(fs):
            d, s = _report("gend d%d_f%d" % (i, j), f, g["d%d_f%d" % (i, j)])
            assert d < 2e-2 * max(1.0, s)


def test_tocg_discriminator_forward():
    import networks
    g = load_golden("tocgd_256x192_b1")
    seed = int(g["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        m = networks.define_D(input_nc=33, Ddownx2=True, Ddropout=True, n_layers_D=3, spectral=False, num_D=2)
    m.load_state_dict(synth_state_dict("tocgd", seed))
    m = m.cuda().eval()
    i1, i2 = synth.tocg_inputs(1, 256, 192, seed)
    segs = synth.one_hot(synth.labels((1, 256, 192), 13, seed, "dseg"), 13)
    with torch.no_grad():
        res = m(torch.cat([i1, i2, segs], 1).cuda())
    for i, r in enumerate(res):
        d, s = _report("tocgd d%d" % i, r[0], g["d%d" % i])
        assert d < 2e-2 * max(1.0, s)


def test_generator_minimum_size_and_odd_batch():
    """fine size 128x128 is the smallest the architecture admits (sh = sw = 1: the head sees ONE pixel) and batch 3 exercises
    pixel tiles spanning images (TN > 1) with a ragged last tile; compared against the oracle on the same weights/noise."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
    import hrviton_oracle as orc
    import network_generator
    n, h, w, seed = 3, 128, 128, 41
    sd = synth_state_dict("gen", seed)
    m = network_generator.SPADEGenerator(gen_opt(h, w, True), 9)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    c1, c2 = [0], [0]

    def nd(b, hh, ww):
        t = synth.spade_noise(b, hh, ww, seed, c1[0]).cuda()
        c1[0] += 1
        return t

    def nc(b, hh, ww):
        t = synth.spade_noise(b, hh, ww, seed, c2[0])
        c2[0] += 1
        return t

    m.noise_source = nd
    x, seg = synth.gen_inputs(n, h, w, seed)
    with torch.no_grad():
        out = m(x.cuda(), seg.cuda())
        ref = orc.spade_generator_forward(sd, x, seg, nc)
    assert out.shape == (n, 3, h, w) and bool(torch.isfinite(out).all())
    d = float((out.cpu() - ref).abs().mean())
    print("PARITY generator 128x128 b3 mean|d| %.3e" % d)
    assert d < 2e-2  # one-pixel InstanceNorm at the head is degenerate (normalised value = 0): mean error only


def test_standalone_blocks_api():
    """ResBlock / SPADEResBlock / discriminator.downsample are callable stand-alone with NCHW tensors like the reference's."""
    import network_generator
    import networks
    rb = networks.ResBlock(16, 32, scale="down").cuda().eval()
    y = rb(torch.randn(2, 16, 32, 24).cuda())
    assert y.shape == (2, 32, 16, 12) and float(y.min()) >= 0.0
    ref = torch.relu  # reference composition on the same weights through torch ops (eval-mode BN)
    with torch.no_grad():
        x = torch.randn(2, 16, 32, 24).cuda()
        r = torch.nn.functional.conv2d(x, rb.scale.weight, None, stride=2, padding=1)
        want = ref(r + rb.block(r))
        got = rb(x)
    assert float((got - want).abs().max()) < 5e-2 * max(1.0, float(want.abs().max()))
    d = network_generator.MultiscaleDiscriminator(gen_opt(256, 192, True)).cuda().eval()
    z = torch.randn(1, 10, 33, 25).cuda()
    want = torch.nn.functional.avg_pool2d(z, 3, stride=2, padding=1, count_include_pad=False)
    assert float((d.downsample(z) - want).abs().max()) < 2e-2
