"""Deterministic synthetic weights and inputs.

Everything is derived from integer PCG64 streams and exact float arithmetic so
that the container that generates ``tests/golden`` and the GPU box that checks
them produce bit-identical tensors (no dependence on libm / SIMD width).
"""
import zlib

import numpy as np
import torch


def _rng(seed, tag):
    return np.random.Generator(np.random.PCG64([int(seed), zlib.crc32(str(tag).encode())]))


def uniform(shape, seed, tag, lo=-1.0, hi=1.0):
    g = _rng(seed, tag)
    u = g.integers(0, 1 << 24, size=tuple(shape), dtype=np.int64).astype(np.float64) / float(1 << 24)
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32))


def normalish(shape, seed, tag, std=1.0, mean=0.0):
    """Irwin-Hall(4) pseudo-normal: exact in fp64, unit variance after scaling."""
    g = _rng(seed, tag)
    u = g.integers(0, 1 << 24, size=(4,) + tuple(shape), dtype=np.int64).astype(np.float64) / float(1 << 24)
    z = (u.sum(0) - 2.0) * np.sqrt(3.0)
    return torch.from_numpy((mean + std * z).astype(np.float32))


def labels(shape_nhw, ncls, seed, tag, block=16):
    """Piecewise-constant label map (N,H,W) int64: random labels on a coarse grid, nearest-upsampled."""
    n, h, w = shape_nhw
    g = _rng(seed, tag)
    ch, cw = max(1, (h + block - 1) // block), max(1, (w + block - 1) // block)
    coarse = g.integers(0, ncls, size=(n, ch, cw), dtype=np.int64)
    full = np.repeat(np.repeat(coarse, block, axis=1), block, axis=2)[:, :h, :w]
    return torch.from_numpy(np.ascontiguousarray(full))


def one_hot(lab, ncls):
    n, h, w = lab.shape
    out = torch.zeros(n, ncls, h, w, dtype=torch.float32)
    out.scatter_(1, lab[:, None], 1.0)
    return out


def fill_state_dict(sd, seed):
    """Overwrite every entry of a (reference-shaped) state_dict in place, keyed by name.

    conv weights ~ N(0, 1/fan_in); biases ~ N(0,0.05); BN weight ~ N(1,0.1);
    running_mean ~ N(0,0.1); running_var ~ U(0.5,1.5); noise_scale ~ N(0,0.1);
    spectral-norm u/v = 3 fp64 power iterations from a deterministic start.
    """
    keys = sorted(sd.keys())
    for k in keys:
        t = sd[k]
        leaf = k.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t.zero_()
        elif leaf == "running_mean":
            t.copy_(normalish(t.shape, seed, k, 0.1))
        elif leaf == "running_var":
            t.copy_(uniform(t.shape, seed, k, 0.5, 1.5))
        elif leaf == "noise_scale":
            t.copy_(normalish(t.shape, seed, k, 0.1))
        elif leaf in ("weight", "weight_orig") and t.dim() == 4:
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            t.copy_(normalish(t.shape, seed, k, float(np.sqrt(1.0 / fan_in))))
        elif leaf == "weight" and t.dim() == 1:
            t.copy_(normalish(t.shape, seed, k, 0.1, 1.0))
        elif leaf == "bias":
            t.copy_(normalish(t.shape, seed, k, 0.05))
        elif leaf in ("weight_u", "weight_v"):
            pass  # below, needs weight_orig
        else:
            raise KeyError("synth.fill_state_dict: unhandled key %s" % k)
    for k in keys:
        if k.endswith(".weight_u"):
            base = k[: -len("weight_u")]
            w = sd[base + "weight_orig"].detach().double().numpy()
            w2 = w.reshape(w.shape[0], -1)
            u = normalish((w2.shape[0],), seed, k).double().numpy()
            u /= np.linalg.norm(u) + 1e-12
            for _ in range(3):
                v = w2.T @ u
                v /= np.linalg.norm(v) + 1e-12
                u = w2 @ v
                u /= np.linalg.norm(u) + 1e-12
            sd[k].copy_(torch.from_numpy(u.astype(np.float32)))
            sd[base + "weight_v"].copy_(torch.from_numpy(v.astype(np.float32)))
    return sd


def tocg_inputs(n, h, w, seed):
    """input1 = cloth(3)+mask(1); input2 = one-hot parse(13)+densepose(3)  (SURVEY §8d)."""
    # a smooth "garment": low-resolution random field, bilinearly enlarged (exact fp32 lerps), plus faint texture —
    # white-noise cloth would turn a 0.03-pixel flow difference into an O(0.1) colour difference
    import torch.nn.functional as F
    coarse = uniform((n, 3, max(2, h // 16), max(2, w // 16)), seed, "cloth")
    cloth = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True) * 0.9 + uniform((n, 3, h, w), seed, "cloth_tex") * 0.05
    mask = (labels((n, h, w), 2, seed, "cmask", block=32)[:, None]).float()
    parse = one_hot(labels((n, h, w), 13, seed, "parse", block=16), 13)
    dense = uniform((n, 3, h, w), seed, "dense")
    return torch.cat([cloth, mask], 1), torch.cat([parse, dense], 1)


def gen_inputs(n, h, w, seed, input_nc=9, seg_nc=7):
    x = uniform((n, input_nc, h, w), seed, "gx")
    seg = one_hot(labels((n, h, w), seg_nc, seed, "gseg", block=16), seg_nc)
    return x, seg


def spade_noise(n, h, w, seed, idx):
    """The idx-th noise draw of a forward, already in (N,H,W) layout.

    Reference: ``torch.randn(b, w, h, 1)`` then ``.transpose(1, 3)`` (network_generator.py:104-107),
    i.e. noise[b, 0, y, x] = draw[b, x, y, 0].
    """
    draw = normalish((n, w, h, 1), seed, "noise%d" % idx)
    return draw.transpose(1, 3)[:, 0].contiguous()
