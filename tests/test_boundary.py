"""CPU: the drop-in boundary — importable names, state_dict keys/shapes identical to the reference's (recorded from
the live reference in tests/golden/state_keys.json), loud failure on CPU tensors, C-ABI exports."""
import contextlib
import ctypes
import io
import json
import os
import re

import pytest
import torch

from helpers import GOLDEN, gen_opt, tocg_opt
from hrviton_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shapes(m):
    return {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()}


def test_state_dict_keys_match_reference():
    import network_generator
    import networks
    ref = json.load(open(os.path.join(GOLDEN, "state_keys.json")))
    with contextlib.redirect_stdout(io.StringIO()):
        mine = {
            "tocg": _shapes(networks.ConditionGenerator(tocg_opt(), 4, 16, 13, ngf=96, norm_layer=torch.nn.BatchNorm2d)),
            "gen": _shapes(network_generator.SPADEGenerator(gen_opt(1024, 768), 9)),
            "gend": _shapes(network_generator.MultiscaleDiscriminator(gen_opt(1024, 768))),
            "tocgd": _shapes(networks.define_D(input_nc=33, Ddownx2=True, Ddropout=True, n_layers_D=3, spectral=False, num_D=2)),
        }
    for k in ref:
        assert list(mine[k].keys()) == list(ref[k].keys()), k
        assert mine[k] == ref[k], k


def test_public_names():
    import network_generator
    import networks
    import sync_batchnorm
    for n in ["ConditionGenerator", "VGGLoss", "GANLoss", "load_checkpoint", "save_checkpoint", "define_D", "make_grid", "ResBlock",
              "Vgg19", "MultiscaleDiscriminator", "NLayerDiscriminator", "weights_init", "get_norm_layer"]:
        assert hasattr(networks, n), n
    for n in ["SPADEGenerator", "MultiscaleDiscriminator", "GANLoss", "BaseNetwork", "SPADENorm", "SPADEResBlock", "MaskNorm",
              "NLayerDiscriminator", "get_nonspade_norm_layer"]:
        assert hasattr(network_generator, n), n
    assert hasattr(sync_batchnorm, "DataParallelWithCallback")


def test_reference_error_conventions():
    import network_generator
    import networks
    o = gen_opt(1024, 768)
    o.num_upsampling_layers = "bogus"
    with pytest.raises(ValueError):
        network_generator.SPADEGenerator(o, 9)
    with pytest.raises(AssertionError):
        networks.ResBlock(4, 8, scale="sideways")
    with pytest.raises(Exception):
        networks.load_checkpoint(torch.nn.Linear(1, 1), "/nonexistent/ckpt.pth")
    with pytest.raises(ValueError):
        network_generator.SPADENorm(o, "aliasbogus", 8, 7)


def test_cpu_tensors_are_refused():
    import network_generator
    import networks
    with contextlib.redirect_stdout(io.StringIO()):
        g = network_generator.SPADEGenerator(gen_opt(128, 128), 9).eval()
        t = networks.ConditionGenerator(tocg_opt(), 4, 16, 13, ngf=96, norm_layer=torch.nn.BatchNorm2d).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            g(torch.zeros(1, 9, 128, 128), torch.zeros(1, 7, 128, 128))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            t(torch.zeros(1, 4, 64, 48), torch.zeros(1, 16, 64, 48))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "hrviton_sm100.h")).read()
    declared = set(re.findall(r"\b(hrv_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"hrv_status", "hrv_dtype"}
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libhrviton_sm100.so does not export %s" % name
    assert set(capi.EXPORTS) == declared
    # the fp16-storage twins (include/hrviton_sm100_f16.h is generated from the main header: same signatures)
    hdr16 = open(os.path.join(ROOT, "include", "hrviton_sm100_f16.h")).read()
    twins = set(re.findall(r"\b(hrv_[a-z0-9_]+_f16)\s*\(", hdr16))
    assert twins == set(capi.EXPORTS_F16) and len(twins) == len(declared) - 3
    for name in sorted(twins):
        assert hasattr(lib, name), "libhrviton_sm100.so does not export %s" % name
    for name in capi.FLAVOURED:  # twin prototypes are textually the main header's with the suffixed name
        a = re.search(r"^(?:int|size_t) %s\(([^;]*?)\);" % name, hdr, flags=re.M | re.S).group(1)
        b = re.search(r"^(?:int|size_t) %s_f16\(([^;]*?)\);" % name, hdr16, flags=re.M | re.S).group(1)
        assert a == b, name
    assert lib.hrv_version() >= 200
