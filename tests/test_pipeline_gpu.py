"""GPU: BASELINE.json configs[4] — the end-to-end try-on inference pipeline (test_generator.py:117-219: tocg -> parse post-processing
-> hi-res warp with occlusion handling -> SPADEGenerator) through hrviton_b200.pipeline.tryon_forward, against the same pipeline
evaluated on the CPU with the oracle networks and the reference's separate torch ops.  Reported as the benchmark asks: PSNR and
max|delta| of the output image (LPIPS needs downloaded weights: unavailable offline)."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import hrviton_oracle as orc  # noqa: E402
from helpers import gen_opt, synth_state_dict, tocg_opt  # noqa: E402
from hrviton_b200 import ops, pipeline, synth, train_step  # noqa: E402

pytestmark = pytest.mark.gpu


def _oracle_pipeline(sdt, sdg, batch, h, w, seed, occlusion, rounding=None):
    cnt = [0]

    def noise(b, hh, ww):
        t = synth.spade_noise(b, hh, ww, seed, cnt[0])
        cnt[0] += 1
        return t

    class _Tocg:
        def __call__(self, i1, i2):
            return orc.tocg_forward(sdt, i1, i2)

    b = dict(batch)
    b["cloth_mask"] = (b["cloth_mask"] > 0.5).float()
    with torch.no_grad(), orc.storage_rounding(rounding):
        g_in, parse = train_step.make_generator_inputs(_Tocg(), b, h, w, occlusion=occlusion, unfused_parse=True)
        out = orc.spade_generator_forward(sdg, g_in, parse, noise)
    return out, g_in[:, 6:], parse


@pytest.mark.parametrize("precision,occlusion", [("bf16", True), ("fp16", True), ("bf16", False)])
def test_tryon_pipeline_matches_oracle_pipeline(precision, occlusion):
    import network_generator
    import networks
    n, h, w, seed = 2, 512, 384, 19
    sdt, sdg = synth_state_dict("tocg", seed), synth_state_dict("gen", seed + 1)
    batch = train_step.synthetic_batch(n, h, w, "cpu", seed=seed)
    out_r, wc_r, parse_r = _oracle_pipeline(sdt, sdg, batch, h, w, seed, occlusion)
    out_q, _, parse_q = _oracle_pipeline(sdt, sdg, batch, h, w, seed, occlusion, {"bf16": torch.bfloat16, "fp16": torch.float16}[precision])
    ops.set_precision(precision)
    try:
        tocg = networks.ConditionGenerator(tocg_opt(True), 4, 16, 13, ngf=96, norm_layer=torch.nn.BatchNorm2d)
        tocg.load_state_dict(sdt)
        G = network_generator.SPADEGenerator(gen_opt(h, w, True), 9)
        G.load_state_dict(sdg)
        tocg, G = tocg.cuda().eval(), G.cuda().eval()
        cnt = [0]

        def noise(b, hh, ww):
            t = synth.spade_noise(b, hh, ww, seed, cnt[0]).cuda()
            cnt[0] += 1
            return t

        G.noise_source = noise
        out, wc, parse = pipeline.tryon_forward(tocg, G, {k: v.cuda() for k, v in batch.items()}, occlusion=occlusion)
        torch.cuda.synchronize()
    finally:
        ops.set_precision("bf16")
    out, wc, parse = out.cpu(), wc.cpu(), parse.cpu()
    agree, agree_q = float((parse == parse_r).float().mean()), float((parse_q == parse_r).float().mean())
    d = (out - out_r).abs()
    dq = (out_q - out_r).abs()
    p, pq = pipeline.psnr(out, out_r), pipeline.psnr(out_q, out_r)
    print("PIPELINE %s occlusion=%s: PSNR %.2f dB (storage-rounded oracle %.2f dB)  max|d| %.3e (%.3e)  mean|d| %.3e (%.3e)  parse agreement %.5f (%.5f)  warped cloth max|d| %.3e"
          % (precision, occlusion, p, pq, float(d.max()), float(dq.max()), float(d.mean()), float(dq.mean()), agree, agree_q, float((wc - wc_r).abs().max())))
    # the arg-max of the blurred class scores flips on a few boundary pixels under 16-bit storage (also for the rounded oracle); where it
    # flips, the generator sees another label and the output differs by O(1): compare where the parse maps agree, and bound the rest
    assert agree > 0.995 and agree >= agree_q - 2e-3
    same = (parse == parse_r).all(1, keepdim=True).float()
    same_q = (parse_q == parse_r).all(1, keepdim=True).float()
    assert float((d * same).sum() / same.sum() / 3) <= 1.1 * float((dq * same_q).sum() / same_q.sum() / 3) + 1e-4
    assert p >= pq - 1.0  # within 1 dB of what the storage type costs the fp32 pipeline itself
    assert float((wc - wc_r).abs().mean()) < 2e-3
