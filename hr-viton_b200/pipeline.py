"""End-to-end try-on inference: the body of the reference's test_generator.py loop (test_generator.py:117-219) as a callable.

    tocg (256x192) -> cloth-mask composition -> parse map (up-sample, 15x15 Gaussian, arg-max, 13->7 regroup) -> hi-res cloth warp
    (+ occlusion handling) -> SPADEGenerator -> image

Both networks run on this repo's kernels; the glue between them is two kernels (hrv_parse_blur_argmax, hrv_flow_warp_nchw) instead of
the reference's ~25 torch ops and three numpy round trips (the `> 0.5` thresholds at test_generator.py:128,163 run on the device).
BASELINE.json configs[4]."""
import torch
import torch.nn.functional as F

from . import ops
from .train_step import GROUP_OF_13, OCCLUSION_CLASSES


def tryon_forward(tocg, generator, inputs, opt=None, occlusion=True, clothmask_composition="warp_grad", datasetting="paired"):
    """inputs: the dict a cp_dataset_test.CPDataLoader batch provides (cloth / cloth_mask may be {setting: tensor} dicts or tensors);
    CUDA fp32 NCHW tensors at the fine resolution.  Returns (output image (N,3,H,W), warped_cloth, parse (N,7,H,W))."""
    pick = lambda v: v[datasetting] if isinstance(v, dict) else v
    clothes, pre_cm = pick(inputs["cloth"]), pick(inputs["cloth_mask"])
    agnostic, densepose, parse_agnostic = inputs["agnostic"], inputs["densepose"], inputs["parse_agnostic"]
    with torch.no_grad():
        pre_cm = (pre_cm > 0.5).float()                                                        # test_generator.py:128 (numpy round trip there)
        input1 = torch.cat([F.interpolate(clothes, size=(256, 192), mode="bilinear"), F.interpolate(pre_cm, size=(256, 192), mode="nearest")], 1)
        input2 = torch.cat([F.interpolate(parse_agnostic, size=(256, 192), mode="nearest"), F.interpolate(densepose, size=(256, 192), mode="bilinear")], 1)
        flow_list, fake_segmap, _, warped_cm = tocg(input1, input2) if opt is None else tocg(opt, input1, input2)
        if clothmask_composition != "no_composition":
            cm = (warped_cm > 0.5).float() if clothmask_composition == "detach" else warped_cm   # test_generator.py:163-176
            fake_segmap = torch.cat([fake_segmap[:, :3], fake_segmap[:, 3:4] * cm, fake_segmap[:, 4:]], 1)
        n, _, ih, iw = clothes.shape
        div = ((96 - 1.0) / 2.0, (128 - 1.0) / 2.0)
        if occlusion:
            _, parse, overlap = ops.parse_blur_argmax(fake_segmap.float(), (ih, iw), group_of=GROUP_OF_13, groups=7, want_idx=False,
                                                      overlap_classes=OCCLUSION_CLASSES)
            warped_cloth, _, _ = ops.flow_warp_nchw(flow_list[-1], clothes.float(), (ih, iw), div, mask=pre_cm, overlap=overlap, composite=True)
        else:
            _, parse = ops.parse_blur_argmax(fake_segmap.float(), (ih, iw), group_of=GROUP_OF_13, groups=7, want_idx=False)
            warped_cloth, _ = ops.flow_warp_nchw(flow_list[-1], clothes.float(), (ih, iw), div)
        output = generator(torch.cat((agnostic, densepose, warped_cloth), 1), parse)
    return output, warped_cloth, parse


def psnr(a, b, peak=2.0):
    """PSNR in dB for images in [-1, 1] (peak-to-peak 2)."""
    mse = float(((a.float() - b.float()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * torch.log10(torch.tensor(peak * peak / mse)).item()
