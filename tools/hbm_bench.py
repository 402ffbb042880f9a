"""Micro-benchmark of the HBM-bound kernels at the SPADE generator's full-resolution shapes: algorithmic GB/s (DESIGN.md §3.2
byte counts) vs the measured copy peak.  Usage: python tools/hbm_bench.py [batch] [reps]  (reps=1 for ncu captures)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hrv_loader; hrv_loader.load()
from hrviton_b200 import ops
from hrviton_b200.ops import Act
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 10
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
dev = "cuda"

def timeit(fn):
    for _ in range(2 if REPS > 1 else 0): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS

def rnd(*shape): return torch.randn(*shape, device=dev).to(torch.bfloat16)

for (c0, c1, h, w, shift) in [(64, 16, 1024, 768, 1), (128, 16, 512, 384, 1), (32, 0, 1024, 768, 0)]:
    C = c0 + c1
    x0 = Act(rnd(B, h >> shift, w >> shift, c0)); x1 = Act(rnd(B, h, w, c1)) if c1 else None
    noise = torch.randn(B, h, w, device=dev); ns = torch.randn(C, device=dev) * 0.1
    px = B * h * w
    ms = timeit(lambda: ops.instnorm_stats(x0, shift, x1, h, w, noise, ns))
    unique = (B * (h >> shift) * (w >> shift) * c0 + px * c1) * 2 + px * 4
    print("instnorm_stats   C=%3d %4dx%-4d shift%d: %7.3f ms  algorithmic(virtual tensor) %6.0f GB/s = %4.1f%% of %d | unique bytes %6.0f GB/s" % (C, h, w, shift, ms, (px * C * 2 + px * 4) / ms / 1e6, 100 * (px * C * 2 + px * 4) / ms / 1e6 / peak, peak, unique / ms / 1e6))
    mean, rstd = ops.instnorm_stats(x0, shift, x1, h, w, noise, ns)
    dh, hh, gm = Act(rnd(B, h, w, C)), Act(rnd(B, h, w, C)), Act(rnd(B, h, w, C))
    ms = timeit(lambda: ops.norm_bwd(dh, hh, gm, x0, shift, x1, noise, ns, mean, rstd, 2, True))
    # reduce: reads dh,h,gamma,x (4C) + noise, writes dgb (2C) + dxn (C); apply: reads dxn, x (+noise), writes dx
    byt = px * C * 2 * 7 + px * 4 + px * C * 2 * 2 + (B * (h >> shift) * (w >> shift) * c0 + px * c1) * 2 + px * 4
    print("norm_bwd (3 launches) C=%3d %4dx%-4d      : %7.3f ms  algorithmic %6.0f GB/s = %4.1f%% of %d" % (C, h, w, ms, byt / ms / 1e6, 100 * byt / ms / 1e6 / peak, peak))
    ms = timeit(lambda: ops.act_bwd_bias(dh, hh, 2))
    print("act_bwd_bias     C=%3d %4dx%-4d      : %7.3f ms  algorithmic %6.0f GB/s = %4.1f%% of %d" % (C, h, w, ms, px * C * 6 / ms / 1e6, 100 * px * C * 6 / ms / 1e6 / peak, peak))
# appearance-flow warp at the tocg's largest feature level (x16 pixels for the 1024x768 stage-1 config)
for (c, h, w) in [(384, 512, 384), (384, 128, 96)]:
    src = Act(rnd(B, h, w, c)); dst = Act.empty(B, h, w, c)
    flow = torch.randn(B, h // 2, w // 2, 2, device=dev) * 3
    ms = timeit(lambda: ops.flow_warp(flow, src, dst))
    byt = B * h * w * (c * 2 * 2 + 8)
    print("flow_warp        C=%3d %4dx%-4d      : %7.3f ms  algorithmic %6.0f GB/s = %4.1f%% of %d" % (c, h, w, ms, byt / ms / 1e6, 100 * byt / ms / 1e6 / peak, peak))
