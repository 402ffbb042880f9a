"""Where does the classic tcgen05 convolution kernel lose time?  Runs representative layers with the kernel's per-CTA stall counters
switched on (hrv_debug_set_conv_stats) and prints, per layer, the share of the MMA warp's life spent waiting for the A ring, the B
(weight) ring and a free TMEM accumulator, the producer warp's waits for free stages, and the first epilogue warp's wait for
accumulators.  Usage: python tools/conv_stall_probe.py [batch]   (env HRV_CONV_HALO etc. select kernel variants)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hrv_loader  # noqa: E402

hrv_loader.load()
from hrviton_b200 import capi, ops  # noqa: E402
from hrviton_b200.ops import Act  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = capi.lib()
L.hrv_debug_set_conv_stats.argtypes = [ctypes.c_void_p]
L.hrv_debug_set_conv_stats.restype = None
stats = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
cases = [  # cin, n_gemm, k, h, w, spade
    (128, 160, 3, 1024, 768, True), (128, 288, 3, 512, 384, True), (128, 544, 3, 256, 192, True), (128, 64, 3, 1024, 768, True),
    (160, 160, 3, 1024, 768, False), (256, 256, 3, 256, 192, False), (512, 512, 3, 128, 96, False), (1040, 512, 3, 64, 48, False),
    (80, 160, 3, 1024, 768, False), (160, 128, 3, 1024, 768, False), (288, 128, 3, 512, 384, False), (128, 128, 3, 512, 384, False),
    (64, 64, 3, 1024, 768, False), (80, 32, 3, 1024, 768, False), (32, 32, 3, 1024, 768, False), (144, 64, 3, 512, 384, False),
    (128, 128, 3, 512, 384, True), (32, 80, 3, 1024, 768, False), (3, 64, 3, 1024, 768, False), (9, 16, 3, 1024, 768, False),
    (64, 3, 3, 1024, 768, False), (3, 32, 3, 1024, 768, False)]
if os.environ.get("HRV_PROBE_CASES"):  # e.g. "9,10,11,3"
    cases = [cases[int(i)] for i in os.environ["HRV_PROBE_CASES"].split(",")]
os.environ.setdefault("HRV_CONV_PIXN", "0")  # default: probe the classic / pair kernels; HRV_CONV_PIXN=1 lets the dispatcher choose
print("%-34s %8s %8s | MMA warp: %6s %6s %6s %6s | producer: %6s %6s | epilogue w4: %6s" %
      ("layer", "ms", "TFLOP/s", "issue", "waitA", "waitB", "waitD", "freeA", "freeB", "waitAcc"))
for cin, ng, k, h, w, spade in cases:
    x = Act(torch.randn(B, h, w, ops.round_up(cin, 8), device="cuda").to(torch.bfloat16), c=cin)
    if spade:
        C = ng // 2
        wt = torch.randn(C, cin, k, k, device="cuda") * 0.05
        pw = ops.pack_weight(wt, (k // 2, k // 2), interleave=wt.clone())
        x0 = Act(torch.randn(B, h, w, C, device="cuda").to(torch.bfloat16))
        mean = torch.zeros(B, C, device="cuda"); rstd = torch.ones(B, C, device="cuda")
        noise = torch.randn(B, h, w, device="cuda"); ns = torch.zeros(C, device="cuda"); sh = torch.zeros(2 * C, device="cuda")
        out = Act.empty(B, h, w, C)
        fn = lambda: ops.conv2d_spade(x, pw, out, x0, 0, None, mean, rstd, noise, ns, sh, 2)
    else:
        wt = torch.randn(ng, cin, k, k, device="cuda") * 0.05
        pw = ops.pack_weight(wt, (k // 2, k // 2))
        out = Act.empty(B, h, w, ng)
        fn = lambda: ops.conv2d(x, pw, out)
    IT = int(os.environ.get("HRV_PROBE_ITERS", "5"))  # 0: exactly one launch per layer (ncu --set full captures)
    for _ in range(3 if IT else 0):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(IT):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / max(IT, 1) if IT else 1.0
    stats.zero_()
    L.hrv_debug_set_conv_stats(ctypes.c_void_p(stats.data_ptr()))
    fn()
    torch.cuda.synchronize()
    L.hrv_debug_set_conv_stats(None)
    s = stats.view(148, 16).double()
    s = s[s[:, 0] > 0]
    if s.shape[0] == 0:  # the dispatcher chose a kernel without stall counters (pixel-N): timing only
        print("%4d->%4d k%d %4dx%-4d %-7s %8.3f %8.1f |   (pixel-N kernel: no stall counters)" %
              (cin, ng, k, h, w, "spade" if spade else "linear", ms, 2.0 * cin * ng * k * k * B * h * w / ms / 1e9), flush=True)
        continue
    tot = s[:, 0].mean()
    f = lambda i: float(s[:, i].mean() / tot)
    fl = 2.0 * cin * ng * k * k * B * h * w
    print("%4d->%4d k%d %4dx%-4d %-7s %8.3f %8.1f |           %5.1f%% %5.1f%% %5.1f%% %5.1f%% |           %5.1f%% %5.1f%% |              %5.1f%%   (%d CTAs, %.0f tiles/CTA, %.0f cyc/tile)"
          % (cin, ng, k, h, w, "spade" if spade else "linear", ms, fl / ms / 1e9, 100 * (1 - f(1) - f(2) - f(3)), 100 * f(1), 100 * f(2), 100 * f(3),
             100 * f(4), 100 * f(5), 100 * float(s[:, 6].mean() / s[:, 7].mean()), s.shape[0], float(s[:, 8].mean()), float(tot / s[:, 8].mean())), flush=True)
    if float(s[:, 11:16].sum()) > 0:
        life = s[:, 7].mean()
        print("      staged epilogue, warp 4: x-load issue %.1f%% | wait accumulator %.1f%% | wait x %.1f%% | drain+barrier %.1f%% | chunks %.1f%% | fence+barrier+store %.1f%%" %
              tuple(100 * float(s[:, i].mean() / life) for i in (11, 12, 13, 14, 15, 9)), flush=True)
    elif float(s[:, 9].sum()) > 0 or float(s[:, 10].sum()) > 0:
        print("      epilogue warp 4 of the pair kernel: %.1f%% of its life issuing the x prefetch, %.1f%% in tcgen05.ld + wait::ld" %
              (100 * float(s[:, 9].mean() / s[:, 7].mean()), 100 * float(s[:, 10].mean() / s[:, 7].mean())), flush=True)
