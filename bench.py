#!/usr/bin/env python
"""bench.py — throughput of the HR-VITON hot path on B200 (driver contract: see the task statement).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm on the host CPU cores (oracle port)

Prints ONE JSON line on rank 0.  Workloads (config.workload):
  train_stage2  (default) one full train_generator.py step per "step": frozen tocg -> warp -> SPADE G fwd+bwd -> D fwd+bwd
                (hinge + feature matching) + VGG loss -> Adam(G), then the D update (2nd G fwd, D fwd+bwd, Adam(D)),
                1024x768, bf16 activations (BASELINE.json configs[3], per-GPU batch --batch)
  train_stage1  one full train_condition.py step (tocg fwd+bwd with train-mode BatchNorm, stage-1 D, L1+VGG+TV+CE+LSGAN, Adam x2),
                1024x768 per-GPU batch 4 (BASELINE.json configs[1])
  gen_fwd       SPADEGenerator inference forward, 1024x768, per-GPU batch 8 (BASELINE.json configs[2])
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1024, 768
GEN_GFLOP_PER_IMG = 1636.4  # conv FLOPs of one SPADEGenerator forward at 1024x768


def gen_opt():
    return types.SimpleNamespace(norm_G="spectralaliasinstance", gen_semantic_nc=7, ngf=64, num_upsampling_layers="most",
                                 fine_height=H, fine_width=W, cuda=True)


def conv_traffic():
    """Average DRAM bytes (read+write) per conv_igemm launch of one step, from the committed ncu pass
    (profiles/r1_conv_dram_traffic.json, written by tools/summarise_ncu_traffic.py); None when that pass was not taken."""
    p = os.path.join(ROOT, "profiles", "r1_conv_dram_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)["avg_dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_generator(device):
    import torch

    import network_generator
    from hrviton_b200 import spade
    torch.manual_seed(0)
    g = network_generator.SPADEGenerator(gen_opt(), 9)
    g.init_weights("xavier", 0.02)
    with torch.no_grad():
        for m in g.modules():
            if isinstance(m, spade.SPADENorm):
                m.noise_scale.normal_(0.0, 0.1)
    g = g.to(device).eval()
    with torch.no_grad():  # realistic spectral-norm u/v (a fresh module holds random vectors)
        for m in g.modules():
            if hasattr(m, "weight_orig"):
                for _ in range(3):
                    spade._sigma(m, True)
    return g


def synth_batch(b, device, seed):
    import torch
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.rand((b, 9, H, W), generator=gen) * 2 - 1
    lab = torch.randint(0, 7, (b, H // 16, W // 16), generator=gen)
    lab = lab.repeat_interleave(16, 1).repeat_interleave(16, 2)
    seg = torch.zeros((b, 7, H, W)).scatter_(1, lab[:, None], 1.0)
    return x, seg


def cpu_baseline_gen(steps=1, warmup=0, budget_s=240.0):
    """The reference algorithm on the host cores: the oracle port (oracle/hrviton_oracle.py — the reference is Python and
    cannot travel to the GPU box) — SPADEGenerator forward, fp32, one 1024x768 image per step."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hrviton_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    import network_generator
    torch.manual_seed(0)
    m = network_generator.SPADEGenerator(gen_opt(), 9)
    m.init_weights("xavier", 0.02)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    x, seg = synth_batch(1, "cpu", 1)
    noise_fn = lambda b, hh, ww: torch.randn(b, hh, ww)
    times = []
    t_begin = time.time()
    done = 0
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.time()
            orc.spade_generator_forward(sd, x, seg, noise_fn)
            dt = time.time() - t0
            if i >= warmup:
                times.append(dt)
                done += 1
            if time.time() - t_begin + dt > budget_s and done >= 1:
                break
    times.sort()
    med = times[len(times) // 2]
    return {"value": 1.0 / med, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": "SPADEGenerator fwd fp32, 1 image 1024x768 per step, %d timed step(s), torch %s CPU" % (len(times), torch.__version__),
            "s_per_image": med, "steps_done": len(times)}


def cpu_baseline_train(budget_s=240.0):
    """Bounded CPU sample of the training workload: the oracle port's SPADEGenerator forward + backward (the 56 % of the
    stage-2 step's FLOPs that dominate it) on ONE 512x384 image, fp32, host cores; reported as
    1024x768-equivalent images/s (x 1/4: the generator is fully convolutional, cost scales with pixels)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hrviton_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(min(cores, 64))
    import network_generator
    torch.manual_seed(0)
    opt = gen_opt()
    opt.fine_height, opt.fine_width = 512, 384
    m = network_generator.SPADEGenerator(opt, 9)
    m.init_weights("xavier", 0.02)
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point() and not k.endswith(("weight_u", "weight_v"))) for k, v in m.state_dict().items()}
    x = torch.rand(1, 9, 512, 384) * 2 - 1
    seg = torch.zeros(1, 7, 512, 384)
    seg[:, 0] = 1
    t0 = time.time()
    out = orc.spade_generator_forward(sd, x, seg, lambda b, hh, ww: torch.randn(b, hh, ww))
    out.mean().backward()
    dt = time.time() - t0
    return {"value": 0.25 / dt, "unit": "images/s", "cores": min(cores, 64), "kind": "port",
            "sample": "oracle SPADEGenerator fwd+bwd fp32 on one 512x384 image (%.1f s), scaled x1/4 to 1024x768; G fwd+bwd is ~56%% of the stage-2 step FLOPs, so the full-step CPU rate is lower still" % dt,
            "s_per_sample": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "train_stage2":
        vals = [cpu_baseline_train() for _ in range(max(1, min(args.steps, 3)))]
        vals.sort(key=lambda c: c["value"])
        cb = vals[len(vals) // 2]
        cb["steps_done"] = len(vals)
        cb["s_per_image"] = cb["s_per_sample"] * 4
        wl = "train_stage2 (bounded sample: generator fwd+bwd only, 512x384, pixel-scaled; reference algorithm on host CPU)"
    else:
        cb = cpu_baseline_gen(steps=max(1, args.steps), warmup=min(args.warmup, 1))
        wl = "gen_fwd: SPADEGenerator inference forward 1024x768 (reference algorithm, host CPU, 1 image per step)"
    line = {"impl": "reference", "metric": "1024x768 try-on images/sec", "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus,
            "steps": cb["steps_done"], "warmup": min(args.warmup, 1), "ms_per_step": cb["s_per_image"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 8 for train_stage2 and gen_fwd, 4 for train_stage1)")
    ap.add_argument("--workload", default="train_stage2", choices=["train_stage2", "train_stage1", "gen_fwd"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="train_stage2: launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--dump-profile", default="", help="write the per-launch CUDA-event profile of one step as CSV")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION (the image's default): keep stdout to the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist

    import hrv_loader
    hrv_loader.load()
    from hrviton_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(3, args.warmup)
    K = args.steps
    train = args.workload in ("train_stage2", "train_stage1")
    stage1 = args.workload == "train_stage1"
    B = args.batch or (4 if stage1 else 8)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if train:
        os.environ.setdefault("HRV_VGG_RANDOM_INIT", "1")  # no network: torchvision weights cannot be downloaded
        import network_generator
        import networks
        from hrviton_b200 import ddp, train_step
        topt = types.SimpleNamespace(warp_feature="T1", out_layer="relu", cuda=True)
        torch.manual_seed(0)
        tocg = networks.ConditionGenerator(topt, 4, 16, 13, ngf=96, norm_layer=torch.nn.BatchNorm2d)
        with torch.no_grad():
            for mod in tocg.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.running_mean.normal_(0, 0.1)
                    mod.running_var.uniform_(0.5, 1.5)
        tocg = tocg.to(dev).eval()
        g = None if stage1 else build_generator(dev).train()
        dopt = gen_opt()
        dopt.ndf, dopt.norm_D, dopt.n_layers_D, dopt.num_D, dopt.no_ganFeat_loss = 64, "spectralinstance", 3, 2, False
        vgg = networks.Vgg19().to(dev).eval()
        reducers = {}
        if stage1:
            import contextlib
            import io
            tocg.train()
            with contextlib.redirect_stdout(io.StringIO()):
                D = networks.define_D(input_nc=33, Ddownx2=True, Ddropout=True, n_layers_D=3, spectral=False, num_D=2)
            D = D.to(dev).train()
            if world > 1:
                reducers = {"G": ddp.GradBucketReducer(list(tocg.parameters())), "D": ddp.GradBucketReducer(list(D.parameters()))}
            tr1 = train_step.Stage1Trainer(tocg, D, vgg, reducers=reducers)

            class _Shim:  # same step/capture/replay surface as Stage2Trainer
                def step(self, b, h, w):
                    return tr1.step(b)

                def capture(self, b, h, w):
                    tr1.capture(b)

                def replay(self, b=None):
                    return tr1.replay(b)
            trainer = _Shim()
            make_batch = train_step.synthetic_batch_stage1
        else:
            D = network_generator.MultiscaleDiscriminator(dopt)
            D.init_weights("xavier", 0.02)
            D = D.to(dev).train()
            if world > 1:
                reducers = {"G": ddp.GradBucketReducer(list(g.parameters())), "D": ddp.GradBucketReducer(list(D.parameters()))}
            trainer = train_step.Stage2Trainer(tocg, g, D, vgg, reducers=reducers)
            make_batch = train_step.synthetic_batch
        batch_h = {k: v.pin_memory() for k, v in make_batch(B, H, W, "cpu", seed=100 + rank).items()}
        batch_d = {k: v.to(dev) for k, v in batch_h.items()}
        loss_h = torch.empty(2, dtype=torch.float32).pin_memory()
        h2d_bytes = int(sum(v.numel() * v.element_size() for v in batch_h.values()))
        d2h_bytes = 8

        # multi-rank runs launch eagerly: capturing the NCCL gradient all-reduce inside the step graph hung on the 2-GPU box
        # (profiles/README.md); the ~10% CPU launch overhead shows up in the N>1 numbers, not in N=1
        use_graph = not args.no_graph and world == 1
        if use_graph:
            try:
                trainer.step(batch_d, H, W)  # first eager step: lazy initialisation (optimizer state, func attributes, caches)
                l_cap = ops.LAUNCHES[0]
                trainer.capture(batch_d, H, W)
                launches_per_replay = (ops.LAUNCHES[0] - l_cap) // 3  # capture() runs 2 warm steps + the captured one
            except Exception as e:  # noqa: BLE001 - report and fall back to eager launches
                import traceback
                sys.stderr.write("CUDA graph capture failed (%s: %s); running eagerly\n%s\n" % (type(e).__name__, e, traceback.format_exc()[-1500:]))
                use_graph = False

        def step_resident():
            if use_graph:
                ops.LAUNCHES[0] += launches_per_replay
                return trainer.replay()
            return trainer.step(batch_d, H, W)

        def step_e2e():
            if use_graph:
                ops.LAUNCHES[0] += launches_per_replay
                out = trainer.replay(batch_h)
            else:
                bd = {k: v.to(dev, non_blocking=True) for k, v in batch_h.items()}
                out = trainer.step(bd, H, W)
            lk = ("loss_g", "loss_d") if stage1 else ("loss_gen", "loss_dis")
            loss_h.copy_(torch.stack([out[lk[0]].float(), out[lk[1]].float()]), non_blocking=True)
    else:
        g = build_generator(dev)
        x_h, seg_h = synth_batch(B, "cpu", 100 + rank)
        x_h, seg_h = x_h.pin_memory(), seg_h.pin_memory()
        x_d, seg_d = x_h.to(dev), seg_h.to(dev)
        out_h = torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory()
        h2d_bytes = int(x_h.numel() * 4 + seg_h.numel() * 4)
        d2h_bytes = int(out_h.numel() * 4)

        def step_resident():
            with torch.no_grad():
                return g(x_d, seg_d)

        def step_e2e():
            with torch.no_grad():
                xd = x_h.to(dev, non_blocking=True)
                sd = seg_h.to(dev, non_blocking=True)
                out = g(xd, sd)
                out_h.copy_(out, non_blocking=True)

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up: nvidia-smi needs ~0.5 s to emit its first sample
    for _ in range(W_):
        step_resident()
    if rank == 0:
        sampler.lines.clear()  # keep only samples taken during the timed region
    l0 = ops.LAUNCHES[0]
    ms = timed(step_resident, K)
    launches = ops.LAUNCHES[0] - l0
    clocks = sampler.stop() if rank == 0 else None
    value = B * world * K / (ms / 1e3)

    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, K)
    e2e_value = B * world * K / (ms_e2e / 1e3)

    # ---- per-kernel profile pass (CUDA events around every C-ABI launch) -> roofline of the dominant kernel
    ops.PROFILE = []
    torch.cuda.synchronize()
    # process-wide start/end range (the backward runs on autograd's own thread, which a push/pop range would not cover):
    # `ncu --nvtx --nvtx-include "hrv_profile_step" ...` = launch list of exactly ONE step of this command
    nvtx_id = torch.cuda.nvtx.range_start("hrv_profile_step")
    if train:
        trainer.step(batch_d, H, W)  # eager (events cannot be recorded inside a graph replay)
    else:
        step_resident()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_end(nvtx_id)
    prof = ops.PROFILE
    ops.PROFILE = None
    agg = {}
    if args.dump_profile and rank == 0:
        with open(args.dump_profile, "w") as f:
            f.write("kind,label,ms,work,rate_T_per_s\n")
            for kind, work, e0, e1, label in prof:
                t = e0.elapsed_time(e1)
                f.write("%s,%s,%.4f,%.4g,%.2f\n" % (kind, label, t, work, work / (t * 1e-3) / 1e12 if t > 0 else 0))
    for kind, work, e0, e1, _label in prof:
        a = agg.setdefault(kind, [0.0, 0.0, 0])
        a[0] += work
        a[1] += e0.elapsed_time(e1)
        a[2] += 1
    peaks = measured_peaks()
    conv_flops = agg.get("conv", [0, 0, 0])[0] + agg.get("conv_spade", [0, 0, 0])[0]
    conv_ms = agg.get("conv", [0, 0, 0])[1] + agg.get("conv_spade", [0, 0, 0])[1]
    conv_launches = agg.get("conv", [0, 0, 0])[2] + agg.get("conv_spade", [0, 0, 0])[2]
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    roofline = {"kernel": "conv_igemm_kernel + conv_pixn_kernel (tcgen05 implicit-GEMM convolution, all %d launches of one step)" % conv_launches,
                "bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tf_sustained"],
                "traffic": conv_traffic() if (train and not stage1 and B == 8) else None,  # the ncu pass was taken on the default workload "peak_source": peaks["src"] + " (bf16 sustained)",
                "avg_launch_ms": conv_ms / max(1, conv_launches), "algorithmic_gflop_per_launch": conv_flops / 1e9 / max(1, conv_launches)}
    total_prof_ms = sum(a[1] for a in agg.values())
    breakdown = {k: {"ms": round(a[1], 3), "launches": a[2], "share": round(a[1] / total_prof_ms, 4)} for k, a in agg.items()}
    if "instnorm_stats" in agg:
        a = agg["instnorm_stats"]
        breakdown["instnorm_stats"]["achieved_GBps"] = round(a[0] / (a[1] * 1e-3) / 1e9, 1)
        breakdown["instnorm_stats"]["frac_of_hbm_peak"] = round(a[0] / (a[1] * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_train() if train else cpu_baseline_gen(steps=1, warmup=0)
        cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": "1024x768 try-on images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W_,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": ("train_stage1: full train_condition.py step (tocg fwd+bwd with train-mode BatchNorm, 3 stage-1 D passes fwd+bwd, VGG loss x5 fwd+dgrad, L1/TV/CE/LSGAN, Adam x2; README flags --Ddownx2 --Ddropout --lasttvonly --interflowloss --occlusion), 1024x768, bf16 activations / fp32 accumulate" if (train and stage1) else
                                        "train_stage2: full train_generator.py step (tocg fwd, G fwd+bwd, D fwd+bwd x2, 2nd G fwd, VGG loss fwd+dgrad, Adam x2), 1024x768, bf16 activations / fp32 accumulate; convs (fwd/dgrad/wgrad), norms, modulation, activations, pooling, weight packing, parse-map glue, VGG L1 on this repo's kernels; hi-res grid_sample, hinge/feature-matching reductions, spectral-norm power iteration, Adam = torch"
                                        if train else "gen_fwd: SPADEGenerator inference forward, 1024x768, bf16 activations, fp32 accumulate"),
                           "launch": ("cuda-graph replay of the whole step" if (train and use_graph) else "eager"), "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                           "l2": "activations per step (>10 GB) exceed the 126 MB L2; no explicit flush",
                           "weights": "xavier(0.02) random init, noise_scale~N(0,0.1)"},
                "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / K,
                        "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "kernel_breakdown": breakdown,
                "model_tflops": (None if stage1 else (8800.0 if train else GEN_GFLOP_PER_IMG) * value / 1e3)}
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
