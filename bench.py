#!/usr/bin/env python
"""Benchmark of the EMLight illumination hot path on B200 (contract: see the task statement / DESIGN.md section 6).

Workload (BASELINE.json configs[2], the one the metric "illumination maps/sec (crop -> 128x256 HDR pano)" describes):
    B=256 LDR crops (3x192x256) per GPU -> DenseNet-BC regression (eval-mode BN, fp32 storage,
    bf16x3 split tensor-core MMAs) -> light composition -> spherical-Gaussian render -> (B,3,128,256) fp32 panoramas.
Weak scaling: every rank processes its own B crops; the path is per-sample independent, so there is no data-path collective.

One JSON line on rank 0.  `--impl reference` times the CPU restatement of the reference (oracle/, the same ATen CPU ops the
reference issues) on the host cores instead.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "illumination maps/sec (256x192 crop -> 128x256 HDR pano)"
N_ANCHORS = 128


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.05)
        except Exception as e:      # noqa: BLE001
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def cpu_reference_step(batch, sd, x, dirs, threads):
    """One pass of the reference path on the CPU: DenseNet forward (eval BN) -> colour composition -> SG render."""
    import numpy as np
    import torch
    from oracle import densenet_oracle as DO, render_oracle as RO
    torch.set_num_threads(threads)
    with torch.no_grad():
        out = DO.densenet_forward(sd, x, training=False)
    # train.py:115-122: colours from the heads, shared anchors, size 0.0025, then the reference's light-by-light render
    cols = (out["distribution"][:, :, None] * (out["intensity"][:, :, None] * 500.0) * out["rgb_ratio"][:, None, :]).reshape(batch, -1)
    sizes = torch.full((batch, N_ANCHORS), 0.0025)
    return RO.convert_to_panorama_torch(torch.from_numpy(dirs).repeat(batch, 1), sizes, cols)


def time_cpu(sample, steps, warmup):
    """Times the CPU path with the intra-op thread count that serves it best (torch's CPU convolutions stop scaling --
    and regress -- well below a large host's core count, so every power of two up to the core count is tried once)."""
    import numpy as np
    import torch
    from oracle import densenet_oracle as DO, render_oracle as RO
    ncpu = os.cpu_count() or 1
    sd = DO.init_state_dict(seed=0, n_anchors=N_ANCHORS)
    x = torch.rand(sample, 3, 192, 256, generator=torch.Generator().manual_seed(1234))
    dirs = RO.sphere_points(N_ANCHORS).astype(np.float32).reshape(1, -1)
    cands = sorted({min(ncpu, c) for c in (8, 16, 32, 64, ncpu)})
    best, threads = None, ncpu
    for c in cands:
        cpu_reference_step(1, sd, x[:1], dirs, c)
        t0 = time.perf_counter()
        cpu_reference_step(sample, sd, x, dirs, c)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, threads = dt, c
    for _ in range(warmup):
        cpu_reference_step(sample, sd, x, dirs, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_step(sample, sd, x, dirs, threads)
    dt = (time.perf_counter() - t0) / steps
    return sample / dt, dt, threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="crops per GPU per step")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--cpu-sample", type=int, default=8, help="crops per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    config = {"workload": "BASELINE configs[2]: DenseNet-BC regression fwd (eval BN) + SG render -> 128x256 pano",
              "batch_per_gpu": args.batch, "global_batch": args.batch * world, "input": "3x192x256 fp32 NCHW",
              "anchors": N_ANCHORS, "parallelism": "dp%d (independent shards, no collective)" % world,
              "l2": "per-step inputs (151 MB) and activations (>10 GB) exceed the 126 MB L2; no explicit flush"}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = args.cpu_sample
        v, dt, threads = time_cpu(sample, max(args.steps, 1), warmup)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "maps/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": dict(config, batch_per_gpu=sample, global_batch=sample),
                "cpu_baseline": {"value": v, "unit": "maps/s", "cores": threads, "kind": "port",
                                 "sample": "best of {8,16,32,64,all} intra-op threads; %d crops per step, oracle/ DenseNet forward + light-by-light SG render, torch CPU ATen ops (eval BN)" % sample},
                "e2e": {"value": v, "unit": "maps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import emlight_b200 as E
    from emlight_b200 import build, parallel
    parallel.init("nccl", dev)
    build.build()
    B = args.batch
    torch.manual_seed(0)
    net = E.DenseNet(n_anchors=N_ANCHORS, precision=args.precision).to(dev).eval()
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(B, 3, 192, 256, generator=gen).pin_memory()
    x = x_host.to(dev)
    dirs = torch.from_numpy(E.sphere_points(N_ANCHORS)).float().to(dev)
    pano_host = torch.empty(B, 3, 128, 256, dtype=torch.float32).pin_memory()

    def step(inp):
        with torch.no_grad():
            o = net(inp)
            return E.render_from_params(o["distribution"], o["intensity"], o["rgb_ratio"], dirs=dirs)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev)

    for _ in range(warmup):
        step(x)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda: step(x), args.steps)
    # end to end through the public API with host buffers: pinned H2D of the crops, D2H of the panoramas, every step
    # Every step: pinned H2D of that step's crops, the public modules, D2H of that step's panoramas.  The copies run on two side
    # streams (full-duplex PCIe) and overlap the neighbouring steps' kernels, the way a serving loop would drive the modules.
    h2d_s, d2h_s = torch.cuda.Stream(), torch.cuda.Stream()
    xin = [torch.empty_like(x) for _ in range(2)]

    def e2e_run(n):
        main = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        h2d_s.wait_stream(main); d2h_s.wait_stream(main)
        with torch.cuda.stream(h2d_s):
            xin[0].copy_(x_host, non_blocking=True); ev_in[0].record(h2d_s)
        for i in range(n):
            if i + 1 < n:
                with torch.cuda.stream(h2d_s):
                    if i >= 1:
                        h2d_s.wait_event(ev_free[(i + 1) % 2])          # step i-1 no longer reads this input buffer
                    xin[(i + 1) % 2].copy_(x_host, non_blocking=True); ev_in[(i + 1) % 2].record(h2d_s)
            main.wait_event(ev_in[i % 2])
            pano = step(xin[i % 2])
            ev_free[i % 2].record(main)
            d2h_s.wait_stream(main)
            with torch.cuda.stream(d2h_s):
                pano_host.copy_(pano, non_blocking=True)
            pano.record_stream(d2h_s)
        main.wait_stream(d2h_s)

    e2e_run(2)
    ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = B * world * args.steps / (ms / 1e3)
    e2e_value = B * world * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel family, measured live with CUDA events on the launching stream
    net.launch_log = []
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); step(x); t1.record()
    torch.cuda.synchronize()
    fam = {}
    for family, name, abytes, flops, a, b in net.launch_log:
        f = fam.setdefault(family, {"ms": 0.0, "bytes": 0, "flops": 0, "launches": 0})
        f["ms"] += a.elapsed_time(b); f["bytes"] += abytes; f["flops"] += flops; f["launches"] += 1
    net.launch_log = None
    step_ms = t0.elapsed_time(t1)
    hbm_peak, tf_peak, peak_kind = peaks()
    top = max(fam, key=lambda k: fam[k]["ms"])
    roofs = {k: {"ms_per_step": round(v["ms"], 3), "launches": v["launches"], "share_of_step": round(v["ms"] / step_ms, 3),
                 "GBps": round(v["bytes"] / v["ms"] / 1e6, 1), "TFLOPs": round(v["flops"] / v["ms"] / 1e9, 2)} for k, v in fam.items()}
    achieved = fam[top]["bytes"] / fam[top]["ms"] / 1e6
    # DRAM traffic of the dominant family per launch, from the committed ncu launch list of this same command (profiles/)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp) and B == 256:
        tj = json.load(open(tp))
        if top in tj["families"]:
            traffic = tj["families"][top]["traffic_bytes_per_launch"]
            traffic_src = "profiles/r01_traffic.json (ncu dram__bytes_read+write, avg per launch); algorithmic bytes per launch = %d" % (
                fam[top]["bytes"] // fam[top]["launches"])
    roofline = {"kernel": {"dense_layer": "dense_layer_kernel<SPLIT> (csrc/dense_layer.cu)",
                           "conv1x1": "conv1x1_persist_kernel<SPLIT,RELU>", "conv3x3": "conv3x3_roll_kernel<48,SPLIT>",
                           "pool1x1": "conv_gemm_kernel<2,SPLIT>"}[top],
                "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s", "launches": fam[top]["launches"],
                "frac": round(achieved / hbm_peak, 4), "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                "traffic": traffic, "traffic_source": traffic_src, "families": roofs,
                "note": "achieved = algorithmic bytes (inputs read once + outputs written once, fp32) / CUDA-event time, summed over the family's launches in one step"}
    fc_launches = 5 if (args.precision != "fp32" and B >= 32) else 1               # bf16 split + 4 GEMM slices, or the SIMT linear
    launches_per_step = 1 + sum(v["launches"] for v in fam.values()) + 1 + fc_launches + 1 + 1   # stem + convs + head_pool + fc + heads + render

    # ---- secondary workloads (reported, not the headline): BASELINE configs[1] and configs[0]
    extra = {}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Bt = 64
        xt = torch.rand(Bt, 3, 192, 256, generator=gen).to(dev)
        yt = torch.softmax(3 * torch.randn(Bt, N_ANCHORS, generator=gen), 1).view(Bt, N_ANCHORS, 1).to(dev)
        loss_fn = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=Bt)
        net.train()

        def cfg1():                     # DenseNet fwd (batch-statistic BN, as train.py runs it) + Sinkhorn-EMD fwd + d/d(dist_pred)
            with torch.no_grad():
                o = net(xt)
            d = o["distribution"].detach().view(Bt, N_ANCHORS, 1).requires_grad_()
            loss_fn(d, yt).sum().backward()
            return d.grad
        for _ in range(3):
            cfg1()
        ms1 = timed(cfg1, 5) / 5
        # BASELINE configs[1]/[3] as train.py runs it: forward + 5-term loss + full backward + Adam step
        sys.path.insert(0, os.path.join(ROOT, "examples"))
        from train_regression_synthetic import synthetic_batch, train_step
        opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999))
        tb = synthetic_batch(Bt, N_ANCHORS, gen, dev)
        l2 = torch.nn.MSELoss()
        for _ in range(2):
            train_step(net, loss_fn, l2, opt, tb, N_ANCHORS, 1)
        ms_tr = timed(lambda: train_step(net, loss_fn, l2, opt, tb, N_ANCHORS, 1), 3) / 3
        net.eval()
        x1 = x[:1].contiguous()
        for _ in range(3):
            step(x1)
        ms0 = timed(lambda: step(x1), 20) / 20
        net.use_cuda_graph = True                       # replay the 104-launch forward as one CUDA graph (emlight_b200/graphs.py)
        for _ in range(3):
            step(x1)
        ms0g = timed(lambda: step(x1), 50) / 50
        net.use_cuda_graph = False
        extra = {"config1_densenet_fwd_trainBN_plus_sinkhorn_fwd_bwd_b64": {"ms_per_step": round(ms1, 3), "maps_per_s": round(Bt / ms1 * 1e3, 1)},
                 "train_step_fwd_bwd_adam_b64": {"ms_per_step": round(ms_tr, 3), "maps_per_s": round(Bt / ms_tr * 1e3, 1)},
                 "config0_single_crop_latency_ms": round(ms0, 3), "config0_single_crop_latency_cuda_graph_ms": round(ms0g, 3)}
        # BASELINE configs[4] (per GPU share of B=512 over 8 GPUs = 64; here the whole 512 on one GPU): needlet j=3 projection + reconstruction
        try:
            from emlight_b200.needlets import NeedletTransform
            del xt, yt, tb
            torch.cuda.empty_cache()
            nt = NeedletTransform(jmax=3, device=dev)
            pn = torch.exp(torch.randn(512, 3, 128, 256, device=dev))
            for _ in range(2):
                nt.reconstruct(nt.project(pn))
            ms_p = timed(lambda: nt.project(pn), 3) / 3
            cf = nt.project(pn)
            ms_r = timed(lambda: nt.reconstruct(cf), 3) / 3
            extra["config4_needlets_j3_b512"] = {"project_ms": round(ms_p, 3), "reconstruct_ms": round(ms_r, 3),
                                                 "maps_per_s_project_plus_reconstruct": round(512 / (ms_p + ms_r) * 1e3, 1),
                                                 "coefficients": nt.n}
            del nt, pn, cf
        except Exception as e:                                  # noqa: BLE001  (secondary workload: never fail the headline line)
            extra["config4_needlets_j3_b512"] = {"error": "%s: %s" % (type(e).__name__, e)}
        # BASELINE configs[3]'s GenProjector part (G step + D step: forward, tape backward of emlight_b200/gp_train.py, Adam).  Opt-in
        # (EML_BENCH_GAN=1) until that path has had its first B200 run -- see DESIGN.md 4.2 / tools/gpu_pending.sh.
        if os.environ.get("EML_BENCH_GAN") == "1":
            try:
                import argparse as _ap
                from train_genprojector_synthetic import synthetic_batch as gan_batch
                torch.cuda.empty_cache()
                gopt = _ap.Namespace(ngf=64, ndf=64, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", norm_D="spectralinstance",
                                     semantic_nc=3, label_nc=3, output_nc=3, num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0,
                                     num_D=2, n_layers_D=4, netD_subarch="n_layer", no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[0],
                                     isTrain=True, gan_mode="hinge", lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False)
                gm = E.Pix2PixModel(gopt)
                gm.train()
                gm.autograd = True
                og, od = gm.create_optimizers(gopt)
                Bg = 4
                gd = gan_batch(Bg, gen, dev)

                def gan_iter():
                    og.zero_grad(); gl, _ = gm(gd, "generator"); sum(gl.values()).mean().backward(); og.step()
                    od.zero_grad(); dl = gm(gd, "discriminator"); sum(dl.values()).mean().backward(); od.step()
                gan_iter()
                ms_g = timed(gan_iter, 2) / 2
                extra["config3_genprojector_G_step_plus_D_step_b4"] = {"ms_per_iteration": round(ms_g, 3), "maps_per_s": round(Bg / ms_g * 1e3, 2)}
                del gm, og, od, gd
            except Exception as e:                              # noqa: BLE001  (secondary workload: never fail the headline line)
                extra["config3_genprojector_G_step_plus_D_step_b4"] = {"error": "%s: %s" % (type(e).__name__, e)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, threads = time_cpu(args.cpu_sample, 3, 1)
        cpu = {"value": v, "unit": "maps/s", "cores": threads, "kind": "port",
               "sample": "best of {8,16,32,64,all} intra-op threads; %d crops per step x 3 steps, oracle/ DenseNet forward + light-by-light SG render, torch CPU ATen ops (eval BN)" % args.cpu_sample}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "maps/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"bf16x3": "f32 storage, bf16x3 split tensor-core MMA (fp32-grade)", "bf16": "f32 storage, bf16 MMA",
                          "fp32": "f32 FFMA"}[args.precision],
                "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": "maps/s", "ms_per_step": ms_e2e / args.steps,
                        "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": pano_host.numel() * 4},
                "gpu_launches": launches_per_step * args.steps, "clocks": sampler.summary(), "roofline": roofline,
                "cpu_baseline": cpu, "other_workloads": extra}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
