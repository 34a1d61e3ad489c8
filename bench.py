"""Benchmark of the volumetric-render hot path (BASELINE.json: rays/s @ 64 samples/ray).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision tc_fp16|fp32]

A "step" = one pass of the hot path over one synthetic batch: at N=1 ONE 512x512 all-hit view of
the synth-313 body (BASELINE.json configs[1]: single B200, 262 144 rays x 64 samples, eval, no
jitter, random-init trained-like decoder).  At N>1 a step is N such views, each ray-sharded over
the N ranks (rank r renders slab r of every view) with one NCCL all-gather per view -- per-GPU
work is fixed (262 144 rays per step) => "scaling": "weak".

`value`  : rays/s with rays, packed volume and packed weights already resident in HBM; only
           nb_render_fwd launches (+ the all-gathers at N>1) are in the timed region.
`e2e`    : the same metric through the public API make_renderer(cfg, net).render(batch) with the
           batch in PINNED HOST memory: H2D of rays/near/far/pose per step, prepare_sp_input,
           weight pack, render, D2H of rgb_map+depth_map inside the timed region.
`--impl reference`: the reference's own CPU implementation of the path (the oracle port of
           /root/reference's if_clight_renderer + latent_xyzc + raw2outputs, validated bit-exact
           against the unmodified reference in the build container), all host threads, each step a
           bounded sample (--ref-rays rays) of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H = W = 512
S = 64
FLOP_PER_SAMPLE_AS_WRITTEN = 859904     # 2 x 429 952 MAC, layers of latent_xyzc.py:20-28
FLOP_PER_SAMPLE_FOLDED = 532224         # exact fold of feature_fc o latent_fc o view_fc[:, :256]
# tensor-core FLOPs the kernel actually ISSUES per sample (dense UMMA tiles incl. bias K-steps, the
# alpha/rgb rows and, in the 3-pass mode, the A_lo*W_hi and A_hi*W_lo correction passes; layer 3 takes the
# lo half of its input only on the 16-row density block)
FLOP_PER_SAMPLE_ISSUED = {"tc_fp16": 2 * 16 * (23 * 256 + 2 * 17 * 256 + 22 * 144 + 9 * 16),
                          "tc_fp16x3": 2 * 16 * (67 * 256 + 2 * 49 * 256 + 22 * 144 + 16 * 16 + 9 * 16), "fp32": 532224}
FLOP_L0_PER_KSTEP = 2 * 16 * 256        # algorithmic FLOPs of one layer-0 K-step (16 of fc_0's 352 inputs), per sample
FLOP_BEYOND_L0 = FLOP_PER_SAMPLE_FOLDED - 22 * FLOP_L0_PER_KSTEP
METRIC = "rays_per_s_512x512_64spp"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "tf_burst": d.get("bf16_tflops", 1590.0),
                "tf_sustained": d.get("bf16_tflops_sustained", 1400.0), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def pick_cpu_threads(fn):
    """The reference's PyTorch CPU path does not scale to every core of a 128-thread host (tiny per-chunk ops):
    time one call at a few thread counts and keep the fastest, so the CPU arm is not handicapped."""
    cores = os.cpu_count() or 1
    best = (None, cores)
    for nt in sorted({min(cores, c) for c in (8, 16, 32, 64, cores)}):
        torch.set_num_threads(nt)
        fn()
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        if best[0] is None or dt < best[0]:
            best = (dt, nt)
    torch.set_num_threads(best[1])
    return best[1]


def build_scene():
    from oracle import synth
    scene = synth.make_scene(H=H, W=W, scale=1.0, all_hit=True)
    assert scene["ray_o"].shape[1] == H * W
    return scene


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's CPU implementation (oracle port) on the host cores; rank 0 only."""
    if rank != 0:
        return
    from oracle import neuralbody_oracle as O
    scene = build_scene()
    n = args.ref_rays
    # a bounded, strided sample of the same 512x512 workload
    idx = torch.arange(0, H * W, (H * W) // n)[:n]
    for k in ("ray_o", "ray_d", "near", "far"):
        scene[k] = scene[k][:, idx].contiguous()
    with torch.no_grad():
        probe = dict(scene)
        for k in ("ray_o", "ray_d", "near", "far"):
            probe[k] = scene[k][:, :2048].contiguous()
        cores = pick_cpu_threads(lambda: O.render(probe, n_samples=S))
        for _ in range(args.warmup):
            O.render(scene, n_samples=S)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.render(scene, n_samples=S)
        dt = time.perf_counter() - t0
    rays_s = n * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": rays_s, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synth-313 512x512 all-hit view, 64 samples/ray, eval (BASELINE configs[1])",
                   "sample": "%d strided rays of the 262144 per step, reference chunking (2048 rays)" % n},
        "cpu_baseline": {"value": rays_s, "unit": "rays/s", "cores": cores, "kind": "port",
                         "sample": "%d rays x %d samples x %d steps, torch %s CPU, %d threads (fastest of 8/16/32/64/all)" % (
                             n, S, args.steps, torch.__version__, cores)},
        "e2e": {"value": rays_s, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ product arm
@torch.no_grad()      # inference, exactly as upstream's run.py:66 (`with torch.no_grad(): renderer.render(batch)`)
def run_product(args, rank, world, local_rank):
    import torch.distributed as dist
    from neuralbody_b200 import capi, dist as nbdist
    from neuralbody_b200.lib.config import cfg
    from neuralbody_b200.lib.networks.make_network import make_network
    from neuralbody_b200.lib.networks.renderer.make_renderer import make_renderer

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    scene = build_scene()
    lib = capi.load()
    precision = args.precision
    if precision == "auto":
        precision = "tc_fp16x3" if lib.nb_has_precision(capi.NB_PRECISION_TC_FP16X3) else "fp32"
    cfg.N_samples, cfg.perturb, cfg.white_bkgd, cfg.raw_noise_std = S, 0.0, False, 0
    cfg.render_precision, cfg.render_volume_dtype, cfg.chunk = precision, "auto", 0
    cfg.render_skip_empty = not args.dense
    cfg.render_return_weights = False     # `weights` (B,n,S) is unused downstream; rgb/depth/acc/disp are written
    cfg.num_train_frame = int(scene["weights"]["latent.weight"].shape[0])
    net = make_network(cfg)
    net.load_state_dict(scene["weights"], strict=False)
    net = net.to(dev).eval()
    net.set_feature_volume([v.to(dev) for v in scene["volumes"]])
    ren = make_renderer(cfg, net)
    ren.stats = torch.zeros(8, dtype=torch.int64, device=dev)   # tiles executed / occupied samples / decoder ns / decoder launches

    keys = ("coord", "out_sh", "bounds", "R", "Th", "latent_index", "ray_o", "ray_d", "near", "far")
    host = {k: scene[k].pin_memory() for k in keys}
    n_views = world                       # N views per step at N GPUs (weak scaling)
    full = {k: host[k].to(dev) for k in keys}
    local, per = nbdist.shard_batch(full, rank, world)
    n_local = local["ray_o"].shape[1]
    sp_input = ren.prepare_sp_input(full)
    vol = net.encode_sparse_voxels(sp_input)
    out = {k: torch.empty((1, n_local) + ((3,) if k == "rgb_map" else ()), dtype=torch.float32, device=dev)
           for k in ("rgb_map", "disp_map", "acc_map", "depth_map")}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def device_step():
        for _ in range(n_views):
            ret = ren.render_rays(local["ray_o"], local["ray_d"], local["near"], local["far"], vol, sp_input, out=out)
            if world > 1:
                nbdist.gather_slabs(nbdist.pack_slab(ret))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches0 = ren.launches
    ren.stats.zero_()
    for s0, s1 in ev:
        flush.fill_(1)                    # untimed L2 flush between timed steps
        barrier()
        s0.record()
        device_step()
        s1.record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    launches = ren.launches - launches0
    stats = [int(v) for v in ren.stats.tolist()]
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e through the public API with host buffers
    pin_rgb = torch.empty((1, n_local, 3), dtype=torch.float32).pin_memory()
    pin_depth = torch.empty((1, n_local), dtype=torch.float32).pin_memory()
    host_local = {k: v.pin_memory() for k, v in nbdist.shard_batch(host, rank, world)[0].items()
                  if torch.is_tensor(v)}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host_local.values()) * n_views
    d2h_bytes = (pin_rgb.numel() + pin_depth.numel()) * 4 * n_views

    def e2e_step():
        for _ in range(n_views):
            batch = {k: v.to(dev, non_blocking=True) for k, v in host_local.items()}
            ret = ren.render(batch)
            if world > 1:
                g = nbdist.unpack_slab(nbdist.gather_slabs(nbdist.pack_slab(ret)), H * W)
            pin_rgb.copy_(ret["rgb_map"], non_blocking=True)
            pin_depth.copy_(ret["depth_map"], non_blocking=True)
        torch.cuda.synchronize(dev)

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    # max over ranks
    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return

    rays_per_step = H * W * n_views       # whole job
    value = rays_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = rays_per_step * args.steps / (e2e_ms * 1e-3)
    peaks = load_peaks()
    # dominant kernel = the fused render kernel; its launch duration = device step time / launches per step
    # (at N=1 the step IS n_views launches of it and nothing else)
    kernel_ms = total_ms / max(1, launches) if world == 1 else None
    kernel_ms_source = "CUDA events around the step / launches per step (the step is that one kernel)"
    kernel_launches = launches
    samples_per_launch = n_local * S
    skipping = precision != "fp32" and not args.dense
    if precision != "fp32" and stats[3] > 0:
        # frame-compacting pipeline: 3 launches per view (classify, decoder, composite).  The decoder kernel is the dominant
        # one; it times itself on the device (%globaltimer: first CTA start -> last CTA end, accumulated in stats[2])
        kernel_launches = stats[3]
        kernel_ms = stats[2] * 1e-6 / stats[3]
        kernel_ms_source = "%globaltimer, first CTA start to last CTA end of render_tc_list_kernel, mean over the timed launches"
    if precision != "fp32" and kernel_launches:
        # only EXECUTED work is credited: 128-row tiles the kernel actually ran (padding rows included), per launch
        samples_per_launch = stats[0] * 128 / kernel_launches
    # layer-0 K-steps the executed tiles actually ran (a tile whose samples see only coarse levels skips the fine levels'
    # K-steps; those multiply exact zeros upstream and are NOT credited): 8 / 16 / 20 / 22 of 22 per tile
    l0_ksteps = (stats[4] / max(1, stats[0])) if (precision != "fp32" and stats[0]) else 22.0
    flop_exec = FLOP_BEYOND_L0 + l0_ksteps * FLOP_L0_PER_KSTEP
    issued = FLOP_PER_SAMPLE_ISSUED[precision]
    if precision != "fp32":
        issued -= 2 * 16 * 256 * (3 if precision == "tc_fp16x3" else 1) * (22.0 - l0_ksteps)
    if kernel_ms:
        tflops_exec = samples_per_launch * flop_exec / (kernel_ms * 1e-3) / 1e12
        tflops_written = samples_per_launch * FLOP_PER_SAMPLE_AS_WRITTEN / (kernel_ms * 1e-3) / 1e12
    else:
        tflops_exec = value * S * flop_exec / world / 1e12
        tflops_written = value * S * FLOP_PER_SAMPLE_AS_WRITTEN / world / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % precision)
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {
        "bound": "tensor", "achieved": tflops_exec, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
        "frac": tflops_exec / peaks["tf_sustained"], "traffic": traffic,
        "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % peaks["src"],
        "frac_of_burst": tflops_exec / peaks["tf_burst"],
        "flop_per_sample_executed": flop_exec, "layer0_ksteps_per_tile": l0_ksteps,
        "note": "achieved/frac count only the ALGORITHMIC folded FLOPs of executed work (532224/sample minus the layer-0 "
                "K-steps a tile skipped); precision-emulation passes, bias K-steps and padding rows the tensor pipe also "
                "executes are reported separately below",
        "tensor_flop_per_sample_issued": issued,
        "tensor_tflops_issued": (tflops_exec * issued / flop_exec),
        "tensor_issued_frac_of_sustained": (tflops_exec * issued / flop_exec) / peaks["tf_sustained"],
        "achieved_if_counted_as_written": tflops_written,
        "kernel": ("render_tc_list_kernel<%d>" % (3 if precision == "tc_fp16x3" else 1)) if precision != "fp32"
                  else "render_f32_kernel (fp32 FFMA pipe, no tensor cores)",
        "kernel_ms": kernel_ms, "kernel_ms_source": kernel_ms_source,
        "kernel_share_of_step": (kernel_ms * kernel_launches / total_ms) if kernel_ms else None,
        "samples_evaluated_per_launch": samples_per_launch, "samples_total_per_launch": n_local * S,
        "empty_sample_skipping": ("exact (sigma_empty < 0): %.1f%% of the samples occupied" % (
            100.0 * stats[1] / max(1, kernel_launches * n_local * S))) if skipping else "off (dense evaluation)",
        "hbm_compulsory_gbs": (n_local * 56 / (kernel_ms * 1e-3) / 1e9) if kernel_ms else None,
    }

    # ---- bounded CPU baseline (oracle port) on the host cores, rank 0, N=1 only
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import neuralbody_oracle as O
        nref = args.ref_rays
        idx = torch.arange(0, H * W, (H * W) // nref)[:nref]
        sub = dict(scene)
        for k in ("ray_o", "ray_d", "near", "far"):
            sub[k] = scene[k][:, idx].contiguous()
        with torch.no_grad():
            probe = dict(sub)
            for k in ("ray_o", "ray_d", "near", "far"):
                probe[k] = sub[k][:, :2048].contiguous()
            cores = pick_cpu_threads(lambda: O.render(probe, n_samples=S))
            O.render(sub, n_samples=S)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                ref = O.render(sub, n_samples=S)
            dt = time.perf_counter() - t0
        cpu_baseline = {"value": nref * reps / dt, "unit": "rays/s", "cores": cores, "kind": "port",
                        "sample": "%d strided rays x %d samples x %d reps of the same 512x512 view, torch CPU, %d threads (fastest of 8/16/32/64/all)"
                                  % (nref, S, reps, cores)}
        # free parity spot-check of the very tensors that were timed
        got = ren.render_rays(full["ray_o"][:, idx].contiguous(), full["ray_d"][:, idx].contiguous(),
                              full["near"][:, idx].contiguous(), full["far"][:, idx].contiguous(), vol, sp_input)
        cpu_baseline["parity_max_abs_rgb"] = float((got["rgb_map"].cpu() - ref["rgb_map"]).abs().max())
        cpu_baseline["parity_max_abs_depth"] = float((got["depth_map"].cpu() - ref["depth_map"]).abs().max())

    line = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"tc_fp16": "f16", "tc_fp16x3": "f16x2 (hi+lo fp16 pairs, fp32 accumulate)", "fp32": "f32"}[precision], "data": "synthetic",
        "frames_per_s_512x512": value / (H * W),
        "config": {"workload": "synth-313 512x512 all-hit view x %d per step, 64 samples/ray, eval, perturb=0 "
                               "(BASELINE configs[1])" % n_views,
                   "precision": precision, "skip_empty": (precision != "fp32" and not args.dense), "pipeline": ("fused single kernel" if precision == "fp32" else "classify -> decoder over the frame's sample list -> composite (3 launches per view)"), "rays_per_step": rays_per_step, "samples_per_ray": S,
                   "parallelism": "ray-sharded x%d (interleaved 256-ray chunks), one all-gather per view" % world if world > 1 else "single GPU",
                   "l2": "256 MiB written between timed steps (untimed) to flush the 126 MB L2",
                   "volume": "fp16 channels-last 69 MB, packed once (cached across views of the frame)"
                             if precision == "tc_fp16" else "fp32 channels-last 137 MB, packed once (cached across views)"},
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "clocks": clocks,
        "step_ms": step_ms,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "tc_fp16x3", "tc_fp16", "fp32"])
    ap.add_argument("--ref-rays", type=int, default=4096, help="rays per step of the CPU arm / baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dense", action="store_true", help="disable the exact empty-sample skipping of the tensor-core kernels")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else max(1, args.warmup)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_product(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
