"""Diagnostics (GPU): BASELINE config 3 -- one N_rand = 1024 training chunk on the synth-313 body, 64 coarse samples +
128 importance samples (cfg.render_importance), gradient path on: forward (exact fp32 kernels with activation record,
coarse + fine) + nb_sample_pdf + backward through both passes.  Prints one JSON line (rays/s for fwd+bwd).
Usage: python tools/bench_train_chunk.py [n_importance=128] [iters=30]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    ni = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    from oracle import synth
    from neuralbody_b200.lib.config import cfg
    import gpu_utils as G
    scene = synth.make_scene(H=512, W=512, scale=1.0, all_hit=True)
    g = torch.Generator().manual_seed(0)
    idx = torch.randperm(scene["ray_o"].shape[1], generator=g)[:1024]
    for k in ("ray_o", "ray_d", "near", "far"):
        scene[k] = scene[k][:, idx].contiguous()
    net, ren = G.make_net_and_renderer(scene)
    cfg.N_samples, cfg.perturb, cfg.white_bkgd, cfg.raw_noise_std, cfg.chunk = 64, 1.0, False, 0, 0
    cfg.render_precision, cfg.render_volume_dtype, cfg.render_importance = "tc_fp16x3", "auto", ni
    net.train()
    vols = [v.cuda().requires_grad_(True) for v in scene["volumes"]]
    net.set_feature_volume(vols)
    batch = {k: scene[k].cuda() for k in G.BATCH_KEYS}
    sp = ren.prepare_sp_input(batch)
    target = torch.rand((1, 1024, 3), device="cuda")

    def step():
        for p in net.parameters():
            p.grad = None
        for v in vols:
            v.grad = None
        out = ren.get_pixel_value(batch["ray_o"], batch["ray_d"], batch["near"], batch["far"], vols, sp, batch)
        loss = ((out["rgb_map"] - target) ** 2).mean()
        if "rgb0" in out:
            loss = loss + ((out["rgb0"] - target) ** 2).mean()          # img_loss0, if_nerf_clight.py:29-32
        loss.backward()
        return float(loss.detach()) if False else None

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(json.dumps({"config": "c3: 1024-ray training chunk, 64 + %d samples, fwd + bwd, exact fp32 kernels" % ni,
                      "ms_per_step": ms, "rays_per_s_fwd_bwd": 1024 / (ms * 1e-3),
                      "grad_norm_fc0": float(dict(net.named_parameters())["fc_0.weight"].grad.norm())}))


if __name__ == "__main__":
    main()
