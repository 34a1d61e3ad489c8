"""Diagnostics (GPU): f-2(ii) end to end -- synth-313 vertices -> dense-PyTorch SparseConvNet emulation -> pack -> render,
i.e. Renderer.render(batch) with nothing supplied but the reference's batch dict.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    from oracle import synth
    from neuralbody_b200.lib.config import cfg
    from neuralbody_b200.lib.networks.make_network import make_network
    from neuralbody_b200.lib.networks.renderer.make_renderer import make_renderer
    import gpu_utils as G
    scene = synth.make_scene(H=128, W=128, scale=1.0, all_hit=True)
    cfg.N_samples, cfg.perturb, cfg.white_bkgd, cfg.raw_noise_std, cfg.chunk = 64, 0.0, False, 0, 0
    cfg.render_precision, cfg.render_importance = "tc_fp16x3", 0
    cfg.num_train_frame = int(scene["weights"]["latent.weight"].shape[0])
    net = make_network(cfg)
    net.load_state_dict(scene["weights"], strict=False)
    net = net.cuda()
    net.attach_dense_encoder()
    net.train()                                   # upstream renders with network.train() (BatchNorm batch statistics)
    ren = make_renderer(cfg, net)
    batch = {k: scene[k].cuda() for k in G.BATCH_KEYS}
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with torch.no_grad():
        sp = ren.prepare_sp_input(batch)
        e[0].record()
        vols = net.encode_sparse_voxels(sp)
        e[1].record()
        out = ren.render(batch)
        e[2].record()
    torch.cuda.synchronize()
    print(json.dumps({"encode_ms": e[0].elapsed_time(e[1]), "encode_plus_render_ms": e[1].elapsed_time(e[2]),
                      "volume_shapes": [list(v.shape) for v in vols],
                      "active_fraction": [float((v.abs().sum(1) > 0).float().mean()) for v in vols],
                      "acc_mean": float(out["acc_map"].mean()), "finite": bool(torch.isfinite(out["rgb_map"]).all()),
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
