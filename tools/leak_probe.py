"""Diagnostics (GPU): which objects of a training step only the cyclic garbage collector can free."""
import gc
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402


def main():
    from oracle import synth
    from neuralbody_b200.lib.config import cfg
    import gpu_utils as G
    scene = synth.make_scene(H=64, W=64, scale=0.5, all_hit=True)
    net, ren = G.make_net_and_renderer(scene)
    batch = {k: scene[k].cuda() for k in G.BATCH_KEYS}
    n = min(1024, batch["ray_o"].shape[1])
    for k in ("ray_o", "ray_d", "near", "far"):
        batch[k] = batch[k][:, :n].contiguous()
    cfg.N_samples, cfg.perturb, cfg.white_bkgd, cfg.raw_noise_std, cfg.chunk = 64, 1.0, False, 0, 0
    cfg.render_importance = int(os.environ.get("NI", "128"))
    net.train()
    vols = [v.cuda().requires_grad_(True) for v in scene["volumes"]]
    net.set_feature_volume(vols)
    sp = ren.prepare_sp_input(batch)

    def step():
        for p in net.parameters():
            p.grad = None
        for v in vols:
            v.grad = None
        out = ren.get_pixel_value(batch["ray_o"], batch["ray_d"], batch["near"], batch["far"], vols, sp, batch)
        loss = (out["rgb_map"] ** 2).mean() + ((out["rgb0"] ** 2).mean() if "rgb0" in out else 0.)
        loss.backward()

    for _ in range(2):
        step()
    gc.collect()
    gc.disable()
    gc.set_debug(gc.DEBUG_SAVEALL)
    a0 = torch.cuda.memory_allocated()
    step()
    a1 = torch.cuda.memory_allocated()
    step()
    a2 = torch.cuda.memory_allocated()
    n_found = gc.collect()
    print("allocated after steps: %d -> %d -> %d bytes; gc found %d objects" % (a0, a1, a2, n_found))
    kinds = {}
    for o in gc.garbage:
        kinds[type(o).__name__] = kinds.get(type(o).__name__, 0) + 1
    print(sorted(kinds.items(), key=lambda kv: -kv[1])[:20])
    shown = 0
    for o in gc.garbage:
        if isinstance(o, torch.Tensor) and shown < 12:
            shown += 1
            refs = [type(r).__name__ + (":" + ",".join(k for k, v in r.items() if v is o) if isinstance(r, dict) else "")
                    for r in gc.get_referrers(o) if r is not gc.garbage]
            print("tensor", tuple(o.shape), o.dtype, "grad_fn", type(o.grad_fn).__name__ if o.grad_fn is not None else None, "<-", refs[:6])
        elif isinstance(o, dict) and shown < 40 and ("save" in o or "args" in o or "rgb_map" in o):
            shown += 1
            print("dict keys", list(o.keys())[:30])


if __name__ == "__main__":
    main()
