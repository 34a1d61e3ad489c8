"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (zju3dv/neuralbody)
from /root/reference so that the oracle restatement (oracle/neuralbody_oracle.py)
can be pinned against it and golden vectors can be generated (oracle/make_golden.py).

/root/reference only exists in the build container, never on the GPU box: nothing in
`tests -m gpu`, `__graft_entry__.smoke()` or `bench.py` may import this module.

Recipe: three module stubs make the render path importable
on CPU without touching the reference sources:
  * `open3d`  -- imported at lib/config/config.py:1, never used on the render path;
  * `imp`     -- removed in Python 3.12; lib/networks/make_network.py:2 and
                 lib/networks/renderer/make_renderer.py:2 use `imp.load_source`;
  * `spconv`  -- lib/networks/latent_xyzc.py:2; the SparseConvNet is built but never
                 run: the dense feature volumes are supplied by the caller.
`lib.config` parses sys.argv and opens the yaml relative to CWD at import time
(lib/config/config.py:176-187), hence the chdir/argv dance below.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("NB_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "networks"))


def _install_stubs():
    import torch.nn as nn

    if "open3d" not in sys.modules:
        sys.modules["open3d"] = types.ModuleType("open3d")

    if "imp" not in sys.modules:
        imp = types.ModuleType("imp")

        def load_source(name, path):
            return importlib.machinery.SourceFileLoader(name, path).load_module()

        imp.load_source = load_source
        sys.modules["imp"] = imp

    if "spconv" not in sys.modules:
        sp = types.ModuleType("spconv")

        class _Inert(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()

        class SparseSequential(nn.Sequential):
            def __init__(self, *mods):
                super().__init__(*[m for m in mods if isinstance(m, nn.Module)])

        class SparseConvTensor:  # never constructed: volumes are supplied
            def __init__(self, *a, **k):
                raise RuntimeError("spconv is stubbed; supply dense volumes")

        sp.SubMConv3d = _Inert
        sp.SparseConv3d = _Inert
        sp.SparseSequential = SparseSequential
        sp.SparseConvTensor = SparseConvTensor
        sys.modules["spconv"] = sp


_loaded = None


def load_reference(cfg_file="configs/snapshot_exp/snapshot_f3c.yaml"):
    """Return (cfg, latent_xyzc module, if_clight_renderer module, nerf_net_utils module)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    old_cwd, old_argv = os.getcwd(), list(sys.argv)
    os.chdir(REFERENCE_ROOT)
    sys.argv = ["x", "--cfg_file", cfg_file]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        from lib.config import cfg  # noqa
        from lib.networks import latent_xyzc  # noqa
        from lib.networks.renderer import if_clight_renderer, nerf_net_utils  # noqa
    finally:
        os.chdir(old_cwd)
        sys.argv = old_argv
    _loaded = (cfg, latent_xyzc, if_clight_renderer, nerf_net_utils)
    return _loaded


def reference_render(scene, n_samples=64, perturb=0.0, training=False, white_bkgd=False,
                     t_rand=None, chunk=2048, num_train_frame=None, grad=False, masks=None):
    """Run the reference renderer on a synthetic scene dict (oracle.synth).

    Follows Renderer.render (if_clight_renderer.py:94-122) literally, except that
    net.encode_sparse_voxels (spconv) is replaced by the supplied dense volumes.
    `t_rand` (B, n, S): when given, torch.rand is patched for the duration of the
    call so the reference's own jitter line (if_clight_renderer.py:22) consumes
    exactly these numbers, chunk by chunk.
    """
    import torch
    cfg, latent_xyzc, if_clight_renderer, _ = load_reference()
    cfg.N_samples = int(n_samples)
    cfg.perturb = float(perturb)
    cfg.white_bkgd = bool(white_bkgd)
    cfg.raw_noise_std = 0
    cfg.voxel_size = [float(v) for v in scene["voxel_size"]]
    if num_train_frame is None:
        num_train_frame = int(scene["weights"]["latent.weight"].shape[0])
    cfg.num_train_frame = int(num_train_frame)

    net = latent_xyzc.Network()
    missing, unexpected = net.load_state_dict(scene["weights"], strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("xyzc_net") or k.startswith("c.") for k in missing), missing
    net.train(training)
    volumes = [v.clone().requires_grad_(grad) for v in scene["volumes"]]
    net.encode_sparse_voxels = lambda sp_input: volumes
    batch = {k: scene[k] for k in ("coord", "out_sh", "bounds", "R", "Th", "latent_index",
                                   "ray_o", "ray_d", "near", "far")}
    if masks is None:
        renderer = if_clight_renderer.Renderer(net)
    elif "R0_snap" in masks:   # f-1, single view: lib/networks/renderer/if_clight_renderer_msk.py
        from lib.networks.renderer import if_clight_renderer_msk
        cfg.H, cfg.W, cfg.ratio = int(masks["mask_H"]), int(masks["mask_W"]), 1.0
        renderer = if_clight_renderer_msk.Renderer(net)
        batch.update({k: masks[k] for k in ("R0_snap", "Th0_snap", "RT", "K", "msk")})
    else:   # f-1: lib/networks/renderer/if_clight_renderer_mmsk.py (H, W come from cfg.H * cfg.ratio)
        from lib.networks.renderer import if_clight_renderer_mmsk
        cfg.H, cfg.W, cfg.ratio = int(masks["mask_H"]), int(masks["mask_W"]), 1.0
        renderer = if_clight_renderer_mmsk.Renderer(net)
        batch.update({k: masks[k] for k in ("RT", "Ks", "msks")})

    state = {"ofs": 0}
    real_rand = torch.rand

    def fake_rand(shape, *a, **k):
        # the reference draws (B, chunk, S) per chunk
        b, n, s = tuple(shape)
        out = t_rand[:, state["ofs"]:state["ofs"] + n, :].clone()
        state["ofs"] += n
        assert out.shape == (b, n, s)
        return out

    # the reference hard-codes chunk = 2048 (if_clight_renderer.py:107)
    assert chunk == 2048
    ctx = torch.enable_grad() if grad else torch.no_grad()
    try:
        if t_rand is not None:
            torch.rand = fake_rand
        with ctx:
            ret = renderer.render(batch)
    finally:
        torch.rand = real_rand
    if grad:
        return ret, net, volumes
    return {k: v.detach() for k, v in ret.items()}


def reference_render_hierarchical(scene, n_samples=64, n_importance=128, perturb=0.0, training=False, white_bkgd=False,
                                  t_rand=None, u=None, chunk=2048, grad=False):
    """f-4 golden generator.  Neural Body has no fine pass of its own, so this composes UNMODIFIED reference
    functions exactly the way the reference's NeRF-baseline renderer does (lib/networks/renderer/volume_renderer.py:60-118):
    Renderer.get_sampling_points / get_density_color (if_clight_renderer.py:11-27,54-60) and Network.calculate_density_color
    for the network, nerf_net_utils.raw2outputs (:6-51) and nerf_net_utils.sample_pdf (:55-90) for the rest.
    `torchsearchsorted` (absent; nerf_net_utils.py:56) is stubbed with torch.searchsorted, which has the same semantics;
    torch.rand is patched so the reference's own draws consume `t_rand` (B,n,S) and `u` (B,n,n_importance)."""
    import torch
    cfg, latent_xyzc, if_clight_renderer, nerf_net_utils = load_reference()
    if "torchsearchsorted" not in sys.modules:
        tss = types.ModuleType("torchsearchsorted")
        tss.searchsorted = lambda a, v, side="left": torch.searchsorted(a, v, right=(side == "right"))
        sys.modules["torchsearchsorted"] = tss
    cfg.N_samples, cfg.perturb, cfg.white_bkgd, cfg.raw_noise_std = int(n_samples), float(perturb), bool(white_bkgd), 0
    cfg.voxel_size = [float(v) for v in scene["voxel_size"]]
    cfg.num_train_frame = int(scene["weights"]["latent.weight"].shape[0])
    net = latent_xyzc.Network()
    missing, unexpected = net.load_state_dict(scene["weights"], strict=False)
    assert not unexpected and all(k.startswith("xyzc_net") or k.startswith("c.") for k in missing)
    net.train(training)
    renderer = if_clight_renderer.Renderer(net)
    batch = {k: scene[k] for k in ("coord", "out_sh", "bounds", "R", "Th", "latent_index")}
    sp_input = renderer.prepare_sp_input(batch)
    vols = [v.clone().requires_grad_(grad) for v in scene["volumes"]]
    decoder = lambda x, vd: net.calculate_density_color(x, vd, vols, sp_input)
    draws = []
    real_rand = torch.rand

    def fake_rand(shape, *a, **k):
        out = draws.pop(0)
        assert tuple(out.shape) == tuple(shape), (tuple(out.shape), tuple(shape))
        return out.clone()

    outs = []
    assert chunk == 2048
    try:
        torch.rand = fake_rand
        with (torch.enable_grad() if grad else torch.no_grad()):
            for i in range(0, scene["ray_o"].shape[1], chunk):
                ro, rd = scene["ray_o"][:, i:i + chunk], scene["ray_d"][:, i:i + chunk]
                near, far = scene["near"][:, i:i + chunk], scene["far"][:, i:i + chunk]
                B, n = ro.shape[:2]
                if perturb > 0 and training:
                    draws.append(t_rand[:, i:i + chunk])
                wpts, z_vals = renderer.get_sampling_points(ro, rd, near, far)
                viewdir = rd / torch.norm(rd, dim=2, keepdim=True)
                raw = renderer.get_density_color(wpts, viewdir, decoder).reshape(-1, n_samples, 4)
                z_vals = z_vals.view(-1, n_samples)
                rays_d = rd.reshape(-1, 3)
                rgb0, disp0, acc0, weights, _ = nerf_net_utils.raw2outputs(raw, z_vals, rays_d, cfg.raw_noise_std, cfg.white_bkgd)
                # volume_renderer.py:84-93
                z_vals_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
                if perturb != 0.:
                    draws.append(u[:, i:i + chunk].reshape(B * n, n_importance))
                z_samples = nerf_net_utils.sample_pdf(z_vals_mid, weights[..., 1:-1], n_importance, det=(perturb == 0.)).detach()
                z_all, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
                S2 = n_samples + n_importance
                pts = ro[:, :, None] + rd[:, :, None] * z_all.view(B, n, S2)[..., None]
                raw = renderer.get_density_color(pts, viewdir, decoder).reshape(-1, S2, 4)
                rgb, disp, acc, w2, depth = nerf_net_utils.raw2outputs(raw, z_all, rays_d, cfg.raw_noise_std, cfg.white_bkgd)
                outs.append({"rgb_map": rgb.view(B, n, 3), "disp_map": disp.view(B, n), "acc_map": acc.view(B, n),
                             "weights": w2.view(B, n, S2), "depth_map": depth.view(B, n), "rgb0": rgb0.view(B, n, 3),
                             "disp0": disp0.view(B, n), "acc0": acc0.view(B, n),
                             "z_std": torch.std(z_samples, dim=-1, unbiased=False).view(B, n), "z_vals": z_all.view(B, n, S2)})
    finally:
        torch.rand = real_rand
    assert not draws
    ret = {k: torch.cat([o[k] for o in outs], dim=1) for k in outs[0]}
    if grad:
        return ret, net, vols
    return {k: v.detach() for k, v in ret.items()}
