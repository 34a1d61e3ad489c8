"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU fp32 restatement of neuralbody's volumetric-render hot path, function by
function, each citing the reference file:line it follows (paths relative to the
reference tree, commit 3c516b9).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` leg may import this module; the
product path (neuralbody_b200/) never does and fails loudly without its CUDA
library.

Parity pinning: the reference ships no tests or golden vectors.
This restatement is pinned instead against outputs of the UNMODIFIED reference
executed in the build container (oracle/ref_harness.py) on the seeded synthetic
scenes of oracle/synth.py; the vectors are committed under tests/golden/
together with the generating script oracle/make_golden.py, and
tests/test_oracle.py re-checks the restatement against them on every run.

Torch is used as the array library (F.grid_sample / conv1d are the third-party
arithmetic the reference itself calls); nothing here touches CUDA.
"""
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ a2 sampling
def get_sampling_points(ray_o, ray_d, near, far, n_samples, perturb=0.0, training=False, t_rand=None):
    """lib/networks/renderer/if_clight_renderer.py:11-27.
    ray_o/ray_d (B,n,3), near/far (B,n) -> pts (B,n,S,3), z_vals (B,n,S).
    `t_rand` stands for the `torch.rand(z_vals.shape)` draw at :22."""
    t_vals = torch.linspace(0., 1., steps=n_samples).to(near)
    z_vals = near[..., None] * (1. - t_vals) + far[..., None] * t_vals
    if perturb > 0. and training:
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        if t_rand is None:
            t_rand = torch.rand(z_vals.shape)
        z_vals = lower + (upper - lower) * t_rand.to(upper)
    pts = ray_o[:, :, None] + ray_d[:, :, None] * z_vals[..., None]
    return pts, z_vals


# ------------------------------------------------------------------ a5 / a6 / a7
def pts_to_can_pts(pts, R, Th):
    """lib/networks/latent_xyzc.py:41-47: (p - Th) @ R.  Th is (B,1,3) or (B,3)."""
    Th = Th.reshape(Th.shape[0], 1, 3)
    return torch.matmul(pts - Th, R)


def get_grid_coords(pts, bounds, out_sh, voxel_size):
    """lib/networks/latent_xyzc.py:49-60 (divides by out_sh, not out_sh-1; xyz order out)."""
    dhw = pts[..., [2, 1, 0]]
    min_dhw = bounds[:, 0, [2, 1, 0]]
    dhw = dhw - min_dhw[:, None]
    dhw = dhw / torch.tensor(voxel_size).to(dhw)
    out_sh = torch.tensor(out_sh).to(dhw)
    dhw = dhw / out_sh * 2 - 1
    return dhw[..., [2, 1, 0]]


def interpolate_features(grid_coords, feature_volume):
    """lib/networks/latent_xyzc.py:62-72: 4x trilinear grid_sample, zeros padding,
    align_corners=True; channel order [L1:32, L2:64, L3:128, L4:128] -> (B,352,P)."""
    g = grid_coords[:, None, None]
    feats = [F.grid_sample(v, g, padding_mode='zeros', align_corners=True) for v in feature_volume]
    feats = torch.cat(feats, dim=1)
    return feats.view(feats.size(0), -1, feats.size(4))


# ------------------------------------------------------------------ a9 embedder
def positional_embed(x, multires):
    """lib/networks/embedder.py:5-50: [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]."""
    freq_bands = 2. ** torch.linspace(0., multires - 1, steps=multires)
    out = [x]
    for freq in freq_bands:
        out.append(torch.sin(x * freq))
        out.append(torch.cos(x * freq))
    return torch.cat(out, -1)


# ------------------------------------------------------------------ a8 decoder
def _conv(w, name, x):
    return F.conv1d(x, w[name + ".weight"], w[name + ".bias"])


def calculate_density_color(w, wpts, viewdir, feature_volume, sp_input, voxel_size, xyz_res=10, view_res=4):
    """lib/networks/latent_xyzc.py:91-126. wpts, viewdir (B,P,3) -> raw (B,P,4) = (rgb logits, sigma)."""
    ppts = pts_to_can_pts(wpts, sp_input['R'], sp_input['Th'])
    grid_coords = get_grid_coords(ppts, sp_input['bounds'], sp_input['out_sh'], voxel_size)
    xyzc_features = interpolate_features(grid_coords, feature_volume)

    net = F.relu(_conv(w, "fc_0", xyzc_features))
    net = F.relu(_conv(w, "fc_1", net))
    net = F.relu(_conv(w, "fc_2", net))
    alpha = _conv(w, "alpha_fc", net)

    features = _conv(w, "feature_fc", net)
    latent = w["latent.weight"][sp_input['latent_index']]
    latent = latent[..., None].expand(*latent.shape, net.size(2))
    features = torch.cat((features, latent), dim=1)
    features = _conv(w, "latent_fc", features)

    vd = positional_embed(viewdir, view_res).transpose(1, 2)
    light_pts = positional_embed(wpts, xyz_res).transpose(1, 2)
    features = torch.cat((features, vd, light_pts), dim=1)
    net = F.relu(_conv(w, "view_fc", features))
    rgb = _conv(w, "rgb_fc", net)
    raw = torch.cat((rgb, alpha), dim=1)
    return raw.transpose(1, 2)


def calculate_density(w, wpts, feature_volume, sp_input, voxel_size):
    """lib/networks/latent_xyzc.py:74-89 (f-3, the mesh renderer's alpha decoder): (B,P,3) -> (B,P,1)."""
    ppts = pts_to_can_pts(wpts, sp_input['R'], sp_input['Th'])
    grid_coords = get_grid_coords(ppts, sp_input['bounds'], sp_input['out_sh'], voxel_size)
    xyzc_features = interpolate_features(grid_coords, feature_volume)
    net = F.relu(_conv(w, "fc_0", xyzc_features))
    net = F.relu(_conv(w, "fc_1", net))
    net = F.relu(_conv(w, "fc_2", net))
    return _conv(w, "alpha_fc", net).transpose(1, 2)


# ------------------------------------------------------------------ a10 composite
def raw2outputs(raw, z_vals, rays_d, white_bkgd=False):
    """lib/networks/renderer/nerf_net_utils.py:6-51 (raw_noise_std = 0 as in every config)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    dists = torch.cat([dists, torch.Tensor([1e10]).expand(dists[..., :1].shape).to(dists)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3])
    alpha = 1. - torch.exp(-F.relu(raw[..., 3]) * dists)
    weights = alpha * torch.cumprod(
        torch.cat([torch.ones((alpha.shape[0], 1)).to(alpha), 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
    acc_map = torch.sum(weights, -1)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, disp_map, acc_map, weights, depth_map


# ------------------------------------------------------------------ a3 / a4 / a1 driver
def prepare_sp_input(batch):
    """lib/networks/renderer/if_clight_renderer.py:29-52 (coord concat omitted: it only
    feeds spconv, which stays on the reference path)."""
    out_sh, _ = torch.max(batch['out_sh'], dim=0)
    return {'out_sh': out_sh.tolist(), 'batch_size': batch['coord'].shape[0], 'bounds': batch['bounds'],
            'R': batch['R'], 'Th': batch['Th'], 'latent_index': batch['latent_index']}


def get_pixel_value(w, ray_o, ray_d, near, far, feature_volume, sp_input, voxel_size, n_samples,
                    perturb=0.0, training=False, white_bkgd=False, t_rand=None):
    """lib/networks/renderer/if_clight_renderer.py:62-92 (+ get_density_color :54-60)."""
    wpts, z_vals = get_sampling_points(ray_o, ray_d, near, far, n_samples, perturb, training, t_rand)
    viewdir = ray_d / torch.norm(ray_d, dim=2, keepdim=True)
    n_batch, n_pixel, n_sample = wpts.shape[:3]
    wpts_flat = wpts.view(n_batch, n_pixel * n_sample, -1)
    vd = viewdir[:, :, None].repeat(1, 1, n_sample, 1).contiguous().view(n_batch, n_pixel * n_sample, -1)
    raw = calculate_density_color(w, wpts_flat, vd, feature_volume, sp_input, voxel_size)
    raw = raw.reshape(-1, n_sample, 4)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(
        raw, z_vals.view(-1, n_sample), ray_d.reshape(-1, 3), white_bkgd)
    return {'rgb_map': rgb_map.view(n_batch, n_pixel, -1), 'disp_map': disp_map.view(n_batch, n_pixel),
            'acc_map': acc_map.view(n_batch, n_pixel), 'weights': weights.view(n_batch, n_pixel, -1),
            'depth_map': depth_map.view(n_batch, n_pixel)}


# ------------------------------------------------------------------ f-1: masked renderer (vis_novel_view / vis_novel_pose)
def prepare_inside_pts(pts, batch, H, W):
    """lib/networks/renderer/if_clight_renderer_mmsk.py:12-45.  pts (1,n,S,3) -> inside (1, n*S) bool:
    a sample is kept only if it projects into the foreground mask of EVERY training view
    (H, W = int(cfg.H * cfg.ratio), int(cfg.W * cfg.ratio) upstream; B = 1 only)."""
    sh = pts.shape
    pts = pts.view(sh[0], -1, sh[3])
    inside = None
    for nv in range(batch['Ks'].size(1)):
        R = batch['RT'][:, nv, :3, :3]
        T = batch['RT'][:, nv, :3, 3]
        pts_ = torch.matmul(pts, R.transpose(2, 1)) + T[:, None]
        pts_ = torch.matmul(pts_, batch['Ks'][:, nv].transpose(2, 1))
        pts2d = pts_[..., :2] / pts_[..., 2:]
        pts2d = pts2d.round().long()
        pts2d[..., 0] = torch.clamp(pts2d[..., 0], 0, W - 1)
        pts2d[..., 1] = torch.clamp(pts2d[..., 1], 0, H - 1)
        pts2d = pts2d[0]
        msk = batch['msks'][0, nv]
        ins = msk[pts2d[:, 1], pts2d[:, 0]][None].bool()
        inside = ins if inside is None else inside * ins
    return inside


def get_pixel_value_mmsk(w, ray_o, ray_d, near, far, feature_volume, sp_input, voxel_size, n_samples, masks,
                         white_bkgd=False):
    """if_clight_renderer_mmsk.py:47-94: decoder only on inside samples, raw = 0 elsewhere."""
    wpts, z_vals = get_sampling_points(ray_o, ray_d, near, far, n_samples)
    if "R0_snap" in masks:   # if_clight_renderer_msk.Renderer overrides prepare_inside_pts, nothing else
        inside = prepare_inside_pts_msk(wpts, masks, masks["mask_H"], masks["mask_W"])
    else:
        inside = prepare_inside_pts(wpts, masks, masks["mask_H"], masks["mask_W"])
    viewdir = ray_d / torch.norm(ray_d, dim=2, keepdim=True)
    n_batch, n_pixel, n_sample = wpts.shape[:3]
    wp = wpts.view(n_batch, n_pixel * n_sample, -1)
    vd = viewdir[:, :, None].repeat(1, 1, n_sample, 1).contiguous().view(n_batch, n_pixel * n_sample, -1)
    full_raw = torch.zeros([n_batch, n_pixel * n_sample, 4]).to(wp)
    if inside.sum() > 0:
        raw = calculate_density_color(w, wp[inside][None], vd[inside][None], feature_volume, sp_input, voxel_size)
        full_raw[inside] = raw[0]
    raw = full_raw.reshape(-1, n_sample, 4)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals.view(-1, n_sample), ray_d.reshape(-1, 3), white_bkgd)
    return {'rgb_map': rgb_map.view(n_batch, n_pixel, -1), 'disp_map': disp_map.view(n_batch, n_pixel),
            'acc_map': acc_map.view(n_batch, n_pixel), 'weights': weights.view(n_batch, n_pixel, -1),
            'depth_map': depth_map.view(n_batch, n_pixel)}


def prepare_inside_pts_msk(wpts, batch, H, W):
    """lib/networks/renderer/if_clight_renderer_msk.py:12-49 (single-view variant of the People-Snapshot demos): world ->
    SMPL frame with the rendered frame's (R, Th) -> world of the snapshot frame (R0_snap, Th0_snap) -> that frame's camera
    (RT (1,3,4), K (1,3,3)) -> foreground test in msk (1,H,W).  wpts (1,n,S,3) -> inside (1, n*S) bool."""
    Th = batch['Th']
    can_pts = wpts - Th[:, None, None]
    can_pts = torch.matmul(can_pts, batch['R'])
    R0 = batch['R0_snap']
    Th0 = batch['Th0_snap']
    sh = can_pts.shape
    can_pts = can_pts.view(sh[0], -1, sh[3])
    pts = torch.matmul(can_pts, R0.transpose(2, 1)) + Th0[:, None]
    R = batch['RT'][..., :3]
    T = batch['RT'][..., 3]
    pts = torch.matmul(pts, R.transpose(2, 1)) + T[:, None]
    pts = torch.matmul(pts, batch['K'].transpose(2, 1))
    pts2d = pts[..., :2] / pts[..., 2:]
    pts2d = pts2d.round().long()
    pts2d[..., 0] = torch.clamp(pts2d[..., 0], 0, W - 1)
    pts2d[..., 1] = torch.clamp(pts2d[..., 1], 0, H - 1)
    pts2d = pts2d[0]
    msk = batch['msk'][0]
    return msk[pts2d[:, 1], pts2d[:, 0]][None].bool()


def render_mmsk(scene, masks, n_samples=64, white_bkgd=False, chunk=2048):
    """Renderer.render (if_clight_renderer.py:94-122) driving the masked get_pixel_value."""
    sp_input = prepare_sp_input(scene)
    n_pixel = scene['ray_o'].shape[1]
    if "R0_snap" in masks:
        masks = dict(masks, R=scene['R'], Th=scene['Th'])
    ret_list = []
    for i in range(0, n_pixel, chunk):
        ret_list.append(get_pixel_value_mmsk(
            scene['weights'], scene['ray_o'][:, i:i + chunk], scene['ray_d'][:, i:i + chunk], scene['near'][:, i:i + chunk],
            scene['far'][:, i:i + chunk], scene['volumes'], sp_input, scene['voxel_size'], n_samples, masks, white_bkgd))
    return {k: torch.cat([r[k] for r in ret_list], dim=1) for k in ret_list[0]}


def render(scene, n_samples=64, perturb=0.0, training=False, white_bkgd=False, t_rand=None, chunk=2048,
           max_rays=None):
    """lib/networks/renderer/if_clight_renderer.py:94-122 with the dense volumes supplied
    (scene['volumes']) in place of net.encode_sparse_voxels."""
    ray_o, ray_d, near, far = scene['ray_o'], scene['ray_d'], scene['near'], scene['far']
    if max_rays is not None:
        ray_o, ray_d, near, far = ray_o[:, :max_rays], ray_d[:, :max_rays], near[:, :max_rays], far[:, :max_rays]
    sp_input = prepare_sp_input(scene)
    n_pixel = ray_o.shape[1]
    ret_list = []
    for i in range(0, n_pixel, chunk):
        tr = None if t_rand is None else t_rand[:, i:i + chunk]
        ret_list.append(get_pixel_value(
            scene['weights'], ray_o[:, i:i + chunk], ray_d[:, i:i + chunk], near[:, i:i + chunk],
            far[:, i:i + chunk], scene['volumes'], sp_input, scene['voxel_size'], n_samples,
            perturb, training, white_bkgd, tr))
    return {k: torch.cat([r[k] for r in ret_list], dim=1) for k in ret_list[0]}


# ------------------------------------------------------------------ f-4: hierarchical (coarse + importance) sampling
# Neural Body's own renderer has no fine pass (`N_importance` is a dead key for it); the spec is the
# reference's NeRF-baseline renderer, whose pieces are restated here and composed with the Neural Body decoder.
def sample_pdf(bins, weights, n_importance, det=False, u=None):
    """Inverse-CDF sampling of lib/networks/renderer/nerf_net_utils.py:55-90, same fp32 operations in the same order.
    bins (N, M) interval mid-points, weights (N, M-1) -> samples (N, n_importance).
      :59-63  pdf = (w + 1e-5) / sum, cdf = [0, cumsum(pdf)]                         (N, M)
      :66-70  u = linspace(0, 1, n) if det else rand(N, n)   (`u` stands for that draw)
      :74-77  hi = searchsorted(cdf, u, side='right') (torchsearchsorted == torch.searchsorted(right=True)),
              lo = max(hi - 1, 0), hi = min(hi, M - 1)
      :82-88  t = (u - cdf[lo]) / (cdf[hi] - cdf[lo], or 1 where that is < 1e-5); sample = bins[lo] + t * (bins[hi] - bins[lo])"""
    w = weights + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    lead = list(cdf.shape[:-1])
    if det:
        u = torch.linspace(0., 1., steps=n_importance).to(cdf).expand(lead + [n_importance])
    elif u is None:
        u = torch.rand(lead + [n_importance]).to(cdf)
    u = u.contiguous()
    hi = torch.searchsorted(cdf, u, right=True)
    lo = torch.clamp(hi - 1, min=0)
    hi = torch.clamp(hi, max=cdf.shape[-1] - 1)
    c_lo, c_hi = torch.gather(cdf, -1, lo), torch.gather(cdf, -1, hi)
    b_lo, b_hi = torch.gather(bins, -1, lo), torch.gather(bins, -1, hi)
    span = c_hi - c_lo
    span = torch.where(span < 1e-5, torch.ones_like(span), span)
    t = (u - c_lo) / span
    return b_lo + t * (b_hi - b_lo)


def importance_z_vals(z_vals, weights, n_importance, det=False, u=None):
    """lib/networks/renderer/volume_renderer.py:84-93: z_vals, weights (N, S) -> (sorted (N, S + n_importance), samples)."""
    z_vals_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
    z_samples = sample_pdf(z_vals_mid, weights[..., 1:-1], n_importance, det=det, u=u).detach()
    z_all, _ = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
    return z_all, z_samples


def get_pixel_value_at(w, ray_o, ray_d, z_vals, feature_volume, sp_input, voxel_size, white_bkgd=False):
    """get_pixel_value (if_clight_renderer.py:62-92) from given depths: pts = ray_o + ray_d * z (:25), decoder, raw2outputs."""
    wpts = ray_o[:, :, None] + ray_d[:, :, None] * z_vals[..., None]
    viewdir = ray_d / torch.norm(ray_d, dim=2, keepdim=True)
    n_batch, n_pixel, n_sample = wpts.shape[:3]
    vd = viewdir[:, :, None].repeat(1, 1, n_sample, 1).contiguous().view(n_batch, n_pixel * n_sample, -1)
    raw = calculate_density_color(w, wpts.view(n_batch, n_pixel * n_sample, -1), vd, feature_volume, sp_input, voxel_size)
    raw = raw.reshape(-1, n_sample, 4)
    rgb_map, disp_map, acc_map, weights, depth_map = raw2outputs(raw, z_vals.view(-1, n_sample), ray_d.reshape(-1, 3), white_bkgd)
    return {'rgb_map': rgb_map.view(n_batch, n_pixel, -1), 'disp_map': disp_map.view(n_batch, n_pixel),
            'acc_map': acc_map.view(n_batch, n_pixel), 'weights': weights.view(n_batch, n_pixel, -1),
            'depth_map': depth_map.view(n_batch, n_pixel)}


def render_hierarchical(scene, n_samples=64, n_importance=128, perturb=0.0, training=False, white_bkgd=False,
                        t_rand=None, u=None, chunk=2048):
    """volume_renderer.py:60-118 with the Neural Body decoder in place of `self.net`: coarse pass, sample_pdf on its
    weights (det = (perturb == 0)), sort-merge, fine pass over the S + n_importance depths with the SAME network.
    Returns the fine maps plus rgb0 / disp0 / acc0 / z_std like the reference dict (:105-118); `u` (B,n,n_importance)."""
    sp_input = prepare_sp_input(scene)
    w, vols, vs = scene['weights'], scene['volumes'], scene['voxel_size']
    outs = []
    for i in range(0, scene['ray_o'].shape[1], chunk):
        ro, rd = scene['ray_o'][:, i:i + chunk], scene['ray_d'][:, i:i + chunk]
        near, far = scene['near'][:, i:i + chunk], scene['far'][:, i:i + chunk]
        tr = None if t_rand is None else t_rand[:, i:i + chunk]
        _, z_vals = get_sampling_points(ro, rd, near, far, n_samples, perturb, training, tr)
        coarse = get_pixel_value_at(w, ro, rd, z_vals, vols, sp_input, vs, white_bkgd)
        B, n = z_vals.shape[:2]
        uu = None if u is None else u[:, i:i + chunk].reshape(B * n, -1)
        z_all, z_samples = importance_z_vals(z_vals.view(B * n, -1), coarse['weights'].view(B * n, -1), n_importance,
                                             det=(perturb == 0.), u=uu)
        fine = get_pixel_value_at(w, ro, rd, z_all.view(B, n, -1), vols, sp_input, vs, white_bkgd)
        fine.update({'rgb0': coarse['rgb_map'], 'disp0': coarse['disp_map'], 'acc0': coarse['acc_map'],
                     'z_std': torch.std(z_samples, dim=-1, unbiased=False).view(B, n), 'z_vals': z_all.view(B, n, -1)})
        outs.append(fine)
    return {k: torch.cat([r[k] for r in outs], dim=1) for k in outs[0]}
