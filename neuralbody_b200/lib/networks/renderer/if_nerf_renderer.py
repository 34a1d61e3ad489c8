"""Drop-in for the reference's Neural Body renderer.

Replaces lib/networks/renderer/if_clight_renderer.py (the file every Neural Body config
selects through `renderer_module/renderer_path`; BASELINE.json calls it
if_nerf_renderer.py): same class name, constructor and `render(batch)` contract
(:94-122), same five output keys/shapes/dtypes (:84-92), same config keys read inside
(`N_samples, perturb, raw_noise_std, white_bkgd` :13,16,82; `voxel_size`
latent_xyzc.py:54).  The body of the per-chunk loop -- get_sampling_points ->
get_density_color -> Network.calculate_density_color -> raw2outputs -- is ONE fused
CUDA launch through the C ABI (include/neuralbody_b200.h); no PyTorch op runs inside
the ray loop and there is no CPU/eager fallback.
"""
import ctypes as C

import torch

from neuralbody_b200 import capi
from neuralbody_b200.lib.config import get_active_cfg

_PRECISIONS = {"fp32": capi.NB_PRECISION_FP32, "tc_fp16": capi.NB_PRECISION_TC_FP16,
               "tc_fp16x3": capi.NB_PRECISION_TC_FP16X3}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f32c(t, device):
    return t.to(device=device, dtype=torch.float32).contiguous()


class _FusedRender(torch.autograd.Function):
    """Autograd boundary of the training path (BASELINE config 3): forward = nb_render_fwd with the exact
    kernel + activation record, backward = nb_render_bwd.  Differentiable outputs: rgb_map, depth_map,
    acc_map (what lib/train/trainers/if_nerf_clight.py:25-32 and depth/mask losses consume); disp_map and
    weights are returned detached.  Differentiable inputs: the four dense volumes (so gradients keep flowing
    into the reference's SparseConvNet / code embedding) and the 17 decoder tensors."""

    @staticmethod
    def forward(ctx, renderer, call, *tensors):
        out = renderer._launch(call, save=True)
        ctx.renderer, ctx.call = renderer, call
        ctx.n_vol = len(call["feature_volume"])
        ctx.save_for_backward(*tensors)
        ctx.mark_non_differentiable(out["disp_map"], out["weights"])
        return out["rgb_map"], out["disp_map"], out["acc_map"], out["depth_map"], out["weights"]

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_depth, d_weights):
        grads = ctx.renderer._launch_bwd(ctx.call, d_rgb, d_depth, d_acc, ctx.needs_input_grad[2:])
        return (None, None) + tuple(grads)


class Renderer:
    def __init__(self, net):
        self.net = net
        self.lib = capi.load()
        self._vol_key = None
        self._vol_blob = None
        self._vol_dims = None
        self._vol_dtype = None
        self._w_key = None
        self._w_blob = None
        self._vol_keep = None
        self._w_keep = None
        self._tvals = {}
        self.launches = 0          # render kernels enqueued so far (bench accounting)

    # ------------------------------------------------------------------ options
    def _opt(self, name, default):
        cfg = get_active_cfg()
        return cfg[name] if name in cfg else default

    def _precision(self):
        name = str(self._opt("render_precision", "tc_fp16x3"))
        if name not in _PRECISIONS:
            raise ValueError("cfg.render_precision must be one of %s" % sorted(_PRECISIONS))
        return _PRECISIONS[name]

    def _train_precision(self, B, n, S):
        """Precision of a call autograd records: 'tc_tf32x3' (default: sample list + tcgen05 TF32 GEMM chains, exact empty-sample
        skipping in forward and backward) or 'fp32' (the exact FFMA kernels)."""
        name = str(self._opt("render_train_precision", "tc_tf32x3"))
        if name not in ("tc_tf32x3", "fp32"):
            raise ValueError("cfg.render_train_precision must be 'tc_tf32x3' or 'fp32'")
        if name == "tc_tf32x3" and S <= 256 and B * n * S < (1 << 28) and self.lib.nb_has_precision(capi.NB_PRECISION_TC_TF32X3):
            return capi.NB_PRECISION_TC_TF32X3
        return capi.NB_PRECISION_FP32

    def _volume_dtype(self, precision):
        name = str(self._opt("render_volume_dtype", "auto"))
        if name == "auto":
            # the fp16 volume alone costs ~1e-3 of depth_map parity: only the 1-pass mode gathers from it
            return capi.NB_DTYPE_F16 if precision == capi.NB_PRECISION_TC_FP16 else capi.NB_DTYPE_F32
        return {"fp32": capi.NB_DTYPE_F32, "fp16": capi.NB_DTYPE_F16}[name]

    # ------------------------------------------------------------------ a2 (host API parity)
    def get_sampling_points(self, ray_o, ray_d, near, far, t_rand=None):
        """if_clight_renderer.py:11-27, kept callable for subclasses.  `render` does not call
        this: the fused kernel generates the same samples in registers."""
        cfg = get_active_cfg()
        t_vals = torch.linspace(0., 1., steps=cfg.N_samples).to(near)
        z_vals = near[..., None] * (1. - t_vals) + far[..., None] * t_vals
        if cfg.perturb > 0. and self.net.training:
            mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
            upper = torch.cat([mids, z_vals[..., -1:]], -1)
            lower = torch.cat([z_vals[..., :1], mids], -1)
            if t_rand is None:
                t_rand = torch.rand(z_vals.shape)
            z_vals = lower + (upper - lower) * t_rand.to(upper)
        pts = ray_o[:, :, None] + ray_d[:, :, None] * z_vals[..., None]
        return pts, z_vals

    # ------------------------------------------------------------------ a3
    def prepare_sp_input(self, batch):
        """if_clight_renderer.py:29-52, unchanged semantics (stays Python; `.tolist()` is the
        one host sync per frame, exactly as upstream)."""
        sp_input = {}
        sh = batch['coord'].shape
        idx = [torch.full([sh[1]], i) for i in range(sh[0])]
        idx = torch.cat(idx).to(batch['coord'])
        coord = batch['coord'].view(-1, sh[-1])
        sp_input['coord'] = torch.cat([idx[:, None], coord], dim=1)
        out_sh, _ = torch.max(batch['out_sh'], dim=0)
        sp_input['out_sh'] = out_sh.tolist()
        sp_input['batch_size'] = sh[0]
        sp_input['bounds'] = batch['bounds']
        sp_input['R'] = batch['R']
        sp_input['Th'] = batch['Th']
        sp_input['latent_index'] = batch['latent_index']
        return sp_input

    def get_density_color(self, wpts, viewdir, raw_decoder):
        """if_clight_renderer.py:54-60 (host API parity for subclasses that pass their own decoder)."""
        n_batch, n_pixel, n_sample = wpts.shape[:3]
        wpts = wpts.view(n_batch, n_pixel * n_sample, -1)
        viewdir = viewdir[:, :, None].repeat(1, 1, n_sample, 1).contiguous()
        viewdir = viewdir.view(n_batch, n_pixel * n_sample, -1)
        return raw_decoder(wpts, viewdir)

    # ------------------------------------------------------------------ once-per-frame packs
    def pack_volume(self, feature_volume, dtype):
        """NCDHW fp32 volumes -> channels-last blob (nb_pack_volume); cached until the tensors change."""
        # the key holds STRONG references to the keyed tensors (self._vol_keep): while an entry is cached, the caching
        # allocator cannot hand the same address to another frame's volumes, so (data_ptr, shape, _version) identifies them
        key = (dtype,) + tuple((v.data_ptr(), tuple(v.shape), v._version) for v in feature_volume)
        if key == self._vol_key:
            return self._vol_blob, self._vol_dims
        if len(feature_volume) != capi.NB_NUM_LEVELS:
            raise ValueError("expected %d feature volumes" % capi.NB_NUM_LEVELS)
        dev = feature_volume[0].device
        if dev.type != "cuda":
            raise RuntimeError("feature volumes must live on a CUDA device (no CPU render path)")
        B = feature_volume[0].shape[0]
        dims = capi.LevelDims()
        levels = (capi.nb_volume_level * capi.NB_NUM_LEVELS)()
        keep = []
        for l, v in enumerate(feature_volume):
            v = v.detach()
            if v.dtype != torch.float32 or not v.is_contiguous():
                v = v.float().contiguous()
            keep.append(v)
            _, c, d, h, w = v.shape
            dims[l][0], dims[l][1], dims[l][2], dims[l][3] = c, d, h, w
            levels[l].data = v.data_ptr()
            levels[l].C, levels[l].D, levels[l].H, levels[l].W = c, d, h, w
        nbytes = self.lib.nb_packed_volume_bytes(dims, B, dtype)
        blob = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        capi.check(self.lib.nb_pack_volume(levels, B, dtype, blob.data_ptr(), nbytes, C.c_void_p(stream)),
                   "nb_pack_volume")
        self._vol_key, self._vol_blob, self._vol_dims, self._vol_dtype = key, blob, dims, dtype
        self._vol_keep = list(feature_volume)
        return blob, dims

    def pack_weights(self, latent_index, device):
        """Fold + re-lay-out the decoder (nb_pack_weights); cached on parameter versions.  nb_pack_weights folds the VALUE
        of latent_index into the blob, so the cache keeps a strong reference to the keyed index tensor (self._w_keep): a
        new frame's freshly allocated index can then never alias the cached one's address."""
        tensors = self.net.decoder_tensors()
        key = tuple((t.data_ptr(), t._version) for t in tensors) + (latent_index.data_ptr(), latent_index._version,
                                                                    tuple(latent_index.shape), str(latent_index.device))
        if key == self._w_key:
            return self._w_blob
        B = int(latent_index.shape[0])
        w, keep = self._weights_struct(tensors, latent_index, device)
        nbytes = self.lib.nb_packed_weights_bytes(B)
        blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device).cuda_stream
        capi.check(self.lib.nb_pack_weights(C.byref(w), blob.data_ptr(), nbytes, C.c_void_p(stream)), "nb_pack_weights")
        self._w_key, self._w_blob = key, blob
        self._w_keep = (list(tensors), latent_index)
        return blob

    def _t_vals(self, S, device):
        k = (S, str(device))
        if k not in self._tvals:
            # upstream: torch.linspace(0., 1., steps=cfg.N_samples).to(near)  -- computed on the CPU, then moved
            self._tvals[k] = torch.linspace(0., 1., steps=S).to(device)
        return self._tvals[k]

    def _draw_t_rand(self, B, n, S, device):
        """The jitter draw of if_clight_renderer.py:22 (`torch.rand(z_vals.shape).to(upper)`, CPU
        generator), issued per 2048-ray chunk like upstream so the RNG stream is identical."""
        parts = [torch.rand((B, min(2048, n - i), S)) for i in range(0, n, 2048)]
        # (a training chunk is ONE part: no CPU torch.cat, whose OpenMP region is a thread hand-off per step -- on a busy
        # host such wake-ups were measured at 30-60 ms, ten times the GPU time of the step)
        host = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)
        return host.to(device).contiguous()

    # ------------------------------------------------------------------ fused launch
    def render_rays(self, ray_o, ray_d, near, far, feature_volume, sp_input, t_rand=None, want_raw=False,
                    out=None, trace=None, masks=None, z_vals=None, want_weights=None):
        """One nb_render_fwd launch for (B,n) rays.  Returns the dict of get_pixel_value.
        When autograd is recording and any volume / decoder tensor requires grad, the call goes through
        the exact kernel and `_FusedRender` so that `loss.backward()` works as it does upstream."""
        cfg = get_active_cfg()
        if int(self._opt("xyz_res", 10)) != 10 or int(self._opt("view_res", 4)) != 4:
            # embedder.py:53-54: the kernels (and view_fc's 346 input columns) are built for PE widths 63 / 27
            raise NotImplementedError("cfg.xyz_res / cfg.view_res other than 10 / 4 are not supported by the fused kernels")
        if float(cfg.raw_noise_std) > 0.:
            # upstream's branch draws CPU randn and would crash on GPU tensors (nerf_net_utils.py:33)
            raise NotImplementedError("raw_noise_std > 0 is not supported (it is 0 in every reference config)")
        dev = ray_o.device
        if dev.type != "cuda":
            raise RuntimeError("Renderer.render needs CUDA tensors: the render path has no CPU implementation")
        B, n = int(ray_o.shape[0]), int(ray_o.shape[1])
        S = int(cfg.N_samples) if z_vals is None else int(z_vals.shape[-1])   # z_vals: caller-supplied depths (fine pass, f-4)
        params = self.net.decoder_tensors()
        needs_grad = torch.is_grad_enabled() and (any(t.requires_grad for t in params) or
                                                  any(v.requires_grad for v in feature_volume))
        precision = self._train_precision(B, n, S) if needs_grad else self._precision()
        skip_empty = bool(self._opt("render_skip_empty", True))
        if precision != capi.NB_PRECISION_FP32 and (S > 1024 or n * S >= (1 << 28)):
            # the tensor-core pipeline works on a frame-wide sample list: rays of up to 1024 samples, < 2^28 samples per frame
            # (every reference config: 64 / 128 samples, <= 1024^2 rays); anything else runs on the exact kernel
            precision = capi.NB_PRECISION_FP32
        if z_vals is not None:
            t_rand = None
        elif t_rand is None and float(cfg.perturb) > 0. and self.net.training:
            t_rand = self._draw_t_rand(B, n, S, dev)
        call = {
            "B": B, "n": n, "S": S, "dev": dev, "precision": precision, "vdtype": self._volume_dtype(precision),
            "ray_o": _f32c(ray_o, dev), "ray_d": _f32c(ray_d, dev), "near": _f32c(near, dev), "far": _f32c(far, dev),
            "R": _f32c(sp_input['R'], dev), "Th": _f32c(sp_input['Th'], dev).reshape(B, 3),   # (B,1,3) or (B,3) upstream
            "bounds": _f32c(sp_input['bounds'], dev), "latent_index": sp_input['latent_index'],
            "out_sh": [int(v) for v in sp_input['out_sh']], "voxel_size": [float(v) for v in cfg.voxel_size],
            "t_rand": None if t_rand is None else _f32c(t_rand, dev), "white_bkgd": bool(cfg.white_bkgd),
            "z_vals": None if z_vals is None else _f32c(z_vals.detach(), dev),
            "feature_volume": list(feature_volume), "want_raw": want_raw or needs_grad, "user_raw": bool(want_raw), "out": out, "trace": trace,
            "want_weights": (bool(self._opt("render_return_weights", True)) if want_weights is None else bool(want_weights))
                            or needs_grad,
            "skip_empty": skip_empty, "stats": getattr(self, "stats", None),
            "masks": None,
        }
        if masks is not None:   # f-1: mask views of if_clight_renderer_mmsk.py (B = 1 only, as upstream)
            if needs_grad:
                raise NotImplementedError("mask views are an inference feature upstream (vis_novel_view / vis_novel_pose)")
            msks = masks["msks"][0].to(device=dev, dtype=torch.uint8).contiguous()
            snap = None
            if masks.get("R0_snap") is not None:   # single-view variant (if_clight_renderer_msk.py): SMPL -> snapshot world
                snap = (_f32c(masks["R0_snap"][0], dev), _f32c(masks["Th0_snap"][0], dev).reshape(3))
            call["masks"] = (msks, _f32c(masks["RT"][0][:, :3, :4], dev), _f32c(masks["Ks"][0], dev), snap)
        if call["t_rand"] is not None:
            assert tuple(call["t_rand"].shape) == (B, n, S)
        if needs_grad:
            rgb, disp, acc, depth, weights = _FusedRender.apply(self, call, *feature_volume, *params)
            ret = {'rgb_map': rgb, 'disp_map': disp, 'acc_map': acc, 'weights': weights, 'depth_map': depth}
            if want_raw:
                ret['raw'] = call["raw"]
            return ret
        return self._launch(call, save=False)

    def _pool_take(self, tag, numel, dtype, dev):
        """A buffer of >= numel elements from the renderer's pool (several can be out at once: the coarse and the fine pass of
        a hierarchical step each hold an activation record until their backward ran)."""
        free = self.__dict__.setdefault("_pool", {}).setdefault((tag, dtype, str(dev)), [])
        fits = [i for i, t in enumerate(free) if t.numel() >= numel]
        if fits:                      # best fit: the coarse pass must not grab the fine pass's (3x larger) record
            return free.pop(min(fits, key=lambda i: free[i].numel()))
        if len(free) >= 4:            # nothing fits and the pool is full: let the allocator have the smallest one back
            free.pop(min(range(len(free)), key=lambda i: free[i].numel()))
        return torch.empty(int(numel), dtype=dtype, device=dev)

    def _pool_give(self, tag, t):
        if t is not None:
            self.__dict__.setdefault("_pool", {}).setdefault((tag, t.dtype, str(t.device)), []).append(t)

    @staticmethod
    def _out_stride(out, B, n):
        """0 for dense output maps; the common ray stride (floats) when the caller's `out` tensors are columns of one fused
        (B,n,stride) record, e.g. the [rgb | disp | acc | depth] slab of neuralbody_b200.dist (nb_render_args.out_ray_stride)."""
        rgb = out['rgb_map']
        if rgb.is_contiguous() and all(out[k].is_contiguous() for k in ('disp_map', 'acc_map', 'depth_map')):
            return 0
        st = int(rgb.stride(1))
        ok = tuple(rgb.shape) == (B, n, 3) and rgb.stride(2) == 1 and rgb.stride(0) == st * n and all(
            tuple(out[k].shape) == (B, n) and out[k].stride(1) == st and out[k].stride(0) == st * n
            for k in ('disp_map', 'acc_map', 'depth_map'))
        if not ok:
            raise ValueError("`out` maps must be dense, or columns of one (B, n, stride) float32 record")
        return st

    def train_listed_samples(self):
        """[(listed, total)] of the most recent training-precision forward calls (the coarse and the fine pass of a hierarchical
        step): how many samples the exact empty-sample skipping left for the GEMM chains.  Synchronises; diagnostics / bench."""
        return [(int(sv[:4].view(torch.int32)[3].item()), total) for sv, total in self.__dict__.get("_train_records", [])]

    def release(self):
        """Drop every pooled / cached device buffer (activation records, backward scratch, workspace, packed blobs).  The pools
        assume ONE stream drives this Renderer: call this only when no launch of it is in flight."""
        self.__dict__.pop("_pool", None)
        self.__dict__.pop("_ws_cache", None)
        self.__dict__.pop("_train_records", None)
        self._vol_key = self._vol_blob = self._vol_dims = self._vol_keep = None
        self._w_key = self._w_blob = self._w_keep = None

    def _workspace(self, nbytes, dev):
        """Scratch for nb_render_fwd, grown on demand and reused by every later call on this device's stream."""
        cache = self.__dict__.setdefault("_ws_cache", {})
        ws = cache.get(dev)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            cache[dev] = ws
        return ws

    def _launch(self, call, save):
        """Pack (cached) + one nb_render_fwd on the current stream."""
        dev, B, n, S = call["dev"], call["B"], call["n"], call["S"]
        precision, vdtype = call["precision"], call["vdtype"]
        with torch.cuda.device(dev), torch.no_grad():
            vol_blob, dims = self.pack_volume(call["feature_volume"], vdtype)
            w_blob = self.pack_weights(call["latent_index"], dev)
            t_vals = self._t_vals(S, dev)
            out = call["out"]
            if out is None:
                out = {
                    'rgb_map': torch.empty((B, n, 3), dtype=torch.float32, device=dev),
                    'disp_map': torch.empty((B, n), dtype=torch.float32, device=dev),
                    'acc_map': torch.empty((B, n), dtype=torch.float32, device=dev),
                    'depth_map': torch.empty((B, n), dtype=torch.float32, device=dev),
                }
                if call["want_weights"]:
                    out['weights'] = torch.empty((B, n, S), dtype=torch.float32, device=dev)
            raw = None
            if call["want_raw"]:
                if save and not call.get("user_raw"):
                    # internal to the autograd node: recycled like the activation record (a fresh 1-3 MB tensor per call comes
                    # out of the caching allocator's large pool, where it splits the blocks the 100 MB volume gradients reuse)
                    raw = self._pool_take("raw", B * n * S * 4, torch.float32, dev)[:B * n * S * 4].view(B, n, S, 4)
                else:
                    raw = torch.empty((B, n, S, 4), dtype=torch.float32, device=dev)
            a = capi.nb_render_args()
            a.batch, a.n_rays, a.n_samples = B, n, S
            a.precision = precision
            sv = None
            if save:
                # the activation record (5.2 KB per sample) and the backward scratch are hundreds of MB per training chunk:
                # they are recycled through a small per-renderer pool instead of going back to the allocator every step
                sv = self._pool_take("save", self.lib.nb_render_save_bytes_for(C.byref(a)) // 4, torch.float32, dev)
            a.ray_o, a.ray_d = call["ray_o"].data_ptr(), call["ray_d"].data_ptr()
            a.near, a.far = call["near"].data_ptr(), call["far"].data_ptr()
            a.t_vals = t_vals.data_ptr()
            a.t_rand = call["t_rand"].data_ptr() if call["t_rand"] is not None else None
            a.z_vals = call["z_vals"].data_ptr() if call["z_vals"] is not None else None
            a.R, a.Th, a.bounds = call["R"].data_ptr(), call["Th"].data_ptr(), call["bounds"].data_ptr()
            for i in range(3):
                a.voxel_size[i] = call["voxel_size"][i]
                a.out_sh[i] = call["out_sh"][i]
            for l in range(capi.NB_NUM_LEVELS):
                for j in range(4):
                    a.level_dims[l][j] = dims[l][j]
            a.volume_blob, a.volume_dtype = vol_blob.data_ptr(), vdtype
            a.weights_blob = w_blob.data_ptr()
            a.white_bkgd = 1 if call["white_bkgd"] else 0
            a.precision = precision
            a.rgb_map, a.disp_map = out['rgb_map'].data_ptr(), out['disp_map'].data_ptr()
            a.acc_map, a.depth_map = out['acc_map'].data_ptr(), out['depth_map'].data_ptr()
            a.out_ray_stride = self._out_stride(out, B, n)
            a.weights = out['weights'].data_ptr() if 'weights' in out else None
            a.raw = raw.data_ptr() if raw is not None else None
            a.save = sv.data_ptr() if sv is not None else None
            a.skip_empty = 1 if call["skip_empty"] else 0
            if call["masks"] is not None:
                msks, RT, Ks, snap = call["masks"]
                a.mask_msks, a.mask_RT, a.mask_Ks = msks.data_ptr(), RT.data_ptr(), Ks.data_ptr()
                if snap is not None:
                    a.mask_R0, a.mask_Th0 = snap[0].data_ptr(), snap[1].data_ptr()
                a.mask_nv, a.mask_H, a.mask_W = int(msks.shape[0]), int(msks.shape[1]), int(msks.shape[2])
            a.stats = call["stats"].data_ptr() if call["stats"] is not None else None
            ws = None
            if precision in (capi.NB_PRECISION_TC_FP16, capi.NB_PRECISION_TC_FP16X3):   # classify -> decoder over the frame's sample list -> composite
                ws = self._workspace(self.lib.nb_render_fwd_workspace_bytes(B, n, S), dev)
                a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
            a.trace = call["trace"].data_ptr() if call["trace"] is not None else None   # diagnostics (tools/trace_timeline.py)
            stream = torch.cuda.current_stream(dev).cuda_stream
            capi.check(self.lib.nb_render_fwd(C.byref(a), C.c_void_p(stream)), "nb_render_fwd")
            self.launches += B * self.lib.nb_render_fwd_launches(precision)
            if save:   # everything nb_render_bwd needs stays alive with the autograd node
                call["args"], call["save"], call["raw"] = a, sv, raw
                if precision == capi.NB_PRECISION_TC_TF32X3:     # diagnostics: list lengths of the last records (train_listed_samples)
                    recs = self.__dict__.setdefault("_train_records", [])
                    recs.append((sv, B * n * S))
                    del recs[:-4]
                # NOT the output tensors: they carry grad_fn -> this call's autograd node -> ctx.call, a reference cycle only the
                # cyclic garbage collector can free (measured: ~7 MB leaked per training step and a 50-130 ms gc pause every few steps)
                call["keep"] = (vol_blob, w_blob, t_vals)
        if raw is not None and call["want_raw"] and not save:
            out = dict(out)
            out['raw'] = raw
        return out

    def _launch_bwd(self, call, d_rgb, d_depth, d_acc, needs):
        """nb_render_bwd: gradients for (volumes..., decoder tensors...) in the order of _FusedRender.apply."""
        dev, B, n, S = call["dev"], call["B"], call["n"], call["S"]
        if call.get("save") is None:
            raise RuntimeError("the activation record of this render call was already consumed by a backward pass "
                               "(backward twice through the same nb_render_fwd is not supported)")
        params = self.net.decoder_tensors()
        vols = call["feature_volume"]
        with torch.cuda.device(dev), torch.no_grad():
            def cf(t):
                return None if t is None else t.to(device=dev, dtype=torch.float32).contiguous()
            d_rgb, d_depth, d_acc = cf(d_rgb), cf(d_depth), cf(d_acc)
            w = self._weights_struct(params, call["latent_index"], dev)
            # 17 small tensors (they stay in the allocator's small pool), zeroed by one multi-tensor launch
            gparams = [torch.empty_like(t, dtype=torch.float32, device=dev) for t in params]
            torch._foreach_zero_(gparams)
            g = capi.nb_decoder_weights()
            names = [f[0] for f in capi.nb_decoder_weights._fields_][:17]
            for name, t in zip(names, gparams):
                setattr(g, name, t.data_ptr())
            g.latent_index, g.num_train_frame, g.batch = w[0].latent_index, w[0].num_train_frame, B
            want_vol = any(needs[:len(vols)])
            gvols = [torch.zeros_like(v, dtype=torch.float32, device=dev) for v in vols] if want_vol else [None] * len(vols)
            nbytes = self.lib.nb_render_bwd_workspace_bytes_for(C.pointer(call["args"]))
            ws = self._pool_take("bwd_ws", nbytes, torch.uint8, dev)
            ba = capi.nb_render_bwd_args()
            ba.fwd = C.pointer(call["args"])
            ba.save, ba.raw = call["save"].data_ptr(), call["raw"].data_ptr()
            ba.d_rgb_map = d_rgb.data_ptr() if d_rgb is not None else None
            ba.d_depth_map = d_depth.data_ptr() if d_depth is not None else None
            ba.d_acc_map = d_acc.data_ptr() if d_acc is not None else None
            ba.weights, ba.grads = C.pointer(w[0]), C.pointer(g)
            for l in range(capi.NB_NUM_LEVELS):
                ba.d_volumes[l] = gvols[l].data_ptr() if want_vol else None
            ba.workspace, ba.workspace_bytes = ws.data_ptr(), nbytes
            stream = torch.cuda.current_stream(dev).cuda_stream
            capi.check(self.lib.nb_render_bwd(C.byref(ba), C.c_void_p(stream)), "nb_render_bwd")
            # stream-ordered reuse: the next forward / backward on this stream runs after the kernels just enqueued
            self._pool_give("bwd_ws", ws)
            self._pool_give("save", call.pop("save"))
            if not call.get("user_raw"):
                r = call.pop("raw")
                self._pool_give("raw", r._base if r._base is not None else r)
        grads = list(gvols) + [gp.view_as(t) for gp, t in zip(gparams, params)]
        return [gr if need else None for gr, need in zip(grads, needs)]

    def _weights_struct(self, tensors, latent_index, device):
        """nb_decoder_weights over the raw parameter tensors (+ the tensors kept alive)."""
        w = capi.nb_decoder_weights()
        names = [f[0] for f in capi.nb_decoder_weights._fields_][:17]
        keep = []
        for name, t in zip(names, tensors):
            t = t.detach()
            if t.device != device or t.dtype != torch.float32 or not t.is_contiguous():
                t = t.to(device=device, dtype=torch.float32).contiguous()
            keep.append(t)
            setattr(w, name, t.data_ptr())
        li = latent_index.to(device=device, dtype=torch.int64).contiguous()
        keep.append(li)
        w.latent_index = li.data_ptr()
        w.num_train_frame = int(self.net.latent.weight.shape[0])
        w.batch = int(latent_index.shape[0])
        return w, keep

    def calculate_density(self, wpts, feature_volume, sp_input):
        """Network.calculate_density (latent_xyzc.py:74-89) on arbitrary world points: (B,P,3) -> (B,P,1).
        f-3: the alpha decoder of the mesh renderer (if_mesh_renderer.py:36-41); exact fp32 kernel."""
        cfg = get_active_cfg()
        dev = wpts.device
        if dev.type != "cuda":
            raise RuntimeError("calculate_density needs CUDA tensors: there is no CPU implementation")
        B, Pn = int(wpts.shape[0]), int(wpts.shape[1])
        with torch.cuda.device(dev), torch.no_grad():
            vol_blob, dims = self.pack_volume(feature_volume, capi.NB_DTYPE_F32)
            w_blob = self.pack_weights(sp_input['latent_index'], dev)
            pts = _f32c(wpts, dev)
            R, Th = _f32c(sp_input['R'], dev), _f32c(sp_input['Th'], dev).reshape(B, 3)
            bounds = _f32c(sp_input['bounds'], dev)
            sigma = torch.empty((B, Pn, 1), dtype=torch.float32, device=dev)
            a = capi.nb_render_args()
            a.batch = B
            a.R, a.Th, a.bounds = R.data_ptr(), Th.data_ptr(), bounds.data_ptr()
            for i in range(3):
                a.voxel_size[i] = float(cfg.voxel_size[i])
                a.out_sh[i] = int(sp_input['out_sh'][i])
            for l in range(capi.NB_NUM_LEVELS):
                for j in range(4):
                    a.level_dims[l][j] = dims[l][j]
            a.volume_blob, a.volume_dtype, a.weights_blob = vol_blob.data_ptr(), capi.NB_DTYPE_F32, w_blob.data_ptr()
            a.precision = capi.NB_PRECISION_FP32
            stream = torch.cuda.current_stream(dev).cuda_stream
            capi.check(self.lib.nb_decode_density(C.byref(a), pts.data_ptr(), Pn, sigma.data_ptr(), C.c_void_p(stream)),
                       "nb_decode_density")
        return sigma

    def get_pixel_value(self, ray_o, ray_d, near, far, feature_volume, sp_input, batch):
        """if_clight_renderer.py:62-92: same signature, same returned dict.  With `cfg.render_importance > 0` the call runs the
        coarse + fine passes of `render_rays_hierarchical` and the dict also carries rgb0 / disp0 / acc0 / z_std."""
        if int(self._opt("render_importance", 0)) > 0:
            return self.render_rays_hierarchical(ray_o, ray_d, near, far, feature_volume, sp_input)
        return self.render_rays(ray_o, ray_d, near, far, feature_volume, sp_input)

    # ------------------------------------------------------------------ f-4: hierarchical (coarse + importance) sampling
    def importance_z_vals(self, near, far, weights, n_samples, n_importance, t_rand=None, u=None):
        """z_vals_mid + sample_pdf + sort-merge (volume_renderer.py:84-93, nerf_net_utils.py:55-90) as one nb_sample_pdf launch.
        weights (B,n,S) from the coarse pass; t_rand its jitter (or None); u (B,n,n_importance) uniforms or None for the
        deterministic branch.  Returns (z_all (B,n,S+n_importance) ascending, z_samples (B,n,n_importance))."""
        dev = weights.device
        B, n, S = int(weights.shape[0]), int(weights.shape[1]), int(n_samples)
        Ni = int(n_importance)
        with torch.cuda.device(dev), torch.no_grad():
            z_all = torch.empty((B, n, S + Ni), dtype=torch.float32, device=dev)
            z_smp = torch.empty((B, n, Ni), dtype=torch.float32, device=dev)
            keep = [_f32c(near, dev), _f32c(far, dev), self._t_vals(S, dev), _f32c(weights.detach(), dev),
                    None if t_rand is None else _f32c(t_rand, dev), None if u is None else _f32c(u, dev)]
            a = capi.nb_importance_args()
            a.n_rays_total, a.n_samples, a.n_importance = B * n, S, Ni
            a.near, a.far, a.t_vals, a.weights = keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), keep[3].data_ptr()
            a.t_rand = keep[4].data_ptr() if keep[4] is not None else None
            a.u = keep[5].data_ptr() if keep[5] is not None else None
            a.z_out, a.z_samples = z_all.data_ptr(), z_smp.data_ptr()
            stream = torch.cuda.current_stream(dev).cuda_stream
            capi.check(self.lib.nb_sample_pdf(C.byref(a), C.c_void_p(stream)), "nb_sample_pdf")
            self.launches += 1
        return z_all, z_smp

    def render_rays_hierarchical(self, ray_o, ray_d, near, far, feature_volume, sp_input, t_rand=None, u=None):
        """Coarse pass (cfg.N_samples) -> importance samples from its weights -> fine pass over the merged depths with the SAME
        network, as the reference's NeRF-baseline renderer does (volume_renderer.py:60-118; Neural Body's own renderer has no
        fine pass).  `cfg.render_importance` = N_importance; det = (cfg.perturb == 0) as upstream.
        Both passes are nb_render_fwd launches (the fine one with nb_render_args.z_vals); under autograd each is a
        `_FusedRender` node and the importance samples are detached, as upstream."""
        cfg = get_active_cfg()
        Ni = int(self._opt("render_importance", 0))
        S = int(cfg.N_samples)
        dev = ray_o.device
        B, n = int(ray_o.shape[0]), int(ray_o.shape[1])
        if t_rand is None and float(cfg.perturb) > 0. and self.net.training:
            t_rand = self._draw_t_rand(B, n, S, dev)
        if u is None and float(cfg.perturb) != 0.:
            u = torch.rand((B * n, Ni)).view(B, n, Ni).to(dev)          # nerf_net_utils.py:70 (CPU generator, like upstream)
        coarse = self.render_rays(ray_o, ray_d, near, far, feature_volume, sp_input, t_rand=t_rand,
                                  want_weights=True)          # the coarse weights drive the importance sampling
        z_all, z_smp = self.importance_z_vals(near, far, coarse['weights'], S, Ni, t_rand=t_rand, u=u)
        fine = dict(self.render_rays(ray_o, ray_d, near, far, feature_volume, sp_input, z_vals=z_all))
        fine['rgb0'], fine['disp0'], fine['acc0'] = coarse['rgb_map'], coarse['disp_map'], coarse['acc_map']
        fine['z_std'] = torch.std(z_smp, dim=-1, unbiased=False)
        return fine

    # ------------------------------------------------------------------ a1
    def render(self, batch):
        """if_clight_renderer.py:94-122.  `cfg.chunk` rays per launch (0 = everything in one
        launch; upstream hard-codes 2048 to bound activation memory, which the fused kernel
        never materialises)."""
        ray_o = batch['ray_o']
        ray_d = batch['ray_d']
        near = batch['near']
        far = batch['far']

        sp_input = self.prepare_sp_input(batch)
        feature_volume = self.net.encode_sparse_voxels(sp_input)

        n_pixel = ray_o.shape[1]
        chunk = int(self._opt("chunk", 0)) or n_pixel
        if chunk >= n_pixel:
            return self.get_pixel_value(ray_o, ray_d, near, far, feature_volume, sp_input, batch)
        ret_list = []
        for i in range(0, n_pixel, chunk):
            ret_list.append(self.get_pixel_value(ray_o[:, i:i + chunk], ray_d[:, i:i + chunk], near[:, i:i + chunk],
                                                 far[:, i:i + chunk], feature_volume, sp_input, batch))
        keys = ret_list[0].keys()
        return {k: torch.cat([r[k] for r in ret_list], dim=1) for k in keys}
