"""f-2: camera rays on the device.  Replaces lib/utils/render_utils.py:120-137 (`image_rays`: get_rays + get_near_far +
mask_at_box compaction), which upstream runs in numpy per view and ships to the GPU (8.4 MB per 512x512 view)."""
import ctypes as C

import numpy as np
import torch

from . import capi


def _camera(RT, K, bounds, H, W):
    RT = np.asarray(RT, dtype=np.float64)
    cam = capi.nb_camera()
    kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    for i, v in enumerate(kinv.reshape(-1)):
        cam.K_inv[i] = float(v)
    for i, v in enumerate(RT[:3, :3].reshape(-1)):
        cam.R[i] = float(v)
    for i, v in enumerate(RT[:3, 3].reshape(-1)):
        cam.T[i] = float(v)
    for i, v in enumerate(np.asarray(bounds, dtype=np.float32).astype(np.float64).reshape(-1)):
        cam.bounds[i] = float(v)
    cam.H, cam.W = int(H), int(W)
    return cam


class ShardedRays:
    """Fixed-shape device buffers for one rank's interleaved shard of an H x W view (nb_gen_rays_sharded): nothing is
    compacted, so generating the rays of the next view never synchronises with the host.  Rays that miss the box (upstream
    drops them, render_utils.py:131-132) are dead rays with mask 0."""

    def __init__(self, H, W, rank=0, world=1, chunk=256, device="cuda:0"):
        self.H, self.W, self.rank, self.world, self.chunk = int(H), int(W), int(rank), int(world), int(chunk)
        n = self.H * self.W
        n_chunks = (n + chunk - 1) // chunk
        self.n_local = ((n_chunks + world - 1) // world) * chunk
        dev = torch.device(device)
        self.ray_o = torch.empty((1, self.n_local, 3), dtype=torch.float32, device=dev)
        self.ray_d = torch.empty((1, self.n_local, 3), dtype=torch.float32, device=dev)
        self.near = torch.empty((1, self.n_local), dtype=torch.float32, device=dev)
        self.far = torch.empty((1, self.n_local), dtype=torch.float32, device=dev)
        self.mask = torch.empty((1, self.n_local), dtype=torch.uint8, device=dev)
        self.lib = capi.load()

    def generate(self, RT, K, bounds):
        """Enqueue the ray generation of one view on the current stream; returns self (ray_o, ray_d, near, far, mask)."""
        cam = _camera(RT, K, bounds, self.H, self.W)
        dev = self.ray_o.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            capi.check(self.lib.nb_gen_rays_sharded(C.byref(cam), self.rank, self.world, self.chunk, self.n_local,
                                                    self.ray_o.data_ptr(), self.ray_d.data_ptr(), self.near.data_ptr(),
                                                    self.far.data_ptr(), self.mask.data_ptr(), C.c_void_p(stream)),
                       "nb_gen_rays_sharded")
        return self


def image_rays(RT, K, bounds, H, W, device="cuda:0"):
    """RT (3,4) or (4,4) world->camera, K (3,3), bounds (2,3) world box -> ray_o, ray_d (n,3), near, far (n,), mask_at_box
    (H*W,) bool, all torch tensors on `device`; n = mask_at_box.sum() (order = row-major pixel order, as upstream)."""
    lib = capi.load()
    dev = torch.device(device)
    cam = _camera(RT, K, bounds, H, W)
    n = int(H) * int(W)
    with torch.cuda.device(dev):
        ray_o = torch.empty((n, 3), dtype=torch.float32, device=dev)
        ray_d = torch.empty((n, 3), dtype=torch.float32, device=dev)
        near = torch.empty((n,), dtype=torch.float32, device=dev)
        far = torch.empty((n,), dtype=torch.float32, device=dev)
        mask = torch.empty((n,), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        capi.check(lib.nb_gen_rays(C.byref(cam), ray_o.data_ptr(), ray_d.data_ptr(), near.data_ptr(), far.data_ptr(),
                                   mask.data_ptr(), C.c_void_p(stream)), "nb_gen_rays")
        m = mask.bool()
        return ray_o[m], ray_d[m], near[m], far[m], m
