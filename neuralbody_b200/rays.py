"""f-2: camera rays on the device.  Replaces lib/utils/render_utils.py:120-137 (`image_rays`: get_rays + get_near_far +
mask_at_box compaction), which upstream runs in numpy per view and ships to the GPU (8.4 MB per 512x512 view)."""
import ctypes as C

import numpy as np
import torch

from . import capi


def image_rays(RT, K, bounds, H, W, device="cuda:0"):
    """RT (3,4) or (4,4) world->camera, K (3,3), bounds (2,3) world box -> ray_o, ray_d (n,3), near, far (n,), mask_at_box
    (H*W,) bool, all torch tensors on `device`; n = mask_at_box.sum() (order = row-major pixel order, as upstream)."""
    lib = capi.load()
    dev = torch.device(device)
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
