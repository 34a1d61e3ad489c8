"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see below).

Brute-force restatement of the parts of spconv 1.2.1 (commit abf0acf30f5526ea93e687e3f424f62d9cd8313a, the version the
reference pins in INSTALL.md:14-22) that the reference's SparseConvNet uses (lib/networks/latent_xyzc.py:166-274): a sparse
tensor is a python dict {(b, z, y, x): feature row}; every operator loops over sites.  spconv itself is not in this image
and there is no network, so this file CANNOT be checked against the real library: it states spconv's published semantics
(the same ones SURVEY.md 8c lists) and is only used to cross-check the dense emulation
neuralbody_b200/lib/networks/sparse_encode.py on inputs small enough for python loops (tests/test_sparse_encode.py).

Semantics restated:
  SubMConv3d(k=3, bias=False)            out[p] = sum_k in[p + k - 1] @ W[k]  for p in the INPUT active set only; absent
                                         neighbours contribute nothing; W is [kD, kH, kW, Cin, Cout]
  SparseConv3d(k=3, s=2, p=1)            out[o] = sum_k in[2 o - 1 + k] @ W[k] for every o with at least one active input in
                                         its window; output extent floor((in - 1) / 2) + 1 per axis
  BatchNorm1d(eps, batch statistics)     over the rows of the active set (biased variance), then ReLU
  .dense()                               (B, C, D, H, W), zeros off the active set
  duplicate input coordinates            order-dependent in spconv; here the LAST row wins (as the emulation defines it)
"""
import itertools

import numpy as np

OFFS = list(itertools.product(range(3), range(3), range(3)))


def from_points(feats, coords):
    """feats (N,C), coords (N,4) int (b,z,y,x) -> dict; later rows overwrite earlier ones at the same site."""
    out = {}
    for f, c in zip(np.asarray(feats, dtype=np.float64), np.asarray(coords)):
        out[tuple(int(v) for v in c)] = f.copy()
    return out


def subm_conv(x, W):
    W = np.asarray(W, dtype=np.float64)
    out = {}
    for (b, z, y, xx) in x:
        acc = np.zeros(W.shape[-1])
        for (kz, ky, kx) in OFFS:
            q = (b, z + kz - 1, y + ky - 1, xx + kx - 1)
            if q in x:
                acc += x[q] @ W[kz, ky, kx]
        out[(b, z, y, xx)] = acc
    return out


def strided_conv(x, W, shape):
    """shape = input (D,H,W) -> (dict, output shape)."""
    W = np.asarray(W, dtype=np.float64)
    oshape = tuple((s - 1) // 2 + 1 for s in shape)
    out = {}
    for (b, z, y, xx), f in x.items():
        for (kz, ky, kx) in OFFS:
            # input i = 2 o - 1 + k  =>  o = (i + 1 - k) / 2 when that is an integer inside the output extent
            num = (z + 1 - kz, y + 1 - ky, xx + 1 - kx)
            if any(v % 2 for v in num):
                continue
            o = tuple(v // 2 for v in num)
            if any(v < 0 or v >= s for v, s in zip(o, oshape)):
                continue
            key = (b,) + o
            out.setdefault(key, np.zeros(W.shape[-1]))
            out[key] += f @ W[kz, ky, kx]
    return out, oshape


def bn_relu(x, gamma, beta, eps=1e-3):
    keys = list(x)
    rows = np.stack([x[k] for k in keys])
    mean, var = rows.mean(0), rows.var(0)
    rows = (rows - mean) / np.sqrt(var + eps) * np.asarray(gamma, dtype=np.float64) + np.asarray(beta, dtype=np.float64)
    return {k: np.maximum(r, 0.0) for k, r in zip(keys, rows)}


def dense(x, batch, channels, shape):
    out = np.zeros((batch, channels) + tuple(shape))
    for (b, z, y, xx), f in x.items():
        out[b, :, z, y, xx] = f
    return out


def sparse_conv_net(params, feats, coords, shape, batch):
    """latent_xyzc.py:166-207 with `params` = the emulation's state_dict (numpy): returns [net1..net4] dense."""
    def block(x, name, n_convs, strided, shp):
        for i in range(n_convs):
            W = params["%s.%d.weight" % (name, 3 * i)]
            if strided:
                x, shp = strided_conv(x, W, shp)
            else:
                x = subm_conv(x, W)
            x = bn_relu(x, params["%s.%d.weight" % (name, 3 * i + 1)], params["%s.%d.bias" % (name, 3 * i + 1)])
        return x, shp

    x = from_points(feats, coords)
    shp = tuple(shape)
    vols = []
    x, shp = block(x, "conv0", 2, False, shp)
    x, shp = block(x, "down0", 1, True, shp)
    x, shp = block(x, "conv1", 2, False, shp)
    vols.append(dense(x, batch, 32, shp))
    x, shp = block(x, "down1", 1, True, shp)
    x, shp = block(x, "conv2", 3, False, shp)
    vols.append(dense(x, batch, 64, shp))
    x, shp = block(x, "down2", 1, True, shp)
    x, shp = block(x, "conv3", 3, False, shp)
    vols.append(dense(x, batch, 128, shp))
    x, shp = block(x, "down3", 1, True, shp)
    x, shp = block(x, "conv4", 3, False, shp)
    vols.append(dense(x, batch, 128, shp))
    return vols
