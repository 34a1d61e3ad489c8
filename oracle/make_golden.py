"""Generate tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_harness.py)
on the seeded synthetic cases of oracle/golden_cases.py.  Run in the build container:

    python -m oracle.make_golden

Each file holds the reference's five outputs (fp32) plus a sha256 of every input tensor, so a
consumer on another machine can prove it rebuilt the identical inputs from the seeds.
TEST INFRASTRUCTURE ONLY."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from neuralbody_b200 import synth
    from oracle import ref_harness, golden_cases
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in golden_cases.CASES:
        scene, rkw = golden_cases.build_case(name)
        ret = ref_harness.reference_render(scene, **rkw)
        arrays = {k: v.numpy().astype(np.float32) for k, v in ret.items()}
        if "masks" in rkw:
            arrays["mask_sha256"] = np.frombuffer(synth.scene_checksum({**scene, "weights": {}, "volumes": [
                rkw["masks"]["RT"], rkw["masks"]["Ks"], rkw["masks"]["msks"].float()]}).encode(), dtype=np.uint8)
        arrays["input_sha256"] = np.frombuffer(synth.scene_checksum(scene).encode(), dtype=np.uint8)
        arrays["torch_version"] = np.frombuffer(torch.__version__.encode(), dtype=np.uint8)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **arrays)
        acc = ret["acc_map"]
        print("%-22s rays=%-5d acc.mean=%.3f nan_disp=%d -> %s (%d KB)" % (
            name, acc.numel(), float(acc.mean()), int(torch.isnan(ret["disp_map"]).sum()), path,
            os.path.getsize(path) // 1024))


def grad_golden():
    """Per-tensor gradient fingerprints from the reference's own autograd (oracle/grad_case.py)."""
    import torch
    from neuralbody_b200 import synth
    from oracle import ref_harness, grad_case
    scene, t_rand, G = grad_case.build()
    ret, net, vols = ref_harness.reference_render(scene, n_samples=grad_case.N_SAMPLES, perturb=1.0, training=True,
                                                  white_bkgd=True, t_rand=t_rand, grad=True)
    grad_case.loss_of(ret, G).backward()
    sd = dict(net.named_parameters())
    arrays = {"input_sha256": np.frombuffer(synth.scene_checksum(scene).encode(), dtype=np.uint8)}
    for k in grad_case.GRAD_KEYS:
        g = sd[k].grad
        arrays["sum:" + k] = np.float64(g.double().sum())
        arrays["abs:" + k] = np.float64(g.double().abs().sum())
        arrays["head:" + k] = g.reshape(-1)[:64].numpy().astype(np.float32)
    for l, v in enumerate(vols):
        arrays["sum:vol%d" % l] = np.float64(v.grad.double().sum())
        arrays["abs:vol%d" % l] = np.float64(v.grad.double().abs().sum())
    path = os.path.join(ROOT, "tests", "golden", "grad_train_s32.npz")
    np.savez_compressed(path, **arrays)
    print("gradient fingerprints ->", path, "|dfc_0.weight|_1 = %.4e" % float(arrays["abs:fc_0.weight"]))


def hier_golden():
    """f-4: coarse + importance render composed from the reference's own functions (ref_harness.reference_render_hierarchical)."""
    import torch
    from neuralbody_b200 import synth
    from oracle import ref_harness, golden_cases
    for name in golden_cases.HIER_CASES:
        scene, rkw = golden_cases.build_hier_case(name)
        ret = ref_harness.reference_render_hierarchical(scene, **rkw)
        arrays = {k: v.numpy().astype(np.float32) for k, v in ret.items()}
        arrays["input_sha256"] = np.frombuffer(synth.scene_checksum(scene).encode(), dtype=np.uint8)
        arrays["torch_version"] = np.frombuffer(torch.__version__.encode(), dtype=np.uint8)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **arrays)
        print("%-22s rays=%-5d acc.mean=%.3f |rgb - rgb0|max=%.3f -> %s (%d KB)" % (
            name, ret["acc_map"].numel(), float(ret["acc_map"].mean()), float((ret["rgb_map"] - ret["rgb0"]).abs().max()),
            path, os.path.getsize(path) // 1024))


def hier_grad_golden():
    """f-4 gradient fingerprints: autograd of the reference's own functions through coarse pass, detached sample_pdf, fine pass."""
    from neuralbody_b200 import synth
    from oracle import ref_harness, grad_case
    scene, t_rand, u, G = grad_case.hier_build()
    ret, net, vols = ref_harness.reference_render_hierarchical(
        scene, n_samples=grad_case.N_SAMPLES, n_importance=grad_case.N_IMPORTANCE, perturb=1.0, training=True, white_bkgd=True,
        t_rand=t_rand, u=u, grad=True)
    grad_case.hier_loss_of(ret, G).backward()
    sd = dict(net.named_parameters())
    arrays = {"input_sha256": np.frombuffer(synth.scene_checksum(scene).encode(), dtype=np.uint8)}
    for k in grad_case.GRAD_KEYS:
        g = sd[k].grad
        arrays["sum:" + k] = np.float64(g.double().sum())
        arrays["abs:" + k] = np.float64(g.double().abs().sum())
        arrays["head:" + k] = g.reshape(-1)[:64].numpy().astype(np.float32)
    for l, v in enumerate(vols):
        arrays["sum:vol%d" % l] = np.float64(v.grad.double().sum())
        arrays["abs:vol%d" % l] = np.float64(v.grad.double().abs().sum())
    path = os.path.join(ROOT, "tests", "golden", "grad_hier_s32_i48.npz")
    np.savez_compressed(path, **arrays)
    print("hierarchical gradient fingerprints ->", path, "|dfc_0.weight|_1 = %.4e" % float(arrays["abs:fc_0.weight"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "hier":
        hier_golden()
        hier_grad_golden()
    else:
        grad_golden()
        main()
        hier_golden()
        hier_grad_golden()
