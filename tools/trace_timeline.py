"""Diagnostics (GPU): per-role timeline of CTA 0 of the tensor-core render kernel, from the clock64 trace
the kernel writes when nb_render_args.trace is set.  Usage: python tools/trace_timeline.py [precision] [first_tile] [n_tiles]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

PROD = {1: "tile begin (list rows loaded)"}
PROD.update({10 + s: "seg%d buf free" % s for s in range(6)})
PROD.update({20 + s: "seg%d gathered" % s for s in range(6)})
MMA = {1: "tile begin", 20: "L0 issued", 21: "L1 issued", 22: "L2 issued", 23: "L3 issued", 31: "h ready L1", 32: "h ready L2",
       33: "h ready L3", 34: "h ready L4"}
MMA.update({10 + s: "seg%d available" % s for s in range(6)})
EPI = {1: "tile begin", 2: "PE written", 10: "acc L0", 11: "acc L1", 12: "acc L2", 13: "acc L3", 14: "acc L4 (rgb)", 20: "epi L0 done",
       21: "epi L1 done", 22: "epi L2 done", 23: "epi L3 done"}


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "tc_fp16x3"
    firsts = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6]
    ntile = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    from oracle import synth
    from neuralbody_b200.lib.config import cfg
    import gpu_utils as G
    scene = synth.make_scene(H=512, W=512, scale=1.0, all_hit=True)
    net, ren = G.make_net_and_renderer(scene)
    cfg.N_samples, cfg.perturb, cfg.white_bkgd, cfg.render_precision, cfg.render_volume_dtype = 64, 0.0, False, prec, "auto"
    cfg.render_skip_empty = not (len(sys.argv) > 4 and sys.argv[4] == "dense")
    net.eval()
    batch = {k: scene[k].cuda() for k in G.BATCH_KEYS}
    sp = ren.prepare_sp_input(batch)
    vol = net.encode_sparse_voxels(sp)
    trace = torch.zeros(4 * 4096, dtype=torch.int64, device="cuda")
    for _ in range(2):
        trace.zero_()
        with torch.no_grad():
            ren.render_rays(batch["ray_o"], batch["ray_d"], batch["near"], batch["far"], vol, sp, trace=trace)
    torch.cuda.synchronize()
    t = trace.cpu().view(4, 4096)
    MMA.update({40: "slot wait", 41: "slot: own half landed", 42: "slot: peer half landed", 44: "h wait", 45: "h ready"})
    LOAD = {1: "tile begin", 2: "push: wait for a free slot", 3: "push: slot free, copy issued"}
    roles = [("PROD", PROD, 13, 1), ("MMA", MMA, None, 1), ("EPI", EPI, None, 1), ("LOAD", LOAD, None, 1)]
    events = []
    for r, (name, names, _, begin_code) in enumerate(roles):
        tile = -1
        for v in t[r].tolist():
            if v == 0:
                break
            code, clk = (v >> 48) & 0xFFFF, v & 0xFFFFFFFFFFFF
            if code == begin_code:
                tile += 1
            events.append((clk, name, tile, names.get(code, str(code))))
    events.sort()
    for first in firsts:
        sel = [e for e in events if first <= e[2] < first + ntile]
        if not sel:
            continue
        t0 = sel[0][0]
        last = {}
        print("---- tiles %d..%d of CTA 0" % (first, first + ntile - 1))
        for clk, name, tile, what in sel:
            d = clk - last.get(name, clk)
            last[name] = clk
            print("%9d  (+%6d)  %-5s tile %-3d %s" % (clk - t0, d, name, tile, what))
    # per-tile period of the MMA role
    begins = [e[0] for e in events if e[1] == "MMA" and e[3] == "tile begin"]
    if len(begins) > 3:
        per = [b - a for a, b in zip(begins[:-1], begins[1:])]
        print("MMA tile period (cycles): mean %.0f  min %d  max %d  over %d tiles" % (sum(per) / len(per), min(per), max(per), len(per)))
        for lo in range(0, len(per), 40):
            chunk = per[lo:lo + 40]
            print("  tiles %3d..%3d: mean period %.0f" % (lo, lo + len(chunk) - 1, sum(chunk) / len(chunk)))


if __name__ == "__main__":
    main()
