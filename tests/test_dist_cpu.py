"""CPU, gloo, world_size 2: the ray-sharding + one-all-gather host logic of neuralbody_b200/dist.py.
The per-rank render function is the oracle (tests may use it as a checker/stand-in); the check is the
one SURVEY 4(v) asks for: the sharded result equals the single-process result BIT FOR BIT."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rays, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import synth
        from neuralbody_b200 import dist as nbdist
        from oracle import neuralbody_oracle as O
        torch.set_num_threads(1)
        scene = synth.make_scene(H=16, W=16, scale=0.25, all_hit=True, n_rays=n_rays)

        def render_fn(batch):
            sc = dict(scene)
            sc.update({k: batch[k] for k in ("ray_o", "ray_d", "near", "far")})
            return O.render(sc, n_samples=16)

        batch = {k: scene[k] for k in ("coord", "out_sh", "bounds", "R", "Th", "latent_index", "ray_o", "ray_d", "near", "far")}
        full = nbdist.render_sharded(render_fn, batch, chunk=8)
        torch.save({k: v.clone() for k, v in full.items()}, os.path.join(out_dir, "rank%d.pt" % rank))
        # the streaming form: double-buffered slabs the render writes into, one gather + un-permute per view
        plan = nbdist.ShardPlan.get(n_rays, world, 8, "cpu")
        local = plan.shard(batch, rank)
        g = nbdist.FrameGatherer(n_rays, world, rank, "cpu", chunk=8, host=True, host_rank="rotate")
        frames = []
        for view in range(3):
            out = g.begin()
            ret = render_fn(local)
            for k in out:
                out[k].copy_(ret[k])
            frames.append(g.finish().clone())
        g.drain()
        torch.save(frames, os.path.join(out_dir, "frames%d.pt" % rank))
        # rotating owner: view v lands in the host buffer of rank v % world (views 0 and 2 on rank 0, view 1 on rank 1)
        torch.save(g.host[rank].clone(), os.path.join(out_dir, "host%d.pt" % rank))
        if rank == 0:
            torch.save(render_fn(batch), os.path.join(out_dir, "single.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [40, 37])     # even split and a ragged last slab (padding dropped)
def test_ray_sharded_render_equals_single_process(tmp_path, n_rays):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_rays, str(tmp_path)), nprocs=world, join=True)
    single = torch.load(os.path.join(tmp_path, "single.pt"))
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, "rank%d.pt" % r))
        for k in ("rgb_map", "disp_map", "acc_map", "depth_map"):
            assert got[k].shape == single[k].shape, k
            assert torch.equal(torch.nan_to_num(got[k]), torch.nan_to_num(single[k])), (r, k)
        from neuralbody_b200 import dist as nbdist
        for frame in torch.load(os.path.join(tmp_path, "frames%d.pt" % r)):
            views = nbdist.slab_views(frame)
            for k in ("rgb_map", "disp_map", "acc_map", "depth_map"):
                assert torch.equal(torch.nan_to_num(views[k]), torch.nan_to_num(single[k])), (r, k)
        hviews = nbdist.slab_views(torch.load(os.path.join(tmp_path, "host%d.pt" % r)))
        for k in ("rgb_map", "disp_map", "acc_map", "depth_map"):
            assert torch.equal(torch.nan_to_num(hviews[k]), torch.nan_to_num(single[k])), ("host", r, k)


def test_interleaved_shards_cover_all_rays():
    from neuralbody_b200.dist import shard_indices
    for n in (1, 7, 300, 262144, 262145):
        for world in (1, 2, 3, 8):
            for chunk in (4, 256):
                seen, pers = [], set()
                for r in range(world):
                    idx, per = shard_indices(n, r, world, chunk)
                    assert idx.numel() == per and int(idx.max()) < n
                    pers.add(per)
                    seen += idx.tolist()
                assert len(pers) == 1                         # equal shard sizes => one plain all-gather
                assert sorted(set(seen)) == list(range(n))   # every ray rendered (padding only duplicates)
