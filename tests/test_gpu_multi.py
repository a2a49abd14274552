"""Multi-GPU path (needs >= 2 GPUs; skipped otherwise): each rank holds one shard in its own fyx context,
culls it, and the NCCL all-gather (overlapped inside fyx_render_prep) gives every rank the visible set of
the whole, unsharded scene as computed by the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
N_NODES, N_UNITS, VERTS = 60000, 32, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import fyrox_b200 as fb
    from fyrox_b200 import camera
    from fyrox_b200.dist import broadcast_bytes
    from fyrox_b200.scenegen import Scene

    sc = Scene(N_NODES, n_units=N_UNITS, verts_per_unit=VERTS, rank=rank, nranks=world)
    ctx = fb.Context(device=rank)
    ctx.set_topology(sc.parent, sc.flags, sc.render_mask, sc.local_aabb, root=0, global_index=sc.global_index)
    ctx.set_local_matrices(sc.local_m16)
    for u in range(sc.n_units):
        verts, bb = sc.unit_vertices(u)
        ctx.add_skinned_surface(sc.unit_mesh_node(u), sc.unit_bone_nodes(u), sc.unit_inv_bind(u), verts)
    uid = broadcast_bytes(fb.Context.comm_unique_id() if rank == 0 else b"\0" * 128, 0, device="cuda")
    ctx.comm_init(world, rank, uid)
    frusta = camera.cube_frusta()
    out = {}
    for mode in ("fused", "separate", "pipelined", "pipelined_own"):
        if mode == "fused":
            ctx.render_prep(update_flags=fb.UPDATE_ALL, frusta=frusta, readback_visible=False, allgather=True)
        elif mode == "separate":
            ctx.update_and_cull(frusta, fb.UPDATE_ALL)
            ctx.allgather_visible()
        else:  # two frames in flight, gathered lists collected by fyx_frame_wait
            # "pipelined_own": only rank 0 takes the whole lists to the host, the others their own (FYX_FRAME_READBACK_OWN)
            own = mode == "pipelined_own" and rank != 0
            for k in range(3):
                ctx.render_prep(update_flags=fb.UPDATE_ALL, frusta=frusta, readback_visible=True, allgather=True, async_=True, readback_own=own)
                if k:
                    ctx.frame_wait()
            ctx.frame_wait()
            if own:
                for f in range(len(frusta)):
                    out[f"own_{f}"] = np.sort(ctx.get_visible(f))
        for f in range(len(frusta)):
            out[f"{mode}_{f}"] = np.sort(ctx.get_visible_gathered(f))  # still complete: fetched from the device copy on demand
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_gpu_cull_and_nccl_allgather_match_the_unsharded_oracle(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import oracle_binding as ob
    from fyrox_b200.scenegen import Scene
    from helpers import cube_frusta

    sc = Scene(N_NODES, n_units=N_UNITS, verts_per_unit=VERTS)
    og = ob.Graph.build(sc.parent, sc.flags, sc.render_mask, sc.local_m16, sc.local_aabb.copy())
    for u in range(sc.n_units):
        og.add_surface(sc.unit_mesh_node(u), sc.unit_bone_nodes(u))
    og.L.orc_graph_drop_messages(og.h)
    og.update_hierarchical_data()
    fos, _ = cube_frusta()
    want = [np.sort(og.from_graph(fo)) for fo in fos]
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        for mode in ("fused", "separate", "pipelined", "pipelined_own"):
            for f in range(len(fos)):
                got = z[f"{mode}_{f}"]
                assert np.array_equal(got, want[f]), f"rank {r} {mode} frustum {f}: {got.size} vs {want[f].size}"
        if r:  # the rank's own lists = the part of the oracle's set that lives in its shard
            shard = Scene(N_NODES, n_units=N_UNITS, verts_per_unit=VERTS, rank=r, nranks=world)
            mine = np.unique(shard.global_index)
            for f in range(len(fos)):
                assert np.array_equal(z[f"own_{f}"], np.intersect1d(want[f], mine))
