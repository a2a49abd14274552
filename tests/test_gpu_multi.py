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
    frusta = camera.cube_frusta()
    out = {}
    for variant, (ex, hs) in VARIANTS.items():
        # the exchange form is chosen by fyx_comm_init from the environment: one context per variant
        os.environ["FYX_EXCHANGE"], os.environ["FYX_HOSTSEG"] = ex, hs
        ctx = fb.Context(device=rank)
        ctx.set_topology(sc.parent, sc.flags, sc.render_mask, sc.local_aabb, root=0, global_index=sc.global_index)
        ctx.set_local_matrices(sc.local_m16)
        for u in range(sc.n_units):
            verts, bb = sc.unit_vertices(u)
            ctx.add_skinned_surface(sc.unit_mesh_node(u), sc.unit_bone_nodes(u), sc.unit_inv_bind(u), verts)
        uid = broadcast_bytes(fb.Context.comm_unique_id() if rank == 0 else b"\0" * 128, 0, device="cuda")
        ctx.comm_init(world, rank, uid)
        _run_modes(ctx, fb, frusta, rank, out, variant)
        out[f"{variant}_mode"] = np.frombuffer(ctx.comm_mode().encode(), dtype=np.uint8)
        dist.barrier()
        ctx.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    dist.barrier()
    dist.destroy_process_group()


VARIANTS = {"peer_seg": ("peer", "1"), "nccl_seg": ("nccl", "1"), "peer_private": ("peer", "0"), "nccl_private": ("nccl", "0")}
MODES = ("fused", "fused_readback", "separate", "pipelined", "pipelined_own")


def _run_modes(ctx, fb, frusta, rank, out, variant):
    for mode in MODES:
        if mode == "fused":
            ctx.render_prep(update_flags=fb.UPDATE_ALL, frusta=frusta, readback_visible=False, allgather=True)
        elif mode == "fused_readback":  # synchronous frame that asks for the lists: every rank publishes its part
            ctx.render_prep(update_flags=fb.UPDATE_ALL, frusta=frusta, readback_visible=True, allgather=True)
            if rank != 0:  # only rank 0 fetches the whole lists below; the others' own lists must still be right
                for f in range(len(frusta)):
                    out[f"{variant}_own2_{f}"] = np.sort(ctx.get_visible(f))
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
                    out[f"{variant}_own_{f}"] = np.sort(ctx.get_visible(f))
        if mode == "fused_readback" and rank != 0:
            continue  # a consumer on one rank only must not need the others to ask
        for f in range(len(frusta)):
            out[f"{variant}_{mode}_{f}"] = np.sort(ctx.get_visible_gathered(f))  # complete on every rank that asks
            p, n = ctx.get_visible_gathered_device(f)
            assert n == out[f"{variant}_{mode}_{f}"].size
        st = ctx.comm_stats()  # the exchange that produced these lists
        assert st["entries_total"] == sum(out[f"{variant}_{mode}_{f}"].size for f in range(len(frusta)))
        assert st["device_ms"] > 0.0 and st["entries_own"] <= st["entries_total"]
        if "peer" in variant:
            assert st["egress_bytes"] == 4 * st["entries_own"]  # 2 ranks: one peer


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
        shard = Scene(N_NODES, n_units=N_UNITS, verts_per_unit=VERTS, rank=r, nranks=world)
        mine = np.unique(shard.global_index)
        for variant, (ex, hs) in VARIANTS.items():
            desc = bytes(z[f"{variant}_mode"]).decode()
            # cudaIpc / memfd must really be in use on the box (a silent fallback to NCCL would pass the set compares)
            assert ("peer stores" in desc) == (ex == "peer"), desc
            assert ("host segment" in desc) == (hs == "1"), desc
            for mode in MODES:
                if mode == "fused_readback" and r != 0:
                    continue
                for f in range(len(fos)):
                    got = z[f"{variant}_{mode}_{f}"]
                    assert np.array_equal(got, want[f]), f"rank {r} {variant} {mode} frustum {f}: {got.size} vs {want[f].size}"
            if r:  # the rank's own lists = the part of the oracle's set that lives in its shard
                for f in range(len(fos)):
                    assert np.array_equal(z[f"{variant}_own_{f}"], np.intersect1d(want[f], mine))
                    assert np.array_equal(z[f"{variant}_own2_{f}"], np.intersect1d(want[f], mine))
