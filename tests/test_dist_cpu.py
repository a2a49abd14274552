"""N>1 host logic on CPU (gloo, world size 2): sector sharding with a replicated root + the variable-length
all-gather of visible lists give exactly the visible set of the unsharded scene."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)

N_NODES, N_UNITS, VERTS = 40000, 24, 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frusta(ob):
    faces = [((1, 0, 0), (0, -1, 0)), ((-1, 0, 0), (0, -1, 0)), ((0, 0, -1), (0, -1, 0))]
    out = []
    for look, up in faces:
        out.append(ob.frustum_from_vp(ob.mat4_mul(ob.perspective(1.0, float(np.pi / 2), 0.01, 120.0), ob.look_at_rh((0, 0, 0), look, up))))
    return out


def _visible_global(ob, sc):
    og = ob.Graph.build(sc.parent, sc.flags, sc.render_mask, sc.local_m16, sc.local_aabb.copy())
    og.update_hierarchical_data()
    return [np.sort(sc.global_index[og.from_graph(f)]) for f in _frusta(ob)]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_binding as ob
    from fyrox_b200.dist import allgather_varlen, broadcast_bytes
    from fyrox_b200.scenegen import Scene

    sc = Scene(N_NODES, n_units=N_UNITS, verts_per_unit=VERTS, rank=rank, nranks=world)
    lists = _visible_global(ob, sc)
    gathered = []
    for v in lists:
        t, counts = allgather_varlen(torch.from_numpy(v.astype(np.int64)))
        assert counts[rank] == v.size and sum(counts) == t.numel()
        gathered.append(np.sort(t.numpy()))
    uid = broadcast_bytes(bytes(range(128)) if rank == 0 else b"", 0) if rank == 0 else broadcast_bytes(b"\0" * 128, 0)
    assert uid == bytes(range(128))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), *gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_cull_plus_allgather_equals_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, HERE)
    import oracle_binding as ob
    from fyrox_b200.scenegen import Scene

    whole = _visible_global(ob, Scene(N_NODES, n_units=N_UNITS, verts_per_unit=VERTS))
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        for f, want in enumerate(whole):
            got = z[f"arr_{f}"]
            assert np.array_equal(got, want), f"rank {r} frustum {f}: {got.size} vs {want.size}"
    assert sum(w.size for w in whole) > 0


def test_host_segment_protocol_with_forked_ranks():
    """fyx_hostseg.hpp (the node-wide host segment every rank DMA-copies its own visible lists into): 4 forked processes,
    400 epochs, two-slot reuse, whole-list verification by the consumer — tests/cpp/test_hostseg.cpp, CPU only."""
    import subprocess

    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp")
    subprocess.run(["make", "-C", d, "test_hostseg"], check=True, capture_output=True)
    out = subprocess.run([os.path.join(d, "test_hostseg")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "hostseg ok" in out.stdout
