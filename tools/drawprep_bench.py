#!/usr/bin/env python
"""N3 at scale: fyx_pack_instances on the visible list of a big static scene (default: the C2 scene, 10 M nodes, one
camera frustum).  Prints one JSON line: call latency (wall clock around the synchronous C-ABI call: three launches,
the counts read-back and the final synchronisation included) and the algorithmic bytes per instance it moves.
Kernel-only durations come from ncu (profiles/)."""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

import fyrox_b200 as fb  # noqa: E402
from fyrox_b200 import camera  # noqa: E402
from fyrox_b200.scenegen import Scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nodes", type=int, default=10_000_000)
ap.add_argument("--bundles", type=int, default=4096)
ap.add_argument("--frusta", type=int, default=1, choices=[1, 6])
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()

sc = Scene(args.nodes)
ctx = fb.Context()
ctx.set_topology(sc.parent, sc.flags, sc.render_mask, sc.local_aabb, root=0)
ctx.set_local_matrices(sc.local_m16)
rng = np.random.default_rng(3)
ctx.set_bundle_ids(rng.integers(0, args.bundles, sc.capacity).astype(np.uint32))
ctx.enable_instances()
frusta = [camera.camera_frustum()] if args.frusta == 1 else camera.cube_frusta()
view = np.eye(4, dtype=np.float32).reshape(16)  # camera at the origin looking down -Z: look_at_rh is the identity rotation
vp = fb.mat4_mul(np.eye(4, dtype=np.float32).reshape(16), view)
ctx.update_and_cull(frusta, fb.UPDATE_ALL)
for _ in range(3):
    inst = ctx.pack_instances(0, view, vp)
n = int(inst["node"].size)
t0 = time.perf_counter()
for _ in range(args.steps):
    ctx._chk(ctx._lib.fyx_pack_instances(ctx._h, 0, view.ctypes.data, vp.ctypes.data))
ms = (time.perf_counter() - t0) / args.steps * 1e3
per_inst = 4 + 4 + 48 + 4 + 4 + 8 + (8 + 4 + 4 + 48 + 4) + 4 + 8 + 128  # pass 1 reads + tmp; pass 3 reads + writes
print(json.dumps({"tool": "drawprep_bench", "nodes": args.nodes, "visible": n, "bundles": int(inst["bundles"].size), "ms_per_call": ms,
                  "approx_bytes_per_instance": per_inst, "GBps_incl_call_overhead": per_inst * n / (ms * 1e-3) / 1e9}))
