#!/usr/bin/env python
"""BASELINE.json configs[0]: 100k static nodes, 1 camera frustum, the reference's CPU Graph::update + cull,
timed on the oracle's port of that path (single thread, pointer tree + recursive DFS like the reference)."""
import ctypes as C
import os
import statistics
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402

import oracle_binding as ob  # noqa: E402
from fyrox_b200.scenegen import Scene  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
sc = Scene(n)
og = ob.Graph.build(sc.parent, sc.flags, sc.render_mask, sc.local_m16, sc.local_aabb.copy())
f = ob.frustum_from_vp(ob.mat4_mul(ob.perspective(16 / 9, float(np.deg2rad(60.0)), 0.1, 150.0), ob.look_at_rh((0, 0, 0), (0, 0, -1), (0, 1, 0))))
vis = np.empty(n, np.uint32)
L = og.L
tu, tc = [], []
for it in range(35):
    t0 = time.perf_counter()
    og.update_hierarchical_data()
    t1 = time.perf_counter()
    nv = L.orc_from_graph(og.h, C.byref(f), 0xFFFFFFFF, 0, vis.ctypes.data_as(C.c_void_p), n)
    t2 = time.perf_counter()
    if it >= 5:
        tu.append(t1 - t0)
        tc.append(t2 - t1)
mu, mc = statistics.median(tu) * 1e3, statistics.median(tc) * 1e3
print(f"nodes={n} visible={nv} update_all={mu:.3f} ms cull={mc:.3f} ms total={mu + mc:.3f} ms nodes/s={n / ((mu + mc) * 1e-3):.3e} (1 thread, host cores={os.cpu_count()})")
