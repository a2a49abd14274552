#!/usr/bin/env python
"""k_skin alone, back to back: is the gap between its ncu duration (alone, cold) and its time inside the benchmark's frames a
property of the frame sequence or of sustained load?  Times `reps` fyx_skin calls (each synchronised; CUDA-event durations from
fyx_get_timings) after one full frame, for a C4-sized vertex set, with idle gaps of `gap_ms` between the calls."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
import fyrox_b200 as fb
from fyrox_b200 import camera
from fyrox_b200.scenegen import Scene


def main():
    units = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    sc = Scene(units * 200, n_units=units, verts_per_unit=5000, bones_per_unit=64, seed=bench.SEED)
    ctx = fb.Context()
    bench.load_scene(ctx, sc, fb, lambda m: None)
    ctx.render_prep(update_flags=fb.UPDATE_ALL, frusta=camera.cube_frusta())
    out = {}
    nbytes = 68 * units * 5000
    for gap_ms in (0, 5, 50):
        t = []
        for _ in range(20):
            ctx.skin()
            t.append(ctx.timings()["skin_ms"])
            if gap_ms:
                time.sleep(gap_ms * 1e-3)
        t = np.array(t[3:])
        out[f"gap_{gap_ms}ms"] = {"mean_ms": float(t.mean()), "min_ms": float(t.min()), "max_ms": float(t.max()), "GBps_mean": nbytes / t.mean() / 1e6}
    print(json.dumps({"units": units, "verts": units * 5000, "skin_alone": out}))


if __name__ == "__main__":
    main()
