"""numpy-facing wrapper of libfyrox_scenegen.so (fyrox_b200/csrc/scenegen.h): deterministic synthetic
Fyrox scenes of the shapes BASELINE.json names.  Input generator only — not part of the hot path."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

VERTEX_BYTES = 68  # AnimatedVertex, scene/mesh/vertex.rs:140-155


def set_threads(n: int):
    """OpenMP threads of the generator (torchrun exports OMP_NUM_THREADS=1)."""
    L.load_scenegen().sg_set_threads(int(n))


class Scene:
    def __init__(self, n_nodes: int, n_units: int = 0, verts_per_unit: int = 5000, bones_per_unit: int = 64, seed: int = 0xF1A0C5,
                 rank: int = 0, nranks: int = 1):
        self._sg = L.load_scenegen()
        cfg = L.sg_config(seed, n_nodes, n_units, bones_per_unit, verts_per_unit, rank, nranks)
        self._h = self._sg.sg_create(C.byref(cfg))
        if not self._h:
            raise ValueError("sg_create failed (n_nodes too small for the requested units?)")
        self.cfg = cfg
        self.capacity = self._sg.sg_capacity(self._h)
        self.n_units = self._sg.sg_n_units(self._h)
        self.n_renderable = self._sg.sg_n_renderable(self._h)
        self.bones_per_unit = bones_per_unit
        self.verts_per_unit = verts_per_unit
        n = self.capacity

        def view(ptr, count, dtype):
            return np.ctypeslib.as_array(ptr, shape=(count,)).view(dtype)

        self.parent = view(self._sg.sg_parent(self._h), n, np.uint32)
        self.flags = view(self._sg.sg_flags(self._h), n, np.uint32)
        self.render_mask = view(self._sg.sg_render_mask(self._h), n, np.uint32)
        self.local_m16 = view(self._sg.sg_local_m16(self._h), n * 16, np.float32).reshape(n, 16)
        self.local_aabb = view(self._sg.sg_local_aabb(self._h), n * 6, np.float32).reshape(n, 6)
        self.global_index = view(self._sg.sg_global_index(self._h), n, np.uint32)

    def close(self):
        if getattr(self, "_h", None):
            self.parent = self.flags = self.render_mask = self.local_m16 = self.local_aabb = self.global_index = None
            self._sg.sg_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # units
    def unit_mesh_node(self, u: int) -> int:
        return self._sg.sg_unit_mesh_node(self._h, u)

    def unit_bone_nodes(self, u: int) -> np.ndarray:
        return np.ctypeslib.as_array(self._sg.sg_unit_bone_nodes(self._h, u), shape=(self.bones_per_unit,))

    def unit_inv_bind(self, u: int) -> np.ndarray:
        return np.ctypeslib.as_array(self._sg.sg_unit_inv_bind(self._h, u), shape=(self.bones_per_unit * 16,)).reshape(-1, 16)

    def unit_vertices(self, u: int):
        """(bytes array [V*68] uint8, aabb[6])"""
        buf = np.empty(self.verts_per_unit * VERTEX_BYTES, dtype=np.uint8)
        aabb = np.empty(6, dtype=np.float32)
        self._sg.sg_unit_vertices(self._h, u, buf.ctypes.data_as(C.c_void_p), aabb.ctypes.data_as(C.c_void_p))
        return buf, aabb

    def units_vertices_into(self, u0: int, count: int, out_addr: int, aabb_out: np.ndarray):
        """Generate `count` units' vertex buffers back to back at raw address out_addr (OpenMP)."""
        self._sg.sg_units_vertices(self._h, u0, count, C.c_void_p(out_addr), aabb_out.ctypes.data_as(C.c_void_p))

    def animate_into(self, frame: int, idx_addr: int, m16_addr: int) -> int:
        return self._sg.sg_animate(self._h, frame, C.c_void_p(idx_addr) if idx_addr else None, C.c_void_p(m16_addr))

    def animate(self, frame: int):
        n = self.n_units * self.bones_per_unit
        idx = np.empty(n, dtype=np.uint32)
        m = np.empty((n, 16), dtype=np.float32)
        self._sg.sg_animate(self._h, frame, idx.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p))
        return idx, m

    def animate_trs_into(self, frame: int, idx_addr: int, trs_addr: int) -> int:
        return self._sg.sg_animate_trs(self._h, frame, C.c_void_p(idx_addr) if idx_addr else None, C.c_void_p(trs_addr))

    def animate_trs(self, frame: int):
        """(node indices, (n,10) f32 rows: position xyz, rotation ijkw, scale xyz)."""
        n = self.n_units * self.bones_per_unit
        idx = np.empty(n, dtype=np.uint32)
        t = np.empty((n, 10), dtype=np.float32)
        self._sg.sg_animate_trs(self._h, frame, idx.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p))
        return idx, t
