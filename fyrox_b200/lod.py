"""Host mirror of the reference's LOD groups (fyrox-impl/src/scene/base.rs:61-160) and the host part of the LOD filter:
the resolution of "an object listed in several levels" that RenderDataBundleStorage::from_graph performs implicitly by
writing ``lod_filter[object]`` in a fixed order (renderer/bundle.rs:898-916).  Pure host logic (no GPU needed);
``resolve_lod_ranges`` produces what ``Context.set_lod_ranges`` / ``fyx_set_lod_ranges`` takes.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Tuple

import numpy as np

_f32 = np.float32


def _clamp01(x: float) -> float:
    x = float(_f32(x))
    return 0.0 if x < 0.0 else (1.0 if x > 1.0 else x)  # f32::clamp(0.0, 1.0); NaN stays NaN


class LevelOfDetail:
    """scene/base.rs:61-127"""

    def __init__(self, begin: float, end: float, objects: Iterable[int]):
        objects = list(objects)
        for o in objects:
            assert o is not None and o != 0xFFFFFFFF, "Invalid handles are not allowed"  # base.rs:75-78
        b, e = float(_f32(begin)), float(_f32(end))
        b = min(b, e)  # base.rs:79-80
        e = max(e, b)
        self._begin, self._end = _clamp01(b), _clamp01(e)
        self.objects: List[int] = objects

    def set_begin(self, percent: float):
        self._begin = _clamp01(percent)
        if self._begin > self._end:
            self._begin, self._end = self._end, self._begin

    def begin(self) -> float:
        return self._begin

    def set_end(self, percent: float):
        self._end = _clamp01(percent)
        if self._end < self._begin:
            self._begin, self._end = self._end, self._begin

    def end(self) -> float:
        return self._end


class LodGroup:
    """scene/base.rs:129-160"""

    def __init__(self, levels: Optional[List[LevelOfDetail]] = None):
        self.levels: List[LevelOfDetail] = list(levels or [])


def resolve_lod_ranges(groups: Dict[int, LodGroup], is_alive: Optional[Callable[[int], bool]] = None) -> Tuple[np.ndarray, np.ndarray]:
    """groups: owner node index -> its LodGroup.  Walks the owners in pool (index) order, their levels and objects in order —
    the order of the loop in from_graph — and keeps, per object, the range written last.  Objects that are not alive
    (`try_get_node` fails) are skipped.  Returns (object indices u32[n] ascending, ranges f32[n, 2])."""
    final: Dict[int, Tuple[float, float]] = {}
    for owner in sorted(groups):
        for level in groups[owner].levels:
            for obj in level.objects:
                if is_alive is not None and not is_alive(obj):
                    continue
                final[obj] = (level.begin(), level.end())
    idx = np.array(sorted(final), np.uint32)
    rng = np.array([final[int(i)] for i in idx], np.float32).reshape(-1, 2)
    return idx, rng
