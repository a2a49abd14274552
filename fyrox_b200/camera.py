"""Host-side observer matrices (inputs of the cull; ≤ dozens per frame, computed on the host like the
reference does in scene/camera.rs:100,459 and renderer/shadow/point.rs:162-178).  f32 numpy."""
from __future__ import annotations

import numpy as np

from .context import frustum_from_view_projection_matrix, mat4_mul

f32 = np.float32


def look_at_rh(eye, target, up) -> np.ndarray:
    """Right-handed look-at view matrix, 16 f32 column-major."""
    eye, target, up = (np.asarray(v, f32) for v in (eye, target, up))
    f = target - eye
    f = f / f32(np.linalg.norm(f))
    s = np.cross(f, up).astype(f32)
    s = s / f32(np.linalg.norm(s))
    u = np.cross(s, f).astype(f32)
    m = np.zeros(16, f32)
    m[0], m[4], m[8], m[12] = s[0], s[1], s[2], -np.dot(s, eye)
    m[1], m[5], m[9], m[13] = u[0], u[1], u[2], -np.dot(u, eye)
    m[2], m[6], m[10], m[14] = -f[0], -f[1], -f[2], np.dot(f, eye)
    m[15] = 1.0
    return m


def perspective(aspect: float, fovy: float, znear: float, zfar: float) -> np.ndarray:
    """nalgebra Perspective3::new, 16 f32 column-major."""
    m = np.zeros(16, f32)
    m22 = f32(1.0) / f32(np.tan(f32(fovy) / f32(2.0)))
    m[5] = m22
    m[0] = m22 / f32(aspect)
    m[10] = (f32(zfar) + f32(znear)) / (f32(znear) - f32(zfar))
    m[14] = f32(zfar) * f32(znear) * f32(2.0) / (f32(znear) - f32(zfar))
    m[11] = -1.0
    return m


def camera_frustum(eye=(0, 0, 0), target=(0, 0, -1), up=(0, 1, 0), aspect=16 / 9, fovy=np.deg2rad(60.0), znear=0.1, zfar=150.0):
    """F=1 observer of SURVEY §8d."""
    return frustum_from_view_projection_matrix(mat4_mul(perspective(aspect, fovy, znear, zfar), look_at_rh(eye, target, up)))


# renderer/utils.rs:49-75: look / up vectors of the six cube-map faces
CUBE_FACES = [((1, 0, 0), (0, -1, 0)), ((-1, 0, 0), (0, -1, 0)), ((0, 1, 0), (0, 0, 1)), ((0, -1, 0), (0, 0, -1)), ((0, 0, 1), (0, -1, 0)), ((0, 0, -1), (0, -1, 0))]


def cube_frusta(origin=(0, 0, 0), radius=120.0):
    """F=6: the point-light shadow pass (renderer/shadow/point.rs:162-178): perspective(1, pi/2, 0.01, R) per face."""
    out = []
    for look, up in CUBE_FACES:
        tgt = tuple(origin[i] + look[i] for i in range(3))
        out.append(camera_frustum(origin, tgt, up, 1.0, np.pi / 2, 0.01, radius))
    return out
