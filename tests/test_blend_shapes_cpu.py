"""N4 blend shapes, CPU side: the oracle's restatement of the shader's blend-shape stage (standard.shader:167-173) is checked
against an independent numpy evaluation (numpy's float16 -> float32 conversion is exact, like a texel fetch), over every
binary16 bit pattern incl. subnormals, zeros of both signs, infinities."""
import ctypes as C

import numpy as np

import oracle_binding as ob


def test_blend_shape_stage_of_the_oracle_equals_numpy():
    L = ob.lib()
    rng = np.random.default_rng(5)
    nv, ns = 4096, 3
    # every finite half appears somewhere: 65536 patterns over (3 shapes x 4096 verts x 9) slots
    pats = np.arange(65536, dtype=np.uint16)
    finite = pats[np.isfinite(pats.view(np.float16))]
    rec = finite[rng.integers(0, len(finite), (ns, nv + 7, 9))].astype(np.uint16)  # layer_stride = nv + 7 (padding texels)
    rec.reshape(-1)[: len(finite)] = finite
    weights = np.array([1.0, 0.375, 0.0], np.float32)
    verts = np.zeros((nv, 17), np.float32)
    verts[:, 0:3] = rng.normal(size=(nv, 3))
    verts[:, 5:8] = rng.normal(size=(nv, 3))
    verts[:, 12:16] = np.array([1.0, 0.0, 0.0, 0.0], np.float32)  # one bone, weight 1, index 0
    vbytes = verts.view(np.uint8).reshape(nv, 68).copy()
    vbytes[:, 64:68] = 0
    pal = np.eye(4, dtype=np.float32).reshape(1, 16)  # identity palette: output = morphed input (x*1 + 0... exact)
    pos = np.empty((nv, 3), np.float32)
    nrm = np.empty((nv, 3), np.float32)
    lay = ob.ANIMATED_VERTEX
    with np.errstate(over="ignore", invalid="ignore"):
        L.orc_skin_vertices_blend(ob.fp(pal.reshape(-1)), nv, vbytes.ctypes.data_as(C.c_void_p), C.byref(lay), ns, rec.ctypes.data_as(C.c_void_p), rec.shape[1],
                                  ob.fp(weights), ob.fp(pos.reshape(-1)), ob.fp(nrm.reshape(-1)))
        p = verts[:, 0:3].copy()
        n = verts[:, 5:8].copy()
        for i in range(ns):
            off = rec[i, :nv].view(np.float16).astype(np.float32)
            p = p + off[:, 0:3] * weights[i]
            n = n + off[:, 3:6] * weights[i]
        # identity palette with weight (1,0,0,0): acc = 0 + ((1*x + 0*y) + 0*z + 0) * 1 + three zero terms -> x (up to -0 -> +0)
        want_p, want_n = p + np.float32(0.0), n + np.float32(0.0)
    assert np.array_equal(pos.view(np.uint32), want_p.view(np.uint32))
    assert np.array_equal(nrm.view(np.uint32), want_n.view(np.uint32))
