"""The device keeps local / global matrices as their three upper rows and claims (fyx_math.cuh) that this loses
nothing: for affine inputs (bottom row exactly (+0,+0,+0,1), finite entries) the nalgebra-ordered 4x4 product
C[i,j] = ((A[i,0]B[0,j] + A[i,1]B[1,j]) + A[i,2]B[2,j]) + A[i,3]B[3,j] has (a) upper rows equal to the 3-row form that
evaluates `+ A[i,3]*(+0)` for j < 3 and `+ A[i,3]` for j = 3, and (b) a bottom row that is again exactly (+0,+0,+0,1).
Checked in numpy float32 (one rounding per operation) on random matrices seeded with zeros of both signs and denormals."""
import numpy as np

f32 = np.float32


def full_product(A, B):
    """A, B: (n, 4, 4) row-indexed [i, j]; nalgebra's accumulation order per element."""
    C = np.empty_like(A)
    for i in range(4):
        for j in range(4):
            C[:, i, j] = ((A[:, i, 0] * B[:, 0, j] + A[:, i, 1] * B[:, 1, j]) + A[:, i, 2] * B[:, 2, j]) + A[:, i, 3] * B[:, 3, j]
    return C


def rows_product(A, B):
    """affine_mul_row of fyx_math.cuh on the three upper rows"""
    C = np.empty((len(A), 3, 4), f32)
    for i in range(3):
        z = A[:, i, 3] * f32(0.0)
        for j in range(3):
            C[:, i, j] = ((A[:, i, 0] * B[:, 0, j] + A[:, i, 1] * B[:, 1, j]) + A[:, i, 2] * B[:, 2, j]) + z
        C[:, i, 3] = ((A[:, i, 0] * B[:, 0, 3] + A[:, i, 1] * B[:, 1, 3]) + A[:, i, 2] * B[:, 2, 3]) + A[:, i, 3]
    return C


def random_affine(rng, n):
    M = (rng.normal(size=(n, 4, 4)) * 10.0 ** rng.uniform(-3, 3, (n, 1, 1))).astype(f32)
    special = np.array([0.0, -0.0, 1.0, -1.0, 1e-45, -1e-45, 1e-38], f32)
    pick = rng.random((n, 4, 4)) < 0.3
    M = np.where(pick, special[rng.integers(0, len(special), (n, 4, 4))], M)
    M[:, 3, :] = np.array([0.0, 0.0, 0.0, 1.0], f32)
    return M


def test_three_row_product_is_the_full_product_and_the_bottom_row_survives():
    rng = np.random.default_rng(11)
    n = 400_000
    A, B = random_affine(rng, n), random_affine(rng, n)
    with np.errstate(under="ignore"):
        C = full_product(A, B)
        R = rows_product(A, B)
        assert np.array_equal(C[:, :3, :].view(np.uint32), R.view(np.uint32))
        bottom = np.array([0.0, 0.0, 0.0, 1.0], f32).view(np.uint32)
        assert (C[:, 3, :].view(np.uint32) == bottom).all()
        # and through a chain (parent * local * local ...): the property is closed under the product
        D = full_product(C, random_affine(rng, n))
        assert (D[:, 3, :].view(np.uint32) == bottom).all()
        # identity parent (root / orphan): I * L equals L up to the sign of zeros (-0 entries may become +0), as DESIGN says
        I = np.tile(np.eye(4, dtype=f32), (n, 1, 1))
        IL = full_product(I, A)
        assert np.array_equal(IL, A) and not np.array_equal(IL.view(np.uint32), A.view(np.uint32))
