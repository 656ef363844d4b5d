"""Opt-in permute kernels with several loads in flight per thread (`tnr_set_option
"permute_unroll"` = 2 | 4, `"permute_tile"` = 32 | 48 | 64: copy_tiled_mlp_kernel<U>, copy_rows_kernel<double2>, csrc/permute.cu)
against numpy and against the default kernels -- pure data movement, bit exact.  Covers ragged
tiles (extents that are no multiple of the 96-element composite run), odd extents (no 16-byte
path), the equal-fastest-leg case and the chi = 24 rotation of hotrg3d.jl:134."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    ((5,), (0,)), ((4, 7), (1, 0)), ((33, 65), (1, 0)), ((3, 4, 5), (2, 0, 1)),
    ((2, 3, 4, 5), (1, 3, 0, 2)), ((6, 5, 4, 3, 2, 7), (5, 3, 1, 2, 0, 4)),
    ((4, 4, 8, 8, 8, 8), (0, 1, 3, 2, 5, 4)), ((8, 8, 8, 8, 8, 8), (1, 3, 5, 0, 2, 4)),
    ((24, 24, 24), (2, 1, 0)), ((1, 9, 1, 4), (3, 2, 1, 0)), ((40, 3, 40), (0, 2, 1)),
    ((12,) * 6, (5, 3, 1, 2, 0, 4)), ((12,) * 6, (0, 5, 4, 3, 1, 2)), ((10, 7, 6, 9), (0, 3, 2, 1)),
    ((100, 130), (1, 0)), ((97, 101), (1, 0)), ((50, 2, 50), (2, 1, 0)), ((200, 3, 5), (1, 2, 0)),
    ((24, 24, 24, 24), (3, 2, 1, 0)), ((16, 6, 16, 6), (2, 3, 0, 1)),
]


@pytest.mark.parametrize("unroll,tile", [(2, 96), (4, 96), (4, 64), (1, 48), (4, 32)])
@pytest.mark.parametrize("dims,perm", CASES)
def test_permute_unrolled_kernels_bit_exact(tk, ctx, dims, perm, unroll, tile):
    rng = np.random.default_rng(len(dims) * 100 + unroll)
    a = rng.standard_normal(dims)
    T = tk.DeviceTensor.from_numpy(a)
    ctx.set_option("permute_unroll", unroll)
    ctx.set_option("permute_tile", tile)
    try:
        got = T.permute(perm).to_numpy()
    finally:
        ctx.set_option("permute_unroll", 1)
        ctx.set_option("permute_tile", 96)
    assert np.array_equal(got, np.transpose(a, perm))
    assert np.array_equal(got, T.permute(perm).to_numpy())


def test_permute_unroll_option_is_validated(tk, ctx):
    with pytest.raises(tk.TNRCudaError):
        ctx.set_option("permute_unroll", 3)
    with pytest.raises(tk.TNRCudaError):
        ctx.set_option("permute_tile", 100)
    ctx.set_option("permute_unroll", 1)
    ctx.set_option("permute_tile", 96)


def test_hotrg3d_step_with_unrolled_permutes(tk, ctx):
    """The whole HOTRG_3D step on the opt-in kernels: same norm list, bit for bit."""
    T = tk.classical_ising_3D(tk.Trivial)
    base = tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(6), tk.maxiter(3), verbosity=0)
    ctx.set_option("permute_unroll", 4)
    try:
        got = tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(6), tk.maxiter(3), verbosity=0)
    finally:
        ctx.set_option("permute_unroll", 1)
    assert got == base
