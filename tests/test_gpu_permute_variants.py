"""Every variant of the permute kernels against numpy -- pure data movement, bit exact:
the TMA-fed tiled copy (`copy_bulk_kernel`, cp.async.bulk reads; default whenever the source
pieces are 16-byte aligned), the register-staged fallbacks with U rows of loads in flight
(`tnr_set_option "permute_unroll"` = 1 | 2 | 4 (default), `"permute_tile"` = 32 | 48 | 64 | 96:
copy_tiled_kernel, copy_tiled_mlp_kernel<U>, copy_rows_kernel<double2>, csrc/permute.cu).  Covers ragged
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


def _defaults(ctx):
    ctx.set_option("permute_bulk", 1)
    ctx.set_option("permute_unroll", 4)
    ctx.set_option("permute_tile", 96)


def _bulk_launches(ctx):
    import ctypes as C

    v = C.c_double()
    ctx.call("tnr_get_counter", b"permute_bulk_launches", C.byref(v))
    return int(v.value)


@pytest.mark.parametrize("unroll,tile", [(1, 96), (2, 96), (4, 96), (4, 64), (1, 48), (4, 32)])
@pytest.mark.parametrize("dims,perm", CASES)
def test_permute_unrolled_kernels_bit_exact(tk, ctx, dims, perm, unroll, tile):
    rng = np.random.default_rng(len(dims) * 100 + unroll)
    a = rng.standard_normal(dims)
    T = tk.DeviceTensor.from_numpy(a)
    ctx.set_option("permute_bulk", 0)
    ctx.set_option("permute_unroll", unroll)
    ctx.set_option("permute_tile", tile)
    try:
        got = T.permute(perm).to_numpy()
    finally:
        _defaults(ctx)
    assert np.array_equal(got, np.transpose(a, perm))
    assert np.array_equal(got, T.permute(perm).to_numpy())


BULK_CASES = [
    ((24,) * 4 + (2, 3), (5, 3, 1, 2, 0, 4)), ((24, 24, 2, 3, 24, 24), (5, 3, 1, 2, 0, 4)),
    ((12,) * 6, (5, 3, 1, 2, 0, 4)), ((12,) * 6, (0, 5, 4, 3, 1, 2)), ((12,) * 6, (1, 3, 5, 0, 2, 4)),
    ((12,) * 6, (3, 0, 1, 2, 4, 5)), ((12,) * 6, (1, 2, 3, 4, 0, 5)), ((16,) * 6, (5, 3, 1, 2, 0, 4)),
    ((24, 6, 5, 4), (0, 3, 2, 1)), ((24, 24, 24, 24), (0, 3, 1, 2)), ((8, 3, 5), (0, 2, 1)),
    ((100, 130), (1, 0)), ((96, 200), (1, 0)), ((98, 102), (1, 0)), ((50, 2, 50), (2, 1, 0)),
    ((24, 24, 24), (2, 1, 0)), ((16, 6, 16, 6), (2, 3, 0, 1)), ((4, 4, 8, 8, 8, 8), (0, 1, 3, 2, 5, 4)),
    ((8,) * 6, (1, 3, 5, 0, 2, 4)), ((400, 6, 10), (0, 2, 1)), ((1000, 4, 6), (0, 2, 1)),
    ((2, 9, 9), (0, 2, 1)), ((6, 10, 14), (2, 1, 0)), ((48,) * 4, (3, 1, 2, 0)),
    ((48,) * 4, (0, 3, 2, 1)), ((20, 30, 40), (1, 2, 0)), ((1728, 1730), (1, 0)),
    ((7, 4, 6), (0, 2, 1)),
    # short source rows fetched densely and re-pitched in shared memory
    ((24, 24, 6, 4), (1, 0, 2, 3)), ((24, 200), (1, 0)), ((8, 30, 6), (1, 0, 2)), ((16, 98), (1, 0)),
    ((62, 40), (1, 0)), ((24,) * 4, (1, 3, 0, 2)),
]


@pytest.mark.parametrize("dims,perm", BULK_CASES)
def test_bulk_permute_kernel_bit_exact(tk, ctx, dims, perm):
    """copy_bulk_kernel (TMA reads): the same cases tests/test_permute_host.py runs thread by
    thread on the CPU; here on the device, and the counter proves the bulk kernel ran."""
    rng = np.random.default_rng(sum(dims))
    a = rng.standard_normal(dims)
    T = tk.DeviceTensor.from_numpy(a)
    _defaults(ctx)
    n0 = _bulk_launches(ctx)
    got = T.permute(perm).to_numpy()
    assert _bulk_launches(ctx) == n0 + 1
    assert np.array_equal(got, np.transpose(a, perm))


@pytest.mark.parametrize("dims,perm", [((7, 5, 3), (2, 1, 0)), ((97, 101), (1, 0)),
                                       ((7, 3, 5), (0, 2, 1)), ((3, 3, 3, 3), (3, 2, 1, 0))])
def test_bulk_permute_declines_unaligned(tk, ctx, dims, perm):
    a = np.random.default_rng(1).standard_normal(dims)
    T = tk.DeviceTensor.from_numpy(a)
    _defaults(ctx)
    n0 = _bulk_launches(ctx)
    got = T.permute(perm).to_numpy()
    assert _bulk_launches(ctx) == n0
    assert np.array_equal(got, np.transpose(a, perm))


def test_bulk_permute_fuzz_against_fallback(tk, ctx):
    """120 random even-extent shapes: default path (bulk where eligible) == round-1 kernel."""
    rng = np.random.default_rng(5)
    _defaults(ctx)
    n0 = _bulk_launches(ctx)
    for case in range(120):
        rank = int(rng.integers(2, 7))
        dims = [int(rng.integers(1, 7)) * 2 for _ in range(rank)]
        if rng.random() < 0.5:
            dims[int(rng.integers(0, rank))] = int(rng.integers(10, 120)) * 2
        while np.prod(dims) > 2_000_000:
            dims[int(np.argmax(dims))] //= 2
        perm = [int(x) for x in rng.permutation(rank)]
        a = rng.standard_normal(dims)
        got = tk.DeviceTensor.from_numpy(a).permute(perm).to_numpy()
        assert np.array_equal(got, np.transpose(a, perm)), (dims, perm)
    assert _bulk_launches(ctx) - n0 >= 60


def test_permute_unroll_option_is_validated(tk, ctx):
    with pytest.raises(tk.TNRCudaError):
        ctx.set_option("permute_unroll", 3)
    with pytest.raises(tk.TNRCudaError):
        ctx.set_option("permute_tile", 100)
    _defaults(ctx)


def test_hotrg3d_step_with_unrolled_permutes(tk, ctx):
    """The whole HOTRG_3D step on every permute variant: same norm list, bit for bit."""
    T = tk.classical_ising_3D(tk.Trivial)
    _defaults(ctx)
    base = tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(6), tk.maxiter(3), verbosity=0)
    for bulk, unroll in [(0, 1), (0, 4)]:
        ctx.set_option("permute_bulk", bulk)
        ctx.set_option("permute_unroll", unroll)
        try:
            got = tk.run(tk.HOTRG_3D(T, shard=False), tk.truncrank(6), tk.maxiter(3), verbosity=0)
        finally:
            _defaults(ctx)
        assert got == base
