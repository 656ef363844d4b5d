"""Host check (g++, no GPU) of the index arithmetic of the permute path: the planner of
`strided_copy` (flat / rows / tiled decision, merged index groups, composite tiles -- the code
the default kernels run on) and the per-thread read / write phases of the opt-in
copy_tiled_mlp_kernel<U> (`tnrkit.jl_b200/csrc/permute_plan.cuh`), executed block by block and
thread by thread on the CPU against numpy.transpose.  Bit exact: pure data movement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ((5,), (0,)), ((4, 7), (1, 0)), ((33, 65), (1, 0)), ((3, 4, 5), (2, 0, 1)),
    ((2, 3, 4, 5), (1, 3, 0, 2)), ((6, 5, 4, 3, 2, 7), (5, 3, 1, 2, 0, 4)),
    ((4, 4, 8, 8, 8, 8), (0, 1, 3, 2, 5, 4)), ((8, 8, 8, 8, 8, 8), (1, 3, 5, 0, 2, 4)),
    ((24, 24, 24), (2, 1, 0)), ((1, 9, 1, 4), (3, 2, 1, 0)), ((40, 3, 40), (0, 2, 1)),
    ((12,) * 6, (5, 3, 1, 2, 0, 4)), ((12,) * 6, (0, 5, 4, 3, 1, 2)), ((10, 7, 6, 9), (0, 3, 2, 1)),
    ((100, 130), (1, 0)), ((97, 101), (1, 0)), ((50, 2, 50), (2, 1, 0)), ((200, 3, 5), (1, 2, 0)),
    ((24, 24, 24, 24), (3, 2, 1, 0)), ((16, 6, 16, 6), (2, 3, 0, 1)), ((7, 5, 3, 2, 4, 6), (3, 1, 5, 0, 4, 2)),
    ((24,) * 4 + (3, 2), (5, 3, 1, 2, 0, 4)),
]


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("perm") / "libpermute_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "permute_host_shim.cpp")], check=True)
    return C.CDLL(so)


def _run(shim, a, perm, unroll, tile):
    dims = a.shape
    flat = np.ascontiguousarray(np.transpose(a).reshape(-1))          # column-major data
    out = np.full(flat.size, np.nan)
    info = (C.c_longlong * 6)()
    kind = shim.permute_host(flat.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                             len(dims), (C.c_longlong * len(dims))(*dims),
                             (C.c_int * len(perm))(*perm), unroll, tile, info)
    od = tuple(dims[p] for p in perm)
    return kind, np.transpose(out.reshape(tuple(reversed(od)))), list(info)


@pytest.mark.parametrize("unroll,tile", [(1, 96), (2, 96), (4, 96), (4, 64), (2, 48), (4, 32)])
@pytest.mark.parametrize("dims,perm", CASES)
def test_planner_and_tiled_phases_on_host(shim, dims, perm, unroll, tile):
    rng = np.random.default_rng(len(dims) * 10 + tile)
    a = rng.standard_normal(dims)
    kind, got, info = _run(shim, a, perm, unroll, tile)
    assert kind in (1, 2, 3), kind
    assert np.array_equal(got, np.transpose(a, perm))
    if kind == 3:
        assert info[0] <= tile and info[1] <= tile and info[4] % 2 == 1
        assert info[3] <= 100 * 1024          # the dynamic shared memory the kernels opt in to


def test_chi24_rotation_plan(shim):
    """The permutation of hotrg3d.jl:134 at a reduced but structurally identical size (legs of 24
    on the four tile legs): composite 96 x 96 tiles, 768-byte runs on both sides."""
    a = np.random.default_rng(0).standard_normal((24, 24, 2, 3, 24, 24))
    kind, got, info = _run(shim, a, (5, 3, 1, 2, 0, 4), 4, 96)
    assert np.array_equal(got, np.transpose(a, (5, 3, 1, 2, 0, 4)))
    assert kind == 3


def test_planner_fuzz_random_permutations(shim):
    """300 random shapes / permutations (ranks 1..6, legs 1..9 with an occasional long leg) through
    the default plan (tile 96) and the opt-in variants."""
    rng = np.random.default_rng(2024)
    kinds = set()
    for case in range(300):
        rank = int(rng.integers(1, 7))
        dims = [int(rng.integers(1, 10)) for _ in range(rank)]
        if rng.random() < 0.3:
            dims[int(rng.integers(0, rank))] = int(rng.integers(30, 130))
        while np.prod(dims) > 400000:
            dims[int(np.argmax(dims))] //= 2
        perm = [int(x) for x in rng.permutation(rank)]
        a = rng.standard_normal(dims)
        unroll, tile = [(1, 96), (4, 96), (2, 64), (4, 48)][case % 4]
        kind, got, _ = _run(shim, a, perm, unroll, tile)
        kinds.add(kind)
        assert np.array_equal(got, np.transpose(a, perm)), (dims, perm, unroll, tile)
    assert kinds == {1, 2, 3}


def test_strided_copy_sub_blocks_fuzz(shim):
    """`tnr_strided_copy` as the block-sparse layer uses it (symmetric.py: matricize, sym_slice,
    sym_scatter): a permuted sub-block of a larger source array into a sub-block of a larger
    destination array, both with non-compact strides."""
    rng = np.random.default_rng(7)
    for case in range(200):
        rank = int(rng.integers(1, 6))
        big_s = [int(rng.integers(2, 12)) for _ in range(rank)]
        big_d_perm = [int(x) for x in rng.permutation(rank)]
        dims = [int(rng.integers(1, b + 1)) for b in big_s]
        off_s = [int(rng.integers(0, b - d + 1)) for b, d in zip(big_s, dims)]
        # destination: the same legs in permuted order inside a larger array
        big_d = [dims[q] + int(rng.integers(0, 4)) for q in big_d_perm]
        off_d = [int(rng.integers(0, b - dims[q] + 1)) for b, q in zip(big_d, big_d_perm)]
        src = rng.standard_normal(big_s)
        dst = np.full(big_d, np.nan)
        want = dst.copy()
        sl_s = tuple(slice(o, o + d) for o, d in zip(off_s, dims))
        sl_d = tuple(slice(o, o + dims[q]) for o, q in zip(off_d, big_d_perm))
        want[sl_d] = np.transpose(src[sl_s], big_d_perm)
        # column-major element strides
        fs = np.ascontiguousarray(np.transpose(src).reshape(-1))
        fd = np.ascontiguousarray(np.transpose(dst).reshape(-1))
        st_s = np.cumprod([1] + big_s[:-1]).tolist()
        st_d_pos = np.cumprod([1] + big_d[:-1]).tolist()
        st_d = [0] * rank
        for pos, q in enumerate(big_d_perm):
            st_d[q] = st_d_pos[pos]
        base_s = sum(o * s for o, s in zip(off_s, st_s))
        base_d = sum(o * s for o, s in zip(off_d, st_d_pos))
        info = (C.c_longlong * 6)()
        unroll, tile = [(1, 96), (4, 96), (2, 48)][case % 3]
        kind = shim.strided_copy_host(
            C.c_void_p(fs.ctypes.data + 8 * base_s), C.c_void_p(fd.ctypes.data + 8 * base_d), rank,
            (C.c_longlong * rank)(*dims), (C.c_longlong * rank)(*st_s), (C.c_longlong * rank)(*st_d),
            unroll, tile, info)
        assert kind in (1, 2, 3)
        got = np.transpose(fd.reshape(tuple(reversed(big_d))))
        assert np.array_equal(np.isnan(got), np.isnan(want)), (big_s, dims, big_d_perm)
        assert np.array_equal(got[sl_d], want[sl_d]), (big_s, dims, big_d_perm)


def test_rows_16_byte_path_on_host(shim):
    """The equal-fastest-leg copy on double2 elements (`rows_vectorize`): taken for even shared
    runs with even outer strides, declined otherwise; same result either way."""
    rng = np.random.default_rng(3)
    took = 0
    for dims, perm in [((24, 6, 5, 4), (0, 3, 2, 1)), ((24, 24, 24, 24), (0, 3, 1, 2)),
                       ((8, 3, 5), (0, 2, 1)), ((7, 4, 6), (0, 2, 1)), ((12,) * 6, (0, 5, 4, 3, 1, 2)),
                       ((2, 9, 9), (0, 2, 1))]:
        a = rng.standard_normal(dims)
        for unroll in (1, 4):
            kind, got, info = _run(shim, a, perm, unroll, 96)
            assert kind == 2 and np.array_equal(got, np.transpose(a, perm))
            if unroll == 4 and info[5] == -1:
                took += 1
                assert dims[0] % 2 == 0
            if dims[0] % 2 == 1:
                assert info[5] != -1
    assert took >= 4


# ---- the TMA-fed kernel (copy_bulk_kernel): planner + both phases on the host ------------------
# unroll = -1 asks the shim for the bulk path; kind 4 = it ran (memcpy stands in for cp.async.bulk
# and checks the instruction's 16-byte rules and the mbarrier byte count), kind 1-3 = the planner
# declined (odd extents, unaligned pieces) and the default kernels ran.
BULK_CASES = [
    ((24,) * 4 + (2, 3), (5, 3, 1, 2, 0, 4)),          # rotation of hotrg3d.jl:134 (reduced legs)
    ((24, 24, 2, 3, 24, 24), (5, 3, 1, 2, 0, 4)),
    ((12,) * 6, (5, 3, 1, 2, 0, 4)), ((12,) * 6, (0, 5, 4, 3, 1, 2)), ((12,) * 6, (1, 3, 5, 0, 2, 4)),
    ((12,) * 6, (3, 0, 1, 2, 4, 5)), ((12,) * 6, (1, 2, 3, 4, 0, 5)),
    ((24, 6, 5, 4), (0, 3, 2, 1)), ((24, 24, 24, 24), (0, 3, 1, 2)), ((8, 3, 5), (0, 2, 1)),
    ((100, 130), (1, 0)), ((96, 200), (1, 0)), ((98, 102), (1, 0)), ((50, 2, 50), (2, 1, 0)),
    ((24, 24, 24), (2, 1, 0)), ((16, 6, 16, 6), (2, 3, 0, 1)), ((4, 4, 8, 8, 8, 8), (0, 1, 3, 2, 5, 4)),
    ((8,) * 6, (1, 3, 5, 0, 2, 4)), ((400, 6, 10), (0, 2, 1)), ((1000, 4, 6), (0, 2, 1)),
    ((2, 9, 9), (0, 2, 1)), ((6, 10, 14), (2, 1, 0)), ((6, 10, 14), (1, 0, 2)),
    ((48,) * 4, (3, 1, 2, 0)), ((48,) * 4, (0, 3, 2, 1)), ((20, 30, 40), (1, 2, 0)),
]


@pytest.mark.parametrize("dims,perm", BULK_CASES)
def test_bulk_kernel_phases_on_host(shim, dims, perm):
    rng = np.random.default_rng(sum(dims))
    a = rng.standard_normal(dims)
    kind, got, info = _run(shim, a, perm, -1, 96)
    assert kind == 4, (kind, info)
    assert np.array_equal(got, np.transpose(a, perm))
    assert info[3] <= 100 * 1024 and info[4] % 2 == 0 and (info[4] // 2) % 2 == 1


def test_bulk_kernel_declines_odd_extents(shim):
    """Pieces that break the 16-byte rule of cp.async.bulk must fall back, never be issued."""
    rng = np.random.default_rng(11)
    for dims, perm in [((7, 5, 3), (2, 1, 0)), ((97, 101), (1, 0)), ((7, 3, 5), (0, 2, 1)),
                       ((5,), (0,)), ((3, 3, 3, 3), (3, 2, 1, 0))]:
        a = rng.standard_normal(dims)
        kind, got, _ = _run(shim, a, perm, -1, 96)
        assert kind in (1, 2, 3)
        assert np.array_equal(got, np.transpose(a, perm))


def test_bulk_kernel_fuzz(shim):
    """400 random shapes / permutations and 200 permuted sub-block copies with non-compact strides
    (what the block-sparse layer issues): whichever path the planner picks, the result is exact
    and nothing is written outside the destination block; the bulk path must be taken often."""
    rng = np.random.default_rng(99)
    kinds = {}
    for case in range(400):
        rank = int(rng.integers(2, 7))
        dims = [int(rng.integers(1, 7)) * 2 for _ in range(rank)]
        if rng.random() < 0.4:
            dims[int(rng.integers(0, rank))] = int(rng.integers(10, 80)) * 2
        if rng.random() < 0.2:
            dims[int(rng.integers(0, rank))] = int(rng.integers(1, 12))    # maybe odd
        while np.prod(dims) > 300000:
            dims[int(np.argmax(dims))] //= 2
        perm = [int(x) for x in rng.permutation(rank)]
        a = rng.standard_normal(dims)
        kind, got, _ = _run(shim, a, perm, -1, 96)
        kinds[kind] = kinds.get(kind, 0) + 1
        assert kind in (1, 2, 3, 4), (dims, perm, kind)
        assert np.array_equal(got, np.transpose(a, perm)), (dims, perm, kind)
    assert kinds.get(4, 0) >= 200, kinds
    took = 0
    for case in range(200):
        rank = int(rng.integers(1, 6))
        big_s = [int(rng.integers(1, 7)) * 2 for _ in range(rank)]
        big_d_perm = [int(x) for x in rng.permutation(rank)]
        dims = [int(rng.integers(1, b // 2 + 1)) * 2 for b in big_s]
        off_s = [int(rng.integers(0, (b - d) // 2 + 1)) * 2 for b, d in zip(big_s, dims)]
        big_d = [dims[q] + 2 * int(rng.integers(0, 3)) for q in big_d_perm]
        off_d = [int(rng.integers(0, (b - dims[q]) // 2 + 1)) * 2 for b, q in zip(big_d, big_d_perm)]
        if case % 5 == 0:   # odd offsets: 8-byte aligned only
            off_s[0] = min(off_s[0] + 1, big_s[0] - dims[0])
        src = rng.standard_normal(big_s)
        dst = np.full(big_d, np.nan)
        want = dst.copy()
        sl_s = tuple(slice(o, o + d) for o, d in zip(off_s, dims))
        sl_d = tuple(slice(o, o + dims[q]) for o, q in zip(off_d, big_d_perm))
        want[sl_d] = np.transpose(src[sl_s], big_d_perm)
        fs = np.ascontiguousarray(np.transpose(src).reshape(-1))
        fd = np.ascontiguousarray(np.transpose(dst).reshape(-1))
        st_s = np.cumprod([1] + big_s[:-1]).tolist()
        st_d_pos = np.cumprod([1] + big_d[:-1]).tolist()
        st_d = [0] * rank
        for pos, q in enumerate(big_d_perm):
            st_d[q] = st_d_pos[pos]
        base_s = sum(o * s for o, s in zip(off_s, st_s))
        base_d = sum(o * s for o, s in zip(off_d, st_d_pos))
        info = (C.c_longlong * 6)()
        kind = shim.strided_copy_host(
            C.c_void_p(fs.ctypes.data + 8 * base_s), C.c_void_p(fd.ctypes.data + 8 * base_d), rank,
            (C.c_longlong * rank)(*dims), (C.c_longlong * rank)(*st_s), (C.c_longlong * rank)(*st_d),
            -1, 96, info)
        assert kind in (1, 2, 3, 4), (big_s, dims, big_d_perm, kind)
        took += kind == 4
        got = np.transpose(fd.reshape(tuple(reversed(big_d))))
        assert np.array_equal(np.isnan(got), np.isnan(want)), (big_s, dims, big_d_perm, kind)
        assert np.array_equal(got[sl_d], want[sl_d]), (big_s, dims, big_d_perm, kind)
    assert took >= 40, took


def test_bulk_dense_repitch_path_is_taken(shim):
    """Short source rows that follow each other contiguously (2-D transposition with a short
    source-contiguous leg: `Qn -> Qk` and the Gram operand of the HOTRG_3D step): ONE bulk piece
    per j2 slice, re-pitched in shared memory.  info[5] >= 100 marks the dense path."""
    rng = np.random.default_rng(21)
    for dims, perm in [((12,) * 6, (1, 3, 5, 0, 2, 4)), ((12,) * 6, (1, 2, 3, 4, 0, 5)),
                       ((24, 24, 6, 4), (1, 0, 2, 3)), ((24, 200), (1, 0)), ((8, 30, 6), (1, 0, 2)),
                       ((16, 98), (1, 0)), ((62, 40), (1, 0))]:
        a = rng.standard_normal(dims)
        kind, got, info = _run(shim, a, perm, -1, 96)
        assert kind == 4 and info[5] >= 100, (dims, perm, kind, info)
        assert np.array_equal(got, np.transpose(a, perm))
    # long rows keep one piece per row
    a = rng.standard_normal((96, 200))
    kind, got, info = _run(shim, a, (1, 0), -1, 96)
    assert kind == 4 and info[5] < 100 and np.array_equal(got, a.T)
