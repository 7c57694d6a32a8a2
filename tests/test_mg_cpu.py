"""Host-side logic of the row-partitioned multi-GPU layer (spgpu_b200/mg.py) on
CPU: world_size-2 (and 3) process groups over gloo; the local block SpMV is done
by the CPU oracle (test infrastructure), the exchange by gloo send/recv.  The
assembled result must equal the single-block oracle result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spgpu_b200 import formats as F, generators as G, mg
from tests import util


def test_row_blocks_are_hack_aligned_and_cover():
    for nrows, world, hs in [(1000, 3, 32), (64, 2, 32), (4096, 8, 64), (31, 2, 32), (100000, 7, 32)]:
        blocks = mg.row_blocks(nrows, world, hs)
        assert blocks[0][0] == 0 and blocks[-1][1] == nrows
        for (a, b), (c, d) in zip(blocks, blocks[1:]):
            assert b == c and b % hs == 0
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 2 * hs or nrows < world * hs


def test_split_hell_remaps_into_x_ext():
    coo = G.laplace3d_7pt(8)                     # 512 rows, plane = 64
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    x = G.random_vector(512, np.float64, 1)
    want = util.oracle_spmv("hell", hell, x, None, 1.0, 0.0)
    for world in (2, 4):
        got = np.zeros(512)
        for r in range(world):
            loc = mg.split_hell(hell, world, r, 64)
            x_ext = np.zeros(loc.ext_len)
            a, b = max(0, loc.lo - 64), min(512, loc.hi + 64)
            x_ext[a - (loc.lo - 64): b - (loc.lo - 64)] = x[a:b]
            A = F.Hell(loc.values, loc.indices, loc.hack_offsets, loc.rs, 32, 0, loc.nrows, loc.ext_len, 0)
            got[loc.lo:loc.hi] = util.oracle_spmv("hell", A, x_ext, None, 1.0, 0.0)
        np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        mg.split_hell(hell, 2, 0, 8)             # halo narrower than the stencil reach
    with pytest.raises(ValueError):              # 32-row blocks cannot feed a 64-entry halo
        mg.MgHellSpmv(3, 16, 32, 64, lambda *a: None, mg.HaloExchange(3, 16, 64, "gloo"))


def test_split_hdia_shifts_offsets_into_x_ext():
    """each rank's HDIA block on x_ext = [halo | owned | halo] gives the rows of the global product,
    including the first / last rank (window cells outside the matrix) and a rectangular matrix"""
    for coo, halo in ((G.stencil3d_27pt(8), 64 + 8 + 1), (G.laplace2d_5pt(40, 23), 40),
                      (G.random_coo(300, 340, (0, 6), 3, np.float64, 0), 340)):
        hdia = F.coo_to_hdia(coo, 32)
        x = G.random_vector(coo.ncols, np.float64, 1, -1, 1)
        want = util.oracle_spmv("hdia", hdia, x, None, 1.0, 0.0)
        for world in (1, 2, 3):
            got = np.zeros(coo.nrows)
            for r in range(world):
                loc = mg.split_hdia(hdia, world, r, halo)
                x_ext = np.zeros(loc.ext_len)                      # window entries outside the matrix stay 0
                a, b = max(0, loc.lo - halo), min(coo.ncols, loc.hi + halo)
                x_ext[a - (loc.lo - halo): b - (loc.lo - halo)] = x[a:b]
                A = F.Hdia(loc.values, loc.offsets, loc.hack_offsets, 32, 0, loc.nrows, loc.ext_len)
                got[loc.lo:loc.hi] = util.oracle_spmv("hdia", A, x_ext, None, 1.0, 0.0)
            np.testing.assert_array_equal(got, want)
    with pytest.raises(ValueError):
        mg.split_hdia(F.coo_to_hdia(G.stencil3d_27pt(8), 32), 2, 0, 8)


def test_device_builder_row_block_equals_split_hdia():
    """device_build.hdia_row_block (torch; what bench.py uses on the GPU) == mg.split_hdia (numpy)"""
    from spgpu_b200 import device_build as DB
    full = DB.hdia_stencil27(32, device="cpu")
    host = F.Hdia(full.values.numpy(), full.offsets.numpy(), full.hack_offsets.numpy(), 32, int(full.offsets.numel()),
                  full.nrows, full.ncols)
    halo = 32 * 32 + 32 + 32
    for world in (2, 4):
        for r in range(world):
            want = mg.split_hdia(host, world, r, halo)
            got = DB.hdia_row_block(full, want.lo, want.hi, halo)
            np.testing.assert_array_equal(got.values.numpy(), want.values)
            np.testing.assert_array_equal(got.offsets.numpy(), want.offsets)
            np.testing.assert_array_equal(got.hack_offsets.numpy(), want.hack_offsets)
            assert got.nrows == want.nrows and got.ncols == want.ext_len


def _hdia_worker(rank, world, port, n, overlap, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        halo = n * n + n + 1                                       # reach of the 27-point stencil
        halo = -(-halo // 32) * 32
        coo = G.stencil3d_27pt(n)
        hdia = F.coo_to_hdia(coo, 32)
        loc = mg.split_hdia(hdia, world, rank, halo)
        x = G.random_vector(coo.nrows, np.float64, 12345)
        x_ext = torch.zeros(loc.ext_len, dtype=torch.float64)
        x_ext[halo:halo + loc.nrows] = torch.from_numpy(x[loc.lo:loc.hi])
        z = torch.full((loc.nrows,), float("nan"), dtype=torch.float64)
        O = util.oracle_lib()
        T = util.TYPES["D"]

        def local_spmv(zt, xt, r0, r1):
            # rows [r0, r1) of the block: hacks from r0/32 on, x shifted by r0 (column = row + offset)
            zz, xx = zt.numpy(), xt.numpy()
            O.Dhdiaspmv(zz.ctypes.data + 8 * r0, None, T.scalar(1.0), util.ptr(loc.values), util.ptr(loc.offsets), 32,
                        loc.hack_offsets.ctypes.data + 4 * (r0 // 32), r1 - r0, loc.ext_len - r0,
                        xx.ctypes.data + 8 * r0, T.scalar(0.0))

        op = mg.MgHellSpmv(rank, world, loc.nrows, halo, local_spmv,
                           mg.HaloExchange(rank, world, halo, "gloo"), None, overlap=overlap)
        for _ in range(2):
            op.apply(z, x_ext)
        np.save(os.path.join(out_dir, f"z{rank}.npy"), z.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,overlap", [(2, False), (2, True)])
def test_partitioned_hdia_spmv_over_gloo(tmp_path, world, overlap):
    n = 12
    mp.spawn(_hdia_worker, args=(world, _free_port(), n, overlap, str(tmp_path)), nprocs=world, join=True)
    coo = G.stencil3d_27pt(n)
    hdia = F.coo_to_hdia(coo, 32)
    x = G.random_vector(coo.nrows, np.float64, 12345)
    want = util.oracle_spmv("hdia", hdia, x, None, 1.0, 0.0)
    got = np.concatenate([np.load(tmp_path / f"z{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, want)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, overlap, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plane = n * n
        coo = G.laplace3d_7pt(n)
        hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
        loc = mg.split_hell(hell, world, rank, plane)
        x = G.random_vector(coo.nrows, np.float64, 12345)
        x_ext = torch.zeros(loc.ext_len, dtype=torch.float64)
        x_ext[plane:plane + loc.nrows] = torch.from_numpy(x[loc.lo:loc.hi])   # owned part only
        z = torch.full((loc.nrows,), float("nan"), dtype=torch.float64)
        A = F.Hell(loc.values, loc.indices, loc.hack_offsets, loc.rs, 32, 0, loc.nrows, loc.ext_len, 0)
        O = util.oracle_lib()
        T = util.TYPES["D"]

        def local_spmv(zt, xt, r0, r1):
            zz, xx = zt.numpy(), xt.numpy()
            O.Dhellspmv(zz.ctypes.data + 8 * r0, None, T.scalar(1.0), util.ptr(A.values), util.ptr(A.indices), 32,
                        A.hack_offsets.ctypes.data + 4 * (r0 // 32), A.rs.ctypes.data + 4 * r0, None, 7,
                        r1 - r0, util.ptr(xx), T.scalar(0.0), 0)

        op = mg.MgHellSpmv(rank, world, loc.nrows, plane, local_spmv,
                           mg.HaloExchange(rank, world, plane, "gloo"), None, overlap=overlap)
        for _ in range(2):                      # twice: the exchange must be repeatable
            op.apply(z, x_ext)
        # partitioned dot: z.z summed over ranks
        part = torch.tensor([float(np.dot(z.numpy(), z.numpy()))], dtype=torch.float64)
        mg.global_dot(part)
        amax = mg.global_amax(torch.tensor([float(np.abs(z.numpy()).max())], dtype=torch.float64))
        nrm2 = mg.global_nrm2(torch.tensor([float(np.dot(z.numpy(), z.numpy()))], dtype=torch.float64))
        np.save(os.path.join(out_dir, f"z{rank}.npy"), z.numpy())
        np.save(os.path.join(out_dir, f"dot{rank}.npy"), np.array([part.item(), amax.item(), nrm2.item()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,overlap", [(2, False), (2, True), (3, True)])
def test_partitioned_spmv_over_gloo(tmp_path, world, overlap):
    n = 12 if world == 3 else 8
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, overlap, str(tmp_path)), nprocs=world, join=True)
    coo = G.laplace3d_7pt(n)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    x = G.random_vector(coo.nrows, np.float64, 12345)
    want = util.oracle_spmv("hell", hell, x, None, 1.0, 0.0)
    got = np.concatenate([np.load(tmp_path / f"z{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, want)
    red = [np.load(tmp_path / f"dot{r}.npy") for r in range(world)]
    dots = [float(r[0]) for r in red]
    assert all(d == dots[0] for d in dots)
    assert abs(dots[0] - float(np.dot(want, want))) <= 1e-12 * float(np.dot(want, want))
    assert all(float(r[1]) == float(np.abs(want).max()) for r in red)                   # amax: max over ranks
    assert all(abs(float(r[2]) - float(np.linalg.norm(want))) <= 1e-12 * float(np.linalg.norm(want)) for r in red)


def _ag_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coo = G.random_coo(1000, 1000, (0, 12), 77, np.float64, 1)      # unstructured, base 1
        hell = F.ell_to_hell(F.coo_to_ell(coo, 1), 32)
        loc = mg.split_hell_allgather(hell, world, rank)
        x = G.random_vector(1000, np.float64, 5)
        own = torch.zeros(loc.widest, dtype=torch.float64)
        own[:loc.nrows] = torch.from_numpy(x[loc.lo:loc.hi])
        x_full = torch.zeros(world * loc.widest, dtype=torch.float64)
        z = torch.full((loc.nrows,), float("nan"), dtype=torch.float64)
        O = util.oracle_lib()
        T = util.TYPES["D"]

        def local_spmv(zt, xt):
            O.Dhellspmv(util.ptr(zt.numpy()), None, T.scalar(1.0), util.ptr(loc.values), util.ptr(loc.indices), 32,
                        util.ptr(loc.hack_offsets), util.ptr(loc.rs), None, 6, loc.nrows, util.ptr(xt.numpy()),
                        T.scalar(0.0), 1)

        mg.MgAllGatherSpmv(world, loc.widest, local_spmv).apply(z, own, x_full)
        np.save(os.path.join(out_dir, f"z{rank}.npy"), z.numpy())
    finally:
        dist.destroy_process_group()


def test_allgather_mode_for_unstructured_columns(tmp_path):
    world = 3
    mp.spawn(_ag_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    coo = G.random_coo(1000, 1000, (0, 12), 77, np.float64, 1)
    hell = F.ell_to_hell(F.coo_to_ell(coo, 1), 32)
    x = G.random_vector(1000, np.float64, 5)
    want = util.oracle_spmv("hell", hell, x, None, 1.0, 0.0)
    got = np.concatenate([np.load(tmp_path / f"z{r}.npy") for r in range(world)])
    np.testing.assert_array_equal(got, want)


def _plan(L, hell, world):
    import ctypes
    bounds = (ctypes.c_int * (world + 1))()
    halo, allg = ctypes.c_int(-7), ctypes.c_int(-7)
    idx = np.ascontiguousarray(hell.indices, np.int32)
    hoff = np.ascontiguousarray(hell.hack_offsets, np.int32)
    rs = np.ascontiguousarray(hell.rs, np.int32)
    st = L.spgpuMgHellPlan(world, idx.ctypes.data, hell.hack_size, hoff.ctypes.data, rs.ctypes.data, hell.nrows,
                           hell.base, bounds, ctypes.byref(halo), ctypes.byref(allg))
    return st, list(bounds), halo.value, allg.value


def test_c_partition_plan_agrees_with_the_python_partition():
    """spgpuMgHellPlan (the host arithmetic of spgpuMg?hellCreate, include/spgpu_mg.h) needs no device: its row
    blocks are mg.row_blocks', its halo is the furthest reach rounded up to 32 and is accepted by mg.split_hell,
    and a pattern mg.split_hell refuses at every width is planned as all-gather."""
    from spgpu_b200 import capi
    L = capi.lib()
    cases = [(G.laplace3d_7pt(8), 32, 0), (G.laplace3d_7pt(16), 32, 1), (G.stencil3d_27pt(8), 64, 0),
             (G.laplace2d_5pt(40, 23), 32, 0), (G.laplace2d_5pt(33, 31), 96, 1)]
    for coo, hs, base in cases:
        if base != coo.base:
            coo = F.Coo(coo.rows + (base - coo.base), coo.cols + (base - coo.base), coo.vals, coo.nrows, coo.ncols, base)
        hell = F.ell_to_hell(F.coo_to_ell(coo, ell_base=base), hs)
        reach = int(np.abs(coo.rows.astype(np.int64) - coo.cols.astype(np.int64)).max())
        for world in (1, 2, 3, 5):
            st, bounds, halo, allg = _plan(L, hell, world)
            assert st == 0
            assert bounds == [a for a, _ in mg.row_blocks(hell.nrows, world, hs)] + [hell.nrows]
            if world == 1:
                assert (halo, allg) == (0, 0)
                continue
            smallest = min(b - a for a, b in zip(bounds, bounds[1:]))
            # the reach OUT of a block is at most the bandwidth; blocks smaller than the halo force all-gather
            if allg:
                assert halo == 0 and smallest < ((reach + 31) // 32) * 32
                with pytest.raises(ValueError):
                    for r in range(world):
                        mg._check_block(bounds[r], bounds[r + 1], ((reach + 31) // 32) * 32, world)
                        mg.split_hell(hell, world, r, ((reach + 31) // 32) * 32)
            else:
                assert halo % 32 == 0 and 0 < halo <= ((reach + 31) // 32) * 32 and smallest >= halo
                for r in range(world):
                    mg.split_hell(hell, world, r, halo)          # accepted: banded within the planned width
                if halo > 32:
                    with pytest.raises(ValueError):              # and no narrower multiple of 32 would do
                        for r in range(world):
                            mg.split_hell(hell, world, r, halo - 32)


def test_c_partition_plan_of_an_unstructured_pattern_is_all_gather():
    from spgpu_b200 import capi
    L = capi.lib()
    coo = G.random_coo(640, 640, (1, 9), 5, np.float64, 0)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    for world in (3, 4):
        st, bounds, halo, allg = _plan(L, hell, world)
        assert st == 0 and allg == 1 and halo == 0 and bounds[-1] == 640
    # two ranks: every column is at most one whole block away, so the halo is the neighbour's entire block
    st, bounds, halo, allg = _plan(L, hell, 2)
    assert (st, allg) == (0, 0) and 0 < halo <= 320
    for r in range(2):
        mg.split_hell(hell, 2, r, halo)
    assert _plan(L, hell, 0)[0] != 0 and _plan(L, hell, 17)[0] != 0
    bad = F.Hell(hell.values, hell.indices, hell.hack_offsets, hell.rs, 48, hell.height, hell.nrows, hell.ncols, 0)
    assert _plan(L, bad, 2)[0] != 0                              # hackSize must be a multiple of 32


def test_c_hdia_partition_plan_agrees_with_the_python_partition():
    """spgpuMgHdiaPlan: same row blocks, the halo is the furthest NON-ZERO cell outside its block rounded up to 32
    (mg.split_hdia accepts it and refuses anything 32 narrower), blocks smaller than the halo do not fit"""
    import ctypes
    from spgpu_b200 import capi
    L = capi.lib()
    for coo, hs in ((G.stencil3d_27pt(8), 32), (G.laplace2d_5pt(40, 23), 32), (G.laplace3d_7pt(12), 64)):
        for dtype in (np.float64, np.complex64):
            hdia = F.coo_to_hdia(F.Coo(coo.rows, coo.cols, coo.vals.astype(dtype), coo.nrows, coo.ncols, coo.base), hs)
            vals = np.ascontiguousarray(hdia.values)
            offs = np.ascontiguousarray(hdia.offsets, np.int32)
            hoff = np.ascontiguousarray(hdia.hack_offsets, np.int32)
            for world in (1, 2, 3, 6):
                bounds = (ctypes.c_int * (world + 1))()
                halo, fits = ctypes.c_int(-1), ctypes.c_int(-1)
                st = L.spgpuMgHdiaPlan(world, capi.TYPES[util.sym_of(dtype)].code, vals.ctypes.data, offs.ctypes.data, hs,
                                       hoff.ctypes.data, hdia.nrows, hdia.ncols, bounds, ctypes.byref(halo), ctypes.byref(fits))
                assert st == 0
                assert list(bounds) == [a for a, _ in mg.row_blocks(hdia.nrows, world, hs)] + [hdia.nrows]
                if world == 1:
                    assert (halo.value, fits.value) == (0, 1)
                    continue
                reach = int(np.abs(coo.rows.astype(np.int64) - coo.cols.astype(np.int64)).max())
                smallest = min(b - a for a, b in zip(bounds, list(bounds)[1:]))
                if not fits.value:
                    assert smallest < halo.value
                    continue
                assert halo.value % 32 == 0 and 0 < halo.value <= ((reach + 31) // 32) * 32 and smallest >= halo.value
                for r in range(world):
                    mg.split_hdia(hdia, world, r, halo.value)
                if halo.value > 32:
                    with pytest.raises(ValueError):
                        for r in range(world):
                            mg.split_hdia(hdia, world, r, halo.value - 32)
    assert L.spgpuMgHdiaPlan(2, capi.TYPES["D"].code, vals.ctypes.data, offs.ctypes.data, 48, hoff.ctypes.data, 10, 10,
                             bounds, ctypes.byref(halo), ctypes.byref(fits)) != 0


def test_fused_kernel_block_schedule_covers_every_block_once_and_marks_every_boundary_row():
    """spgpuHaloBlockPlan / spgpuHaloBlockOf (csrc/ext_halo.cu: the arithmetic the fused SpMV + halo kernel runs on):
    for a sweep of (rows, haloN, neighbours) -- rows not a multiple of 128, haloN not a multiple of 128, blocks that
    read BOTH zones, a block as small as its halo -- the CTA -> row-block map is a permutation, every block holding a
    row of [0, haloN) waits for the lower zone, every block holding a row of [rows - haloN, rows) for the upper one
    (round 1's schedule missed one: ADVICE r1 'high'), no interior block waits, and the boundary blocks come after
    the `early` interior ones"""
    import ctypes
    from spgpu_b200 import capi
    L = capi.lib()
    rng = np.random.default_rng(3)
    cases = [(1000, 128), (1000, 96), (872, 128), (4096, 4096), (4096, 2048), (130, 64), (129, 129), (128, 128), (1, 1),
             (262144 * 64, 262144), (5037, 128), (64, 32), (0, 0), (300, 0)]
    cases += [(int(r), int(h)) for r, h in zip(rng.integers(1, 200000, 60), rng.integers(1, 3000, 60))]
    for rows, halo in cases:
        halo = min(halo, rows)                      # a rank with a neighbour owns at least haloN rows
        for lo, hi in ((1, 1), (1, 0), (0, 1), (0, 0)):
            for sms in (148, 4):
                plan = (ctypes.c_uint * 6)()
                assert L.spgpuHaloBlockPlan(rows, halo, lo, hi, sms, plan) == 0
                head, first_hi, early, n_lo, n_hi, hi_start = list(plan)
                blocks = -(-rows // 128)
                if blocks <= 20000:
                    order = [L.spgpuHaloBlockOf(plan, c) for c in range(blocks)]
                    assert sorted(order) == list(range(blocks)), (rows, halo, lo, hi)
                    pos = {b: c for c, b in enumerate(order)}
                else:                                  # the 512^3 slab: spot-check the map instead of enumerating 2 M blocks
                    probe = [0, 1, early - 1, early, early + n_lo - 1, early + n_lo, early + n_lo + n_hi - 1,
                             early + n_lo + n_hi, blocks - 1]
                    order = None
                    seen = {L.spgpuHaloBlockOf(plan, c) for c in probe if 0 <= c < blocks}
                    assert len(seen) == len({c for c in probe if 0 <= c < blocks}) and max(seen) < blocks
                need_lo = lambda b: b < head
                need_hi = lambda b: b >= first_hi
                if lo and halo > 0:
                    assert all(need_lo(i // 128) for i in {0, halo - 1, halo // 2})
                    assert head == -(-halo // 128)
                else:
                    assert head == 0
                if hi and halo > 0:
                    assert all(need_hi(i // 128) for i in {rows - halo, rows - 1, rows - 1 - halo // 2})
                    # ... and no block below the first upper-zone row's block waits for it
                    assert first_hi == (rows - halo) // 128
                else:
                    assert first_hi == 0xffffffff
                if order is not None and blocks:
                    boundary = [b for b in range(blocks) if need_lo(b) or need_hi(b)]
                    interior = [b for b in range(blocks) if not (need_lo(b) or need_hi(b))]
                    assert len(interior) == (hi_start - n_lo if (lo or hi) and halo > 0 else blocks) or not boundary
                    if boundary and interior:
                        assert min(pos[b] for b in boundary) >= early
                        assert sorted(pos[b] for b in interior)[:early] == list(range(early))
    assert L.spgpuHaloBlockPlan(-1, 0, 0, 0, 148, (ctypes.c_uint * 6)()) == -1
