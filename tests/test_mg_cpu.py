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
        np.save(os.path.join(out_dir, f"z{rank}.npy"), z.numpy())
        np.save(os.path.join(out_dir, f"dot{rank}.npy"), part.numpy())
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
    dots = [float(np.load(tmp_path / f"dot{r}.npy")[0]) for r in range(world)]
    assert all(d == dots[0] for d in dots)
    assert abs(dots[0] - float(np.dot(want, want))) <= 1e-12 * float(np.dot(want, want))


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
