"""Single-GPU checks of the multi-GPU device code (include/spgpu_ext.h).  Kernels of
different ranks must never wait on each other on ONE GPU, so the neighbours are emulated:
their "pushes" are pre-filled halo zones + pre-set ready flags, and this rank's pushes land
in scratch buffers standing in for the neighbours' memory.  The real multi-rank runs are
bench.py --verify under torchrun (bit-exact against global columns) and the gloo tests."""
import ctypes

import numpy as np
import pytest

from spgpu_b200 import formats as F, generators as G, mg
from tests import util

pytestmark = pytest.mark.gpu


def _middle_block(n=24, world=3, rank=1):
    coo = G.laplace3d_7pt(n)
    plane = n * n
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    loc = mg.split_hell(hell, world, rank, plane)
    x = G.random_vector(coo.nrows, np.float64, 12345)
    x_ext = x[loc.lo - plane: loc.hi + plane].copy()
    want = util.oracle_spmv("hell", hell, x, None, 1.0, 0.0)[loc.lo:loc.hi]
    return coo, loc, plane, x, x_ext, want


@pytest.mark.parametrize("seq", [1, 5])
def test_spmv_fused_with_halo_exchange(ours, gpu_handle, seq):
    import torch
    coo, loc, plane, x, x_ext, want = _middle_block()
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    dz = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
    my_flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    my_flags[0] = seq; my_flags[1] = seq          # both neighbours' halos "have arrived"
    my_flags[2] = seq - 1; my_flags[3] = seq - 1  # and they acknowledged my previous pushes
    peer_lo_flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    peer_hi_flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    peer_lo_halo = torch.full((plane,), float("nan"), dtype=torch.float64, device="cuda")
    peer_hi_halo = torch.full((plane,), float("nan"), dtype=torch.float64, device="cuda")
    ours.spgpuDhellspmvHalo(gpu_handle, dz.data_ptr(), 0, 1.0, dv.data_ptr(), di.data_ptr(), 32, dho.data_ptr(),
                            drs.data_ptr(), 7, loc.nrows, dx.data_ptr(), 0.0, 0, plane,
                            peer_lo_halo.data_ptr(), peer_hi_halo.data_ptr(), my_flags.data_ptr(),
                            peer_lo_flags.data_ptr(), peer_hi_flags.data_ptr(), seq)
    torch.cuda.synchronize()
    util.assert_rows_close(dz.cpu().numpy(), want, np.full(loc.nrows, 12.0), "D", "fused spmv+halo")
    # my first / last owned plane landed in the neighbours' halo zones
    np.testing.assert_array_equal(peer_lo_halo.cpu().numpy(), x_ext[plane:2 * plane])
    np.testing.assert_array_equal(peer_hi_halo.cpu().numpy(), x_ext[loc.nrows:loc.nrows + plane])
    lo_f, hi_f = peer_lo_flags.cpu().numpy(), peer_hi_flags.cpu().numpy()
    assert lo_f[1] == seq and lo_f[3] == seq       # lower neighbour: ready-from-above, ack-from-above
    assert hi_f[0] == seq and hi_f[2] == seq       # upper neighbour: ready-from-below, ack-from-below
    assert lo_f[0] == 0 and hi_f[1] == 0


@pytest.mark.parametrize("seq", [1, 4])
@pytest.mark.parametrize("rank", [0, 1, 2])
def test_hdia_spmv_fused_with_halo_exchange(ours, gpu_handle, seq, rank):
    """spgpuDhdiaspmvHalo on the first / a middle / the last block of a 27-point stencil split in three
    (emulated neighbours): rows equal the global product, boundary entries land in the neighbours'
    halo zones, flags carry seq; also equal to the plain spgpuDhdiaspmv on the same block bit for bit"""
    import torch
    n, world = 16, 3
    coo = G.stencil3d_27pt(n)
    hdia = F.coo_to_hdia(coo, 32)
    halo = -(-(n * n + n + 1) // 32) * 32
    loc = mg.split_hdia(hdia, world, rank, halo)
    x = G.random_vector(coo.nrows, np.float64, 12345, -1, 1)
    y = G.random_vector(coo.nrows, np.float64, 777, -1, 1)
    x_ext = np.zeros(loc.ext_len)
    a, b = max(0, loc.lo - halo), min(coo.nrows, loc.hi + halo)
    x_ext[a - (loc.lo - halo): b - (loc.lo - halo)] = x[a:b]
    want = util.oracle_spmv("hdia", hdia, x, y, 1.5, -0.5)[loc.lo:loc.hi]
    dv, doff, dho = (util.to_dev(t) for t in (loc.values, loc.offsets, loc.hack_offsets))
    dx, dy = util.to_dev(x_ext), util.to_dev(y[loc.lo:loc.hi])
    dz = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
    has_lo, has_hi = rank > 0, rank < world - 1
    my_flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    my_flags[0] = seq; my_flags[1] = seq
    my_flags[2] = seq - 1; my_flags[3] = seq - 1
    pf = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(2)]
    ph = [torch.full((halo,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(2)]
    ours.spgpuDhdiaspmvHalo(gpu_handle, dz.data_ptr(), dy.data_ptr(), 1.5, dv.data_ptr(), doff.data_ptr(), 32, dho.data_ptr(),
                            loc.nrows, loc.ext_len, dx.data_ptr(), -0.5, halo,
                            ph[0].data_ptr() if has_lo else 0, ph[1].data_ptr() if has_hi else 0, my_flags.data_ptr(),
                            pf[0].data_ptr() if has_lo else 0, pf[1].data_ptr() if has_hi else 0, seq)
    torch.cuda.synchronize()
    scale = util.row_scale(coo, x, y, 1.5, -0.5)[loc.lo:loc.hi]
    util.assert_rows_close(dz.cpu().numpy(), want, scale, "D", "fused hdia spmv+halo")
    if has_lo:
        np.testing.assert_array_equal(ph[0].cpu().numpy(), x_ext[halo:2 * halo])
        f = pf[0].cpu().numpy()
        assert f[1] == seq and f[3] == seq and f[0] == 0
    if has_hi:
        np.testing.assert_array_equal(ph[1].cpu().numpy(), x_ext[loc.nrows:loc.nrows + halo])
        f = pf[1].cpu().numpy()
        assert f[0] == seq and f[2] == seq and f[1] == 0
    dz2 = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
    T = util.TYPES["D"]
    ours.spgpuDhdiaspmv(gpu_handle, dz2.data_ptr(), dy.data_ptr(), T.scalar(1.5), dv.data_ptr(), doff.data_ptr(), 32,
                        dho.data_ptr(), loc.nrows, loc.ext_len, dx.data_ptr(), T.scalar(-0.5))
    torch.cuda.synchronize()
    assert torch.equal(dz, dz2)


def test_spmv_fused_without_neighbours_is_the_plain_kernel(ours, gpu_handle):
    import torch
    coo = G.laplace3d_7pt(20)
    A = F.ell_to_hell(F.coo_to_ell(coo), 32)
    dA = util.upload(A)
    x = G.random_vector(coo.nrows, np.float64, 3)
    y = G.random_vector(coo.nrows, np.float64, 4)
    plain = util.dev_spmv(ours, gpu_handle, "hell", A, dA, x, y, 1.5, -0.5)
    dx, dy = util.to_dev(x), util.to_dev(y)
    dz = torch.full((coo.nrows,), float("nan"), dtype=torch.float64, device="cuda")
    flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    ours.spgpuDhellspmvHalo(gpu_handle, dz.data_ptr(), dy.data_ptr(), 1.5, dA["values"].data_ptr(), dA["indices"].data_ptr(),
                            32, dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), 7, coo.nrows, dx.data_ptr(), -0.5, 0, 0,
                            0, 0, flags.data_ptr(), 0, 0, 1)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dz.cpu().numpy(), plain)


def test_halo_exchange_kernel_and_ack(ours, gpu_handle):
    """spgpuDhaloExchange / spgpuHaloAck / spgpuDhaloPush / spgpuWaitFlag with emulated neighbours"""
    import torch
    n = 5000
    src = torch.arange(3 * n, dtype=torch.float64, device="cuda")
    dst_lo = torch.zeros(n, dtype=torch.float64, device="cuda")
    dst_hi = torch.zeros(n, dtype=torch.float64, device="cuda")
    mine = torch.zeros(16, dtype=torch.int32, device="cuda")
    peer = torch.zeros(16, dtype=torch.int32, device="cuda")
    mine[0] = 1; mine[1] = 1                       # neighbours already signalled seq 1
    m, p = mine.data_ptr(), peer.data_ptr()
    ours.spgpuDhaloExchange(gpu_handle, dst_lo.data_ptr(), src.data_ptr(), dst_hi.data_ptr(), src.data_ptr() + 8 * 2 * n, n,
                            m + 8, m + 12, p + 4, p + 0, m + 0, m + 4, 1)
    ours.spgpuHaloAck(gpu_handle, p + 12, p + 8, 1)
    torch.cuda.synchronize()
    assert torch.equal(dst_lo, src[:n]) and torch.equal(dst_hi, src[2 * n:])
    assert peer.cpu().numpy()[:4].tolist() == [1, 1, 1, 1]
    # the two-kernel form
    dst = torch.zeros(n + 1, dtype=torch.float64, device="cuda")
    flag = torch.zeros(4, dtype=torch.int32, device="cuda")
    ours.spgpuDhaloPush(gpu_handle, dst.data_ptr() + 8, src.data_ptr() + 8, n, flag.data_ptr(), 7)   # unaligned -> scalar path
    ours.spgpuWaitFlag(gpu_handle, flag.data_ptr(), 7)
    torch.cuda.synchronize()
    assert torch.equal(dst[1:], src[1:n + 1]) and int(flag[0].item()) == 7


def test_peer_allreduce_kernel_with_emulated_peer(ours, gpu_handle):
    """spgpuAllreduceSumDev, world = 3, this rank = 1: the two peers' contributions are pre-filled in
    this rank's table; this rank's (value, seq) must land in slot 1 of every table and the sum must
    be taken in rank order"""
    import ctypes
    import struct
    import torch
    world, me = 3, 1
    for seq in (1, 2, 7):
        tabs = [torch.zeros(2 * world * 4, dtype=torch.int32, device="cuda") for _ in range(world)]
        parity = seq & 1
        vals = {0: 0.125, 2: -3.5}
        host = np.zeros(2 * world * 4, dtype=np.int32)
        for r, v in vals.items():
            lo, hi = struct.unpack("<ii", struct.pack("<d", v))
            base = (parity * world + r) * 4
            host[base:base + 4] = [lo, hi, seq, 0]
        tabs[me].copy_(torch.from_numpy(host))
        value = torch.tensor([10.0], dtype=torch.float64, device="cuda")
        ptrs = (ctypes.c_void_p * world)(*[t.data_ptr() for t in tabs])
        ours.spgpuAllreduceSumDev(gpu_handle, value.data_ptr(), world, me, ptrs, seq)
        torch.cuda.synchronize()
        assert float(value.item()) == (0.125 + 10.0) + -3.5
        for t in tabs:
            got = t.cpu().numpy()[(parity * world + me) * 4:(parity * world + me) * 4 + 3]
            assert struct.unpack("<d", struct.pack("<ii", int(got[0]), int(got[1])))[0] == 10.0 and got[2] == seq


def test_spmv_halo_dot(ours, gpu_handle):
    """spgpuDhellspmvHaloDot: fused exchange + SpMV + this rank's share of p.Ap"""
    import torch
    coo, loc, plane, x, x_ext, want = _middle_block()
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    dz = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
    my_flags = torch.zeros(16, dtype=torch.int32, device="cuda")
    my_flags[0] = 1; my_flags[1] = 1
    pf = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(2)]
    ph = [torch.zeros(plane, dtype=torch.float64, device="cuda") for _ in range(2)]
    dres = torch.full((1,), float("nan"), dtype=torch.float64, device="cuda")
    ours.spgpuDhellspmvHaloDot(gpu_handle, dz.data_ptr(), dv.data_ptr(), di.data_ptr(), 32, dho.data_ptr(), drs.data_ptr(),
                               7, loc.nrows, dx.data_ptr(), 0, plane, ph[0].data_ptr(), ph[1].data_ptr(),
                               my_flags.data_ptr(), pf[0].data_ptr(), pf[1].data_ptr(), 1, dres.data_ptr())
    torch.cuda.synchronize()
    util.assert_rows_close(dz.cpu().numpy(), want, np.full(loc.nrows, 12.0), "D", "fused spmv+halo+dot")
    own = x_ext[plane:plane + loc.nrows]
    ref = float(np.dot(own, want))
    assert abs(float(dres.item()) - ref) <= 1e-12 * float(np.sum(np.abs(own) * np.abs(want)))


def test_device_resident_sequence_numbers(ours, gpu_handle):
    """seq == 0: the fused halo kernel and the all-reduce take their sequence number from device counters
    (spgpuSetSeqCounters); spgpuHaloSeqAdvance / the all-reduce itself advance them -- what lets a partitioned
    iteration be replayed from a CUDA graph.  Emulated neighbours as above."""
    import struct
    import torch
    coo, loc, plane, x, x_ext, want = _middle_block()
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    counters = torch.tensor([6, 10], dtype=torch.int32, device="cuda")       # 6 exchanges, 10 all-reduces done so far
    assert ours.spgpuSetSeqCounters(gpu_handle, counters.data_ptr(), counters.data_ptr() + 4) == 0
    try:
        my_flags = torch.zeros(16, dtype=torch.int32, device="cuda")
        pf = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(2)]
        ph = [torch.zeros(plane, dtype=torch.float64, device="cuda") for _ in range(2)]
        for k in (7, 8):                                                     # two exchanges: 7 and 8
            my_flags[0] = k; my_flags[1] = k; my_flags[2] = k - 1; my_flags[3] = k - 1
            dz = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
            ours.spgpuDhellspmvHalo(gpu_handle, dz.data_ptr(), 0, 1.0, dv.data_ptr(), di.data_ptr(), 32, dho.data_ptr(),
                                    drs.data_ptr(), 7, loc.nrows, dx.data_ptr(), 0.0, 0, plane, ph[0].data_ptr(),
                                    ph[1].data_ptr(), my_flags.data_ptr(), pf[0].data_ptr(), pf[1].data_ptr(), 0)
            ours.spgpuHaloSeqAdvance(gpu_handle)
            torch.cuda.synchronize()
            util.assert_rows_close(dz.cpu().numpy(), want, np.full(loc.nrows, 12.0), "D", "fused spmv+halo, device seq")
            assert int(counters[0].item()) == k
            assert pf[0].cpu().numpy()[1] == k and pf[0].cpu().numpy()[3] == k
            assert pf[1].cpu().numpy()[0] == k and pf[1].cpu().numpy()[2] == k
        # all-reduce number 11, world 2, this rank 0: the peer's slot is pre-filled
        world, me, seq = 2, 0, 11
        tabs = [torch.zeros(2 * world * 4, dtype=torch.int32, device="cuda") for _ in range(world)]
        host = np.zeros(2 * world * 4, dtype=np.int32)
        lo, hi = struct.unpack("<ii", struct.pack("<d", 2.5))
        base = ((seq & 1) * world + 1) * 4
        host[base:base + 4] = [lo, hi, seq, 0]
        tabs[me].copy_(torch.from_numpy(host))
        value = torch.tensor([4.0], dtype=torch.float64, device="cuda")
        ptrs = (ctypes.c_void_p * world)(*[t.data_ptr() for t in tabs])
        ours.spgpuAllreduceSumDev(gpu_handle, value.data_ptr(), world, me, ptrs, 0)
        torch.cuda.synchronize()
        assert float(value.item()) == 6.5 and int(counters[1].item()) == 11
    finally:
        ours.spgpuSetSeqCounters(gpu_handle, 0, 0)
