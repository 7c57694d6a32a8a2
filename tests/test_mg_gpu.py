"""Single-GPU checks of the multi-GPU device code (include/spgpu_ext.h).  Kernels of
different ranks must never wait on each other on ONE GPU, so the neighbours are emulated:
their "pushes" are pre-filled halo zones + pre-set ready flags (or set LATE from a second
stream), and this rank's pushes land in scratch buffers standing in for the neighbours'
memory.  The real multi-rank runs are tests/test_mg_multi_gpu.py (needs >= 2 GPUs), bench.py
under torchrun (bit-exact against global columns, on by default) and the gloo tests."""
import ctypes
import struct

import numpy as np
import pytest

from spgpu_b200 import capi, formats as F, generators as G, mg
from tests import util

pytestmark = pytest.mark.gpu

TORCH_OF = {"S": "float32", "D": "float64", "C": "complex64", "Z": "complex128"}


def _middle_block(n=24, world=3, rank=1, sym="D"):
    coo = G.laplace3d_7pt(n)
    dt = util.TYPES[sym].np_dtype
    if sym != "D":
        rng = np.random.default_rng(5)
        vals = coo.vals.astype(dt)
        if util.TYPES[sym].is_complex:
            vals = vals + 1j * rng.uniform(-1, 1, vals.shape[0]).astype(util.real_of(dt))
        coo = F.Coo(coo.rows, coo.cols, vals.astype(dt), coo.nrows, coo.ncols, coo.base)
    plane = n * n
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    loc = mg.split_hell(hell, world, rank, plane)
    x = G.random_vector(coo.nrows, dt, 12345, -1, 1)
    x_ext = x[loc.lo - plane: loc.hi + plane].copy()
    want = util.oracle_spmv("hell", hell, x, None, 1.0, 0.0)[loc.lo:loc.hi]
    scale = util.row_scale(coo, x, None, 1.0, 0.0)[loc.lo:loc.hi]
    return coo, loc, plane, x, x_ext, want, scale


class Emulated:
    """One rank's view with emulated neighbours: flag blocks, the neighbours' zones (as plain local buffers) and the
    spgpuHaloLinks struct pointing at them.  Fused protocol words: [4] ready-from-below [5] ready-from-above
    [6] ack-from-below [7] ack-from-above."""

    def __init__(self, halo, tdtype, has_lo=True, has_hi=True):
        import torch
        nan = float("nan")
        self.my_flags = torch.zeros(16, dtype=torch.int32, device="cuda")
        self.pf = [torch.zeros(16, dtype=torch.int32, device="cuda") for _ in range(2)]
        self.pz = [torch.full((max(halo, 1),), nan, dtype=tdtype, device="cuda") for _ in range(2)]   # where my pushes land
        lk = capi.HaloLinks()
        if has_lo:
            lk.peerLoUpperZone, lk.peerFlagsLo = self.pz[0].data_ptr(), self.pf[0].data_ptr()
        if has_hi:
            lk.peerHiLowerZone, lk.peerFlagsHi = self.pz[1].data_ptr(), self.pf[1].data_ptr()
        lk.myFlags = self.my_flags.data_ptr()
        self.links, self.halo = lk, halo

    def ref(self):
        return ctypes.byref(self.links)

    def arrived(self, seq):
        """the neighbours' entries of exchange `seq` are in place and they have consumed mine of seq - 1"""
        self.my_flags[4] = seq; self.my_flags[5] = seq
        self.my_flags[6] = seq - 1; self.my_flags[7] = seq - 1


@pytest.mark.parametrize("seq", [1, 4])
@pytest.mark.parametrize("sym", ["S", "D", "C", "Z"])
def test_spmv_fused_with_halo_exchange(ours, gpu_handle, seq, sym):
    import torch
    coo, loc, plane, x, x_ext, want, scale = _middle_block(sym=sym)
    tdt = getattr(torch, TORCH_OF[sym])
    t = util.TYPES[sym]
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    em = Emulated(plane, tdt)
    em.arrived(seq)
    dz = torch.full((loc.nrows,), float("nan"), dtype=tdt, device="cuda")
    getattr(ours, f"spgpu{sym}hellspmvHalo")(gpu_handle, dz.data_ptr(), 0, t.scalar(1.0), dv.data_ptr(), di.data_ptr(), 32,
                                             dho.data_ptr(), drs.data_ptr(), 7, loc.nrows, dx.data_ptr(), t.scalar(0.0), 0,
                                             plane, em.ref(), seq)
    torch.cuda.synchronize()
    util.assert_rows_close(dz.cpu().numpy(), want, scale, sym, f"fused spmv+halo {sym}")
    # my first / last owned plane landed in the neighbours' zones
    np.testing.assert_array_equal(em.pz[0].cpu().numpy(), x_ext[plane:2 * plane])
    np.testing.assert_array_equal(em.pz[1].cpu().numpy(), x_ext[loc.nrows:loc.nrows + plane])
    lo_f, hi_f = em.pf[0].cpu().numpy(), em.pf[1].cpu().numpy()
    assert lo_f[5] == seq and lo_f[7] == seq - 1        # lower neighbour: ready-from-above; ack-from-above of the PREVIOUS
    assert hi_f[4] == seq and hi_f[6] == seq - 1        # exchange (this kernel started = that one is over); upper likewise
    assert lo_f[4] == 0 and hi_f[5] == 0 and not lo_f[:4].any() and not hi_f[:4].any()
    # the next exchange: the same call again
    em.arrived(seq + 1)
    getattr(ours, f"spgpu{sym}hellspmvHalo")(gpu_handle, dz.data_ptr(), 0, t.scalar(1.0), dv.data_ptr(), di.data_ptr(), 32,
                                             dho.data_ptr(), drs.data_ptr(), 7, loc.nrows, dx.data_ptr(), t.scalar(0.0), 0,
                                             plane, em.ref(), seq + 1)
    torch.cuda.synchronize()
    util.assert_rows_close(dz.cpu().numpy(), want, scale, sym, f"fused spmv+halo {sym}, second exchange")
    assert em.pf[0].cpu().numpy()[7] == seq and em.pf[1].cpu().numpy()[6] == seq
    assert ours.spgpuGetDeviceStatus(gpu_handle, 0) == 0


def _banded_block(nrows_total=3000, world=3, rank=1, halo=128, seed=3):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(nrows_total), 9)
    cols = rows + rng.integers(-halo, halo + 1, rows.shape[0])
    keep = (cols >= 0) & (cols < nrows_total)
    rows, cols = rows[keep], cols[keep]
    key = np.unique(rows.astype(np.int64) * nrows_total + cols)
    rows, cols = (key // nrows_total).astype(np.int32), (key % nrows_total).astype(np.int32)
    vals = rng.uniform(-1, 1, rows.shape[0])
    coo = F.Coo(rows, cols, vals, nrows_total, nrows_total, 0)
    hell = F.ell_to_hell(F.coo_to_ell(coo), 32)
    loc = mg.split_hell(hell, world, rank, halo)
    x = G.random_vector(nrows_total, np.float64, 99, -1, 1)
    want = util.oracle_spmv("hell", hell, x, None, 1.0, 0.0)[loc.lo:loc.hi]
    scale = util.row_scale(coo, x, None, 1.0, 0.0)[loc.lo:loc.hi]
    return loc, x[loc.lo - halo: loc.hi + halo].copy(), want, scale


@pytest.mark.parametrize("seq", [2, 3])
def test_fused_rows_not_a_multiple_of_128_and_late_flags(ours, gpu_handle, seq):
    """992 rows, halo 128: rows 864..991 read the upper zone, i.e. row blocks 6 AND 7 (block 6 holds rows
    768..895).  The neighbours' entries and ready flags arrive LATE, from a second stream: every row that reads a
    zone must have waited (a block that did not would multiply the NaN the zones hold before)."""
    import torch
    halo = 128
    loc, x_ext, want, scale = _banded_block(halo=halo)
    assert loc.nrows % 128 != 0 and (loc.nrows - halo) % 128 != 0
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    em = Emulated(halo, torch.float64)
    lo_vals = torch.from_numpy(x_ext[:halo].copy()).cuda()
    hi_vals = torch.from_numpy(x_ext[halo + loc.nrows:].copy()).cuda()
    dz = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
    late = torch.cuda.Stream()

    def neighbours_arrive():
        dx[:halo] = lo_vals
        dx[halo + loc.nrows:] = hi_vals
        em.my_flags[4:6] = seq
    with torch.cuda.stream(late):
        # run everything the late stream will do ONCE beforehand: the first launch of a kernel loads its module
        # (CUDA lazy loading), and a module load waits for running kernels -- i.e. for the spinning SpMV
        torch.cuda._sleep(1000)
        neighbours_arrive()
        em.my_flags.zero_()
        em.my_flags[6:8] = seq - 1                        # my previous pushes have been consumed
        dx[:halo] = float("nan"); dx[halo + loc.nrows:] = float("nan")
    torch.cuda.synchronize()
    ours.spgpuSetTuning(gpu_handle, b"spinTimeoutMs", 5000)
    ours.spgpuDhellspmvHalo(gpu_handle, dz.data_ptr(), 0, 1.0, dv.data_ptr(), di.data_ptr(), 32, dho.data_ptr(),
                            drs.data_ptr(), 9, loc.nrows, dx.data_ptr(), 0.0, 0, halo, em.ref(), seq)
    with torch.cuda.stream(late):
        torch.cuda._sleep(200_000_000)                    # ~0.1 s: the boundary blocks are spinning by now
        neighbours_arrive()
    torch.cuda.synchronize()
    ours.spgpuSetTuning(gpu_handle, b"spinTimeoutMs", 20000)
    assert ours.spgpuGetDeviceStatus(gpu_handle, 1) == 0
    util.assert_rows_close(dz.cpu().numpy(), want, scale, "D", "fused spmv+halo, late flags")


@pytest.mark.parametrize("seq", [1, 4])
@pytest.mark.parametrize("rank", [0, 1, 2])
def test_hdia_spmv_fused_with_halo_exchange(ours, gpu_handle, seq, rank):
    """spgpuDhdiaspmvHalo on the first / a middle / the last block of a 27-point stencil split in three
    (emulated neighbours): rows equal the global product, boundary entries land in the neighbours'
    zones, ready and ack words carry seq; also equal to the plain spgpuDhdiaspmv on the same block bit for bit"""
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
    em = Emulated(halo, torch.float64, has_lo, has_hi)
    em.arrived(seq)
    ours.spgpuDhdiaspmvHalo(gpu_handle, dz.data_ptr(), dy.data_ptr(), 1.5, dv.data_ptr(), doff.data_ptr(), 32, dho.data_ptr(),
                            loc.nrows, loc.ext_len, dx.data_ptr(), -0.5, halo, em.ref(), seq)
    torch.cuda.synchronize()
    scale = util.row_scale(coo, x, y, 1.5, -0.5)[loc.lo:loc.hi]
    util.assert_rows_close(dz.cpu().numpy(), want, scale, "D", "fused hdia spmv+halo")
    if has_lo:
        np.testing.assert_array_equal(em.pz[0].cpu().numpy(), x_ext[halo:2 * halo])
        f = em.pf[0].cpu().numpy()
        assert f[5] == seq and f[7] == seq - 1
    if has_hi:
        np.testing.assert_array_equal(em.pz[1].cpu().numpy(), x_ext[loc.nrows:loc.nrows + halo])
        f = em.pf[1].cpu().numpy()
        assert f[4] == seq and f[6] == seq - 1
    # the plain kernel on the same block
    dx2 = dx
    dz2 = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
    T = util.TYPES["D"]
    ours.spgpuDhdiaspmv(gpu_handle, dz2.data_ptr(), dy.data_ptr(), T.scalar(1.5), dv.data_ptr(), doff.data_ptr(), 32,
                        dho.data_ptr(), loc.nrows, loc.ext_len, dx2.data_ptr(), T.scalar(-0.5))
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
    for links in (None, Emulated(0, torch.float64, False, False).ref()):
        dz = torch.full((coo.nrows,), float("nan"), dtype=torch.float64, device="cuda")
        ours.spgpuDhellspmvHalo(gpu_handle, dz.data_ptr(), dy.data_ptr(), 1.5, dA["values"].data_ptr(), dA["indices"].data_ptr(),
                                32, dA["hack_offsets"].data_ptr(), dA["rs"].data_ptr(), 7, coo.nrows, dx.data_ptr(), -0.5, 0, 0,
                                links, 1)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(dz.cpu().numpy(), plain)


def test_halo_exchange_kernel_and_ack(ours, gpu_handle):
    """spgpuDhaloExchange / spgpuHaloAck / spgpuDhaloPush / spgpuHaloPush / spgpuWaitFlag with emulated neighbours"""
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
    ours.spgpuDhaloPush(gpu_handle, dst.data_ptr() + 8, src.data_ptr() + 8, n, flag.data_ptr(), 7)   # 8-byte aligned only
    ours.spgpuWaitFlag(gpu_handle, flag.data_ptr(), 7)
    torch.cuda.synchronize()
    assert torch.equal(dst[1:], src[1:n + 1]) and int(flag[0].item()) == 7
    # byte form: odd length, odd alignment
    bsrc = torch.arange(1003, dtype=torch.uint8, device="cuda")
    bdst = torch.zeros(1003, dtype=torch.uint8, device="cuda")
    ours.spgpuHaloPush(gpu_handle, bdst.data_ptr() + 1, bsrc.data_ptr() + 1, 1001, flag.data_ptr() + 4, 9)
    torch.cuda.synchronize()
    assert torch.equal(bdst[1:1002], bsrc[1:1002]) and int(bdst[0].item()) == 0 and int(bdst[1002].item()) == 0
    assert int(flag[1].item()) == 9


def test_wait_timeout_sets_the_sticky_status(ours, gpu_handle):
    """a wait for a peer that never comes gives up after spinTimeoutMs, records SPGPU_DEVSTATUS_TIMEOUT in the
    handle (host-readable), and later waits of the handle return at once until the status is cleared"""
    import time
    import torch
    flag = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert ours.spgpuGetDeviceStatus(gpu_handle, 1) >= 0
    ours.spgpuSetTuning(gpu_handle, b"spinTimeoutMs", 30)
    try:
        ours.spgpuWaitFlag(gpu_handle, flag.data_ptr(), 1)
        torch.cuda.synchronize()
        assert ours.spgpuGetDeviceStatus(gpu_handle, 0) == 1
        t0 = time.perf_counter()
        for _ in range(20):
            ours.spgpuWaitFlag(gpu_handle, flag.data_ptr(), 1)
        torch.cuda.synchronize()
        assert time.perf_counter() - t0 < 0.3              # 20 full timeouts would be >= 0.6 s
        assert ours.spgpuGetDeviceStatus(gpu_handle, 1) == 1
        assert ours.spgpuGetDeviceStatus(gpu_handle, 0) == 0
    finally:
        ours.spgpuSetTuning(gpu_handle, b"spinTimeoutMs", 20000)


def _slot_words(a, b, seq):
    alo, ahi = struct.unpack("<ii", struct.pack("<d", a))
    blo, bhi = struct.unpack("<ii", struct.pack("<d", b))
    return [alo, ahi, blo, bhi, seq, 0, 0, 0]


def _tables(world, me, seq, contributions):
    """device tables of all ranks; this rank's own holds the peers' contributions for `seq`"""
    import torch
    W = capi.AR_SLOT_BYTES // 4
    tabs = [torch.zeros(2 * world * W, dtype=torch.int32, device="cuda") for _ in range(world)]
    host = np.zeros(2 * world * W, dtype=np.int32)
    for r, (a, b) in contributions.items():
        base = ((seq & 1) * world + r) * W
        host[base:base + W] = _slot_words(a, b, seq)
    tabs[me].copy_(torch.from_numpy(host))
    ptrs = (ctypes.c_void_p * world)(*[t.data_ptr() for t in tabs])
    args = capi.PeerAllreduceArgs(world, me, ctypes.cast(ptrs, ctypes.POINTER(ctypes.c_void_p)), seq)
    return tabs, ptrs, args


@pytest.mark.parametrize("sym", ["S", "D", "C", "Z"])
def test_peer_allreduce_kernel_with_emulated_peer(ours, gpu_handle, sym):
    """spgpu?allreduceSumDev, world = 3, this rank = 1: the two peers' contributions are pre-filled in
    this rank's table; this rank's (value, seq) must land in slot 1 of every table and the sum must
    be taken in rank order, in double"""
    import torch
    world, me = 3, 1
    tdt = getattr(torch, TORCH_OF[sym])
    W = capi.AR_SLOT_BYTES // 4
    cplx = util.TYPES[sym].is_complex
    for seq in (1, 2, 7):
        tabs, ptrs, args = _tables(world, me, seq, {0: (0.125, 1.0), 2: (-3.5, 0.25)})
        mine = (10.0 - 2.0j) if cplx else 10.0
        value = torch.tensor([mine], dtype=tdt, device="cuda")
        getattr(ours, f"spgpu{sym}allreduceSumDev")(gpu_handle, value.data_ptr(), ctypes.byref(args))
        torch.cuda.synchronize()
        got = value.cpu().numpy()[0]
        assert got.real == (0.125 + 10.0) + -3.5
        if cplx:
            assert got.imag == (1.0 + -2.0) + 0.25
        for t in tabs:
            w = t.cpu().numpy()[((seq & 1) * world + me) * W:((seq & 1) * world + me + 1) * W]
            assert w.tolist()[:5] == _slot_words(10.0, -2.0 if cplx else 0.0, seq)[:5]


@pytest.mark.parametrize("sym", ["D", "Z", "S"])
def test_spmv_halo_dot_with_fused_allreduce(ours, gpu_handle, sym):
    """spgpu?hellspmvHaloDot: fused exchange + SpMV + this rank's share of p.Ap, all-reduced by the fold kernel's
    last CTA (emulated peers, world 3)"""
    import torch
    coo, loc, plane, x, x_ext, want, scale = _middle_block(sym=sym)
    tdt = getattr(torch, TORCH_OF[sym])
    cplx = util.TYPES[sym].is_complex
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    own = x_ext[plane:plane + loc.nrows]
    local = np.sum(own.astype(np.complex128) * want.astype(np.complex128))          # unconjugated, like spgpu?dot
    mag = float(np.sum(np.abs(own) * np.abs(want)))
    tol = 1e-12 if sym in "DZ" else 2e-6
    for seq, ar_on in ((1, False), (2, True)):
        em = Emulated(plane, tdt)
        em.arrived(seq)
        dz = torch.full((loc.nrows,), float("nan"), dtype=tdt, device="cuda")
        dres = torch.full((1,), float("nan"), dtype=tdt, device="cuda")
        tabs, ptrs, args = _tables(3, 1, 5, {0: (100.0, 7.0), 2: (-50.0, 1.0)})
        getattr(ours, f"spgpu{sym}hellspmvHaloDot")(gpu_handle, dz.data_ptr(), dv.data_ptr(), di.data_ptr(), 32, dho.data_ptr(),
                                                    drs.data_ptr(), 7, loc.nrows, dx.data_ptr(), 0, plane, em.ref(), seq,
                                                    dres.data_ptr(), ctypes.byref(args) if ar_on else None)
        torch.cuda.synchronize()
        util.assert_rows_close(dz.cpu().numpy(), want, scale, sym, "fused spmv+halo+dot")
        got = complex(dres.cpu().numpy()[0])
        ref = local + ((50.0 + (8.0j if cplx else 0.0)) if ar_on else 0.0)
        if not cplx:
            ref = ref.real
        assert abs(got - ref) <= tol * (mag + 160.0), (got, ref)
    assert ours.spgpuGetDeviceStatus(gpu_handle, 0) == 0


def test_device_resident_sequence_numbers(ours, gpu_handle):
    """seq == 0: the fused halo kernel and the all-reduce take their sequence number from device counters
    (spgpuSetSeqCounters); spgpuHaloSeqAdvance / the all-reduce itself advance them -- what lets a partitioned
    iteration be replayed from a CUDA graph.  Emulated neighbours as above."""
    import torch
    coo, loc, plane, x, x_ext, want, scale = _middle_block()
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    counters = torch.tensor([6, 10], dtype=torch.int32, device="cuda")       # 6 exchanges, 10 all-reduces done so far
    assert ours.spgpuSetSeqCounters(gpu_handle, counters.data_ptr(), counters.data_ptr() + 4) == 0
    try:
        em = Emulated(plane, torch.float64)
        for k in (7, 8):                                                     # two exchanges: 7 and 8
            em.arrived(k)
            dz = torch.full((loc.nrows,), float("nan"), dtype=torch.float64, device="cuda")
            ours.spgpuDhellspmvHalo(gpu_handle, dz.data_ptr(), 0, 1.0, dv.data_ptr(), di.data_ptr(), 32, dho.data_ptr(),
                                    drs.data_ptr(), 7, loc.nrows, dx.data_ptr(), 0.0, 0, plane, em.ref(), 0)
            ours.spgpuHaloSeqAdvance(gpu_handle)
            torch.cuda.synchronize()
            util.assert_rows_close(dz.cpu().numpy(), want, scale, "D", "fused spmv+halo, device seq")
            assert int(counters[0].item()) == k
            assert em.pf[0].cpu().numpy()[5] == k and em.pf[1].cpu().numpy()[4] == k
            assert em.pf[0].cpu().numpy()[7] == k - 1 and em.pf[1].cpu().numpy()[6] == k - 1
        # all-reduce number 11, world 2, this rank 0: the peer's slot is pre-filled
        tabs, ptrs, args = _tables(2, 0, 11, {1: (2.5, 0.0)})
        args.seq = 0
        value = torch.tensor([4.0], dtype=torch.float64, device="cuda")
        ours.spgpuDallreduceSumDev(gpu_handle, value.data_ptr(), ctypes.byref(args))
        torch.cuda.synchronize()
        assert float(value.item()) == 6.5 and int(counters[1].item()) == 11
    finally:
        ours.spgpuSetSeqCounters(gpu_handle, 0, 0)


def test_halo_trace(ours, gpu_handle):
    """haloTrace tuning key + spgpuHaloTraceRead: timestamps of one fused exchange are recorded and ordered"""
    import torch
    coo, loc, plane, x, x_ext, want, scale = _middle_block()
    dv, di, dho, drs = (util.to_dev(a) for a in (loc.values, loc.indices, loc.hack_offsets, loc.rs))
    dx = util.to_dev(x_ext)
    em = Emulated(plane, torch.float64)
    em.arrived(2)
    dz = torch.zeros(loc.nrows, dtype=torch.float64, device="cuda")
    assert ours.spgpuSetTuning(gpu_handle, b"haloTrace", 1) == 0
    try:
        ours.spgpuDhellspmvHalo(gpu_handle, dz.data_ptr(), 0, 1.0, dv.data_ptr(), di.data_ptr(), 32, dho.data_ptr(),
                                drs.data_ptr(), 7, loc.nrows, dx.data_ptr(), 0.0, 0, plane, em.ref(), 2)
        buf = (ctypes.c_ulonglong * 8)()
        assert ours.spgpuHaloTraceRead(gpu_handle, buf, 2, 1) == 0
        t = list(buf)
        assert t[0] > 0 and t[1] >= t[0] and t[6] > 0 and t[7] >= t[6]
    finally:
        assert ours.spgpuSetTuning(gpu_handle, b"haloTrace", 0) == 0
