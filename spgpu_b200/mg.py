"""Row-partitioned multi-GPU SpMV (north_star item 3; no reference counterpart:
the reference is one handle per device with no communication, core.h:88-93).

One process per GPU (torchrun), `torch.distributed` for the plumbing.  The
matrix is split in contiguous row blocks whose boundaries are multiples of the
hack size, so a block is a self-contained HELL matrix (its hackOffsets re-base
by subtraction).  Each rank owns x[lo:hi) and keeps it inside

    x_ext = [ lower halo (w) | owned (hi-lo) | upper halo (w) ]

with the block's column indices remapped to x_ext positions, so the UNCHANGED
single-GPU kernel (spgpu?hellspmv through the C ABI) runs on it.  Before a SpMV
the two halo zones are filled from the neighbours' boundary entries:

  * mode "nccl":  grouped ncclSend/ncclRecv (torch.distributed P2P ops);
  * mode "fused" (default on GPUs): ONE kernel per SpMV, spgpu?hellspmvHalo: its
    first CTAs store this rank's boundary entries straight into the neighbours' halo
    zones through CUDA-IPC peer pointers over NVLink and publish a sequence number;
    the interior row blocks are scheduled first and the row blocks that read a halo
    zone late (they wait on the local ready flag).  A rank acknowledges an exchange
    when its NEXT fused kernel starts (stream order says the rows that read the zones
    have finished), so the row blocks carry no tickets, fences or barriers for it.
    Neighbours may drift most of a kernel apart.  Transfer and multiply overlap inside
    one launch.
  * mode "push":  each rank's spgpuDhaloPush kernel stores its boundary entries
    straight into the neighbour's halo zone through a CUDA-IPC peer pointer over
    NVLink and release-stores a sequence number into the neighbour's flag word;
    the consumer's stream waits on its own flag (spgpuWaitFlag).  An "ack" flag
    in the other direction keeps a producer from overwriting a halo that is
    still being read.
  * mode "allgather": for unstructured columns every rank gathers the full x.

Rows that reference no halo entry ("interior") are multiplied while the
exchange is in flight; the boundary rows run after it.  Dots are local partial
reductions + an all-reduce of one scalar.

The same class runs on CPU tensors with the gloo backend and a caller-supplied
local SpMV (the tests pass the CPU oracle) to check the partition / remap /
exchange logic without a GPU.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


# --------------------------------------------------------------------------- #
# partitioning (pure host logic)
# --------------------------------------------------------------------------- #

def row_blocks(nrows: int, world: int, align: int):
    """Contiguous, near-equal row blocks whose boundaries are multiples of `align`."""
    units = (nrows + align - 1) // align
    bounds = [min(nrows, ((units * r) // world) * align) for r in range(world)] + [nrows]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def _check_block(lo: int, hi: int, halo: int, world: int):
    """The exchange sends a rank's first / last `halo` OWNED entries to its neighbours: a block
    must own at least that many rows (otherwise a halo would need entries of a rank two hops away)."""
    if world > 1 and hi - lo < halo:
        raise ValueError(f"row block [{lo}, {hi}) is smaller than the halo width {halo}: use fewer ranks or mode 'allgather'")


@dataclass
class LocalHell:
    """One rank's block in HELL layout with columns remapped into x_ext."""
    values: object
    indices: object
    hack_offsets: object
    rs: object
    hack_size: int
    nrows: int
    lo: int
    hi: int
    halo: int          # entries in each halo zone
    base: int
    nnz: int

    @property
    def ext_len(self):
        return self.nrows + 2 * self.halo


def _live_slots(rs: np.ndarray, hack_offsets: np.ndarray, hs: int) -> np.ndarray:
    """element positions of every stored entry of a HELL block: entry k of row r sits at
    hackOffsets[r / hackSize] + k * hackSize + r % hackSize, for k < rS[r]"""
    rs64 = rs.astype(np.int64)
    rows = np.repeat(np.arange(rs.shape[0], dtype=np.int64), rs64)
    first = np.cumsum(rs64) - rs64
    k = np.arange(rows.shape[0], dtype=np.int64) - first[rows]
    return np.asarray(hack_offsets, dtype=np.int64)[rows // hs] + k * hs + rows % hs


def split_hell(hell, world: int, rank: int, halo: int) -> LocalHell:
    """Cut rank's block out of a global host-side formats.Hell and remap its
    column indices to x_ext positions.  Raises if a row references a column
    outside [lo-halo, hi+halo) (then use mode 'allgather')."""
    hs = hell.hack_size
    lo, hi = row_blocks(hell.nrows, world, hs)[rank]
    h0, h1 = lo // hs, (hi + hs - 1) // hs
    hoff = np.asarray(hell.hack_offsets)
    e0 = int(hoff[h0]) if h0 < hoff.shape[0] else int(hell.values.shape[0])
    e1 = int(hoff[h1]) if h1 < hoff.shape[0] else int(hell.values.shape[0])
    values = np.array(hell.values[e0:e1], copy=True)
    indices = np.array(hell.indices[e0:e1], copy=True)
    rs = np.array(hell.rs[lo:hi], copy=True)
    local_hoff = (hoff[h0:h1] - e0).astype(np.int32)
    # remap only the slots that exist (padding is undefined and stays so)
    sl = _live_slots(rs, local_hoff, hs)
    g = indices[sl].astype(np.int64) - hell.base
    if ((g < lo - halo) | (g >= hi + halo)).any():
        raise ValueError("column outside the halo window; use mode='allgather'")
    # The overlapped and the fused exchange multiply rows [halo, nrows - halo) BEFORE the halos have arrived
    # (they are "interior"): only the first `halo` rows may read the lower zone and only the last `halo` rows
    # the upper one.  A banded matrix (|col - row| <= halo) always satisfies this; anything else is refused.
    rs64 = rs.astype(np.int64)
    row_of = np.repeat(np.arange(hi - lo, dtype=np.int64), rs64)
    if world > 1 and (((g < lo) & (row_of >= halo)) | ((g >= hi) & (row_of < (hi - lo) - halo))).any():
        raise ValueError("a row outside the first / last `halo` rows of the block references a halo column "
                         "(the pattern is not banded within the halo width); use mode='allgather'")
    indices[sl] = (g - (lo - halo) + hell.base).astype(np.int32)
    return LocalHell(values, indices, local_hoff, rs, hs, hi - lo, lo, hi, halo, hell.base, int(sl.size))


def split_hell_allgather(hell, world: int, rank: int) -> LocalHell:
    """Cut rank's block out of a global HELL matrix for the ALL-GATHER mode (unstructured
    columns): every rank gathers the whole x into x_full = [block 0 | block 1 | ...], each
    block padded to the largest block, and the column indices are remapped to that padded
    layout.  Works for any sparsity pattern; costs an all-gather of x per SpMV."""
    hs = hell.hack_size
    blocks = row_blocks(hell.nrows, world, hs)
    lo, hi = blocks[rank]
    h0, h1 = lo // hs, (hi + hs - 1) // hs
    hoff = np.asarray(hell.hack_offsets)
    end = int(hell.values.shape[0])
    e0 = int(hoff[h0]) if h0 < hoff.shape[0] else end
    e1 = int(hoff[h1]) if h1 < hoff.shape[0] else end
    values = np.array(hell.values[e0:e1], copy=True)
    indices = np.array(hell.indices[e0:e1], copy=True)
    rs = np.array(hell.rs[lo:hi], copy=True)
    local_hoff = (hoff[h0:h1] - e0).astype(np.int32)
    starts = np.array([b[0] for b in blocks], dtype=np.int64)
    widest = max(b[1] - b[0] for b in blocks)
    sl = _live_slots(rs, local_hoff, hs)
    g = indices[sl].astype(np.int64) - hell.base
    owner = np.searchsorted(starts, g, side="right") - 1
    indices[sl] = (owner * widest + (g - starts[owner]) + hell.base).astype(np.int32)
    out = LocalHell(values, indices, local_hoff, rs, hs, hi - lo, lo, hi, 0, hell.base, int(sl.size))
    out.widest = widest
    return out


@dataclass
class LocalHdia:
    """One rank's block in HDIA layout, diagonal offsets shifted so that row r of the
    block reads x_ext[r + offset] (x_ext = [halo | owned | halo])."""
    values: object
    offsets: object
    hack_offsets: object
    hack_size: int
    nrows: int
    lo: int
    hi: int
    halo: int

    @property
    def ext_len(self):
        return self.nrows + 2 * self.halo


def split_hdia(hdia, world: int, rank: int, halo: int) -> LocalHdia:
    """Cut rank's block out of a global host-side formats.Hdia.

    HDIA addresses x relative to the row (column = row + offset), so a row block is the
    same hacks with hackOffsets re-based by subtraction; adding `halo` to every offset
    makes local row r read x_ext[r + offset + halo], and spgpu?hdiaspmv's own range test
    0 <= r + offset' < cols with cols = ext_len then guards the window instead of the
    matrix.  Cells whose GLOBAL column is outside [0, ncols) are zeroed in the copy (the
    kernel would now see them inside the window; the conversions store 0 there anyway).
    Raises if a non-zero cell needs a column outside [lo-halo, hi+halo)."""
    hs = hdia.hack_size
    lo, hi = row_blocks(hdia.nrows, world, hs)[rank]
    h0, h1 = lo // hs, (hi + hs - 1) // hs
    hoff = np.asarray(hdia.hack_offsets)
    d0, d1 = int(hoff[h0]), int(hoff[h1])
    values = np.array(hdia.values[d0 * hs:d1 * hs], copy=True).reshape(d1 - d0, hs)
    offs = np.asarray(hdia.offsets[d0:d1]).astype(np.int64)
    local_hoff = (hoff[h0:h1 + 1] - d0).astype(np.int32)
    # local row of cell (d, j): first row of d's hack + j
    hack_of_diag = np.repeat(np.arange(h1 - h0, dtype=np.int64), np.diff(local_hoff))
    rows = hack_of_diag[:, None] * hs + np.arange(hs, dtype=np.int64)[None, :]
    cols = lo + rows + offs[:, None]                              # global column of every cell
    outside_matrix = (cols < 0) | (cols >= hdia.ncols) | (rows >= hi - lo)
    values[outside_matrix] = 0
    outside_window = ~outside_matrix & ((cols < lo - halo) | (cols >= hi + halo))
    if (values[outside_window] != 0).any():
        raise ValueError("diagonal reaches outside the halo window")
    values[outside_window] = 0
    # same band requirement as split_hell: interior rows are multiplied before the halos arrive
    early = ~outside_matrix & (((cols < lo) & (rows >= halo)) | ((cols >= hi) & (rows < (hi - lo) - halo)))
    if world > 1 and (values[early] != 0).any():
        raise ValueError("a row outside the first / last `halo` rows of the block reads a halo column "
                         "(the pattern is not banded within the halo width)")
    return LocalHdia(values.reshape(-1), (offs + halo).astype(np.int32), local_hoff, hs, hi - lo, lo, hi, halo)


class MgAllGatherSpmv:
    """z_owned = A_block * x with x all-gathered from every rank's (padded) owned slice."""

    def __init__(self, world, widest, local_spmv, group=None):
        self.world, self.widest, self.local_spmv, self.group = world, widest, local_spmv, group

    def apply(self, z, x_owned_padded, x_full):
        if self.world > 1:
            dist.all_gather_into_tensor(x_full, x_owned_padded, group=self.group)
        else:
            x_full.copy_(x_owned_padded)
        self.local_spmv(z, x_full)


# --------------------------------------------------------------------------- #
# communicator
# --------------------------------------------------------------------------- #

class HaloExchange:
    """Fills the two halo zones of x_ext from the neighbours (1-D chain of ranks)."""

    def __init__(self, rank, world, halo, mode="nccl", group=None):
        self.rank, self.world, self.halo, self.mode, self.group = rank, world, halo, mode, group
        self.seq = 0
        self.push = None

    # -- NCCL / gloo point-to-point ------------------------------------------
    def start(self, x_ext: torch.Tensor):
        """Begin the exchange; returns a list of work handles (may be empty)."""
        w, n = self.halo, x_ext.numel() - 2 * self.halo
        if self.world == 1 or w == 0:
            return []
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, x_ext[w:2 * w], self.rank - 1, self.group))
            ops.append(dist.P2POp(dist.irecv, x_ext[0:w], self.rank - 1, self.group))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, x_ext[n:n + w], self.rank + 1, self.group))
            ops.append(dist.P2POp(dist.irecv, x_ext[n + w:n + 2 * w], self.rank + 1, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(works):
        for wk in works:
            wk.wait()


class PeerHalo:
    """NVLink peer stores + flags between the rank processes (CUDA IPC): the separate-kernel exchange
    (exchange / wait / ack, exchange_fused / ack_fused) and the links of the kernels that carry the
    exchange inside the SpMV (spgpu?{hell,hdia}spmvHalo[Dot]: their own flag words and sequence numbers)."""

    def __init__(self, L, handle, rank, world, x_ext_ptr, ext_len, halo, group=None, itemsize=8):
        from .capi import HALO_FLAG_WORDS, HaloLinks
        self.L, self.h, self.rank, self.world, self.halo = L, handle, rank, world, halo
        self.x_ptr, self.ext_len, self.isz = x_ext_ptr, ext_len, itemsize
        self.seq = 0            # exchanges of the separate-kernel protocol
        self.fseq = 0           # exchanges of the fused protocol
        flags = ctypes.c_void_p()
        assert L.spgpuDeviceAlloc(ctypes.byref(flags), 4 * HALO_FLAG_WORDS) == 0
        self.flags = flags.value
        self._flag_view = _as_tensor(self.flags, HALO_FLAG_WORDS, torch.int32)
        self._flag_view.zero_()
        torch.cuda.synchronize()
        hx, hf = ((ctypes.c_char * 64)() for _ in range(2))
        assert L.spgpuIpcGetHandle(x_ext_ptr, hx) == 0
        assert L.spgpuIpcGetHandle(self.flags, hf) == 0
        mine = (bytes(hx), bytes(hf), ext_len)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        self.peer = {}
        for nb in (rank - 1, rank + 1):
            if 0 <= nb < world:
                opened = []
                for blob in (everyone[nb][0], everyone[nb][1]):
                    q = ctypes.c_void_p()
                    rc = L.spgpuIpcOpenHandle((ctypes.c_char * 64).from_buffer_copy(blob), ctypes.byref(q))
                    assert rc == 0, f"cudaIpcOpenMemHandle -> {rc}"
                    opened.append(q.value)
                self.peer[nb] = (opened[0], opened[1], everyone[nb][2])
        lo, hi = self.peer.get(rank - 1), self.peer.get(rank + 1)
        w, b = halo, itemsize
        lk = HaloLinks()
        if lo:      # the lower neighbour's UPPER zone: the end of its x_ext
            lk.peerLoUpperZone = lo[0] + b * (lo[2] - w)
            lk.peerFlagsLo = lo[1]
        if hi:      # the upper neighbour's LOWER zone: the start of its x_ext
            lk.peerHiLowerZone = hi[0]
            lk.peerFlagsHi = hi[1]
        lk.myFlags = self.flags
        self.links = lk
        dist.barrier(group=group)

    def exchange(self):
        """push my boundary entries to both neighbours, then make my stream wait
        until both of mine have arrived (all stream-ordered, no host sync)."""
        L, h, w, b = self.L, self.h, self.halo, self.isz
        n = self.ext_len - 2 * w
        self.seq += 1
        seq = self.seq
        lo_nb, hi_nb = self.rank - 1, self.rank + 1
        # wait until the neighbours have consumed the previous halo I sent them
        if seq > 1:
            if lo_nb in self.peer:
                L.spgpuWaitFlag(h, self.flags + 4 * 2, seq - 1)
            if hi_nb in self.peer:
                L.spgpuWaitFlag(h, self.flags + 4 * 3, seq - 1)
        if lo_nb in self.peer:       # my first w owned entries -> their upper halo, their flag[1]
            px, pf, plen = self.peer[lo_nb]
            L.spgpuHaloPush(h, px + b * (plen - w), self.x_ptr + b * w, w * b, pf + 4 * 1, seq)
        if hi_nb in self.peer:       # my last w owned entries -> their lower halo, their flag[0]
            px, pf, plen = self.peer[hi_nb]
            L.spgpuHaloPush(h, px, self.x_ptr + b * n, w * b, pf + 4 * 0, seq)

    def wait(self):
        L, h, seq = self.L, self.h, self.seq
        if self.rank - 1 in self.peer:
            L.spgpuWaitFlag(h, self.flags + 4 * 0, seq)
        if self.rank + 1 in self.peer:
            L.spgpuWaitFlag(h, self.flags + 4 * 1, seq)

    def ack(self):
        """tell both neighbours their halo data has been consumed (after the SpMV)."""
        L, h, seq = self.L, self.h, self.seq
        for nb, word in ((self.rank - 1, 3), (self.rank + 1, 2)):
            if nb in self.peer:
                pf = self.peer[nb][1]
                L.spgpuHaloPush(h, 0, 0, 0, pf + 4 * word, seq)

    def exchange_fused(self):
        """ONE kernel: wait acks, push both boundary runs, signal, wait for my own halos."""
        L, h, w, b = self.L, self.h, self.halo, self.isz
        n = self.ext_len - 2 * w
        self.seq += 1
        lo, hi = self.peer.get(self.rank - 1), self.peer.get(self.rank + 1)
        f = self.flags
        L.spgpuHaloExchange(
            h,
            (lo[0] + b * (lo[2] - w)) if lo else 0, self.x_ptr + b * w,      # -> lower neighbour's upper halo
            hi[0] if hi else 0, self.x_ptr + b * n,                           # -> upper neighbour's lower halo
            w * b,
            (f + 4 * 2) if lo else 0, (f + 4 * 3) if hi else 0,               # acks I wait for
            (lo[1] + 4 * 1) if lo else 0, (hi[1] + 4 * 0) if hi else 0,       # their ready flags
            (f + 4 * 0) if lo else 0, (f + 4 * 1) if hi else 0,               # my ready flags
            self.seq)

    def ack_fused(self):
        lo, hi = self.peer.get(self.rank - 1), self.peer.get(self.rank + 1)
        self.L.spgpuHaloAck(self.h, (lo[1] + 4 * 3) if lo else 0, (hi[1] + 4 * 2) if hi else 0, self.seq)

    def links_ref(self):
        """the spgpuHaloLinks* argument of spgpu?{hell,hdia}spmvHalo[Dot]"""
        return ctypes.byref(self.links)

    def next_seq(self):
        """sequence number of the next exchange of the FUSED protocol"""
        self.fseq += 1
        return self.fseq

    # -- sequence numbers in device memory (CUDA-graph replay, include/spgpu_ext.h) --------------
    def to_device_seq(self, counter: torch.Tensor):
        """hand the fused-exchange count to a 1-element int32 device counter (the handle must have been given
        it with spgpuSetSeqCounters): from now on the fused kernels are called with seq = 0"""
        counter.fill_(self.fseq)

    def from_device_seq(self, counter: torch.Tensor):
        self.fseq = int(counter.item())

    def close(self):
        torch.cuda.synchronize()
        for px, pf, _ in self.peer.values():
            self.L.spgpuIpcCloseHandle(px)
            self.L.spgpuIpcCloseHandle(pf)
        self.peer = {}


class PeerAllreduce:
    """Sum all-reduce of one value over NVLink peer memory (spgpu?allreduceSumDev, or inside the last CTA of
    the kernel that produces the value -- `next_ref()` is the spgpuPeerAllreduce* argument of those entry
    points): every rank's 2*world-slot table is CUDA-IPC mapped into every other rank."""

    def __init__(self, L, handle, rank, world, group=None):
        from .capi import AR_SLOT_BYTES, PeerAllreduceArgs
        assert world <= 16
        self.L, self.h, self.rank, self.world = L, handle, rank, world
        self.seq = 0
        nbytes = 2 * world * AR_SLOT_BYTES
        p = ctypes.c_void_p()
        assert L.spgpuDeviceAlloc(ctypes.byref(p), nbytes) == 0
        self.table = p.value
        _as_tensor(self.table, nbytes // 4, torch.int32).zero_()
        torch.cuda.synchronize()
        hb = (ctypes.c_char * 64)()
        assert L.spgpuIpcGetHandle(self.table, hb) == 0
        everyone = [None] * world
        dist.all_gather_object(everyone, bytes(hb), group=group)
        self.tables = (ctypes.c_void_p * world)()
        self.opened = []
        for r in range(world):
            if r == rank:
                self.tables[r] = self.table
            else:
                q = ctypes.c_void_p()
                buf = (ctypes.c_char * 64).from_buffer_copy(everyone[r])
                rc = L.spgpuIpcOpenHandle(buf, ctypes.byref(q))
                assert rc == 0, f"cudaIpcOpenMemHandle -> {rc}"
                self.tables[r] = q.value
                self.opened.append(q.value)
        self.args = PeerAllreduceArgs(world, rank, ctypes.cast(self.tables, ctypes.POINTER(ctypes.c_void_p)), 0)
        dist.barrier(group=group)

    def next_ref(self):
        """spgpuPeerAllreduce* for the NEXT all-reduce (the struct is read during the call that takes it)"""
        if self.device_seq:
            self.args.seq = 0
        else:
            self.seq += 1
            self.args.seq = self.seq
        return ctypes.byref(self.args)

    def __call__(self, t: torch.Tensor):
        """in-place sum of a 1-element device tensor across the ranks (stream-ordered, its own kernel)"""
        sym = {torch.float32: "S", torch.float64: "D", torch.complex64: "C", torch.complex128: "Z"}[t.dtype]
        getattr(self.L, f"spgpu{sym}allreduceSumDev")(self.h, t.data_ptr(), self.next_ref())

    device_seq = False

    def to_device_seq(self, counter: torch.Tensor):
        """continue the sequence from a device counter (spgpuSetSeqCounters): calls pass seq = 0 and the
        kernel advances the counter itself, so the call can be replayed from a CUDA graph"""
        counter.fill_(self.seq)
        self.device_seq = True

    def from_device_seq(self, counter: torch.Tensor):
        self.seq = int(counter.item())
        self.device_seq = False

    def close(self):
        torch.cuda.synchronize()
        for q in self.opened:
            self.L.spgpuIpcCloseHandle(q)
        self.opened = []


class _RawCuda:
    """__cuda_array_interface__ carrier so torch can view a raw cudaMalloc block."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                         "version": 2, "strides": None}


def _as_tensor(ptr: int, n: int, dtype: torch.dtype) -> torch.Tensor:
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4", torch.complex64: "<c8",
               torch.complex128: "<c16"}[dtype]
    return torch.as_tensor(_RawCuda(ptr, n, typestr), device="cuda")


def raw_device_vector(L, n: int, dtype=torch.float64):
    """(ptr, torch view) of a dedicated cudaMalloc allocation (IPC-exportable)."""
    p = ctypes.c_void_p()
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    rc = L.spgpuDeviceAlloc(ctypes.byref(p), nbytes)
    assert rc == 0, f"cudaMalloc({nbytes}) -> {rc}"
    return p.value, _as_tensor(p.value, n, dtype)


# --------------------------------------------------------------------------- #
# the partitioned operator
# --------------------------------------------------------------------------- #

class MgHellSpmv:
    """z_owned = A_block * x_ext with halo exchange, one rank's view (HELL or HDIA blocks:
    the format only lives inside local_spmv).

    local_spmv(z, x_ext, row0, row1) multiplies rows [row0, row1) of the block
    (used to split interior / boundary work); on the GPU it is a closure around
    spgpuDhellspmv with offset pointers (same trick as the reference's
    large-vector loop, hell_spmv_base.cuh:121-137)."""

    def __init__(self, rank, world, nrows, halo, local_spmv, exchange: HaloExchange,
                 peer: PeerHalo | None = None, overlap=True, align=32, fused_spmv=None):
        self.rank, self.world, self.nrows, self.halo = rank, world, nrows, halo
        _check_block(0, nrows, halo, world)
        self.local_spmv, self.ex, self.peer = local_spmv, exchange, peer
        # rows [0, head) and [tail, nrows) may touch a halo zone; both cuts sit on
        # hack boundaries so each piece is a valid HELL sub-matrix
        self.head = min(nrows, -(-halo // align) * align)
        self.tail = max(self.head, ((nrows - halo) // align) * align)
        self.overlap = overlap and world > 1 and halo > 0 and self.tail > self.head
        # fused_spmv(seq): the SpMV kernel that carries its own halo exchange (spgpu?hellspmvHalo); a row block
        # of it that reads both zones (fewer rows than two halo widths) simply waits for both neighbours
        self.fused_spmv = fused_spmv

    def apply(self, z, x_ext):
        w, n = self.halo, self.nrows
        if self.world == 1 or w == 0:
            self.local_spmv(z, x_ext, 0, n)
            return
        if self.peer is not None and self.fused_spmv is not None:
            self.fused_spmv(self.peer.next_seq())      # ONE kernel: push + multiply + ack
            return
        if self.peer is not None:
            # NVLink push: everything is ordered on the handle's stream
            if self.overlap:
                self.peer.exchange()
                self.local_spmv(z, x_ext, self.head, self.tail)   # interior rows need no halo
                self.peer.wait()
                self.local_spmv(z, x_ext, 0, self.head)
                self.local_spmv(z, x_ext, self.tail, n)
                self.peer.ack()
            else:
                self.peer.exchange_fused()          # returns when both halos are in place
                self.local_spmv(z, x_ext, 0, n)
                self.peer.ack_fused()
            return
        works = self.ex.start(x_ext)
        if self.overlap:
            self.local_spmv(z, x_ext, self.head, self.tail)
            HaloExchange.finish(works)
            self.local_spmv(z, x_ext, 0, self.head)
            self.local_spmv(z, x_ext, self.tail, n)
        else:
            HaloExchange.finish(works)
            self.local_spmv(z, x_ext, 0, n)


def global_dot(local_value: torch.Tensor, group=None) -> torch.Tensor:
    """sum over ranks of a 1-element tensor holding the local partial dot."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_value, op=dist.ReduceOp.SUM, group=group)
    return local_value


def global_amax(local_value: torch.Tensor, group=None) -> torch.Tensor:
    """max over ranks of a 1-element tensor holding the local amax (amax reduces with max)."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_value, op=dist.ReduceOp.MAX, group=group)
    return local_value


def global_nrm2(local_sum_of_squares: torch.Tensor, group=None) -> torch.Tensor:
    """sqrt of the sum over ranks of the local sums of squares (spgpu?nrm2sqDev leaves them on the device)."""
    return torch.sqrt(global_dot(local_sum_of_squares, group))
