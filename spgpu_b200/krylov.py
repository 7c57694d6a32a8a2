"""Conjugate-gradient iteration built on the C ABI (BASELINE configs[4]: "plus CG
step (spmv+dot+axpby)").  Caller-side code -- the reference ships no solver, its
users write this loop around spgpu?hellspmv / spgpu?dot / spgpu?axpby.

Two flavours of the same recurrence:

* `step_blocking`  what a spGPU user writes today: spgpuDhellspmv, spgpuDdot
  (blocks the host, returns the value), spgpuDaxpby with host scalars --
  6 kernels and 2 host synchronisations per iteration.
* `step_device`    the additive entry points of include/spgpu_ext.h: the SpMV is
  fused with p.Ap (spgpuDhellspmvDot); the x and r updates and r.r are one pass
  over the four vectors (spgpuDcgUpdateDev); the p update reads beta = rr'/rr
  from device memory (spgpuDaxpbyDev) -- 4 launches, no host synchronisation,
  CUDA-graph capturable.  With several GPUs the halo of the new p travels inside
  the SpMV kernel (spgpuDhellspmvHaloDot, spgpu_b200/mg.py) and the two scalars are
  all-reduced over NVLink peer memory by the LAST CTA of the kernel that produces
  them (`ar_fused`, a mg.PeerAllreduce) -- still 4 launches; with `allreduce`
  only (e.g. NCCL) each scalar costs a separate collective.

Vectors of a partition: p lives inside p_ext = [halo | owned | halo]; x, r, Ap
are owned-length.
"""
from __future__ import annotations

import torch

from .capi import TYPES


class CgState:
    def __init__(self, n_owned, halo, device, p_ext=None):
        f64 = torch.float64
        self.n, self.halo = n_owned, halo
        self.x = torch.zeros(n_owned, dtype=f64, device=device)
        self.r = torch.zeros(n_owned, dtype=f64, device=device)
        self.ap = torch.zeros(n_owned, dtype=f64, device=device)
        self.p_ext = p_ext if p_ext is not None else torch.zeros(n_owned + 2 * halo, dtype=f64, device=device)
        self.p = self.p_ext[halo:halo + n_owned]
        # device scalars: [rr, pAp, rr_new, spare]
        self.s = torch.zeros(4, dtype=f64, device=device)
        self.rr_host = 0.0


class Cg:
    """apply_A(z_tensor, x_ext_tensor) must compute z = A * x_ext (incl. halo exchange);
    apply_A_dot(z, x_ext, d_out_ptr, ar_ref) optionally the fused SpMV + p.Ap, where ar_ref is None or the
    spgpuPeerAllreduce* to hand to spgpu?h{ell,dia}spmvHaloDot.  allreduce(t): in-place sum of a 1-element
    tensor over the ranks (a collective of its own); ar_fused: a mg.PeerAllreduce whose exchange the
    producing kernels run themselves."""

    def __init__(self, L, handle, state: CgState, apply_A, apply_A_dot=None, allreduce=None, ar_fused=None,
                 pingpong=False):
        self.L, self.h, self.st = L, handle, state
        self.apply_A, self.apply_A_dot, self.allreduce, self.ar_fused = apply_A, apply_A_dot, allreduce, ar_fused
        self.T = TYPES["D"]
        # pingpong: r.r lives alternately in scalar slot 0 and slot 2, so that "rr <- rr'" is no kernel at all (4 launches
        # per iteration instead of 5).  Off by default because a CUDA graph of ONE iteration freezes the slots; capture
        # two iterations, or leave it off, when replaying a graph.
        self.pingpong = pingpong
        self.k = 0

    def start(self, b: torch.Tensor):
        """x = 0, r = p = b, rr = b.b"""
        st, L, h, n = self.st, self.L, self.h, self.st.n
        st.x.zero_()
        st.r.copy_(b)
        st.p.copy_(b)
        L.spgpuDdotDev(h, n, st.r.data_ptr(), st.r.data_ptr(), st.s.data_ptr())
        if self.allreduce:
            self.allreduce(st.s[0:1])
        st.rr_host = float(st.s[0].item())
        self.k = 0
        return st.rr_host

    # ---- reference-style: blocking dots, host scalars -----------------------
    def step_blocking(self):
        st, L, h, n, T = self.st, self.L, self.h, self.st.n, self.T
        self.apply_A(st.ap, st.p_ext)
        pap = L.spgpuDdot(h, n, st.p.data_ptr(), st.ap.data_ptr())
        if self.allreduce:
            t = torch.tensor([pap], dtype=torch.float64, device=st.x.device)
            self.allreduce(t)
            pap = float(t.item())
        alpha = st.rr_host / pap
        L.spgpuDaxpby(h, st.x.data_ptr(), n, T.scalar(1.0), st.x.data_ptr(), T.scalar(alpha), st.p.data_ptr())
        L.spgpuDaxpby(h, st.r.data_ptr(), n, T.scalar(1.0), st.r.data_ptr(), T.scalar(-alpha), st.ap.data_ptr())
        rr_new = L.spgpuDdot(h, n, st.r.data_ptr(), st.r.data_ptr())
        if self.allreduce:
            t = torch.tensor([rr_new], dtype=torch.float64, device=st.x.device)
            self.allreduce(t)
            rr_new = float(t.item())
        beta = rr_new / st.rr_host
        L.spgpuDaxpby(h, st.p.data_ptr(), n, T.scalar(beta), st.p.data_ptr(), T.scalar(1.0), st.r.data_ptr())
        st.rr_host = rr_new
        return rr_new

    # ---- device scalars: no host synchronisation ----------------------------
    def step_device(self, mark=None):
        """one iteration; mark(label), if given, is called after each phase has been queued (bench.py records an
        event there to time the phases)"""
        st, L, h, n = self.st, self.L, self.h, self.st.n
        s = st.s.data_ptr()
        cur = 2 * (self.k & 1) if self.pingpong else 0       # slot of r.r of the current iterate
        nxt = 2 - cur
        rr, pap, rrn = s + 8 * cur, s + 8, s + 8 * nxt
        fused = self.ar_fused
        mark = mark or (lambda _label: None)
        if self.apply_A_dot is not None:
            # SpMV (+ halo exchange) fused with p.Ap (+ its all-reduce in the fold kernel's last CTA)
            self.apply_A_dot(st.ap, st.p_ext, pap, fused.next_ref() if fused else None)
            if self.allreduce and not fused:
                self.allreduce(st.s[1:2])
            mark("spmv+halo+p.Ap (+fold, all-reduce)")
        else:
            self.apply_A(st.ap, st.p_ext)
            L.spgpuDdotDev(h, n, st.p.data_ptr(), st.ap.data_ptr(), pap)
            if self.allreduce:
                self.allreduce(st.s[1:2])
            mark("spmv, p.Ap, all-reduce")
        # x += (rr/pAp) p ;  r -= (rr/pAp) Ap ;  rr' = r.r   -- one pass (spgpuDcgUpdateDev)
        L.spgpuDcgUpdateDev(h, st.x.data_ptr(), st.r.data_ptr(), st.p.data_ptr(), st.ap.data_ptr(), n, rr, pap, rrn,
                            fused.next_ref() if fused else None)
        if self.allreduce and not fused:
            self.allreduce(st.s[nxt:nxt + 1])
        mark("x, r update + r.r (+all-reduce)")
        # p = r + (rr'/rr) p ; then rr <- rr'
        L.spgpuDaxpbyDev(h, st.p.data_ptr(), n, rrn, rr, 1.0, st.p.data_ptr(), 0, 0, 1.0, st.r.data_ptr())
        if self.pingpong:
            self.k += 1                                       # the slots swap roles: nothing to copy
        else:
            # through the library so that it is ordered on the HANDLE's stream whatever torch's current stream is
            L.spgpuDscal(h, rr, 1, self.T.scalar(1.0), rrn)
        mark("p update")

    def residual_norm2(self):
        return float(self.st.s[2 * (self.k & 1) if self.pingpong else 0].item())


CG_BYTES_PER_ROW_VECTOR_OPS = {
    # algorithmic bytes per row of the vector part of one iteration (double):
    # blocking: p.Ap 16 + x update 24 + r update 24 + r.r 8 + p update 24
    "blocking": 96,
    # device: p.Ap fused into the SpMV epilogue (p re-read 8) + fused x/r update 48 + p update 24
    "device": 80,
}
