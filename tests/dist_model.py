"""CPU model of the slab-partitioned V-cycle (test infrastructure).

It executes, rank by rank over `torch.distributed` (gloo), the SAME schedule
libmgb runs over NCCL (multigrid_parallel_b200/csrc/api.cu: halo_after_sweep,
the deep-halo fetch before the fused residual+restriction, the coarse-rhs halo,
the gather onto rank 0, the broadcast of the agglomerated correction), with the
serial oracle (oracle/mg_oracle.c) doing the arithmetic on each slab.  The slab
planner is the library's own (mgb_plan_slab / mgb_plan_first_dist_level, pure
host arithmetic).  If the schedule is right the slabs equal the serial oracle's
arrays bit for bit."""
import ctypes as C
import math

import numpy as np
import torch
import torch.distributed as dist


def plan_slab(lib, ni, nranks, rank):
    a, b = C.c_int(), C.c_int()
    assert lib.mgb_plan_slab(ni, nranks, rank, a, b) == 0
    return a.value, b.value


class SlabLevel:
    def __init__(self, lib, shape, rank, world, distributed, h):
        self.ni, self.nj, self.nk = shape
        self.h = h
        self.dist = distributed
        if distributed:
            self.own_lo, self.own_hi = plan_slab(lib, self.ni, world, rank)
            self.lower = 2 if rank > 0 else 0
            self.upper = 1 if rank < world - 1 else 0
        else:
            self.own_lo, self.own_hi, self.lower, self.upper = 0, self.ni, 0, 0
        self.i0 = self.own_lo - self.lower
        self.li = self.own_hi - self.own_lo + self.lower + self.upper
        self.u = np.zeros((self.li, self.nj, self.nk))
        self.d = np.zeros((self.li, self.nj, self.nk))

    def loc(self, plane):
        return plane - self.i0

    @property
    def sweep(self):
        return max(self.own_lo, 1), min(self.own_hi, self.ni - 1)


class SlabMG:
    """mirror of mgb_create_dist + enqueue_cycle for one rank"""

    def __init__(self, lib, orc, coarse, levels, gs, rank, world, min_planes=2, min_points=0,
                 shortcuts=False, fused=False):
        # fused: the schedule of the P2P path (csrc/halo.cuh, api.cu) -- every half-sweep
        # hands the swept COLOUR of its planes own_hi-2, own_hi-1 to the upper neighbour and
        # of own_lo to the lower one (so no extra exchange precedes the restriction), the
        # first unpartitioned level's right-hand side is all-gathered and the levels below
        # are computed redundantly on every rank (no broadcast)
        self.fused = fused
        # shortcuts: the two exact short-cuts libmgb's cycle takes -- coarse levels are
        # not zeroed (first RED half-sweep with the guess taken as 0), the
        # prolongation corrects RED points only (api.cu: enqueue_cycle, q_prolong)
        self.shortcuts = shortcuts and gs >= 1
        self.lib, self.orc, self.rank, self.world, self.gs = lib, orc, rank, world, gs
        self.L = levels
        self.LD = lib.mgb_plan_first_dist_level(*coarse, levels, world, min_planes, min_points)
        assert 1 <= self.LD < levels
        shapes = [tuple((c - 1) * (1 << l) + 1 for c in coarse) for l in range(levels)]
        hf = 1.0 / (shapes[-1][2] - 1)
        self.lv = []
        for l in range(levels):
            h = hf
            for _ in range(levels - 1 - l):
                h = 2 * h
            self.lv.append(SlabLevel(lib, shapes[l], rank, world, l >= self.LD, h))
        n0 = int(np.prod(shapes[0]))
        self.lu = orc.coarse_matrix(shapes[0], self.lv[0].h)
        orc.lu_factor(self.lu)
        self.n0 = n0

    # -- communication (gloo), same pattern as halo_step ------------------
    def _xchg(self, arr, lv, send_up, recv_low, send_down, recv_up):
        ops = []
        has_low, has_up = self.rank > 0, self.rank < self.world - 1
        bufs = []
        if has_up and send_up is not None:
            t = torch.from_numpy(arr[lv.loc(send_up)].copy())
            ops.append(dist.P2POp(dist.isend, t, self.rank + 1))
        if has_low and recv_low is not None:
            t = torch.empty((lv.nj, lv.nk), dtype=torch.float64)
            bufs.append((t, recv_low))
            ops.append(dist.P2POp(dist.irecv, t, self.rank - 1))
        if has_low and send_down is not None:
            t = torch.from_numpy(arr[lv.loc(send_down)].copy())
            ops.append(dist.P2POp(dist.isend, t, self.rank - 1))
        if has_up and recv_up is not None:
            t = torch.empty((lv.nj, lv.nk), dtype=torch.float64)
            bufs.append((t, recv_up))
            ops.append(dist.P2POp(dist.irecv, t, self.rank + 1))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for t, plane in bufs:
            arr[lv.loc(plane)] = t.numpy()

    def works_on(self, q):
        return q >= self.LD or self.rank == 0 or self.fused

    def _xchg_colour(self, lv, colour):
        """fused schedule: the swept colour of the slab's boundary planes"""
        has_low, has_up = self.rank > 0, self.rank < self.world - 1
        ops, bufs = [], []
        if has_up:
            for pl in (lv.own_hi - 1, lv.own_hi - 2):
                t = torch.from_numpy(lv.u[lv.loc(pl)].copy())
                ops.append(dist.P2POp(dist.isend, t, self.rank + 1))
        if has_low:
            for pl in (lv.own_lo - 1, lv.own_lo - 2):
                t = torch.empty((lv.nj, lv.nk), dtype=torch.float64)
                bufs.append((t, pl))
                ops.append(dist.P2POp(dist.irecv, t, self.rank - 1))
            t = torch.from_numpy(lv.u[lv.loc(lv.own_lo)].copy())
            ops.append(dist.P2POp(dist.isend, t, self.rank - 1))
        if has_up:
            t = torch.empty((lv.nj, lv.nk), dtype=torch.float64)
            bufs.append((t, lv.own_hi))
            ops.append(dist.P2POp(dist.irecv, t, self.rank + 1))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for t, plane in bufs:
            # only the INTERIOR points of the swept colour travel (the kernels mirror exactly
            # the stores they make)
            sel = self._colour_mask(lv, plane, plane + 1, colour)[0]
            sel[0, :] = sel[-1, :] = False
            sel[:, 0] = sel[:, -1] = False
            lv.u[lv.loc(plane)][sel] = t.numpy()[sel]

    def _after_sweep(self, lv, colour):
        if not lv.dist:
            return
        if self.fused:
            self._xchg_colour(lv, colour)
        else:
            self._xchg(lv.u, lv, lv.own_hi - 1, lv.own_lo - 1, lv.own_lo, lv.own_hi)

    # -- operators ---------------------------------------------------------
    def half_sweep(self, q, colour):
        if not self.works_on(q):
            return
        lv = self.lv[q]
        lo, hi = lv.sweep
        if hi > lo:
            a, b = lv.loc(lo - 1), lv.loc(hi + 1)
            # the oracle colours by LOCAL i: shift by the parity of the box origin
            self.orc.half_sweep(lv.u[a:b], lv.d[a:b], lv.h, colour ^ ((lo - 1) & 1))
        self._after_sweep(lv, colour)

    def _colour_mask(self, lv, lo, hi, colour):
        """points of global colour `colour` on local planes [lo, hi) (global indices)"""
        i, j, k = np.meshgrid(np.arange(lo, hi), np.arange(lv.nj), np.arange(lv.nk), indexing="ij")
        return ((i + j + k) & 1) == colour

    def first_sweep_zero(self, q, colour):
        """k_first_sweep_zero: the update with all six neighbours 0, owned interior
        planes only, nothing of the old array is read"""
        if not self.works_on(q):
            return
        lv = self.lv[q]
        lo, hi = lv.sweep
        if hi > lo:
            s = np.zeros((hi - lo, lv.nj, lv.nk))
            for _ in range(5):
                s = s + 0.0
            new = (1.0 / 6) * (s - (lv.h * lv.h) * lv.d[lv.loc(lo):lv.loc(hi)])
            sel = self._colour_mask(lv, lo, hi, colour)
            sel[:, 0, :] = sel[:, -1, :] = False
            sel[:, :, 0] = sel[:, :, -1] = False
            lv.u[lv.loc(lo):lv.loc(hi)][sel] = new[sel]
        self._after_sweep(lv, colour)

    def smooth(self, q, first_red, zero_guess=False):
        for it in range(self.gs):
            if zero_guess and it == 0:
                self.first_sweep_zero(q, 1 if first_red else 0)
            else:
                self.half_sweep(q, 1 if first_red else 0)
            self.half_sweep(q, 0 if first_red else 1)

    def residual_sumsq(self, q):
        lv = self.lv[q]
        lo, hi = lv.sweep
        a, b = lv.loc(lo - 1), lv.loc(hi + 1)
        n = self.orc.residual(lv.u[a:b], lv.d[a:b], lv.h) if self.works_on(q) else 0.0
        t = torch.tensor([n * n], dtype=torch.float64)
        if lv.dist:
            dist.all_reduce(t)
        return float(t[0])

    def residual_restrict(self, q):
        if not self.works_on(q):
            return
        f, c = self.lv[q], self.lv[q - 1]
        if not f.dist:
            r = np.zeros_like(f.u)
            self.orc.residual(f.u, f.d, f.h, r)
            self.orc.restrict(r, c.d)
            return
        if not self.fused:
            self._xchg(f.u, f, f.own_hi - 2, f.own_lo - 2, None, None)  # deep halo
        Ilo, Ihi = plan_slab(self.lib, c.ni, self.world, self.rank)
        Im0, Im1 = max(Ilo, 1), min(Ihi, c.ni - 1)  # interior coarse planes
        # fine box [2*Im0-2, 2*Im1]: its interior planes carry the residuals needed
        fa, fb = f.loc(2 * Im0 - 2), f.loc(2 * Im1) + 1
        rbox = np.zeros((fb - fa, f.nj, f.nk))
        self.orc.residual(f.u[fa:fb], f.d[fa:fb], f.h, rbox)
        cbox = np.zeros((Im1 - Im0 + 2, c.nj, c.nk))
        self.orc.restrict(rbox, cbox)
        c.d[c.loc(Im0):c.loc(Im1)] = cbox[1:-1]
        for I in (0, c.ni - 1):  # injected boundary residual = 0
            if Ilo <= I < Ihi:
                c.d[c.loc(I)] = 0.0
        if c.dist:
            self._xchg(c.d, c, c.own_hi - 1, c.own_lo - 1, None, None)
        elif self.fused:  # all-gather: every rank computes the levels below redundantly
            for r in range(self.world):
                lo, hi = plan_slab(self.lib, c.ni, self.world, r)
                t = torch.from_numpy(c.d[lo:hi].copy()) if r == self.rank else \
                    torch.empty((hi - lo, c.nj, c.nk), dtype=torch.float64)
                dist.broadcast(t, r)
                c.d[lo:hi] = t.numpy()
        else:  # gather on rank 0
            if self.rank > 0:
                dist.send(torch.from_numpy(c.d[Ilo:Ihi].copy()), 0)
            else:
                for r in range(1, self.world):
                    lo, hi = plan_slab(self.lib, c.ni, self.world, r)
                    t = torch.empty((hi - lo, c.nj, c.nk), dtype=torch.float64)
                    dist.recv(t, r)
                    c.d[lo:hi] = t.numpy()

    def prolong(self, q):
        if not self.works_on(q):
            return
        f = self.lv[q]
        if not self.shortcuts:
            self._prolong_full(q)
            return
        # red-only: run the reference's operation on a copy, keep its RED points on the
        # planes the rank updates; black face points of the finest level get `+ 0.`
        lo = f.own_lo - (1 if (f.dist and self.rank > 0) else 0)
        hi = f.own_hi + (1 if (f.dist and self.rank < self.world - 1) else 0)
        before = f.u.copy()
        self._prolong_full(q)
        after = f.u.copy()
        f.u[...] = before
        red = self._colour_mask(f, lo, hi, 1)
        f.u[f.loc(lo):f.loc(hi)][red] = after[f.loc(lo):f.loc(hi)][red]
        if q == self.L - 1:
            a, b = f.own_lo, f.own_hi
            face = np.zeros((b - a, f.nj, f.nk), bool)
            face[:, 0, :] = face[:, -1, :] = True
            face[:, :, 0] = face[:, :, -1] = True
            for I in (0, f.ni - 1):
                if a <= I < b:
                    face[I - a] = True
            sel = face & self._colour_mask(f, a, b, 0)
            blk = f.u[f.loc(a):f.loc(b)]
            blk[sel] = blk[sel] + 0.0

    def _prolong_full(self, q):
        f, c = self.lv[q], self.lv[q - 1]
        if not f.dist:
            self.orc.prolong_correct(c.u, f.u)
            return
        lo = f.own_lo - (1 if self.rank > 0 else 0)
        hi = f.own_hi + (1 if self.rank < self.world - 1 else 0)
        e0 = lo & ~1                      # even fine plane at or below lo
        e1 = (hi - 1) + ((hi - 1) & 1)    # even fine plane at or above hi-1
        fbox = f.u[f.loc(e0):f.loc(e1) + 1] if f.loc(e0) >= 0 and f.loc(e1) < f.li else None
        cb = c.u[c.loc(e0 // 2):c.loc(e1 // 2) + 1]
        if fbox is None:  # pad the box with scratch planes outside the local range
            tmp = np.zeros((e1 - e0 + 1, f.nj, f.nk))
            a, b = max(e0, f.i0), min(e1 + 1, f.i0 + f.li)
            tmp[a - e0:b - e0] = f.u[f.loc(a):f.loc(b)]
            self.orc.prolong_correct(cb, tmp)
            f.u[f.loc(lo):f.loc(hi)] = tmp[lo - e0:hi - e0]
        else:
            keep_lo = fbox[:lo - e0].copy()
            keep_hi = fbox[hi - e0:].copy()
            self.orc.prolong_correct(cb, fbox)
            fbox[:lo - e0] = keep_lo
            fbox[hi - e0:] = keep_hi

    def cycle_level(self, q):
        if not self.works_on(q):
            return
        lv = self.lv[q]
        zero_guess = self.shortcuts and 0 < q < self.L - 1
        if q < self.L - 1 and not zero_guess:
            lv.u[...] = 0.0
        if q == 0:
            lv.u[...] = self.orc.lu_solve(self.lu, lv.d.reshape(-1)).reshape(lv.u.shape)
            return
        self.smooth(q, True, zero_guess)
        self.residual_restrict(q)
        self.cycle_level(q - 1)
        if q == self.LD and not self.fused:
            t = torch.from_numpy(self.lv[q - 1].u)
            dist.broadcast(t, 0)
        self.prolong(q)
        self.smooth(q, False)

    def vcycle(self):
        self.cycle_level(self.L - 1)
        return math.sqrt(self.residual_sumsq(self.L - 1))
