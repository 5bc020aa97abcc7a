"""numpy emulation of the device algorithms driven by the symbolic structures the
library exports (test infrastructure: lets the CPU-only suite validate the gather
map, ordering, supernode structures, extend-add maps and panel offsets that the
CUDA kernels consume, before any GPU time is spent)."""
import numpy as np


class Sym:
    def __init__(self, h):
        g = h.symbolic
        self.n = int(h.info("n"))
        self.perm = g("perm"); self.sfirst = g("sfirst"); self.sparent = g("sparent")
        self.rowptr = g("rowptr"); self.rowidx = g("rowidx"); self.rel = g("rel")
        self.Loff = g("Loff"); self.CBoff = g("CBoff"); self.level = g("level")
        self.amap = g("amap"); self.dpos = g("dpos"); self.Mp = g("Mp"); self.Mi = g("Mi")
        self.nsuper = len(self.sfirst) - 1
        self.nnzL = int(h.info("nnzL"))
        self.cb_total = int(h.info("cb_total"))
        self.gptr = g("gptr"); self.gsrc = g("gsrc")


def assemble_M(h, Jx, Hx, y, s):
    """tril(J'DJ + H) through the gather map, in the kernel's operation order."""
    pp = h.symbolic("pair_ptr"); A = h.symbolic("pairA"); B = h.symbolic("pairB"); hm = h.symbolic("hmap")
    sig = y / s
    n = int(h.info("n"))
    # row of every J entry
    T = None
    nnzM = len(hm)
    M = np.zeros(nnzM)
    return pp, A, B, hm, sig


def assemble_M_values(h, Jp, Ji, Jx, Hx, y, s):
    pp = h.symbolic("pair_ptr"); A = h.symbolic("pairA"); B = h.symbolic("pairB"); hm = h.symbolic("hmap")
    sig = y / s
    T = Jx * sig[Ji]
    nnzM = len(hm)
    M = np.zeros(nnzM)
    prod = T[A] * Jx[B]
    for e in range(nnzM):
        t0, t1 = pp[e], pp[e + 1]
        have = False
        acc = 0.0
        if t1 > t0:
            acc = prod[t0]
            for t in range(t0 + 1, t1):
                acc = acc + prod[t]
            have = True
        if hm[e] >= 0:
            acc = acc + Hx[hm[e]] if have else Hx[hm[e]]
        M[e] = acc
    return M


def factor(S, Mval, delta, mode="chol"):
    """Multifrontal factorisation with the exported maps.  Returns (ok, Lval)."""
    L = np.zeros(S.nnzL)
    L[S.amap] = Mval
    L[S.dpos] += delta
    # update blocks live in ONE arena at CBoff (level-lifetime allocator, symbolic.cpp).  The device
    # runs the supernodes of a level concurrently, so per level every front first READS its
    # children's blocks and only then are the level's own blocks WRITTEN.
    CB = np.full(max(S.cb_total, 1), np.nan)
    children = [[] for _ in range(S.nsuper)]
    for s in range(S.nsuper):
        if S.sparent[s] >= 0:
            children[S.sparent[s]].append(s)
    nlev = int(S.level.max()) + 1 if S.nsuper else 0
    by_level = [np.nonzero(S.level == l)[0] for l in range(nlev)]
    for lev in by_level:
        produced = []
        for s in lev:
            f = S.sfirst[s]; c = S.sfirst[s + 1] - f
            r = S.rowptr[s + 1] - S.rowptr[s]; N = c + r
            F = np.zeros((N, N))
            ld = (N + 1) & ~1     # panel leading dimension (symbolic.h: panel_ld)
            F[:, :c] = L[S.Loff[s]:S.Loff[s] + ld * c].reshape(c, ld).T[:N]
            for ch in children[s]:
                rel = S.rel[S.rowptr[ch]:S.rowptr[ch + 1]]
                rc = len(rel)
                cb = CB[S.CBoff[ch]:S.CBoff[ch] + rc * rc].reshape(rc, rc).T     # column-major, ld = rc
                assert not np.isnan(np.tril(cb)).any(), "child update block was overwritten or never written"
                idx = np.ix_(rel, rel)
                F[idx] += np.tril(cb)
            for j in range(c):
                d = F[j, j]
                if mode == "chol":
                    if not (d > 0):
                        return False, L
                    ljj = np.sqrt(d)
                    F[j + 1:, j] /= ljj
                    F[j, j] = ljj
                    col = F[j + 1:, j]
                    F[j + 1:, j + 1:] -= np.tril(np.outer(col, col))
                else:
                    if d == 0 or d != d:
                        return False, L
                    w = F[j + 1:, j].copy()
                    F[j + 1:, j] = w / d
                    F[j + 1:, j + 1:] -= np.tril(np.outer(F[j + 1:, j], w))
            P = np.zeros((ld, c)); P[:N] = F[:, :c]
            L[S.Loff[s]:S.Loff[s] + ld * c] = P.T.reshape(-1)
            if S.sparent[s] >= 0 and r > 0:
                produced.append((s, F[c:, c:].copy()))
        for s, blk in produced:
            r = blk.shape[0]
            CB[S.CBoff[s]:S.CBoff[s] + r * r] = blk.T.reshape(-1)
        # a reused region now holds another block's values: the comparison of the resulting factor
        # with the oracle (tests/test_symbolic.py) catches any clobbering of a live block
    return True, L


def solve(S, L, b, mode="chol"):
    n = S.n
    x = b[S.perm].astype(float).copy()
    uflat = np.zeros(S.rowptr[-1])
    children = [[] for _ in range(S.nsuper)]
    for s in range(S.nsuper):
        if S.sparent[s] >= 0:
            children[S.sparent[s]].append(s)
    order = np.argsort(S.level, kind="stable")
    panels = {}
    for s in order:
        f = S.sfirst[s]; c = S.sfirst[s + 1] - f
        r = S.rowptr[s + 1] - S.rowptr[s]; N = c + r
        ld = (N + 1) & ~1
        P = L[S.Loff[s]:S.Loff[s] + ld * c].reshape(c, ld).T[:N]
        panels[s] = P
        # children's update vectors through the gather lists (gptr/gsrc), as fwd_kernel does
        gb = S.rowptr[s] + f
        acc = np.array([uflat[S.gsrc[S.gptr[gb + d]:S.gptr[gb + d + 1]]].sum() for d in range(N)])
        x[f:f + c] += acc[:c]
        us = acc[c:].copy()
        L11 = np.tril(P[:c, :c])
        if mode != "chol":
            L11 = np.tril(L11, -1) + np.eye(c)
        yv = np.linalg.solve(L11, x[f:f + c]) if c else x[f:f + c]
        x[f:f + c] = yv
        us -= P[c:, :c] @ yv
        uflat[S.rowptr[s]:S.rowptr[s + 1]] = us
    for s in order[::-1]:
        f = S.sfirst[s]; c = S.sfirst[s + 1] - f
        r = S.rowptr[s + 1] - S.rowptr[s]
        P = panels[s]
        rows = S.rowidx[S.rowptr[s]:S.rowptr[s + 1]]
        L11 = np.tril(P[:c, :c])
        rhs = x[f:f + c].copy()
        if mode != "chol":
            rhs = rhs / np.diag(P[:c, :c])
            L11 = np.tril(L11, -1) + np.eye(c)
        rhs -= P[c:, :c].T @ x[rows]
        x[f:f + c] = np.linalg.solve(L11.T, rhs)
    out = np.empty(n)
    out[S.perm] = x
    return out
