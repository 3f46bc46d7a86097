"""Prototype (numpy) of the multi-CTA factorisation planned for ba_cholesky_kernel.

The reduced camera system of a trajectory is banded (KF-interleaved ordering, block band B): today ONE CTA walks the whole
band (42.5 of 45 ms per iteration at configs[4], 75 k unknowns x B ~ 165).  Substructuring removes the chain:

  columns = I_0 | s_0 | I_1 | s_1 | ... | s_{P-2} | I_{P-1}        separators s_k of width >= B: interiors do not see each other

  per interior (one CTA each, all in parallel):
      A_II = L L^T                                   banded Cholesky of the interior
      W    = L^-1 [A_I,left | A_I,right]             only the first / last B rows of the interior couple to a separator
      y    = L^-1 b_I
      S_ll -= W_l^T W_l ; S_rr -= W_r^T W_r ; S_lr -= W_l^T W_r      (S_lr is fill: the two separators now see each other)
      b_l  -= W_l^T y   ; b_r  -= W_r^T y
  separator system: block tridiagonal, B x B blocks, P-1 of them -> block cyclic reduction (log2 P levels, parallel per level)
  per interior again:  x_I = L^-T (y - W x_sep)

This file checks the algebra against numpy.linalg.solve on random banded SPD systems (including the LM-damped form the solver
uses) and prints the work / critical-path model that motivates it.  It is a design aid, not product code.
"""
import numpy as np


def banded_spd(n, band, rng, damping=1e-3):
    a = np.zeros((n, n))
    for i in range(n):
        lo = max(0, i - band)
        a[i, lo:i + 1] = rng.normal(size=i + 1 - lo)
    a = np.tril(a)
    s = a @ a.T                                   # SPD, half-bandwidth = band
    s[np.abs(np.subtract.outer(np.arange(n), np.arange(n))) > band] = 0.0      # (a a^T is already banded; keep it exact)
    return s + damping * np.diag(np.diag(s))


def partition(n, band, parts):
    """Column ranges: interiors and separators (width = band), interiors as equal as possible."""
    seps = parts - 1
    interior_total = n - seps * band
    assert interior_total >= parts * band, "segments shorter than the band: use fewer parts"
    sizes = [interior_total // parts + (1 if k < interior_total % parts else 0) for k in range(parts)]
    ranges, pos = [], 0
    for k in range(parts):
        ranges.append(("I", pos, pos + sizes[k])); pos += sizes[k]
        if k < seps:
            ranges.append(("S", pos, pos + band)); pos += band
    assert pos == n
    return ranges


def block_cyclic_reduction(diag, off, rhs):
    """Solve the SPD block-tridiagonal system  diag[k] x_k + off[k-1]^T... by recursive odd/even elimination.
    diag[k]: (B,B); off[k]: coupling block between k and k+1 stored as S[k+1 rows, k cols] (lower); rhs[k]: (B,)."""
    m = len(diag)
    if m == 1:
        return [np.linalg.solve(diag[0], rhs[0])]
    # eliminate the even-indexed unknowns 0, 2, 4, ... (independent of each other) into their odd neighbours
    d2, o2, r2, keep = [], [], [], list(range(1, m, 2))
    inv = {k: np.linalg.inv(diag[k]) for k in range(0, m, 2)}          # the kernel would hold Cholesky factors instead
    for j, k in enumerate(keep):
        d = diag[k].copy(); r = rhs[k].copy()
        lo = off[k - 1]                                  # block (k, k-1)
        d -= lo @ inv[k - 1] @ lo.T; r -= lo @ inv[k - 1] @ rhs[k - 1]
        if k + 1 < m:
            up = off[k]                                  # block (k+1, k): coupling of k with the even k+1
            d -= up.T @ inv[k + 1] @ up; r -= up.T @ inv[k + 1] @ rhs[k + 1]
        d2.append(d); r2.append(r)
        if j + 1 < len(keep):                            # new coupling (k+2, k) through the eliminated k+1
            o2.append(-off[k + 1] @ inv[k + 1] @ off[k])
    x_odd = block_cyclic_reduction(d2, o2, r2)
    x = [None] * m
    for j, k in enumerate(keep):
        x[k] = x_odd[j]
    for k in range(0, m, 2):
        r = rhs[k].copy()
        if k - 1 >= 0:
            r -= off[k - 1] @ x[k - 1]                   # block (k, k-1) times x_{k-1}
        if k + 1 < m:
            r -= off[k].T @ x[k + 1]                     # block (k+1, k)^T times x_{k+1}
        x[k] = inv[k] @ r
    return x


def solve_substructured(s, b, band, parts):
    n = len(b)
    ranges = partition(n, band, parts)
    interiors = [(lo, hi) for kind, lo, hi in ranges if kind == "I"]
    seps = [(lo, hi) for kind, lo, hi in ranges if kind == "S"]
    m = len(seps)
    diag = [s[lo:hi, lo:hi].copy() for lo, hi in seps]
    off = [np.zeros((band, band)) for _ in range(m - 1)]              # filled by the interiors between two separators
    rhs = [b[lo:hi].copy() for lo, hi in seps]
    saved = []
    for k, (lo, hi) in enumerate(interiors):
        L = np.linalg.cholesky(s[lo:hi, lo:hi])
        y = np.linalg.solve(L, b[lo:hi])
        Wl = Wr = None
        if k > 0:                                                       # left separator k-1
            sl, sh = seps[k - 1]
            Wl = np.linalg.solve(L, s[lo:hi, sl:sh])
            diag[k - 1] -= Wl.T @ Wl; rhs[k - 1] -= Wl.T @ y
        if k < m:                                                       # right separator k
            sl, sh = seps[k]
            Wr = np.linalg.solve(L, s[lo:hi, sl:sh])
            diag[k] -= Wr.T @ Wr; rhs[k] -= Wr.T @ y
        if Wl is not None and Wr is not None:
            off[k - 1] -= Wr.T @ Wl                                     # block (sep k rows, sep k-1 cols)
        saved.append((L, y, Wl, Wr))
    xs = block_cyclic_reduction(diag, off, rhs) if m else []
    x = np.zeros(n)
    for (lo, hi), xk in zip(seps, xs):
        x[lo:hi] = xk
    for k, (lo, hi) in enumerate(interiors):
        L, y, Wl, Wr = saved[k]
        r = y.copy()
        if Wl is not None: r -= Wl @ xs[k - 1]
        if Wr is not None: r -= Wr @ xs[k]
        x[lo:hi] = np.linalg.solve(L.T, r)
    return x


def model(n, band, parts, gflops_per_cta=60.0):
    """Flop / critical-path model (double precision, one CTA sustaining `gflops_per_cta` on these small dense blocks)."""
    seq = n * band * band                                               # banded Cholesky, one CTA: n B^2 flops
    n_i = (n - (parts - 1) * band) / parts
    interior = n_i * band * band + 2 * (2 * band) * n_i * band          # factor + W for 2B right-hand sides through the band
    schur = 3 * 2 * band ** 3                                           # S_ll, S_rr, S_lr: W is dense in ~B rows at each end of the interior
    levels = int(np.ceil(np.log2(max(2, parts - 1))))
    per_level = (1.0 / 3 + 2 + 2) * band ** 3                           # one factorisation + two solves/products of B x B per surviving block
    crit = interior + schur + levels * per_level + 2 * n_i * band * 2   # + the interior back-substitution
    t_seq = seq / (gflops_per_cta * 1e9); t_par = crit / (gflops_per_cta * 1e9)
    return dict(n=n, band=band, parts=parts, sequential_ms=1e3 * t_seq, critical_path_ms=1e3 * t_par, levels=levels, interior_cols=n_i)


if __name__ == "__main__":
    rng = np.random.default_rng(7)
    for n, band, parts in ((600, 30, 4), (1500, 45, 8), (2000, 33, 13), (900, 60, 5)):
        s = banded_spd(n, band, rng)
        b = rng.normal(size=n)
        x = solve_substructured(s, b, band, parts)
        ref = np.linalg.solve(s, b)
        err = np.max(np.abs(x - ref)) / np.max(np.abs(ref))
        print("n %5d band %3d parts %2d  rel err vs numpy %.2e" % (n, band, parts, err))
        assert err < 1e-8
    # the two workloads of BASELINE.json this is for
    # configs[4] (5 k keyframes x 15 unknowns, band ~ 165): measured 42.5 ms for the one-CTA band walk.  Windows (configs[1], [3])
    # are dense (band = n): nothing to split there, they need the cluster / DSMEM variant instead.
    for parts in (16, 64, 148):
        print("configs[4]", model(75000, 165, parts))
