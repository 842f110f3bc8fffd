"""Synthetic SPD test matrices and the geometric nested-dissection ordering for regular meshes.

Definitions follow the reference's own MATLAB tools (restated in numpy, not translated line by line):
  * stencils and ``A = npoints*I - Adj``:  MATLAB_Tools/MESHND/meshsparse.m:64-90,144
  * recursive middle-plane bisection, stop when max dim <= 2, separator ordered last:
    MATLAB_Tools/MESHND/meshnd.m:79-114 (``nd2``)
Grid index is column-major like MATLAB's ``G(:)``: node (i,j,k) of an m-by-n-by-k mesh has index i + m*(j + n*k).
The hex-FEM elasticity matrix is NOT defined by the reference; we define it here.
All matrices are returned as scipy CSC, upper triangle (``stype=+1`` in CHOLMOD terms), int64 indices, sorted.
"""
from __future__ import annotations
import sys
import numpy as np
import scipy.sparse as sp


def mesh_index(m: int, n: int, k: int) -> np.ndarray:
    return np.arange(m * n * k, dtype=np.int64).reshape((m, n, k), order="F")


def meshnd_perm(m: int, n: int, k: int = 1) -> np.ndarray:
    """Nested-dissection permutation p (0-based): new position t holds old node p[t]."""
    G = mesh_index(m, n, k)
    out = np.empty(m * n * k, dtype=np.int64)
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))

    def rec(g: np.ndarray, pos: int) -> int:
        mm, nn, kk = g.shape
        if max(mm, nn, kk) <= 2:
            cnt = g.size
            out[pos:pos + cnt] = g.ravel(order="F")
            return pos + cnt
        if kk >= max(mm, nn):
            s = (kk + 1) // 2          # ceil(k/2), 1-based middle slice -> 0-based s-1
            a, b, mid = g[:, :, :s - 1], g[:, :, s:], g[:, :, s - 1]
        elif nn >= max(mm, kk):
            s = (nn + 1) // 2
            a, b, mid = g[:, :s - 1, :], g[:, s:, :], g[:, s - 1, :]
        else:
            s = (mm + 1) // 2
            a, b, mid = g[:s - 1, :, :], g[s:, :, :], g[s - 1, :, :]
        if a.size:
            pos = rec(a, pos)
        if b.size:
            pos = rec(b, pos)
        cnt = mid.size
        out[pos:pos + cnt] = mid.ravel(order="F")
        return pos + cnt

    end = rec(G, 0)
    assert end == out.size
    return out


def _stencil(npoints: int):
    if npoints == 7:
        return [(-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, -1), (0, 0, 1)]
    if npoints == 27:
        return [(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1) if (i, j, k) != (0, 0, 0)]
    if npoints == 5:
        return [(-1, 0, 0), (1, 0, 0), (0, 1, 0), (0, -1, 0)]
    if npoints == 9:
        return [(i, j, 0) for i in (-1, 0, 1) for j in (-1, 0, 1) if (i, j) != (0, 0)]
    raise ValueError("stencil must be 5, 7, 9 or 27")


def laplacian(N: int | tuple, stencil: int = 7) -> sp.csc_matrix:
    """``A = npts*I - Adj`` on an N^3 (or (m,n,k)) mesh, upper triangle, CSC int64 sorted."""
    m, n, k = (N, N, N) if np.isscalar(N) else N
    G = mesh_index(m, n, k)
    offs = _stencil(stencil)
    npts = len(offs)
    rows, cols = [], []
    for (di, dj, dk) in offs:
        # keep only the upper triangle: neighbour index larger than own index
        if (di + m * (dj + n * dk)) <= 0:
            continue
        i0, i1 = max(0, -di), min(m, m - di)
        j0, j1 = max(0, -dj), min(n, n - dj)
        k0, k1 = max(0, -dk), min(k, k - dk)
        g1 = G[i0:i1, j0:j1, k0:k1]
        g2 = G[i0 + di:i1 + di, j0 + dj:j1 + dj, k0 + dk:k1 + dk]
        rows.append(g1.ravel())
        cols.append(g2.ravel())
    nn = m * n * k
    rows.append(np.arange(nn, dtype=np.int64))
    cols.append(np.arange(nn, dtype=np.int64))
    r = np.concatenate(rows)
    c = np.concatenate(cols)
    v = np.full(r.shape, -1.0)
    v[-nn:] = float(npts)
    A = sp.csc_matrix((v, (r, c)), shape=(nn, nn))
    A.sort_indices()
    A.indices = A.indices.astype(np.int64)
    A.indptr = A.indptr.astype(np.int64)
    return A


def _hex_element_stiffness(E: float = 1.0, nu: float = 0.3) -> np.ndarray:
    """24x24 stiffness of a unit trilinear (Q1) hexahedron, 2x2x2 Gauss quadrature, isotropic."""
    lam = E * nu / ((1 + nu) * (1 - 2 * nu))
    mu = E / (2 * (1 + nu))
    D = np.zeros((6, 6))
    D[:3, :3] = lam
    D[np.arange(3), np.arange(3)] += 2 * mu
    D[np.arange(3, 6), np.arange(3, 6)] = mu
    corners = np.array([(i, j, k) for k in (0, 1) for j in (0, 1) for i in (0, 1)], dtype=float)  # local node a = i+2j+4k
    gp = 0.5 + np.array([-1, 1]) * 0.5 / np.sqrt(3.0)
    Ke = np.zeros((24, 24))
    for x in gp:
        for y in gp:
            for z in gp:
                pt = np.array([x, y, z])
                dN = np.zeros((8, 3))
                for a in range(8):
                    f = [(pt[d] if corners[a, d] == 1 else 1 - pt[d]) for d in range(3)]
                    g = [(1.0 if corners[a, d] == 1 else -1.0) for d in range(3)]
                    dN[a] = [g[0] * f[1] * f[2], f[0] * g[1] * f[2], f[0] * f[1] * g[2]]
                B = np.zeros((6, 24))
                for a in range(8):
                    dx, dy, dz = dN[a]
                    B[0, 3 * a] = dx; B[1, 3 * a + 1] = dy; B[2, 3 * a + 2] = dz
                    B[3, 3 * a] = dy; B[3, 3 * a + 1] = dx
                    B[4, 3 * a + 1] = dz; B[4, 3 * a + 2] = dy
                    B[5, 3 * a] = dz; B[5, 3 * a + 2] = dx
                Ke += 0.125 * (B.T @ D @ B)       # weight 1/8 each on the unit cube
    return Ke


def elasticity(N: int, E: float = 1.0, nu: float = 0.3) -> sp.csc_matrix:
    """3-DOF/node Q1 hex-FEM linear elasticity on an N^3 node grid (N-1)^3 unit elements, face i=0 clamped by
    replacing the clamped DOFs' rows/columns with identity.  DOF order 3*node+c.  Upper triangle, CSC int64."""
    Ke = _hex_element_stiffness(E, nu)
    m = n = k = N
    G = mesh_index(m, n, k)
    loc = [(i, j, kk) for kk in (0, 1) for j in (0, 1) for i in (0, 1)]
    nn = m * n * k
    # accumulate per node-offset 3x3 blocks: W[delta][c][d][node_a]
    acc = {}
    for a, (ai, aj, ak) in enumerate(loc):
        for b, (bi, bj, bk) in enumerate(loc):
            dlt = (bi - ai, bj - aj, bk - ak)
            lin = dlt[0] + m * (dlt[1] + n * dlt[2])
            if lin < 0:
                continue               # upper triangle at node level (diagonal block handled below)
            W = acc.setdefault(dlt, np.zeros((3, 3, m, n, k)))
            # elements e=(ex,ey,ez), 0<=e<N-1; node_a = e + loc[a]
            sl = (slice(ai, ai + m - 1), slice(aj, aj + n - 1), slice(ak, ak + k - 1))
            for c in range(3):
                for d in range(3):
                    W[c, d][sl] += Ke[3 * a + c, 3 * b + d]
    clamped = np.zeros((m, n, k), dtype=bool)
    clamped[0, :, :] = True
    rows, cols, vals = [], [], []
    for dlt, W in acc.items():
        di, dj, dk = dlt
        i0, i1 = max(0, -di), min(m, m - di)
        j0, j1 = max(0, -dj), min(n, n - dj)
        k0, k1 = max(0, -dk), min(k, k - dk)
        ga = G[i0:i1, j0:j1, k0:k1]
        gb = G[i0 + di:i1 + di, j0 + dj:j1 + dj, k0 + dk:k1 + dk]
        free = ~(clamped[i0:i1, j0:j1, k0:k1] | clamped[i0 + di:i1 + di, j0 + dj:j1 + dj, k0 + dk:k1 + dk])
        ga_f, gb_f = ga[free], gb[free]
        for c in range(3):
            for d in range(3):
                if dlt == (0, 0, 0) and d < c:
                    continue           # upper triangle inside the diagonal block
                w = W[c, d][i0:i1, j0:j1, k0:k1][free]
                rows.append(3 * ga_f + c); cols.append(3 * gb_f + d); vals.append(w)
    cl = G[clamped]
    for c in range(3):
        rows.append(3 * cl + c); cols.append(3 * cl + c); vals.append(np.ones(cl.size))
    r = np.concatenate(rows); c_ = np.concatenate(cols); v = np.concatenate(vals)
    A = sp.csc_matrix((v, (r, c_)), shape=(3 * nn, 3 * nn))
    A.sort_indices()
    A.indices = A.indices.astype(np.int64)
    A.indptr = A.indptr.astype(np.int64)
    return A


def graded_laplacian(N: int, contrast: float = 1e10, seed: int = 42):
    """Ill-conditioned SPD test matrix: 7-point variable-coefficient diffusion on an N^3 grid, node coefficients
    c = contrast**u with u graded along x plus a seeded random part, edge weight = harmonic mean of its two nodes,
    Dirichlet closure (a missing neighbour contributes the node's own coefficient to the diagonal).
    cond(A) ~ contrast * N^2.  Returns (A_upper_csc, nested-dissection permutation)."""
    m = n = k = N
    G = mesh_index(m, n, k)
    rng = np.random.default_rng(seed)
    u = np.linspace(0.0, 1.0, m)[:, None, None] * 0.7 + 0.3 * rng.random((m, n, k))
    c = contrast ** (u - 0.5)
    nn = m * n * k
    diag = np.zeros((m, n, k))
    rows, cols, vals = [], [], []
    for ax in range(3):
        sl_a = [slice(None)] * 3; sl_b = [slice(None)] * 3
        sl_a[ax] = slice(0, -1); sl_b[ax] = slice(1, None)
        ca, cb = c[tuple(sl_a)], c[tuple(sl_b)]
        w = 2.0 * ca * cb / (ca + cb)
        diag[tuple(sl_a)] += w; diag[tuple(sl_b)] += w
        rows.append(G[tuple(sl_a)].ravel()); cols.append(G[tuple(sl_b)].ravel()); vals.append(-w.ravel())
        # Dirichlet closure on the two faces of this axis
        f0 = [slice(None)] * 3; f1 = [slice(None)] * 3
        f0[ax] = 0; f1[ax] = -1
        diag[tuple(f0)] += c[tuple(f0)]; diag[tuple(f1)] += c[tuple(f1)]
    rows.append(np.arange(nn, dtype=np.int64)); cols.append(np.arange(nn, dtype=np.int64)); vals.append(diag.ravel(order="F"))
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nn, nn))
    A.sort_indices()
    A.indices = A.indices.astype(np.int64); A.indptr = A.indptr.astype(np.int64)
    return A, meshnd_perm(m, n, k)


def expand_perm(p: np.ndarray, ndof: int) -> np.ndarray:
    """Node permutation -> DOF permutation (DOF = ndof*node + c)."""
    return (ndof * p[:, None] + np.arange(ndof, dtype=np.int64)[None, :]).ravel()


def make_problem(kind: str, N: int):
    """Returns (A_upper_csc, perm) for kind in {'lap7','lap27','elas'}."""
    if kind == "lap7":
        return laplacian(N, 7), meshnd_perm(N, N, N)
    if kind == "lap27":
        return laplacian(N, 27), meshnd_perm(N, N, N)
    if kind == "elas":
        return elasticity(N), expand_perm(meshnd_perm(N, N, N), 3)
    raise ValueError(kind)
