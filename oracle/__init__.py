"""CPU oracle for the SpareNet hot path -- TEST INFRASTRUCTURE ONLY.

`oracle/` is the checker: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  Nothing under sparenet_b200/ does.  The arithmetic lives in
sparenet_oracle.c (each function cites the reference file:line it restates); this module only
builds it with the system gcc and marshals torch CPU tensors through ctypes.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "sparenet_oracle.c")
_lib = None


def build(force=False):
    """Compile sparenet_oracle.c -> oracle/_build/liboracle.so (gcc, OpenMP when available)."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    base = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-mavx2", "-fno-math-errno",
            "-fvisibility=hidden", "-o", _SO, _SRC, "-lm"]
    last = None
    for cc in ("/usr/bin/gcc", "gcc", "cc"):
        for omp in (["-fopenmp"], []):
            try:
                subprocess.run([cc] + omp + base, check=True, capture_output=True)
                return _SO
            except (subprocess.CalledProcessError, FileNotFoundError) as e:  # try the next recipe
                last = e
    raise RuntimeError(f"could not build the oracle: {getattr(last, 'stderr', last)}")


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_emd_fwd.restype = ctypes.c_int
        _lib.orc_expansion_fwd.restype = ctypes.c_int
        _lib.orc_mds_check.restype = ctypes.c_int
        _lib.orc_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return lib().orc_num_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))


def _f(t):
    assert t.dtype == torch.float32 and t.device.type == "cpu"
    return t.contiguous()


def _i(t):
    assert t.dtype == torch.int32 and t.device.type == "cpu"
    return t.contiguous()


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


# ------------------------------------------------------------------ Chamfer
def chamfer_fwd(xyz1, xyz2):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    d1, d2 = torch.empty(B, N), torch.empty(B, M)
    i1, i2 = torch.empty(B, N, dtype=torch.int32), torch.empty(B, M, dtype=torch.int32)
    lib().orc_chamfer_fwd(_p(xyz1), _p(xyz2), B, N, M, _p(d1), _p(d2), _p(i1), _p(i2))
    return d1, d2, i1, i2


def chamfer_bwd(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, xyz2, g1, g2 = _f(xyz1), _f(xyz2), _f(g1), _f(g2)
    idx1, idx2 = _i(idx1), _i(idx2)
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    gx1, gx2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    lib().orc_chamfer_bwd(_p(xyz1), _p(xyz2), B, N, M, _p(idx1), _p(idx2), _p(g1), _p(g2), _p(gx1), _p(gx2))
    return gx1, gx2


# ------------------------------------------------------------------ EMD
def emd_fwd(xyz1, xyz2, eps, iters, return_evals=False):
    xyz1, xyz2 = _f(xyz1), _f(xyz2)
    B, N, _ = xyz1.shape
    assert xyz2.shape == xyz1.shape
    dist = torch.empty(B, N)
    ass = torch.empty(B, N, dtype=torch.int32)
    ev = torch.zeros(B, dtype=torch.float64)
    rc = lib().orc_emd_fwd(_p(xyz1), _p(xyz2), B, N, ctypes.c_float(eps), int(iters), _p(dist), _p(ass), _p(ev))
    if rc != 0:
        raise ValueError("emd: bad input (n % 1024 != 0 or B > 512)")
    return (dist, ass, ev) if return_evals else (dist, ass)


def emd_bwd(xyz1, xyz2, gdist, assignment):
    xyz1, xyz2, gdist, assignment = _f(xyz1), _f(xyz2), _f(gdist), _i(assignment)
    B, N, _ = xyz1.shape
    g = torch.empty_like(xyz1)
    lib().orc_emd_bwd(_p(xyz1), _p(xyz2), B, N, _p(gdist), _p(assignment), _p(g))
    return g


# ------------------------------------------------------------------ expansion penalty
def expansion_fwd(xyz, primitive_size, alpha):
    xyz = _f(xyz)
    B, N, _ = xyz.shape
    dist = torch.empty(B, N)
    idx = torch.empty(B, N, dtype=torch.int32)
    mml = torch.empty(B)
    rc = lib().orc_expansion_fwd(_p(xyz), B, N, int(primitive_size), ctypes.c_float(alpha), _p(dist), _p(idx), _p(mml))
    if rc != 0:
        raise ValueError("expansion: primitive_size must be a power of two <= 512 dividing n")
    return dist, idx, mml


def expansion_bwd(xyz, gdist, idx):
    xyz, gdist, idx = _f(xyz), _f(gdist), _i(idx)
    B, N, _ = xyz.shape
    g = torch.empty_like(xyz)
    lib().orc_expansion_bwd(_p(xyz), B, N, _p(gdist), _p(idx), _p(g))
    return g


# ------------------------------------------------------------------ MDS + gather
def mds(xyz, npoint, mean_mst_length):
    xyz, mml = _f(xyz), _f(mean_mst_length)
    B, n, _ = xyz.shape
    out = torch.zeros(B, npoint, dtype=torch.int32)
    lib().orc_mds(_p(xyz), B, n, int(npoint), _p(mml), _p(out))
    return out


def mds_check(xyz_one, mml_one, idx_one, rel_tol=1e-5):
    """Replay one sample's sampled sequence; returns (violations, steps where host-expf argmin differs)."""
    xyz_one, idx_one = _f(xyz_one), _i(idx_one)
    mism = ctypes.c_int(0)
    bad = lib().orc_mds_check(_p(xyz_one), xyz_one.shape[0], idx_one.shape[0], ctypes.c_float(float(mml_one)),
                              _p(idx_one), ctypes.c_double(rel_tol), ctypes.byref(mism))
    return bad, mism.value


def gather_fwd(features, idx):
    features, idx = _f(features), _i(idx)
    B, C, n = features.shape
    m = idx.shape[1]
    out = torch.empty(B, C, m)
    lib().orc_gather_fwd(_p(features), _p(idx), B, C, n, m, _p(out))
    return out


def gather_bwd(gout, idx, n):
    gout, idx = _f(gout), _i(idx)
    B, C, m = gout.shape
    g = torch.empty(B, C, n)
    lib().orc_gather_bwd(_p(gout), _p(idx), B, C, n, m, _p(g))
    return g


# ------------------------------------------------------------------ p2i (pixel-space points)
def _suf(t):
    return {torch.float32: ("f32", ctypes.c_float), torch.float64: ("f64", ctypes.c_double)}[t.dtype]


def p2i_max_fwd(points, feat, batch_inds, background, radius):
    suf, cty = _suf(points)
    points, feat, background = points.contiguous(), feat.contiguous(), background.contiguous()
    binds = _i(batch_inds)
    B, C, H, W = background.shape
    out = torch.empty_like(background)
    ids = torch.empty(B, C, H, W, dtype=torch.int32)
    getattr(lib(), "orc_p2i_max_fwd_" + suf)(_p(points), _p(feat), _p(binds), _p(background), points.shape[0],
                                             B, C, H, W, cty(radius), _p(out), _p(ids))
    return out, ids


def p2i_max_bwd(gout, ids, points, feat, radius):
    suf, cty = _suf(points)
    gout, points, feat, ids = gout.contiguous(), points.contiguous(), feat.contiguous(), _i(ids)
    B, C, H, W = gout.shape
    gp, gf, gb = torch.empty_like(points), torch.empty_like(feat), torch.empty_like(gout)
    getattr(lib(), "orc_p2i_max_bwd_" + suf)(_p(gout), _p(ids), _p(points), _p(feat), points.shape[0],
                                             B, C, H, W, cty(radius), _p(gp), _p(gf), _p(gb))
    return gp, gf, gb


def p2i_sum_fwd(points, feat, batch_inds, background, radius):
    suf, cty = _suf(points)
    points, feat, background = points.contiguous(), feat.contiguous(), background.contiguous()
    binds = _i(batch_inds)
    B, C, H, W = background.shape
    out = torch.empty_like(background)
    getattr(lib(), "orc_p2i_sum_fwd_" + suf)(_p(points), _p(feat), _p(binds), _p(background), points.shape[0],
                                             B, C, H, W, cty(radius), _p(out))
    return out


def p2i_sum_bwd(gout, points, feat, batch_inds, radius):
    suf, cty = _suf(points)
    gout, points, feat, binds = gout.contiguous(), points.contiguous(), feat.contiguous(), _i(batch_inds)
    B, C, H, W = gout.shape
    gp, gf = torch.empty_like(points), torch.empty_like(feat)
    getattr(lib(), "orc_p2i_sum_bwd_" + suf)(_p(gout), _p(points), _p(feat), _p(binds), points.shape[0],
                                             B, C, H, W, cty(radius), _p(gp), _p(gf))
    return gp, gf


# ------------------------------------------------------------------ kNN
def knn(x, k, return_dist=False):
    x = _f(x)
    B, C, N = x.shape
    idx = torch.empty(B, N, k, dtype=torch.int32)
    dist = torch.empty(B, N, k)
    lib().orc_knn(_p(x), B, C, N, int(k), _p(idx), _p(dist))
    return (idx, dist) if return_dist else idx


# ------------------------------------------------------------------ gridding (GRNet)
def gridding_fwd(pts, bounds):
    """pts [B,n,3] already scaled; bounds = (minx, maxx, miny, maxy, minz, maxz) -> grid [B,V], weights [B,n,8,3], indexes [B,n,8]"""
    pts = _f(pts)
    B, n, _ = pts.shape
    lens = [int(bounds[2 * i + 1] - bounds[2 * i] + 1) for i in range(3)]
    V = lens[0] * lens[1] * lens[2]
    grid, w, ix = torch.empty(B, V), torch.empty(B, n, 8, 3), torch.empty(B, n, 8, dtype=torch.int32)
    lib().orc_gridding_fwd(_p(pts), B, n, *[ctypes.c_float(float(v)) for v in bounds], _p(grid), _p(w), _p(ix))
    return grid, w, ix


def gridding_bwd(weights, indexes, ggrid):
    weights, indexes, ggrid = _f(weights), _i(indexes), _f(ggrid)
    B, n = indexes.shape[:2]
    g = torch.empty(B, n, 3)
    lib().orc_gridding_bwd(_p(weights), _p(indexes), _p(ggrid), B, n, ctypes.c_size_t(ggrid.shape[1]), _p(g))
    return g


def gridding_rev_fwd(grid, scale):
    grid = _f(grid).reshape(grid.shape[0], -1)
    B = grid.shape[0]
    pts = torch.empty(B, scale ** 3, 3)
    lib().orc_gridding_rev_fwd(_p(grid), B, int(scale), _p(pts))
    return pts


def gridding_rev_bwd(pts, grid, gpts, scale):
    pts, gpts = _f(pts), _f(gpts)
    grid = _f(grid).reshape(grid.shape[0], -1)
    B = grid.shape[0]
    g = torch.empty(B, scale ** 3)
    lib().orc_gridding_rev_bwd(_p(pts), _p(grid), _p(gpts), B, int(scale), _p(g))
    return g


# ------------------------------------------------------------------ gridding loss grid / cubic feature sampling (GRNet)
def gridding_dist_fwd(pts, bounds):
    pts = _f(pts)
    B, n, _ = pts.shape
    lens = [int(bounds[2 * i + 1] - bounds[2 * i] + 1) for i in range(3)]
    V = lens[0] * lens[1] * lens[2]
    grid, w, ix = torch.empty(B, V, 8), torch.empty(B, n, 8, 3), torch.empty(B, n, 8, dtype=torch.int32)
    lib().orc_gridding_dist_fwd(_p(pts), B, n, *[ctypes.c_float(float(v)) for v in bounds], _p(grid), _p(w), _p(ix))
    return grid, w, ix


def gridding_dist_bwd(weights, indexes, ggrid):
    """Same routing as gridding_bwd; the gradient grid is [B, V, 8]."""
    return gridding_bwd(weights, indexes, _f(ggrid).reshape(ggrid.shape[0], -1))


def cubic_sampling_fwd(pts, feat, ns):
    pts, feat = _f(pts), _f(feat)
    B, n, _ = pts.shape
    C, S = feat.shape[1], feat.shape[2]
    V = (2 * ns) ** 3
    out, ix = torch.empty(B, n, V, C), torch.empty(B, n, V, dtype=torch.int32)
    lib().orc_cubic_sampling_fwd(_p(pts), _p(feat), B, n, C, S, int(ns), _p(out), _p(ix))
    return out, ix


def cubic_sampling_bwd(gout, indexes, S, ns):
    gout, indexes = _f(gout), _i(indexes)
    B, n, V, C = gout.shape
    g = torch.empty(B, C, S, S, S)
    lib().orc_cubic_sampling_bwd(_p(gout), _p(indexes), B, n, C, int(S), int(ns), _p(g))
    return g
