"""Raw (non-autograd) host wrappers: torch CUDA tensors in, C-ABI call, torch CUDA tensors out.

PyTorch is plumbing here: it owns device memory (outputs and scratch come from its caching allocator) and
the current stream.  All arithmetic happens in libsparenet_b200.so; there is no fallback path.
"""
import os

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _cuda_f32(t, name):
    if not t.is_cuda:
        raise SnbValueError(f"{name} must be a CUDA tensor (sparenet_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise SnbValueError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _cuda_i32(t, name):
    if not t.is_cuda or t.dtype != torch.int32:
        raise SnbValueError(f"{name} must be a CUDA int32 tensor")
    return t.contiguous()


class SnbValueError(ValueError):
    pass


# ---- instrumentation: kernel-launch counter (bench.py's gpu_launches) and optional per-op CUDA-event timing --------
LAUNCHES = {"count": 0}
PROFILE = None  # set to a dict {op: [(start_event, end_event), ...]} by bench.py to time ops on the current stream
FLOPS = {}      # op -> floating-point operations issued while PROFILE is set (tensor-core GEMMs: 2*M*N*K per product)


class _op:
    """Counts the kernels a C-ABI call launches and, when PROFILE is set, brackets it with CUDA events."""
    def __init__(self, name, launches, nbytes=None, flops=None):
        self.name, self.launches, self.nbytes = name, launches, nbytes   # nbytes: algorithmic HBM bytes of the call (roofline)
        self.flops = flops

    def __enter__(self):
        LAUNCHES["count"] += self.launches
        if PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            b = torch.cuda.Event(enable_timing=True)
            b.record()
            PROFILE.setdefault(self.name, []).append((self.a, b, self.nbytes))
            if self.flops:
                FLOPS[self.name] = FLOPS.get(self.name, 0) + self.flops
        return False


def _ws(nbytes, device):
    # 16-byte aligned scratch from the caching allocator (allocations are 512-byte aligned)
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


# ----------------------------------------------------------------------------- Chamfer
_CHAMFER_MEMO = {}
_CHAMFER_REUSE = {"depth": 0}
CHAMFER_PRUNED = True   # False forces the brute-force kernel (tests compare the two bit for bit)


class chamfer_reuse:
    """Opt-in de-duplication of IDENTICAL Chamfer searches inside the block: the training loss evaluates Chamfer(refine, gt) twice in a
    row (ChamferDistanceMean, then the consistency term, runners/sparenet_runner.py:87,103).  Within the scope a second
    chamfer_forward on the very same, unmodified tensors (data pointer, version counter, shape, stream) returns the first search's
    result instead of repeating the N x M search -- same values, one search less.  Off by default: outside a scope every call
    searches.  The memo is dropped when the outermost scope exits, so it never pins tensors beyond the loss computation."""
    def __enter__(self):
        _CHAMFER_REUSE["depth"] += 1
        return self

    def __exit__(self, *exc):
        _CHAMFER_REUSE["depth"] -= 1
        if _CHAMFER_REUSE["depth"] == 0:
            _CHAMFER_MEMO.clear()
        return False


def chamfer_forward(xyz1, xyz2):
    xyz1, xyz2 = _cuda_f32(xyz1, "xyz1"), _cuda_f32(xyz2, "xyz2")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.size(2) != 3 or xyz2.size(2) != 3 or xyz1.size(0) != xyz2.size(0):
        raise SnbValueError("chamfer: expected [B,N,3] and [B,M,3]")
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    dev = xyz1.device
    # Inside a chamfer_reuse() scope a one-entry memo on the very same, unmodified tensors skips a repeated N x M search.  The entry
    # keeps references to its inputs (their storage cannot be recycled while it is alive) and remembers the outputs' version
    # counters: an in-place edit of a returned tensor invalidates it.
    c = _CHAMFER_MEMO
    reuse = _CHAMFER_REUSE["depth"] > 0
    key = (xyz1.data_ptr(), xyz1._version, tuple(xyz1.shape), xyz2.data_ptr(), xyz2._version, tuple(xyz2.shape),
           torch.cuda.current_stream(dev).cuda_stream)
    if reuse and c.get("key") == key and all(t._version == v for t, v in zip(c["out"], c["out_versions"])):
        return tuple(t.detach() for t in c["out"])     # fresh aliases: each autograd node owns its output objects
    d1 = torch.empty(B, N, device=dev)
    d2 = torch.empty(B, M, device=dev)
    i1 = torch.empty(B, N, dtype=torch.int32, device=dev)
    i2 = torch.empty(B, M, dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = lib.snb_chamfer_workspace_bytes(B, N, M) if CHAMFER_PRUNED else 0
    ws = _ws(nbytes, dev) if nbytes else None
    with torch.cuda.device(dev), _op("chamfer_fwd", 2 if nbytes else 1):
        check(lib.snb_chamfer_fwd(ptr(xyz1), ptr(xyz2), B, N, M, ptr(d1), ptr(d2), ptr(i1), ptr(i2), ptr(ws), nbytes, stream_ptr()), "chamfer_fwd")
    c.clear()
    if reuse:
        # detached aliases share storage and version counter but not the autograd graph: the inputs carry their history, and the
        # outputs get the calling autograd.Function's grad_fn attached in place once it returns -- holding either would keep the
        # previous step's whole graph (and its AccumulateGrad nodes) alive
        out = (d1.detach(), d2.detach(), i1, i2)
        c.update(key=key, keep=(xyz1.detach(), xyz2.detach()), out=out, out_versions=tuple(t._version for t in out))
    return d1, d2, i1, i2


def chamfer_backward(xyz1, xyz2, idx1, idx2, g1, g2):
    xyz1, xyz2 = _cuda_f32(xyz1, "xyz1"), _cuda_f32(xyz2, "xyz2")
    g1, g2 = _cuda_f32(g1, "grad_dist1"), _cuda_f32(g2, "grad_dist2")
    idx1, idx2 = _cuda_i32(idx1, "idx1"), _cuda_i32(idx2, "idx2")
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    gx1, gx2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    with torch.cuda.device(xyz1.device), _op("chamfer_bwd", 2):
        check(_lib.load().snb_chamfer_bwd(ptr(xyz1), ptr(xyz2), B, N, M, ptr(idx1), ptr(idx2), ptr(g1), ptr(g2), ptr(gx1), ptr(gx2),
                                          stream_ptr()), "chamfer_bwd")
    return gx1, gx2


# ----------------------------------------------------------------------------- EMD
def emd_forward(xyz1, xyz2, eps, iters, exhaustive=False):
    """exhaustive=True: the Bid pass that evaluates every (bidder, object) pair (snb_emd_fwd_scan) instead of the box-pruned one;
    identical results, kept for clouds above 16384 points and as the A/B check of the pruning."""
    xyz1, xyz2 = _cuda_f32(xyz1, "xyz1"), _cuda_f32(xyz2, "xyz2")
    B, N, _ = xyz1.shape
    dev = xyz1.device
    dist = torch.empty(B, N, device=dev)
    ass = torch.empty(B, N, dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = lib.snb_emd_workspace_bytes(B, N)
    ws = _ws(nbytes, dev)
    with torch.cuda.device(dev), _op("emd_fwd", 1):
        fwd = lib.snb_emd_fwd_scan if exhaustive else lib.snb_emd_fwd
        check(fwd(ptr(xyz1), ptr(xyz2), B, N, float(eps), int(iters), ptr(dist), ptr(ass), ptr(ws), nbytes, stream_ptr()), "emd_fwd")
    return dist, ass


def emd_backward(xyz1, xyz2, grad_dist, assignment):
    xyz1, xyz2, grad_dist = _cuda_f32(xyz1, "xyz1"), _cuda_f32(xyz2, "xyz2"), _cuda_f32(grad_dist, "grad_dist")
    assignment = _cuda_i32(assignment, "assignment")
    B, N, _ = xyz1.shape
    g = torch.empty_like(xyz1)
    with torch.cuda.device(xyz1.device), _op("emd_bwd", 1):
        check(_lib.load().snb_emd_bwd(ptr(xyz1), ptr(xyz2), B, N, ptr(grad_dist), ptr(assignment), ptr(g), stream_ptr()), "emd_bwd")
    return g


# ----------------------------------------------------------------------------- expansion penalty
def expansion_forward(xyz, primitive_size, alpha):
    xyz = _cuda_f32(xyz, "xyz")
    B, N, _ = xyz.shape
    dev = xyz.device
    dist = torch.empty(B, N, device=dev)
    ass = torch.empty(B, N, dtype=torch.int32, device=dev)
    mml = torch.empty(B, device=dev)
    lib = _lib.load()
    nbytes = lib.snb_expansion_workspace_bytes(B, N, int(primitive_size))
    ws = _ws(nbytes, dev)
    with torch.cuda.device(dev), _op("expansion_fwd", 2):
        check(lib.snb_expansion_fwd(ptr(xyz), B, N, int(primitive_size), float(alpha), ptr(dist), ptr(ass), ptr(mml), ptr(ws), nbytes,
                                    stream_ptr()), "expansion_fwd")
    return dist, ass, mml


def expansion_backward(xyz, grad_dist, assignment):
    xyz, grad_dist, assignment = _cuda_f32(xyz, "xyz"), _cuda_f32(grad_dist, "grad_dist"), _cuda_i32(assignment, "assignment")
    B, N, _ = xyz.shape
    g = torch.empty_like(xyz)
    with torch.cuda.device(xyz.device), _op("expansion_bwd", 1):
        check(_lib.load().snb_expansion_bwd(ptr(xyz), B, N, ptr(grad_dist), ptr(assignment), ptr(g), stream_ptr()), "expansion_bwd")
    return g


# ----------------------------------------------------------------------------- MDS + gather
def mds_sample(xyz, npoint, mean_mst_length):
    xyz, mml = _cuda_f32(xyz, "xyz"), _cuda_f32(mean_mst_length, "mean_mst_length")
    B, n, _ = xyz.shape
    idx = torch.empty(B, int(npoint), dtype=torch.int32, device=xyz.device)
    lib = _lib.load()
    nbytes = lib.snb_mds_workspace_bytes(B, n, int(npoint))
    ws = _ws(nbytes, xyz.device)
    with torch.cuda.device(xyz.device), _op("mds_sample", 1):
        check(lib.snb_mds_sample(ptr(xyz), B, n, int(npoint), ptr(mml), ptr(idx), ptr(ws), nbytes, stream_ptr()), "mds_sample")
    return idx


def gather_forward(features, idx):
    features, idx = _cuda_f32(features, "features"), _cuda_i32(idx, "idx")
    B, C, n = features.shape
    m = idx.shape[1]
    out = torch.empty(B, C, m, device=features.device)
    with torch.cuda.device(features.device), _op("gather_fwd", 1):
        check(_lib.load().snb_gather_fwd(ptr(features), ptr(idx), B, C, n, m, ptr(out), stream_ptr()), "gather_fwd")
    return out


def gather_backward(grad_out, idx, n):
    grad_out, idx = _cuda_f32(grad_out, "grad_out"), _cuda_i32(idx, "idx")
    B, C, m = grad_out.shape
    g = torch.empty(B, C, int(n), device=grad_out.device)
    with torch.cuda.device(grad_out.device), _op("gather_bwd", 1):
        check(_lib.load().snb_gather_bwd(ptr(grad_out), ptr(idx), B, C, int(n), m, ptr(g), stream_ptr()), "gather_bwd")
    return g


# ----------------------------------------------------------------------------- p2i
def _p2i_dtype(points):
    if not points.is_cuda:
        raise SnbValueError("p2i: tensors must live on a CUDA device")
    if points.dtype not in (torch.float32, torch.float64):
        raise SnbValueError(f"p2i: float32 or float64 expected, got {points.dtype}")
    return 1 if points.dtype == torch.float64 else 0


def depthmaps_forward(data, view_matrix, H, W, radius):
    """data [B,N,3] float32 CUDA, view_matrix: 16 python floats (row-major 4x4) -> (out [B,1,H,W], ids [B,1,H,W] int32, workspace).
    The workspace must be handed back to depthmaps_backward (it keeps the per-point pixel coordinates and the depth range)."""
    import ctypes
    data = _cuda_f32(data, "data")
    B, N = data.shape[0], data.shape[1]
    dev = data.device
    out = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
    ids = torch.empty(B, 1, H, W, dtype=torch.int32, device=dev)
    lib = _lib.load()
    nbytes = lib.snb_depthmaps_workspace_bytes(B, N, H, W)
    ws = _ws(nbytes, dev)
    vm = (ctypes.c_float * 16)(*[float(v) for v in view_matrix])
    with torch.cuda.device(dev), _op("depthmaps_fwd", 4):
        check(lib.snb_depthmaps_fwd(ptr(data), B, N, vm, H, W, float(radius), ptr(out), ptr(ids), ptr(ws), nbytes, stream_ptr()), "depthmaps_fwd")
    return out, ids, ws


def depthmaps_backward(grad_out, ids, data, view_matrix, radius, ws):
    import ctypes
    data = _cuda_f32(data, "data")
    grad_out = _cuda_f32(grad_out, "grad_out")
    B, N = data.shape[0], data.shape[1]
    H, W = grad_out.shape[-2], grad_out.shape[-1]
    gdata = torch.empty_like(data)
    vm = (ctypes.c_float * 16)(*[float(v) for v in view_matrix])
    with torch.cuda.device(data.device), _op("depthmaps_bwd", 3):
        check(_lib.load().snb_depthmaps_bwd(ptr(grad_out), ptr(ids), ptr(data), B, N, vm, H, W, float(radius), ptr(ws), ws.numel(),
                                            ptr(gdata), stream_ptr()), "depthmaps_bwd")
    return gdata


def p2i_max_forward(points, feat, batch_inds, background, kernel_kind, radius):
    dbl = _p2i_dtype(points)
    points, feat, background = points.contiguous(), feat.to(points.dtype).contiguous(), background.to(points.dtype).contiguous()
    binds = _cuda_i32(batch_inds, "batch_inds")
    B, C, H, W = background.shape
    out = torch.empty_like(background)
    ids = torch.empty(B, C, H, W, dtype=torch.int32, device=points.device)
    lib = _lib.load()
    nbytes = lib.snb_p2i_workspace_bytes(B, C, H, W, dbl)
    ws = _ws(nbytes, points.device)
    with torch.cuda.device(points.device), _op("p2i_max_fwd", 3):
        check(lib.snb_p2i_max_fwd(ptr(points), ptr(feat), ptr(binds), ptr(background), points.shape[0], B, C, H, W, int(kernel_kind),
                                  float(radius), dbl, ptr(out), ptr(ids), ptr(ws), nbytes, stream_ptr()), "p2i_max_fwd")
    return out, ids


def p2i_max_backward(grad_out, ids, points, feat, kernel_kind, radius):
    dbl = _p2i_dtype(points)
    grad_out, points, feat = grad_out.to(points.dtype).contiguous(), points.contiguous(), feat.to(points.dtype).contiguous()
    ids = _cuda_i32(ids, "out_point_ids")
    B, C, H, W = grad_out.shape
    gp, gf, gb = torch.empty_like(points), torch.empty_like(feat), torch.empty_like(grad_out)
    with torch.cuda.device(points.device), _op("p2i_max_bwd", 1):
        check(_lib.load().snb_p2i_max_bwd(ptr(grad_out), ptr(ids), ptr(points), ptr(feat), points.shape[0], B, C, H, W, int(kernel_kind),
                                          float(radius), dbl, ptr(gp), ptr(gf), ptr(gb), stream_ptr()), "p2i_max_bwd")
    return gp, gf, gb


def p2i_sum_forward(points, feat, batch_inds, background, kernel_kind, radius):
    dbl = _p2i_dtype(points)
    points, feat, background = points.contiguous(), feat.to(points.dtype).contiguous(), background.to(points.dtype).contiguous()
    binds = _cuda_i32(batch_inds, "batch_inds")
    B, C, H, W = background.shape
    out = torch.empty_like(background)
    with torch.cuda.device(points.device), _op("p2i_sum_fwd", 1):
        check(_lib.load().snb_p2i_sum_fwd(ptr(points), ptr(feat), ptr(binds), ptr(background), points.shape[0], B, C, H, W, int(kernel_kind),
                                          float(radius), dbl, ptr(out), stream_ptr()), "p2i_sum_fwd")
    return out


def p2i_sum_backward(grad_out, points, feat, batch_inds, kernel_kind, radius):
    dbl = _p2i_dtype(points)
    grad_out, points, feat = grad_out.to(points.dtype).contiguous(), points.contiguous(), feat.to(points.dtype).contiguous()
    binds = _cuda_i32(batch_inds, "batch_inds")
    B, C, H, W = grad_out.shape
    gp, gf = torch.empty_like(points), torch.empty_like(feat)
    with torch.cuda.device(points.device), _op("p2i_sum_bwd", 1):
        check(_lib.load().snb_p2i_sum_bwd(ptr(grad_out), ptr(points), ptr(feat), ptr(binds), points.shape[0], B, C, H, W, int(kernel_kind),
                                          float(radius), dbl, ptr(gp), ptr(gf), stream_ptr()), "p2i_sum_bwd")
    return gp, gf


# ----------------------------------------------------------------------------- kNN
def knn_indices_pruned(x, k):
    """Same result as knn_indices for wide features: the TF32 Gram matrix X^T X from the tcgen05 GEMM (snb_gemm_tf32: A = the
    point-major copy, B = the channel-major features) prunes, snb_knn_pruned re-evaluates the surviving candidates exactly
    (csrc/knn_prune.cu).  Indices identical to knn_indices (tests/test_gpu_ops.py::test_knn_pruned_identical_to_brute_force)."""
    from . import gemm
    x = _cuda_f32(x, "x").detach()                       # indices carry no gradient: keep the GEMM out of the autograd graph
    B, C, N = x.shape
    idx = torch.empty(B, N, int(k), dtype=torch.int32, device=x.device)
    lib = _lib.load()
    nbytes = lib.snb_knn_pruned_workspace_bytes(B, N)
    ws = _ws(nbytes, x.device)
    with torch.cuda.device(x.device), _op("knn", 3):     # transpose + norms + prune; the op's time includes the Gram GEMM (counted by its own op)
        xT = torch.empty(B, N, C, device=x.device, dtype=torch.float32)
        check(lib.snb_transpose_cn(ptr(x), B, C, N, ptr(xT), stream_ptr()), "transpose_cn")
        gram, _ = gemm.conv_fwd(x, xT)                   # [B,N,N]: one "weight" per sample = its own points
        check(lib.snb_knn_pruned(ptr(xT), ptr(gram), B, C, N, int(k), ptr(idx), ptr(ws), nbytes, stream_ptr()), "knn_pruned")
    return idx


_KNN_PRUNE_CHECKED = {}   # (C, N, k) -> bool: the pruned path reproduced the brute-force indices on its first use in this process


def _knn_brute(x, k):
    B, C, N = x.shape
    idx = torch.empty(B, N, int(k), dtype=torch.int32, device=x.device)
    lib = _lib.load()
    nbytes = lib.snb_knn_workspace_bytes(B, N)
    ws = _ws(nbytes, x.device)
    with torch.cuda.device(x.device), _op("knn", 2):
        check(lib.snb_knn(ptr(x), B, C, N, int(k), ptr(idx), ptr(ws), nbytes, stream_ptr()), "knn")
    return idx


def knn_indices(x, k):
    """x: [B, C, N] float32 (channel-major, as the encoder holds it) -> idx [B, N, k] int32."""
    x = _cuda_f32(x, "x")
    B, C, N = x.shape
    # Wide features: the TF32 Gram matrix (our tensor-core GEMM) prunes, exact fp32 re-evaluation decides (identical indices;
    # parity-tested on B200).  On by default from C >= 256, where the brute-force kernel is FP32-issue bound;
    # SNB_KNN_PRUNE=1 / 0 forces it on (for C >= 64) / off.  Shapes the GEMM does not tile (N % 32, C % 4) take the brute-force kernels.
    force = os.environ.get("SNB_KNN_PRUNE")
    if (C & 3) == 0 and (N & 31) == 0 and force != "0" and ((force == "1" and C >= 64) or C >= 256):
        key = (C, N, int(k), x.device.index)
        ok = _KNN_PRUNE_CHECKED.get(key)
        if ok is None and not torch.cuda.is_current_stream_capturing():
            # First use of a shape on a device in this process: both paths once, compared on the device; should the indices ever
            # differ the brute-force kernels keep serving that shape -- loudly.
            a, b = knn_indices_pruned(x, k), _knn_brute(x, k)
            ok = bool(torch.equal(a, b))
            _KNN_PRUNE_CHECKED[key] = ok
            if not ok:
                import warnings
                warnings.warn(f"sparenet_b200: Gram-pruned kNN disagreed with the brute-force kernels for C={C}, N={N}, k={k}; "
                              "using the brute-force kernels for this shape", RuntimeWarning)
            return b
        if ok is not False:
            return knn_indices_pruned(x, k)
    return _knn_brute(x, k)


# ----------------------------------------------------------------------------- gridding (GRNet)
def gridding_forward(ptcloud, bounds):
    """ptcloud [B,n,3] already scaled, bounds (min_x, max_x, min_y, max_y, min_z, max_z) -> grid [B,V], weights [B,n,8,3], idx [B,n,8]."""
    ptcloud = _cuda_f32(ptcloud, "ptcloud")
    B, n, _ = ptcloud.shape
    lens = [int(bounds[2 * i + 1] - bounds[2 * i] + 1) for i in range(3)]
    V = lens[0] * lens[1] * lens[2]
    dev = ptcloud.device
    grid = torch.empty(B, V, device=dev)
    w = torch.empty(B, n, 8, 3, device=dev)
    ix = torch.empty(B, n, 8, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _op("gridding_fwd", 1):
        check(_lib.load().snb_gridding_fwd(ptr(ptcloud), B, n, *[float(v) for v in bounds], ptr(grid), ptr(w), ptr(ix), stream_ptr()), "gridding_fwd")
    return grid, w, ix


def gridding_backward(weights, indexes, grad_grid):
    weights, grad_grid, indexes = _cuda_f32(weights, "grid_pt_weights"), _cuda_f32(grad_grid, "grad_grid"), _cuda_i32(indexes, "grid_pt_indexes")
    B, n = indexes.shape[:2]
    g = torch.empty(B, n, 3, device=weights.device)
    with torch.cuda.device(weights.device), _op("gridding_bwd", 1):
        check(_lib.load().snb_gridding_bwd(ptr(weights), ptr(indexes), ptr(grad_grid), B, n, grad_grid.shape[1], ptr(g), stream_ptr()), "gridding_bwd")
    return g


def gridding_reverse_forward(grid, scale):
    grid = _cuda_f32(grid, "grid")
    B = grid.shape[0]
    pts = torch.empty(B, int(scale) ** 3, 3, device=grid.device)
    with torch.cuda.device(grid.device), _op("gridding_rev_fwd", 1):
        check(_lib.load().snb_gridding_rev_fwd(ptr(grid), B, int(scale), ptr(pts), stream_ptr()), "gridding_rev_fwd")
    return pts


def gridding_reverse_backward(ptcloud, grid, grad_ptcloud, scale):
    ptcloud, grid, grad_ptcloud = _cuda_f32(ptcloud, "ptcloud"), _cuda_f32(grid, "grid"), _cuda_f32(grad_ptcloud, "grad_ptcloud")
    B = grid.shape[0]
    g = torch.empty(B, int(scale) ** 3, device=grid.device)
    with torch.cuda.device(grid.device), _op("gridding_rev_bwd", 1):
        check(_lib.load().snb_gridding_rev_bwd(ptr(ptcloud), ptr(grid), ptr(grad_ptcloud), B, int(scale), ptr(g), stream_ptr()), "gridding_rev_bwd")
    return g


# ----------------------------------------------------------------------------- gridding loss / cubic feature sampling (GRNet)
def gridding_dist_forward(ptcloud, bounds):
    """Like gridding_forward with eight accumulators per vertex (one per corner role): grid [B,V,8], weights [B,n,8,3], idx [B,n,8]."""
    ptcloud = _cuda_f32(ptcloud, "ptcloud")
    B, n, _ = ptcloud.shape
    lens = [int(bounds[2 * i + 1] - bounds[2 * i] + 1) for i in range(3)]
    V = lens[0] * lens[1] * lens[2]
    dev = ptcloud.device
    grid = torch.empty(B, V, 8, device=dev)
    w = torch.empty(B, n, 8, 3, device=dev)
    ix = torch.empty(B, n, 8, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _op("gridding_dist_fwd", 1):
        check(_lib.load().snb_gridding_dist_fwd(ptr(ptcloud), B, n, *[float(v) for v in bounds], ptr(grid), ptr(w), ptr(ix), stream_ptr()),
              "gridding_dist_fwd")
    return grid, w, ix


def gridding_dist_backward(weights, indexes, grad_grid):
    weights, grad_grid, indexes = _cuda_f32(weights, "grid_pt_weights"), _cuda_f32(grad_grid, "grad_grid"), _cuda_i32(indexes, "grid_pt_indexes")
    B, n = indexes.shape[:2]
    g = torch.empty(B, n, 3, device=weights.device)
    with torch.cuda.device(weights.device), _op("gridding_dist_bwd", 1):
        check(_lib.load().snb_gridding_dist_bwd(ptr(weights), ptr(indexes), ptr(grad_grid), B, n, grad_grid.shape[1], ptr(g), stream_ptr()),
              "gridding_dist_bwd")
    return g


def cubic_sampling_forward(ptcloud, cubic_features, neighborhood_size):
    """ptcloud [B,n,3] in grid units, cubic_features [B,C,S,S,S] -> point_features [B,n,(2 ns)^3,C], grid_pt_indexes [B,n,(2 ns)^3]."""
    ptcloud, cubic_features = _cuda_f32(ptcloud, "ptcloud"), _cuda_f32(cubic_features, "cubic_features")
    B, n, _ = ptcloud.shape
    C, S = cubic_features.shape[1], cubic_features.shape[2]
    V = (2 * int(neighborhood_size)) ** 3
    dev = ptcloud.device
    out = torch.empty(B, n, V, C, device=dev)
    ix = torch.empty(B, n, V, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _op("cubic_sampling_fwd", 2):
        check(_lib.load().snb_cubic_sampling_fwd(ptr(ptcloud), ptr(cubic_features), B, n, C, S, int(neighborhood_size), ptr(out), ptr(ix), stream_ptr()),
              "cubic_sampling_fwd")
    return out, ix


def cubic_sampling_backward(grad_point_features, indexes, scale, neighborhood_size):
    g = _cuda_f32(grad_point_features, "grad_point_features")
    indexes = _cuda_i32(indexes, "grid_pt_indexes")
    B, n, V, C = g.shape
    S = int(scale)
    gf = torch.empty(B, C, S, S, S, device=g.device)
    with torch.cuda.device(g.device), _op("cubic_sampling_bwd", 1):
        check(_lib.load().snb_cubic_sampling_bwd(ptr(g), ptr(indexes), B, n, C, S, int(neighborhood_size), ptr(gf), stream_ptr()), "cubic_sampling_bwd")
    return gf
