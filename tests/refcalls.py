"""Restated calling protocols of the reference's pybind extensions (oracle/_ref/*.so), i.e. what the
reference's own Python wrappers allocate and pass (cited per function).  Test infrastructure only."""
import torch


def chamfer_fwd(ext, x, y):  # cuda/chamfer_dist/__init__.py:8-12
    d1, d2, i1, i2 = ext.forward(x, y)
    return d1, d2, i1, i2


def chamfer_bwd(ext, x, y, i1, i2, g1, g2):  # cuda/chamfer_dist/__init__.py:15-18
    return ext.backward(x, y, i1, i2, g1, g2)


def emd_fwd(ext, x1, x2, eps, iters):  # cuda/emd/emd_module.py:31-76
    B, n, _ = x1.shape
    dev = x1.device

    def z(*s, dt=torch.float32):
        return torch.zeros(*s, device=dev, dtype=dt)

    dist = z(B, n)
    assignment = z(B, n, dt=torch.int32) - 1
    assignment_inv = z(B, n, dt=torch.int32) - 1
    price, bid, bid_inc, max_inc = z(B, n), z(B, n, dt=torch.int32), z(B, n), z(B, n)
    unass_idx, max_idx = z(B * n, dt=torch.int32), z(B * n, dt=torch.int32)
    unass_cnt, unass_cnt_sum, cnt_tmp = z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32)
    ext.forward(x1, x2, dist, assignment, price, assignment_inv, bid, bid_inc, max_inc, unass_idx, unass_cnt, unass_cnt_sum,
                cnt_tmp, max_idx, eps, iters)
    return dist, assignment


def emd_bwd(ext, x1, x2, gdist, assignment):  # emd_module.py:79-87
    g = torch.zeros_like(x1)
    ext.backward(x1, x2, g, gdist, assignment)
    return g


def expansion_fwd(ext, xyz, p, alpha):  # cuda/expansion_penalty/expansion_penalty_module.py:26-40
    B, n, _ = xyz.shape
    dev = xyz.device
    dist = torch.zeros(B, n, device=dev)
    assignment = torch.zeros(B, n, device=dev, dtype=torch.int32) - 1
    neighbor = torch.zeros(B, n * 512, device=dev, dtype=torch.int32)
    cost = torch.zeros(B, n * 512, device=dev)
    mml = torch.zeros(B, device=dev)
    ext.forward(xyz, p, assignment, dist, alpha, neighbor, cost, mml)
    return dist, assignment, mml / (n / p)


def expansion_bwd(ext, xyz, gdist, assignment):  # :43-48
    g = torch.zeros_like(xyz)
    ext.backward(xyz, g, gdist, assignment)
    return g


def mds(ext, xyz, m, mml):  # cuda/MDS/MDS_module.py:29-33
    idx = torch.zeros(xyz.shape[0], m, device=xyz.device, dtype=torch.int32)
    ext.minimum_density_sampling(xyz, m, mml, idx)
    return idx
