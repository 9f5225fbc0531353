"""INTEGRATION.md section B, as an executable file: rebinding the reference's cuda/chamfer_dist/__init__.py:8-18 to the C ABI.
tests/test_abi.py checks the argument list against include/sparenet_b200.h; tests/test_gpu_ops.py runs it on the GPU."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = ctypes.CDLL(os.path.join(_HERE, "..", "sparenet_b200", "lib", "libsparenet_b200.so"))
_P, _I = ctypes.c_void_p, ctypes.c_int
# int snb_chamfer_fwd(const float* xyz1, const float* xyz2, int B, int N, int M, float* dist1, float* dist2, int* idx1, int* idx2,
#                     void* workspace, size_t workspace_bytes, void* stream);           -- 12 arguments
_lib.snb_chamfer_fwd.argtypes = [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, ctypes.c_size_t, _P]
_lib.snb_chamfer_fwd.restype = _I
_lib.snb_chamfer_workspace_bytes.argtypes = [_I, _I, _I]
_lib.snb_chamfer_workspace_bytes.restype = ctypes.c_size_t
_lib.snb_strerror.argtypes = [_I]
_lib.snb_strerror.restype = ctypes.c_char_p


def chamfer_forward(xyz1, xyz2):                      # replaces chamfer.forward(xyz1, xyz2) -> [dist1, dist2, idx1, idx2]
    B, N, _ = xyz1.shape
    M = xyz2.shape[1]
    dev = xyz1.device
    d1, d2 = torch.empty(B, N, device=dev), torch.empty(B, M, device=dev)
    i1, i2 = torch.empty(B, N, dtype=torch.int32, device=dev), torch.empty(B, M, dtype=torch.int32, device=dev)
    nbytes = _lib.snb_chamfer_workspace_bytes(B, N, M)          # 0: the brute-force kernel needs no scratch (workspace may be NULL)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dev)
    rc = _lib.snb_chamfer_fwd(_P(xyz1.data_ptr()), _P(xyz2.data_ptr()), B, N, M, _P(d1.data_ptr()), _P(d2.data_ptr()),
                              _P(i1.data_ptr()), _P(i2.data_ptr()), _P(ws.data_ptr()) if nbytes else None, nbytes,
                              _P(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(_lib.snb_strerror(rc).decode())
    return d1, d2, i1, i2
