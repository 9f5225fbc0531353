"""sparenet_b200 -- B200-native (sm_100a) implementation of SpareNet's per-batch point-cloud hot path.

Layout: csrc/ (CUDA kernels + C ABI, built into lib/libsparenet_b200.so), functional.py (raw host wrappers),
dropin/ (mirror of the reference's operator/module interface: cuda.chamfer_dist, cuda.chamfer_distance,
cuda.emd.emd_module, cuda.expansion_penalty.expansion_penalty_module, cuda.MDS.MDS_module, cuda.p2i_op,
utils.p2i_utils, models.sparenet_generator, knn_cuda).  Put `sparenet_b200.dropin_path()` at the front of
sys.path to import those modules under the reference's own names.
"""
import os

__version__ = "0.2.0"


def dropin_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
