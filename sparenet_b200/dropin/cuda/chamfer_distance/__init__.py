from .chamfer_distance import ChamferDistance, ChamferDistanceMean, ChamferDistanceFunction  # noqa: F401
