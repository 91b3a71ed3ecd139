"""drop-in for the reference's model/posendf.py: same names, fused sm_100a kernel underneath."""
from posendf_b200.module import PoseNDF, gradient  # noqa: F401
