"""drop-in for the reference's model/network/net_modules.py (parameter containers of the fused engine)."""
from posendf_b200.module import BoneMLP, DFNet, StructureEncoder  # noqa: F401
