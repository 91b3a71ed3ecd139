"""drop-in for the names the hot path uses from the reference's model/network/net_utils.py."""
from posendf_b200.module import gradient  # noqa: F401
from posendf_b200.synth import PARENTS


def get_parent_mapping(model_type):
    if model_type != "smpl":
        print("Model hierarchy not defined.....")
        return None
    return list(PARENTS)
