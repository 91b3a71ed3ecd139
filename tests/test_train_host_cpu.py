"""CPU: host-side plumbing of the native train step -- the flat parameter / gradient buffers of posendf_b200.PoseNDF that the
weight-gradient kernels, the fused optimizer and the data-parallel all-reduce work on (the arithmetic itself is CUDA only and
is checked on the GPU in tests/test_gpu_train.py)."""
import numpy as np
import pytest
import torch

from posendf_b200 import PoseNDF, synth, train

DIMS = [256, 512, 1024, 512, 256, 64]


def _opt(use_enc=True, act="lrelu"):
    return {"train": {"device": "cpu", "loss_type": "l1", "batch_size": 4},
            "model": {"StrEnc": {"use": use_enc, "act": act, "beta": 100},
                      "DFNet": {"in_dim": 126 if use_enc else 84, "dims": DIMS, "act": act, "beta": 100}}}


@pytest.mark.parametrize("use_enc", [True, False])
def test_flat_gradient_views_follow_the_reference_parameter_order(use_enc):
    net = PoseNDF(_opt(use_enc))
    flat = net.flat_grad()
    off = 0
    for (n, p), v, (name, shape) in zip(net.named_parameters(), net._grad_views, synth.param_shapes(126 if use_enc else 84, use_enc=use_enc)):
        assert n == name and tuple(p.shape) == tuple(shape) == tuple(v.shape)
        assert v.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()
    assert off == flat.numel() == (1365565 if use_enc else 1365565 - 3516 - 42 * 256)
    assert not net.grads_attached()
    net.attach_grads()
    assert net.grads_attached() and all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(net.parameters(), net._grad_views))
    net.zero_grad()                       # torch default: set_to_none
    assert not net.grads_attached()


def test_flatten_parameters_keeps_values_state_dict_and_reference_order():
    net = PoseNDF(_opt())
    params = synth.make_params(3)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    before = {k: v.clone() for k, v in net.state_dict().items()}
    flat = net.flatten_parameters_()
    assert flat.numel() == 1365565 and net.flatten_parameters_() is flat          # idempotent
    np.testing.assert_array_equal(flat.numpy(), synth.flatten_params(params))      # == the C ABI's pndf_set_weights layout
    off = 0
    for p in net.parameters():
        assert p.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()
    for k, v in net.state_dict().items():
        assert torch.equal(v, before[k])
    # load_state_dict copies in place: the views survive
    other = synth.make_params(4)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in other.items()})
    np.testing.assert_array_equal(flat.numpy(), synth.flatten_params(other))
    # a write through one parameter is a write into the flat vector
    with torch.no_grad():
        net.dfnet.lin6.bias.fill_(0.25)
    assert flat[-1].item() == 0.25


def test_export_column_map_matches_the_kernel_header():
    """train.py's column constants (softplus second-order chain) and csrc/pndf_kernel.cuh / pndf_capi.cu must agree"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "posendf_b200", "csrc", "pndf_kernel.cuh")).read()
    assert int(re.search(r"kDumpRows = (\d+)", hdr).group(1)) == train.DUMP_ROWS
    capi = open(os.path.join(root, "posendf_b200", "csrc", "pndf_capi.cu")).read()
    z = [int(v) for v in re.search(r"z_col\[7\] = \{([^}]*)\}", capi).group(1).split(",")]
    a = [int(v) for v in re.search(r"a_col\[6\] = \{([^}]*)\}", capi).group(1).split(",")]
    assert z == [r[0] for r in train.Z_ROWS] and a == [r[0] for r in train.A_ROWS]
    assert "dump_dev + 5376" in capi and train.G0_ROW == 5376
