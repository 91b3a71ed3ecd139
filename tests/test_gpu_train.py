"""GPU: training step on the fused path (posendf_b200/train.py) against the reference's own autograd
(oracle/posendf_torch.py, pinned to the reference's golden gradients in tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

from autograd_crosscheck import train_forward_autograd
from conftest import case_cfg, load_golden
from oracle import posendf_torch as otorch
from posendf_b200 import synth

pytestmark = pytest.mark.gpu


def _opt(cfg, loss="l1"):
    return {"train": {"device": "cuda", "loss_type": loss, "batch_size": 4},
            "model": {"StrEnc": {"use": cfg["use_enc"], "act": cfg["enc_act"], "beta": cfg["enc_beta"]},
                      "DFNet": {"in_dim": 126 if cfg["use_enc"] else 84, "dims": [256, 512, 1024, 512, 256, 64],
                                "act": cfg["df_act"], "beta": cfg["df_beta"]}}}


def _batch(seed, B):
    tp = synth.make_poses(2000 + seed, B, kind="noisy", sigma=0.25)
    tm = synth.make_poses(3000 + seed, B, kind="randn")
    tgt = (synth.uniform01(4000 + seed, B) * 0.5).astype(np.float32)
    return tp, tgt, tm


@pytest.mark.parametrize("name,B", [("lrelu_enc_s1", 96), ("relu_enc_s2", 70), ("softplus_enc_s3", 96),
                                    ("softplus_b5_enc_s4", 64), ("lrelu_enc_l2_s8", 64), ("mixed_relu_softplus_s7", 64)])
def test_train_losses_and_parameter_gradients_vs_reference_autograd(name, B):
    from posendf_b200 import PoseNDF
    meta, _ = load_golden(name)
    cfg = case_cfg(meta)
    loss_type = meta.get("loss_type", "l1")
    params = synth.make_params(meta["seed"], sensitised=meta["sensitised"])
    net = PoseNDF(_opt(cfg, loss_type))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    net.train()
    tp, tgt, tm = _batch(meta["seed"], B)
    weights = {"dist": 1.0, "man_loss": 0.7, "eikonal": 1.3}
    net.zero_grad()
    loss, ld = net(torch.from_numpy(tp), torch.from_numpy(tgt), torch.from_numpy(tm), train=True, eikonal=1.0)
    assert set(ld) == {"dist", "man_loss", "eikonal"} and loss is ld["dist"]
    sum(weights[k] * v for k, v in ld.items()).backward()
    # oracle: the reference's computation in fp64
    tp64 = otorch.to_torch_params(params, torch.float64, requires_grad=True)
    _, ld_ref, g_ref = otorch.train_step_grads(tp64, torch.from_numpy(tp).double(), torch.from_numpy(tgt).double(),
                                                torch.from_numpy(tm).double(), cfg, weights=weights, loss_type=loss_type)
    for k in ("dist", "man_loss", "eikonal"):
        assert abs(ld[k].item() - ld_ref[k].item()) < 2e-5 * max(1.0, abs(ld_ref[k].item())), k
    worst = 0.0
    for n, p in net.named_parameters():
        ref = g_ref[n].numpy()
        got = p.grad.cpu().numpy().astype(np.float64)
        err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-12)
        worst = max(worst, err)
        # kink flips (relu/lrelu) perturb single samples; norm-wise 5e-3 per tensor, most are ~1e-6
        assert err < (5e-3 if cfg["df_act"] != "softplus" else 2e-4), (n, err)
    print(name, "worst per-tensor relative gradient error", worst)


def test_eikonal_off_returns_dist_only_and_matches():
    from posendf_b200 import PoseNDF
    meta, _ = load_golden("lrelu_enc_s1")
    cfg = case_cfg(meta)
    params = synth.make_params(1)
    net = PoseNDF(_opt(cfg))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    tp, tgt, tm = _batch(1, 64)
    loss, ld = net(torch.from_numpy(tp), torch.from_numpy(tgt), torch.from_numpy(tm), train=True, eikonal=0.0)
    assert set(ld) == {"dist"}
    loss.backward()
    tp64 = otorch.to_torch_params(params, torch.float64, requires_grad=True)
    tot, _, g_ref = otorch.train_step_grads(tp64, torch.from_numpy(tp).double(), torch.from_numpy(tgt).double(),
                                             torch.from_numpy(tm).double(), cfg, weights={"dist": 1.0}, eikonal=0.0)
    assert abs(loss.item() - tot.item()) < 1e-5
    for n, p in net.named_parameters():
        ref = g_ref[n].numpy()
        assert np.linalg.norm(p.grad.cpu().numpy() - ref) < 5e-3 * max(np.linalg.norm(ref), 1e-12), n


def _forward(net, fused, tp, tgt, tm, eikonal=1.0):
    args = (torch.from_numpy(tp), torch.from_numpy(tgt), torch.from_numpy(tm))
    return net(*args, train=True, eikonal=eikonal) if fused else train_forward_autograd(net, *args, eikonal)


def test_fused_train_step_equals_torch_autograd_path_and_adam_step_runs():
    """one trainer step (model/train_posendf.py:93-99) on both paths from the same init: losses and gradient norm agree."""
    from posendf_b200 import PoseNDF
    meta, _ = load_golden("softplus_enc_s3")
    cfg = case_cfg(meta)
    params = synth.make_params(3)
    tp, tgt, tm = _batch(3, 128)
    outs = []
    for fused in (True, False):
        net = PoseNDF(_opt(cfg))
        net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        opt = torch.optim.Adam(net.parameters(), lr=1e-5, weight_decay=1e-4)
        opt.zero_grad()
        _, ld = _forward(net, fused, tp, tgt, tm)
        sum(ld.values()).backward()
        gn = torch.sqrt(sum((p.grad ** 2).sum() for p in net.parameters())).item()
        opt.step()
        outs.append((gn, {k: v.item() for k, v in ld.items()}))
    assert abs(outs[0][0] - outs[1][0]) < 1e-3 * outs[1][0]
    for k in outs[0][1]:
        assert abs(outs[0][1][k] - outs[1][1][k]) < 1e-5 * max(1.0, abs(outs[1][1][k]))


@pytest.mark.parametrize("act,B", [("lrelu", 32768), ("softplus", 8192)])
def test_large_batch_split_k_path_matches_autograd_cross_check(act, B, monkeypatch):
    """real batch sizes take many K-splits of the weight-gradient kernel and (with a small CHUNK) several export launches per
    batch: every parameter gradient must agree with torch autograd over cuBLAS on the same GPU (fp32 both)."""
    from posendf_b200 import PoseNDF, train
    monkeypatch.setattr(train, "CHUNK", B // 2)
    cfg = dict(use_enc=True, enc_act=act, enc_beta=100.0, df_act=act, df_beta=100.0)
    params = synth.make_params(5)
    tp, tgt, tm = _batch(5, B)
    grads = []
    for fused in (True, False):
        net = PoseNDF(_opt(cfg))
        net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        _, ld = _forward(net, fused, tp, tgt, tm)
        (1.0 * ld["dist"] + 0.5 * ld["man_loss"] + 2.0 * ld["eikonal"]).backward()
        grads.append({n: p.grad.double().cpu() for n, p in net.named_parameters()})
    worst = 0.0
    for n in grads[0]:
        err = (grads[0][n] - grads[1][n]).norm().item() / max(grads[1][n].norm().item(), 1e-12)
        worst = max(worst, err)
        # relu / lrelu: units on their kink flip between the two fp32 evaluations (observed 1e-4); softplus ~1e-6
        assert err < (1e-3 if act != "softplus" else 2e-5), (n, err)
    print(act, B, "worst per-tensor relative difference fused vs autograd", worst)


def test_device_weight_repack_equals_host_upload_and_follows_optimizer_steps():
    """pndf_set_weights_device (gather kernel, model/train_posendf.py:99 -> next forward) == pndf_set_weights bit for bit"""
    from posendf_b200 import PoseNDF
    from posendf_b200.engine import Engine
    cfg = dict(use_enc=True, enc_act="lrelu", enc_beta=100.0, df_act="lrelu", df_beta=100.0)
    params = synth.make_params(9)
    flat = synth.flatten_params(params)
    x = torch.from_numpy(synth.make_poses(9, 96)).cuda()
    a, b = Engine(device=0, **{k: cfg[k] for k in ("enc_act", "df_act")}), Engine(device=0, **{k: cfg[k] for k in ("enc_act", "df_act")})
    a.set_weights_flat(flat)
    b.set_weights_device(torch.from_numpy(flat).cuda())
    da, ga = a.forward_grad(x)
    db, gb = b.forward_grad(x)
    assert torch.equal(da, db) and torch.equal(ga, gb)
    # the module re-packs after every optimizer step, on a side stream as well
    net = PoseNDF(_opt(cfg))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    opt = torch.optim.SGD(net.parameters(), lr=1e-2)
    tp, tgt, tm = _batch(9, 64)
    d0 = net(x, train=False)["dist_pred"].clone()
    _, ld = net(torch.from_numpy(tp), torch.from_numpy(tgt), torch.from_numpy(tm), train=True, eikonal=1.0)
    sum(ld.values()).backward()
    opt.step()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        d1 = net(x, train=False)["dist_pred"].clone()
    side.synchronize()
    c = Engine(device=0, **{k: cfg[k] for k in ("enc_act", "df_act")})
    c.set_weights_flat(np.concatenate([p.detach().cpu().numpy().ravel() for p in net.parameters()]))
    assert torch.equal(d1, c.forward(x)) and not torch.equal(d0, d1)


@pytest.mark.parametrize("act", ["lrelu", "softplus"])
def test_tangent_launch_with_mask_handoff_equals_two_pass_tangent(act):
    """pndf_forward_tangent_export given launch 1's activation derivatives (bit masks / fp32 for softplus; skips its
    primal pass) == the self-contained two-pass launch"""
    from posendf_b200 import train
    from posendf_b200.engine import Engine
    eng = Engine(device=0, enc_act=act, df_act=act)
    eng.set_weights_flat(synth.flatten_params(synth.make_params(4)))
    B = 200
    x = torch.from_numpy(synth.make_poses(4, B, kind="noisy", sigma=0.25)).cuda()
    v = torch.from_numpy(synth.make_poses(40, B, kind="raw")).cuda() * 1e-3
    outs = []
    for want in (True, False):
        ex = train._Exports(eng, x, True, want_masks=want)
        ex.v = v.contiguous()
        ex.tangent_launch(eng)
        outs.append(ex.dump_t[:B, :train.Z_END].clone())
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1])
    assert outs[0].abs().max().item() > 0


def test_dead_poses_zero_gradient_norm_stay_finite_and_match_autograd():
    """default-init amass.yaml weights with a ReLU output give d == 0 and an all-zero pose gradient for
    (almost) every pose; torch's norm backward takes the zero sub-gradient there.  The native step must stay finite and
    agree with the reference's autograd (fp64 oracle) -- the Eikonal tangent kernel masks |g| == 0."""
    from posendf_b200 import PoseNDF
    cfg = dict(use_enc=True, enc_act="lrelu", enc_beta=100.0, df_act="lrelu", df_beta=100.0)
    params = synth.make_params(1)
    net = PoseNDF(_opt(cfg))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    tp, tgt, tm = _batch(0, 96)
    # shift the output bias so that the last pre-activation straddles zero: about half of the poses get d == 0 (ReLU output)
    # and with it an all-zero pose gradient -- what default-init weights or near-manifold poses of a trained net produce
    d0 = net(torch.from_numpy(tp), train=False)["dist_pred"]
    ds = np.sort(d0.cpu().numpy().ravel())
    cut = 0.5 * (float(ds[47]) + float(ds[48]))          # between two poses: nobody sits ON the output kink (fp32 vs fp64 would flip it)
    params["dfnet.lin6.bias"] = (params["dfnet.lin6.bias"] - np.float32(cut)).astype(np.float32)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    dead = (net(torch.from_numpy(tp), train=False)["dist_pred"] == 0).float().mean().item()
    assert 0.2 < dead < 0.8
    _, ld = net(torch.from_numpy(tp), torch.from_numpy(tgt), torch.from_numpy(tm), train=True, eikonal=1.0)
    sum(ld.values()).backward()
    tp64 = otorch.to_torch_params(params, torch.float64, requires_grad=True)
    _, ld_ref, g_ref = otorch.train_step_grads(tp64, torch.from_numpy(tp).double(), torch.from_numpy(tgt).double(),
                                                torch.from_numpy(tm).double(), cfg, weights={"dist": 1.0, "man_loss": 1.0, "eikonal": 1.0})
    for k in ("dist", "man_loss", "eikonal"):
        assert np.isfinite(ld[k].item()) and abs(ld[k].item() - ld_ref[k].item()) < 2e-5 * max(1.0, abs(ld_ref[k].item())), k
    for n, p in net.named_parameters():
        assert torch.isfinite(p.grad).all(), n
        ref = g_ref[n].numpy()
        assert np.linalg.norm(p.grad.cpu().numpy() - ref) <= 5e-3 * np.linalg.norm(ref) + 1e-12, n


def test_gradient_accumulation_semantics_of_the_flat_buffer():
    """p.grad are views of ONE flat buffer: None grads are overwritten, attached ones accumulated, foreign tensors carried over"""
    from posendf_b200 import PoseNDF
    cfg = dict(use_enc=True, enc_act="lrelu", enc_beta=100.0, df_act="lrelu", df_beta=100.0)
    net = PoseNDF(_opt(cfg))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_params(2).items()})
    tp, tgt, tm = _batch(2, 64)

    def run():
        _, ld = _forward(net, True, tp, tgt, tm)
        sum(ld.values()).backward()

    run()
    g1 = net.flat_grad().clone()
    assert net.grads_attached() and g1.abs().sum().item() > 0
    run()                                                     # attached -> accumulates
    # (g1 + pose part) + manifold part vs 2 (pose part + manifold part): equal up to fp32 rounding of the running sum
    assert torch.allclose(net.flat_grad(), 2 * g1, rtol=1e-5, atol=1e-6 * g1.abs().max().item())
    net.zero_grad()                                           # torch default: grads -> None
    run()
    assert torch.equal(net.flat_grad(), g1)                   # overwritten, deterministic
    for p in net.parameters():                                # foreign gradient tensors are carried over
        p.grad = torch.ones_like(p)
    run()
    assert torch.allclose(net.flat_grad(), g1 + 1.0, rtol=1e-5, atol=1e-6) and net.grads_attached()


def test_fused_adam_matches_torch_adam_and_keeps_the_packed_weights_current():
    """posendf_b200.optim.FusedAdam (one kernel) vs torch.optim.Adam(lr, weight_decay=1e-4) (model/train_posendf.py:30) fed the
    same gradients for 4 steps; afterwards the engine's packed weights equal a fresh host upload of the parameters bit for bit."""
    from posendf_b200 import PoseNDF
    from posendf_b200.engine import Engine
    from posendf_b200.optim import FusedAdam
    cfg = dict(use_enc=True, enc_act="lrelu", enc_beta=100.0, df_act="lrelu", df_beta=100.0)
    params = synth.make_params(6)
    net = PoseNDF(_opt(cfg))
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    opt = FusedAdam(net, lr=1e-3, weight_decay=1e-4)
    ref_p = torch.from_numpy(synth.flatten_params(params)).cuda().requires_grad_(True)
    ref_opt = torch.optim.Adam([ref_p], lr=1e-3, weight_decay=1e-4)
    gen = torch.Generator(device="cuda").manual_seed(0)
    for t in range(4):
        g = torch.randn(ref_p.numel(), device="cuda", generator=gen) * (10.0 ** (t - 2))
        g[::7] = 0.0
        net.flat_grad().copy_(g)
        net.attach_grads()
        ref_p.grad = g.clone()
        opt.step()
        ref_opt.step()
        got = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
        assert (got - ref_p.detach()).abs().max().item() <= 2e-7 * max(1.0, ref_p.detach().abs().max().item()), t
    st = ref_opt.state[ref_p]
    assert torch.allclose(opt.exp_avg, st["exp_avg"], rtol=1e-6, atol=1e-12)
    assert torch.allclose(opt.exp_avg_sq, st["exp_avg_sq"], rtol=1e-6, atol=1e-20)
    x = torch.from_numpy(synth.make_poses(6, 96)).cuda()
    d, gr = net.engine().forward_grad(x)
    fresh = Engine(device=0)
    fresh.set_weights_flat(torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu().numpy())
    d2, gr2 = fresh.forward_grad(x)
    assert torch.equal(d, d2) and torch.equal(gr, gr2)


def test_data_parallel_step_single_process_equals_manual_sequence():
    """posendf_b200.dist.DataParallelStep (zero_grad, native losses, backward, FusedAdam) == the same calls written out with
    torch.optim.Adam on a twin module: same losses, parameters agree to fp32 round-off after 2 steps"""
    from posendf_b200 import PoseNDF
    from posendf_b200.dist import DataParallelStep
    cfg = dict(use_enc=True, enc_act="lrelu", enc_beta=100.0, df_act="lrelu", df_beta=100.0)
    params = synth.make_params(7)
    tp, tgt, tm = (torch.from_numpy(a).cuda() for a in _batch(7, 256))
    a, b = PoseNDF(_opt(cfg)), PoseNDF(_opt(cfg))
    for net in (a, b):
        net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    trainer = DataParallelStep(a, lr=1e-4, weight_decay=1e-4, weights=(1.0, 0.5, 2.0))
    ref_opt = torch.optim.Adam(b.parameters(), lr=1e-4, weight_decay=1e-4)
    for _ in range(2):
        la = trainer.step(tp, tgt, tm)
        ref_opt.zero_grad()
        _, lb = b(tp, tgt, tm, train=True, eikonal=2.0)
        (1.0 * lb["dist"] + 0.5 * lb["man_loss"] + 2.0 * lb["eikonal"]).backward()
        ref_opt.step()
        for k in la:
            assert abs(la[k].item() - lb[k].item()) <= 1e-6 * max(1.0, abs(lb[k].item())), k
    pa = torch.cat([p.detach().reshape(-1) for p in a.parameters()])
    pb = torch.cat([p.detach().reshape(-1) for p in b.parameters()])
    assert (pa - pb).abs().max().item() < 1e-6
    assert not torch.equal(pa, torch.from_numpy(synth.flatten_params(params)).cuda())
