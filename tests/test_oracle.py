"""CPU tests that pin the oracle: committed golden fixtures, invariants, independent solvers."""
import math

import os

import numpy as np
import pytest

from oracle import dls_pnp as D
from oracle import netvlad as NV
from oracle import search as S
from tests import golden_io, synth


# ------------------------------------------------------------------ DLS-PnP
def _T(C, t):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = C, t
    return T


@pytest.mark.parametrize("seed", range(5))
def test_dls_noise_free_recovery(seed):
    rng = np.random.default_rng(seed)
    X, uv, T, _ = D.synth_candidate(rng, n=15, noise=0.0, outlier_frac=0.0)
    sols = D.dls_pnp(X, uv)
    assert len(sols) >= 1
    err = min(D.pose_error(_T(*s), T) for s in sols)
    assert err[0] < 1e-9 and err[1] < 1e-9


def test_dls_all_27_roots_satisfy_gradient():
    rng = np.random.default_rng(3)
    X, uv, _, _ = D.synth_candidate(rng, n=15, noise=1e-3, outlier_frac=0.0)
    _, coef, _ = D.dls_setup(X, uv)
    s, real = D.dls_roots(coef)
    assert s.shape == (27, 3)
    for j in range(27):
        mono = np.array([s[j, 0] ** e[0] * s[j, 1] ** e[1] * s[j, 2] ** e[2] for e in D.M3])
        scale = np.abs(coef).max() * max(1.0, np.abs(mono).max())
        assert np.abs(coef @ mono).max() < 1e-6 * scale
    assert real.sum() >= 1


def test_macaulay_block_triangular_structure():
    """The 93x93 block ordered by descending degree is block upper-triangular with diagonal
    blocks 36,27,18,9,3 -- the structure the CUDA elimination exploits."""
    rng = np.random.default_rng(0)
    X, uv, _, _ = D.synth_candidate(rng, n=15)
    _, coef, _ = D.dls_setup(X, uv)
    M11 = D.macaulay_matrix(coef)[27:, 27:]
    deg = np.array([sum(e) for e in D.NONRED])
    assert [int((deg == p).sum()) for p in (7, 6, 5, 4, 3)] == [36, 27, 18, 9, 3]
    assert np.all(M11[deg[:, None] < deg[None, :]] == 0)


def test_dls_agrees_with_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    X, uv, T, _ = D.synth_candidate(rng, n=60, noise=1e-3, outlier_frac=0.0)
    sols = D.dls_pnp(X, uv)
    assert len(sols) == 1
    ok, rvec, tvec = cv2.solvePnP(X, uv, np.eye(3), None, flags=cv2.SOLVEPNP_SQPNP)
    R, _ = cv2.Rodrigues(rvec)
    e = D.pose_error(_T(*sols[0]), _T(R, tvec.ravel()))
    assert e[0] < 1e-3 and e[1] < 1e-2


def test_sampler_is_deterministic_and_distinct():
    a = D.sample_indices(5, 3, 17, 200)
    b = D.sample_indices(5, 3, 17, 200)
    assert np.array_equal(a, b) and len(set(a.tolist())) == 15 and a.min() >= 0 and a.max() < 200
    assert not np.array_equal(a, D.sample_indices(5, 3, 18, 200))
    small = D.sample_indices(1, 0, 0, 15)
    assert sorted(small.tolist()) == list(range(15))
    # known-answer vector so the CUDA sampler can be pinned bit for bit
    assert D.sample_indices(99, 0, 0, 200).tolist() == golden_io.load("pnp_golden.npz")["c0_tab"][0].tolist()


def test_ransac_golden_and_truth():
    g = golden_io.load("pnp_golden.npz")
    for c in range(6):
        X, uv, tab = g["c%d_X" % c], g["c%d_uv" % c], g["c%d_tab" % c]
        r = D.ransac_pnp(X, uv, tab)
        meta = g["c%d_adaptive_meta" % c]
        assert r["num_iterations"] == int(meta[1]) and r["n_inliers"] == int(meta[2]) and r["best_hyp"] == int(meta[3])
        assert np.allclose(r["T"], g["c%d_adaptive_T" % c], atol=1e-9)
        assert abs(r["confidence"] - meta[0]) < 1e-12
        if r["confidence"] > 0.5:  # candidates with 30-40 % outliers legitimately fail at 50 iterations
            e = D.pose_error(r["T"], g["c%d_Ttrue" % c])
            assert e[0] < 5e-3 and e[1] < 5e-2


def test_ransac_refuses_few_points_and_bookkeeping():
    rng = np.random.default_rng(4)
    X, uv, _, _ = D.synth_candidate(rng, n=19)
    assert D.ransac_pnp(X, uv, D.sample_table(0, 0, 50, 19))["confidence"] == -1.0  # DlsPnpWithRansac.cpp:136-139
    X, uv, _, _ = D.synth_candidate(rng, n=200, outlier_frac=0.0)
    r = D.ransac_pnp(X, uv, D.sample_table(0, 0, 50, 200))
    # all inliers: ratio 1.0 -> bound collapses to min_iterations (5)
    assert r["num_iterations"] == 5 and r["n_inliers"] == 200 and r["confidence"] == 1.0
    p = D.RansacParameters(adaptive=False, max_iterations=20)
    r = D.ransac_pnp(X, uv, D.sample_table(0, 0, 20, 200), p)
    assert r["num_iterations"] == 20


# ------------------------------------------------------------------ search
def test_last_argmax_tie_rule():
    assert S.last_argmax(np.array([0.1, 0.9, 0.3, 0.9, 0.2])) == 3  # Cerebro.cpp:1039-1043


def test_naive_stream_finds_planted_loops():
    d, n = 256, 300
    base = synth.unit_rows(n, d, seed=1)
    desc = base.copy()
    for i in range(30):
        desc[200 + i] = synth.planted_queries(base, [50 + i], seed=10 + i, score=0.95)[0]
    found = S.naive_stream(desc.astype(np.float64), list(range(3, n + 1, 3)))
    assert found and all(200 <= a < 230 and abs((a - 150) - b) <= 0 for a, b, _ in found)
    assert all(s > 0.85 for *_, s in found)
    # lag: nothing can be found while k = l - 50 <= 5
    assert S.naive_step(np.ascontiguousarray(desc.T.astype(np.float64)), 55) is None


def test_index_flat_ip_contract():
    db = synth.unit_rows(100, 128, seed=2)
    ix = S.IndexFlatIP(128)
    ix.add(db[:40])
    ix.add(db[40:])
    assert ix.ntotal == 100
    Dd, I = ix.search(db[7], 5)
    assert I[0, 0] == 7 and np.all(np.diff(Dd[0]) <= 0)
    Dd, I = ix.search(db[7], 5, limit_rows=3)
    assert set(I[0, :3]) == {0, 1, 2} and np.all(I[0, 3:] == -1) and np.all(np.isinf(Dd[0, 3:]))


def test_faiss_naive_stream_runs():
    d, n = 256, 400
    base = synth.unit_rows(n, d, seed=5)
    desc = base.copy()
    for i in range(30):
        desc[330 + i] = synth.planted_queries(base, [20 + i], seed=40 + i, score=0.95)[0]
    found = S.faiss_naive_stream(desc, list(range(3, n + 1, 3)))
    assert found and all(abs((a - 310) - b) <= 2 for a, b, _ in found)


# ------------------------------------------------------------------ NetVLAD
@pytest.mark.parametrize("model,c,dim", [("mobilenet_conv7", 3, 8192), ("gray_conv6", 1, 4096),
                                         ("mobilenetv2_block9_gray", 1, 1024), ("mobilenet_pw6", 3, 8192)])
def test_netvlad_golden_and_invariants(model, c, dim):
    w = golden_io.raw_weights(model)
    gold = golden_io.load("netvlad_golden.npz")["%s_96x128_desc64" % model]
    imgs = synth.band_limited_images(2, 96, 128, c, seed=96 + c)
    d64 = NV.describe(imgs, w, dtype="float64")
    assert d64.shape == (2, dim)
    assert np.allclose(d64, gold, atol=1e-12)
    d32 = NV.describe(imgs, w, dtype="float32")
    assert np.abs(d32 - d64).max() < 2e-5
    # predict_utils.py:59-61: unit norm; every K-major block was intra-normalised, so all blocks
    # share one norm -- except clusters whose soft-assignment mass is so small that
    # tf.nn.l2_normalize's epsilon (1e-12 on the squared norm) clamps them below it.
    assert np.allclose(np.linalg.norm(d64, axis=1), 1.0, atol=1e-12)
    K = 16
    bn = np.linalg.norm(d64.reshape(2, K, dim // K), axis=2)
    top = bn.max(axis=1, keepdims=True)
    assert np.all((np.abs(bn - top) < 1e-12) | (bn < top))
    assert np.all((np.abs(bn - top) < 1e-12).sum(axis=1) >= 4)


def test_netvlad_plus_centers_sign():
    """predict_utils.py:47 adds the cluster centres (x + C); flipping the sign must change the output."""
    w = golden_io.raw_weights("gray_conv6")
    imgs = synth.band_limited_images(1, 96, 128, 1, seed=1)
    a = NV.describe(imgs, w, dtype="float64")
    w2 = dict(w)
    w2["net_vlad_layer_1/cluster_centers"] = -w["net_vlad_layer_1/cluster_centers"]
    b = NV.describe(imgs, w2, dtype="float64")
    assert np.abs(a - b).max() > 1e-4


def test_bn_folding_matches_unfolded_oracle():
    """The product folds BN into conv weights on the host; check that algebra on CPU."""
    import torch
    import torch.nn.functional as F

    from cerebro_b200.keras_weights import fold_mobilenet_netvlad

    w = golden_io.raw_weights("gray_conv6")
    net = fold_mobilenet_netvlad(w)
    imgs = synth.band_limited_images(1, 64, 96, 1, seed=2)
    x = NV.preprocess(imgs, torch.float64)
    _, acts = NV.backbone(x, w, return_all=True)
    k = torch.as_tensor(net["conv1_w"], dtype=torch.float64).permute(3, 2, 0, 1)
    y = F.conv2d(F.pad(x, (0, 1, 0, 1)), k, stride=2) + torch.as_tensor(net["conv1_b"], dtype=torch.float64).view(1, -1, 1, 1)
    y = torch.clamp(y, 0, 6)
    assert torch.allclose(y, acts[0], atol=1e-5)


def test_mobilenetv2_folding_and_cbw_roundtrip(tmp_path):
    """June2019 MobileNetV2 model: BN folding of the inverted-residual blocks (expand / depthwise / linear project / Add)
    reproduces the un-folded oracle, and the .cbw container round-trips the folded network."""
    import torch
    import torch.nn.functional as F

    from cerebro_b200 import keras_weights as KW

    w = golden_io.raw_weights("mobilenetv2_block9_gray")
    assert KW.is_mobilenetv2(w)
    net = KW.fold_model(w)
    assert net["arch"] == "mobilenetv2" and len(net["ir_blocks"]) == 10
    assert [b["stride"] for b in net["ir_blocks"]] == [1, 2, 1, 2, 1, 1, 2, 1, 1, 1]
    assert [b["residual"] for b in net["ir_blocks"]] == [0, 0, 1, 0, 1, 1, 0, 1, 1, 1]  # block_{2,4,5,7,8,9}_add in the model_config
    path = os.path.join(str(tmp_path), "v2.cbw")
    KW.save_cbw(path, net)
    net = KW.load_model(path)
    imgs = synth.band_limited_images(1, 64, 96, 1, seed=2)
    dt = torch.float64
    x = NV.preprocess(imgs, dt)
    ref, acts = NV.backbone_v2(x, w, return_all=True)

    def t(a):
        return torch.as_tensor(a, dtype=dt)

    def pw(x, wt, b):
        return F.conv2d(x, t(wt).t()[:, :, None, None]) + t(b).view(1, -1, 1, 1)

    def dw(x, wt, b, stride):
        k = t(wt).permute(2, 0, 1)[:, None]
        if stride == 2:
            y = F.conv2d(F.pad(x, (0, 1, 0, 1)), k, stride=2, groups=k.shape[0])
        else:
            y = F.conv2d(x, k, padding=1, groups=k.shape[0])
        return y + t(b).view(1, -1, 1, 1)

    y = F.conv2d(F.pad(x, (0, 1, 0, 1)), t(net["conv1_w"]).permute(3, 2, 0, 1), stride=2) + t(net["conv1_b"]).view(1, -1, 1, 1)
    y = torch.clamp(y, 0, 6)
    for b in net["ir_blocks"]:
        inp = y
        if b["expand_w"] is not None:
            y = torch.clamp(pw(y, b["expand_w"], b["expand_b"]), 0, 6)
        y = torch.clamp(dw(y, b["dw_w"], b["dw_b"], b["stride"]), 0, 6)
        y = pw(y, b["project_w"], b["project_b"])
        if b["residual"]:
            y = y + inp
    assert y.shape == ref.shape
    assert torch.allclose(y, ref, atol=1e-4), float((y - ref).abs().max())


# ------------------------------------------------------------------ candidate-logic variants
def test_clique_accumulation_signed_locality_quirk():
    """Cerebro.cpp:645: the duplicate test is ``(key - label) < 7`` WITHOUT abs over the ascending std::map: any retained
    key smaller than the new label absorbs it; a label more than 7 below every key opens a new clique."""
    r = {}
    S.clique_accumulate(r, np.float32([0.95, 0.9, 0.86, 0.5, 0.4]), [606, 168, 172, 11, 12])
    # 606 first; 168: 606-168=438 >= 7 -> new key; 172: smallest key 168: 168-172=-4 < 7 -> 168 absorbs; 0.5 < thresh: stop
    assert r == {606: 1, 168: 2}
    S.clique_accumulate(r, np.float32([0.99]), [900])  # 168 - 900 < 7: the smallest key absorbs even a far larger label
    assert r == {606: 1, 168: 3}
    S.clique_accumulate(r, np.float32([0.99]), [100])  # 168-100=68, 606-100=506: new clique
    assert r == {606: 1, 168: 3, 100: 1}


def test_faiss_clique_stream_finds_planted_revisit():
    n, d = 260, 256
    desc = synth.unit_rows(n, d, seed=21)
    for i in range(12):
        desc[220 + i] = synth.planted_queries(desc, [30 + i], seed=70 + i, score=0.95)[0]
    arrivals = list(range(2, n + 1, 2))
    found = S.faiss_clique_stream(desc, arrivals)
    assert found, "no clique candidates"
    assert all(sc == 0.9 for *_, sc in found)
    assert all(28 <= b <= 45 for _, b, _ in found) and all(a >= 219 for a, _, _ in found)
    # rand() returning 99 drops every candidate of a multi-clique reset, single-clique resets are always kept
    kept = S.faiss_clique_stream(desc, arrivals, rand=lambda: 99)
    assert len(kept) <= len(found)


def test_hypothesis_manager_ttl_and_chaining():
    hm = S.HypothesisManager()
    hm.add_node(500, 100, 0.9)
    assert len(hm.active_hyp) == 1 and hm.active_hyp[0].time_to_live == 20
    hm.add_node(503, 104, 0.91)  # within 7 of the newest node on both ends -> same hypothesis, ttl + 1
    assert len(hm.active_hyp) == 1 and hm.active_hyp[0].time_to_live == 21 and len(hm.active_hyp[0].nodes) == 2
    hm.add_node(503, 300, 0.9)  # b too far -> new hypothesis
    assert len(hm.active_hyp) == 2
    hm.digest()
    assert [h.time_to_live for h in hm.active_hyp] == [17, 16]
    for _ in range(10):
        hm.digest()
    assert [h.time_to_live for h in hm.active_hyp] == [0, 0] and not hm.active_hyp[0].is_hypothesis_active()
    hm.add_node(505, 106, 0.9)  # an expired hypothesis is never removed and still absorbs matching nodes
    assert len(hm.active_hyp) == 2 and hm.active_hyp[0].time_to_live == 1
    h = S.Hypothesis(0, 0, 1.0)
    h.time_to_live = 100
    h.increment_ttl()  # > 100 after the first increment -> a second one (HypothesisManager.h:115-117)
    assert h.time_to_live == 102


def test_ghostvlad_oracle_drops_ghost_clusters():
    """GhostVLADLayer.call (predict_utils.py:110-141): K + G clusters share the softmax, only the first K reach the output;
    each kept cluster block has norm 1/sqrt(K) and the whole descriptor norm 1."""
    import torch

    from oracle import netvlad as NV

    rng = np.random.default_rng(0)
    D, K, G = 64, 6, 3
    w = {"gv/kernel": rng.standard_normal((1, 1, D, K + G)).astype(np.float32) * 0.1,
         "gv/bias": rng.standard_normal((1, 1, K + G)).astype(np.float32) * 0.1,
         "gv/cluster_centers": rng.standard_normal((1, 1, 1, D, K + G)).astype(np.float32)}
    x = torch.as_tensor(rng.uniform(0, 6, (2, D, 5, 7)))
    full = NV.netvlad(x, w).numpy()
    ghost = NV.netvlad(x, w, num_ghost_clusters=G).numpy()
    assert full.shape == (2, (K + G) * D) and ghost.shape == (2, K * D)
    assert np.allclose(np.linalg.norm(ghost, axis=1), 1.0)
    blocks = ghost.reshape(2, K, D)
    assert np.allclose(np.linalg.norm(blocks, axis=2), 1.0 / np.sqrt(K))
    # same directions as the un-ghosted layer's first K blocks (the softmax is shared), different global scale
    fb = full.reshape(2, K + G, D)[:, :K]
    assert np.allclose(blocks * np.sqrt(K), fb * np.sqrt(K + G), atol=1e-12)


def test_register_resident_elimination_prototype_matches_dense_oracle():
    """numpy prototype of the control flow of dls_eliminate2_kernel (csrc/pnp.cu): per-row assembly from the template tables,
    Gauss-Jordan with the (high word, low word) arg-max and lowest-row tie rule, pivot scaling deferred to the read-out,
    D^T Y = E_J for the degree-7 block and the contraction with its right-hand sides -- against the dense 93 x 93 Schur
    complement of the oracle.  Guards the tables of csrc/dls_tables.h and the algebra the kernel relies on (CPU only)."""
    import re
    import struct

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "cerebro_b200", "csrc", "dls_tables.h")).read()

    def tab(name, shape):
        m = re.search(name + r"(\[\d+\])+ = \{([^}]*)\}", src)
        return np.array([float(x) for x in m.group(2).split(",")]).reshape(shape)

    row_poly, row_terms = tab("kRowPoly", (93,)).astype(int), tab("kRowTerms", (93, 20)).astype(int)
    border7, f0_terms, f0 = tab("kBorder7", (3,)).astype(int), tab("kF0Terms", (27, 4)).astype(int), tab("kF0", (4,))
    off = [0, 3, 12, 30, 57, 93]

    def key(x):
        b = struct.unpack("<Q", struct.pack("<d", abs(x)))[0]
        return b >> 32, b & 0xFFFFFFFF

    def gauss_jordan(A, n):
        used, myk, mypiv = [False] * n, [-1] * n, [1.0] * n
        for k in range(n):
            cands = [(key(A[i, k]), -i) for i in range(n) if not used[i] and A[i, k] == A[i, k]]
            p = -max(cands)[1]
            used[p], myk[p], mypiv[p] = True, k, A[p, k]
            pr, inv = A[p].copy(), 1.0 / A[p, k]
            for i in range(n):
                if i != p:
                    A[i, k + 1:] -= (A[i, k] * inv) * pr[k + 1:]
        return myk, mypiv

    def action(coef):
        cf, N = coef.reshape(60), np.zeros((60, 27))
        for blk in range(4):
            o0, n = off[blk], off[blk + 1] - off[blk]
            A = np.zeros((n, n + 27))
            for r in range(n):
                for t in range(20):
                    c, cd = cf[row_poly[o0 + r] * 20 + t], row_terms[o0 + r][t]
                    if cd < 27:
                        A[r, n + cd] += -c
                    elif cd - 27 >= o0:
                        A[r, cd - 27 - o0] = c
                    else:
                        A[r, n:] -= c * N[cd - 27]
            myk, mypiv = gauss_jordan(A, n)
            for i in range(n):
                N[o0 + myk[i]] = A[i, n:] / mypiv[i]
        o0, n = 57, 36
        A = np.zeros((36, 39))
        for r in range(n):
            for t in range(20):
                cd = row_terms[o0 + r][t]
                if cd >= 27 and cd - 27 >= o0:
                    A[cd - 27 - o0, r] = cf[row_poly[o0 + r] * 20 + t]
        for t in range(3):
            A[border7[t], 36 + t] = 1.0
        myk, mypiv = gauss_jordan(A, n)
        Y = np.zeros((36, 3))
        for i in range(n):
            Y[myk[i]] = A[i, 36:] / mypiv[i]
        for r in range(n):
            rr = np.zeros(27)
            for t in range(20):
                c, cd = cf[row_poly[o0 + r] * 20 + t], row_terms[o0 + r][t]
                if cd < 27:
                    rr[cd] -= c
                elif cd - 27 < o0:
                    rr -= c * N[cd - 27]
            N[57:60] += np.outer(Y[r], rr)
        S = np.zeros((27, 27))
        for b in range(27):
            for k in range(4):
                cd = f0_terms[b][k]
                if cd < 27:
                    S[b, cd] += f0[k]
                else:
                    S[b] += f0[k] * N[cd - 27]
        return S

    rng = np.random.default_rng(7)
    for _ in range(6):
        X, uv, _, _ = D.synth_candidate(rng, n=15, noise=1e-3, outlier_frac=0.0)
        coef = np.asarray(D.dls_setup(X, uv)[1])
        S, S0 = action(coef), D.action_matrix(coef)
        assert np.abs(S - S0).max() < 1e-8 * max(1.0, np.abs(S0).max())
