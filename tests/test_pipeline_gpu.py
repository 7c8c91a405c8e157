"""End-to-end parity of the Python Cerebro mirror on the GPU: keyframe images -> NetVLAD -> DB -> loop candidates
-> PnP, against the oracle pipeline (oracle descriptors -> oracle naive_stream -> oracle RANSAC) on the same inputs."""
import numpy as np
import pytest

from tests import golden_io, synth

pytestmark = pytest.mark.gpu


def test_python_cerebro_stream_matches_oracle(native_lib, cuda_device):
    from cerebro_b200.descriptor import NetvladDescriptor
    from cerebro_b200.keras_weights import fold_mobilenet_netvlad
    from cerebro_b200.loop_detector import Cerebro
    from cerebro_b200.pnp import default_params
    from oracle import dls_pnp as D
    from oracle import netvlad as NV
    from oracle.search import naive_stream

    raw = golden_io.raw_weights("gray_conv6")
    rows, cols, n = 96, 128, 96
    places = synth.band_limited_images(64, rows, cols, 1, seed=60)
    rng = np.random.default_rng(61)
    imgs = np.empty((n, rows, cols, 1), dtype=np.uint8)
    imgs[:64] = places
    for i in range(32):
        imgs[64 + i] = np.clip(places[3 + i].astype(np.int16) + rng.integers(-5, 6, places[0].shape), 0, 255).astype(np.uint8)
    nd = NetvladDescriptor(fold_mobilenet_netvlad(raw), rows, cols, 1, max_batch=3)
    cer = Cerebro(nd, capacity=200)
    stamps = [0.1 * (i + 1) for i in range(n)]
    tracked = [100] * n
    tracked[10] = 5  # < 20 tracked features: skipped (Cerebro.cpp:206-210)
    found = []
    for a in range(0, n, 3):
        cer.descriptor_step(stamps[a : a + 3], imgs[a : a + 3], tracked[a : a + 3])
        e = cer.run_step()
        if e is not None:
            found.append(e)
    keep = [i for i in range(n) if i != 10]
    assert cer.wholeImageComputedList_size() == len(keep)
    desc = NV.describe(imgs[keep], raw, dtype="float32").astype(np.float64)
    arrivals, l = [], 0
    for a in range(0, n, 3):
        l += sum(1 for i in range(a, min(a + 3, n)) if i != 10)
        arrivals.append(l)
    expected = naive_stream(desc, arrivals)
    assert len(expected) >= 5
    got = [(cer._whole.index(a), cer._whole.index(b)) for a, b, _ in found]
    assert got == [(a, b) for a, b, _ in expected]
    assert cer.foundLoops_count() == len(expected) and "dotprodt" in cer.foundLoops_as_JSON()
    # verify every candidate with synthetic correspondences
    rng2 = np.random.default_rng(62)
    cands = [D.synth_candidate(rng2, n=160, outlier_frac=0.1) for _ in found]
    out = cer.loopcandidate_consumer_step([(c[0], c[1]) for c in cands], default_params(seed=9))
    assert cer.processedLoops_count() == len(found)
    for j, rec in enumerate(out):
        o = D.ransac_pnp(cands[j][0], cands[j][1], D.sample_table(9, j, 50, 160))
        assert abs(rec["goodness"] - o["confidence"]) < 1e-6
        e = D.pose_error(rec["b_T_a"], o["T"])
        assert e[0] < 1e-3 and e[1] < 1e-2
    nd.close()


def test_faiss_naive_rule_matches_oracle(native_lib, cuda_device):
    """Cerebro.faiss_naive_step (top-5 search with a 150-keyframe lag) vs oracle.faiss_naive_stream on descriptors
    fed straight into the index (the descriptor stage is covered above)."""
    from cerebro_b200.loop_detector import Cerebro
    from oracle.search import faiss_naive_stream

    class _FakeDesc:  # the rule only needs descriptor_size
        dim = 1024

    d, n = 1024, 420
    base = synth.unit_rows(n, d, seed=70)
    desc = base.copy()
    for i in range(45):
        desc[360 + i] = synth.planted_queries(base, [40 + i], seed=200 + i, score=0.95)[0]
    cer = Cerebro(_FakeDesc(), capacity=n)
    found = []
    for a in range(0, n, 3):
        cer.index.add(desc[a : a + 3])
        cer._whole.extend(range(a, a + 3))
        e = cer.faiss_naive_step()
        if e is not None:
            found.append(e)
    expected = faiss_naive_stream(desc, list(range(3, n + 1, 3)))
    assert len(expected) >= 5
    assert [(a, b) for a, b, _ in found] == [(a, b) for a, b, _ in expected]
    assert np.allclose([s for *_, s in found], [s for *_, s in expected], atol=1e-6)


def test_descriptor_row_stride(native_lib, cuda_device):
    """cb_descriptor_compute with padded image rows (cv::Mat step > cols*channels, Cerebro.cpp:243-256)."""
    import ctypes as C

    from cerebro_b200 import _lib
    from cerebro_b200.descriptor import NetvladDescriptor
    from cerebro_b200.keras_weights import fold_mobilenet_netvlad

    nd = NetvladDescriptor(fold_mobilenet_netvlad(golden_io.raw_weights("gray_conv6")), 96, 128, 1, max_batch=2)
    imgs = synth.band_limited_images(2, 96, 128, 1, seed=3)
    ref = nd.compute(imgs)
    padded = np.zeros((2, 96, 160), dtype=np.uint8)
    padded[:, :, :128] = imgs[..., 0]
    out = np.empty((2, nd.dim), dtype=np.float32)
    _lib.check(_lib.load().cb_descriptor_compute(nd._h, 2, _lib.ptr(padded), 160, _lib.ptr(out)))
    assert np.array_equal(out, ref)
    nd.close()


def test_pageable_pinned_and_f64_entry_points_agree(native_lib, cuda_device):
    """cb_descriptor_compute from pageable memory (the library's pinned staging ring, more frames than one ring buffer and a
    padded row stride) == from pinned memory (uploaded in place) == cb_descriptor_compute_f64 narrowed back (the service's
    float64[] reply, srv/WholeImageDescriptorCompute.srv:4)."""
    import torch

    from cerebro_b200 import _lib
    from cerebro_b200.descriptor import NetvladDescriptor
    from cerebro_b200.keras_weights import fold_mobilenet_netvlad

    n = 21  # > 2 x 8 staged frames: every ring buffer is reused
    nd = NetvladDescriptor(fold_mobilenet_netvlad(golden_io.raw_weights("gray_conv6")), 96, 128, 1, max_batch=n)
    imgs = synth.textured_scenes(n, 96, 128, 1, seed=8)
    pageable = nd.compute(imgs)
    pinned_in = torch.from_numpy(imgs).pin_memory()
    pinned = nd.compute(pinned_in.numpy())
    assert np.array_equal(pageable, pinned)
    f64 = nd.compute_f64(imgs)
    assert f64.dtype == np.float64 and np.array_equal(f64.astype(np.float32), pageable) and np.array_equal(f64, pageable.astype(np.float64))
    padded = np.zeros((n, 96, 160), dtype=np.uint8)
    padded[:, :, :128] = imgs[..., 0]
    out = np.empty((n, nd.dim), dtype=np.float64)
    _lib.check(_lib.load().cb_descriptor_compute_f64(nd._h, n, _lib.ptr(padded), 160, _lib.ptr(out)))
    assert np.array_equal(out, f64)
    nd.close()


def test_clique_and_multihypothesis_rules_match_oracle(native_lib, cuda_device):
    """faiss_clique_loopcandidate_generator / faiss_multihypothesis_tracking (Cerebro.cpp:506-885) on the device index,
    one batched top-5 search per wake-up, against the oracle's one-query-at-a-time replay."""
    from cerebro_b200.loop_detector import Cerebro
    from oracle import search as S

    class FakeDesc:
        dim = 256

    n = 300
    desc = synth.unit_rows(n, 256, seed=21)
    for i in range(14):
        desc[230 + i] = synth.planted_queries(desc, [30 + i], seed=70 + i, score=0.95)[0]
    for i in range(6):
        desc[270 + i] = synth.planted_queries(desc, [100 + i], seed=90 + i, score=0.93)[0]
    arrivals = [2, 5, 6, 9] + list(range(12, n + 1, 3))
    stamps = [1000 + 7 * i for i in range(n)]
    state = {"x": 12345}

    def lcg():  # deterministic stand-in for rand(), same sequence for product and oracle
        state["x"] = (1103515245 * state["x"] + 12345) & 0x7FFFFFFF
        return state["x"]

    expect = S.faiss_clique_stream(desc, arrivals, rand=lcg)
    assert len(expect) >= 3
    state["x"] = 12345
    c = Cerebro(FakeDesc(), capacity=n)
    got, fed = [], 0
    for l in arrivals:
        c.index.add(desc[fed:l])
        c._whole.extend(stamps[fed:l])
        fed = l
        got += c.faiss_clique_step(rand=lcg)
    assert [(a, b, s) for a, b, s in got] == [(stamps[a], stamps[b], s) for a, b, s in expect]
    assert c.foundLoops_count() == len(expect)

    hm_o = S.faiss_multihypothesis_stream(desc, arrivals)
    c2 = Cerebro(FakeDesc(), capacity=n)
    fed = 0
    for l in arrivals:
        c2.index.add(desc[fed:l])
        c2._whole.extend(stamps[fed:l])
        fed = l
        c2.faiss_multihypothesis_step()
    hm = c2.hyp_manager
    assert len(hm.active_hyp) == len(hm_o.active_hyp) >= 2
    for h, ho in zip(hm.active_hyp, hm_o.active_hyp):
        assert h.get_ttl() == ho.time_to_live
        assert [(a, b) for a, b, _ in h.list_of_nodes_in_this_hypothesis] == [(a, b) for a, b, _ in ho.nodes]
        assert np.allclose([s for *_, s in h.list_of_nodes_in_this_hypothesis], [s for *_, s in ho.nodes], atol=1e-6)


def _stream_inputs(c, rows=480, cols=640, n_places=64, n_revisit=32, seed=160):
    """A keyframe stream at the benchmarked image size: textured scenes (whole-image descriptors that differ -- band-limited
    noise all looks alike to NetVLAD), then revisits of places 3.. with strong sensor noise."""
    places = synth.textured_scenes(n_places, rows, cols, c, seed=seed)
    rng = np.random.default_rng(seed + 1)
    imgs = np.empty((n_places + n_revisit, rows, cols, c), dtype=np.uint8)
    imgs[:n_places] = places
    for i in range(n_revisit):
        imgs[n_places + i] = np.clip(places[3 + i].astype(np.int16) + rng.integers(-25, 26, places[0].shape), 0, 255).astype(np.uint8)
    return imgs


@pytest.mark.parametrize("model,c,tie_tol,min_exact", [("gray_conv6", 1, 0.0, 1.0), ("mobilenet_conv7", 3, 1e-4, 0.75)])
def test_bench_config_descriptors_give_identical_candidates(native_lib, cuda_device, model, c, tie_tol, min_exact):
    """north_star's parity criterion at the benchmarked configuration (480x640, the default 8192-D model and the 4096-D
    gray model): device descriptors -> device search give the SAME foundLoops list and the same top-5 labels as oracle
    (fp32, the reference's arithmetic) descriptors -> oracle search.  No score carve-out.

    The 4096-D model separates the scenes (scores -0.2 .. 0.6, revisits 0.98+): its top-5 must be identical, full stop.
    The default mobilenet_conv7_allpairloss checkpoint maps every image onto nearly the same descriptor (oracle scores of
    unrelated scenes 0.98 .. 0.99999, rank gaps ~1e-5, i.e. below what two fp32 implementations agree on): for it a
    rank may differ only between labels whose ORACLE scores are within 1e-4 of each other, and 75 % of the queries must
    match exactly (measured 27 of 32)."""
    from cerebro_b200.descriptor import NetvladDescriptor
    from cerebro_b200.index import TIE_LOW_LABEL, IndexFlatIP
    from cerebro_b200.keras_weights import fold_model
    from cerebro_b200.loop_detector import Cerebro
    from oracle import netvlad as NV
    from oracle import search as S

    raw = golden_io.raw_weights(model)
    imgs = _stream_inputs(c)
    n = imgs.shape[0]
    nd = NetvladDescriptor(fold_model(raw), 480, 640, c, max_batch=3)
    cer = Cerebro(nd, capacity=200)
    found = []
    for a in range(0, n, 3):
        cer.descriptor_step([0.1 * (i + 1) for i in range(a, a + 3)], imgs[a : a + 3])
        e = cer.run_step()
        if e is not None:
            found.append(e)
    ref = NV.describe(imgs, raw, dtype="float32")
    dev = cer.index.get_rows(0, n)
    err = np.linalg.norm(dev.astype(np.float64) - ref.astype(np.float64), axis=1)
    print("%s 480x640: descriptor L2 err max %.2e mean %.2e" % (model, err.max(), err.mean()))
    assert err.max() < 2e-3
    expected = S.naive_stream(ref.astype(np.float64), list(range(3, n + 1, 3)))
    assert len(expected) >= 8
    assert [(cer._whole.index(a), cer._whole.index(b)) for a, b, _ in found] == [(a, b) for a, b, _ in expected]
    assert np.allclose([s for *_, s in found], [s for *_, s in expected], atol=2e-4)
    # top-5 of every revisit against the rows that precede it by the reference's 50-keyframe lag
    o = S.IndexFlatIP(nd.dim)
    o.add(ref)
    exact = 0
    queries = list(range(64, n))
    for i in queries:
        Dg, Ig = cer.index.search(dev[i : i + 1], 5, limit_rows=i - 50, tie=TIE_LOW_LABEL)
        so = ref[: i - 50].astype(np.float64) @ ref[i].astype(np.float64)
        Io = np.argsort(-so, kind="stable")[:5]
        if np.array_equal(Ig[0], Io):
            exact += 1
            continue
        assert tie_tol > 0.0, "query %d: top-5 %s != oracle %s" % (i, Ig[0], Io)
        for r in range(5):
            if Ig[0, r] != Io[r]:
                assert abs(so[Ig[0, r]] - so[Io[r]]) < tie_tol, "query %d rank %d: %d vs %d, oracle scores %.6f / %.6f" % (
                    i, r, Ig[0, r], Io[r], so[Ig[0, r]], so[Io[r]])
    print("%s: %d / %d top-5 lists identical" % (model, exact, len(queries)))
    assert exact >= min_exact * len(queries)
    nd.close()
