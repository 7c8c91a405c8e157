"""GPU parity: CUDA index (through the C ABI) vs the numpy oracle on seeded inputs."""
import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu


def _oracle_index(db):
    from oracle.search import IndexFlatIP as O

    o = O(db.shape[1])
    o.add(db)
    return o


@pytest.mark.parametrize("n,d,nq,k", [(1000, 4096, 1, 5), (2500, 8192, 3, 5), (10000, 4096, 64, 5), (777, 1024, 17, 32), (40, 128, 5, 8)])
def test_topk_matches_oracle(native_lib, cuda_device, n, d, nq, k):
    from cerebro_b200.index import IndexFlatIP

    db = synth.unit_rows(n, d, seed=42)
    rng = np.random.default_rng(7)
    targets = rng.integers(0, n, nq)
    xq = synth.planted_queries(db, targets, seed=8)
    ix = IndexFlatIP(d, capacity=n + 10)
    # add in uneven pieces, like keyframes trickling in
    cuts = [0, n // 3, n // 3 + 1, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        ix.add(db[a:b])
    assert ix.ntotal == n
    D, I, S = ix.search(xq, k, return_f64=True)
    Do, Io = _oracle_index(db).search(xq, k)
    assert np.array_equal(I, Io), "top-k labels differ from the oracle"
    assert np.all(I[:, 0] == targets)
    # fp64 re-scored values agree with the fp64 oracle to rounding
    ref = (xq.astype(np.float64) @ db.astype(np.float64).T)
    assert np.allclose(S, np.take_along_axis(ref, I, 1), rtol=0, atol=1e-12)
    assert np.allclose(D, Do, rtol=0, atol=1e-6)
    ix.close()


def test_limit_rows_and_padding(native_lib, cuda_device):
    from cerebro_b200.index import IndexFlatIP

    n, d = 300, 1024
    db = synth.unit_rows(n, d, seed=1)
    xq = synth.planted_queries(db, [250, 10], seed=2)
    ix = IndexFlatIP(d, capacity=n)
    ix.add(db)
    o = _oracle_index(db)
    for lim in (0, 3, 5, 100, 251, 300, 1000):
        D, I = ix.search(xq, 5, limit_rows=lim)
        Do, Io = o.search(xq, 5, limit_rows=lim)
        assert np.array_equal(I, Io), lim
        assert np.array_equal(np.isinf(D), np.isinf(Do))
    ix.close()


def test_tie_rules(native_lib, cuda_device):
    """Duplicate rows give exactly equal scores: LOW label first (FAISS-like) vs the
    reference's last-index-wins arg-max (src/Cerebro.cpp:1039-1043)."""
    from cerebro_b200.index import IndexFlatIP, TIE_HIGH_LABEL, TIE_LOW_LABEL

    n, d = 500, 1024
    db = synth.unit_rows(n, d, seed=3)
    db[400] = db[100]
    db[450] = db[100]
    xq = db[100:101].copy()
    ix = IndexFlatIP(d, capacity=n)
    ix.add(db)
    _, I = ix.search(xq, 3, tie=TIE_LOW_LABEL)
    assert list(I[0]) == [100, 400, 450]
    _, I = ix.search(xq, 3, tie=TIE_HIGH_LABEL)
    assert list(I[0]) == [450, 400, 100]
    ix.close()


def test_add_f64_and_get_rows(native_lib, cuda_device):
    from cerebro_b200.index import IndexFlatIP

    db = synth.unit_rows(64, 256, seed=5)
    ix = IndexFlatIP(256, capacity=64)
    ix.add(db.astype(np.float64))
    assert np.array_equal(ix.get_rows(0, 64), db)
    ix.close()


def test_sharded_search_single_process(native_lib, cuda_device):
    """World-4 sharding emulated on one GPU: 4 shard handles + device merge == unsharded oracle."""
    import torch

    from cerebro_b200.index import IndexFlatIP, merge_topk_device

    n, d, nq, k, world = 4001, 1024, 9, 5, 4
    db = synth.unit_rows(n, d, seed=11)
    xq = synth.planted_queries(db, np.arange(nq) * 400 + 3, seed=12)
    shards = [IndexFlatIP(d, capacity=n // world + 2, rank=r, world=world) for r in range(world)]
    for a in range(0, n, 1000):
        for s in shards:
            s.add(db[a : a + 1000])
    assert [s.nlocal for s in shards] == [len(range(r, n, world)) for r in range(world)]
    xq_dev = torch.from_numpy(xq).cuda()
    for lim in (None, 2000, 7):
        outs = [s.search_device(xq_dev, k, limit_rows=lim) for s in shards]
        gs = torch.stack([o[0] for o in outs])
        gl = torch.stack([o[1] for o in outs])
        S, L = merge_topk_device(gs, gl, k)
        torch.cuda.synchronize()
        Do, Io = _oracle_index(db).search(xq, k, limit_rows=lim)
        assert np.array_equal(L.cpu().numpy(), Io), lim
    for s in shards:
        s.close()


def test_naive_loop_candidate_stream(native_lib, cuda_device):
    """Replay Cerebro::descrip_N__dot__descrip_0_N over a synthetic trajectory that revisits
    earlier places; foundLoops must equal the oracle's list."""
    from cerebro_b200.index import IndexFlatIP
    from oracle.search import naive_stream

    d, n = 1024, 400
    rng = np.random.default_rng(21)
    base = synth.unit_rows(n, d, seed=20)
    desc = base.copy()
    # frames 300..339 revisit frames 100..139 (score ~0.93), in order
    for i in range(40):
        desc[300 + i] = synth.planted_queries(base, [100 + i], seed=100 + i, score=0.93)[0]
    arrivals = list(range(3, n + 1, 3))
    expected = naive_stream(desc.astype(np.float64), arrivals)
    assert len(expected) > 5
    ix = IndexFlatIP(d, capacity=n)
    found = []
    last = 0
    for l in arrivals:
        ix.add(desc[last:l])
        last = l
        ok, prev, score, _ = ix.naive_candidate(l)
        if ok:
            found.append((l - 1, prev, score))
    assert [(a, b) for a, b, _ in found] == [(a, b) for a, b, _ in expected]
    assert np.allclose([s for *_, s in found], [s for *_, s in expected], atol=1e-12)
    ix.close()


def test_full_size_properties(native_lib, cuda_device):
    """BASELINE config 4 size (100k x 8192): planted neighbours are retrieved; scores are
    invariant to how the DB is sharded (checksum of top-k over 1 vs 8 shards)."""
    import torch

    from cerebro_b200.index import IndexFlatIP, merge_topk_device

    n, d, nq, k = 100_000, 8192, 16, 5
    g = torch.Generator(device="cuda").manual_seed(5)
    db = torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)
    db /= db.norm(dim=1, keepdim=True)
    targets = torch.arange(nq, device="cuda") * 6000 + 17
    noise = torch.randn((nq, d), generator=g, device="cuda")
    noise /= noise.norm(dim=1, keepdim=True)
    xq = 0.9 * db[targets] + (1 - 0.81) ** 0.5 * noise
    xq /= xq.norm(dim=1, keepdim=True)
    xq = xq.contiguous()
    one = IndexFlatIP(d, capacity=n)
    one.add(db)
    S1, L1 = one.search_device(xq, k)
    torch.cuda.synchronize()
    assert torch.equal(L1[:, 0], targets)
    assert torch.all(S1[:, 0] > 0.85) and torch.all(S1[:, 1] < 0.2)
    one.close()
    world = 8
    shards = [IndexFlatIP(d, capacity=n // world + 1, rank=r, world=world) for r in range(world)]
    for s in shards:
        s.add(db)
    outs = [s.search_device(xq, k) for s in shards]
    S8, L8 = merge_topk_device(torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]), k)
    torch.cuda.synchronize()
    assert torch.equal(L8, L1)
    assert torch.equal(S8, S1)  # fp64 re-scoring is shard-independent, bit for bit


@pytest.mark.parametrize("n,d,nq", [(1500, 1024, 64), (300, 8192, 9), (5000, 4096, 40), (4500, 2048, 128), (2000, 1024, 200), (700, 512, 65)])
def test_sweep_variants_agree(native_lib, cuda_device, monkeypatch, n, d, nq):
    """The tensor-core sweep on tiled planes (default), on row-major planes (CB_TC_V1=1) and the fp32 CUDA-core sweep
    (CB_NO_TC=1) must return the same labels and fp64 scores as the oracle -- with rows trickling in between searches
    (planes extended incrementally), ragged tile tails and a row limit."""
    from cerebro_b200.index import IndexFlatIP

    db = synth.unit_rows(n, d, seed=11)
    rng = np.random.default_rng(3)
    half = n // 2 + 7
    t1 = rng.integers(0, half, nq)
    t2 = rng.integers(0, n, nq)
    q1 = synth.planted_queries(db, t1, seed=4)
    q2 = synth.planted_queries(db, t2, seed=5)
    o = _oracle_index(db)
    o_half = _oracle_index(db[:half])
    for env in ({}, {"CB_TC_V1": "1"}, {"CB_NO_TC": "1"}, {"CB_TOPK_ONE_PASS": "1"}, {"CB_NO_TC": "1", "CB_TOPK_ONE_PASS": "1"}, {"CB_TC_Q64": "1"}):
        for key in ("CB_TC_V1", "CB_NO_TC", "CB_TOPK_ONE_PASS", "CB_TC_Q64"):
            monkeypatch.delenv(key, raising=False)
        for key, val in env.items():
            monkeypatch.setenv(key, val)
        ix = IndexFlatIP(d, capacity=n + 3)
        ix.add(db[:half])
        D, I, S = ix.search(q1, 5, return_f64=True)
        assert np.array_equal(I, o_half.search(q1, 5)[1]), env
        ix.add(db[half:])
        D, I, S = ix.search(q2, 5, return_f64=True)
        assert np.array_equal(I, o.search(q2, 5)[1]), env
        ref = q2.astype(np.float64) @ db.astype(np.float64).T
        assert np.allclose(S, np.take_along_axis(ref, I, 1), rtol=0, atol=1e-12), env
        lim = n - 129
        D, I = ix.search(q2, 5, limit_rows=lim)
        assert np.array_equal(I, o.search(q2, 5, limit_rows=lim)[1]), env
        ix.close()


@pytest.mark.parametrize("nq", [3, 20])
def test_many_exact_ties_above_the_bound(native_lib, cuda_device, nq):
    """100 identical rows tie for the best score (more than the 32 candidates a list carries) in a DB large enough for
    the two-pass selection: the returned labels must follow the requested tie rule exactly, as in the oracle."""
    from cerebro_b200.index import TIE_HIGH_LABEL, TIE_LOW_LABEL, IndexFlatIP

    n, d = 9000, 512
    db = synth.unit_rows(n, d, seed=31)
    db[4000:4100] = db[123]  # 101 copies of row 123 in total
    db[8990:8995] = db[77]
    xq = np.concatenate([db[123][None], db[77][None], synth.planted_queries(db, np.arange(nq - 2) * 50 + 5000, seed=3)])
    ix = IndexFlatIP(d, capacity=n)
    ix.add(db)
    o = _oracle_index(db)
    D, I = ix.search(xq, 8, tie=TIE_LOW_LABEL)
    assert np.array_equal(I, o.search(xq, 8)[1])
    assert list(I[0]) == [123] + list(range(4000, 4007)) and list(I[1][:6]) == [77, 8990, 8991, 8992, 8993, 8994]
    D, I = ix.search(xq, 8, tie=TIE_HIGH_LABEL)
    assert list(I[0]) == list(range(4099, 4091, -1)) and list(I[1][:6]) == [8994, 8993, 8992, 8991, 8990, 77]
    ix.close()
