"""One rank of the multi-GPU parity run (launched by tests/test_sharded_nccl_gpu.py and tools/ under torchrun):
the sharded search behind the C ABI (cb_comm_* + cb_index_search_sharded[_device]) against the numpy oracle on the
whole database, and the streaming merge of BASELINE config 3 (keyframes of several sessions arriving on different ranks,
DB sharded round-robin, lag enforced with limit_rows) against a single-index oracle replay.  Exits non-zero on mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def oracle_topk(db, q, k, limit, tie_high):
    """fp64 scores of q against db[:limit], descending, ties by label (low first, or high first)."""
    n = len(db) if limit is None else min(limit, len(db))
    s = q.astype(np.float64) @ db[:n].astype(np.float64).T
    D = np.full((q.shape[0], k), -np.inf)
    I = np.full((q.shape[0], k), -1, dtype=np.int64)
    lab = np.arange(n)
    for j in range(q.shape[0]):
        order = np.lexsort(((-lab if tie_high else lab), -s[j]))[:k]
        D[j, : len(order)] = s[j, order]
        I[j, : len(order)] = order
    return D, I


def main():
    import torch
    import torch.distributed as dist

    from cerebro_b200 import synthetic
    from cerebro_b200.index import TIE_HIGH_LABEL, TIE_LOW_LABEL, Comm, ShardedIndex

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = Comm.from_torch_group(local)
    d, n = 1024, 6000 + 7  # not a multiple of the world size: ragged shards
    db = synthetic.unit_rows(n, d, seed=31)
    ix = ShardedIndex(d, n // world + 8, local, comm=comm)
    ix.add(db)  # every rank sees the global rows and keeps g % world == rank
    assert ix.local.nlocal == len(range(rank, n, world))
    # ---- (1) every rank searches its OWN queries: planted neighbours on every shard + exact duplicates (ties across shards)
    nq = 37
    targets = (np.arange(nq) * 131 + 17 * rank) % n
    q = synthetic.planted_queries(db, targets, seed=100 + rank, score=0.93)
    q[5] = db[(7 + rank) % n]  # exact copy of a row
    for tie in (TIE_LOW_LABEL, TIE_HIGH_LABEL):
        for limit in (None, n - 50, 40):
            Do, Io = oracle_topk(db, q, 5, limit, tie == TIE_HIGH_LABEL)
            s, l = ix.search_sharded_device(torch.from_numpy(q).to(dev), 5, limit_rows=limit, tie=tie)
            torch.cuda.synchronize()
            assert np.array_equal(l.cpu().numpy(), Io), (rank, tie, limit, l.cpu().numpy()[:3], Io[:3])
            assert np.allclose(s.cpu().numpy(), Do, rtol=0, atol=1e-12)
            Dh, Ih, Sh = ix.search_sharded(q, 5, limit_rows=limit, tie=tie)  # host-pointer entry point
            assert np.array_equal(Ih, Io) and np.allclose(Sh, Do, rtol=0, atol=1e-12)
    # the torch.distributed path of round 1 (same queries on every rank) agrees with the C-ABI collective
    q0 = torch.from_numpy(q).to(dev)
    dist.broadcast(q0, src=0)
    s_t, l_t = ix.search_device(q0, 5)
    s_c, l_c = ix.search_sharded_device(q0, 5)
    torch.cuda.synchronize()
    assert torch.equal(l_t, l_c) and torch.equal(s_t, s_c)
    # ---- (2) BASELINE config 3: live merge.  `world` sessions stream keyframes, 3 per rank per step; each step is ONE
    # collective search (lag 50 via limit_rows) followed by appending the gathered block.  Global arrival order of a step
    # is rank-major.  The loop candidates (top-1 above 0.85) must equal a single-index replay of the same order.
    steps, per = 40, 3
    total = steps * per * world
    base = synthetic.unit_rows(total, d, seed=77)
    stream = base.copy()
    for g in range(total // 2, total):  # the second half revisits the first half
        if g % 3 == 0:
            stream[g] = synthetic.planted_queries(base, [g - total // 2], seed=500 + g, score=0.95)[0]
    live = ShardedIndex(d, total // world + 8, local, comm=comm)
    found = []
    for t in range(steps):
        g0 = t * per * world
        mine = stream[g0 + rank * per : g0 + (rank + 1) * per]
        limit = max(g0 - 50, 0)
        s, l = live.search_sharded_device(torch.from_numpy(mine).to(dev), 1, limit_rows=limit, tie=TIE_HIGH_LABEL)
        live.add_gathered_device(per * world)
        torch.cuda.synchronize()
        for j in range(per):
            if limit > 0 and s[j, 0].item() > 0.85:
                found.append((g0 + rank * per + j, int(l[j, 0].item()), float(s[j, 0].item())))
    assert live.ntotal == total
    rows = live.local.get_rows(0, live.local.nlocal)
    assert np.array_equal(rows, stream[rank::world])  # every shard holds exactly its round-robin rows
    expect = []
    for t in range(steps):
        g0 = t * per * world
        limit = max(g0 - 50, 0)
        if limit == 0:
            continue
        for j in range(per):
            g = g0 + rank * per + j
            sc = stream[:limit].astype(np.float64) @ stream[g].astype(np.float64)
            a = int(np.flatnonzero(sc == sc.max())[-1])  # last-index tie rule (Cerebro.cpp:1039-1043)
            if sc[a] > 0.85:
                expect.append((g, a, float(sc[a])))
    assert [(a, b) for a, b, _ in found] == [(a, b) for a, b, _ in expect], (rank, found[:4], expect[:4])
    assert np.allclose([x[2] for x in found], [x[2] for x in expect], rtol=0, atol=1e-12)
    cnt = torch.tensor([len(found)], device=dev)
    dist.all_reduce(cnt)
    if rank == 0:
        print("sharded parity ok: world %d, NCCL %d, %d loop candidates in the live merge" % (world, comm._lib.cb_comm_nccl_version(), int(cnt.item())))
    dist.barrier()
    ix.local.close()
    live.local.close()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
