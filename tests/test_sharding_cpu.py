"""world_size-2 gloo test (CPU) of the multi-GPU host logic: round-robin row ownership, the lag mask in
global labels, one all-gather of per-shard top-k, and the deterministic merge rule.  The per-shard engine
here is the numpy oracle; on the GPU the same flow runs through libcerebro_b200 (tests/test_search_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cerebro_b200.index import TIE_HIGH_LABEL, TIE_LOW_LABEL, merge_topk_host, shard_rows
from tests import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, d, nq, k, limit, tie, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    db = synth.unit_rows(n, d, seed=42)
    db[300] = db[100]  # exact tie across shards (300 % 2 == 100 % 2 -> same shard), and
    db[301] = db[100]  # across different shards
    xq = np.concatenate([synth.planted_queries(db, [100, 7, n - 1], seed=3), db[100:101]])
    mine = shard_rows(n, rank, world)
    lim = n if limit is None else limit
    mine = mine[mine < lim]  # lag mask applied locally, in GLOBAL labels
    s = xq.astype(np.float64) @ db[mine].astype(np.float64).T
    S = np.full((nq, k), -np.inf)
    L = np.full((nq, k), -1, dtype=np.int64)
    for q in range(nq):
        order = np.lexsort(((-mine if tie == TIE_HIGH_LABEL else mine), -s[q]))[:k]
        S[q, : len(order)] = s[q, order]
        L[q, : len(order)] = mine[order]
    gs = [torch.empty((nq, k), dtype=torch.float64) for _ in range(world)]
    gl = [torch.empty((nq, k), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gs, torch.from_numpy(S))
    dist.all_gather(gl, torch.from_numpy(L))
    Sm, Lm = merge_topk_host(torch.stack(gs).numpy(), torch.stack(gl).numpy(), k, tie)
    if rank == 0:
        np.savez(out_path, S=Sm, L=Lm)
    # every rank must hold the identical merged list
    chk = torch.from_numpy(Lm.copy())
    dist.broadcast(chk, src=0)
    assert np.array_equal(chk.numpy(), Lm)
    dist.destroy_process_group()


@pytest.mark.parametrize("limit,tie", [(None, TIE_LOW_LABEL), (250, TIE_LOW_LABEL), (None, TIE_HIGH_LABEL), (3, TIE_LOW_LABEL)])
def test_two_rank_sharded_topk_equals_unsharded(tmp_path, limit, tie):
    n, d, nq, k, world = 501, 128, 4, 5, 2
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(world, _free_port(), n, d, nq, k, limit, tie, out), nprocs=world, join=True)
    got = np.load(out)
    db = synth.unit_rows(n, d, seed=42)
    db[300] = db[100]
    db[301] = db[100]
    xq = np.concatenate([synth.planted_queries(db, [100, 7, n - 1], seed=3), db[100:101]])
    lim = n if limit is None else limit
    s = xq.astype(np.float64) @ db[:lim].astype(np.float64).T
    labels = np.arange(lim)
    for q in range(nq):
        order = np.lexsort(((-labels if tie == TIE_HIGH_LABEL else labels), -s[q]))[:k]
        kk = len(order)
        assert np.array_equal(got["L"][q, :kk], order), (q, got["L"][q], order)
        assert np.all(got["L"][q, kk:] == -1)
    if limit is None:
        exp = [300, 301] if tie == TIE_HIGH_LABEL else [100, 300]
        assert list(got["L"][3, :2]) == ([301, 300] if tie == TIE_HIGH_LABEL else [100, 300]), got["L"][3]


def test_shard_rows_partition():
    for world in (1, 2, 3, 8):
        allrows = np.sort(np.concatenate([shard_rows(1001, r, world) for r in range(world)]))
        assert np.array_equal(allrows, np.arange(1001))
