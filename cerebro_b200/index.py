"""Host-side mirror of the descriptor index the reference searches.

``IndexFlatIP`` keeps the names and argument meaning of the slice of ``faiss::IndexFlatIP``
that Cerebro uses (src/Cerebro.cpp:390 ``IndexFlatIP(d)``, :431 ``add(1, x)``, :460
``search(1, x, 5, distances, labels)``, ``ntotal``); all arithmetic happens in
``libcerebro_b200.so``.  ``ShardedIndex`` is the multi-GPU form: one process per GPU, rows
round-robin over ranks, per-shard top-k all-gathered over NCCL and merged on the device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr

TIE_LOW_LABEL = 0
TIE_HIGH_LABEL = 1


class IndexFlatIP:
    def __init__(self, d: int, capacity: int = 29000, device: int = 0, rank: int = 0, world: int = 1):
        # capacity default: the reference preallocates 29000 columns (src/Cerebro.cpp:946)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.d = int(d)
        self.device = device
        self.rank, self.world = rank, world
        check(self._lib.cb_index_create(C.byref(self._h), d, capacity, device, rank, world))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.cb_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ntotal(self) -> int:
        return int(self._lib.cb_index_ntotal(self._h))

    @property
    def nlocal(self) -> int:
        return int(self._lib.cb_index_nlocal(self._h))

    def reset(self):
        check(self._lib.cb_index_reset(self._h))

    # ---- add -------------------------------------------------------------------------
    def add(self, x) -> None:
        """x: [n, d] float32 (FAISS contract) or float64 (the VectorXd the reference stores,
        src/DataNode.cpp:427-444) host array, or a CUDA float32 torch tensor."""
        if hasattr(x, "is_cuda"):
            if not x.is_cuda:
                x = x.numpy()
            else:
                import torch

                assert x.dtype == torch.float32 and x.is_contiguous()
                n = x.numel() // self.d
                check(self._lib.cb_index_add_device(self._h, n, ptr(x), _lib.current_stream_ptr()))
                return
        x = np.ascontiguousarray(x).reshape(-1, self.d)
        if x.dtype == np.float64:
            check(self._lib.cb_index_add_f64(self._h, x.shape[0], ptr(x)))
        else:
            x = np.ascontiguousarray(x, dtype=np.float32)
            check(self._lib.cb_index_add(self._h, x.shape[0], ptr(x)))

    def add_local(self, x) -> None:
        """Append CUDA float32 rows straight into this shard's storage (bulk load of a shard)."""
        import torch

        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
        check(self._lib.cb_index_add_local_device(self._h, x.numel() // self.d, ptr(x), _lib.current_stream_ptr()))

    def set_timing(self, on: bool) -> None:
        check(self._lib.cb_index_set_timing(self._h, 1 if on else 0))

    def sweep_timing(self):
        """(total_ms, n_launches) of the sweep kernel since the last call."""
        ms = C.c_double(0.0)
        n = C.c_int(0)
        check(self._lib.cb_index_get_sweep_timing(self._h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # ---- search ----------------------------------------------------------------------
    def search(self, xq, k: int, limit_rows: int | None = None, tie: int = TIE_LOW_LABEL, return_f64: bool = False):
        """Host arrays in, host arrays out: (distances float32 [nq,k], labels int64 [nq,k])."""
        xq = np.ascontiguousarray(xq, dtype=np.float32).reshape(-1, self.d)
        nq = xq.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        S = np.empty((nq, k), dtype=np.float64)
        check(
            self._lib.cb_index_search(
                self._h, nq, ptr(xq), k, -1 if limit_rows is None else int(limit_rows), tie, ptr(D), ptr(I), ptr(S)
            )
        )
        return (D, I, S) if return_f64 else (D, I)

    def search_device(self, xq, k: int, limit_rows: int | None = None, tie: int = TIE_LOW_LABEL, out=None):
        """CUDA tensors in/out, asynchronous on torch's current stream.
        Returns (scores float64 [nq,k], labels int64 [nq,k]) for THIS shard."""
        import torch

        assert xq.is_cuda and xq.dtype == torch.float32 and xq.is_contiguous()
        nq = xq.numel() // self.d
        if out is None:
            s = torch.empty((nq, k), dtype=torch.float64, device=xq.device)
            l = torch.empty((nq, k), dtype=torch.int64, device=xq.device)
        else:
            s, l = out
        check(
            self._lib.cb_index_search_device(
                self._h, nq, ptr(xq), k, -1 if limit_rows is None else int(limit_rows), tie, ptr(s), ptr(l),
                _lib.current_stream_ptr(),
            )
        )
        return s, l

    def naive_candidate(self, l: int, lag: int = 50, locality_thresh: int = 12, dot_thresh: float = 0.85):
        """One iteration of Cerebro::descrip_N__dot__descrip_0_N (src/Cerebro.cpp:1019-1081).
        Returns (found, prev, score, (argmax_v, argmax_vm, argmax_vmm))."""
        found = C.c_int(0)
        prev = C.c_int64(-1)
        score = C.c_double(0.0)
        am = (C.c_int64 * 3)(-1, -1, -1)
        check(
            self._lib.cb_index_naive_candidate(
                self._h, l, lag, locality_thresh, dot_thresh, C.byref(found), C.byref(prev), C.byref(score), am
            )
        )
        return bool(found.value), int(prev.value), float(score.value), tuple(int(v) for v in am)

    def get_rows(self, first_local: int, n: int) -> np.ndarray:
        out = np.empty((n, self.d), dtype=np.float32)
        check(self._lib.cb_index_get_rows(self._h, first_local, n, ptr(out)))
        return out


def merge_topk_device(scores, labels, k_out: int, tie: int = TIE_LOW_LABEL):
    """scores/labels: CUDA tensors [n_lists, nq, k_in] -> ([nq,k_out] f64, [nq,k_out] i64)."""
    import torch

    lib = _lib.load()
    n_lists, nq, k_in = scores.shape
    assert scores.dtype == torch.float64 and labels.dtype == torch.int64
    scores = scores.contiguous()
    labels = labels.contiguous()
    os_ = torch.empty((nq, k_out), dtype=torch.float64, device=scores.device)
    ol = torch.empty((nq, k_out), dtype=torch.int64, device=scores.device)
    check(lib.cb_topk_merge_device(n_lists, nq, k_in, ptr(scores), ptr(labels), k_out, tie, ptr(os_), ptr(ol), _lib.current_stream_ptr()))
    return os_, ol


def shard_rows(n_total: int, rank: int, world: int) -> np.ndarray:
    """Global labels owned by ``rank`` (round-robin by insertion index)."""
    return np.arange(rank, n_total, world)


COMM_ID_BYTES = 128


class Comm:
    """cb_comm: the NCCL communicator behind the C ABI.  ``Comm.from_torch_group()`` creates the id on rank 0 and hands it
    round with torch.distributed (any backend) -- a C++ host would use its own transport for those 128 bytes."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int):
        self._lib = _lib.load()
        assert len(unique_id) == COMM_ID_BYTES
        self._id = np.frombuffer(unique_id, dtype=np.uint8).copy()
        h = C.c_void_p()
        check(self._lib.cb_comm_create(C.byref(h), ptr(self._id), rank, world, device))
        self._h = h
        self.rank, self.world, self.device = rank, world, device

    @staticmethod
    def unique_id() -> bytes:
        buf = np.zeros(COMM_ID_BYTES, dtype=np.uint8)
        check(_lib.load().cb_comm_get_unique_id(ptr(buf)))
        return buf.tobytes()

    @classmethod
    def from_torch_group(cls, device: int, group=None):
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(box[0], rank, world, device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.cb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedIndex:
    """Descriptor DB sharded over the GPUs of one box (one process per GPU).

    Every rank calls ``add`` with the same global rows (or only its own via ``add_local``).  ``search_sharded_device`` is the
    product path: every rank passes its OWN queries and the C ABI does query all-gather -> local sweep -> ONE ncclAllGather
    of the packed per-shard top-k -> merge (``cb_index_search_sharded_device``).  ``search_device`` (the same queries on
    every rank, lists gathered with torch.distributed) is kept as the cross-check the tests compare it with."""

    def __init__(self, d: int, capacity_per_shard: int, device: int, group=None, comm: "Comm | None" = None):
        import torch.distributed as dist

        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = IndexFlatIP(d, capacity_per_shard, device, self.rank, self.world)
        self.d = d
        self.comm = None
        if comm is not None:
            self.attach_comm(comm)

    def attach_comm(self, comm: "Comm"):
        check(self.local._lib.cb_index_attach_comm(self.local._h, comm._h))
        self.comm = comm

    @property
    def ntotal(self):
        return self.local.ntotal

    def add(self, x):
        self.local.add(x)

    def search_sharded_device(self, xq_local, k: int, limit_rows: int | None = None, tie: int = TIE_LOW_LABEL, out=None):
        """CUDA tensor [nq_local, d] of THIS rank's queries in; merged global top-k of those queries out
        (scores float64 [nq_local, k], labels int64 [nq_local, k]).  Collective: every rank must call it."""
        import torch

        assert xq_local.is_cuda and xq_local.dtype == torch.float32 and xq_local.is_contiguous()
        nq = xq_local.numel() // self.d
        if out is None:
            s = torch.empty((nq, k), dtype=torch.float64, device=xq_local.device)
            l = torch.empty((nq, k), dtype=torch.int64, device=xq_local.device)
        else:
            s, l = out
        check(self.local._lib.cb_index_search_sharded_device(
            self.local._h, nq, ptr(xq_local), k, -1 if limit_rows is None else int(limit_rows), tie, ptr(s), ptr(l),
            _lib.current_stream_ptr()))
        return s, l

    def search_sharded(self, xq_local, k: int, limit_rows: int | None = None, tie: int = TIE_LOW_LABEL):
        """Host arrays in / out through ``cb_index_search_sharded`` (blocking; the call a C++ host makes)."""
        xq = np.ascontiguousarray(xq_local, dtype=np.float32).reshape(-1, self.d)
        nq = xq.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        S = np.empty((nq, k), dtype=np.float64)
        check(self.local._lib.cb_index_search_sharded(
            self.local._h, nq, ptr(xq), k, -1 if limit_rows is None else int(limit_rows), tie, ptr(D), ptr(I), ptr(S)))
        return D, I, S

    def add_gathered_device(self, n_rows: int):
        """Append the block the last ``search_sharded_device`` gathered (world * nq_local new descriptors in global,
        rank-major order) to the sharded DB: every rank keeps its own rows, no second exchange."""
        p = self.local._lib.cb_index_gathered_queries(self.local._h)
        check(self.local._lib.cb_index_add_device(self.local._h, n_rows, p, _lib.current_stream_ptr()))

    def search_device(self, xq, k: int, limit_rows: int | None = None, tie: int = TIE_LOW_LABEL):
        import torch

        s, l = self.local.search_device(xq, k, limit_rows, tie)
        if self.world == 1:
            return s, l
        gs = torch.empty((self.world,) + tuple(s.shape), dtype=s.dtype, device=s.device)
        gl = torch.empty((self.world,) + tuple(l.shape), dtype=l.dtype, device=l.device)
        self.dist.all_gather_into_tensor(gs, s, group=self.group)
        self.dist.all_gather_into_tensor(gl, l, group=self.group)
        return merge_topk_device(gs, gl, k, tie)


def merge_topk_host(scores: np.ndarray, labels: np.ndarray, k_out: int, tie: int = TIE_LOW_LABEL):
    """Host restatement of the merge rule used after the all-gather (exercised by the gloo
    world_size-2 CPU tests of the sharding logic; the product path merges on the device).
    scores/labels: [n_lists, nq, k_in]."""
    n_lists, nq, k_in = scores.shape
    S = np.full((nq, k_out), -np.inf)
    L = np.full((nq, k_out), -1, dtype=np.int64)
    for q in range(nq):
        s = scores[:, q, :].reshape(-1)
        l = labels[:, q, :].reshape(-1)
        keep = l >= 0
        s, l = s[keep], l[keep]
        order = np.lexsort(((-l if tie == TIE_HIGH_LABEL else l), -s))[:k_out]
        S[q, : len(order)] = s[order]
        L[q, : len(order)] = l[order]
    return S, L
