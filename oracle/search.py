"""ORACLE (test infrastructure, not product): CPU restatement of the reference's
descriptor search / loop-candidate generation.

PARITY UNPINNED against reference golden vectors (the reference has none for
this path; FAISS is not installable here).  The arithmetic is a dot product and
an arg-max, so the restatement is exact up to floating-point summation order;
tests use inputs whose score gaps are far above that.

Follows:
  * Cerebro::descrip_N__dot__descrip_0_N           src/Cerebro.cpp:903-1103
      constants LOCALITY_THRESH 12, DOT_PROD_THRESH 0.85, lag 50   (:912-914)
      acts only when >= 3 new descriptors                           (:962)
      k = l - 50, requires k > 5                                    (:1019-1022)
      three fp64 GEMVs over M[:, :k]                                (:1026-1028)
      maxCoeff + linear scan keeping the LAST index equal to max    (:1035-1043)
      acceptance rule                                               (:1056)
      foundLoops.push_back( (t[l-1], t[argmax], max) )              (:1078-1081)
  * faiss::IndexFlatIP contract used at             src/Cerebro.cpp:390-460
      add(n, x) appends rows; search(nq, xq, k, D, I) returns the k largest inner
      products in descending order with labels = insertion index.
  * Cerebro::faiss__naive_loopcandidate_generator  src/Cerebro.cpp:366-492
      lag 150, LOCALITY 12, threshold 0.9, top-5 search per new descriptor.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module.
"""
from __future__ import annotations

import numpy as np

LOCALITY_THRESH = 12  # Cerebro.cpp:912
DOT_PROD_THRESH = 0.85  # Cerebro.cpp:913  (declared float in the reference)
LAG = 50  # Cerebro.cpp:914


def last_argmax(u: np.ndarray) -> int:
    """Cerebro.cpp:1039-1043: scan forward, keep the last index whose value == max."""
    m = u.max()
    return int(np.nonzero(u == m)[0][-1])


def naive_step(M: np.ndarray, l: int, lag: int = LAG):
    """One iteration of the reference loop body for list length ``l``.

    M: [D, >=l] float64 column-major in the reference; here any [D, n] array whose
    column j is descriptor j.  Returns None (nothing to do / rejected) or
    (curr=l-1, prev=argmax, score).
    """
    k = l - lag
    if k <= 5:
        return None
    v, vm, vmm = M[:, l - 1], M[:, l - 2], M[:, l - 3]
    Mk = M[:, :k]
    u, um, umm = v @ Mk, vm @ Mk, vmm @ Mk
    a, am, amm = last_argmax(u), last_argmax(um), last_argmax(umm)
    umax = u.max()
    if abs(a - am) < LOCALITY_THRESH and abs(a - amm) < LOCALITY_THRESH and umax > np.float32(DOT_PROD_THRESH):
        return (l - 1, a, float(umax))
    return None


def naive_stream(desc: np.ndarray, arrivals, lag: int = LAG):
    """Replay the thread: ``arrivals`` is the sequence of list lengths l seen at each
    10 Hz wake-up (strictly increasing).  desc: [n, D] float64 rows = descriptors in
    arrival order.  Returns the foundLoops list [(curr, prev, score), ...]."""
    M = np.ascontiguousarray(desc.T.astype(np.float64))
    found = []
    last_l = 0
    for l in arrivals:
        if l - last_l < 3:  # :962
            continue
        r = naive_step(M, l, lag)
        if r is not None:
            found.append(r)
        last_l = l
    return found


class IndexFlatIP:
    """The slice of faiss::IndexFlatIP the reference uses (Cerebro.cpp:390,431,460)."""

    def __init__(self, d: int):
        self.d = d
        self.x = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return self.x.shape[0]

    def add(self, x: np.ndarray) -> None:
        x = np.asarray(x, dtype=np.float32).reshape(-1, self.d)
        self.x = np.concatenate([self.x, x], axis=0)

    def search(self, xq: np.ndarray, k: int, limit_rows: int | None = None, accumulate: str = "float64"):
        """Descending inner products.  Ties: lower label first (stable).
        ``accumulate='float64'`` makes the oracle's ranking independent of summation
        order; the values returned are rounded to float32 like FAISS's."""
        xq = np.asarray(xq, dtype=np.float32).reshape(-1, self.d)
        n = self.ntotal if limit_rows is None else min(limit_rows, self.ntotal)
        D = np.full((xq.shape[0], k), -np.inf, dtype=np.float32)
        I = np.full((xq.shape[0], k), -1, dtype=np.int64)
        if n == 0:
            return D, I
        acc = np.float64 if accumulate == "float64" else np.float32
        s = xq.astype(acc, copy=False) @ self.x[:n].astype(acc, copy=False).T
        kk = min(k, n)
        for q in range(xq.shape[0]):
            order = np.argsort(-s[q], kind="stable")[:kk]
            D[q, :kk] = s[q, order].astype(np.float32)
            I[q, :kk] = order
        return D, I


def faiss_naive_stream(desc: np.ndarray, arrivals, lag: int = 150, thresh: float = 0.9):
    """Cerebro.cpp:366-492 replayed over the same arrival schedule."""
    index = IndexFlatIP(desc.shape[1])
    found = []
    last_l = 0
    added = 0
    for l in arrivals:
        if l - last_l < 3:  # :403
            continue
        if l > lag:  # :415
            if l - lag > added:
                index.add(desc[added : l - lag])
            added = max(added, l - lag)
        tmp, tmp_i = [], []
        for li in range(last_l, l):  # :441
            if index.ntotal < 5:  # :449
                continue
            D, I = index.search(desc[li], 5)
            tmp.append(float(D[0, 0]))
            tmp_i.append(int(I[0, 0]))
        n = len(tmp)
        if (
            n == 3
            and tmp[n - 1] > np.float32(thresh)
            and abs(tmp_i[0] - tmp_i[1]) < LOCALITY_THRESH
            and abs(tmp_i[0] - tmp_i[2]) < LOCALITY_THRESH
        ):  # :476
            found.append((l - 1, tmp_i[2], tmp[2]))
        last_l = l
    return found


# --------------------------------------------------------------------------------------------
# faiss_clique_loopcandidate_generator (Cerebro.cpp:506-722) and faiss_multihypothesis_tracking
# (Cerebro.cpp:731-885) + HypothesisManager (HypothesisManager.cpp:15-87, HypothesisManager.h:26-131),
# replayed over an arrival schedule.  Both are alternates of run() (Cerebro.cpp:352-355).
# --------------------------------------------------------------------------------------------
CLIQUE_LAG = 150  # start_adding_descriptors_to_index_after, :513
CLIQUE_K = 5  # :514
CLIQUE_THRESH = 0.85  # :515
CLIQUE_LOCALITY = 7  # :516
CLIQUE_RESET = 4  # reset_accumulation_every_n_frames, :517


def clique_accumulate(retained: dict, D, I) -> None:
    """:633-661.  ``retained`` maps label -> votes and is walked in ascending key order (std::map); the duplicate
    test is the reference's SIGNED difference ``(key - label) < LOCALITY`` (no abs), first hit wins."""
    for g in range(len(D)):
        if D[g] < np.float32(CLIQUE_THRESH):
            break
        dup = -1
        for key in sorted(retained):
            if key - int(I[g]) < CLIQUE_LOCALITY:
                dup = key
                break
        if dup != -1:
            retained[dup] += 1
        else:
            retained[int(I[g])] = 1


def faiss_clique_stream(desc: np.ndarray, arrivals, rand=None):
    """Returns foundLoops [(curr, prev, 0.9), ...].  ``rand`` stands in for libc ``rand()`` (:703); default: a
    generator that always returns 0 (every candidate retained)."""
    rand = rand or (lambda: 0)
    index = IndexFlatIP(desc.shape[1])
    found, retained = [], {}
    last_l = added = 0
    for l in arrivals:
        if l <= last_l:  # :544
            continue
        if l > CLIQUE_LAG:  # :559
            if l - CLIQUE_LAG > added:
                index.add(desc[added : l - CLIQUE_LAG])
            added = max(added, l - CLIQUE_LAG)
        for li in range(last_l, l):  # :593
            if index.ntotal < CLIQUE_K:  # :600 break
                break
            D, I = index.search(desc[li], CLIQUE_K)
            clique_accumulate(retained, D[0], I[0])
            if len(retained) > 0 and li % CLIQUE_RESET == 0:  # :664
                if len(retained) == 1:
                    found.append((l - 1, next(iter(retained)), 0.9))
                else:
                    percent = int(100.0 / len(retained))
                    for key in sorted(retained):
                        if rand() % 100 < percent:
                            found.append((l - 1, key, 0.9))
                retained.clear()
        last_l = l
    return found


class Hypothesis:  # HypothesisManager.h:26-131
    def __init__(self, a, b, prod):
        self.nodes = [(a, b, prod)]
        self.time_to_live = 20

    def decrement_ttl(self):
        if self.time_to_live > 0:
            self.time_to_live -= 1

    def increment_ttl(self):
        self.time_to_live += 1
        if self.time_to_live > 100:
            self.time_to_live += 1

    def is_hypothesis_active(self):
        return self.time_to_live > 0


class HypothesisManager:  # HypothesisManager.cpp:15-87
    def __init__(self):
        self.active_hyp = []

    def add_node(self, a, b, dot_prod):
        for h in self.active_hyp:  # note: expired hypotheses (ttl 0) stay in the list and still absorb nodes
            for (_a, _b, _) in reversed(h.nodes):
                if abs(a - _a) < 7 and abs(b - _b) < 7:
                    h.nodes.append((a, b, dot_prod))
                    h.increment_ttl()
                    return
        self.active_hyp.append(Hypothesis(a, b, dot_prod))

    def digest(self):
        for h in self.active_hyp:
            for _ in range(4):
                h.decrement_ttl()


def faiss_multihypothesis_stream(desc: np.ndarray, arrivals):
    """Returns the HypothesisManager after the stream (Cerebro.cpp:731-885)."""
    index = IndexFlatIP(desc.shape[1])
    hm = HypothesisManager()
    last_l = added = 0
    for l in arrivals:
        if l <= last_l:
            continue
        if l > CLIQUE_LAG:
            if l - CLIQUE_LAG > added:
                index.add(desc[added : l - CLIQUE_LAG])
            added = max(added, l - CLIQUE_LAG)
        for li in range(last_l, l):
            if index.ntotal < CLIQUE_K:
                break
            D, I = index.search(desc[li], CLIQUE_K)
            for g in range(CLIQUE_K):
                if D[0, g] > np.float32(0.85):  # :857
                    hm.add_node(li, int(I[0, g]), float(D[0, g]))
            hm.digest()
        last_l = l
    return hm
