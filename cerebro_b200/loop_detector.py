"""ROS-free mirror of the loop-closure brain's hot path (src/Cerebro.{h,cpp}).

``Cerebro`` keeps the method names other reference code calls (Cerebro.h:66-233):
``wholeImageComputedList_size/_at``, ``foundLoops_count/_i/_as_JSON``, ``processedLoops_count/_i``
and the three thread bodies, here as explicit ``*_step`` functions a host loop (or the reference's
own threads) would call:

  descriptor_computer_thread   (Cerebro.cpp:47-303)   -> descriptor_step(stamps, images)
  descrip_N__dot__descrip_0_N  (Cerebro.cpp:903-1103) -> run_step()
  loopcandiate_consumer_thread (Cerebro.cpp:1185-1281)-> loopcandidate_consumer_step(correspondences)

All arithmetic runs behind the C ABI (descriptor / index / pnp handles); this file is bookkeeping.
``LoopPipeline`` is the batched throughput form used by bench.py: B keyframes per step through
desc -> search -> PnP on one GPU, optionally with the DB sharded over the ranks of a process group.
"""
from __future__ import annotations

import json
import math

import numpy as np

from .descriptor import NetvladDescriptor
from .index import TIE_HIGH_LABEL, TIE_LOW_LABEL, IndexFlatIP, ShardedIndex
from .pnp import PnpBatch, default_params


class LoopEdge:
    """cerebro::LoopEdge (msg/LoopEdge.msg:1-5)."""

    def __init__(self, timestamp0, timestamp1, pose_1T0, weight, description):
        self.timestamp0, self.timestamp1 = timestamp0, timestamp1
        self.pose_1T0 = pose_1T0  # 4x4; the ROS shim converts with eigenmat_to_geometry_msgs_Pose
        self.weight = float(weight)
        self.description = description


def R2ypr(R: np.ndarray) -> np.ndarray:
    """PoseManipUtils::R2ypr (src/utils/PoseManipUtils.cpp:148-163): yaw, pitch, roll in DEGREES."""
    import math

    n, o, a = R[:, 0], R[:, 1], R[:, 2]
    y = math.atan2(n[1], n[0])
    p = math.atan2(-n[2], n[0] * math.cos(y) + n[1] * math.sin(y))
    r = math.atan2(a[0] * math.sin(y) - a[1] * math.cos(y), -o[0] * math.sin(y) + o[1] * math.cos(y))
    return np.array([y, p, r]) * (180.0 / math.pi)


class ProcessedLoopCandidate:
    """The slice of src/ProcessedLoopCandidate.{h,cpp} that consumes the verifier's output: the three poses
    (Option A PnP, Option B PnP with roles swapped and inverted, Option C 3D-3D ICP), their goodness values
    and the decision whether a LoopEdge is published."""

    def __init__(self, idx_from_raw_candidates_list, t_1, t_2, idx_1=-1, idx_2=-1):
        self.idx_from_raw_candidates_list = idx_from_raw_candidates_list
        self.t_1, self.t_2 = t_1, t_2  # node_1->getT(), node_2->getT() in seconds
        self.idx_from_datamanager_1, self.idx_from_datamanager_2 = idx_1, idx_2
        self.opX_b_T_a = []
        self.opX_goodness = []
        self.pf_matches = 0
        self.isSet_3d2d__2T1 = False
        self._3d2d__2T1 = None
        self._3d2d__2T1__ransac_confidence = 0.0

    def makeLoopEdgeMsg(self):  # ProcessedLoopCandidate.cpp:16-36
        if not self.isSet_3d2d__2T1:
            return None
        desc = "%d<=>%d    this pose is: %d_T_%d" % (
            self.idx_from_datamanager_1, self.idx_from_datamanager_2, self.idx_from_datamanager_2, self.idx_from_datamanager_1)
        return LoopEdge(self.t_1, self.t_2, self._3d2d__2T1, self._3d2d__2T1__ransac_confidence, desc)

    def makeLoopEdgeMsgWithConsistencyCheck(self):  # ProcessedLoopCandidate.cpp:40-125
        if len(self.opX_b_T_a) != 3:
            return None
        # :49-56 -- ros::Duration::sec is floor-normalised (nsec >= 0): -9.5 s reads as sec = -10
        if abs(math.floor(self.t_1 - self.t_2)) < 10:
            return None
        op1, op2, icp = self.opX_b_T_a
        op1_m_op2 = np.linalg.inv(op1) @ op2
        op1_m_icp = np.linalg.inv(op1) @ icp
        op2_m_icp = np.linalg.inv(op2) @ icp
        is_consistent_ypr = (
            np.abs(R2ypr(op1_m_op2[:3, :3])).max() < 5.0
            and np.abs(R2ypr(op1_m_icp[:3, :3])).max() < 5.0
            and np.abs(R2ypr(op2_m_icp[:3, :3])).max() < 5.0
        )  # :77-81
        # :83-87 -- the reference tests op1-icp twice and never the op1-op2 translation; kept as is
        is_consistent_tr = np.abs(op1_m_icp[:3, 3]).max() < 0.2 and np.abs(op1_m_icp[:3, 3]).max() < 0.2 and np.abs(op2_m_icp[:3, 3]).max() < 0.2
        if self.pf_matches > 800 and is_consistent_ypr and is_consistent_tr:  # :110
            self._3d2d__2T1 = self.opX_b_T_a[0]
            self.isSet_3d2d__2T1 = True
            self._3d2d__2T1__ransac_confidence = max(self.opX_goodness)
            return self.makeLoopEdgeMsg()
        return None


def consistent_pose_compute(fe, pnp, K, img3d_a, img3d_b, match_results, stamps, node_indices=None, seed=0):
    """``Cerebro::process_loop_candidate_imagepair_consistent_pose_compute`` (src/Cerebro.cpp:1414-1771) for a BATCH of
    loop candidates, from the point where both frames' ORB features and stereo 3-D images exist.

    ``fe`` is a ``FrontEnd`` on which ``match_gms`` has just been called for the batch (``match_results`` = its return
    value); img3d_a / img3d_b [n, H, W, 3] float32; stamps[p] = (t_a, t_b) in seconds; node_indices[p] = (idx_1, idx_2).
    Per candidate, as the reference: reject when fewer than 150 GMS matches (:1484-1491); Option A
    ``PNP(3d(a), uv(b)) -> b_T_a`` (:1514-1522), Option B ``PNP(3d(b), uv(a)) -> a_T_b``, inverted (:1562-1575), Option C
    ``P3P_ICP(3d(a), 3d(b)) -> b_T_a`` (:1617-1626); reject on NaN (:1672); then
    ``ProcessedLoopCandidate::makeLoopEdgeMsgWithConsistencyCheck``.  The three solves of the whole batch are two device
    calls (one PnP batch of 2n candidates, one ICP batch of n).  Returns a list of (ProcessedLoopCandidate | None,
    LoopEdge | None)."""
    n = len(match_results)
    sets_a = fe.make_3d_2d_collection(K, img3d_a)                 # (uv_a, uv_b, X_a)
    sets_b = fe.make_3d_2d_collection(K, img3d_b, swapped=True)   # (uv_a, uv_b, X_b)
    sets_c = fe.make_3d_3d_collection(img3d_a, img3d_b)           # (X_a, Y_b)
    r_pnp = pnp.solve([sa[2] for sa in sets_a] + [sb[2] for sb in sets_b], [sa[1] for sa in sets_a] + [sb[0] for sb in sets_b],
                      default_params(seed=seed))
    r_icp = pnp.icp([sc[0] for sc in sets_c], [sc[1] for sc in sets_c], default_params(seed=seed + 1, error_thresh=0.1))
    out = []
    for p in range(n):
        pf = int(match_results[p]["n_inliers"])
        if pf < 150:  # "too few gms matches ... rejecting this loopcandidate"
            out.append((None, None))
            continue
        # A solver that refused (< 20 valid-depth points: confidence -1, DlsPnpWithRansac.cpp:136-139) or found no model
        # (best_hyp < 0) leaves the reference's output matrix uninitialised, which then fails its NaN / consistency test;
        # the device returns identity for both cases, so they are rejected here explicitly.
        solved = all(
            float(r["confidence"][q]) >= 0.0 and int(r["best_hyp"][q]) >= 0
            for r, q in ((r_pnp, p), (r_pnp, n + p), (r_icp, p))
        )
        if not solved:
            out.append((None, None))
            continue
        op1 = r_pnp["T"][p]
        op2 = np.linalg.inv(r_pnp["T"][n + p])
        icp = r_icp["T"][p]
        if np.isnan(op1).any() or np.isnan(op2).any() or np.isnan(icp).any():
            out.append((None, None))
            continue
        i1, i2 = node_indices[p] if node_indices is not None else (-1, -1)
        cand = ProcessedLoopCandidate(p, stamps[p][0], stamps[p][1], i1, i2)
        cand.pf_matches = pf
        cand.opX_b_T_a = [op1, op2, icp]
        cand.opX_goodness = [float(r_pnp["confidence"][p]), float(r_pnp["confidence"][n + p]), float(r_icp["confidence"][p])]
        cand.opX_b_T_a_name = ["op1__b_T_a", "op2__b_T_a", "icp_b_T_a"]
        out.append((cand, cand.makeLoopEdgeMsgWithConsistencyCheck()))
    return out


def process_loop_candidates_from_raw_stereo(feat, fe, pnp, K, Q, left_a, right_a, left_b, right_b, stamps, node_indices=None,
                                            warp_slots=None, n_orb_feat=5000, seed=0):
    """``Cerebro::process_loop_candidate_imagepair_consistent_pose_compute`` (src/Cerebro.cpp:1414-1771) for a batch of loop
    candidates FROM THE RAW STEREO PAIRS, every stage on the device:

      raw images --cv::remap x2--> stereo-rectified (StereoGeometry, CameraGeometry.cpp:42, 381-382; ``feat.remap``)
      --StereoBM(64, 21)--> disparity --disparity_to_3DPoints--> the two 3-D images (:81, 410-418, 459-520; ``fe.stereo_bm``)
      left images --cv::ORB(n_orb_feat, FAST threshold 0)--> keypoints + descriptors (PointFeatureMatching.cpp:16-22; ``feat.orb``)
      --BFMatcher(HAMMING) + GMS--> matches (:40-52; ``fe.match_gms``) --> Options A / B / C + consistency check
      (``consistent_pose_compute``).

    left_a / right_a / left_b / right_b: [n, rows, cols] uint8.  warp_slots = ((undistort_left, rectify_left),
    (undistort_right, rectify_right)) slot numbers of ``feat.set_remap``, or None when the images are already rectified.
    K: 3x3 intrinsics of the rectified left camera, Q: 4x4 reprojection matrix (both from cv::stereoRectify, host set-up).
    Returns the list ``consistent_pose_compute`` returns plus the per-candidate match results."""
    la, ra, lb, rb = (np.ascontiguousarray(x, dtype=np.uint8) for x in (left_a, right_a, left_b, right_b))
    n = la.shape[0]
    if warp_slots is not None:
        (ul, rl), (ur, rr) = warp_slots
        lefts = np.concatenate([feat.remap(la[i : i + feat.max_images], ul, rl) for i in range(0, n, feat.max_images)] +
                               [feat.remap(lb[i : i + feat.max_images], ul, rl) for i in range(0, n, feat.max_images)])
        rights = np.concatenate([feat.remap(ra[i : i + feat.max_images], ur, rr) for i in range(0, n, feat.max_images)] +
                                [feat.remap(rb[i : i + feat.max_images], ur, rr) for i in range(0, n, feat.max_images)])
    else:
        lefts, rights = np.concatenate([la, lb]), np.concatenate([ra, rb])
    disp = fe.stereo_bm(lefts, rights)          # [2n, rows, cols] int16
    img3d = fe.disparity_to_3d(disp, Q)         # [2n, rows, cols, 3] float32
    orb = []
    for i in range(0, 2 * n, feat.max_images):
        orb += feat.orb(lefts[i : i + feat.max_images], n_orb_feat)
    rows, cols = la.shape[1:3]
    kp1, d1 = [orb[p]["pt"] for p in range(n)], [orb[p]["desc"] for p in range(n)]
    kp2, d2 = [orb[n + p]["pt"] for p in range(n)], [orb[n + p]["desc"] for p in range(n)]
    matches = fe.match_gms(kp1, d1, kp2, d2, (cols, rows), (cols, rows))
    out = consistent_pose_compute(fe, pnp, K, img3d[:n], img3d[n:], matches, stamps, node_indices, seed=seed)
    return out, matches


# cv::cvtColor(CV_BGR2GRAY) on 8-bit images is fixed point: OpenCV 4 uses 15 bits, (3735 B + 19235 G + 9798 R + 2^14) >> 15
# (bit-exact against the installed cv2, tests/test_host_logic.py); OpenCV 3 -- the reference's era -- used 14 bits,
# (1868 B + 9617 G + 4899 R + 2^13) >> 14, which differs by one grey level on ~0.2 % of the pixels.
_BGR2GRAY = {15: (3735, 19235, 9798), 14: (1868, 9617, 4899)}


def convert_channels(images_u8: np.ndarray, n_channels: int, fixed_point_bits: int = 15) -> np.ndarray:
    """The channel fix-up the descriptor thread applies before the service call (Cerebro.cpp:229-234): a 1-channel image
    for a 3-channel model goes through ``cv::cvtColor(CV_GRAY2BGR)`` (replicate), a 3-channel (BGR) image for a
    1-channel model through ``cv::cvtColor(CV_BGR2GRAY)`` (fixed point, see above) -- anything else passes through.
    [n, rows, cols(, c)] uint8 in, [n, rows, cols, n_channels] out."""
    x = np.asarray(images_u8)
    if x.ndim == 3:
        x = x[..., None]
    assert x.dtype == np.uint8 and x.ndim == 4
    c = x.shape[3]
    if n_channels == 3 and c == 1:
        return np.ascontiguousarray(np.repeat(x, 3, axis=3))
    if n_channels == 1 and c == 3:
        cb, cg, cr = _BGR2GRAY[fixed_point_bits]
        b, g, r = (x[..., i].astype(np.int64) for i in range(3))
        return ((cb * b + cg * g + cr * r + (1 << (fixed_point_bits - 1))) >> fixed_point_bits).astype(np.uint8)[..., None]
    return np.ascontiguousarray(x)


class Hypothesis:
    """HypothesisManager.h:26-131: a chain of (a, b, dot product) nodes with a time-to-live."""

    def __init__(self, a, b, prod):
        self.list_of_nodes_in_this_hypothesis = [(a, b, prod)]
        self.time_to_live = 20  # :32

    def decrement_ttl(self):  # :103-108
        if self.time_to_live <= 0:
            return
        self.time_to_live -= 1

    def increment_ttl(self):  # :110-121
        self.time_to_live += 1
        if self.time_to_live > 100:
            self.time_to_live += 1

    def get_ttl(self):
        return self.time_to_live

    def is_hypothesis_active(self):
        return self.time_to_live > 0

    def n_elements_in_list(self):
        return len(self.list_of_nodes_in_this_hypothesis)


class HypothesisManager:
    """HypothesisManager.cpp:15-87 (the monitoring thread that prints to /dev/pts/1 is not part of the path)."""

    def __init__(self):
        self.active_hyp = []

    def add_node(self, a, b, dot_prod):
        for h in self.active_hyp:
            for (_a, _b, _) in reversed(h.list_of_nodes_in_this_hypothesis):  # :42-58, newest node first
                if abs(a - _a) < 7 and abs(b - _b) < 7:
                    h.list_of_nodes_in_this_hypothesis.append((a, b, dot_prod))
                    h.increment_ttl()
                    return True
        self.active_hyp.append(Hypothesis(a, b, dot_prod))  # :65-68
        return True

    def digest(self):  # :74-87: four decrements per searched descriptor
        for h in self.active_hyp:
            for _ in range(4):
                h.decrement_ttl()


class Cerebro:
    LOCALITY_THRESH = 12  # Cerebro.cpp:912
    DOT_PROD_THRESH = 0.85  # Cerebro.cpp:913
    LAG = 50  # Cerebro.cpp:914

    def __init__(self, descriptor: NetvladDescriptor, capacity: int = 29000, device: int = 0):
        self.descriptor = descriptor
        self.descriptor_size = descriptor.dim  # learnt by the probe call in the reference (Cerebro.cpp:113-120)
        self.index = IndexFlatIP(self.descriptor_size, capacity=capacity, device=device)
        self.pnp = PnpBatch(max_candidates=16, max_points_total=16 * 5000, max_hypotheses=50, device=device)
        self._whole = []  # wholeImageComputedList (Cerebro.h:101-106): stamps in arrival order
        self._found = []  # foundLoops (Cerebro.h:157-158): (t_curr, t_prev, score)
        self._processed = []  # processedloopcandi_list (Cerebro.h:199-200)
        self._last_l = 0
        # dynamic-skip state of descriptor_computer_thread (Cerebro.cpp:117, 166-170)
        self.estimated_descriptor_compute_time_ms = 0
        self._last_proc_timestamp = 0.0  # ros::Time()
        self._n_considered = 0
        self._last_consumed = 0
        self.hyp_manager = HypothesisManager()  # faiss_multihypothesis_tracking's (Cerebro.cpp:757)

    # ---- wholeImageComputedList_* (Cerebro.cpp:305-330)
    def wholeImageComputedList_size(self):
        return len(self._whole)

    def wholeImageComputedList_at(self, k):
        return self._whole[k]

    # ---- foundLoops_* (Cerebro.cpp:1113-1164)
    def foundLoops_count(self):
        return len(self._found)

    def foundLoops_i(self, i):
        return self._found[i]

    def foundLoops_as_JSON(self):
        out = []
        for i, (a, b, s) in enumerate(self._found):
            out.append({"time_sec_a": a, "time_sec_b": b, "dotprodt": s, "global_a": self._whole.index(a), "global_b": self._whole.index(b), "count": i})
        return json.dumps(out)

    def processedLoops_count(self):
        return len(self._processed)

    def processedLoops_i(self, i):
        return self._processed[i]

    # ---- resume / persist (DataManager::loadStateFromDisk + Cerebro.cpp:128-160; DataManager::saveStateToDisk)
    def load_state(self, save_folder_name):
        """Re-list every stored descriptor (ascending stamp) into wholeImageComputedList and bulk-load the rows into
        the device index with one call.  Stamps are the nodes' ``stampNSec``.  Returns the number of rows loaded."""
        from . import state_io

        stamps, descs, _ = state_io.load_state(save_folder_name)
        if not stamps:
            return 0
        if descs.shape[1] != self.descriptor_size:
            raise ValueError("state.json holds %d-D descriptors, the model produces %d-D" % (descs.shape[1], self.descriptor_size))
        self.index.add(np.ascontiguousarray(descs, dtype=np.float64))  # cb_index_add_f64: narrowed to fp32 on the device
        self._whole.extend(stamps)
        self._last_l = len(self._whole)  # nothing already listed is re-searched as "new"
        return len(stamps)

    def save_state(self, save_folder_name):
        """state.json with one data node per listed keyframe, descriptors read back from the device index."""
        from . import state_io

        n = len(self._whole)
        rows = self.index.get_rows(0, n) if n else np.zeros((0, self.descriptor_size), dtype=np.float32)
        return state_io.save_state(save_folder_name, self._whole, rows.astype(np.float64))

    # ---- descriptor_computer_thread body (Cerebro.cpp:169-298), for the keyframes handed in
    RAND_MAX = 2147483647

    def _dynamic_skip(self, stamp, rand):
        """The reference's "dynamic skip" (Cerebro.cpp:189-203): when keyframes arrive faster than a descriptor takes,
        drop them with probability ``1 - incoming_diff_ms / estimated_descriptor_compute_time_ms`` (after the first four;
        ``last_proc_timestamp`` advances for skipped keyframes too).  Stamps in seconds.  An estimate of 0 ms -- a batch
        that took less than a millisecond per keyframe -- gives skip_frac = -inf, i.e. never skip."""
        self._n_considered += 1
        incoming_diff_ms = min(int((stamp - self._last_proc_timestamp) * 1000.0), 2147483647)  # :193, int truncation
        est = np.float32(self.estimated_descriptor_compute_time_ms)
        with np.errstate(divide="ignore", invalid="ignore"):
            skip_frac = np.float32(1.0) - np.float32(incoming_diff_ms) / est  # :194
        self._last_proc_timestamp = stamp  # :196
        return self._n_considered > 4 and np.float32(rand() / np.float32(self.RAND_MAX)) < skip_frac  # :199

    def descriptor_step(self, stamps, images_u8, n_tracked=None, rand=None):
        """images_u8 [n, rows, cols, chnls].  Per keyframe, in arrival order: the dynamic random skip when ``rand`` is given
        (a ``rand()`` stand-in returning 0..RAND_MAX, Cerebro.cpp:193-203; without it nothing is skipped -- one batched
        device call outruns any keyframe rate), then keyframes with < 20 tracked features are dropped
        (Cerebro.cpp:206-210).  Descriptors are appended to the device DB in arrival order;
        ``estimated_descriptor_compute_time_ms`` becomes the batch's wall time per keyframe (int ms, :281)."""
        import time

        keep = []
        for i in range(len(stamps)):
            if rand is not None and self._dynamic_skip(stamps[i], rand):
                continue
            if n_tracked is not None and n_tracked[i] < 20:
                continue
            keep.append(i)
        if not keep:
            return 0
        t0 = time.perf_counter()
        imgs = np.asarray(images_u8)[keep]
        n_ch = getattr(self.descriptor, "chnls", None)
        if n_ch in (1, 3):  # Cerebro.cpp:229-234: gray <-> BGR fix-up when the stored image and the model disagree
            imgs = convert_channels(imgs, n_ch)
        desc = self.descriptor.compute(np.ascontiguousarray(imgs))
        self.index.add(desc)
        for i in keep:
            self._whole.append(stamps[i])
        self.estimated_descriptor_compute_time_ms = int((time.perf_counter() - t0) * 1000.0 / len(keep))
        return len(keep)

    # ---- one wake-up of descrip_N__dot__descrip_0_N (Cerebro.cpp:956-1100)
    def run_step(self):
        l = self.wholeImageComputedList_size()
        if l - self._last_l < 3:  # :962
            return None
        found, prev, score, _ = self.index.naive_candidate(l, self.LAG, self.LOCALITY_THRESH, self.DOT_PROD_THRESH)
        self._last_l = l
        if found:
            edge = (self._whole[l - 1], self._whole[prev], score)  # :1078-1081
            self._found.append(edge)
            return edge
        return None

    # ---- one wake-up of faiss__naive_loopcandidate_generator (Cerebro.cpp:366-492, compiled under HAVE_FAISS)
    FAISS_LAG = 150  # Cerebro.cpp:374
    FAISS_DOT_PROD_THRESH = 0.9  # Cerebro.cpp:376

    def faiss_naive_step(self):
        """The FAISS-flavoured candidate rule on the same device index: descriptors enter the searchable set with
        a 150-keyframe lag (expressed as ``limit_rows``; the reference delays ``index.add`` instead, :415-433), every
        new descriptor is searched top-5 (:460), a loop is reported when exactly 3 new descriptors were searched, the
        newest top-1 score exceeds 0.9 and the three top-1 labels lie within LOCALITY_THRESH (:476)."""
        l = self.wholeImageComputedList_size()
        if l - self._last_l < 3:  # :403
            return None
        last_l = self._last_l
        self._last_l = l
        limit = l - self.FAISS_LAG
        if limit < 5:  # index.ntotal < 5 -> nothing is searched (:449)
            return None
        if l - last_l != 3:  # tmp_ only has 3 entries when exactly 3 descriptors arrived (:476 `_n == 3`)
            return None
        q = self.index.get_rows(last_l, 3)
        D, I = self.index.search(q, 5, limit_rows=limit, tie=TIE_LOW_LABEL)
        tmp, tmp_i = [float(D[j, 0]) for j in range(3)], [int(I[j, 0]) for j in range(3)]
        if (tmp[2] > np.float32(self.FAISS_DOT_PROD_THRESH) and abs(tmp_i[0] - tmp_i[1]) < self.LOCALITY_THRESH
                and abs(tmp_i[0] - tmp_i[2]) < self.LOCALITY_THRESH):
            edge = (self._whole[l - 1], self._whole[tmp_i[2]], tmp[2])  # :484
            self._found.append(edge)
            return edge
        return None

    # ---- shared by the two top-5 generators below: one wake-up's searches as ONE batched device call
    CLIQUE_LAG = 150  # start_adding_descriptors_to_index_after (Cerebro.cpp:513, :743)
    CLIQUE_K = 5
    CLIQUE_THRESH = 0.85
    CLIQUE_LOCALITY = 7
    CLIQUE_RESET = 4

    def _top5_of_new(self, last_l, l):
        """index.search of every descriptor in [last_l, l) against rows [0, l-150) (the reference adds with a 150-frame
        lag before searching, :559-581): the whole wake-up is one batched sweep.  None when ntotal < 5 (:600 break)."""
        limit = l - self.CLIQUE_LAG
        if limit < self.CLIQUE_K:
            return None
        q = self.index.get_rows(last_l, l - last_l)
        return self.index.search(q, self.CLIQUE_K, limit_rows=limit, tie=TIE_LOW_LABEL)

    # ---- one wake-up of faiss_clique_loopcandidate_generator (Cerebro.cpp:506-722)
    def faiss_clique_step(self, rand=None):
        """Top-5 neighbours above 0.85 vote into ``retained`` (label -> votes; the duplicate test is the reference's
        signed ``(key - label) < 7`` over the ascending std::map, first hit wins, :633-661); every 4th list index the
        accumulated cliques are pushed to foundLoops with score 0.9 -- all of them when there is one, otherwise each
        with probability 1/len via ``rand() % 100 < 100/len`` (:664-713; ``rand`` defaults to "always keep")."""
        rand = rand or (lambda: 0)
        l = self.wholeImageComputedList_size()
        last_l = self._last_l
        if l <= last_l:  # :544
            return []
        self._last_l = l
        res = self._top5_of_new(last_l, l)
        if res is None:
            return []
        D, I = res
        if not hasattr(self, "_retained"):
            self._retained = {}
        retained, out = self._retained, []
        for j, li in enumerate(range(last_l, l)):
            for g in range(self.CLIQUE_K):
                if D[j, g] < np.float32(self.CLIQUE_THRESH):
                    break
                dup = -1
                for key in sorted(retained):
                    if key - int(I[j, g]) < self.CLIQUE_LOCALITY:
                        dup = key
                        break
                if dup != -1:
                    retained[dup] += 1
                else:
                    retained[int(I[j, g])] = 1
            if retained and li % self.CLIQUE_RESET == 0:
                if len(retained) == 1:
                    keys = list(retained)
                else:
                    percent = int(100.0 / len(retained))
                    keys = [k for k in sorted(retained) if rand() % 100 < percent]
                for k in keys:
                    edge = (self._whole[l - 1], self._whole[k], 0.9)
                    self._found.append(edge)
                    out.append(edge)
                retained.clear()
        return out

    # ---- one wake-up of faiss_multihypothesis_tracking (Cerebro.cpp:731-885)
    def faiss_multihypothesis_step(self):
        l = self.wholeImageComputedList_size()
        last_l = self._last_l
        if l <= last_l:
            return self.hyp_manager
        self._last_l = l
        res = self._top5_of_new(last_l, l)
        if res is None:
            return self.hyp_manager
        D, I = res
        for j, li in enumerate(range(last_l, l)):
            for g in range(self.CLIQUE_K):
                if D[j, g] > np.float32(0.85):  # :857
                    self.hyp_manager.add_node(li, int(I[j, g]), float(D[j, g]))
            self.hyp_manager.digest()
        return self.hyp_manager

    # ---- loopcandiate_consumer_thread body (Cerebro.cpp:1203-1277), geometry supplied by the caller
    def loopcandidate_consumer_step(self, correspondences, params=None):
        """correspondences[j] = (w_X [n,3], c_uv [n,2]) for every not-yet-consumed foundLoops entry
        (the stereo/GMS front-end that produces them is outside this path)."""
        new = self._found[self._last_consumed :]
        assert len(correspondences) == len(new)
        if not new:
            return []
        r = self.pnp.solve([c[0] for c in correspondences], [c[1] for c in correspondences], params or default_params())
        out = []
        for j, (a, b, s) in enumerate(new):
            rec = dict(t_curr=a, t_prev=b, score=s, b_T_a=r["T"][j], goodness=float(r["confidence"][j]))
            self._processed.append(rec)
            out.append(rec)
        self._last_consumed = len(self._found)
        return out


class LoopPipeline:
    """B keyframes per step through desc -> search -> PnP, device-resident (bench.py)."""

    def __init__(self, net, rows, cols, chnls, batch, db_rows_local, device=0, sharded=False, n_corr=200, hypotheses=50):
        import torch

        self.torch = torch
        self.batch = batch
        self.desc = NetvladDescriptor(net, rows, cols, chnls, max_batch=batch, device=device)
        self.dim = self.desc.dim
        self.sharded = sharded
        if sharded:
            from .index import Comm

            self.index = ShardedIndex(self.dim, db_rows_local, device)
            self.world = self.index.world
            if self.world > 1:  # the collective search lives behind the C ABI (cb_index_search_sharded_device)
                self.comm = Comm.from_torch_group(device)
                self.index.attach_comm(self.comm)
        else:
            self.index = IndexFlatIP(self.dim, db_rows_local, device)
            self.world = 1
        self.pnp = PnpBatch(max_candidates=batch, max_points_total=batch * n_corr, max_hypotheses=hypotheses, device=device)
        self.params = default_params(max_iterations=hypotheses, seed=7)
        self.k = 5  # FAISS path searches top-5 (Cerebro.cpp:460)
        self._side = None  # second CUDA stream for the verifier

    def step_device(self, images_dev, offsets_dev, X_dev, uv_dev, bufs):
        """All inputs already in HBM.  Returns (labels [B,k] -- the merged global top-k of this rank's keyframes --,
        scores, pnp outputs).

        The verifier does not depend on the descriptor/search chain of the same step (the reference runs them
        in different threads, cerebro_node.cpp:487-509), so it is issued on a second CUDA stream: its
        latency-bound warp-per-hypothesis kernels fill the SMs' idle issue slots while the HBM-bound sweep runs."""
        torch = self.torch
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        self._side.wait_stream(main)
        with torch.cuda.stream(self._side):
            out = self.pnp.solve_device(offsets_dev, X_dev, uv_dev, self.params, out=bufs["pnp"])
        d = self.desc.compute_device(images_dev, out=bufs["desc"])
        if self.sharded and self.world > 1:
            # one C-ABI call: query all-gather -> shard sweep -> ONE ncclAllGather of the packed top-k -> merge
            s, l = self.index.search_sharded_device(d, self.k, out=bufs.get("search_out"))
        else:
            s, l = self.index.search_device(d, self.k)
        main.wait_stream(self._side)
        return l, s, out
