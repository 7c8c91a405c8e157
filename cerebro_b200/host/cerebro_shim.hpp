// ROS-free C++ host shim over the C ABI (include/cerebro_b200.h).
//
// Mirrors the slice of the reference's class surface that sits on the loop-detection hot path, with the
// same method names and semantics, so that the three thread bodies of src/Cerebro.cpp can be pointed at
// libcerebro_b200.so by replacing only their compute calls (see INTEGRATION.md for the exact patch):
//
//   Cerebro::descriptor_computer_thread   src/Cerebro.cpp:47-303    -> cb_descriptor_compute
//   Cerebro::descrip_N__dot__descrip_0_N  src/Cerebro.cpp:903-1103  -> cb_index_add_f64 + cb_index_naive_candidate
//   Cerebro::loopcandiate_consumer_thread src/Cerebro.cpp:1185-1281 -> cb_pnp_solve_batch
//   StaticTheiaPoseCompute::PNP           src/DlsPnpWithRansac.h:169-179
//   ProcessedLoopCandidate::makeLoopEdgeMsgWithConsistencyCheck  src/ProcessedLoopCandidate.cpp:40-125 (host logic, no device call)
//   PoseManipUtils::{R2ypr, eigenmat_to_rawyprt, eigenmat_to_geometry_msgs_Pose}  src/utils/PoseManipUtils.cpp:31-43, 148-163
//
// ros::Time, DataNode and DataManager are reduced to what those bodies touch (stamp, keyframe flag,
// tracked-feature count, image, VectorXd descriptor as std::vector<double>).  No Eigen/OpenCV/ROS needed.
#pragma once
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <functional>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/cerebro_b200.h"

namespace cerebro_b200 {

struct Time {  // stands in for ros::Time (only ordering, toSec and nsec are used on the path)
  int64_t nsec = 0;
  double toSec() const { return 1e-9 * (double)nsec; }
  bool operator<(const Time& o) const { return nsec < o.nsec; }
  bool operator==(const Time& o) const { return nsec == o.nsec; }
};

class DataNode {  // src/DataNode.h:52-119, the members the hot path touches
 public:
  explicit DataNode(Time t) : stamp(t) {}
  Time getT() const { return stamp; }
  bool isKeyFrame() const { return is_keyframe; }
  int getNumberOfSuccessfullyTrackedFeatures() const { return n_tracked; }
  void setWholeImageDescriptor(const std::vector<double>& vec) {  // src/DataNode.cpp:427-433
    std::lock_guard<std::mutex> lk(m);
    desc = vec;
    desc_available = true;
  }
  std::vector<double> getWholeImageDescriptor() {  // by-value copy, src/DataNode.cpp:435-440
    std::lock_guard<std::mutex> lk(m);
    return desc;
  }
  bool isWholeImageDescriptorAvailable() const { return desc_available; }
  // test-harness fields
  bool is_keyframe = true;
  int n_tracked = 100;
  std::vector<uint8_t> left_image;  // rows*cols*chnls, what ImageDataManager::getImage("left_image") returns

 private:
  Time stamp;
  std::mutex m;
  std::vector<double> desc;
  std::atomic<bool> desc_available{false};
};

class DataManager {  // src/DataManager.h:96-117: only getDataMapRef is used by the path
 public:
  std::map<Time, DataNode*>* getDataMapRef() { return &data_map; }
  ~DataManager() {
    for (auto& kv : data_map) delete kv.second;
  }
  std::map<Time, DataNode*> data_map;
};

class StaticTheiaPoseCompute {  // src/DlsPnpWithRansac.h:169-179
 public:
  // w_X: n x 3, c_uv_normalized: n x 2 (row-major doubles); c_T_w: 16 doubles row-major 4x4.
  // Returns summary.confidence, or -1 when fewer than 20 points are supplied (DlsPnpWithRansac.cpp:136-139).
  static float PNP(cb_pnp* solver, const std::vector<double>& w_X, const std::vector<double>& c_uv_normalized,
                   double* c_T_w, std::string& pnp__msg, uint64_t seed = 0) {
    const int n = (int)(w_X.size() / 3);
    if (n < 20) return -1.f;
    pnp__msg = "";
    cb_ransac_params prm;
    cb_ransac_params_default(&prm);  // error_thresh 0.03, min_inlier_ratio 0.7, 50/5 iterations, MLE (:207-212)
    prm.seed = seed;
    int32_t off[2] = {0, n};
    float conf = -1.f;
    int32_t nit = 0, ninl = 0, bh = -1;
    const int rc = cb_pnp_solve_batch(solver, 1, off, w_X.data(), c_uv_normalized.data(), &prm, nullptr, c_T_w, &conf,
                                      &nit, &ninl, &bh);
    if (rc != CB_OK) {
      pnp__msg = std::string("cb_pnp_solve_batch failed: ") + cb_last_error();
      return -1.f;
    }
    std::ostringstream ss;
    ss << "DlsPnpWithRansac (best_rel_pose.b_T_a): num_iterations=" << nit << "  confidence=" << conf << ";";
    pnp__msg += ss.str();
    return conf;
  }
};

// ---- 4x4 pose algebra (row-major double[16]); stands in for the Eigen calls of the consistency check
struct Matrix4d {
  double m[16];
  double& operator()(int r, int c) { return m[4 * r + c]; }
  double operator()(int r, int c) const { return m[4 * r + c]; }
  static Matrix4d Identity() {
    Matrix4d I{};
    for (int i = 0; i < 4; ++i) I(i, i) = 1.0;
    return I;
  }
  static Matrix4d from(const double* p) {
    Matrix4d M;
    for (int i = 0; i < 16; ++i) M.m[i] = p[i];
    return M;
  }
  Matrix4d operator*(const Matrix4d& o) const {
    Matrix4d R{};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0.0;
        for (int k = 0; k < 4; ++k) s += (*this)(i, k) * o(k, j);
        R(i, j) = s;
      }
    return R;
  }
  // general inverse (Gauss-Jordan, partial pivoting), as Eigen's Matrix4d::inverse() is: the reference does not assume
  // that the solver returned a rigid transform
  Matrix4d inverse() const {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        a[i][j] = (*this)(i, j);
        a[i][4 + j] = i == j ? 1.0 : 0.0;
      }
    for (int k = 0; k < 4; ++k) {
      int p = k;
      for (int i = k + 1; i < 4; ++i)
        if (std::fabs(a[i][k]) > std::fabs(a[p][k])) p = i;
      if (p != k)
        for (int j = 0; j < 8; ++j) std::swap(a[k][j], a[p][j]);
      const double inv = 1.0 / a[k][k];
      for (int j = 0; j < 8; ++j) a[k][j] *= inv;
      for (int i = 0; i < 4; ++i)
        if (i != k) {
          const double f = a[i][k];
          for (int j = 0; j < 8; ++j) a[i][j] -= f * a[k][j];
        }
    }
    Matrix4d R;
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) R(i, j) = a[i][4 + j];
    return R;
  }
  bool hasNaN() const {
    for (double v : m)
      if (v != v) return true;
    return false;
  }
};

struct PoseManipUtils {
  // yaw, pitch, roll in DEGREES of R = Rz Ry Rx (src/utils/PoseManipUtils.cpp:148-163)
  static void R2ypr(const Matrix4d& T, double ypr[3]) {
    const double n0 = T(0, 0), n1 = T(1, 0), n2 = T(2, 0), o0 = T(0, 1), o1 = T(1, 1), a0 = T(0, 2), a1 = T(1, 2);
    const double y = std::atan2(n1, n0);
    const double p = std::atan2(-n2, n0 * std::cos(y) + n1 * std::sin(y));
    const double r = std::atan2(a0 * std::sin(y) - a1 * std::cos(y), -o0 * std::sin(y) + o1 * std::cos(y));
    ypr[0] = y / M_PI * 180.0, ypr[1] = p / M_PI * 180.0, ypr[2] = r / M_PI * 180.0;
  }
  static void eigenmat_to_rawyprt(const Matrix4d& T, double ypr[3], double t[3]) {
    R2ypr(T, ypr);
    t[0] = T(0, 3), t[1] = T(1, 3), t[2] = T(2, 3);
  }
  // position + orientation (x, y, z, w) = Eigen::Quaterniond(R): the branch on the trace and the largest diagonal element
  // that Eigen's rotation-matrix constructor uses (src/utils/PoseManipUtils.cpp:31-43)
  static void eigenmat_to_geometry_msgs_Pose(const Matrix4d& T, double position[3], double orientation_xyzw[4]) {
    position[0] = T(0, 3), position[1] = T(1, 3), position[2] = T(2, 3);
    double q[4];  // x y z w
    double t = T(0, 0) + T(1, 1) + T(2, 2);
    if (t > 0.0) {
      t = std::sqrt(t + 1.0);
      q[3] = 0.5 * t;
      t = 0.5 / t;
      q[0] = (T(2, 1) - T(1, 2)) * t;
      q[1] = (T(0, 2) - T(2, 0)) * t;
      q[2] = (T(1, 0) - T(0, 1)) * t;
    } else {
      int i = 0;
      if (T(1, 1) > T(0, 0)) i = 1;
      if (T(2, 2) > T(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(T(i, i) - T(j, j) - T(k, k) + 1.0);
      q[i] = 0.5 * t;
      t = 0.5 / t;
      q[3] = (T(k, j) - T(j, k)) * t;
      q[j] = (T(j, i) + T(i, j)) * t;
      q[k] = (T(k, i) + T(i, k)) * t;
    }
    for (int i = 0; i < 4; ++i) orientation_xyzw[i] = q[i];
  }
  static double linf3(const double v[3]) {
    const double a = std::fabs(v[0]), b = std::fabs(v[1]), c = std::fabs(v[2]);
    return a > b ? (a > c ? a : c) : (b > c ? b : c);
  }
};

struct LoopEdge {  // msg/LoopEdge.msg:1-5
  Time timestamp0, timestamp1;
  double position[3] = {0, 0, 0};
  double orientation_xyzw[4] = {0, 0, 0, 1};  // geometry_msgs/Pose pose_1T0
  float weight = 0.f;
  std::string description;
};

// The slice of src/ProcessedLoopCandidate.{h,cpp} that consumes the verifier's three poses (Option A PnP, Option B PnP with
// the roles swapped and inverted, Option C 3D-3D alignment) and decides whether a LoopEdge is published.
class ProcessedLoopCandidate {
 public:
  ProcessedLoopCandidate(int idx_from_raw_candidates_list_, Time t_1_, Time t_2_, int idx_1 = -1, int idx_2 = -1)
      : idx_from_raw_candidates_list(idx_from_raw_candidates_list_), t_1(t_1_), t_2(t_2_), idx_from_datamanager_1(idx_1),
        idx_from_datamanager_2(idx_2) {}

  bool makeLoopEdgeMsg(LoopEdge& msg) const {  // ProcessedLoopCandidate.cpp:16-36
    if (!isSet_3d2d__2T1) return false;
    msg.timestamp0 = t_1;
    msg.timestamp1 = t_2;
    PoseManipUtils::eigenmat_to_geometry_msgs_Pose(_3d2d__2T1, msg.position, msg.orientation_xyzw);
    msg.weight = _3d2d__2T1__ransac_confidence;
    msg.description = std::to_string(idx_from_datamanager_1) + "<=>" + std::to_string(idx_from_datamanager_2);
    msg.description += "    this pose is: " + std::to_string(idx_from_datamanager_2) + "_T_" + std::to_string(idx_from_datamanager_1);
    return true;
  }

  bool makeLoopEdgeMsgWithConsistencyCheck(LoopEdge& msg) {  // ProcessedLoopCandidate.cpp:40-125
    if (opX_b_T_a.size() != 3) return false;
    // :49-56 ros::Duration is floor-normalised (0 <= nsec < 1e9), so diff.sec = floor(diff)
    const int64_t dn = t_1.nsec - t_2.nsec;
    int64_t dsec = dn / 1000000000LL;
    if (dn % 1000000000LL < 0) --dsec;
    if ((dsec < 0 ? -dsec : dsec) < 10) return false;
    const Matrix4d &op1 = opX_b_T_a[0], &op2 = opX_b_T_a[1], &icp = opX_b_T_a[2];
    const Matrix4d op1_m_op2 = op1.inverse() * op2, op1_m_icp = op1.inverse() * icp, op2_m_icp = op2.inverse() * icp;
    double y12[3], t12[3], y1i[3], t1i[3], y2i[3], t2i[3];
    PoseManipUtils::eigenmat_to_rawyprt(op1_m_op2, y12, t12);
    PoseManipUtils::eigenmat_to_rawyprt(op1_m_icp, y1i, t1i);
    PoseManipUtils::eigenmat_to_rawyprt(op2_m_icp, y2i, t2i);
    const bool is_consistent_ypr =
        PoseManipUtils::linf3(y12) < 5.0 && PoseManipUtils::linf3(y1i) < 5.0 && PoseManipUtils::linf3(y2i) < 5.0;  // :77-81
    // :83-87 the reference tests op1-icp twice and never the op1-op2 translation; kept as is
    const bool is_consistent_tr = PoseManipUtils::linf3(t1i) < .2 && PoseManipUtils::linf3(t1i) < .2 && PoseManipUtils::linf3(t2i) < .2;
    (void)t12;
    if (pf_matches > 800 && is_consistent_ypr && is_consistent_tr) {  // :110
      _3d2d__2T1 = opX_b_T_a[0];
      isSet_3d2d__2T1 = true;
      _3d2d__2T1__ransac_confidence = std::max(std::max(opX_goodness[0], opX_goodness[1]), opX_goodness[2]);
      return makeLoopEdgeMsg(msg);
    }
    return false;
  }

  int idx_from_raw_candidates_list;
  Time t_1, t_2;  // node_1->getT(), node_2->getT()
  int idx_from_datamanager_1, idx_from_datamanager_2;
  std::vector<Matrix4d> opX_b_T_a;
  std::vector<float> opX_goodness;
  int pf_matches = 0;
  bool isSet_3d2d__2T1 = false;
  Matrix4d _3d2d__2T1 = Matrix4d::Identity();
  float _3d2d__2T1__ransac_confidence = 0.f;
};

class Cerebro {
 public:
  // constants of descrip_N__dot__descrip_0_N (src/Cerebro.cpp:912-914)
  static constexpr int LOCALITY_THRESH = 12;
  static constexpr float DOT_PROD_THRESH = 0.85f;
  static constexpr int start_adding_descriptors_to_index_after = 50;

  Cerebro(cb_descriptor* desc, int rows, int cols, int chnls, int64_t capacity = 29000 /* Cerebro.cpp:946 */, int device = 0)
      : desc_(desc), rows_(rows), cols_(cols), chnls_(chnls) {
    descriptor_size = cb_descriptor_dim(desc_);  // learnt by the probe call in the reference (:113-120)
    descriptor_size_available = descriptor_size > 0;
    ok_ = cb_index_create(&index_, descriptor_size, capacity, device, 0, 1) == CB_OK &&
          cb_pnp_create(&pnp_, 16, 16 * 5000, 50, device) == CB_OK;
  }
  ~Cerebro() {
    cb_index_destroy(index_);
    cb_pnp_destroy(pnp_);
  }
  bool ok() const { return ok_; }
  void setDataManager(DataManager* dm) {
    dataManager = dm;
    m_dataManager_available = true;
  }

  // ---- wholeImageComputedList (src/Cerebro.h:101-106, Cerebro.cpp:305-330)
  int wholeImageComputedList_size() {
    std::lock_guard<std::mutex> lk(m_wholeImageComputedList);
    return (int)wholeImageComputedList.size();
  }
  Time wholeImageComputedList_at(int k) {
    std::lock_guard<std::mutex> lk(m_wholeImageComputedList);
    return wholeImageComputedList.at(k);
  }

  // ---- foundLoops (src/Cerebro.h:152-158, Cerebro.cpp:1113-1124)
  int foundLoops_count() {
    std::lock_guard<std::mutex> lk(m_foundLoops);
    return (int)foundLoops.size();
  }
  std::tuple<Time, Time, double> foundLoops_i(int i) {
    std::lock_guard<std::mutex> lk(m_foundLoops);
    return foundLoops.at(i);
  }
  std::string foundLoops_as_JSON() {
    std::lock_guard<std::mutex> lk(m_foundLoops);
    std::ostringstream ss;
    ss << "[";
    for (size_t i = 0; i < foundLoops.size(); ++i) {
      ss << (i ? "," : "") << "{\"time_sec_a\":" << std::get<0>(foundLoops[i]).toSec()
         << ",\"time_sec_b\":" << std::get<1>(foundLoops[i]).toSec() << ",\"dotprodt\":" << std::get<2>(foundLoops[i]) << "}";
    }
    ss << "]";
    return ss.str();
  }
  int processedLoops_count() const { return (int)processed_.size(); }

  // ---- one pass of descriptor_computer_thread's loop body (src/Cerebro.cpp:169-298)
  // Every keyframe without a descriptor and newer than the last processed stamp is sent through the
  // descriptor handle (the reference does one blocking service call per keyframe, :263).
  int descriptor_computer_step() {
    if (!m_dataManager_available) return -1;
    auto* data_map = dataManager->getDataMapRef();
    int done = 0;
    for (auto& kv : *data_map) {
      DataNode* node = kv.second;
      if (!node->isKeyFrame() || node->isWholeImageDescriptorAvailable()) continue;  // :189
      if (last_processed_set_ && !(last_processed_ < kv.first)) continue;
      // :189-203 "dynamic skip": keyframes arriving faster than a descriptor takes are dropped at random.  With the
      // device path the estimate is 0-1 ms, so skip_frac <= 0 at any realistic keyframe rate and nothing is dropped.
      ++n_considered_;
      const double diff_ms = (kv.first.toSec() - (last_processed_set_ ? last_processed_.toSec() : 0.0)) * 1000.;
      const int incoming_diff_ms = diff_ms > 2147483647. ? 2147483647 : (int)diff_ms;  // first keyframe: ros::Time() = 0
      const float skip_frac = 1.0f - incoming_diff_ms / float(estimated_descriptor_compute_time_ms);
      last_processed_ = kv.first;  // advances for skipped keyframes too (:196)
      last_processed_set_ = true;
      if (dynamic_skip_enabled && n_considered_ > 4 && (rand_fn_() / float(RAND_MAX)) < skip_frac) continue;
      if (node->getNumberOfSuccessfullyTrackedFeatures() < 20) continue;  // :206-210
      const auto t_begin = std::chrono::steady_clock::now();
      // :229-234: the stored image may have the other channel count (CV_GRAY2BGR replicates, CV_BGR2GRAY is OpenCV 4's 8-bit
      // fixed point (3735 B + 19235 G + 9798 R + 2^14) >> 15; OpenCV 3 used (1868, 9617, 4899) >> 14, at most one level off)
      const size_t px = (size_t)rows_ * cols_;
      const uint8_t* img = node->left_image.data();
      std::vector<uint8_t> conv;
      if (node->left_image.size() == px && chnls_ == 3) {
        conv.resize(px * 3);
        for (size_t i = 0; i < px; ++i) conv[3 * i] = conv[3 * i + 1] = conv[3 * i + 2] = img[i];
        img = conv.data();
      } else if (node->left_image.size() == px * 3 && chnls_ == 1) {
        conv.resize(px);
        for (size_t i = 0; i < px; ++i)
          conv[i] = (uint8_t)((3735 * (int)img[3 * i] + 19235 * (int)img[3 * i + 1] + 9798 * (int)img[3 * i + 2] + 16384) >> 15);
        img = conv.data();
      } else if (node->left_image.size() != px * (size_t)chnls_) {
        continue;
      }
      std::vector<double> vec(descriptor_size);
      const int rc = cb_descriptor_compute_f64(desc_, 1, img, 0, vec.data());  // replaces client.call(srv), :263; the reply IS float64[] (:268-271)
      if (rc != CB_OK) {
        std::fprintf(stderr, "[descriptor_computer_thread] %s\n", cb_last_error());  // ROS_ERROR and continue, :288-290
        continue;
      }
      node->setWholeImageDescriptor(vec);               // :274
      {
        std::lock_guard<std::mutex> lk(m_wholeImageComputedList);
        wholeImageComputedList.push_back(kv.first);  // :275
      }
      estimated_descriptor_compute_time_ms =
          (int)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t_begin).count();  // :281
      ++done;
    }
    return done;
  }

  // ---- one wake-up of descrip_N__dot__descrip_0_N (src/Cerebro.cpp:956-1100)
  bool run_step() {
    auto* data_map = dataManager->getDataMapRef();
    const int l = wholeImageComputedList_size();
    if (l - last_l_ < 3) return false;  // :962
    // "Fill descriptors [last_l, l) into M" :1005-1012 -> device DB.  Rows are counted as they go in, so a failure part-way
    // (e.g. the index is full: the reference's M has 29000 columns) neither re-adds rows on the next wake-up nor lets
    // index labels drift from wholeImageComputedList; it is reported, not swallowed.
    for (; n_added_ < l; ++n_added_) {
      std::vector<double> v = data_map->at(wholeImageComputedList_at(n_added_))->getWholeImageDescriptor();
      if (cb_index_add_f64(index_, 1, v.data()) != CB_OK) {
        last_status_ = CB_EINVAL;
        std::fprintf(stderr, "[descrip_N__dot__descrip_0_N] row %d not added: %s\n", n_added_, cb_last_error());
        return false;
      }
    }
    int found = 0;
    int64_t prev = -1, am[3];
    double score = 0;
    const int rc = cb_index_naive_candidate(index_, l, start_adding_descriptors_to_index_after, LOCALITY_THRESH,
                                            DOT_PROD_THRESH, &found, &prev, &score, am);  // :1019-1056
    last_l_ = l;
    last_status_ = rc;
    if (rc != CB_OK) {
      std::fprintf(stderr, "[descrip_N__dot__descrip_0_N] %s\n", cb_last_error());
      return false;
    }
    if (!found) return false;
    std::lock_guard<std::mutex> lk(m_foundLoops);
    foundLoops.push_back(std::make_tuple(wholeImageComputedList_at(l - 1), wholeImageComputedList_at((int)prev), score));  // :1078-1081
    return true;
  }

  // ---- the PnP part of process_loop_candidate_imagepair_consistent_pose_compute (src/Cerebro.cpp:1518)
  float verify(const std::vector<double>& w_X, const std::vector<double>& uv, double* b_T_a, std::string& msg) {
    const float g = StaticTheiaPoseCompute::PNP(pnp_, w_X, uv, b_T_a, msg);
    processed_.push_back(g);
    return g;
  }

  int descriptor_size = -1;
  bool descriptor_size_available = false;
  int estimated_descriptor_compute_time_ms = 0;  // Cerebro.cpp:117, :281
  bool dynamic_skip_enabled = true;
  void set_rand(std::function<int()> f) { rand_fn_ = std::move(f); }
  int last_status() const { return last_status_; }  // CB_OK, or the error of the last run_step

 private:
  cb_descriptor* desc_ = nullptr;
  cb_index* index_ = nullptr;
  cb_pnp* pnp_ = nullptr;
  int rows_, cols_, chnls_;
  bool ok_ = false;
  DataManager* dataManager = nullptr;
  std::atomic<bool> m_dataManager_available{false};
  std::mutex m_wholeImageComputedList, m_foundLoops;
  std::vector<Time> wholeImageComputedList;
  std::vector<std::tuple<Time, Time, double>> foundLoops;
  std::vector<float> processed_;
  Time last_processed_;
  bool last_processed_set_ = false;
  int last_l_ = 0;
  int n_added_ = 0;       // rows of wholeImageComputedList already in the device index
  int last_status_ = 0;
  int n_considered_ = 0;  // n_computed of Cerebro.cpp:168
  std::function<int()> rand_fn_ = [] { return std::rand(); };
};

}  // namespace cerebro_b200
