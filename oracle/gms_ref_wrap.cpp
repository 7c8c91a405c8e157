// ORACLE (test infrastructure): C entry point around the UNMODIFIED reference GMS matcher, compiled from
// /root/reference/src/utils/GMSMatcher/gms_matcher.cpp by oracle/Makefile into oracle/_ref/libgms_ref.so.
// Call shape = StaticPointFeatureMatching::gms_point_feature_matches (src/utils/PointFeatureMatching.cpp:50-52):
//   gms_matcher gms(kp1, size1, kp2, size2, matches_all); gms.GetInlierMask(vbInliers, false, false);
#include "gms_matcher.h"

extern "C" int gms_ref_inlier_mask(int n1, const float* kp1_xy, int w1, int h1, int n2, const float* kp2_xy, int w2, int h2,
                                   int n_matches, const int* query_idx, const int* train_idx, int with_scale,
                                   int with_rotation, unsigned char* mask_out) {
  std::vector<cv::KeyPoint> k1(n1), k2(n2);
  for (int i = 0; i < n1; ++i) k1[i].pt = cv::Point2f(kp1_xy[2 * i], kp1_xy[2 * i + 1]);
  for (int i = 0; i < n2; ++i) k2[i].pt = cv::Point2f(kp2_xy[2 * i], kp2_xy[2 * i + 1]);
  std::vector<cv::DMatch> m(n_matches);
  for (int i = 0; i < n_matches; ++i) {
    m[i].queryIdx = query_idx[i];
    m[i].trainIdx = train_idx[i];
  }
  gms_matcher gms(k1, cv::Size(w1, h1), k2, cv::Size(w2, h2), m);
  std::vector<bool> inl;
  const int n = gms.GetInlierMask(inl, with_scale != 0, with_rotation != 0);
  for (size_t i = 0; i < inl.size() && i < (size_t)n_matches; ++i) mask_out[i] = inl[i] ? 1 : 0;
  return n;
}
