// ORACLE helper (test infrastructure): KeyPointsFilter::retainBest (OpenCV features2d keypoint.cpp) on bare
// (response, index) records.  The permutation std::nth_element / std::partition leave behind is part of cv::ORB's
// observable output (keypoint order), and it is a property of libstdc++'s introselect, which both OpenCV and this helper
// call -- the records are moved by the same sequence of swaps as cv::KeyPoint structs would be.
#include <algorithm>
#include <cstdint>
#include <vector>

namespace {
struct Rec {
  float response;
  int32_t idx;
};
struct Greater {
  bool operator()(const Rec& a, const Rec& b) const { return a.response > b.response; }
};
struct GreaterEq {
  float v;
  bool operator()(const Rec& a) const { return a.response >= v; }
};
}  // namespace

extern "C" int orb_retain_best(const float* response, int n, int n_points, int32_t* out_idx) {
  std::vector<Rec> k((size_t)n);
  for (int i = 0; i < n; ++i) k[(size_t)i] = Rec{response[i], i};
  if (n_points >= 0 && k.size() > (size_t)n_points) {
    if (n_points == 0) return 0;
    std::nth_element(k.begin(), k.begin() + n_points - 1, k.end(), Greater());
    const float ambiguous = k[(size_t)n_points - 1].response;
    auto new_end = std::partition(k.begin() + n_points, k.end(), GreaterEq{ambiguous});
    k.resize((size_t)(new_end - k.begin()));
  }
  for (size_t i = 0; i < k.size(); ++i) out_idx[i] = k[i].idx;
  return (int)k.size();
}
