// Host-side pieces of cv::ORB's pipeline shared by the CUDA orchestration (csrc/features.cu) and the CPU emulation
// (host/orb_emul.cpp): pyramid geometry, the INTER_LINEAR_EXACT coefficient tables, the per-level feature budget, the
// umax table of the circular patch and KeyPointsFilter::retainBest.  Plain C++ (no CUDA).  OpenCV file references:
// modules/features2d/src/orb.cpp (ORB_Impl::detectAndCompute, computeKeyPoints), keypoint.cpp (retainBest),
// modules/imgproc/src/resize.cpp (interpolationLinear::getCoeffs).
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "orb_core.h"

namespace orb {

struct Level {
  int w = 0, h = 0;
  size_t off = 0;   // pixel offset of the level inside one image's pyramid buffer
  float scale = 1;  // layerScale
};

// getScale(level, 0, 1.2f) and Size(cvRound(cols / scale), cvRound(rows / scale)); returns the pixels of one pyramid
inline size_t level_geometry(int rows, int cols, Level lv[kLevels]) {
  const double scale_factor = (double)1.2f;
  size_t off = 0;
  for (int l = 0; l < kLevels; ++l) {
    lv[l].scale = (float)pow(scale_factor, (double)l);
    lv[l].w = (int)lrintf((float)cols / lv[l].scale);
    lv[l].h = (int)lrintf((float)rows / lv[l].scale);
    lv[l].off = off;
    off += (size_t)lv[l].w * lv[l].h;
  }
  return off;
}

// interpolationLinear<uint8_t>::getCoeffs for every destination index: source offset + 8.8 weight of the next sample;
// destinations below `lo` / from `hi` on replicate the first / last source sample
inline void lin_coeffs(int src, int dst, std::vector<int>& ofs, std::vector<int>& c1, int& lo, int& hi) {
  const double inv = (double)dst / (double)src;
  const double scale = 1.0 / inv;
  ofs.assign((size_t)dst, 0);
  c1.assign((size_t)dst, 0);
  lo = 0;
  hi = dst;
  for (int v = 0; v < dst; ++v) {
    const double fval = scale * ((double)v + 0.5) - 0.5;
    const int ival = (int)floor(fval);
    if (ival >= 0 && src > 1) {
      if (ival < src - 1) {
        ofs[(size_t)v] = ival;
        c1[(size_t)v] = (int)lrint((fval - (double)ival) * 256.0);
      } else {
        ofs[(size_t)v] = src - 1;
        hi = std::min(hi, v);
      }
    } else {
      lo = std::max(lo, v + 1);
    }
  }
}

inline void features_per_level(int nfeatures, int out[kLevels]) {
  const float factor = (float)(1.0 / (double)1.2f);
  float nd = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)kLevels));
  int sum = 0;
  for (int l = 0; l < kLevels - 1; ++l) {
    out[l] = (int)lrintf(nd);
    sum += out[l];
    nd *= factor;
  }
  out[kLevels - 1] = std::max(nfeatures - sum, 0);
}

inline void umax_table(int umax[kHalfPatch + 2]) {
  for (int i = 0; i < kHalfPatch + 2; ++i) umax[i] = 0;
  const int vmax = (int)floorf((float)kHalfPatch * sqrtf(2.f) / 2 + 1);
  const int vmin = (int)ceilf((float)kHalfPatch * sqrtf(2.f) / 2);
  for (int v = 0; v <= vmax; ++v) umax[v] = (int)lrint(sqrt((double)kHalfPatch * kHalfPatch - (double)v * v));
  for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
    while (umax[v0] == umax[v0 + 1]) ++v0;
    umax[v] = v0;
    ++v0;
  }
}

// KeyPointsFilter::retainBest on (response, index) records: the order std::nth_element / std::partition leave behind is
// part of cv::ORB's output order, and it is libstdc++'s -- the same routine OpenCV calls.
struct Rec {
  float response;
  int idx;
};
inline void retain_best(std::vector<Rec>& k, int n_points) {
  if (n_points >= 0 && k.size() > (size_t)n_points) {
    if (n_points == 0) {
      k.clear();
      return;
    }
    std::nth_element(k.begin(), k.begin() + n_points - 1, k.end(), [](const Rec& a, const Rec& b) { return a.response > b.response; });
    const float ambiguous = k[(size_t)n_points - 1].response;
    auto new_end = std::partition(k.begin() + n_points, k.end(), [ambiguous](const Rec& a) { return a.response >= ambiguous; });
    k.resize((size_t)(new_end - k.begin()));
  }
}

}  // namespace orb
