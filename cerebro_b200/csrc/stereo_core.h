// Per-thread bodies of the stereo block-matching kernels (cv::StereoBM with its default parameters, as the reference
// creates it at src/utils/CameraGeometry.cpp:81 and runs it at :410-418; algorithm: OpenCV modules/calib3d/src/stereobm.cpp)
// and of StereoGeometry::disparity_to_3DPoints (src/utils/CameraGeometry.cpp:459-520).
//
// The functions are plain scalar code marked CB_HD so that the SAME source runs inside the CUDA kernels (frontend.cu) and
// inside a g++-compiled emulation that walks the kernels' (block, thread) space on the CPU (host/stereo_emul.cpp,
// tests/test_stereo.py): indexing, border clamps, tie rules and the integer arithmetic are validated bit-exactly against
// the oracle without a GPU; the kernels add only the thread mapping and one shared-memory hand-off.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define CB_HD __host__ __device__ __forceinline__
#else
#define CB_HD inline
#endif

struct SbmGeom {
  int h, w;          // image size
  int ndisp, wsz;    // numDisparities (multiple of 16, <= 256), SADWindowSize (odd)
  int mindisp;       // minDisparity (0 in the reference)
  int cap;           // preFilterCap (31)
  int texture_threshold, uniqueness_ratio;  // 10, 15
  int wsz2, lofs, rofs, width1;             // derived, see sbm_make_geom
  int filtered;                             // (mindisp - 1) << 4
};

CB_HD SbmGeom sbm_make_geom(int h, int w, int ndisp, int wsz) {
  SbmGeom g;
  g.h = h, g.w = w, g.ndisp = ndisp, g.wsz = wsz;
  g.mindisp = 0, g.cap = 31, g.texture_threshold = 10, g.uniqueness_ratio = 15;
  g.wsz2 = wsz / 2;
  const int m = ndisp - 1 + g.mindisp;
  g.lofs = m > 0 ? m : 0;
  g.rofs = m < 0 ? -m : 0;
  g.width1 = w - g.rofs - ndisp + 1;
  g.filtered = (g.mindisp - 1) * 16;
  return g;
}

CB_HD int sbm_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// prefilterXSobel: clamp(dx(up) + 2 dx(y) + dx(down), -cap, cap) + cap, dx = I[x+1] - I[x-1].  OpenCV produces rows in
// pairs, which fixes the borders: above row 0 is row 1; below the last row of the image is the row above it; the
// last row of an odd-height image and the first / last column are `cap`.
CB_HD uint8_t sbm_prefilter_px(const uint8_t* img, int h, int w, int y, int x, int cap) {
  if (x <= 0 || x >= w - 1) return (uint8_t)cap;
  if ((h & 1) && y == h - 1) return (uint8_t)cap;
  if (h < 2) return (uint8_t)cap;
  const int up = y > 0 ? y - 1 : y + 1;
  const int down = y + 1 <= h - 1 ? y + 1 : y - 1;
  const uint8_t *r0 = img + (size_t)up * w, *r1 = img + (size_t)y * w, *r2 = img + (size_t)down * w;
  const int v = ((int)r0[x + 1] - (int)r0[x - 1]) + 2 * ((int)r1[x + 1] - (int)r1[x - 1]) + ((int)r2[x + 1] - (int)r2[x - 1]);
  return (uint8_t)(sbm_clampi(v, -cap, cap) + cap);
}

// |L - R| of window column c (relative to output column 0) for candidate d in row y of the pre-filtered images
CB_HD int sbm_diff(const uint8_t* PL, const uint8_t* PR, const SbmGeom& g, int y, int c, int d) {
  const int lc = g.lofs + sbm_clampi(c, -g.lofs, g.w - g.lofs - 1);
  const int rc = g.rofs + sbm_clampi(c, -g.rofs, g.w - g.rofs - g.ndisp) + d;
  const int v = (int)PL[(size_t)y * g.w + lc] - (int)PR[(size_t)y * g.w + rc];
  return v < 0 ? -v : v;
}
CB_HD int sbm_text(const uint8_t* PL, const SbmGeom& g, int y, int c) {
  const int lc = g.lofs + sbm_clampi(c, -g.lofs, g.w - g.lofs - 1);
  const int v = (int)PL[(size_t)y * g.w + lc] - g.cap;
  return v < 0 ? -v : v;
}

// Horizontal pass, one thread = (row y, candidate d, segment of output columns [x0, x1)): sliding window sums over
// columns x - wsz2 .. x + wsz2.  hsad [h][width1][ndisp] u16, htext [h][width1] (written by the d == 0 thread).
CB_HD void sbm_hsad_thread(const uint8_t* PL, const uint8_t* PR, const SbmGeom& g, int y, int d, int x0, int x1, uint16_t* hsad,
                           int* htext) {
  if (x0 >= x1) return;
  int s = 0, t = 0;
  for (int c = x0 - g.wsz2; c <= x0 + g.wsz2; ++c) {
    s += sbm_diff(PL, PR, g, y, c, d);
    if (d == 0) t += sbm_text(PL, g, y, c);
  }
  for (int x = x0;; ++x) {
    hsad[((size_t)y * g.width1 + x) * g.ndisp + d] = (uint16_t)s;
    if (d == 0) htext[(size_t)y * g.width1 + x] = t;
    if (x + 1 >= x1) break;
    s += sbm_diff(PL, PR, g, y, x + 1 + g.wsz2, d) - sbm_diff(PL, PR, g, y, x - g.wsz2, d);
    if (d == 0) t += sbm_text(PL, g, y, x + 1 + g.wsz2) - sbm_text(PL, g, y, x - g.wsz2);
  }
}

// Vertical pass, one thread = (output column x, candidate d): window rows y - wsz2 .. y + wsz2 with the row index clamped
CB_HD int sbm_vsad_init(const uint16_t* hsad, const SbmGeom& g, int x, int d, int y) {
  int s = 0;
  for (int j = -g.wsz2; j <= g.wsz2; ++j) s += hsad[((size_t)sbm_clampi(y + j, 0, g.h - 1) * g.width1 + x) * g.ndisp + d];
  return s;
}
CB_HD int sbm_vsad_step(const uint16_t* hsad, const SbmGeom& g, int x, int d, int y /*row just finished*/, int s) {
  return s + hsad[((size_t)sbm_clampi(y + 1 + g.wsz2, 0, g.h - 1) * g.width1 + x) * g.ndisp + d] -
         hsad[((size_t)sbm_clampi(y - g.wsz2, 0, g.h - 1) * g.width1 + x) * g.ndisp + d];
}
CB_HD int sbm_vtext_init(const int* htext, const SbmGeom& g, int x, int y) {
  int s = 0;
  for (int j = -g.wsz2; j <= g.wsz2; ++j) s += htext[(size_t)sbm_clampi(y + j, 0, g.h - 1) * g.width1 + x];
  return s;
}
CB_HD int sbm_vtext_step(const int* htext, const SbmGeom& g, int x, int y, int s) {
  return s + htext[(size_t)sbm_clampi(y + 1 + g.wsz2, 0, g.h - 1) * g.width1 + x] - htext[(size_t)sbm_clampi(y - g.wsz2, 0, g.h - 1) * g.width1 + x];
}

// Decision for one pixel from its ndisp SAD values: first minimum, texture and uniqueness rejection, parabola sub-pixel
// fit, disparity * 16 (dispDescale<short>).  `sad` points at candidate 0 of an array with one writable guard element on
// either side (sad[-1], sad[ndisp]).
CB_HD int sbm_decide(int* sad, const SbmGeom& g, int tsum) {
  int minsad = 0x7fffffff, mind = -1;
  for (int d = 0; d < g.ndisp; ++d)
    if (sad[d] < minsad) {
      minsad = sad[d];
      mind = d;
    }
  if (tsum < g.texture_threshold) return g.filtered;
  if (g.uniqueness_ratio > 0) {
    const int thresh = minsad + (minsad * g.uniqueness_ratio / 100);
    for (int d = 0; d < g.ndisp; ++d)
      if ((d < mind - 1 || d > mind + 1) && sad[d] <= thresh) return g.filtered;
  }
  sad[-1] = sad[1];
  sad[g.ndisp] = sad[g.ndisp - 2];
  const int p = sad[mind + 1], n = sad[mind - 1];
  const int dd = p + n - 2 * sad[mind] + (p > n ? p - n : n - p);
  return (((g.ndisp - mind - 1 + g.mindisp) * 256 + (dd != 0 ? (p - n) * 256 / dd : 0) + 15) >> 4);
}

// getValidDisparityROI with empty roi1 / roi2: is image pixel (y, X) inside?
CB_HD bool sbm_in_roi(const SbmGeom& g, int y, int X) {
  const int maxd = g.mindisp + g.ndisp - 1;
  const int xmin = (maxd > 0 ? maxd : 0) + g.wsz2, xmax = (g.w < g.w - g.mindisp ? g.w : g.w - g.mindisp) - g.wsz2;
  const int ymin = g.wsz2, ymax = g.h - g.wsz2;
  if (xmax - xmin <= 0 || ymax - ymin <= 0) return true;  // empty ROI: OpenCV crops nothing
  return y >= ymin && y < ymax && X >= xmin && X < xmax;
}

// StereoGeometry::disparity_to_3DPoints, CameraGeometry.cpp:500-520
CB_HD void sbm_point3d(int16_t disp16, int i, int j, float Q03, float Q13, float Q23, float Q32, float Q33, float* out) {
  const float d = (float)disp16;
#ifdef __CUDA_ARCH__  // no FMA contraction: the reference's (and the oracle's) operations are rounded one by one
  const double den = __dadd_rn(__dadd_rn(__dmul_rn((double)d / 16., (double)Q32), (double)Q33), 1E-6);
#else
  const double den = (double)d / 16. * (double)Q32 + (double)Q33 + 1E-6;
#endif
  const float pw = (float)(1.0 / den);  // 1.0f / double -> float
  out[0] = ((float)j + Q03) * pw;
  out[1] = ((float)i + Q13) * pw;
  out[2] = Q23 * pw;
}
