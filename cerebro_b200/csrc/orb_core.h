// Per-pixel / per-keypoint bodies of cv::ORB (nfeatures, scaleFactor 1.2, 8 levels, edgeThreshold 31, HARRIS_SCORE, patch 31,
// fastThreshold 0) as the reference calls it (src/utils/PointFeatureMatching.cpp:16-22), and of cv::remap(INTER_LINEAR) with
// CV_32FC1 maps (src/utils/CameraGeometry.cpp:42, 381-382).  Compiled twice: into the CUDA kernels of csrc/features.cu, and by
// g++ into a CPU emulation that walks the same item space (host/orb_emul.cpp) so that indexing, integer arithmetic, rounding
// order and tie rules are checked bit-exactly against the oracle / OpenCV without a GPU.  Plain pointers, no CUDA types.
//
// Arithmetic follows OpenCV's modules/features2d/src/{orb,fast,fast_score}.cpp and imgproc's resize_bitExact / remapBilinear:
// every floating-point expression is evaluated in the same order with separate roundings (no fused multiply-add).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CB_UNROLL _Pragma("unroll")
#else
#define CB_UNROLL
#endif

#if defined(__CUDA_ARCH__)
#define CB_HD __host__ __device__ __forceinline__
#define CB_FMUL(a, b) __fmul_rn((a), (b))
#define CB_FADD(a, b) __fadd_rn((a), (b))
#define CB_FSUB(a, b) __fsub_rn((a), (b))
#define CB_FDIV(a, b) __fdiv_rn((a), (b))
#define CB_RINT(a) __float2int_rn(a)
#elif defined(__CUDACC__)
#define CB_HD __host__ __device__ inline
#define CB_FMUL(a, b) ((a) * (b))
#define CB_FADD(a, b) ((a) + (b))
#define CB_FSUB(a, b) ((a) - (b))
#define CB_FDIV(a, b) ((a) / (b))
#define CB_RINT(a) ((int)lrintf(a))
#else  // g++ -ffp-contract=off
#define CB_HD inline
#define CB_FMUL(a, b) ((a) * (b))
#define CB_FADD(a, b) ((a) + (b))
#define CB_FSUB(a, b) ((a) - (b))
#define CB_FDIV(a, b) ((a) / (b))
#define CB_RINT(a) ((int)lrintf(a))
#endif

namespace orb {

constexpr int kLevels = 8;
constexpr int kEdge = 31;        // edgeThreshold
constexpr int kHalfPatch = 15;   // patchSize 31
constexpr int kHarrisBlock = 7;  // HarrisResponses(..., 7, HARRIS_K)

// ---- INTER_LINEAR_EXACT: one destination pixel.  ox / cx (oy / cy): source offset and 8.8 weight of the right (lower)
// neighbour per destination column (row); columns outside [minx, maxx) replicate the first / last source column.
CB_HD int resize_hor(const uint8_t* row, int sw, int x, const int* ox, const int* cx, int minx, int maxx) {
  if (x < minx) return (int)row[0] * 256;
  if (x >= maxx) return (int)row[sw - 1] * 256;
  const int o = ox[x], c = cx[x];
  return (256 - c) * (int)row[o] + c * (int)row[o + 1];
}
CB_HD uint8_t resize_px(const uint8_t* src, int sw, int sh, int x, int y, const int* ox, const int* cx, int minx, int maxx,
                        const int* oy, const int* cy, int miny, int maxy) {
  if (y < miny) return (uint8_t)((resize_hor(src, sw, x, ox, cx, minx, maxx) + 128) >> 8);
  if (y >= maxy) return (uint8_t)((resize_hor(src + (size_t)(sh - 1) * sw, sw, x, ox, cx, minx, maxx) + 128) >> 8);
  const int o = oy[y], c = cy[y];
  const long long v = (long long)(256 - c) * resize_hor(src + (size_t)o * sw, sw, x, ox, cx, minx, maxx) +
                      (long long)c * resize_hor(src + (size_t)(o + 1) * sw, sw, x, ox, cx, minx, maxx);
  const long long r = (v + (1 << 15)) >> 16;
  return (uint8_t)(r > 255 ? 255 : r);
}

// ---- FAST-9/16, threshold 0: cornerScore<16> of the pixel when it is a corner, else 0 (what fast.cpp's row buffers hold)
CB_HD uint8_t fast_score(const uint8_t* img, int w, int h, int x, int y) {
  if (x < 3 || y < 3 || x >= w - 3 || y >= h - 3) return 0;
  const uint8_t* p = img + (size_t)y * w + x;
  const int v = p[0];
  int d[16];
  d[0] = v - p[3 * w];
  d[1] = v - p[3 * w + 1];
  d[2] = v - p[2 * w + 2];
  d[3] = v - p[w + 3];
  d[4] = v - p[3];
  d[5] = v - p[-w + 3];
  d[6] = v - p[-2 * w + 2];
  d[7] = v - p[-3 * w + 1];
  d[8] = v - p[-3 * w];
  d[9] = v - p[-3 * w - 1];
  d[10] = v - p[-2 * w - 2];
  d[11] = v - p[-w - 3];
  d[12] = v - p[-3];
  d[13] = v - p[w - 3];
  d[14] = v - p[2 * w - 2];
  d[15] = v - p[3 * w - 1];
  // cornerScore<16> = (largest t such that nine contiguous circle pixels are all darker than v - t + 1 ... i.e. all d >= t, or
  // all brighter, all d <= -t) - 1: bisection on t over 16-bit masks of the circle.  [A straightforward min / max over the 16
  // arcs compiles to chains of 3-input VIMNMX3 on sm_100a that returned max|d| - 1 on the B200 -- measured with a standalone
  // probe, nvcc 12.9, also with inline-asm min.s32 and ptxas -O1 -- so the score is computed without integer min / max.]
  int lo = 0, hi = 255;  // invariant: a corner exists at threshold lo (0 = none known), none at hi + 1
  while (lo < hi) {
    const int t = (lo + hi + 1) >> 1;
    unsigned mb = 0, md = 0;
CB_UNROLL
    for (int k = 0; k < 16; ++k) {
      mb |= (unsigned)(d[k] >= t) << k;
      md |= (unsigned)(d[k] <= -t) << k;
    }
    mb |= mb << 16;
    md |= md << 16;
    unsigned rb = mb, rd = md;
CB_UNROLL
    for (int i = 1; i < 9; ++i) {
      rb &= mb >> i;
      rd &= md >> i;
    }
    if (((rb | rd) & 0xffffu) != 0)
      lo = t;
    else
      hi = t - 1;
  }
  return lo > 0 ? (uint8_t)(lo - 1) : (uint8_t)0;
}

// ---- non-maximum suppression (strictly greater than the 8 neighbours) + KeyPointsFilter::runByImageBorder(edgeThreshold)
CB_HD bool nms_keep(const uint8_t* score, int w, int h, int x, int y) {
  if (x < kEdge || y < kEdge || x >= w - kEdge || y >= h - kEdge) return false;
  const uint8_t* s = score + (size_t)y * w + x;
  const int c = s[0];
  return c > s[-1] && c > s[1] && c > s[-w - 1] && c > s[-w] && c > s[-w + 1] && c > s[w - 1] && c > s[w] && c > s[w + 1];
}

// ---- HarrisResponses (orb.cpp): 7 x 7 block of Sobel products
CB_HD float harris(const uint8_t* img, int w, int x0, int y0) {
  const int r = kHarrisBlock / 2;
  int a = 0, b = 0, c = 0;
  for (int i = -r; i <= r; ++i)
    for (int j = -r; j <= r; ++j) {
      const uint8_t* p = img + (size_t)(y0 + i) * w + (x0 + j);
      const int Ix = ((int)p[1] - (int)p[-1]) * 2 + ((int)p[-w + 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[w - 1]);
      const int Iy = ((int)p[w] - (int)p[-w]) * 2 + ((int)p[w - 1] - (int)p[-w - 1]) + ((int)p[w + 1] - (int)p[-w + 1]);
      a += Ix * Ix;
      b += Iy * Iy;
      c += Ix * Iy;
    }
  const float scale = CB_FDIV(1.f, CB_FMUL((float)((1 << 2) * kHarrisBlock), 255.f));
  const float s4 = CB_FMUL(CB_FMUL(CB_FMUL(scale, scale), scale), scale);
  const float fa = (float)a, fb = (float)b, fc = (float)c;
  const float sab = CB_FADD(fa, fb);
  const float t = CB_FSUB(CB_FSUB(CB_FMUL(fa, fb), CB_FMUL(fc, fc)), CB_FMUL(CB_FMUL(0.04f, sab), sab));
  return CB_FMUL(t, s4);
}

// ---- cv::fastAtan2 (core mathfuncs, scalar atan_f32), degrees
CB_HD float fast_atan2(float y, float x) {
  const float p1 = 57.2836266f, p3 = -18.6674461f, p5 = 8.91400051f, p7 = -2.53972459f;  // 0.99978784f*(float)(180/CV_PI) ...
  const float eps = 2.22044605e-16f;
  const float ax = fabsf(x), ay = fabsf(y);
  float a, c, c2;
  if (ax >= ay) {
    c = CB_FDIV(ay, CB_FADD(ax, eps));
    c2 = CB_FMUL(c, c);
    a = CB_FMUL(CB_FADD(CB_FMUL(CB_FADD(CB_FMUL(CB_FADD(CB_FMUL(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = CB_FDIV(ax, CB_FADD(ay, eps));
    c2 = CB_FMUL(c, c);
    a = CB_FSUB(90.f, CB_FMUL(CB_FADD(CB_FMUL(CB_FADD(CB_FMUL(CB_FADD(CB_FMUL(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0.f) a = CB_FSUB(180.f, a);
  if (y < 0.f) a = CB_FSUB(360.f, a);
  return a;
}

// ---- ICAngles: intensity centroid over the circular patch of radius 15; umax[v] = half width of row v
CB_HD float ic_angle(const uint8_t* img, int w, int x0, int y0, const int* umax) {
  const uint8_t* c = img + (size_t)y0 * w + x0;
  int m01 = 0, m10 = 0;
  for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * (int)c[u];
  for (int v = 1; v <= kHalfPatch; ++v) {
    int vsum = 0;
    const int d = umax[v];
    for (int u = -d; u <= d; ++u) {
      const int vp = c[u + v * w], vm = c[u - v * w];
      vsum += vp - vm;
      m10 += u * (vp + vm);
    }
    m01 += v * vsum;
  }
  return fast_atan2((float)m01, (float)m10);
}

// ---- GaussianBlur(7x7, sigma 2) in float, row pass then column pass, BORDER_REFLECT_101; taps = getGaussianKernel(7, 2, CV_32F)
CB_HD int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
CB_HD float gauss_tap(int i) {
  const float k[7] = {0.07015932f, 0.13107488f, 0.19071282f, 0.21610594f, 0.19071282f, 0.13107488f, 0.07015932f};
  return k[i];
}
CB_HD float blur_row(const uint8_t* img, int w, int x, int y) {
  const uint8_t* r = img + (size_t)y * w;
  float acc = CB_FMUL(gauss_tap(0), (float)r[reflect101(x - 3, w)]);
CB_UNROLL
  for (int i = 1; i < 7; ++i) acc = CB_FADD(acc, CB_FMUL(gauss_tap(i), (float)r[reflect101(x - 3 + i, w)]));
  return acc;
}
CB_HD uint8_t blur_col(const float* hor, int w, int h, int x, int y) {
  float acc = CB_FMUL(gauss_tap(0), hor[(size_t)reflect101(y - 3, h) * w + x]);
CB_UNROLL
  for (int i = 1; i < 7; ++i) acc = CB_FADD(acc, CB_FMUL(gauss_tap(i), hor[(size_t)reflect101(y - 3 + i, h) * w + x]));
  const int v = CB_RINT(acc);
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ---- computeOrbDescriptors, WTA_K = 2: 256 binary tests on the blurred level image around (x0, y0), pattern rotated by `angle`
CB_HD void descriptor(const uint8_t* blur, int w, int x0, int y0, float angle_deg, const signed char* pattern, uint8_t* out) {
  const float ang = CB_FMUL(angle_deg, 0.0174532924f);  // (float)(CV_PI / 180.f)
  const float a = (float)cos((double)ang), b = (float)sin((double)ang);
  const uint8_t* c = blur + (size_t)y0 * w + x0;
  for (int i = 0; i < 32; ++i) {
    int val = 0;
CB_UNROLL
    for (int j = 0; j < 8; ++j) {
      const signed char* pp = pattern + ((i * 8 + j) * 4);
      const float x1 = (float)pp[0], y1 = (float)pp[1], x2 = (float)pp[2], y2 = (float)pp[3];
      const int ix1 = CB_RINT(CB_FSUB(CB_FMUL(x1, a), CB_FMUL(y1, b))), iy1 = CB_RINT(CB_FADD(CB_FMUL(x1, b), CB_FMUL(y1, a)));
      const int ix2 = CB_RINT(CB_FSUB(CB_FMUL(x2, a), CB_FMUL(y2, b))), iy2 = CB_RINT(CB_FADD(CB_FMUL(x2, b), CB_FMUL(y2, a)));
      const int t0 = c[iy1 * w + ix1], t1 = c[iy2 * w + ix2];
      val |= (t0 < t1) << j;
    }
    out[i] = (uint8_t)val;
  }
}

// ---- cv::remap, INTER_LINEAR, CV_32FC1 maps, BORDER_CONSTANT 0 (imgwarp.cpp remapBilinear<FixedPtCast<int,uchar,15>>):
// coordinates to 1/32 pixel (cvRound), integer weights summing to 2^15
CB_HD uint8_t remap_px(const uint8_t* src, int w, int h, float mx, float my) {
  const int sx = CB_RINT(CB_FMUL(mx, 32.f)), sy = CB_RINT(CB_FMUL(my, 32.f));
  const int fx = sx & 31, fy = sy & 31;
  int ix = sx >> 5, iy = sy >> 5;
  ix = ix < -32768 ? -32768 : (ix > 32767 ? 32767 : ix);  // saturate_cast<short>
  iy = iy < -32768 ? -32768 : (iy > 32767 ? 32767 : iy);
  const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
  const bool x0 = ix >= 0 && ix < w, x1 = ix + 1 >= 0 && ix + 1 < w, y0 = iy >= 0 && iy < h, y1 = iy + 1 >= 0 && iy + 1 < h;
  const int p00 = (x0 && y0) ? src[(size_t)iy * w + ix] : 0, p01 = (x1 && y0) ? src[(size_t)iy * w + ix + 1] : 0;
  const int p10 = (x0 && y1) ? src[(size_t)(iy + 1) * w + ix] : 0, p11 = (x1 && y1) ? src[(size_t)(iy + 1) * w + ix + 1] : 0;
  return (uint8_t)((w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11 + (1 << 14)) >> 15);
}

}  // namespace orb
