// CPU emulation of the ORB / remap kernels of csrc/features.cu over the shared per-item bodies (csrc/orb_core.h): the same
// stages in the same order, plain loops instead of launches.  TEST INFRASTRUCTURE (never a fallback of the product path):
// it lets tests/test_orb.py check indexing, rounding and ordering bit-exactly against the oracle and the installed OpenCV
// on a machine without a GPU.  Built by cerebro_b200/build.py with g++ -ffp-contract=off.
#include <cstring>
#include <vector>

#include "../csrc/orb_pattern.h"
#include "../csrc/orb_pipeline.h"

extern "C" {

// returns the number of keypoints written (<= max_kp), or -1 when max_kp is too small
int orb_emul_detect_and_compute(const uint8_t* img, int rows, int cols, int nfeatures, int max_kp, float* xy, float* size,
                                float* angle, float* response, int32_t* octave, uint8_t* desc) {
  using namespace orb;
  Level lv[kLevels];
  const size_t total = level_geometry(rows, cols, lv);
  std::vector<uint8_t> pyr(total), score(total), blur(total);
  std::vector<float> hor(total);
  memcpy(pyr.data(), img, (size_t)rows * cols);
  for (int l = 1; l < kLevels; ++l) {
    std::vector<int> ox, cx, oy, cy;
    int minx, maxx, miny, maxy;
    lin_coeffs(lv[l - 1].w, lv[l].w, ox, cx, minx, maxx);
    lin_coeffs(lv[l - 1].h, lv[l].h, oy, cy, miny, maxy);
    for (int y = 0; y < lv[l].h; ++y)
      for (int x = 0; x < lv[l].w; ++x)
        pyr[lv[l].off + (size_t)y * lv[l].w + x] = resize_px(pyr.data() + lv[l - 1].off, lv[l - 1].w, lv[l - 1].h, x, y, ox.data(),
                                                             cx.data(), minx, maxx, oy.data(), cy.data(), miny, maxy);
  }
  int npl[kLevels], umax[kHalfPatch + 2];
  features_per_level(nfeatures, npl);
  umax_table(umax);
  struct Kp {
    int x, y, level;
    float response;
  };
  std::vector<Kp> stage1;
  std::vector<int> stage1_count(kLevels, 0);
  for (int l = 0; l < kLevels; ++l) {
    const uint8_t* im = pyr.data() + lv[l].off;
    uint8_t* sc = score.data() + lv[l].off;
    const int w = lv[l].w, h = lv[l].h;
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) sc[(size_t)y * w + x] = fast_score(im, w, h, x, y);
    std::vector<Rec> cand;
    std::vector<Kp> pts;
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        if (nms_keep(sc, w, h, x, y)) {
          cand.push_back(Rec{(float)sc[(size_t)y * w + x], (int)pts.size()});
          pts.push_back(Kp{x, y, l, 0.f});
        }
    retain_best(cand, 2 * npl[l]);
    for (const Rec& r : cand) stage1.push_back(pts[(size_t)r.idx]);
    stage1_count[(size_t)l] = (int)cand.size();
  }
  for (Kp& k : stage1) k.response = harris(pyr.data() + lv[k.level].off, lv[k.level].w, k.x, k.y);
  std::vector<Kp> fin;
  size_t o = 0;
  for (int l = 0; l < kLevels; ++l) {
    std::vector<Rec> cand((size_t)stage1_count[(size_t)l]);
    for (size_t i = 0; i < cand.size(); ++i) cand[i] = Rec{stage1[o + i].response, (int)i};
    retain_best(cand, npl[l]);
    for (const Rec& r : cand) fin.push_back(stage1[o + (size_t)r.idx]);
    o += (size_t)stage1_count[(size_t)l];
  }
  if ((int)fin.size() > max_kp) return -1;
  for (int l = 0; l < kLevels; ++l) {
    const int w = lv[l].w, h = lv[l].h;
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) hor[lv[l].off + (size_t)y * w + x] = blur_row(pyr.data() + lv[l].off, w, x, y);
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) blur[lv[l].off + (size_t)y * w + x] = blur_col(hor.data() + lv[l].off, w, h, x, y);
  }
  for (size_t i = 0; i < fin.size(); ++i) {
    const Kp& k = fin[i];
    const Level& L = lv[k.level];
    const float ang = ic_angle(pyr.data() + L.off, L.w, k.x, k.y, umax);
    xy[2 * i] = (float)k.x * L.scale;
    xy[2 * i + 1] = (float)k.y * L.scale;
    size[i] = (float)(2 * kHalfPatch + 1) * L.scale;
    angle[i] = ang;
    response[i] = k.response;
    octave[i] = k.level;
    // computeOrbDescriptors recovers the level coordinates from the scaled point: cvRound(pt * (1.f / layerScale))
    const float inv = 1.f / L.scale;
    const int cx = (int)lrintf(xy[2 * i] * inv), cy = (int)lrintf(xy[2 * i + 1] * inv);
    descriptor(blur.data() + L.off, L.w, cx, cy, ang, kOrbPattern, desc + 32 * i);
  }
  return (int)fin.size();
}

void orb_emul_remap(const uint8_t* src, int rows, int cols, const float* map_x, const float* map_y, uint8_t* dst) {
  for (int y = 0; y < rows; ++y)
    for (int x = 0; x < cols; ++x)
      dst[(size_t)y * cols + x] = orb::remap_px(src, cols, rows, map_x[(size_t)y * cols + x], map_y[(size_t)y * cols + x]);
}

}  // extern "C"
