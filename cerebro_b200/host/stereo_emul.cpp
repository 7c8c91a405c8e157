// CPU emulation of the stereo kernels of csrc/frontend.cu: the same per-thread bodies (csrc/stereo_core.h), walked over the
// kernels' (block, thread) space in plain loops.  Built by cerebro_b200/build.py into _native/libstereo_emul.so and used
// only by tests/test_stereo.py to validate the kernel logic against the oracle without a GPU.  Not a fallback: nothing in
// the product path loads it.
#include <vector>

#include "../csrc/stereo_core.h"

extern "C" int sbm_emulate(const uint8_t* left, const uint8_t* right, int h, int w, int ndisp, int wsz, int seg, int stripe,
                           int16_t* disp) {
  const SbmGeom g = sbm_make_geom(h, w, ndisp, wsz);
  for (size_t i = 0; i < (size_t)h * w; ++i) disp[i] = (int16_t)g.filtered;  // fill kernel
  if (g.lofs >= w || g.rofs >= w || g.width1 < 1) return 0;
  std::vector<uint8_t> PL((size_t)h * w), PR((size_t)h * w);
  for (int y = 0; y < h; ++y)  // sbm_prefilter_kernel: thread per pixel
    for (int x = 0; x < w; ++x) {
      PL[(size_t)y * w + x] = sbm_prefilter_px(left, h, w, y, x, g.cap);
      PR[(size_t)y * w + x] = sbm_prefilter_px(right, h, w, y, x, g.cap);
    }
  std::vector<uint16_t> hsad((size_t)h * g.width1 * ndisp);
  std::vector<int> htext((size_t)h * g.width1);
  const int n_seg = (g.width1 + seg - 1) / seg;
  for (int y = 0; y < h; ++y)  // sbm_hsad_kernel: grid (n_seg, h), block = ndisp threads
    for (int s = 0; s < n_seg; ++s)
      for (int d = 0; d < ndisp; ++d) {
        const int x0 = s * seg, x1 = (x0 + seg < g.width1) ? x0 + seg : g.width1;
        sbm_hsad_thread(PL.data(), PR.data(), g, y, d, x0, x1, hsad.data(), htext.data());
      }
  const int n_stripe = (h + stripe - 1) / stripe;
  std::vector<int> s_sad(ndisp + 2), run(ndisp);
  for (int x = 0; x < g.width1; ++x)  // sbm_vsad_kernel: grid (width1, n_stripe), block = ndisp threads
    for (int st = 0; st < n_stripe; ++st) {
      const int y0 = st * stripe, y1 = (y0 + stripe < h) ? y0 + stripe : h;
      for (int d = 0; d < ndisp; ++d) run[d] = sbm_vsad_init(hsad.data(), g, x, d, y0);
      int tsum = sbm_vtext_init(htext.data(), g, x, y0);
      for (int y = y0; y < y1; ++y) {
        for (int d = 0; d < ndisp; ++d) s_sad[d + 1] = run[d];
        const int v = sbm_decide(s_sad.data() + 1, g, tsum);  // thread 0
        disp[(size_t)y * w + g.lofs + x] = (int16_t)(sbm_in_roi(g, y, g.lofs + x) ? v : g.filtered);
        for (int d = 0; d < ndisp; ++d) run[d] = sbm_vsad_step(hsad.data(), g, x, d, y, run[d]);
        tsum = sbm_vtext_step(htext.data(), g, x, y, tsum);
      }
    }
  return 0;
}

extern "C" void sbm_emulate_3d(const int16_t* disp, int h, int w, float Q03, float Q13, float Q23, float Q32, float Q33, float* out) {
  for (int i = 0; i < h; ++i)
    for (int j = 0; j < w; ++j) sbm_point3d(disp[(size_t)i * w + j], i, j, Q03, Q13, Q23, Q32, Q33, out + ((size_t)i * w + j) * 3);
}
