// cb_features_*: the image-side front end of a loop candidate on the device:
//   * cv::remap(INTER_LINEAR) with CV_32FC1 maps -- the undistortion and stereo-rectification warps the reference applies to
//     both images of a stereo pair (src/utils/CameraGeometry.cpp:42, 381-382; the maps themselves come from camodocal /
//     cv::initUndistortRectifyMap once at start-up, :20-34, :348-349, and stay host-side set-up)
//   * cv::ORB::create(n) + setFastThreshold(0) + detectAndCompute (src/utils/PointFeatureMatching.cpp:16-22): 8-level
//     INTER_LINEAR_EXACT pyramid, FAST-9/16 score map, non-maximum suppression + border filter, Harris re-ranking, intensity
//     centroid orientation, 7x7 Gaussian, 256-bit rBRIEF.
// Integer / byte work, HBM-bound and tiny (1.1 M pyramid pixels per 480 x 752 image): one thread per pixel or per keypoint
// over the bodies of orb_core.h, which a CPU emulation shares (host/orb_emul.cpp) -- so every stage is checked bit-exactly
// against OpenCV without a GPU, and the kernels only add the item -> thread mapping.  The two KeyPointsFilter::retainBest
// selections run on the host between kernels: cv::ORB's keypoint ORDER is the permutation libstdc++'s std::nth_element leaves
// behind, which only libstdc++ reproduces (orb_pipeline.h); they see a few ten thousand 5-byte records per image.
#include "common.cuh"
#include "orb_pattern.h"
#include "orb_pipeline.h"

#include <vector>

namespace {

using orb::kLevels;

struct LevelDev {
  int w, h;
  unsigned off;       // pixel offset in the pyramid
  unsigned cand_off;  // offset of the level's candidate slots
  unsigned row0;      // first global row index of the level
  float scale;
};

struct PyrDesc {
  LevelDev lv[kLevels];
  unsigned total_px, cand_total, rows_total;
};

__constant__ signed char c_pattern[1024];

__global__ void remap_kernel(const uint8_t* __restrict__ src, const float* __restrict__ mx, const float* __restrict__ my, int rows,
                             int cols, uint8_t* __restrict__ dst) {
  const size_t n = (size_t)rows * cols;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t img = blockIdx.y;
  dst[img * n + i] = orb::remap_px(src + img * n, cols, rows, mx[i], my[i]);
}

__global__ void resize_kernel(uint8_t* __restrict__ pyr, PyrDesc pd, int l, const int* __restrict__ ox, const int* __restrict__ cx,
                              int minx, int maxx, const int* __restrict__ oy, const int* __restrict__ cy, int miny, int maxy) {
  const LevelDev d = pd.lv[l], s = pd.lv[l - 1];
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned)(d.w * d.h)) return;
  uint8_t* base = pyr + (size_t)blockIdx.y * pd.total_px;
  const int x = i % d.w, y = i / d.w;
  base[d.off + i] = orb::resize_px(base + s.off, s.w, s.h, x, y, ox, cx, minx, maxx, oy, cy, miny, maxy);
}

__device__ __forceinline__ int level_of(const PyrDesc& pd, unsigned px) {
  int l = 0;
#pragma unroll
  for (int k = 1; k < kLevels; ++k) l += px >= pd.lv[k].off;
  return l;
}

__global__ void fast_kernel(const uint8_t* __restrict__ pyr, PyrDesc pd, uint8_t* __restrict__ score) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pd.total_px) return;
  const LevelDev d = pd.lv[level_of(pd, i)];
  const unsigned r = i - d.off;
  const size_t base = (size_t)blockIdx.y * pd.total_px;
  score[base + i] = orb::fast_score(pyr + base + d.off, d.w, d.h, r % d.w, r / d.w);
}

// One block per pyramid row: ordered count / write of the pixels that survive non-maximum suppression and the border filter.
// pass 0 writes the row's count; pass 1 (after the row offsets are known) writes the candidates in row-major order -- the
// order cv::FAST emits them in.
template <int PASS>
__global__ void __launch_bounds__(128) compact_kernel(const uint8_t* __restrict__ score, PyrDesc pd, unsigned* __restrict__ row_count,
                                                      const unsigned* __restrict__ row_off, unsigned* __restrict__ cand_xy,
                                                      uint8_t* __restrict__ cand_sc) {
  __shared__ unsigned warp_sum[4];
  __shared__ unsigned running;
  const unsigned grow = blockIdx.x, img = blockIdx.y;
  int l = 0;
#pragma unroll
  for (int k = 1; k < kLevels; ++k) l += grow >= pd.lv[k].row0;
  const LevelDev d = pd.lv[l];
  const int y = grow - d.row0;
  const uint8_t* sc = score + (size_t)img * pd.total_px + d.off;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const unsigned out0 = PASS ? (d.cand_off + row_off[(size_t)img * pd.rows_total + grow]) : 0u;
  for (int x0 = 0; x0 < d.w; x0 += 128) {
    const int x = x0 + threadIdx.x;
    const bool keep = x < d.w && orb::nms_keep(sc, d.w, d.h, x, y);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_sum[warp] = __popc(b);
    __syncthreads();
    unsigned before = running;
    for (int w2 = 0; w2 < warp; ++w2) before += warp_sum[w2];
    if (PASS && keep) {
      const unsigned o = out0 + before + __popc(b & ((1u << lane) - 1u));
      cand_xy[(size_t)img * pd.cand_total + o] = (unsigned)x | ((unsigned)y << 16);
      cand_sc[(size_t)img * pd.cand_total + o] = sc[(size_t)y * d.w + x];
    }
    __syncthreads();
    if (threadIdx.x == 0) running += warp_sum[0] + warp_sum[1] + warp_sum[2] + warp_sum[3];
    __syncthreads();
  }
  if (!PASS && threadIdx.x == 0) row_count[(size_t)img * pd.rows_total + grow] = running;
}

// per image: exclusive scan of the row counts inside every level, level totals
__global__ void scan_rows_kernel(const unsigned* __restrict__ row_count, PyrDesc pd, unsigned* __restrict__ row_off,
                                 unsigned* __restrict__ level_count) {
  const unsigned img = blockIdx.x;
  const int l = threadIdx.x;
  if (l >= kLevels) return;
  const LevelDev d = pd.lv[l];
  unsigned acc = 0;
  for (int y = 0; y < d.h; ++y) {
    const size_t i = (size_t)img * pd.rows_total + d.row0 + y;
    row_off[i] = acc;
    acc += row_count[i];
  }
  level_count[img * kLevels + l] = acc;
}

__global__ void harris_kernel(const uint8_t* __restrict__ pyr, PyrDesc pd, const unsigned* __restrict__ sel_xy,
                              const uint8_t* __restrict__ sel_level, const int* __restrict__ n_sel, int cap, float* __restrict__ resp) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
  if ((int)i >= n_sel[img]) return;
  const unsigned xy = sel_xy[(size_t)img * cap + i];
  const LevelDev d = pd.lv[sel_level[(size_t)img * cap + i]];
  resp[(size_t)img * cap + i] = orb::harris(pyr + (size_t)img * pd.total_px + d.off, d.w, xy & 0xffff, xy >> 16);
}

__global__ void blur_row_kernel(const uint8_t* __restrict__ pyr, PyrDesc pd, float* __restrict__ hor) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pd.total_px) return;
  const LevelDev d = pd.lv[level_of(pd, i)];
  const unsigned r = i - d.off;
  const size_t base = (size_t)blockIdx.y * pd.total_px;
  hor[base + i] = orb::blur_row(pyr + base + d.off, d.w, r % d.w, r / d.w);
}

__global__ void blur_col_kernel(const float* __restrict__ hor, PyrDesc pd, uint8_t* __restrict__ blur) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pd.total_px) return;
  const LevelDev d = pd.lv[level_of(pd, i)];
  const unsigned r = i - d.off;
  const size_t base = (size_t)blockIdx.y * pd.total_px;
  blur[base + i] = orb::blur_col(hor + base + d.off, d.w, d.h, r % d.w, r / d.w);
}

struct UmaxTab {
  int u[orb::kHalfPatch + 2];
};

// per final keypoint: orientation on the level image, the KeyPoint fields in level-0 coordinates, the rBRIEF descriptor on the
// blurred level image
__global__ void describe_kernel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, PyrDesc pd, UmaxTab um,
                                const unsigned* __restrict__ fin_xy, const uint8_t* __restrict__ fin_level, const int* __restrict__ n_fin,
                                int cap, int max_kp, float* __restrict__ o_xy, float* __restrict__ o_size, float* __restrict__ o_angle,
                                int* __restrict__ o_octave, uint8_t* __restrict__ o_desc) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x, img = blockIdx.y;
  if ((int)i >= n_fin[img]) return;
  const unsigned xy = fin_xy[(size_t)img * cap + i];
  const int l = fin_level[(size_t)img * cap + i];
  const LevelDev d = pd.lv[l];
  const int x = xy & 0xffff, y = xy >> 16;
  const size_t base = (size_t)img * pd.total_px + d.off;
  const float ang = orb::ic_angle(pyr + base, d.w, x, y, um.u);
  const size_t o = (size_t)img * max_kp + i;
  const float fx = __fmul_rn((float)x, d.scale), fy = __fmul_rn((float)y, d.scale);
  o_xy[2 * o] = fx;
  o_xy[2 * o + 1] = fy;
  o_size[o] = __fmul_rn((float)(2 * orb::kHalfPatch + 1), d.scale);
  o_angle[o] = ang;
  o_octave[o] = l;
  const float inv = __fdiv_rn(1.f, d.scale);
  const int cx = __float2int_rn(__fmul_rn(fx, inv)), cy = __float2int_rn(__fmul_rn(fy, inv));
  orb::descriptor(blur + base, d.w, cx, cy, ang, c_pattern, o_desc + 32 * o);
}

}  // namespace

struct cb_features {
  int device = 0, rows = 0, cols = 0, max_images = 0;
  cudaStream_t stream = nullptr;
  PyrDesc pd;
  orb::Level lv[kLevels];
  UmaxTab um;
  // resize tables per level (device) + their bounds
  int* tab[kLevels][4] = {};
  int bounds[kLevels][4] = {};
  uint8_t *img = nullptr, *pyr = nullptr, *score = nullptr, *blur = nullptr, *dst = nullptr;
  float* hor = nullptr;
  unsigned *row_count = nullptr, *row_off = nullptr, *level_count = nullptr, *cand_xy = nullptr, *sel_xy = nullptr;
  uint8_t *cand_sc = nullptr, *sel_level = nullptr;
  float* resp = nullptr;
  int* n_sel = nullptr;
  // outputs
  int max_kp = 0;
  float *o_xy = nullptr, *o_size = nullptr, *o_angle = nullptr;
  int* o_octave = nullptr;
  uint8_t* o_desc = nullptr;
  // pinned host staging
  unsigned *h_level_count = nullptr, *h_cand_xy = nullptr, *h_sel_xy = nullptr;
  uint8_t *h_cand_sc = nullptr, *h_sel_level = nullptr;
  float* h_resp = nullptr;
  int* h_n = nullptr;
  // remap slots
  float* map_x[4] = {};
  float* map_y[4] = {};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  float last_orb_ms = 0.f;
};

extern "C" {

int cb_features_destroy(cb_features* f) {
  if (!f) return CB_OK;
  cb::DeviceGuard g(f->device);
  if (f->stream) cb::sync_stream(f->stream);
  void* dev[] = {f->img,     f->pyr,    f->score,  f->blur,     f->dst,       f->hor,  f->row_count, f->row_off, f->level_count,
                 f->cand_xy, f->sel_xy, f->cand_sc, f->sel_level, f->resp,     f->n_sel, f->o_xy,      f->o_size,  f->o_angle,
                 f->o_octave, f->o_desc, f->map_x[0], f->map_x[1], f->map_x[2], f->map_x[3], f->map_y[0], f->map_y[1], f->map_y[2], f->map_y[3]};
  for (void* p : dev)
    if (p) cudaFree(p);
  for (int l = 0; l < kLevels; ++l)
    for (int k = 0; k < 4; ++k)
      if (f->tab[l][k]) cudaFree(f->tab[l][k]);
  void* host[] = {f->h_level_count, f->h_cand_xy, f->h_sel_xy, f->h_cand_sc, f->h_sel_level, f->h_resp, f->h_n};
  for (void* p : host)
    if (p) cudaFreeHost(p);
  for (cudaEvent_t e : f->ev)
    if (e) cudaEventDestroy(e);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
  return CB_OK;
}

int cb_features_create(cb_features** out, int rows, int cols, int max_images, int max_keypoints, int device) {
  if (!out) return cb::fail(CB_EINVAL, "out is NULL");
  *out = nullptr;
  if (rows < 96 || cols < 96 || rows > 8192 || cols > 8192 || max_images < 1 || max_keypoints < 1)
    return cb::fail(CB_EINVAL, "bad image size / batch for cb_features_create (%d x %d, %d images, %d keypoints)", rows, cols, max_images, max_keypoints);
  int rc = cb::select_device(device, nullptr);
  if (rc) return rc;
  cb::DeviceGuard g(device);
  cb_features* f = new cb_features();
  f->device = device;
  f->rows = rows;
  f->cols = cols;
  f->max_images = max_images;
  f->max_kp = max_keypoints;
  const size_t total = orb::level_geometry(rows, cols, f->lv);
  orb::umax_table(f->um.u);
  unsigned cand = 0, row0 = 0;
  for (int l = 0; l < kLevels; ++l) {
    LevelDev& d = f->pd.lv[l];
    d.w = f->lv[l].w;
    d.h = f->lv[l].h;
    d.off = (unsigned)f->lv[l].off;
    d.scale = f->lv[l].scale;
    d.cand_off = cand;
    d.row0 = row0;
    cand += (unsigned)(((d.w + 1) / 2) * ((d.h + 1) / 2));  // non-maximum suppression keeps at most one pixel of every 2 x 2
    row0 += (unsigned)d.h;
  }
  f->pd.total_px = (unsigned)total;
  f->pd.cand_total = cand;
  f->pd.rows_total = row0;
  const size_t n = (size_t)max_images;
  cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
#define CB_DEV(ptr, bytes) \
  if (e == cudaSuccess) e = cudaMalloc((void**)&(ptr), (bytes))
#define CB_HOST(ptr, bytes) \
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&(ptr), (bytes), cudaHostAllocDefault)
  CB_DEV(f->img, n * rows * cols);
  CB_DEV(f->dst, n * rows * cols);
  CB_DEV(f->pyr, n * total);
  CB_DEV(f->score, n * total);
  CB_DEV(f->blur, n * total);
  CB_DEV(f->hor, n * total * sizeof(float));
  CB_DEV(f->row_count, n * row0 * sizeof(unsigned));
  CB_DEV(f->row_off, n * row0 * sizeof(unsigned));
  CB_DEV(f->level_count, n * kLevels * sizeof(unsigned));
  CB_DEV(f->cand_xy, n * cand * sizeof(unsigned));
  CB_DEV(f->cand_sc, n * cand);
  CB_DEV(f->sel_xy, n * cand * sizeof(unsigned));
  CB_DEV(f->sel_level, n * cand);
  CB_DEV(f->resp, n * cand * sizeof(float));
  CB_DEV(f->n_sel, n * sizeof(int));
  CB_DEV(f->o_xy, n * max_keypoints * 2 * sizeof(float));
  CB_DEV(f->o_size, n * max_keypoints * sizeof(float));
  CB_DEV(f->o_angle, n * max_keypoints * sizeof(float));
  CB_DEV(f->o_octave, n * max_keypoints * sizeof(int));
  CB_DEV(f->o_desc, n * max_keypoints * 32);
  CB_HOST(f->h_level_count, n * kLevels * sizeof(unsigned));
  CB_HOST(f->h_cand_xy, n * cand * sizeof(unsigned));
  CB_HOST(f->h_cand_sc, n * cand);
  CB_HOST(f->h_sel_xy, n * cand * sizeof(unsigned));
  CB_HOST(f->h_sel_level, n * cand);
  CB_HOST(f->h_resp, n * cand * sizeof(float));
  CB_HOST(f->h_n, n * sizeof(int));
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreate(&f->ev[i]);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_pattern, kOrbPattern, sizeof(kOrbPattern));
  for (int l = 1; l < kLevels && e == cudaSuccess; ++l) {
    std::vector<int> t[4];
    orb::lin_coeffs(f->lv[l - 1].w, f->lv[l].w, t[0], t[1], f->bounds[l][0], f->bounds[l][1]);
    orb::lin_coeffs(f->lv[l - 1].h, f->lv[l].h, t[2], t[3], f->bounds[l][2], f->bounds[l][3]);
    for (int k = 0; k < 4 && e == cudaSuccess; ++k) {
      e = cudaMalloc((void**)&f->tab[l][k], t[k].size() * sizeof(int));
      if (e == cudaSuccess) e = cudaMemcpy(f->tab[l][k], t[k].data(), t[k].size() * sizeof(int), cudaMemcpyHostToDevice);
    }
  }
#undef CB_DEV
#undef CB_HOST
  if (e != cudaSuccess) {
    cb_features_destroy(f);
    return cb::fail(CB_ENOMEM, "cb_features_create: %s", cudaGetErrorString(e));
  }
  *out = f;
  return CB_OK;
}

int cb_features_set_remap(cb_features* f, int slot, const float* map_x, const float* map_y) {
  if (!f || !map_x || !map_y) return cb::fail(CB_EINVAL, "NULL argument to cb_features_set_remap");
  if (slot < 0 || slot >= 4) return cb::fail(CB_EINVAL, "remap slot %d outside [0,4)", slot);
  cb::DeviceGuard g(f->device);
  const size_t bytes = (size_t)f->rows * f->cols * sizeof(float);
  if (!f->map_x[slot]) CB_CUDA(cudaMalloc((void**)&f->map_x[slot], bytes));
  if (!f->map_y[slot]) CB_CUDA(cudaMalloc((void**)&f->map_y[slot], bytes));
  CB_CUDA(cudaMemcpyAsync(f->map_x[slot], map_x, bytes, cudaMemcpyHostToDevice, f->stream));
  CB_CUDA(cudaMemcpyAsync(f->map_y[slot], map_y, bytes, cudaMemcpyHostToDevice, f->stream));
  CB_CUDA(cb::sync_stream(f->stream));
  return CB_OK;
}

// two warps in sequence when slot_b >= 0 (raw -> undistorted -> rectified, rounding to 8 bits in between as the reference's
// two cv::remap calls do)
int cb_features_remap(cb_features* f, int n, const uint8_t* src, int slot_a, int slot_b, uint8_t* dst) {
  if (!f || !src || !dst) return cb::fail(CB_EINVAL, "NULL argument to cb_features_remap");
  if (n < 1 || n > f->max_images) return cb::fail(CB_EINVAL, "n %d outside [1,%d]", n, f->max_images);
  if (slot_a < 0 || slot_a >= 4 || !f->map_x[slot_a] || slot_b >= 4 || (slot_b >= 0 && !f->map_x[slot_b]))
    return cb::fail(CB_EINVAL, "remap slot not set (cb_features_set_remap)");
  cb::DeviceGuard g(f->device);
  const size_t px = (size_t)f->rows * f->cols;
  CB_CUDA(cudaMemcpyAsync(f->img, src, n * px, cudaMemcpyHostToDevice, f->stream));
  const dim3 grid((unsigned)((px + 255) / 256), (unsigned)n);
  remap_kernel<<<grid, 256, 0, f->stream>>>(f->img, f->map_x[slot_a], f->map_y[slot_a], f->rows, f->cols, f->dst);
  CB_LAUNCH_CHECK();
  const uint8_t* res = f->dst;
  if (slot_b >= 0) {
    remap_kernel<<<grid, 256, 0, f->stream>>>(f->dst, f->map_x[slot_b], f->map_y[slot_b], f->rows, f->cols, f->img);
    CB_LAUNCH_CHECK();
    res = f->img;
  }
  CB_CUDA(cudaMemcpyAsync(dst, res, n * px, cudaMemcpyDeviceToHost, f->stream));
  CB_CUDA(cb::sync_stream(f->stream));
  return CB_OK;
}

int cb_features_orb(cb_features* f, int n, const uint8_t* images, int n_features, int32_t* n_keypoints, float* kp_xy, float* kp_size,
                    float* kp_angle, float* kp_response, int32_t* kp_octave, uint8_t* descriptors) {
  if (!f || !images || !n_keypoints || !kp_xy || !descriptors) return cb::fail(CB_EINVAL, "NULL argument to cb_features_orb");
  if (n < 1 || n > f->max_images) return cb::fail(CB_EINVAL, "n %d outside [1,%d]", n, f->max_images);
  if (n_features < 1) return cb::fail(CB_EINVAL, "n_features must be positive");
  cb::DeviceGuard g(f->device);
  cudaStream_t st = f->stream;
  const PyrDesc& pd = f->pd;
  const size_t px0 = (size_t)f->rows * f->cols;
  const unsigned cap = pd.cand_total;
  int npl[kLevels];
  orb::features_per_level(n_features, npl);
  CB_CUDA(cudaEventRecord(f->ev[0], st));
  // ---- pyramid: level 0 = the image, level l resized from level l - 1
  CB_CUDA(cudaMemcpy2DAsync(f->pyr, pd.total_px, images, px0, px0, (size_t)n, cudaMemcpyHostToDevice, st));
  for (int l = 1; l < kLevels; ++l) {
    const unsigned npx = (unsigned)(pd.lv[l].w * pd.lv[l].h);
    resize_kernel<<<dim3((npx + 255) / 256, n), 256, 0, st>>>(f->pyr, pd, l, f->tab[l][0], f->tab[l][1], f->bounds[l][0], f->bounds[l][1],
                                                              f->tab[l][2], f->tab[l][3], f->bounds[l][2], f->bounds[l][3]);
    CB_LAUNCH_CHECK();
  }
  const dim3 gpx((pd.total_px + 255) / 256, n);
  fast_kernel<<<gpx, 256, 0, st>>>(f->pyr, pd, f->score);
  CB_LAUNCH_CHECK();
  compact_kernel<0><<<dim3(pd.rows_total, n), 128, 0, st>>>(f->score, pd, f->row_count, nullptr, nullptr, nullptr);
  CB_LAUNCH_CHECK();
  scan_rows_kernel<<<n, 32, 0, st>>>(f->row_count, pd, f->row_off, f->level_count);
  CB_LAUNCH_CHECK();
  compact_kernel<1><<<dim3(pd.rows_total, n), 128, 0, st>>>(f->score, pd, nullptr, f->row_off, f->cand_xy, f->cand_sc);
  CB_LAUNCH_CHECK();
  // the blur does not depend on the keypoints: it runs while the host selects
  blur_row_kernel<<<gpx, 256, 0, st>>>(f->pyr, pd, f->hor);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemcpyAsync(f->h_level_count, f->level_count, (size_t)n * kLevels * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cb::sync_stream(st));
  for (int i = 0; i < n; ++i)
    for (int l = 0; l < kLevels; ++l) {
      const unsigned c = f->h_level_count[i * kLevels + l];
      if (!c) continue;
      const size_t o = (size_t)i * cap + pd.lv[l].cand_off;
      CB_CUDA(cudaMemcpyAsync(f->h_cand_xy + o, f->cand_xy + o, c * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
      CB_CUDA(cudaMemcpyAsync(f->h_cand_sc + o, f->cand_sc + o, c, cudaMemcpyDeviceToHost, st));
    }
  blur_col_kernel<<<gpx, 256, 0, st>>>(f->hor, pd, f->blur);
  CB_LAUNCH_CHECK();
  CB_CUDA(cb::sync_stream(st));
  // ---- KeyPointsFilter::retainBest(2 x budget) on the FAST scores, per level (orb.cpp computeKeyPoints)
  std::vector<std::vector<int>> stage1_count((size_t)n, std::vector<int>(kLevels, 0));
  std::vector<orb::Rec> rec;
  int max_sel = 0;
  for (int i = 0; i < n; ++i) {
    int ns = 0;
    for (int l = 0; l < kLevels; ++l) {
      const unsigned c = f->h_level_count[i * kLevels + l];
      const size_t o = (size_t)i * cap + pd.lv[l].cand_off;
      rec.resize(c);
      for (unsigned k = 0; k < c; ++k) rec[k] = orb::Rec{(float)f->h_cand_sc[o + k], (int)k};
      orb::retain_best(rec, 2 * npl[l]);
      for (const orb::Rec& r : rec) {
        f->h_sel_xy[(size_t)i * cap + ns] = f->h_cand_xy[o + (size_t)r.idx];
        f->h_sel_level[(size_t)i * cap + ns] = (uint8_t)l;
        ++ns;
      }
      stage1_count[(size_t)i][(size_t)l] = (int)rec.size();
    }
    f->h_n[i] = ns;
    max_sel = ns > max_sel ? ns : max_sel;
    CB_CUDA(cudaMemcpyAsync(f->sel_xy + (size_t)i * cap, f->h_sel_xy + (size_t)i * cap, (size_t)ns * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(f->sel_level + (size_t)i * cap, f->h_sel_level + (size_t)i * cap, (size_t)ns, cudaMemcpyHostToDevice, st));
  }
  CB_CUDA(cudaMemcpyAsync(f->n_sel, f->h_n, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  if (max_sel > 0) {
    harris_kernel<<<dim3((max_sel + 127) / 128, n), 128, 0, st>>>(f->pyr, pd, f->sel_xy, f->sel_level, f->n_sel, (int)cap, f->resp);
    CB_LAUNCH_CHECK();
    for (int i = 0; i < n; ++i)
      CB_CUDA(cudaMemcpyAsync(f->h_resp + (size_t)i * cap, f->resp + (size_t)i * cap, (size_t)f->h_n[i] * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CB_CUDA(cb::sync_stream(st));
  // ---- retainBest(budget) on the Harris responses, per level; the survivors, in this order, are the keypoints
  int max_fin = 0;
  for (int i = 0; i < n; ++i) {
    int nf = 0;
    size_t o = (size_t)i * cap;
    for (int l = 0; l < kLevels; ++l) {
      const int c = stage1_count[(size_t)i][(size_t)l];
      rec.resize((size_t)c);
      for (int k = 0; k < c; ++k) rec[(size_t)k] = orb::Rec{f->h_resp[o + (size_t)k], k};
      orb::retain_best(rec, npl[l]);
      if (nf + (int)rec.size() > f->max_kp)
        return cb::fail(CB_ENOMEM, "image %d yields more than max_keypoints = %d keypoints (ties at the response threshold are all kept)", i, f->max_kp);
      for (const orb::Rec& r : rec) {
        // the final list re-uses the candidate staging buffers
        f->h_cand_xy[(size_t)i * cap + nf] = f->h_sel_xy[o + (size_t)r.idx];
        f->h_cand_sc[(size_t)i * cap + nf] = (uint8_t)l;
        if (kp_response) kp_response[(size_t)i * f->max_kp + nf] = r.response;
        ++nf;
      }
      o += (size_t)c;
    }
    f->h_n[i] = nf;
    n_keypoints[i] = nf;
    max_fin = nf > max_fin ? nf : max_fin;
    CB_CUDA(cudaMemcpyAsync(f->sel_xy + (size_t)i * cap, f->h_cand_xy + (size_t)i * cap, (size_t)nf * sizeof(unsigned), cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(f->sel_level + (size_t)i * cap, f->h_cand_sc + (size_t)i * cap, (size_t)nf, cudaMemcpyHostToDevice, st));
  }
  CB_CUDA(cudaMemcpyAsync(f->n_sel, f->h_n, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, st));
  if (max_fin > 0) {
    describe_kernel<<<dim3((max_fin + 63) / 64, n), 64, 0, st>>>(f->pyr, f->blur, pd, f->um, f->sel_xy, f->sel_level, f->n_sel, (int)cap, f->max_kp,
                                                                  f->o_xy, f->o_size, f->o_angle, f->o_octave, f->o_desc);
    CB_LAUNCH_CHECK();
    for (int i = 0; i < n; ++i) {
      const size_t o = (size_t)i * f->max_kp, c = (size_t)f->h_n[i];
      CB_CUDA(cudaMemcpyAsync(kp_xy + 2 * o, f->o_xy + 2 * o, c * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
      if (kp_size) CB_CUDA(cudaMemcpyAsync(kp_size + o, f->o_size + o, c * sizeof(float), cudaMemcpyDeviceToHost, st));
      if (kp_angle) CB_CUDA(cudaMemcpyAsync(kp_angle + o, f->o_angle + o, c * sizeof(float), cudaMemcpyDeviceToHost, st));
      if (kp_octave) CB_CUDA(cudaMemcpyAsync(kp_octave + o, f->o_octave + o, c * sizeof(int), cudaMemcpyDeviceToHost, st));
      CB_CUDA(cudaMemcpyAsync(descriptors + 32 * o, f->o_desc + 32 * o, c * 32, cudaMemcpyDeviceToHost, st));
    }
  }
  CB_CUDA(cudaEventRecord(f->ev[1], st));
  CB_CUDA(cb::sync_stream(st));
  cudaEventElapsedTime(&f->last_orb_ms, f->ev[0], f->ev[1]);
  return CB_OK;
}

float cb_features_last_orb_ms(const cb_features* f) { return f ? f->last_orb_ms : -1.f; }

// parity tests: the pyramid (what = 0), FAST score map (1) or blurred pyramid (2) of image 0 of the last cb_features_orb call
// ([total pixels] bytes, levels back to back); returns the number of bytes or a negative error
int64_t cb_features_debug_read(cb_features* f, int what, uint8_t* out, int64_t max_bytes) {
  if (!f || !out) return cb::fail(CB_EINVAL, "NULL argument to cb_features_debug_read");
  if (max_bytes < (int64_t)f->pd.total_px) return cb::fail(CB_EINVAL, "buffer too small: need %u bytes", f->pd.total_px);
  cb::DeviceGuard g(f->device);
  const uint8_t* src = what == 0 ? f->pyr : (what == 1 ? f->score : f->blur);
  CB_CUDA(cudaMemcpy(out, src, f->pd.total_px, cudaMemcpyDeviceToHost));
  return (int64_t)f->pd.total_px;
}

}  // extern "C"
