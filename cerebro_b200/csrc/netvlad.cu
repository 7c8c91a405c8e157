// NetVLAD whole-image descriptor forward pass (sm_100a).
//
// Replaces (reference): HDF5ModelImageDescriptor.handle_req -> model.predict
// (scripts/whole_image_desc_compute_server.py:596-650), i.e. the MobileNet-v1 prefix listed in
// scripts/keras.models/model.json followed by NetVLADLayer.call (scripts/predict_utils.py:36-64).
//
// Data layout in HBM: activations NHWC fp16 ([n*h*w][C], channels contiguous), two ping-pong
// buffers sized for max_batch frames; pointwise weights fp16 [Cout][Cin] (K-major), BN folded on
// the host (cerebro_b200/keras_weights.py); depthwise / stem / VLAD weights fp32.
//
// Kernels
//   conv1_kernel        u8 image -> (x-128)*2/255 -> 3x3 s2 (pad bottom/right) -> +b, ReLU6 -> fp16
//   dw_kernel<S>        3x3 depthwise (s1 'same' | pad bottom/right + s2 'valid') +b, ReLU6; HBM-bound
//   pw_gemm_kernel      1x1 conv as GEMM [pixels x Cin] x [Cin x Cout] on tcgen05: A and B tiles land in
//                       128B/64B-swizzled shared memory via TMA, one elected thread issues
//                       tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16) accumulating fp32 in
//                       TMEM, four epilogue warps read TMEM with tcgen05.ld, add bias, ReLU6, store fp16
//   pw_simt_kernel      the same contraction on CUDA cores; used for layers tcgen05 tiling does not
//                       cover (Cin not a multiple of 32) and as an on-device cross-check (CB_PW_SIMT=1)
//   vlad_assign / vlad_aggregate / vlad_norm   soft-assignment softmax, residual aggregation
//                       (x + C, PLUS as in predict_utils.py:47), intra-norm, flatten K-major, L2 norm
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <vector>

namespace {

using cb::FULL;

__device__ __forceinline__ float relu6(float x) { return fminf(fmaxf(x, 0.f), 6.f); }

// ---------------------------------------------------------------------------------------------
// stem: u8 -> normalise -> 3x3 stride-2 conv (pad bottom/right) -> ReLU6 -> fp16 NHWC
// thread = (4 consecutive output pixels along x, group of 8 output channels): the 3 x 9 input patch is
// converted once and every weight vector fetched from shared memory is reused for the 4 pixels.
// ---------------------------------------------------------------------------------------------
constexpr int kPX = 4;

template <int CIN>
__global__ void __launch_bounds__(256) conv1_kernel(const uint8_t* __restrict__ img, int n, int H, int W, int Ho, int Wo,
                                                   const float* __restrict__ w /*[3][3][CIN][32]*/,
                                                   const float* __restrict__ b, __half* __restrict__ out) {
  __shared__ __align__(16) float sw[9 * CIN * 32];
  __shared__ float sb[32];
  for (int i = threadIdx.x; i < 9 * CIN * 32; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 32) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const int Wg = (Wo + kPX - 1) / kPX;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * Ho * Wg * 4;
  if (t >= total) return;
  const int cg = (int)(t & 3);
  long long p = t >> 2;
  const int xg = (int)(p % Wg);
  p /= Wg;
  const int y = (int)(p % Ho);
  const int f = (int)(p / Ho);
  const int x0 = xg * kPX;
  float acc[kPX][8];
#pragma unroll
  for (int q = 0; q < kPX; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[q][j] = sb[cg * 8 + j];
  const uint8_t* base = img + (size_t)f * H * W * CIN;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = 2 * y + ky;
    if (iy >= H) continue;  // ZeroPadding2D((0,1),(0,1)): the pad row contributes 0 AFTER normalisation
    float v[2 * kPX + 1][CIN];
#pragma unroll
    for (int cx = 0; cx < 2 * kPX + 1; ++cx) {
      const int ix = 2 * x0 + cx;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        v[cx][ci] = ix < W ? ((float)base[((size_t)iy * W + ix) * CIN + ci] - 128.f) * 2.0f / 255.f : 0.f;  // server.py:629
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * CIN + ci) * 32 + cg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + ((ky * 3 + kx) * CIN + ci) * 32 + cg * 8 + 4);
#pragma unroll
        for (int q = 0; q < kPX; ++q) {
          const float xv = v[2 * q + kx][ci];
          acc[q][0] = fmaf(xv, w0.x, acc[q][0]);
          acc[q][1] = fmaf(xv, w0.y, acc[q][1]);
          acc[q][2] = fmaf(xv, w0.z, acc[q][2]);
          acc[q][3] = fmaf(xv, w0.w, acc[q][3]);
          acc[q][4] = fmaf(xv, w1.x, acc[q][4]);
          acc[q][5] = fmaf(xv, w1.y, acc[q][5]);
          acc[q][6] = fmaf(xv, w1.z, acc[q][6]);
          acc[q][7] = fmaf(xv, w1.w, acc[q][7]);
        }
      }
  }
  __half* orow = out + (((size_t)f * Ho + y) * Wo + x0) * 32 + cg * 8;
#pragma unroll
  for (int q = 0; q < kPX; ++q) {
    if (x0 + q >= Wo) break;
    __half2 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(relu6(acc[q][2 * j]), relu6(acc[q][2 * j + 1]));
    *reinterpret_cast<uint4*>(orow + (size_t)q * 32) = *reinterpret_cast<uint4*>(h);
  }
}

// ---------------------------------------------------------------------------------------------
// depthwise 3x3: thread = (4 consecutive output pixels along x, 8 channels); the 72 weights live in
// registers, input columns are shared between neighbouring outputs.
// ---------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(256) dw_kernel(const __half* __restrict__ in, int n, int H, int W, int C, int Ho, int Wo,
                                                const float* __restrict__ w /*[3][3][C]*/, const float* __restrict__ b,
                                                __half* __restrict__ out) {
  const int cgs = C >> 3;
  const int Wg = (Wo + kPX - 1) / kPX;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * Ho * Wg * cgs;
  if (t >= total) return;
  const int cg = (int)(t % cgs);
  long long p = t / cgs;
  const int xg = (int)(p % Wg);
  p /= Wg;
  const int y = (int)(p % Ho);
  const int f = (int)(p / Ho);
  const int x0 = xg * kPX;
  constexpr int P = (S == 1) ? 1 : 0;  // 'same' for stride 1; ZeroPadding2D((0,1),(0,1)) + 'valid' for stride 2
  constexpr int NC = (kPX - 1) * S + 3;  // input columns feeding kPX outputs
  float acc[kPX][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + cg * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + cg * 8 + 4));
#pragma unroll
    for (int q = 0; q < kPX; ++q) {
      acc[q][0] = b0.x, acc[q][1] = b0.y, acc[q][2] = b0.z, acc[q][3] = b0.w;
      acc[q][4] = b1.x, acc[q][5] = b1.y, acc[q][6] = b1.z, acc[q][7] = b1.w;
    }
  }
  const __half* base = in + (size_t)f * H * W * C + cg * 8;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = y * S + ky - P;
    if (iy < 0 || iy >= H) continue;
    float wk[3][8];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float* wp = w + (ky * 3 + kx) * C + cg * 8;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp)), w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
      wk[kx][0] = w0.x, wk[kx][1] = w0.y, wk[kx][2] = w0.z, wk[kx][3] = w0.w;
      wk[kx][4] = w1.x, wk[kx][5] = w1.y, wk[kx][6] = w1.z, wk[kx][7] = w1.w;
    }
    const __half* rowp = base + (size_t)iy * W * C;
#pragma unroll
    for (int cx = 0; cx < NC; ++cx) {
      const int ix = x0 * S + cx - P;
      if (ix < 0 || ix >= W) continue;
      const uint4 raw = *reinterpret_cast<const uint4*>(rowp + (size_t)ix * C);
      const __half2* hv = reinterpret_cast<const __half2*>(&raw);
      float xv[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f2 = __half22float2(hv[j]);
        xv[2 * j] = f2.x;
        xv[2 * j + 1] = f2.y;
      }
#pragma unroll
      for (int q = 0; q < kPX; ++q) {
        const int kx = cx - q * S;  // compile-time after unrolling
        if (kx >= 0 && kx < 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[q][j] = fmaf(xv[j], wk[kx][j], acc[q][j]);
        }
      }
    }
  }
  __half* orow = out + (((size_t)f * Ho + y) * Wo + x0) * C + cg * 8;
#pragma unroll
  for (int q = 0; q < kPX; ++q) {
    if (x0 + q >= Wo) break;
    __half2 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(relu6(acc[q][2 * j]), relu6(acc[q][2 * j + 1]));
    *reinterpret_cast<uint4*>(orow + (size_t)q * C) = *reinterpret_cast<uint4*>(h);
  }
}

// ---------------------------------------------------------------------------------------------
// pointwise conv on tcgen05.  One CTA = one 128-pixel x N_TILE output tile.
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : TMEM allocator + MMA issuer (one elected lane)
//   warps 2-5: epilogue (TMEM lanes 32*(warp%4) .. +31)
// ---------------------------------------------------------------------------------------------
constexpr int kGemmThreads = 192;
constexpr int kStages = 2;

template <int N_TILE, int KB>
struct GemmSmem {
  static constexpr int kABytes = 128 * KB * 2;
  static constexpr int kBBytes = N_TILE * KB * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*align slack*/ + 64 /*barriers*/;
};

template <int N_TILE, int KB>
__global__ void __launch_bounds__(kGemmThreads) pw_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB,
                                                              const float* __restrict__ bias, __half* __restrict__ out,
                                                              int M_total, int N_total, int K,
                                                              const __half* __restrict__ residual, int relu) {
  // residual != NULL: out = acc + bias + residual (MobileNetV2's Add); relu == 0: linear bottleneck (no ReLU6)
  using SM = GemmSmem<N_TILE, KB>;
  constexpr int SWZ = KB * 2;  // bytes per tile row = swizzle span (128 or 64)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * SM::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int n0 = blockIdx.y * N_TILE;
  const int nkb = K / KB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, N_TILE);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t round = kb / kStages;
        mbar_wait(&empty_bar[s], (round & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], SM::kStageBytes);
        uint8_t* sa = smem + s * SM::kStageBytes;
        tma_load_2d(&tmA, &full_bar[s], sa, kb * KB, m0);
        tma_load_2d(&tmB, &full_bar[s], sa + SM::kABytes, kb * KB, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=f16, both K-major, N, M=128
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N_TILE >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        const uint32_t round = kb / kStages;
        mbar_wait(&full_bar[s], round & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * SM::kStageBytes);
        const uint32_t b_addr = a_addr + SM::kABytes;
#pragma unroll
        for (int k = 0; k < KB / 16; ++k) {
          const uint64_t ad = make_kmajor_desc<SWZ>(a_addr + k * 32);
          const uint64_t bd = make_kmajor_desc<SWZ>(b_addr + k * 32);
          umma_f16(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < M_total;
    __half* orow = out + (size_t)row * N_total + n0;
#pragma unroll 1
    for (int c = 0; c < N_TILE; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + j));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c + j + 4));
          float f[8] = {__uint_as_float(v[j + 0]) + b0.x, __uint_as_float(v[j + 1]) + b0.y, __uint_as_float(v[j + 2]) + b0.z,
                        __uint_as_float(v[j + 3]) + b0.w, __uint_as_float(v[j + 4]) + b1.x, __uint_as_float(v[j + 5]) + b1.y,
                        __uint_as_float(v[j + 6]) + b1.z, __uint_as_float(v[j + 7]) + b1.w};
          if (residual) {  // kernel-uniform
            const uint4 rr = *reinterpret_cast<const uint4*>(residual + (size_t)row * N_total + n0 + c + j);
            const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 t = __half22float2(rh[i]);
              f[2 * i] += t.x;
              f[2 * i + 1] += t.y;
            }
          }
          if (relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = relu6(f[i]);
          }
          __half2 h[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
          *reinterpret_cast<uint4*>(orow + c + j) = *reinterpret_cast<uint4*>(h);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, N_TILE);
  }
}

// ---------------------------------------------------------------------------------------------
// fused depthwise -> pointwise block, persistent: one CTA per SM loops over 128-pixel output tiles.
//   warp 0     : TMA producer of the pointwise-weight (B) tiles; the whole weight matrix stays resident in
//                shared memory when it fits (blocks 1-5), else a 2-stage ring of 64-channel K blocks
//   warp 1     : TMEM allocator + tcgen05.mma issuer (two TMEM accumulator stages when COUT <= 256, so the
//                epilogue of tile i overlaps the MMAs of tile i+1)
//   warps 2-5  : epilogue (TMEM -> +bias -> ReLU6 -> fp16 -> global)
//   warps 6-21 : depthwise producers: compute the 3x3 depthwise (+bias, ReLU6) for the tile straight from the
//                previous layer's activation in global memory and write it, as fp16, into the 128B/64B-swizzled
//                K-major A tile in shared memory (the layout TMA would have produced); fence.proxy.async +
//                mbarrier hand each K block to the MMA warp; a ring of A slots lets them run ahead.
// The depthwise output never touches HBM: per block the activation traffic drops from
// (dw in + dw out + pw in + pw out) to (dw in + pw out).
// ---------------------------------------------------------------------------------------------
constexpr int kProdWarps = 16;
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kFusedThreads = 192 + kProdThreads;  // 704
constexpr int kASlots = 3;

template <int CIN, int COUT>
struct FusedSmem {
  static constexpr int KB = CIN < 64 ? CIN : 64;  // channels per K block (32 only for the first block)
  static constexpr int NKB = CIN / KB;
  static constexpr int kABytes = 128 * KB * 2;
  static constexpr int kBBytes = COUT * KB * 2;                    // one K block of the weights
  static constexpr bool kResident = NKB * kBBytes <= 128 * 1024;   // whole [COUT][CIN] matrix in smem
  static constexpr int kBS = kResident ? NKB : 2;                  // B slots
  static constexpr int kAcc = COUT <= 256 ? 2 : 1;                 // TMEM accumulator stages
  static constexpr int kTotal = kASlots * kABytes + kBS * kBBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int CIN, int COUT, int S>
__global__ void __launch_bounds__(kFusedThreads, 1) dwpw_kernel(const __half* __restrict__ in, int H, int W, int Ho, int Wo,
                                                               const float* __restrict__ dw_w /*[3][3][CIN]*/,
                                                               const float* __restrict__ dw_b,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const float* __restrict__ bias, __half* __restrict__ out,
                                                               int M_total, int n_tiles) {
  using SM = FusedSmem<CIN, COUT>;
  constexpr int KB = SM::KB, NKB = SM::NKB;
  constexpr int SWZ = KB * 2;
  constexpr int N_MMA = COUT > 256 ? 256 : COUT;  // columns per tcgen05.mma (two halves when COUT = 512)
  constexpr int kAcc = SM::kAcc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kASlots * SM::kABytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_b + SM::kBS * SM::kBBytes);
  uint64_t* a_empty = a_full + kASlots;
  uint64_t* b_full = a_empty + kASlots;   // [2] (ring) or [1] (resident)
  uint64_t* b_empty = b_full + 2;
  uint64_t* tmem_full_bar = b_empty + 2;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kASlots; ++s) {
      mbar_init(&a_full[s], kProdWarps);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, kAcc * COUT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      if (SM::kResident) {
        mbar_expect_tx(&b_full[0], NKB * SM::kBBytes);
        for (int kb = 0; kb < NKB; ++kb) {
          uint8_t* sb = smem_b + kb * SM::kBBytes;
          tma_load_2d(&tmB, &b_full[0], sb, kb * KB, 0);
          if (COUT > 256) tma_load_2d(&tmB, &b_full[0], sb + 256 * KB * 2, kb * KB, 256);
        }
      } else {
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
          for (int kb = 0; kb < NKB; ++kb, ++g) {
            const int s = g & 1;
            mbar_wait(&b_empty[s], ((g >> 1) & 1) ^ 1);
            mbar_expect_tx(&b_full[s], SM::kBBytes);
            uint8_t* sb = smem_b + s * SM::kBBytes;
            tma_load_2d(&tmB, &b_full[s], sb, kb * KB, 0);
            if (COUT > 256) tma_load_2d(&tmB, &b_full[s], sb + 256 * KB * 2, kb * KB, 256);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N_MMA >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t g = 0, ti = 0;
      if (SM::kResident) mbar_wait(&b_full[0], 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const int acc = ti % kAcc;
        mbar_wait(&tmem_empty_bar[acc], ((ti / kAcc) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * COUT;
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          const int sa = g % kASlots;
          const int sb = SM::kResident ? kb : (int)(g & 1);
          if (!SM::kResident) mbar_wait(&b_full[sb], (g >> 1) & 1);
          mbar_wait(&a_full[sa], (g / kASlots) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + sa * SM::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + sb * SM::kBBytes);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            const uint64_t ad = make_kmajor_desc<SWZ>(a_addr + k * 32);
#pragma unroll
            for (int h = 0; h < COUT / N_MMA; ++h) {
              const uint64_t bd = make_kmajor_desc<SWZ>(b_addr + h * (256 * KB * 2) + k * 32);
              umma_f16(d_tmem + h * 256, ad, bd, idesc, (kb | k) ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[sa]);
          if (!SM::kResident) umma_commit(&b_empty[sb]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    uint32_t ti = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti % kAcc;
      mbar_wait(&tmem_full_bar[acc], (ti / kAcc) & 1);
      tc_fence_after();
      const long long row = (long long)tile * 128 + q * 32 + lane;
      const bool row_ok = row < M_total;
      __half* orow = out + (size_t)row * COUT;
#pragma unroll 1
      for (int c = 0; c < COUT; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + acc * COUT + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c + j));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c + j + 4));
            __half2 h[4];
            h[0] = __floats2half2_rn(relu6(__uint_as_float(v[j + 0]) + b0.x), relu6(__uint_as_float(v[j + 1]) + b0.y));
            h[1] = __floats2half2_rn(relu6(__uint_as_float(v[j + 2]) + b0.z), relu6(__uint_as_float(v[j + 3]) + b0.w));
            h[2] = __floats2half2_rn(relu6(__uint_as_float(v[j + 4]) + b1.x), relu6(__uint_as_float(v[j + 5]) + b1.y));
            h[3] = __floats2half2_rn(relu6(__uint_as_float(v[j + 6]) + b1.z), relu6(__uint_as_float(v[j + 7]) + b1.w));
            *reinterpret_cast<uint4*>(orow + c + j) = *reinterpret_cast<uint4*>(h);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  } else {
    // ---- depthwise producers: task = (pixel of the tile, group of 8 channels of the current K block)
    const int pt = threadIdx.x - 192;        // 0..511
    constexpr int CGB = KB / 8;              // channel groups per K block (8, or 4 for the 32-channel block)
    constexpr int PIX_PER_PASS = kProdThreads / CGB;
    constexpr int PASSES = 128 / PIX_PER_PASS;  // 2 (1 for the 32-channel block)
    constexpr int P = (S == 1) ? 1 : 0;
    const int cgl = pt % CGB;
    const int pl = pt / CGB;
    // tap offsets (in elements) relative to the pixel's top-left tap; identical for every thread
    int tap_off[9];
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) tap_off[t9] = ((t9 / 3) * W + (t9 % 3)) * CIN;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      // geometry once per tile (not per K block): pointer to the top-left tap and a 9-bit validity mask
      const __half* pb[PASSES];
      uint32_t vmask[PASSES];
#pragma unroll
      for (int u = 0; u < PASSES; ++u) {
        const int m = tile * 128 + u * PIX_PER_PASS + pl;
        const bool ok = m < M_total;
        const int mm = ok ? m : 0;
        const int x = mm % Wo;
        const int t2 = mm / Wo;
        const int y = t2 % Ho;
        const int f = t2 / Ho;
        const int iy0 = y * S - P, ix0 = x * S - P;
        uint32_t rmask = 0, cmask = 0;
#pragma unroll
        for (int k3 = 0; k3 < 3; ++k3) {
          rmask |= (uint32_t)(iy0 + k3 >= 0 && iy0 + k3 < H) << k3;
          cmask |= (uint32_t)(ix0 + k3 >= 0 && ix0 + k3 < W) << k3;
        }
        uint32_t vm = 0;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
          if ((rmask >> ky) & 1u) vm |= cmask << (3 * ky);
        vmask[u] = ok ? vm : 0u;
        pb[u] = in + ((long long)(f * H + iy0) * W + ix0) * CIN;  // may point before the frame: only valid taps are read
      }
      for (int kb = 0; kb < NKB; ++kb, ++g) {
        const int sa = g % kASlots;
        mbar_wait(&a_empty[sa], ((g / kASlots) & 1) ^ 1);
        uint8_t* at = smem_a + sa * SM::kABytes;
        const int ch0 = kb * KB + cgl * 8;
        float acc[PASSES][8];
        {
          const float4 bb0 = __ldg(reinterpret_cast<const float4*>(dw_b + ch0));
          const float4 bb1 = __ldg(reinterpret_cast<const float4*>(dw_b + ch0 + 4));
#pragma unroll
          for (int u = 0; u < PASSES; ++u) {
            acc[u][0] = bb0.x, acc[u][1] = bb0.y, acc[u][2] = bb0.z, acc[u][3] = bb0.w;
            acc[u][4] = bb1.x, acc[u][5] = bb1.y, acc[u][6] = bb1.z, acc[u][7] = bb1.w;
          }
        }
        // interior fast path: when every tap of every pixel handled by this warp is inside the image (the common
        // case) the loads carry no predicates and no zero-fill selects
        bool all_in = true;
#pragma unroll
        for (int u = 0; u < PASSES; ++u) all_in = all_in && (vmask[u] == 0x1FFu);
        all_in = __all_sync(FULL, all_in);
#define CB_DW_TAPS(LOAD)                                                                                   \
  _Pragma("unroll") for (int t9 = 0; t9 < 9; ++t9) {                                                       \
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(dw_w + t9 * CIN + ch0));                       \
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(dw_w + t9 * CIN + ch0 + 4));                   \
    uint4 raw[PASSES];                                                                                     \
    _Pragma("unroll") for (int u = 0; u < PASSES; ++u) raw[u] = LOAD;                                      \
    _Pragma("unroll") for (int u = 0; u < PASSES; ++u) {                                                   \
      const __half2* hv = reinterpret_cast<const __half2*>(&raw[u]);                                       \
      const float2 v0 = __half22float2(hv[0]), v1 = __half22float2(hv[1]), v2 = __half22float2(hv[2]),    \
                   v3 = __half22float2(hv[3]);                                                             \
      acc[u][0] = fmaf(v0.x, w0.x, acc[u][0]);                                                             \
      acc[u][1] = fmaf(v0.y, w0.y, acc[u][1]);                                                             \
      acc[u][2] = fmaf(v1.x, w0.z, acc[u][2]);                                                             \
      acc[u][3] = fmaf(v1.y, w0.w, acc[u][3]);                                                             \
      acc[u][4] = fmaf(v2.x, w1.x, acc[u][4]);                                                             \
      acc[u][5] = fmaf(v2.y, w1.y, acc[u][5]);                                                             \
      acc[u][6] = fmaf(v3.x, w1.z, acc[u][6]);                                                             \
      acc[u][7] = fmaf(v3.y, w1.w, acc[u][7]);                                                             \
    }                                                                                                      \
  }
        if (all_in) {
          CB_DW_TAPS(*reinterpret_cast<const uint4*>(pb[u] + tap_off[t9] + ch0))
        } else {
          CB_DW_TAPS((((vmask[u] >> t9) & 1u) ? *reinterpret_cast<const uint4*>(pb[u] + tap_off[t9] + ch0)
                                              : make_uint4(0u, 0u, 0u, 0u)))
        }
#undef CB_DW_TAPS
        bool pok[PASSES];
#pragma unroll
        for (int u = 0; u < PASSES; ++u) pok[u] = tile * 128 + u * PIX_PER_PASS + pl < M_total;
#pragma unroll
        for (int u = 0; u < PASSES; ++u) {
          const int r = u * PIX_PER_PASS + pl;
          __half2 h[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            h[j] = pok[u] ? __floats2half2_rn(relu6(acc[u][2 * j]), relu6(acc[u][2 * j + 1])) : __floats2half2_rn(0.f, 0.f);
          // K-major swizzled tile: row r, 16-byte chunk cgl -> chunk ^ (row bits), Swizzle<3,4,3> (128B) / <2,4,3> (64B)
          const int chunk = (SWZ == 128) ? (cgl ^ (r & 7)) : (cgl ^ ((r >> 1) & 3));
          *reinterpret_cast<uint4*>(at + r * SWZ + chunk * 16) = *reinterpret_cast<uint4*>(h);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[sa]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAcc * COUT);
  }
}

// ---------------------------------------------------------------------------------------------
// fused depthwise -> pointwise, TMA-halo version (the default).
// The M tile is a 16 x 8 patch of output pixels of one frame; its (15 S + 3) x (7 S + 3) input halo of one
// K block of channels is fetched by ONE 4-D TMA box (channels, x, y, frame) into a shared-memory ring --
// out-of-image taps are zero-filled by the TMA unit, which IS the layer's zero padding -- so the depthwise
// producers never wait on global memory: they read their 9 taps with conflict-free LDS.128, accumulate with
// packed FFMA2 and write the swizzled K-major A tile.  Roles:
//   warps 0-15  : depthwise producers                                  warp 24     : TMA of the halos
//   warp 25     : TMA of the pointwise weights (resident or ring)      warp 26     : TMEM allocator + tcgen05.mma issuer
//   warps 16-23 : epilogue (bias + ReLU6 + fp16 / q15 -> 2 KB staging -> TMA store), two groups of four warps: with
//                 two accumulator stages group g drains stage g (alternate tiles), with one stage (COUT = 512) group g
//                 drains columns [256 g, 256 g + 256) of every tile.  [Round 1 had one group: ncu showed the 16 producer
//                 warps stalled ~50-70 % on a_empty behind it -- the kernel was epilogue-bound.]
// ---------------------------------------------------------------------------------------------
// epilogue warps: two groups of four (one per TMEM lane quarter) where the epilogue is the longer pole (COUT = 2 CIN), one
// group where producers and epilogue balance (COUT = CIN: 736 threads leave 80 registers per thread instead of 72)
constexpr int halo_threads(int epi_warps) { return 32 * (kProdWarps + epi_warps + 3); }
constexpr int kTH = 16, kTW = 8;

// MODE bits of the halo kernel (the "precise" descriptor path):
//   kInQ   the block input is stored as q15 (unsigned 15-bit fixed point of the ReLU6 range: u = round(x * 32767 / 6));
//          a PRMT drops u into the mantissa of a float in [2, 4) (f = 2 + u * 2^-14) -- on the ALU pipe, where the fp16
//          decode (HADD2.F32) competes with the FFMA2 taps for the FMA pipe -- and the affine map back to x is folded
//          into the depthwise weights and bias on the host (w' = w * 2^14 * 6 / 32767, b' = b - 2 * sum w').  TMA's
//          zero fill of the padding decodes to f = 2, i.e. x = 0.
//   kSplit the depthwise output (MMA A operand) and the pointwise weights (B operand) are each kept as fp16 hi + lo;
//          three accumulating MMAs  A_hi B_hi + A_lo B_hi + A_hi B_lo  (A_lo is stored negated: h - x is one FHADD;
//          the instruction descriptor's a_negate bit flips it back)
//   kOutQ  the epilogue writes q15 instead of fp16
//   kSplitA only the A operand is split (blocks with 512 output channels: a second copy of their 32 KB weight stages
//          does not fit beside the halo ring): A_hi B + A_lo B
constexpr int kInQ = 1, kSplit = 2, kOutQ = 4, kSplitA = 8;
constexpr int kSatPackMaxCin = 64;  // fused blocks up to this many input channels use q15_pack_sat (their bias is uploaded / 6)
// q15: u = round(x * 32767 / 6)
constexpr double kQ15DecodeW = 16384.0 * 6.0 / 32767.0;  // (f - 2) -> x

template <int CIN, int COUT, int S, bool SPLIT = false, bool SPLIT_B = SPLIT, int EW = 8>
struct HaloCfg {
  static constexpr int KB = 32;  // channels per K block (= per halo box): 64-byte pixels, 64-byte swizzled A/B rows
  static constexpr int NKB = CIN / KB;
  static constexpr int SWZ = KB * 2;
  static constexpr int HH = (kTH - 1) * S + 3, HW = (kTW - 1) * S + 3;
  static constexpr int kHaloTx = HH * HW * KB * 2;                 // bytes one box delivers
  static constexpr int kHaloBytes = (kHaloTx + 127) / 128 * 128;   // ring stride (TMA destinations are 128B aligned)
  static constexpr int kATile = 128 * KB * 2;
  static constexpr int kABytes = (SPLIT ? 2 : 1) * kATile;          // [hi | -lo]
  static constexpr int kBTile = COUT * KB * 2;
  static constexpr int kBBytes = (SPLIT_B ? 2 : 1) * kBTile;        // [hi | lo]
  static constexpr bool kResident = NKB * kBBytes <= 64 * 1024;    // whole [COUT][CIN] matrix in smem, else a ring
  static constexpr int kBS = kResident ? NKB : ((S == 2 && kBBytes >= 32 * 1024) ? 2 : 3);
  static constexpr int kStrip = S == 2 ? 2 : 4;                    // output rows per producer task
  static constexpr int kTasks = (KB / 8) * kTW * (kTH / kStrip);   // producer tasks per K block (128 or 256)
  static constexpr int kTeams = kProdThreads / kTasks;             // K blocks in production concurrently (4 or 2)
  static constexpr int kTeamWarps = kTasks / 32;
  static constexpr int kASmin = (SPLIT && S == 2) ? 2 : 3;          // the big stride-2 halos need the room
  static constexpr int kAS = kTeams > kASmin ? kTeams : kASmin;
  static constexpr int kStageBytes = EW * 32 * 64;                 // epilogue staging: 32 pixels x 32 channels per warp
  static constexpr int kBudget = 232448 - 1024 - 512 - kStageBytes;
  static constexpr int kFit = (kBudget - kBS * kBBytes - kAS * kABytes) / kHaloBytes;
  // halo stages beyond the number of teams are the prefetch depth: a team's next box is already in flight while it
  // works on the current one
  static constexpr int kHS = kFit > 8 ? 8 : kFit;
  static constexpr int kAcc = COUT <= 256 ? 2 : 1;
  // a team waits on ring slot g % K with parity (g / K) & 1: that is only well defined while it cannot run more than one
  // phase ahead of the slot, i.e. while the number of teams does not exceed the ring depth
  static_assert(kTeams <= kAS && kTeams <= kHS, "more producer teams than ring slots");
  static_assert(!SPLIT_B || COUT <= 256, "the weight split is built for the blocks up to 256 output channels");
  static_assert(!SPLIT_B || SPLIT, "B is only split together with A");
  static constexpr int kTotal = kAS * kABytes + kBS * kBBytes + kStageBytes + kHS * kHaloBytes + 1024 + 512;
};

// two fp32 accumulators (acc * 32767 / 6 + bias * 32767 / 6 already applied by one FFMA2) -> two q15 values: clamp to
// [0, 32767] on the ALU pipe, round to nearest-even through the 2^23 magic add, keep the low 16 bits of each.  Used where
// the depthwise producers already fill the FMA pipe (CIN >= 128 and the stem).
__device__ __forceinline__ uint32_t q15_pack(unsigned long long y) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(y));
  lo = fminf(fmaxf(lo, 0.f), 32767.f);
  hi = fminf(fmaxf(hi, 0.f), 32767.f);
  const unsigned long long m = fadd2(pack_f32x2(lo, hi), pack_f32x2(8388608.f, 8388608.f));
  uint32_t a, b;
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(m));
  return __byte_perm(a, b, 0x5410);
}
constexpr float kQ15Scale = 32767.0f / 6.0f;  // x -> u

// Two fp32 accumulators -> two q15 values in five instructions: FFMA.SAT maps ReLU6 onto [0, 1] (acc / 6 + bias / 6, the
// saturation IS the clamp), a second FFMA scales by 32767 and adds 2^23 so that the round-to-nearest-even integer lands in
// the low mantissa bits, one PRMT packs the two low halves.  Used by the blocks whose epilogue is the longer pole (CIN <= 64:
// 32 -> 64 went 263 -> 239 us per 64 frames); it moves the clamp from the ALU pipe to the FMA pipe, which the depthwise
// taps of the wider blocks already fill (128 -> 128 got 25 us SLOWER with it), so those keep q15_pack.  `bias` is pre-scaled
// by 32767 / 6; the 1 / 32767 that maps it onto [0, 1] is folded into the first FFMA's constant.
__device__ __forceinline__ uint32_t q15_pack_sat(float a0, float a1, float b0_6, float b1_6) {
  const float t0 = __saturatef(fmaf(a0, 1.0f / 6.0f, b0_6)), t1 = __saturatef(fmaf(a1, 1.0f / 6.0f, b1_6));  // b*_6 = bias / 6
  const float m0 = fmaf(t0, 32767.f, 8388608.f), m1 = fmaf(t1, 32767.f, 8388608.f);
  return __byte_perm(__float_as_uint(m0), __float_as_uint(m1), 0x5410);
}

// eight q15 values (one 16-byte chunk) -> four packed float pairs f = 2 + u * 2^-14
__device__ __forceinline__ void q15_decode8(const uint4& raw, unsigned long long v[4]) {
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t lo = __byte_perm(w[j], 0x40000000u, 0x7104), hi = __byte_perm(w[j], 0x40000000u, 0x7324);
    asm("mov.b64 %0, {%1, %2};" : "=l"(v[j]) : "r"(lo), "r"(hi));
  }
}

// ReLU6 of two fp32 values -> fp16 hi pair and the NEGATED fp16 lo pair (h - x: one mixed-precision FHADD per value)
__device__ __forceinline__ void relu6_split_h2(unsigned long long v, uint32_t& hi_out, uint32_t& nlo_out) {
  float x0, x1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(v));
  x0 = fminf(fmaxf(x0, 0.f), 6.f);
  x1 = fminf(fmaxf(x1, 0.f), 6.f);
  const __half2 h = __floats2half2_rn(x0, x1);
  const unsigned short h0 = __half_as_ushort(__low2half(h)), h1 = __half_as_ushort(__high2half(h));
  float n0, n1;
  asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(n0) : "h"(h0), "f"(x0));
  asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(n1) : "h"(h1), "f"(x1));
  const __half2 nl = __floats2half2_rn(n0, n1);
  hi_out = *reinterpret_cast<const uint32_t*>(&h);
  nlo_out = *reinterpret_cast<const uint32_t*>(&nl);
}

template <int CIN, int COUT, int S, int MODE, int EW>
__global__ void __launch_bounds__(halo_threads(EW), 1) dwpw_halo_kernel(const __grid_constant__ CUtensorMap tmH,
                                                                   const float* __restrict__ dw_w /*[3][3][CIN]*/,
                                                                   const float* __restrict__ dw_b,
                                                                   const __grid_constant__ CUtensorMap tmB,
                                                                   const float* __restrict__ bias,
                                                                   const __grid_constant__ CUtensorMap tmO, int tiles_x,
                                                                   int tiles_per_frame, int n_tiles) {
  constexpr bool INQ = (MODE & kInQ) != 0, SPLIT_B = (MODE & kSplit) != 0, SPLIT = SPLIT_B || (MODE & kSplitA) != 0,
                 OUTQ = (MODE & kOutQ) != 0;
  using SM = HaloCfg<CIN, COUT, S, SPLIT, SPLIT_B, EW>;
  constexpr int kEpiWarps = EW;
  constexpr int KB = SM::KB, NKB = SM::NKB, SWZ = SM::SWZ;
  constexpr int kAS = SM::kAS, kHS = SM::kHS, kAcc = SM::kAcc;
  constexpr int N_MMA = COUT > 256 ? 256 : COUT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + kAS * SM::kABytes;
  uint8_t* smem_o = smem_b + SM::kBS * SM::kBBytes;  // epilogue staging, 2 KB per epilogue warp (1024-aligned)
  uint8_t* smem_h = smem_o + SM::kStageBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_h + kHS * SM::kHaloBytes);
  uint64_t* a_empty = a_full + 4;
  uint64_t* h_full = a_empty + 4;  // [8]
  uint64_t* h_empty = h_full + 8;
  uint64_t* b_full = h_empty + 8;  // [kBS <= 3] (ring) or [1] (resident)
  uint64_t* b_empty = b_full + 4;
  uint64_t* tmem_full_bar = b_empty + 4;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Roles by PHYSICAL warp id.  The SM sub-partition arbiter favours the highest warp id among eligible warps, so the
  // warps whose few instructions gate everybody else (MMA issue, TMA issue) sit at the top, the epilogue next, and the
  // always-eligible depthwise producers at the bottom.  (With the MMA issuer as warp 1 the producers and the epilogue
  // both showed up waiting on it in ncu: a_empty 27-48 %, tmem_full 6-22 % of their samples.)
  constexpr int kWarpEpi0 = kProdWarps, kWarpHalo = kProdWarps + kEpiWarps, kWarpB = kWarpHalo + 1, kWarpMma = kWarpHalo + 2;

  if (warp == kWarpB && lane == 0) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < 4; ++s) {
      mbar_init(&a_full[s], SM::kTeamWarps);
      mbar_init(&a_empty[s], 1);
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int s = 0; s < 8; ++s) {
      mbar_init(&h_full[s], 1);
      mbar_init(&h_empty[s], SM::kTeamWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], (kAcc == 2 || EW == 4) ? 4 : kEpiWarps);
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == kWarpMma) tmem_alloc(tmem_ptr, kAcc * COUT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kWarpB) {
    if (lane == 0) {
      if (SM::kResident) {
        mbar_expect_tx(&b_full[0], NKB * SM::kBBytes);
        for (int kb = 0; kb < NKB; ++kb) {
          uint8_t* sb = smem_b + kb * SM::kBBytes;
          tma_load_2d(&tmB, &b_full[0], sb, kb * KB, 0);
          if (COUT > 256) tma_load_2d(&tmB, &b_full[0], sb + 256 * KB * 2, kb * KB, 256);
          if (SPLIT_B) tma_load_2d(&tmB, &b_full[0], sb + SM::kBTile, kb * KB, COUT);  // lo rows follow the hi rows
        }
      } else {
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
          for (int kb = 0; kb < NKB; ++kb, ++g) {
            const int s = g % SM::kBS;
            mbar_wait(&b_empty[s], ((g / SM::kBS) & 1) ^ 1);
            mbar_expect_tx(&b_full[s], SM::kBBytes);
            uint8_t* sb = smem_b + s * SM::kBBytes;
            tma_load_2d(&tmB, &b_full[s], sb, kb * KB, 0);
            if (COUT > 256) tma_load_2d(&tmB, &b_full[s], sb + 256 * KB * 2, kb * KB, 256);
            if (SPLIT_B) tma_load_2d(&tmB, &b_full[s], sb + SM::kBTile, kb * KB, COUT);
          }
        }
      }
    }
  } else if (warp == kWarpMma) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(N_MMA >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t g = 0, ti = 0;
      if (SM::kResident) mbar_wait(&b_full[0], 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const int acc = ti % kAcc;
        mbar_wait(&tmem_empty_bar[acc], ((ti / kAcc) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * COUT;
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          const int sa = g % kAS;
          const int sb = SM::kResident ? kb : (int)(g % SM::kBS);
          if (!SM::kResident) mbar_wait(&b_full[sb], (g / SM::kBS) & 1);
          mbar_wait(&a_full[sa], (g / kAS) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + sa * SM::kABytes);
          const uint32_t b_addr = smem_u32(smem_b + sb * SM::kBBytes);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            const uint64_t ad = make_kmajor_desc<SWZ>(a_addr + k * 32);
#pragma unroll
            for (int h = 0; h < COUT / N_MMA; ++h) {
              const uint64_t bd = make_kmajor_desc<SWZ>(b_addr + h * (256 * KB * 2) + k * 32);
              umma_f16(d_tmem + h * 256, ad, bd, idesc, (kb | k) ? 1u : 0u);
              if (SPLIT)  // + A_lo B_hi (A_lo is stored negated: a_negate, descriptor bit 13)
                umma_f16(d_tmem + h * 256, make_kmajor_desc<SWZ>(a_addr + SM::kATile + k * 32), bd, idesc | (1u << 13), 1u);
              if (SPLIT_B)  // + A_hi B_lo
                umma_f16(d_tmem + h * 256, ad, make_kmajor_desc<SWZ>(b_addr + SM::kBTile + k * 32), idesc, 1u);
            }
          }
          umma_commit(&a_empty[sa]);
          if (!SM::kResident) umma_commit(&b_empty[sb]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp >= kWarpEpi0 && warp < kWarpHalo) {
    // ---- epilogue: TMEM -> bias + ReLU6 -> fp16 / q15 -> 64B-swizzled staging rows -> TMA store of a (32 ch, 8 x, 4 y) box;
    //      the TMA unit clips the parts of the patch that lie outside the frame
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int grp = (warp - kWarpEpi0) >> 2;     // epilogue group
    const uint32_t stage = smem_u32(smem_o + (warp - kWarpEpi0) * 2048);
    const uint32_t srow = stage + lane * 64;
    const int sw = (lane >> 1) & 3;      // Swizzle<2,4,3>: 16-byte chunk ^= row bits [1,3)
    constexpr bool kByTile = kAcc == 2 && EW == 8;      // two groups, two accumulator stages: group g drains stage g
    constexpr bool kByCols = kAcc == 1 && EW == 8;      // two groups, one stage: group g owns columns [256 g, 256 g + 256)
    constexpr int C_BEGIN_STEP = kByCols ? 256 : 0;
    constexpr int C_COUNT = kByCols ? 256 : COUT;
    uint32_t ti = kByTile ? (uint32_t)grp : 0u;
    const int tile_step = kByTile ? 2 : 1;
    for (int tile = blockIdx.x + (kByTile ? grp : 0) * (int)gridDim.x; tile < n_tiles; tile += tile_step * (int)gridDim.x, ti += tile_step) {
      const int acc = ti % kAcc;
      const int f = tile / tiles_per_frame, rem = tile - f * tiles_per_frame;
      const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
      mbar_wait(&tmem_full_bar[acc], (ti / kAcc) & 1);
      tc_fence_after();
      const int c_begin = grp * C_BEGIN_STEP;
#pragma unroll 1
      for (int c = c_begin; c < c_begin + C_COUNT; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + acc * COUT + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        // Blocks whose epilogue is the longer pole (CIN <= 64: few K blocks per tile) convert all 32 channels of the lane's
        // pixel BEFORE waiting for the staging rows of the previous store; the others stream (fewer live registers).
        constexpr bool kLateWait = CIN <= kSatPackMaxCin;
        uint4 o[kLateWait ? 4 : 1];
        if (!kLateWait) {
          if (lane == 0) bulk_wait_read0();  // the previous store has finished reading the staging rows
          __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const ulonglong2 b0 = __ldg(reinterpret_cast<const ulonglong2*>(bias + c + j));
          const ulonglong2 b1 = __ldg(reinterpret_cast<const ulonglong2*>(bias + c + j + 4));
          uint4& oo = o[kLateWait ? (j >> 3) : 0];
          if (OUTQ && CIN <= kSatPackMaxCin) {  // `bias` is pre-divided by 6
            const float* bf0 = reinterpret_cast<const float*>(&b0);
            const float* bf1 = reinterpret_cast<const float*>(&b1);
            oo.x = q15_pack_sat(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]), bf0[0], bf0[1]);
            oo.y = q15_pack_sat(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]), bf0[2], bf0[3]);
            oo.z = q15_pack_sat(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]), bf1[0], bf1[1]);
            oo.w = q15_pack_sat(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]), bf1[2], bf1[3]);
          } else if (OUTQ) {  // `bias` is pre-multiplied by 32767 / 6
            const unsigned long long sc = pack_f32x2(kQ15Scale, kQ15Scale);
            oo.x = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1])), sc, b0.x));
            oo.y = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), sc, b0.y));
            oo.z = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])), sc, b1.x));
            oo.w = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])), sc, b1.y));
          } else {
            oo.x = relu6_pack_h2(fadd2(pack_f32x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1])), b0.x));
            oo.y = relu6_pack_h2(fadd2(pack_f32x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), b0.y));
            oo.z = relu6_pack_h2(fadd2(pack_f32x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])), b1.x));
            oo.w = relu6_pack_h2(fadd2(pack_f32x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])), b1.y));
          }
          if (!kLateWait) sts128(srow + (((j >> 3) ^ sw) << 4), oo);
        }
        if (kLateWait) {
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) sts128(srow + ((j ^ sw) << 4), o[kLateWait ? j : 0]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmO, stage, c, txi * kTW, tyi * kTH + q * 4, f);
          bulk_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (lane == 0) bulk_wait0();
  } else if (warp == kWarpHalo) {
    if (lane == 0) {
      constexpr int P = (S == 1) ? 1 : 0;
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int f = tile / tiles_per_frame, rem = tile - f * tiles_per_frame;
        const int tyi = rem / tiles_x, txi = rem - tyi * tiles_x;
        const int x0 = txi * kTW * S - P, y0 = tyi * kTH * S - P;
        for (int kb = 0; kb < NKB; ++kb, ++g) {
          const int hs = g % kHS;
          mbar_wait(&h_empty[hs], ((g / kHS) & 1) ^ 1);
          mbar_expect_tx(&h_full[hs], SM::kHaloTx);
          tma_load_4d(&tmH, &h_full[hs], smem_h + hs * SM::kHaloBytes, kb * KB, x0, y0, f);
        }
      }
    }
  } else {
    // ---- depthwise producers.  Task = (group of 8 channels, column tx, strip of kStrip output rows): the (kStrip-1)S+3 input rows
    //      of a column are each read once per kx and feed up to three output rows from registers; the 3 weights of
    //      the kx column are fetched once per task.  A team of kTasks threads produces one K block; the kTeams teams
    //      work on consecutive K blocks concurrently.
    const int pt = threadIdx.x;  // 0..511
    constexpr int CGB = KB / 8;
    constexpr int PXB = KB * 2;  // bytes per halo pixel
    constexpr int kStrip = SM::kStrip;
    constexpr int NR = (kStrip - 1) * S + 3;
    const int team = pt / SM::kTasks, tt = pt % SM::kTasks;
    const int cgl = tt % CGB;
    const int tx = (tt / CGB) % kTW;
    const int strip = tt / (CGB * kTW);
    const uint32_t hoff = (uint32_t)((strip * kStrip * S * SM::HW + tx * S) * PXB + cgl * 16);
    uint32_t aoff[kStrip];
#pragma unroll
    for (int i = 0; i < kStrip; ++i) {
      const int r = (strip * kStrip + i) * kTW + tx;
      const int chunk = (SWZ == 128) ? (cgl ^ (r & 7)) : (cgl ^ ((r >> 1) & 3));
      aoff[i] = (uint32_t)(r * SWZ + chunk * 16);
    }
    const uint32_t smem_h_u32 = smem_u32(smem_h), smem_a_u32 = smem_u32(smem_a);
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t total = (uint32_t)my_tiles * NKB;
    for (uint32_t g = team; g < total; g += SM::kTeams) {
      const int kb = g % NKB;
      const int sa = g % kAS, hs = g % kHS;
      const int ch0 = kb * KB + cgl * 8;
      unsigned long long acc[kStrip][4];
      {
        const ulonglong2 bb0 = __ldg(reinterpret_cast<const ulonglong2*>(dw_b + ch0));
        const ulonglong2 bb1 = __ldg(reinterpret_cast<const ulonglong2*>(dw_b + ch0 + 4));
#pragma unroll
        for (int i = 0; i < kStrip; ++i) acc[i][0] = bb0.x, acc[i][1] = bb0.y, acc[i][2] = bb1.x, acc[i][3] = bb1.y;
      }
      mbar_wait(&h_full[hs], (g / kHS) & 1);
      const uint32_t hb = smem_h_u32 + hs * SM::kHaloBytes + hoff;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        unsigned long long w[3][4];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const ulonglong2 w0 = __ldg(reinterpret_cast<const ulonglong2*>(dw_w + (ky * 3 + kx) * CIN + ch0));
          const ulonglong2 w1 = __ldg(reinterpret_cast<const ulonglong2*>(dw_w + (ky * 3 + kx) * CIN + ch0 + 4));
          w[ky][0] = w0.x, w[ky][1] = w0.y, w[ky][2] = w1.x, w[ky][3] = w1.y;
        }
#pragma unroll
        for (int ir = 0; ir < NR; ++ir) {
          const uint4 raw = lds128(hb + (ir * SM::HW + kx) * PXB);
          unsigned long long v[4];
          if (INQ) {
            q15_decode8(raw, v);
          } else {
            const __half2* hv = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 fv = __half22float2(hv[j]);
              v[j] = pack_f32x2(fv.x, fv.y);
            }
          }
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int d = ir - ky;
            if (d >= 0 && d % S == 0 && d / S < kStrip) {
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[d / S][j] = ffma2(v[j], w[ky][j], acc[d / S][j]);
            }
          }
        }
      }
      mbar_wait(&a_empty[sa], ((g / kAS) & 1) ^ 1);
      const uint32_t at = smem_a_u32 + sa * SM::kABytes;
#pragma unroll
      for (int i = 0; i < kStrip; ++i) {
        uint4 o;
        if (SPLIT) {
          uint4 l;
          relu6_split_h2(acc[i][0], o.x, l.x);
          relu6_split_h2(acc[i][1], o.y, l.y);
          relu6_split_h2(acc[i][2], o.z, l.z);
          relu6_split_h2(acc[i][3], o.w, l.w);
          sts128(at + SM::kATile + aoff[i], l);
        } else {
          o.x = relu6_pack_h2(acc[i][0]);
          o.y = relu6_pack_h2(acc[i][1]);
          o.z = relu6_pack_h2(acc[i][2]);
          o.w = relu6_pack_h2(acc[i][3]);
        }
        sts128(at + aoff[i], o);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&a_full[sa]);
        mbar_arrive(&h_empty[hs]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAcc * COUT);
  }
}

// ---------------------------------------------------------------------------------------------
// stem on tcgen05: implicit GEMM  [pixels x 32 (27 taps, zero padded)] x [32 x 32 output channels].
//   warp 0    : TMEM allocator + tcgen05.mma issuer
//   warps 1-4 : im2col producers (one output pixel per thread: 27 byte loads -> exact fp16 integers x-128
//               -> 64B-swizzled K-major row), then the epilogue of the same tile
// The 2/255 input scale is folded into the weights, which are split hi + lo (two accumulating MMAs) so the
// only rounding left is the fp16 store of the result.  kConvTiles tiles per CTA amortise the set-up.
// ---------------------------------------------------------------------------------------------
constexpr int kConvThreads = 160;
constexpr int kConvTiles = 4;

// Four image bytes -> four exact fp16 values (b - 128) as two half2 words: 0x6400 | b is the fp16 number 1024 + b, and
// (1024 + b) - 1152 is exact.  One PRMT + one HSUB2 per pair instead of a byte load + I2F per value.
__device__ __forceinline__ void bytes4_to_centered_h2(uint32_t w, uint32_t& lo, uint32_t& hi) {
  const __half2 off = __floats2half2_rn(1152.f, 1152.f);
  uint32_t a = __byte_perm(w, 0x64646464u, 0x4140), b = __byte_perm(w, 0x64646464u, 0x4342);
  __half2 ha = __hsub2(*reinterpret_cast<__half2*>(&a), off), hb = __hsub2(*reinterpret_cast<__half2*>(&b), off);
  lo = *reinterpret_cast<uint32_t*>(&ha);
  hi = *reinterpret_cast<uint32_t*>(&hb);
}

// The 9 bytes (3 pixels x 3 channels) of one image row that feed output column x of the 3-channel stem, as bytes 0..8 of
// (v0, v1, v2 & 0xff), through three aligned 32-bit loads.  `rowp` = start of the row, 4-byte aligned; `avail` = bytes
// readable from rowp (to the end of the frame).  Taps beyond the right border read as 128 (-> 0 after centring).
__device__ __forceinline__ void load_row9(const uint8_t* __restrict__ rowp, int x, int W, long long avail, uint32_t& v0,
                                          uint32_t& v1, uint32_t& v2) {
  const int o = 6 * x, a = o & ~3, sh = (o - a) * 8;
  uint32_t w0, w1, w2;
  if ((long long)a + 12 <= avail) {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(rowp + a);
    w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
  } else {  // the last few bytes of a frame: assemble the window from single bytes
    uint32_t b[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) b[i] = (a + i < avail) ? (uint32_t)__ldg(rowp + a + i) : 128u;
    w0 = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
    w1 = b[4] | (b[5] << 8) | (b[6] << 16) | (b[7] << 24);
    w2 = b[8] | (b[9] << 8) | (b[10] << 16) | (b[11] << 24);
  }
  v0 = __funnelshift_r(w0, w1, sh);
  v1 = __funnelshift_r(w1, w2, sh);
  v2 = w2 >> sh;
  if (2 * x + 2 >= W) {  // third pixel is the zero padding column (ZeroPadding2D ((0,1),(0,1)))
    v1 = (v1 & 0x0000ffffu) | 0x80800000u;
    v2 = 0x80u;
  }
}

template <int CIN, bool OUTQ>
__global__ void __launch_bounds__(kConvThreads) conv1_tc_kernel(const uint8_t* __restrict__ img, int n, int H, int W, int Ho, int Wo,
                                                               const __half* __restrict__ w_hi /*[32][32] K-major*/,
                                                               const __half* __restrict__ w_lo, const float* __restrict__ bias,
                                                               __half* __restrict__ out, int M_total) {
  __shared__ __align__(1024) uint8_t sA[128 * 64];
  __shared__ __align__(1024) uint8_t sB[2][32 * 64];
  __shared__ uint64_t a_full, tmem_full;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&a_full, 4);
    mbar_init(&tmem_full, 1);
    fence_barrier_init();
  }
  __syncwarp();  // warp 0 must be converged for the .sync.aligned allocation
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  if (threadIdx.x >= 32) {  // weights: 32 rows x 4 chunks of 16 B per matrix, 64B-swizzled like the A tile
    const int t = threadIdx.x - 32, r = t >> 2, c = t & 3;
    const int cs = c ^ ((r >> 1) & 3);
    *reinterpret_cast<uint4*>(sB[0] + r * 64 + cs * 16) = *reinterpret_cast<const uint4*>(w_hi + r * 32 + c * 8);
    *reinterpret_cast<uint4*>(sB[1] + r * 64 + cs * 16) = *reinterpret_cast<const uint4*>(w_lo + r * 32 + c * 8);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  // rows start 4-byte aligned: the 3-channel producers can use aligned 32-bit loads
  const bool wide_loads = CIN == 3 && ((W * 3) & 3) == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0;
  // the three image rows (9 bytes each) under output pixel m; issued one tile ahead so that the global-load latency hides
  // behind the previous tile's MMA + epilogue (the kernel was latency-bound: 0.27 of HBM with ~50 % of the warps resident)
  uint32_t rv[3][3];
  bool have_rv = false;
  auto load_rows = [&](long long m, uint32_t (&o)[3][3]) {
    const int x = (int)(m % Wo);
    const long long t2 = m / Wo;
    const int y = (int)(t2 % Ho);
    const long long frame_bytes = (long long)H * W * 3;
    const uint8_t* base = img + (size_t)(t2 / Ho) * frame_bytes;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * y + ky;
      if (iy < H) {
        const long long roff = (long long)iy * W * 3;
        load_row9(base + roff, x, W, frame_bytes - roff, o[ky][0], o[ky][1], o[ky][2]);
      } else {  // zero padding row
        o[ky][0] = o[ky][1] = 0x80808080u;
        o[ky][2] = 0x80u;
      }
    }
  };

  for (int it = 0; it < kConvTiles; ++it) {
    const long long m0 = ((long long)blockIdx.x * kConvTiles + it) * 128;
    if (m0 >= M_total) break;  // CTA-uniform
    if (warp == 0) {
      if (lane == 0) {
        mbar_wait(&a_full, it & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA);
#pragma unroll
        for (int hl = 0; hl < 2; ++hl) {
          const uint32_t b_addr = smem_u32(sB[hl]);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            umma_f16(tmem_base, make_kmajor_desc<64>(a_addr + k * 32), make_kmajor_desc<64>(b_addr + k * 32), idesc,
                     (hl | k) ? 1u : 0u);
        }
        umma_commit(&tmem_full);
      }
    } else {
      const int r = threadIdx.x - 32;  // row of the tile = TMEM lane
      const long long m = m0 + r;
      const bool ok = m < M_total;
      __half hv[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) hv[i] = __ushort_as_half((unsigned short)0);
      if (CIN == 3 && wide_loads) {
        if (ok) {  // 9 aligned 32-bit loads + byte permutes instead of 27 byte loads + conversions
          if (!have_rv) load_rows(m, rv);  // else: prefetched while the previous tile's MMA and epilogue ran
          // 27 bytes (+ one padding byte) as 7 words, tap order (ky, kx, ci)
          uint32_t wd[7];
          wd[0] = rv[0][0];
          wd[1] = rv[0][1];
          wd[2] = (rv[0][2] & 0xffu) | (rv[1][0] << 8);
          wd[3] = (rv[1][0] >> 24) | (rv[1][1] << 8);
          wd[4] = (rv[1][1] >> 24) | ((rv[1][2] & 0xffu) << 8) | (rv[2][0] << 16);
          wd[5] = (rv[2][0] >> 16) | (rv[2][1] << 16);
          wd[6] = (rv[2][1] >> 16) | ((rv[2][2] & 0xffu) << 16) | 0x80000000u;
          uint32_t* hw = reinterpret_cast<uint32_t*>(hv);
#pragma unroll
          for (int i = 0; i < 7; ++i) bytes4_to_centered_h2(wd[i], hw[2 * i], hw[2 * i + 1]);
        }
      } else if (ok) {
        const int x = (int)(m % Wo);
        const long long t2 = m / Wo;
        const int y = (int)(t2 % Ho);
        const uint8_t* base = img + (size_t)(t2 / Ho) * H * W * CIN;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const int iy = 2 * y + ky;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int ix = 2 * x + kx;
            if (iy < H && ix < W) {
#pragma unroll
              for (int ci = 0; ci < CIN; ++ci)
                hv[(ky * 3 + kx) * CIN + ci] = __int2half_rn((int)__ldg(base + ((size_t)iy * W + ix) * CIN + ci) - 128);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cs = c ^ ((r >> 1) & 3);
        *reinterpret_cast<uint4*>(sA + r * 64 + cs * 16) = *reinterpret_cast<uint4*>(&hv[c * 8]);
      }
      fence_proxy_async();
      tc_fence_before();  // orders the previous tile's tcgen05.ld before the next MMA overwrites TMEM
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full);
      have_rv = false;
      if (CIN == 3 && wide_loads && it + 1 < kConvTiles && m + 128 < M_total) {
        load_rows(m + 128, rv);
        have_rv = true;
      }
      // ---- epilogue of this tile
      mbar_wait(&tmem_full, it & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16), v);
      // TMEM lane quarter of this warp is (warp & 3); its rows are 32*(warp&3)+lane
      const long long mr = m0 + (warp & 3) * 32 + lane;
      if (mr < M_total) {
        __half* orow = out + (size_t)mr * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + j));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + j + 4));
          if (OUTQ) {  // q15 output for the halo kernel; `bias` is pre-multiplied by 32767 / 6
            const unsigned long long sc = pack_f32x2(kQ15Scale, kQ15Scale);
            uint4 o;
            o.x = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1])), sc, pack_f32x2(b0.x, b0.y)));
            o.y = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), sc, pack_f32x2(b0.z, b0.w)));
            o.z = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])), sc, pack_f32x2(b1.x, b1.y)));
            o.w = q15_pack(ffma2(pack_f32x2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])), sc, pack_f32x2(b1.z, b1.w)));
            *reinterpret_cast<uint4*>(orow + j) = o;
          } else {
            __half2 h[4];
            h[0] = __floats2half2_rn(relu6(__uint_as_float(v[j + 0]) + b0.x), relu6(__uint_as_float(v[j + 1]) + b0.y));
            h[1] = __floats2half2_rn(relu6(__uint_as_float(v[j + 2]) + b0.z), relu6(__uint_as_float(v[j + 3]) + b0.w));
            h[2] = __floats2half2_rn(relu6(__uint_as_float(v[j + 4]) + b1.x), relu6(__uint_as_float(v[j + 5]) + b1.y));
            h[3] = __floats2half2_rn(relu6(__uint_as_float(v[j + 6]) + b1.z), relu6(__uint_as_float(v[j + 7]) + b1.w));
            *reinterpret_cast<uint4*>(orow + j) = *reinterpret_cast<uint4*>(h);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// CUDA-core version of the same contraction: thread = (pixel, 8 output channels)
__global__ void __launch_bounds__(256) pw_simt_kernel(const __half* __restrict__ in, const __half* __restrict__ w /*[N][K]*/,
                                                     const float* __restrict__ bias, __half* __restrict__ out,
                                                     long long M_total, int N, int K,
                                                     const __half* __restrict__ residual, int relu) {
  const int ngs = N >> 3;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M_total * ngs) return;
  const int ng = (int)(t % ngs);
  const long long m = t / ngs;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = bias[ng * 8 + j];
  const __half* a = in + (size_t)m * K;
  for (int k = 0; k < K; k += 8) {
    const uint4 ar = *reinterpret_cast<const uint4*>(a + k);
    const __half2* ah = reinterpret_cast<const __half2*>(&ar);
    float av[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(ah[i]);
      av[2 * i] = f.x;
      av[2 * i + 1] = f.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 wr = __ldg(reinterpret_cast<const uint4*>(w + (size_t)(ng * 8 + j) * K + k));
      const __half2* wh = reinterpret_cast<const __half2*>(&wr);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(wh[i]);
        acc[j] = fmaf(av[2 * i], f.x, acc[j]);
        acc[j] = fmaf(av[2 * i + 1], f.y, acc[j]);
      }
    }
  }
  if (residual) {
    const uint4 rr = *reinterpret_cast<const uint4*>(residual + (size_t)m * N + ng * 8);
    const __half2* rh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(rh[i]);
      acc[2 * i] += t.x;
      acc[2 * i + 1] += t.y;
    }
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = relu6(acc[j]);
  }
  __half2 h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
  *reinterpret_cast<uint4*>(out + (size_t)m * N + ng * 8) = *reinterpret_cast<uint4*>(h);
}

// ---------------------------------------------------------------------------------------------
// NetVLAD head (predict_utils.py:36-64), K = 16 clusters
// ---------------------------------------------------------------------------------------------
constexpr int kK = 16;

// warp per pixel: s = x W + b, a = softmax_K(s).  Shared memory holds W transposed [16][D] so that
// lanes (consecutive channel pairs) read consecutive words.
__global__ void __launch_bounds__(256) vlad_assign_kernel(const __half* __restrict__ x, long long P_total, int D,
                                                         const float* __restrict__ W /*[D][16]*/,
                                                         const float* __restrict__ bvec, float* __restrict__ a_out) {
  extern __shared__ float sWt[];  // [16][D]
  for (int i = threadIdx.x; i < D * kK; i += blockDim.x) {
    const int dd = i / kK, k = i % kK;
    sWt[k * D + dd] = W[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = (long long)blockIdx.x * 8 + warp;
  if (p >= P_total) return;
  float s[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) s[k] = 0.f;
  const __half2* xp = reinterpret_cast<const __half2*>(x + (size_t)p * D);
  for (int d2 = lane; d2 < (D >> 1); d2 += 32) {
    const float2 f = __half22float2(xp[d2]);
#pragma unroll
    for (int k = 0; k < kK; ++k) {
      const float2 w2 = *reinterpret_cast<const float2*>(sWt + k * D + 2 * d2);
      s[k] = fmaf(f.x, w2.x, fmaf(f.y, w2.y, s[k]));
    }
  }
#pragma unroll
  for (int k = 0; k < kK; ++k) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s[k] += __shfl_xor_sync(FULL, s[k], off);
    s[k] += bvec[k];
  }
  float mx = s[0];
#pragma unroll
  for (int k = 1; k < kK; ++k) mx = fmaxf(mx, s[k]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < kK; ++k) {
    s[k] = expf(s[k] - mx);
    sum += s[k];
  }
  const float inv = 1.f / sum;
  if (lane < kK) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < kK; ++k)
      if (lane == k) v = s[k];
    a_out[(size_t)p * kK + lane] = v * inv;
  }
}

// soft-assignment on tcgen05: s = x W + b as a GEMM [pixels x D] x [D x 16] (W split hi + lo, two accumulating
// MMAs per K step), epilogue = bias + softmax over the 16 clusters straight out of TMEM.
__global__ void __launch_bounds__(kGemmThreads) vlad_assign_tc_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                     const __grid_constant__ CUtensorMap tmWhi,
                                                                     const __grid_constant__ CUtensorMap tmWlo,
                                                                     const float* __restrict__ bvec, float* __restrict__ a_out,
                                                                     int P_total, int D) {
  constexpr int kA = 128 * 64 * 2, kB = 16 * 64 * 2, kStage = kA + 2 * kB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStage);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int nkb = D / 64;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmWhi);
    tma_prefetch_desc(&tmWlo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&empty_bar[s], ((kb / kStages) & 1) ^ 1);
        mbar_expect_tx(&full_bar[s], kStage);
        uint8_t* sa = smem + s * kStage;
        tma_load_2d(&tmX, &full_bar[s], sa, kb * 64, m0);
        tma_load_2d(&tmWhi, &full_bar[s], sa + kA, kb * 64, 0);
        tma_load_2d(&tmWlo, &full_bar[s], sa + kA + kB, kb * 64, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&full_bar[s], (kb / kStages) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * kStage);
#pragma unroll
        for (int hl = 0; hl < 2; ++hl) {
          const uint32_t b_addr = a_addr + kA + hl * kB;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, make_kmajor_desc<128>(a_addr + k * 32), make_kmajor_desc<128>(b_addr + k * 32), idesc,
                     (kb | hl | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16), v);
    if (row < P_total) {
      float sv[kK];
      float mx = -3.0e38f;
#pragma unroll
      for (int k = 0; k < kK; ++k) {
        sv[k] = __uint_as_float(v[k]) + __ldg(bvec + k);
        mx = fmaxf(mx, sv[k]);
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < kK; ++k) {
        sv[k] = expf(sv[k] - mx);
        sum += sv[k];
      }
      const float inv = 1.f / sum;
      float4* o = reinterpret_cast<float4*>(a_out + (size_t)row * kK);
#pragma unroll
      for (int k = 0; k < kK; k += 4) o[k / 4] = make_float4(sv[k] * inv, sv[k + 1] * inv, sv[k + 2] * inv, sv[k + 3] * inv);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

// ---------------------------------------------------------------------------------------------
// VLAD aggregation: partial V[k][d] = sum over a slice of the pixels of a[p][k] x[p][d].
// grid (kAggSplit, frames), 256 threads = 4 pixel lanes x 64 channel groups: a thread owns 8 channels x all 16
// clusters (64 packed-FFMA2 accumulators), reads each x value ONCE (one 16-byte load per pixel) and the pixel's 16
// soft-assignments through warp-uniform (broadcast) loads.  The 4 pixel lanes are summed through shared memory in a
// fixed order; the kAggSplit partials are summed in a fixed order by vlad_norm_kernel, which also adds the
// centre term (sum_p a[p][k]) * C[d][k]  (x + C, predict_utils.py:47).  Deterministic: no atomics.
// ---------------------------------------------------------------------------------------------
constexpr int kAggSplit = 8;
constexpr int kAggSmem = 128 * 64 * 4;

__global__ void __launch_bounds__(256) vlad_aggregate_kernel(const __half* __restrict__ x, const float* __restrict__ a,
                                                            int P, int D, float* __restrict__ Vp /*[frames][split][16][D]*/,
                                                            float* __restrict__ Ap /*[frames][split][16]*/) {
  extern __shared__ float red[];  // [128 accumulators][64 channel groups]; during the pixel loop: the slice's soft-assignments
  const int f = blockIdx.y, sp = blockIdx.x;
  const int dq = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int d0 = dq * 8;
  const bool active = d0 < D;
  const int p_begin = (int)((long long)P * sp / kAggSplit), p_end = (int)((long long)P * (sp + 1) / kAggSplit);
  const int n_s = p_end - p_begin;
  const __half* xf = x + (size_t)f * P * D + d0;
  const float* af = a + ((size_t)f * P + p_begin) * kK;
  // The slice's assignments a[p][16] go to shared memory once (coalesced), zero-padded by one unrolled step, so the pixel loop
  // has a single kind of global load -- the 16-byte x vector -- and keeps kAggUnroll of them in flight per thread (the
  // kernel is latency-bound: 8 warps per SM).  Slices too long for the buffer read the assignments from global memory.
  constexpr int kAggUnroll = 8;
  const bool a_in_smem = (size_t)(n_s + 4 * kAggUnroll) * kK * sizeof(float) <= (size_t)kAggSmem;
  if (a_in_smem) {
    const int n4 = n_s * (kK / 4);
    for (int i = threadIdx.x; i < n4 + 4 * kAggUnroll * (kK / 4); i += 256)
      reinterpret_cast<float4*>(red)[i] = i < n4 ? __ldg(reinterpret_cast<const float4*>(af) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  unsigned long long acc[kK][4];
#pragma unroll
  for (int k = 0; k < kK; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0ull;
  if (active) {
#pragma unroll 1
    for (int p = pl; p < n_s; p += 4 * kAggUnroll) {
      uint4 raw[kAggUnroll];
#pragma unroll
      for (int u = 0; u < kAggUnroll; ++u) {
        const int pp = p + 4 * u;
        raw[u] = pp < n_s ? __ldg(reinterpret_cast<const uint4*>(xf + (size_t)(p_begin + pp) * D)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int u = 0; u < kAggUnroll; ++u) {
        const int pp = p + 4 * u;
        float av[kK];
        if (a_in_smem) {  // padded with zeros: no bounds test
          const float4* ap = reinterpret_cast<const float4*>(red + (size_t)pp * kK);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = ap[q];
            av[4 * q] = t.x, av[4 * q + 1] = t.y, av[4 * q + 2] = t.z, av[4 * q + 3] = t.w;
          }
        } else {
          const float4* ap = reinterpret_cast<const float4*>(af + (size_t)(pp < n_s ? pp : 0) * kK);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 t = __ldg(ap + q);
            av[4 * q] = t.x, av[4 * q + 1] = t.y, av[4 * q + 2] = t.z, av[4 * q + 3] = t.w;
          }
        }
        const __half2* hv = reinterpret_cast<const __half2*>(&raw[u]);  // zero beyond the slice: contributes nothing
        unsigned long long v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 fv = __half22float2(hv[j]);
          v[j] = pack_f32x2(fv.x, fv.y);
        }
#pragma unroll
        for (int k = 0; k < kK; ++k) {
          const unsigned long long ak = pack_f32x2(av[k], av[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[k][j] = ffma2(ak, v[j], acc[k][j]);
        }
      }
    }
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kK) {  // sum_p a[p][k] of this slice, same order as before (ascending p)
    const int k = threadIdx.x - 64;
    float asum = 0.f;
    if (a_in_smem)
      for (int p = 0; p < n_s; ++p) asum += red[(size_t)p * kK + k];
    else
      for (int p = 0; p < n_s; ++p) asum += af[(size_t)p * kK + k];
    Ap[((size_t)f * kAggSplit + sp) * kK + k] = asum;
  }
  __syncthreads();  // the assignments are dead: `red` becomes the reduction buffer
  float r[kK * 8];
#pragma unroll
  for (int k = 0; k < kK; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) asm("mov.b64 {%0, %1}, %2;" : "=f"(r[k * 8 + 2 * j]), "=f"(r[k * 8 + 2 * j + 1]) : "l"(acc[k][j]));
  for (int lane_r = 1; lane_r < 4; ++lane_r) {  // fixed-order sum of the pixel lanes
    if (pl == lane_r) {
#pragma unroll
      for (int i = 0; i < kK * 8; ++i) red[i * 64 + dq] = r[i];
    }
    __syncthreads();
    if (pl == 0) {
#pragma unroll
      for (int i = 0; i < kK * 8; ++i) r[i] += red[i * 64 + dq];
    }
    __syncthreads();
  }
  if (pl == 0 && active) {
#pragma unroll
    for (int k = 0; k < kK; ++k) {
      float* vo = Vp + (((size_t)f * kAggSplit + sp) * kK + k) * D + d0;
      *reinterpret_cast<float4*>(vo) = make_float4(r[k * 8], r[k * 8 + 1], r[k * 8 + 2], r[k * 8 + 3]);
      *reinterpret_cast<float4*>(vo + 4) = make_float4(r[k * 8 + 4], r[k * 8 + 5], r[k * 8 + 6], r[k * 8 + 7]);
    }
  }
}

// CTA per frame: sum the pixel-split partials, add the centre term, intra-normalise each cluster over D, flatten
// K-major, L2-normalise (eps 1e-12 on the squared norm, like tf.nn.l2_normalize).  D <= 512.
// GhostVLADLayer (predict_utils.py:110-141): only the first k_out clusters are kept (`v[:,0:num_clusters,:]`, :133) -- the ghost
// clusters took part in the softmax, their rows never reach the norms or the output.
__global__ void __launch_bounds__(512) vlad_norm_kernel(const float* __restrict__ Vp, const float* __restrict__ Ap,
                                                       const float* __restrict__ Cc /*[D][16]*/, int D,
                                                       float* __restrict__ out, int k_out) {
  __shared__ float s_ss[kK];
  const int f = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 16 warps, one per cluster
  float asum = 0.f;
#pragma unroll
  for (int sp = 0; sp < kAggSplit; ++sp) asum += Ap[((size_t)f * kAggSplit + sp) * kK + warp];
  float v[16];  // D / 32 values per lane (D <= 512)
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = lane + 32 * i;
    float val = 0.f;
    if (d < D) {
#pragma unroll
      for (int sp = 0; sp < kAggSplit; ++sp) val += Vp[(((size_t)f * kAggSplit + sp) * kK + warp) * D + d];
      val += asum * Cc[(size_t)d * kK + warp];
    }
    v[i] = val;
    ss = fmaf(val, val, ss);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(FULL, ss, off);
  if (lane == 0) s_ss[warp] = ss;
  __syncthreads();
  const float inv_k = rsqrtf(fmaxf(s_ss[warp], 1e-12f));
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < kK; ++k) {
    const float ik = rsqrtf(fmaxf(s_ss[k], 1e-12f));
    if (k < k_out) tot += s_ss[k] * ik * ik;
  }
  const float inv_t = rsqrtf(fmaxf(tot, 1e-12f));
  if (warp >= k_out) return;
  float* o = out + (size_t)f * k_out * D + (size_t)warp * D;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int d = lane + 32 * i;
    if (d < D) o[d] = v[i] * inv_k * inv_t;
  }
}

struct Block {
  int C = 0, Cout = 0, stride = 1;
  int Hin = 0, Win = 0, Ho = 0, Wo = 0;
  bool has_pw = false;
  float* dw_w = nullptr;
  float* dw_b = nullptr;
  __half* pw_w = nullptr;  // [Cout][C]
  float* pw_b = nullptr;
  bool use_tc = false;
  int n_tile = 0, kb = 0;
  CUtensorMap tmA[2], tmB;  // A map per ping-pong buffer (the GEMM input may live in either)
  bool fused = false;  // depthwise folded into the tcgen05 GEMM producer
  CUtensorMap tmBf;    // weight map with a <=256-row box for the fused kernel
  bool halo = false;   // TMA-halo version of the fused kernel available
  CUtensorMap tmH[2];  // 4-D halo maps of the block's input, one per ping-pong buffer
  CUtensorMap tmO[2];  // 4-D store maps of the block's output (64 ch x 8 x 4 boxes, 128B swizzle)
  CUtensorMap tmBh;    // weight map with the halo kernel's K block
  // "precise" halo path (MODE bits of dwpw_halo_kernel)
  int mode = 0;              // kInQ | kSplit | kOutQ, fixed at create time
  int epi_warps = 4;         // epilogue warps of the COUT == CIN blocks (CB_DESC_EPI); the COUT = 2 CIN blocks always use 8
  bool runs_halo = false;    // this block is executed by dwpw_halo_kernel (decides the storage format of its input)
  float* dw_wq = nullptr;    // depthwise weights / bias with the q15 decode folded in
  float* dw_bq = nullptr;
  float* pw_bq = nullptr;    // pointwise bias * 32767 / 6 (q15 epilogue), or bias / 6 where the epilogue uses q15_pack_sat
  __half* pw_whl = nullptr;  // [2][Cout][C]: fp16 hi rows, then lo rows
  CUtensorMap tmBs;          // map over pw_whl
};

// One 1x1 convolution of the MobileNetV2 path (channel counts padded to multiples of 64 with zero weights / biases, so
// every layer runs on the tcgen05 GEMM; padded channels stay exactly zero through ReLU6, the linear bottlenecks and Add).
struct PwLayer {
  int C = 0, Cout = 0;  // padded
  __half* w = nullptr;  // [Cout][C] fp16, K-major B operand
  float* b = nullptr;   // [Cout]
  int n_tile = 0, kb = 0;
  bool use_tc = false;
  CUtensorMap tmA[3], tmB;  // A map per activation buffer
};

struct IrBlock {  // inverted-residual block: [expand 1x1 + ReLU6] -> depthwise 3x3 + ReLU6 -> project 1x1 (linear) [+ input]
  int Hin = 0, Win = 0, Ho = 0, Wo = 0, stride = 1, residual = 0;
  int Cin = 0, Cexp = 0, Cout = 0;  // padded
  bool has_expand = false;
  PwLayer expand, project;
  float* dw_w = nullptr;  // [3][3][Cexp]
  float* dw_b = nullptr;
};

}  // namespace

struct cb_descriptor {
  int device = 0, sm_count = 0;
  int rows = 0, cols = 0, chnls = 0, max_batch = 0;
  int H1 = 0, W1 = 0;
  float* conv1_w = nullptr;
  float* conv1_b = nullptr;
  __half* conv1_hi = nullptr;  // [32][32] K-major, scale 2/255 folded in, hi/lo split (tcgen05 stem)
  __half* conv1_lo = nullptr;
  float* conv1_bq = nullptr;   // stem bias * 32767 / 6 (q15 output)
  bool stem_q = false;         // the stem writes q15 (block 0 runs on the halo kernel in precise mode)
  bool fp16_legacy = false;    // CB_DESC_FP16=1: round-1 arithmetic (fp16 storage and operands everywhere)
  int split_blocks = 99;       // CB_DESC_SPLIT: leading blocks whose MMA operands are split hi + lo (default: all)
  std::vector<uint8_t> layer_q;  // per layer: output stored as q15
  std::vector<Block> blocks;
  bool v2 = false;            // MobileNetV2 prefix (cb_descriptor_create_v2): `ir` instead of `blocks`, three buffers
  std::vector<IrBlock> ir;
  int K = 16, D = 0, Hf = 0, Wf = 0;
  float* vlad_w = nullptr;
  float* vlad_b = nullptr;
  float* vlad_c = nullptr;
  __half* vlad_whi = nullptr;  // [16][D] K-major hi/lo split of the soft-assignment weights (tcgen05 path)
  __half* vlad_wlo = nullptr;
  CUtensorMap tmX[3], tmWhi, tmWlo;
  bool vlad_tc = false;
  __half* act[3] = {nullptr, nullptr, nullptr};  // ping-pong (+ a third buffer for MobileNetV2's skip connections)
  size_t act_elems = 0;
  float* assign = nullptr;
  float* Vraw = nullptr;
  uint8_t* img_dev = nullptr;
  float* out_dev = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // host-API uploads, overlapped with the forward pass chunk by chunk
  // pinned staging ring for pageable callers (a cv::Mat's data, Cerebro.cpp:243-256): two buffers of `stage_frames` frames;
  // the CPU copy of chunk i+1 into one overlaps the DMA + forward pass of chunk i from the other
  uint8_t* stage_host[2] = {nullptr, nullptr};
  cudaEvent_t stage_free[2] = {nullptr, nullptr};
  int stage_frames = 8;
  float* out_host = nullptr;  // pinned read-back buffer of cb_descriptor_compute_f64 (max_batch descriptors)
  std::vector<cudaEvent_t> ev_copy;
  bool force_simt = false;
  bool no_fuse = false;  // CB_NO_FUSE=1: separate depthwise + GEMM kernels
  int upload_chunk = 64;  // frames per upload / forward chunk of the host API (CB_DESC_CHUNK): batches above it overlap upload and compute chunk by chunk
  bool no_halo = false;  // CB_NO_HALO=1: fused kernel whose producers read the activations straight from global memory
  int stop_layer = -1;  // CB_DEBUG_STOP_LAYER: stop the forward pass after this layer (bring-up / parity tests)
  int last_buf = 0;                 // ping-pong buffer holding the most recent layer output
  std::vector<size_t> layer_elems;  // per-frame elements of layer l's output
};

namespace {

int upload_f32(float** dst, const float* src, size_t n) {
  CB_CUDA(cudaMalloc((void**)dst, n * sizeof(float)));
  CB_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return CB_OK;
}

template <int N_TILE, int KB>
int launch_gemm(const Block& b, long long M, int in_buf, __half* out, cudaStream_t st) {
  using SM = GemmSmem<N_TILE, KB>;
  auto kern = pw_gemm_kernel<N_TILE, KB>;
  CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)(b.Cout / N_TILE));
  kern<<<grid, kGemmThreads, SM::kTotal, st>>>(b.tmA[in_buf], b.tmB, b.pw_b, out, (int)M, b.Cout, b.C, nullptr, 1);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

bool fused_shape_supported(int cin, int cout, int stride) {
  return (cin == 32 && cout == 64 && stride == 1) || (cin == 64 && cout == 128 && stride == 2) ||
         (cin == 128 && cout == 128 && stride == 1) || (cin == 128 && cout == 256 && stride == 2) ||
         (cin == 256 && cout == 256 && stride == 1) || (cin == 256 && cout == 512 && stride == 2) ||
         (cin == 512 && cout == 512 && stride == 1);
}

template <int CIN, int COUT, int S>
int launch_fused(const Block& b, int n, int sm_count, const __half* in, __half* out, cudaStream_t st) {
  using SM = FusedSmem<CIN, COUT>;
  auto kern = dwpw_kernel<CIN, COUT, S>;
  CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  const long long M = (long long)n * b.Ho * b.Wo;
  const int n_tiles = (int)((M + 127) / 128);
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;  // persistent: one CTA per SM
  kern<<<grid, kFusedThreads, SM::kTotal, st>>>(in, b.Hin, b.Win, b.Ho, b.Wo, b.dw_w, b.dw_b, b.tmBf, b.pw_b, out, (int)M, n_tiles);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int run_fused(const Block& b, int n, int sm, const __half* in, __half* out, cudaStream_t st) {
  if (b.C == 32) return launch_fused<32, 64, 1>(b, n, sm, in, out, st);
  if (b.C == 64) return launch_fused<64, 128, 2>(b, n, sm, in, out, st);
  if (b.C == 128 && b.Cout == 128) return launch_fused<128, 128, 1>(b, n, sm, in, out, st);
  if (b.C == 128) return launch_fused<128, 256, 2>(b, n, sm, in, out, st);
  if (b.C == 256 && b.Cout == 256) return launch_fused<256, 256, 1>(b, n, sm, in, out, st);
  if (b.C == 256) return launch_fused<256, 512, 2>(b, n, sm, in, out, st);
  return launch_fused<512, 512, 1>(b, n, sm, in, out, st);
}

template <int CIN, int COUT, int S, int MODE, int EW>
int launch_halo_ew(const Block& b, int n, int sm_count, int in_buf, cudaStream_t st) {
  using SM = HaloCfg<CIN, COUT, S, (MODE & (kSplit | kSplitA)) != 0, (MODE & kSplit) != 0, EW>;
  auto kern = dwpw_halo_kernel<CIN, COUT, S, MODE, EW>;
  CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  const int tiles_x = (b.Wo + kTW - 1) / kTW, tiles_y = (b.Ho + kTH - 1) / kTH;
  const int per_frame = tiles_x * tiles_y;
  const int n_tiles = n * per_frame;
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;  // persistent: one CTA per SM
  kern<<<grid, halo_threads(EW), SM::kTotal, st>>>(b.tmH[in_buf], (MODE & kInQ) ? b.dw_wq : b.dw_w, (MODE & kInQ) ? b.dw_bq : b.dw_b,
                                               (MODE & kSplit) ? b.tmBs : b.tmBh, (MODE & kOutQ) ? b.pw_bq : b.pw_b,
                                               b.tmO[in_buf ^ 1], tiles_x, per_frame, n_tiles);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

template <int CIN, int COUT, int S, int MODE>
int launch_halo_mode(const Block& b, int n, int sm_count, int in_buf, cudaStream_t st) {
  if constexpr (COUT == CIN) {  // balanced blocks: the epilogue width is a run-time choice (CB_DESC_EPI), default one group
    if (b.epi_warps == 8) return launch_halo_ew<CIN, COUT, S, MODE, 8>(b, n, sm_count, in_buf, st);
    return launch_halo_ew<CIN, COUT, S, MODE, 4>(b, n, sm_count, in_buf, st);
  } else {
    return launch_halo_ew<CIN, COUT, S, MODE, 8>(b, n, sm_count, in_buf, st);
  }
}

template <int CIN, int COUT, int S>
int launch_halo(const Block& b, int n, int sm_count, int in_buf, cudaStream_t st) {
  switch (b.mode) {
    case 0: return launch_halo_mode<CIN, COUT, S, 0>(b, n, sm_count, in_buf, st);
    case kInQ: return launch_halo_mode<CIN, COUT, S, kInQ>(b, n, sm_count, in_buf, st);
    case kInQ | kOutQ: return launch_halo_mode<CIN, COUT, S, kInQ | kOutQ>(b, n, sm_count, in_buf, st);
    default: break;
  }
  if constexpr (COUT <= 256) {
    if (b.mode == (kInQ | kSplit)) return launch_halo_mode<CIN, COUT, S, kInQ | kSplit>(b, n, sm_count, in_buf, st);
    if (b.mode == (kInQ | kSplit | kOutQ)) return launch_halo_mode<CIN, COUT, S, kInQ | kSplit | kOutQ>(b, n, sm_count, in_buf, st);
  } else {
    if (b.mode == (kInQ | kSplitA)) return launch_halo_mode<CIN, COUT, S, kInQ | kSplitA>(b, n, sm_count, in_buf, st);
    if (b.mode == (kInQ | kSplitA | kOutQ)) return launch_halo_mode<CIN, COUT, S, kInQ | kSplitA | kOutQ>(b, n, sm_count, in_buf, st);
  }
  return cb::fail(CB_EINVAL, "halo kernel mode %d is not built for %d -> %d channels", b.mode, CIN, COUT);
}

bool split_shape_supported(int cout) { return cout <= 256; }

int run_halo(const Block& b, int n, int sm, int in_buf, cudaStream_t st) {
  if (b.C == 32) return launch_halo<32, 64, 1>(b, n, sm, in_buf, st);
  if (b.C == 64) return launch_halo<64, 128, 2>(b, n, sm, in_buf, st);
  if (b.C == 128 && b.Cout == 128) return launch_halo<128, 128, 1>(b, n, sm, in_buf, st);
  if (b.C == 128) return launch_halo<128, 256, 2>(b, n, sm, in_buf, st);
  if (b.C == 256 && b.Cout == 256) return launch_halo<256, 256, 1>(b, n, sm, in_buf, st);
  if (b.C == 256) return launch_halo<256, 512, 2>(b, n, sm, in_buf, st);
  return launch_halo<512, 512, 1>(b, n, sm, in_buf, st);
}

int run_pw(cb_descriptor* d, const Block& b, long long M, int in_buf, __half* out, cudaStream_t st) {
  const __half* in = d->act[in_buf];
  if (b.use_tc && !d->force_simt) {
    if (b.kb == 64) {
      switch (b.n_tile) {
        case 64: return launch_gemm<64, 64>(b, M, in_buf, out, st);
        case 128: return launch_gemm<128, 64>(b, M, in_buf, out, st);
        default: return launch_gemm<256, 64>(b, M, in_buf, out, st);
      }
    } else {
      switch (b.n_tile) {
        case 64: return launch_gemm<64, 32>(b, M, in_buf, out, st);
        case 128: return launch_gemm<128, 32>(b, M, in_buf, out, st);
        default: return launch_gemm<256, 32>(b, M, in_buf, out, st);
      }
    }
  }
  const long long threads = M * (b.Cout / 8);
  pw_simt_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(in, b.pw_w, b.pw_b, out, M, b.Cout, b.C, nullptr, 1);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int run_stem(cb_descriptor* d, int n, const uint8_t* img_dev, int cur, cudaStream_t st) {
  if (!d->force_simt && !d->no_fuse) {
    const long long M = (long long)n * d->H1 * d->W1;
    const unsigned grid = (unsigned)((M + 128 * kConvTiles - 1) / (128 * kConvTiles));
    if (d->chnls == 1 && d->stem_q)
      conv1_tc_kernel<1, true><<<grid, kConvThreads, 0, st>>>(img_dev, n, d->rows, d->cols, d->H1, d->W1, d->conv1_hi, d->conv1_lo,
                                                              d->conv1_bq, d->act[cur], (int)M);
    else if (d->chnls == 1)
      conv1_tc_kernel<1, false><<<grid, kConvThreads, 0, st>>>(img_dev, n, d->rows, d->cols, d->H1, d->W1, d->conv1_hi, d->conv1_lo,
                                                               d->conv1_b, d->act[cur], (int)M);
    else if (d->stem_q)
      conv1_tc_kernel<3, true><<<grid, kConvThreads, 0, st>>>(img_dev, n, d->rows, d->cols, d->H1, d->W1, d->conv1_hi, d->conv1_lo,
                                                              d->conv1_bq, d->act[cur], (int)M);
    else
      conv1_tc_kernel<3, false><<<grid, kConvThreads, 0, st>>>(img_dev, n, d->rows, d->cols, d->H1, d->W1, d->conv1_hi, d->conv1_lo,
                                                               d->conv1_b, d->act[cur], (int)M);
    CB_LAUNCH_CHECK();
  } else {
    const long long threads = (long long)n * d->H1 * ((d->W1 + kPX - 1) / kPX) * 4;
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (d->chnls == 1)
      conv1_kernel<1><<<grid, 256, 0, st>>>(img_dev, n, d->rows, d->cols, d->H1, d->W1, d->conv1_w, d->conv1_b, d->act[cur]);
    else
      conv1_kernel<3><<<grid, 256, 0, st>>>(img_dev, n, d->rows, d->cols, d->H1, d->W1, d->conv1_w, d->conv1_b, d->act[cur]);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

int run_vlad_head(cb_descriptor* d, int n, int cur, float* out_dev, cudaStream_t st);
int forward_v2(cb_descriptor* d, int n, const uint8_t* img_dev, float* out_dev, cudaStream_t st);

int forward(cb_descriptor* d, int n, const uint8_t* img_dev, float* out_dev, cudaStream_t st) {
  if (d->v2) return forward_v2(d, n, img_dev, out_dev, st);
  int cur = 0;
  {
    int rc = run_stem(d, n, img_dev, cur, st);
    if (rc) return rc;
  }
  int layer = 0;
  d->last_buf = cur;
  if (d->stop_layer == layer) return CB_OK;
  for (const Block& b : d->blocks) {
    const int dw_layer = layer + 1, pw_layer = layer + 2;
    if (b.has_pw && b.use_tc && b.fused && !d->no_fuse && !d->force_simt && d->stop_layer != dw_layer) {
      // fused depthwise -> pointwise: one kernel, one ping-pong flip (the depthwise output never exists)
      int rc = b.runs_halo ? run_halo(b, n, d->sm_count, cur, st)
                           : run_fused(b, n, d->sm_count, d->act[cur], d->act[cur ^ 1], st);
      if (rc) return rc;
      cur ^= 1;
      d->last_buf = cur;
      layer = pw_layer;
      if (d->stop_layer == layer) return CB_OK;
      continue;
    }
    const long long threads = (long long)n * b.Ho * ((b.Wo + kPX - 1) / kPX) * (b.C / 8);
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (b.stride == 1)
      dw_kernel<1><<<grid, 256, 0, st>>>(d->act[cur], n, b.Hin, b.Win, b.C, b.Ho, b.Wo, b.dw_w, b.dw_b, d->act[cur ^ 1]);
    else
      dw_kernel<2><<<grid, 256, 0, st>>>(d->act[cur], n, b.Hin, b.Win, b.C, b.Ho, b.Wo, b.dw_w, b.dw_b, d->act[cur ^ 1]);
    CB_LAUNCH_CHECK();
    cur ^= 1;
    d->last_buf = cur;
    if (d->stop_layer == ++layer) return CB_OK;
    if (b.has_pw) {
      const long long M = (long long)n * b.Ho * b.Wo;
      int rc = run_pw(d, b, M, cur, d->act[cur ^ 1], st);
      if (rc) return rc;
      cur ^= 1;
      d->last_buf = cur;
      if (d->stop_layer == ++layer) return CB_OK;
    }
  }
  return run_vlad_head(d, n, cur, out_dev, st);
}

int run_vlad_head(cb_descriptor* d, int n, int cur, float* out_dev, cudaStream_t st) {
  const int P = d->Hf * d->Wf;
  const long long Ptot = (long long)n * P;
  if (d->vlad_tc && !d->force_simt && !d->no_fuse) {
    constexpr int kSm = kStages * (128 * 64 * 2 + 2 * 16 * 64 * 2) + 1024 + 64;
    CB_CUDA(cudaFuncSetAttribute(vlad_assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSm));
    vlad_assign_tc_kernel<<<(unsigned)((Ptot + 127) / 128), kGemmThreads, kSm, st>>>(d->tmX[cur], d->tmWhi, d->tmWlo, d->vlad_b,
                                                                                  d->assign, (int)Ptot, d->D);
    CB_LAUNCH_CHECK();
  } else {
    vlad_assign_kernel<<<(unsigned)((Ptot + 7) / 8), 256, (size_t)d->D * kK * sizeof(float), st>>>(
        d->act[cur], Ptot, d->D, d->vlad_w, d->vlad_b, d->assign);
    CB_LAUNCH_CHECK();
  }
  float* Ap = d->Vraw + (size_t)d->max_batch * kAggSplit * kK * d->D;  // [frames][split][16] after the partial V block
  CB_CUDA(cudaFuncSetAttribute(vlad_aggregate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAggSmem));
  vlad_aggregate_kernel<<<dim3(kAggSplit, n), 256, kAggSmem, st>>>(d->act[cur], d->assign, P, d->D, d->Vraw, Ap);
  CB_LAUNCH_CHECK();
  vlad_norm_kernel<<<n, 512, 0, st>>>(d->Vraw, Ap, d->vlad_c, d->D, out_dev, d->K);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

template <int N_TILE, int KB>
int launch_gemm_layer(const PwLayer& L, long long M, int in_buf, __half* out, const __half* residual, int relu, cudaStream_t st) {
  using SM = GemmSmem<N_TILE, KB>;
  auto kern = pw_gemm_kernel<N_TILE, KB>;
  CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal));
  dim3 grid((unsigned)((M + 127) / 128), (unsigned)(L.Cout / N_TILE));
  kern<<<grid, kGemmThreads, SM::kTotal, st>>>(L.tmA[in_buf], L.tmB, L.b, out, (int)M, L.Cout, L.C, residual, relu);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int run_pw_layer(cb_descriptor* d, const PwLayer& L, long long M, int in_buf, __half* out, const __half* residual, int relu,
                 cudaStream_t st) {
  if (L.use_tc && !d->force_simt) {
    if (L.kb == 64) {
      switch (L.n_tile) {
        case 64: return launch_gemm_layer<64, 64>(L, M, in_buf, out, residual, relu, st);
        case 128: return launch_gemm_layer<128, 64>(L, M, in_buf, out, residual, relu, st);
        default: return launch_gemm_layer<256, 64>(L, M, in_buf, out, residual, relu, st);
      }
    } else {
      switch (L.n_tile) {
        case 64: return launch_gemm_layer<64, 32>(L, M, in_buf, out, residual, relu, st);
        case 128: return launch_gemm_layer<128, 32>(L, M, in_buf, out, residual, relu, st);
        default: return launch_gemm_layer<256, 32>(L, M, in_buf, out, residual, relu, st);
      }
    }
  }
  const long long threads = M * (L.Cout / 8);
  pw_simt_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d->act[in_buf], L.w, L.b, out, M, L.Cout, L.C, residual, relu);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

// MobileNetV2 prefix: stem, then per block [expand GEMM] -> depthwise -> project GEMM (+ skip), rotating three buffers:
// the block input stays untouched until the project epilogue has added it.
int forward_v2(cb_descriptor* d, int n, const uint8_t* img_dev, float* out_dev, cudaStream_t st) {
  int cur = 0;
  int rc = run_stem(d, n, img_dev, cur, st);
  if (rc) return rc;
  int layer = 0;
  d->last_buf = cur;
  if (d->stop_layer == layer) return CB_OK;
  for (const IrBlock& b : d->ir) {
    int x = cur;
    if (b.has_expand) {
      x = (cur + 1) % 3;
      rc = run_pw_layer(d, b.expand, (long long)n * b.Hin * b.Win, cur, d->act[x], nullptr, 1, st);
      if (rc) return rc;
      d->last_buf = x;
      if (d->stop_layer == ++layer) return CB_OK;
    }
    const int dwb = b.has_expand ? (cur + 2) % 3 : (cur + 1) % 3;
    const long long threads = (long long)n * b.Ho * ((b.Wo + kPX - 1) / kPX) * (b.Cexp / 8);
    const unsigned grid = (unsigned)((threads + 255) / 256);
    if (b.stride == 1)
      dw_kernel<1><<<grid, 256, 0, st>>>(d->act[x], n, b.Hin, b.Win, b.Cexp, b.Ho, b.Wo, b.dw_w, b.dw_b, d->act[dwb]);
    else
      dw_kernel<2><<<grid, 256, 0, st>>>(d->act[x], n, b.Hin, b.Win, b.Cexp, b.Ho, b.Wo, b.dw_w, b.dw_b, d->act[dwb]);
    CB_LAUNCH_CHECK();
    d->last_buf = dwb;
    if (d->stop_layer == ++layer) return CB_OK;
    const int ob = b.has_expand ? x : (cur + 2) % 3;  // never the block input (the skip) nor the depthwise output
    rc = run_pw_layer(d, b.project, (long long)n * b.Ho * b.Wo, dwb, d->act[ob], b.residual ? d->act[cur] : nullptr, 0, st);
    if (rc) return rc;
    cur = ob;
    d->last_buf = cur;
    if (d->stop_layer == ++layer) return CB_OK;
  }
  return run_vlad_head(d, n, cur, out_dev, st);
}

inline int conv_out_s2(int h) { return (h + 1 - 3) / 2 + 1; }  // ZeroPadding2D((0,1),(0,1)) + 3x3 'valid' stride 2
inline int pad64(int c) { return (c + 63) / 64 * 64; }

// natural [c][cout] fp32 weights + [cout] bias -> zero-padded fp16 [Cout][C] + fp32 [Cout], tensor maps for the three buffers
int setup_pw_layer(cb_descriptor* d, PwLayer& L, const float* w, const float* b, int c, int cout, int C, int Cout, uint64_t Mmax) {
  L.C = C;
  L.Cout = Cout;
  std::vector<__half> tmp((size_t)C * Cout, __float2half_rn(0.f));
  for (int k = 0; k < c; ++k)
    for (int nn = 0; nn < cout; ++nn) tmp[(size_t)nn * C + k] = __float2half_rn(w[(size_t)k * cout + nn]);
  std::vector<float> bias((size_t)Cout, 0.f);
  for (int nn = 0; nn < cout; ++nn) bias[nn] = b[nn];
  cudaError_t e = cudaMalloc((void**)&L.w, tmp.size() * sizeof(__half));
  if (e == cudaSuccess) e = cudaMemcpy(L.w, tmp.data(), tmp.size() * sizeof(__half), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return cb::fail(CB_ENOMEM, "pointwise weight upload failed: %s", cudaGetErrorString(e));
  int rc = upload_f32(&L.b, bias.data(), bias.size());
  if (rc) return rc;
  L.kb = (C % 64 == 0) ? 64 : ((C % 32 == 0) ? 32 : 0);
  L.n_tile = (Cout % 256 == 0) ? 256 : ((Cout % 128 == 0) ? 128 : ((Cout % 64 == 0) ? 64 : 0));
  L.use_tc = L.kb != 0 && L.n_tile != 0;
  if (L.use_tc) {
    for (int i = 0; i < 3 && !rc; ++i) rc = make_map_2d(&L.tmA[i], d->act[i], Mmax, (uint64_t)C, 128, (uint32_t)L.kb);
    if (!rc) rc = make_map_2d(&L.tmB, L.w, (uint64_t)Cout, (uint64_t)C, (uint32_t)L.n_tile, (uint32_t)L.kb);
  }
  return rc;
}

// activation buffers (n_bufs of max_elems * max_batch halves), head scratch, host-API staging, streams
int alloc_common(cb_descriptor* d, size_t max_elems, int n_bufs) {
  const int max_batch = d->max_batch;
  d->act_elems = max_elems * (size_t)max_batch;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < n_bufs && e == cudaSuccess; ++i) e = cudaMalloc((void**)&d->act[i], d->act_elems * sizeof(__half));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d->assign, (size_t)max_batch * d->Hf * d->Wf * kK * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d->Vraw, (size_t)max_batch * kAggSplit * kK * (d->D + 1) * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d->img_dev, (size_t)max_batch * d->rows * d->cols * d->chnls);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d->out_dev, (size_t)max_batch * kK * d->D * sizeof(float));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) return cb::fail(CB_ENOMEM, "descriptor allocation failed: %s", cudaGetErrorString(e));
  return CB_OK;
}

// stem weights (fp32 for the CUDA-core kernel, hi/lo fp16 with the 2/255 scale folded in for the tcgen05 one) and the
// NetVLAD head (fp32 + hi/lo fp16 K-major soft-assignment weights, one feature-map tensor map per activation buffer)
int setup_stem_and_head(cb_descriptor* d, const float* conv1_w, const float* conv1_b, const float* vlad_w_in, const float* vlad_b_in,
                        const float* vlad_c_in, int k_total) {
  const int chnls = d->chnls, max_batch = d->max_batch;
  // the head kernels are built for 16 soft-assignment columns: fewer clusters (or clusters + ghosts) are padded with dead
  // ones -- zero weights and centres, bias -1e30, so their softmax weight is exactly 0
  std::vector<float> wpad((size_t)d->D * kK, 0.f), bpad(kK, -1e30f), cpad((size_t)d->D * kK, 0.f);
  for (int dd = 0; dd < d->D; ++dd)
    for (int k = 0; k < k_total; ++k) {
      wpad[(size_t)dd * kK + k] = vlad_w_in[(size_t)dd * k_total + k];
      cpad[(size_t)dd * kK + k] = vlad_c_in[(size_t)dd * k_total + k];
    }
  for (int k = 0; k < k_total; ++k) bpad[k] = vlad_b_in[k];
  const float *vlad_w = wpad.data(), *vlad_b = bpad.data(), *vlad_c = cpad.data();
  int rc = upload_f32(&d->conv1_w, conv1_w, (size_t)9 * chnls * 32);
  if (!rc) {
    std::vector<__half> hi(32 * 32, __float2half_rn(0.f)), lo(32 * 32, __float2half_rn(0.f));
    for (int k = 0; k < 9 * chnls; ++k)
      for (int nn = 0; nn < 32; ++nn) {
        const float v = conv1_w[(size_t)k * 32 + nn] * (2.0f / 255.0f);  // server.py:629 folded into the weights
        const __half h = __float2half_rn(v);
        hi[nn * 32 + k] = h;
        lo[nn * 32 + k] = __float2half_rn(v - __half2float(h));
      }
    cudaError_t e2 = cudaMalloc((void**)&d->conv1_hi, hi.size() * sizeof(__half));
    if (e2 == cudaSuccess) e2 = cudaMalloc((void**)&d->conv1_lo, lo.size() * sizeof(__half));
    if (e2 == cudaSuccess) e2 = cudaMemcpy(d->conv1_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e2 == cudaSuccess) e2 = cudaMemcpy(d->conv1_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e2 != cudaSuccess) rc = cb::fail(CB_ENOMEM, "stem weight upload failed: %s", cudaGetErrorString(e2));
  }
  if (!rc) rc = upload_f32(&d->conv1_b, conv1_b, 32);
  if (!rc) rc = upload_f32(&d->vlad_w, vlad_w, (size_t)d->D * kK);
  if (!rc) rc = upload_f32(&d->vlad_b, vlad_b, kK);
  if (!rc) rc = upload_f32(&d->vlad_c, vlad_c, (size_t)d->D * kK);
  if (!rc && d->D % 64 == 0) {
    std::vector<__half> hi((size_t)kK * d->D), lo((size_t)kK * d->D);
    for (int dd = 0; dd < d->D; ++dd)
      for (int k = 0; k < kK; ++k) {
        const float v = vlad_w[(size_t)dd * kK + k];
        const __half h = __float2half_rn(v);
        hi[(size_t)k * d->D + dd] = h;
        lo[(size_t)k * d->D + dd] = __float2half_rn(v - __half2float(h));
      }
    cudaError_t e3 = cudaMalloc((void**)&d->vlad_whi, hi.size() * sizeof(__half));
    if (e3 == cudaSuccess) e3 = cudaMalloc((void**)&d->vlad_wlo, lo.size() * sizeof(__half));
    if (e3 == cudaSuccess) e3 = cudaMemcpy(d->vlad_whi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e3 == cudaSuccess) e3 = cudaMemcpy(d->vlad_wlo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice);
    if (e3 != cudaSuccess) rc = cb::fail(CB_ENOMEM, "VLAD weight upload failed: %s", cudaGetErrorString(e3));
    const uint64_t Pmax = (uint64_t)max_batch * d->Hf * d->Wf;
    for (int i = 0; i < 3 && !rc; ++i)
      if (d->act[i]) rc = make_map_2d(&d->tmX[i], d->act[i], Pmax, (uint64_t)d->D, 128, 64);
    if (!rc) rc = make_map_2d(&d->tmWhi, d->vlad_whi, (uint64_t)kK, (uint64_t)d->D, kK, 64);
    if (!rc) rc = make_map_2d(&d->tmWlo, d->vlad_wlo, (uint64_t)kK, (uint64_t)d->D, kK, 64);
    d->vlad_tc = !rc;
  }
  return rc;
}

int upload_f32_padded(float** dst, const float* src, size_t rows, size_t c, size_t C) {  // [rows][c] -> [rows][C], zero fill
  std::vector<float> tmp(rows * C, 0.f);
  for (size_t r = 0; r < rows; ++r)
    for (size_t k = 0; k < c; ++k) tmp[r * C + k] = src[r * c + k];
  return upload_f32(dst, tmp.data(), tmp.size());
}

}  // namespace

extern "C" {

int cb_descriptor_create(cb_descriptor** out, const cb_netvlad_weights* w, int rows, int cols, int chnls, int max_batch,
                         int device) {
  if (!out) return cb::fail(CB_EINVAL, "out is NULL");
  *out = nullptr;
  if (!w) return cb::fail(CB_EINVAL, "weights is NULL");
  if (chnls != w->in_channels || (chnls != 1 && chnls != 3))
    return cb::fail(CB_EINVAL, "image channels %d do not match the model's %d (must be 1 or 3)", chnls, w->in_channels);
  if (rows < 32 || cols < 32 || max_batch < 1) return cb::fail(CB_EINVAL, "bad image size / batch");
  if (w->vlad_k < 1 || w->vlad_k > kK || w->vlad_ghost < 0 || w->vlad_ghost >= w->vlad_k)
    return cb::fail(CB_EINVAL, "NetVLAD heads are built for up to 16 soft-assignment clusters, ghosts included (got K=%d, ghosts=%d)",
                    w->vlad_k, w->vlad_ghost);
  int sm = 0;
  int rc = cb::select_device(device, &sm);
  if (rc) return rc;
  cb::DeviceGuard g(device);
  cb_descriptor* d = new cb_descriptor();
  d->device = device;
  d->sm_count = sm;
  d->rows = rows;
  d->cols = cols;
  d->chnls = chnls;
  d->max_batch = max_batch;
  const char* env = getenv("CB_PW_SIMT");
  d->force_simt = env && env[0] == '1';
  const char* env3 = getenv("CB_NO_FUSE");
  d->no_fuse = env3 && env3[0] == '1';
  const char* env4 = getenv("CB_NO_HALO");
  d->no_halo = env4 && env4[0] == '1';
  const char* env5 = getenv("CB_DESC_FP16");
  d->fp16_legacy = env5 && env5[0] == '1';
  if (const char* env6 = getenv("CB_DESC_SPLIT")) {
    const int nsp = atoi(env6);
    if (nsp >= 0) d->split_blocks = nsp;
  }
  const char* env2 = getenv("CB_DEBUG_STOP_LAYER");
  d->stop_layer = env2 ? atoi(env2) : -1;
  if (const char* envc = getenv("CB_DESC_CHUNK")) {
    const int c = atoi(envc);
    if (c >= 1) d->upload_chunk = c;
  }
  d->H1 = conv_out_s2(rows);
  d->W1 = conv_out_s2(cols);
  size_t max_elems = (size_t)d->H1 * d->W1 * 32;
  d->layer_elems.push_back(max_elems);
  int h = d->H1, wd = d->W1, c = 32;
  for (int i = 0; i < w->n_blocks; ++i) {
    Block b;
    b.C = c;
    b.stride = w->dw_stride[i];
    b.Hin = h;
    b.Win = wd;
    b.Ho = b.stride == 2 ? conv_out_s2(h) : h;
    b.Wo = b.stride == 2 ? conv_out_s2(wd) : wd;
    b.has_pw = w->pw_w[i] != nullptr;
    b.Cout = b.has_pw ? w->channels_out[i] : c;
    if (c % 8 || b.Cout % 8) {
      cb_descriptor_destroy(d);
      return cb::fail(CB_EINVAL, "channel counts must be multiples of 8");
    }
    h = b.Ho;
    wd = b.Wo;
    const size_t e_dw = (size_t)h * wd * c;
    d->layer_elems.push_back(e_dw);
    if (e_dw > max_elems) max_elems = e_dw;
    if (b.has_pw) {
      const size_t e_pw = (size_t)h * wd * b.Cout;
      d->layer_elems.push_back(e_pw);
      if (e_pw > max_elems) max_elems = e_pw;
    }
    c = b.Cout;
    d->blocks.push_back(b);
  }
  d->Hf = h;
  d->Wf = wd;
  d->D = c;
  d->K = w->vlad_k - w->vlad_ghost;  // clusters that reach the output
  if (w->vlad_d != c || c % 64 || c > 512) {
    cb_descriptor_destroy(d);
    return cb::fail(CB_EINVAL, "NetVLAD input dim %d does not match backbone output %d (or not a multiple of 64)", w->vlad_d, c);
  }
  rc = alloc_common(d, max_elems, 2);
  if (!rc) rc = setup_stem_and_head(d, w->conv1_w, w->conv1_b, w->vlad_w, w->vlad_b, w->vlad_c, w->vlad_k);
  cudaError_t e = cudaSuccess;
  for (size_t i = 0; i < d->blocks.size() && !rc; ++i) {
    Block& b = d->blocks[i];
    rc = upload_f32(&b.dw_w, w->dw_w[i], (size_t)9 * b.C);
    if (!rc) rc = upload_f32(&b.dw_b, w->dw_b[i], b.C);
    if (b.has_pw && !rc) {
      // host: [C][Cout] fp32 -> device [Cout][C] fp16 (K-major B operand)
      std::vector<__half> tmp((size_t)b.C * b.Cout);
      for (int k = 0; k < b.C; ++k)
        for (int nn = 0; nn < b.Cout; ++nn) tmp[(size_t)nn * b.C + k] = __float2half_rn(w->pw_w[i][(size_t)k * b.Cout + nn]);
      e = cudaMalloc((void**)&b.pw_w, tmp.size() * sizeof(__half));
      if (e == cudaSuccess) e = cudaMemcpy(b.pw_w, tmp.data(), tmp.size() * sizeof(__half), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) rc = cb::fail(CB_ENOMEM, "pointwise weight upload failed: %s", cudaGetErrorString(e));
      if (!rc) rc = upload_f32(&b.pw_b, w->pw_b[i], b.Cout);
      // tcgen05 tiling: K block 64 (128B swizzle) when Cin % 64 == 0, else 32 (64B swizzle)
      b.kb = (b.C % 64 == 0) ? 64 : ((b.C % 32 == 0) ? 32 : 0);
      b.n_tile = (b.Cout % 256 == 0) ? 256 : ((b.Cout % 128 == 0) ? 128 : ((b.Cout % 64 == 0) ? 64 : 0));
      b.use_tc = b.kb != 0 && b.n_tile != 0;
      if (b.use_tc && !rc) {
        const uint64_t Mmax = (uint64_t)max_batch * b.Ho * b.Wo;
        rc = make_map_2d(&b.tmA[0], d->act[0], Mmax, (uint64_t)b.C, 128, (uint32_t)b.kb);
        if (!rc) rc = make_map_2d(&b.tmA[1], d->act[1], Mmax, (uint64_t)b.C, 128, (uint32_t)b.kb);
        if (!rc) rc = make_map_2d(&b.tmB, b.pw_w, (uint64_t)b.Cout, (uint64_t)b.C, (uint32_t)b.n_tile, (uint32_t)b.kb);
        b.fused = fused_shape_supported(b.C, b.Cout, b.stride);
        if (b.fused && !rc)
          rc = make_map_2d(&b.tmBf, b.pw_w, (uint64_t)b.Cout, (uint64_t)b.C, (uint32_t)(b.Cout > 256 ? 256 : b.Cout),
                           (uint32_t)(b.C < 64 ? b.C : 64));
        if (b.fused && !rc) {
          const uint32_t kbh = 32;  // HaloCfg::KB
          const uint32_t hh = (kTH - 1) * b.stride + 3, hw = (kTW - 1) * b.stride + 3;
          int rh = make_map_2d(&b.tmBh, b.pw_w, (uint64_t)b.Cout, (uint64_t)b.C, (uint32_t)(b.Cout > 256 ? 256 : b.Cout), kbh);
          for (int i2 = 0; i2 < 2 && !rh; ++i2)
            rh = make_map_halo(&b.tmH[i2], d->act[i2], (uint64_t)b.C, (uint64_t)b.Win, (uint64_t)b.Hin, (uint64_t)max_batch, kbh,
                               hw, hh);
          for (int i2 = 0; i2 < 2 && !rh; ++i2)
            rh = make_map_halo(&b.tmO[i2], d->act[i2], (uint64_t)b.Cout, (uint64_t)b.Wo, (uint64_t)b.Ho, (uint64_t)max_batch, 32, kTW,
                               4, 64);
          b.halo = rh == CB_OK;  // a frame too small for the box keeps the global-memory producers
        }
        if (b.fused && b.halo && !rc) {
          // precise path: q15 decode folded into the depthwise weights (f = 2 + u 2^-14 -> x = (f - 2) * kQ15DecodeW),
          // q15-scaled pointwise bias, hi / lo split of the pointwise weights
          std::vector<float> wq((size_t)9 * b.C), bq((size_t)b.C), pbq((size_t)b.Cout);
          for (int ch = 0; ch < b.C; ++ch) {
            double sum = 0.0;
            for (int t9 = 0; t9 < 9; ++t9) {
              const float v = (float)((double)w->dw_w[i][(size_t)t9 * b.C + ch] * kQ15DecodeW);
              wq[(size_t)t9 * b.C + ch] = v;
              sum += (double)v;
            }
            bq[ch] = (float)((double)w->dw_b[i][ch] - 2.0 * sum);
          }
          for (int nn = 0; nn < b.Cout; ++nn) pbq[nn] = (float)((double)w->pw_b[i][nn] * (b.C <= kSatPackMaxCin ? 1.0 / 6.0 : 32767.0 / 6.0));
          rc = upload_f32(&b.dw_wq, wq.data(), wq.size());
          if (!rc) rc = upload_f32(&b.dw_bq, bq.data(), bq.size());
          if (!rc) rc = upload_f32(&b.pw_bq, pbq.data(), pbq.size());
          if (!rc && split_shape_supported(b.Cout)) {
            std::vector<__half> hl((size_t)2 * b.C * b.Cout);
            for (int k = 0; k < b.C; ++k)
              for (int nn = 0; nn < b.Cout; ++nn) {
                const float v = w->pw_w[i][(size_t)k * b.Cout + nn];
                const __half h = __float2half_rn(v);
                hl[(size_t)nn * b.C + k] = h;
                hl[(size_t)(b.Cout + nn) * b.C + k] = __float2half_rn(v - __half2float(h));
              }
            e = cudaMalloc((void**)&b.pw_whl, hl.size() * sizeof(__half));
            if (e == cudaSuccess) e = cudaMemcpy(b.pw_whl, hl.data(), hl.size() * sizeof(__half), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) rc = cb::fail(CB_ENOMEM, "pointwise hi/lo weight upload failed: %s", cudaGetErrorString(e));
            if (!rc) rc = make_map_2d(&b.tmBs, b.pw_whl, (uint64_t)2 * b.Cout, (uint64_t)b.C, (uint32_t)b.Cout, 32);
          }
        }
      }
    }
  }
  if (!rc) {
    // which blocks run on the halo kernel, and from that the storage format of every tensor between them
    const bool precise = !d->fp16_legacy && !d->no_fuse && !d->force_simt && !d->no_halo;
    int layer = 0;
    for (Block& b : d->blocks) {
      const int dw_layer = layer + 1;
      b.runs_halo = b.has_pw && b.use_tc && b.fused && b.halo && !d->no_fuse && !d->force_simt && !d->no_halo && d->stop_layer != dw_layer;
      layer += b.has_pw ? 2 : 1;
    }
    d->layer_q.assign(d->layer_elems.size(), 0);
    bool prev_q = precise;
    layer = 0;
    for (size_t i = 0; i < d->blocks.size(); ++i) {
      Block& b = d->blocks[i];
      b.mode = 0;
      if (b.runs_halo && prev_q) {
        b.mode = kInQ;
        if ((int)i < d->split_blocks) b.mode |= (split_shape_supported(b.Cout) && b.pw_whl) ? kSplit : kSplitA;
        if (i + 1 < d->blocks.size() && d->blocks[i + 1].runs_halo) b.mode |= kOutQ;
        d->layer_q[layer] = 1;  // this block's input
      } else {
        prev_q = false;
      }
      layer += b.has_pw ? 2 : 1;
    }
    // measured per shape (tools/bench_desc.py --layers, 64 frames of 480 x 640): 128->128 207 us with two epilogue groups vs 228
    // with one, 512->512 102 vs 116, but 256->256 137 with ONE group vs 176 with two
    for (Block& b : d->blocks) b.epi_warps = (b.C == 256 && b.Cout == 256) ? 4 : 8;
    if (const char* ee = getenv("CB_DESC_EPI"))
      for (Block& b : d->blocks) b.epi_warps = atoi(ee) == 8 ? 8 : 4;
    d->stem_q = !d->blocks.empty() && (d->blocks[0].mode & kInQ);
    if (d->stem_q) {
      std::vector<float> bq(32);
      for (int nn = 0; nn < 32; ++nn) bq[nn] = (float)((double)w->conv1_b[nn] * (32767.0 / 6.0));
      rc = upload_f32(&d->conv1_bq, bq.data(), 32);
    }
  }
  if (rc) {
    cb_descriptor_destroy(d);
    return rc;
  }
  *out = d;
  return CB_OK;
}

int cb_descriptor_create_v2(cb_descriptor** out, const cb_netvlad_v2_weights* w, int rows, int cols, int chnls, int max_batch,
                            int device) {
  if (!out) return cb::fail(CB_EINVAL, "out is NULL");
  *out = nullptr;
  if (!w || !w->blocks || w->n_blocks < 1) return cb::fail(CB_EINVAL, "weights is NULL / no blocks");
  if (chnls != w->in_channels || (chnls != 1 && chnls != 3))
    return cb::fail(CB_EINVAL, "image channels %d do not match the model's %d (must be 1 or 3)", chnls, w->in_channels);
  if (rows < 32 || cols < 32 || max_batch < 1) return cb::fail(CB_EINVAL, "bad image size / batch");
  if (w->vlad_k < 1 || w->vlad_k > kK || w->vlad_ghost < 0 || w->vlad_ghost >= w->vlad_k)
    return cb::fail(CB_EINVAL, "NetVLAD heads are built for up to 16 soft-assignment clusters, ghosts included (got K=%d, ghosts=%d)",
                    w->vlad_k, w->vlad_ghost);
  int sm = 0;
  int rc = cb::select_device(device, &sm);
  if (rc) return rc;
  cb::DeviceGuard g(device);
  cb_descriptor* d = new cb_descriptor();
  d->v2 = true;
  d->device = device;
  d->sm_count = sm;
  d->rows = rows;
  d->cols = cols;
  d->chnls = chnls;
  d->max_batch = max_batch;
  const char* env = getenv("CB_PW_SIMT");
  d->force_simt = env && env[0] == '1';
  const char* env3 = getenv("CB_NO_FUSE");
  d->no_fuse = env3 && env3[0] == '1';
  const char* env2 = getenv("CB_DEBUG_STOP_LAYER");
  d->stop_layer = env2 ? atoi(env2) : -1;
  if (const char* envc = getenv("CB_DESC_CHUNK")) {
    const int c = atoi(envc);
    if (c >= 1) d->upload_chunk = c;
  }
  d->H1 = conv_out_s2(rows);
  d->W1 = conv_out_s2(cols);
  size_t max_elems = (size_t)d->H1 * d->W1 * 32;
  d->layer_elems.push_back(max_elems);
  int h = d->H1, wd = d->W1, c = 32, c_nat = 32;  // stem output: 32 channels, unpadded
  for (int i = 0; i < w->n_blocks; ++i) {
    const cb_ir_block& src = w->blocks[i];
    const bool has_expand = src.expand_w != nullptr;
    const bool bad = src.c_in != c_nat || src.c_in % 8 || src.c_exp % 8 || src.c_out % 8 || (src.stride != 1 && src.stride != 2) ||
                     (!has_expand && src.c_exp != src.c_in) || (src.residual && (src.stride != 1 || src.c_in != src.c_out)) ||
                     !src.dw_w || !src.dw_b || !src.project_w || !src.project_b || (has_expand && !src.expand_b);
    if (bad) {
      cb_descriptor_destroy(d);
      return cb::fail(CB_EINVAL, "inverted-residual block %d is inconsistent (channels %d -> %d -> %d, stride %d, residual %d)", i,
                      src.c_in, src.c_exp, src.c_out, src.stride, src.residual);
    }
    IrBlock b;
    b.has_expand = has_expand;
    b.stride = src.stride;
    b.residual = src.residual;
    b.Cin = c;
    b.Cexp = has_expand ? pad64(src.c_exp) : c;
    b.Cout = pad64(src.c_out);
    b.Hin = h;
    b.Win = wd;
    b.Ho = b.stride == 2 ? conv_out_s2(h) : h;
    b.Wo = b.stride == 2 ? conv_out_s2(wd) : wd;
    if (has_expand) d->layer_elems.push_back((size_t)h * wd * b.Cexp);
    d->layer_elems.push_back((size_t)b.Ho * b.Wo * b.Cexp);
    d->layer_elems.push_back((size_t)b.Ho * b.Wo * b.Cout);
    for (int k = 0; k < 3; ++k) {
      const size_t e = d->layer_elems[d->layer_elems.size() - 1 - k];
      if (e > max_elems) max_elems = e;
    }
    h = b.Ho;
    wd = b.Wo;
    c = b.Cout;
    c_nat = src.c_out;
    d->ir.push_back(b);
  }
  d->Hf = h;
  d->Wf = wd;
  d->D = c;
  d->K = w->vlad_k - w->vlad_ghost;  // clusters that reach the output
  if (w->vlad_d != c_nat || c_nat % 64 || c > 512) {
    cb_descriptor_destroy(d);
    return cb::fail(CB_EINVAL, "NetVLAD input dim %d does not match backbone output %d (or not a multiple of 64)", w->vlad_d, c_nat);
  }
  rc = alloc_common(d, max_elems, 3);
  if (!rc) rc = setup_stem_and_head(d, w->conv1_w, w->conv1_b, w->vlad_w, w->vlad_b, w->vlad_c, w->vlad_k);
  c_nat = 32;
  for (int i = 0; i < w->n_blocks && !rc; ++i) {
    const cb_ir_block& src = w->blocks[i];
    IrBlock& b = d->ir[i];
    if (b.has_expand)
      rc = setup_pw_layer(d, b.expand, src.expand_w, src.expand_b, src.c_in, src.c_exp, b.Cin, b.Cexp,
                          (uint64_t)max_batch * b.Hin * b.Win);
    if (!rc) rc = upload_f32_padded(&b.dw_w, src.dw_w, 9, (size_t)src.c_exp, (size_t)b.Cexp);
    if (!rc) rc = upload_f32_padded(&b.dw_b, src.dw_b, 1, (size_t)src.c_exp, (size_t)b.Cexp);
    if (!rc)
      rc = setup_pw_layer(d, b.project, src.project_w, src.project_b, src.c_exp, src.c_out, b.Cexp, b.Cout,
                          (uint64_t)max_batch * b.Ho * b.Wo);
  }
  if (rc) {
    cb_descriptor_destroy(d);
    return rc;
  }
  *out = d;
  return CB_OK;
}

int cb_descriptor_destroy(cb_descriptor* d) {
  if (!d) return CB_OK;
  cb::DeviceGuard g(d->device);
  if (d->stream) cb::sync_stream(d->stream);
  for (IrBlock& b : d->ir) {
    void* ps[] = {b.expand.w, b.expand.b, b.project.w, b.project.b, b.dw_w, b.dw_b};
    for (void* q : ps)
      if (q) cudaFree(q);
  }
  for (Block& b : d->blocks) {
    cudaFree(b.dw_w);
    cudaFree(b.dw_b);
    cudaFree(b.pw_w);
    cudaFree(b.pw_b);
    void* qs[] = {b.dw_wq, b.dw_bq, b.pw_bq, b.pw_whl};
    for (void* q : qs)
      if (q) cudaFree(q);
  }
  if (d->conv1_bq) cudaFree(d->conv1_bq);
  void* ptrs[] = {d->vlad_whi, d->vlad_wlo, d->conv1_hi, d->conv1_lo, d->conv1_w, d->conv1_b, d->vlad_w, d->vlad_b, d->vlad_c, d->act[0], d->act[1],
                  d->act[2], d->assign,  d->Vraw,    d->img_dev, d->out_dev};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (int i = 0; i < 2; ++i) {
    if (d->stage_host[i]) cudaFreeHost(d->stage_host[i]);
    if (d->stage_free[i]) cudaEventDestroy(d->stage_free[i]);
  }
  if (d->out_host) cudaFreeHost(d->out_host);
  if (d->stream) cudaStreamDestroy(d->stream);
  if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
  for (cudaEvent_t ev : d->ev_copy) cudaEventDestroy(ev);
  delete d;
  return CB_OK;
}

int cb_descriptor_dim(const cb_descriptor* d) { return d ? d->K * d->D : -1; }

int cb_descriptor_compute_device(cb_descriptor* d, int n, const uint8_t* images_dev, float* out_dev, void* stream) {
  if (!d || !images_dev || !out_dev) return cb::fail(CB_EINVAL, "NULL argument to cb_descriptor_compute_device");
  if (n < 1 || n > d->max_batch) return cb::fail(CB_EINVAL, "batch %d outside [1,%d]", n, d->max_batch);
  cb::DeviceGuard g(d->device);
  return forward(d, n, images_dev, out_dev, (cudaStream_t)stream);
}

namespace {

bool host_pointer_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// uploads + forward passes of `n` frames; descriptors land in d->out_dev (no read-back, no synchronisation)
int compute_to_device(cb_descriptor* d, int n, const uint8_t* images, int64_t row_stride_bytes) {
  const size_t rowb = (size_t)d->cols * d->chnls;
  if (row_stride_bytes == 0) row_stride_bytes = (int64_t)rowb;
  if ((size_t)row_stride_bytes < rowb) return cb::fail(CB_EINVAL, "row stride smaller than a row");
  // server.py:614-619 asserts the image shape; here the shape is fixed at create time
  const size_t frame_bytes = rowb * d->rows;
  const bool pinned = host_pointer_is_pinned(images);
  // Pinned caller memory: frames go up in chunks straight from the caller's buffer on a copy stream while the previous
  // chunk is being processed (the 0.9 MB/frame upload is the largest cost of the host path).  Pageable memory (what a ROS
  // callback holds): the same pipeline through the library's own pinned ring, so the DMA never waits for the driver's
  // internal staging and the CPU copy of the next chunk overlaps the device work of the current one.
  const int chunk = pinned ? d->upload_chunk : d->stage_frames;
  if (!pinned && !d->stage_host[0]) {
    for (int i = 0; i < 2; ++i) {
      CB_CUDA(cudaHostAlloc((void**)&d->stage_host[i], (size_t)d->stage_frames * frame_bytes, cudaHostAllocDefault));
      CB_CUDA(cudaEventCreateWithFlags(&d->stage_free[i], cudaEventDisableTiming));
    }
  }
  int ci = 0;
  for (int c0 = 0; c0 < n; c0 += chunk, ++ci) {
    const int nc = n - c0 < chunk ? n - c0 : chunk;
    if ((int)d->ev_copy.size() <= ci) {
      cudaEvent_t ev;
      CB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      d->ev_copy.push_back(ev);
    }
    const uint8_t* src = images + (size_t)c0 * d->rows * (size_t)row_stride_bytes;
    uint8_t* dst = d->img_dev + (size_t)c0 * frame_bytes;
    if (pinned) {
      CB_CUDA(cudaMemcpy2DAsync(dst, rowb, src, (size_t)row_stride_bytes, rowb, (size_t)nc * d->rows, cudaMemcpyHostToDevice, d->copy_stream));
    } else {
      const int sb = ci & 1;
      if (ci >= 2) CB_CUDA(cudaEventSynchronize(d->stage_free[sb]));  // the DMA that last read this buffer has finished
      uint8_t* stg = d->stage_host[sb];
      if ((size_t)row_stride_bytes == rowb) {
        memcpy(stg, src, (size_t)nc * frame_bytes);
      } else {
        for (size_t r = 0; r < (size_t)nc * d->rows; ++r) memcpy(stg + r * rowb, src + r * (size_t)row_stride_bytes, rowb);
      }
      CB_CUDA(cudaMemcpyAsync(dst, stg, (size_t)nc * frame_bytes, cudaMemcpyHostToDevice, d->copy_stream));
      CB_CUDA(cudaEventRecord(d->stage_free[sb], d->copy_stream));
    }
    CB_CUDA(cudaEventRecord(d->ev_copy[ci], d->copy_stream));
    CB_CUDA(cudaStreamWaitEvent(d->stream, d->ev_copy[ci], 0));
    int rc = forward(d, nc, dst, d->out_dev + (size_t)c0 * d->K * d->D, d->stream);
    if (rc) return rc;
  }
  return CB_OK;
}

}  // namespace

int cb_descriptor_compute(cb_descriptor* d, int n, const uint8_t* images, int64_t row_stride_bytes, float* out) {
  if (!d || !images || !out) return cb::fail(CB_EINVAL, "NULL argument to cb_descriptor_compute");
  if (n < 1 || n > d->max_batch) return cb::fail(CB_EINVAL, "batch %d outside [1,%d]", n, d->max_batch);
  cb::DeviceGuard g(d->device);
  int rc = compute_to_device(d, n, images, row_stride_bytes);
  if (rc) return rc;
  CB_CUDA(cudaMemcpyAsync(out, d->out_dev, (size_t)n * d->K * d->D * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  CB_CUDA(cb::sync_stream(d->stream));
  return CB_OK;
}

int cb_descriptor_compute_f64(cb_descriptor* d, int n, const uint8_t* images, int64_t row_stride_bytes, double* out) {
  if (!d || !images || !out) return cb::fail(CB_EINVAL, "NULL argument to cb_descriptor_compute_f64");
  if (n < 1 || n > d->max_batch) return cb::fail(CB_EINVAL, "batch %d outside [1,%d]", n, d->max_batch);
  cb::DeviceGuard g(d->device);
  const size_t ne = (size_t)n * d->K * d->D;
  if (!d->out_host) CB_CUDA(cudaHostAlloc((void**)&d->out_host, (size_t)d->max_batch * d->K * d->D * sizeof(float), cudaHostAllocDefault));
  int rc = compute_to_device(d, n, images, row_stride_bytes);
  if (rc) return rc;
  // fp32 crosses the bus (half the bytes of the service's float64[]), the widening of server.py:648 / Cerebro.cpp:268-271
  // happens while writing the caller's buffer
  CB_CUDA(cudaMemcpyAsync(d->out_host, d->out_dev, ne * sizeof(float), cudaMemcpyDeviceToHost, d->stream));
  CB_CUDA(cb::sync_stream(d->stream));
  for (size_t i = 0; i < ne; ++i) out[i] = (double)d->out_host[i];
  return CB_OK;
}

int64_t cb_descriptor_get_activation(cb_descriptor* d, int layer, float* out, int64_t max_floats) {
  if (!d || !out) return cb::fail(CB_EINVAL, "NULL argument to cb_descriptor_get_activation");
  if (layer < 0 || layer >= (int)d->layer_elems.size()) return cb::fail(CB_EINVAL, "layer %d out of range", layer);
  // returns the MOST RECENT layer output (the forward pass stops after `CB_DEBUG_STOP_LAYER`, or runs to the
  // final feature map); `layer` only selects the element count
  const size_t ne = d->layer_elems[layer];
  if ((int64_t)ne > max_floats) return cb::fail(CB_EINVAL, "buffer too small: need %zu floats", ne);
  cb::DeviceGuard g(d->device);
  std::vector<__half> tmp(ne);
  CB_CUDA(cb::sync_stream(d->stream));
  CB_CUDA(cudaMemcpy(tmp.data(), d->act[d->last_buf], ne * sizeof(__half), cudaMemcpyDeviceToHost));
  const bool q = (size_t)layer < d->layer_q.size() && d->layer_q[layer];  // q15 storage between halo blocks
  if (q) {
    const uint16_t* u = reinterpret_cast<const uint16_t*>(tmp.data());
    for (size_t i = 0; i < ne; ++i) out[i] = (float)((double)u[i] * (6.0 / 32767.0));
  } else {
    for (size_t i = 0; i < ne; ++i) out[i] = __half2float(tmp[i]);
  }
  return (int64_t)ne;
}

}  // extern "C"
