// Correspondence front end of the geometric verifier (sm_100a): ORB descriptors + keypoints + depth images of a loop
// candidate's two frames  ->  brute-force Hamming matches  ->  GMS inlier mask  ->  3D-2D / 3D-3D sets for the pose solvers.
//
// Replaces (reference), in StaticPointFeatureMatching (src/utils/PointFeatureMatching.cpp):
//   cv::BFMatcher(cv::NORM_HAMMING).match(d1, d2)                              :40-43
//   gms_matcher(kp1, size1, kp2, size2, matches).GetInlierMask(mask, false, false)  :50-52
//        (src/utils/GMSMatcher/gms_matcher.h:51-243, gms_matcher.cpp:5-181)
//   make_3d_2d_collection__using__pfmatches_and_disparity                       :96-153
//   make_3d_3d_collection__using__pfmatches_and_disparity                       :158-196
// ORB detection and the stereo block matcher stay OpenCV's on the host.
//
// Kernels
//   hamming_tc_kernel    : the all-pairs Hamming distance as a binary GEMM on tcgen05 (kind::i8), see below; default
//   hamming_match_kernel : CUDA-core version (CB_MATCH_SIMT=1): thread = one query descriptor (8 x u32 in registers); the CTA stages 256 train descriptors at a
//                          time in shared memory and every thread scans them with broadcast 128-bit loads, xor + popc.
//                          The train set is split over gridDim.y; partial minima meet in a packed 64-bit atomicMin
//                          (distance << 32 | train index), which also reproduces OpenCV's first-minimum tie rule.
//   gms_kernel           : one CTA per image pair.  The 400 x 400 motion-statistics matrix of the reference is never
//                          formed: matches are bucketed by left grid cell with a shared-memory counting sort, a warp per
//                          cell finds the most frequent right cell (first maximum) and the 3 x 3 neighbourhood support
//                          by scanning the (short) buckets; four shifted grids, integer counts, fp64 threshold
//                          6 * sqrt(mean points per cell) exactly as gms_matcher.cpp:139-146.
//   collect_kernel       : one CTA per pair: stable compaction of the GMS inliers (match order), K^-1 normalisation,
//                          (int)-truncated depth-image lookup, 0.1 m <= z <= 25 m gate.
#include "common.cuh"
#include "ptx.cuh"
#include "stereo_core.h"

#include <stdlib.h>

namespace {

using cb::FULL;

constexpr int kDescWords = 8;       // 256-bit ORB descriptor
constexpr int kMatchThreads = 128;  // queries per CTA
constexpr int kTrainTile = 256;     // train descriptors staged per step
constexpr int kTrainChunk = 1024;   // train descriptors per gridDim.y slice
constexpr int kGrid = 20;           // gms_matcher.h:63
constexpr int kCells = kGrid * kGrid;
constexpr int kGmsThreads = 256;
constexpr int kGmsWarps = kGmsThreads / 32;

__global__ void __launch_bounds__(kMatchThreads)
hamming_match_kernel(const uint32_t* __restrict__ d1, const uint32_t* __restrict__ d2, const int* __restrict__ off1,
                     const int* __restrict__ off2, unsigned long long* __restrict__ best) {
  __shared__ uint4 tile[kTrainTile * 2];
  const int pair = blockIdx.z;
  const int q0 = off1[pair], n1 = off1[pair + 1] - q0;
  const int t0 = off2[pair], n2 = off2[pair + 1] - t0;
  const int qi = blockIdx.x * kMatchThreads + threadIdx.x;
  const int c0 = blockIdx.y * kTrainChunk;
  if (blockIdx.x * kMatchThreads >= n1 || c0 >= n2) return;  // CTA-uniform
  const int c1 = min(c0 + kTrainChunk, n2);
  uint32_t q[kDescWords];
  const bool active = qi < n1;
  {
    const uint4* qp = reinterpret_cast<const uint4*>(d1 + (size_t)(q0 + (active ? qi : 0)) * kDescWords);
    const uint4 a = qp[0], b = qp[1];
    q[0] = a.x, q[1] = a.y, q[2] = a.z, q[3] = a.w, q[4] = b.x, q[5] = b.y, q[6] = b.z, q[7] = b.w;
  }
  unsigned best_d = 0xffffffffu, best_i = 0;
  for (int base = c0; base < c1; base += kTrainTile) {
    const int nt = min(kTrainTile, c1 - base);
    __syncthreads();
    for (int i = threadIdx.x; i < nt * 2; i += kMatchThreads)
      tile[i] = reinterpret_cast<const uint4*>(d2 + (size_t)(t0 + base) * kDescWords)[i];
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < nt; ++j) {
      const uint4 a = tile[2 * j], b = tile[2 * j + 1];  // broadcast loads
      const unsigned d = __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) +
                         __popc(q[4] ^ b.x) + __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
      if (d < best_d) {  // strict: the first minimum wins (OpenCV batchDistance)
        best_d = d;
        best_i = (unsigned)(base + j);
      }
    }
  }
  if (active && best_d != 0xffffffffu)
    atomicMin(best + q0 + qi, ((unsigned long long)best_d << 32) | (unsigned long long)best_i);
}

// ---------------------------------------------------------------------------------------------
// Tensor-core matcher.  For bit vectors  hamming(a, b) = |a| + |b| - 2 a.b, and a.b over all query x train pairs is a
// GEMM: the descriptors are expanded once to one byte per bit (0 / 1), and  [128 queries x 256] x [256 x 256 train]
// tiles run as tcgen05.mma.kind::i8 (unsigned 8-bit operands, exact s32 accumulators in TMEM).  The popc pipe
// bounds the SIMT kernel (8 xor + 8 popc per pair of descriptors); here the per-pair work left for the CUDA cores is one
// integer multiply-add and one compare in the epilogue.
//   CTA = 192 threads: warp 0 TMA, warp 1 MMA, warps 2-5 epilogue (thread = one query row = one TMEM lane).
//   The 32 KB query tile stays resident; 64 KB train tiles (2 K blocks of 256 rows x 128 B, 128-byte swizzle) stream through
//   a 2-stage ring; two 256-column accumulators let the epilogue of tile i overlap the MMAs of tile i+1.
//   gridDim = (query tiles, train splits, pairs); partial minima meet in the same packed 64-bit atomicMin as the SIMT kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kHtThreads = 192;
constexpr int kHtM = 128, kHtN = 256;
constexpr int kHtABytes = 2 * kHtM * 128;  // both K blocks of the query tile
constexpr int kHtBBytes = 2 * kHtN * 128;  // both K blocks of one train tile
constexpr int kHtSmem = kHtABytes + 2 * kHtBBytes + 1024 + 256 + 2 * kHtN * 4;

// descriptor bits -> one byte per bit, |d| -> pop[]
__global__ void expand_bits_kernel(const uint32_t* __restrict__ d, int n, uint8_t* __restrict__ e, int* __restrict__ pop) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one 32-bit word -> 32 bytes
  const bool valid = i < n * kDescWords;                // no early exit: the warp shuffles below need every lane
  const uint32_t w = valid ? d[i] : 0u;
  uint32_t o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t nib = (w >> (4 * j)) & 0xfu;
    o[j] = (nib & 1u) | ((nib & 2u) << 7) | ((nib & 4u) << 14) | ((nib & 8u) << 21);
  }
  if (valid) {
    uint4* dst = reinterpret_cast<uint4*>(e + (size_t)i * 32);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
  }
  // eight consecutive threads hold one descriptor
  int c = __popc(w);
  c += __shfl_xor_sync(FULL, c, 1);
  c += __shfl_xor_sync(FULL, c, 2);
  c += __shfl_xor_sync(FULL, c, 4);
  if (valid && (i & 7) == 0) pop[i >> 3] = c;
}

__global__ void __launch_bounds__(kHtThreads, 1)
hamming_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmT, const int* __restrict__ off1,
                  const int* __restrict__ off2, const int* __restrict__ pop1, const int* __restrict__ pop2, int tiles_per_split,
                  unsigned long long* __restrict__ best) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kHtABytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + 2 * kHtBBytes);
  uint64_t* full = a_full + 1;         // [2]
  uint64_t* empty = full + 2;          // [2]
  uint64_t* tmem_full_bar = empty + 2;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  int* nb_s = reinterpret_cast<int*>(sB + 2 * kHtBBytes + 256);  // [2][kHtN] |train descriptor|, huge beyond the pair's set
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.z;
  const int q0 = off1[pair], n1 = off1[pair + 1] - q0;
  const int t0 = off2[pair], n2 = off2[pair + 1] - t0;
  const int m0 = blockIdx.x * kHtM;
  const int n_tiles_all = (n2 + kHtN - 1) / kHtN;
  const int tile_begin = blockIdx.y * tiles_per_split;
  const int tile_end = min(tile_begin + tiles_per_split, n_tiles_all);
  if (m0 >= n1 || tile_begin >= tile_end) return;  // CTA-uniform

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmT);
    mbar_init(a_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 4);
    }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(a_full, kHtABytes);
      tma_load_2d(&tmQ, a_full, sA, 0, q0 + m0);
      tma_load_2d(&tmQ, a_full, sA + kHtM * 128, 128, q0 + m0);
      uint32_t g = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++g) {
        const int st = g & 1;
        mbar_wait(&empty[st], ((g >> 1) & 1) ^ 1);
        mbar_expect_tx(&full[st], kHtBBytes);
        uint8_t* dst = sB + st * kHtBBytes;
        tma_load_2d(&tmT, &full[st], dst, 0, t0 + tile * kHtN);
        tma_load_2d(&tmT, &full[st], dst + kHtN * 128, 128, t0 + tile * kHtN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D = s32, A = B = unsigned 8-bit, both K-major, N = 256, M = 128
      const uint32_t idesc = (2u << 4) | ((uint32_t)(kHtN >> 3) << 17) | ((uint32_t)(kHtM >> 4) << 24);
      mbar_wait(a_full, 0);
      tc_fence_after();
      uint32_t g = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++g) {
        const int st = g & 1, acc = g & 1;
        mbar_wait(&tmem_empty_bar[acc], ((g >> 1) & 1) ^ 1);
        mbar_wait(&full[st], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB + st * kHtBBytes);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_i8(tmem_base + acc * kHtN, make_kmajor_desc<128>(a_addr + kb * kHtM * 128 + k * 32),
                    make_kmajor_desc<128>(b_addr + kb * kHtN * 128 + k * 32), idesc, (kb | k) ? 1u : 0u);
        umma_commit(&empty[st]);
        umma_commit(&tmem_full_bar[acc]);
      }
    }
  } else {
    const int qtr = warp & 3;
    const int et = threadIdx.x - 64;  // 0..127
    const int row = m0 + qtr * 32 + lane;
    int best_h = 0x7fffffff, best_i = 0;
    uint32_t g = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++g) {
      const int acc = g & 1;
      int* nb = nb_s + acc * kHtN;
      for (int c = et; c < kHtN; c += 128) {
        const int col = tile * kHtN + c;
        nb[c] = col < n2 ? pop2[t0 + col] : (1 << 24);  // columns beyond the pair's train set never win
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(&tmem_full_bar[acc], (g >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < kHtN; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + acc * kHtN + c + ((uint32_t)(qtr * 32) << 16), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int h = nb[c + j] - 2 * (int)v[j];  // + |query| at the end
          if (h < best_h) {  // strict and in ascending column order: the first minimum wins
            best_h = h;
            best_i = tile * kHtN + c + j;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
    if (row < n1 && best_h < (1 << 23)) {
      const unsigned long long key = ((unsigned long long)(unsigned)(best_h + pop1[q0 + row]) << 32) | (unsigned long long)(unsigned)best_i;
      atomicMin(best + q0 + row, key);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void unpack_matches_kernel(const unsigned long long* __restrict__ best, int n, int* __restrict__ train_idx,
                                      int* __restrict__ dist) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long k = best[i];
  const bool none = k == ~0ull;  // a pair without train descriptors
  train_idx[i] = none ? -1 : (int)(k & 0xffffffffu);
  dist[i] = none ? -1 : (int)(k >> 32);
}

// GetGridIndexLeft, gms_matcher.h:147-189.  px, py: normalised float32 coordinates; float * int is a float32 product, the
// + 0.5 is a double addition.
__device__ __forceinline__ int grid_index_left(float px, float py, int type) {
  const float fx = __fmul_rn(px, (float)kGrid), fy = __fmul_rn(py, (float)kGrid);
  int x, y;
  if (type == 1) {
    x = (int)floorf(fx), y = (int)floorf(fy);
    if (y >= kGrid || x >= kGrid) return -1;
  } else if (type == 2) {
    x = (int)floor((double)fx + 0.5), y = (int)floorf(fy);
    if (x >= kGrid || x < 1) return -1;
  } else if (type == 3) {
    x = (int)floorf(fx), y = (int)floor((double)fy + 0.5);
    if (y >= kGrid || y < 1) return -1;
  } else {
    x = (int)floor((double)fx + 0.5), y = (int)floor((double)fy + 0.5);
    if (y >= kGrid || y < 1 || x >= kGrid || x < 1) return -1;
  }
  return x + y * kGrid;
}

// neighbour k (0..8, row-major 3x3 around the cell) of `cell`, -1 outside the grid: GetNB9, gms_matcher.h:199-220
__device__ __forceinline__ int nb9(int cell, int k) {
  const int x = cell % kGrid + (k % 3 - 1), y = cell / kGrid + (k / 3 - 1);
  return (x < 0 || x >= kGrid || y < 0 || y >= kGrid) ? -1 : x + y * kGrid;
}

// grid (pairs, 4): one CTA per (pair, shifted grid) -- the four grid types of gms_matcher::run only meet in the OR of their
// inlier marks, so they run in parallel; `mask` is zeroed beforehand, every CTA stores 1s, and the last of a pair's four
// CTAs to finish (ticket counter) counts the inliers.
// Shared memory: rg[nm] u16 | sorted[nm] u16 | lgm[nm] i16 | count[400] | start[401] | cursor[400] | pairc[400]
__global__ void __launch_bounds__(kGmsThreads)
gms_kernel(const float* __restrict__ kp1, const float* __restrict__ kp2, const int* __restrict__ off1, const int* __restrict__ off2,
           const int* __restrict__ train_idx, int w1, int h1, int w2, int h2, int max_nm, unsigned char* __restrict__ mask,
           int* __restrict__ n_inliers, unsigned int* __restrict__ tickets) {
  extern __shared__ unsigned char gsm[];
  unsigned short* rg = reinterpret_cast<unsigned short*>(gsm);
  unsigned short* sorted = rg + max_nm;
  short* lgm = reinterpret_cast<short*>(sorted + max_nm);
  int* count = reinterpret_cast<int*>(lgm + max_nm + (max_nm & 1));
  int* start = count + kCells;
  int* cursor = start + kCells + 1;
  int* pairc = cursor + kCells;
  __shared__ int s_total;
  __shared__ unsigned int s_ticket;
  const int pair = blockIdx.x;
  const int q0 = off1[pair], nm = off1[pair + 1] - q0;  // one match per query descriptor
  const int t0 = off2[pair];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr unsigned short kBad = 0xffffu;

  // right grid cell of every match (GetGridIndexRight, gms_matcher.h:191-196; assigned once, gms_matcher.cpp:86-88);
  // scale 0: the right grid is 20 x 20 as well (SetScale(0), gms_matcher.cpp:11)
  for (int i = tid; i < nm; i += kGmsThreads) {
    const int t = train_idx[q0 + i];
    unsigned short r = kBad;
    if (t >= 0) {
      const float px = __fdiv_rn(kp2[2 * (size_t)(t0 + t)], (float)w2), py = __fdiv_rn(kp2[2 * (size_t)(t0 + t) + 1], (float)h2);
      const int x = (int)floorf(__fmul_rn(px, (float)kGrid)), y = (int)floorf(__fmul_rn(py, (float)kGrid));
      const int g = x + y * kGrid;
      // the reference indexes mMotionStatistics with g unchecked (out of bounds for a point on the right / bottom border);
      // such matches are left out here
      if (g >= 0 && g < kCells && x >= 0 && x < kGrid) r = (unsigned short)g;
    }
    rg[i] = r;
  }
  {
    const int type = (int)blockIdx.y + 1;
    for (int i = tid; i < kCells; i += kGmsThreads) count[i] = 0, cursor[i] = 0;
    __syncthreads();
    // AssignMatchPairs, gms_matcher.cpp:75-100
    for (int i = tid; i < nm; i += kGmsThreads) {
      const float px = __fdiv_rn(kp1[2 * (size_t)(q0 + i)], (float)w1), py = __fdiv_rn(kp1[2 * (size_t)(q0 + i) + 1], (float)h1);
      const int lg = grid_index_left(px, py, type);
      lgm[i] = (short)lg;
      if (lg >= 0 && rg[i] != kBad) atomicAdd(&count[lg], 1);
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan of 400 counts
      int carry = 0;
      for (int b = 0; b < kCells; b += 32) {
        const int v = (b + lane < kCells) ? count[b + lane] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int u = __shfl_up_sync(FULL, s, o);
          if (lane >= o) s += u;
        }
        if (b + lane < kCells) start[b + lane] = carry + s - v;
        carry += __shfl_sync(FULL, s, 31);
      }
      if (lane == 0) start[kCells] = carry;
    }
    __syncthreads();
    for (int i = tid; i < nm; i += kGmsThreads) {
      const int lg = lgm[i];
      if (lg >= 0 && rg[i] != kBad) sorted[start[lg] + atomicAdd(&cursor[lg], 1)] = rg[i];
    }
    __syncthreads();
    // VerifyCellPairs, gms_matcher.cpp:102-149: warp per left cell
    for (int cell = warp; cell < kCells; cell += kGmsWarps) {
      const int b0 = start[cell], m = count[cell];
      int result = -1;
      if (m > 0) {
        // most frequent right cell, smallest index among equals (strict > scanning upwards, :113-121)
        unsigned key = 0;
        for (int e = lane; e < m; e += 32) {
          const unsigned short v = sorted[b0 + e];
          unsigned c = 0;
          for (int f = 0; f < m; ++f) c += (sorted[b0 + f] == v);
          const unsigned k = (c << 16) | (0xffffu - v);
          key = max(key, k);
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) key = max(key, __shfl_xor_sync(FULL, key, o));
        const int j = (int)(0xffffu - (key & 0xffffu));
        // neighbourhood support (rotation pattern 1 = identity), :128-137
        int score = 0, tsum = 0, numpair = 0;
        for (int k = 0; k < 9; ++k) {
          const int ll = nb9(cell, k), rr = nb9(j, k);
          if (ll == -1 || rr == -1) continue;
          const int lb = start[ll], lm = count[ll];
          int c = 0;
          for (int e = lane; e < lm; e += 32) c += (sorted[lb + e] == (unsigned short)rr);
          score += c;
          tsum += lm;
          ++numpair;
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) score += __shfl_xor_sync(FULL, score, o);
        const double thresh = 6.0 * sqrt((double)tsum / (double)numpair);  // THRESH_FACTOR, :139
        result = ((double)score < thresh) ? -2 : j;
      }
      if (lane == 0) pairc[cell] = result;
    }
    __syncthreads();
    for (int i = tid; i < nm; i += kGmsThreads) {  // gms_matcher.cpp:169-177
      const int lg = lgm[i];
      if (lg >= 0 && rg[i] != kBad && pairc[lg] == (int)rg[i]) mask[q0 + i] = 1;
    }
  }
  // the last of the pair's CTAs counts the union of the four grids' marks
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    s_ticket = atomicAdd(&tickets[pair], 1u);
    s_total = 0;
  }
  __syncthreads();
  if (s_ticket != gridDim.y - 1) return;
  __threadfence();
  int c = 0;
  for (int i = tid; i < nm; i += kGmsThreads) c += __ldcg(mask + q0 + i);
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  if (lane == 0) atomicAdd(&s_total, c);
  __syncthreads();
  if (tid == 0) {
    n_inliers[pair] = s_total;
    tickets[pair] = 0;  // ready for the next launch
  }
}

// One CTA per pair.  Walks the matches in order, 256 at a time, and appends the kept ones (ballot + prefix) so the output
// order is the reference's (match order = ascending query index).
//   mode 0: 3D-2D  (world point of frame a at (int)uv, normalised uv of a and uv_d of b)   PointFeatureMatching.cpp:96-153
//   mode 1: 3D-3D  (points of both frames, both depths gated)                               PointFeatureMatching.cpp:158-196
//   mode 2: 3D-2D with the frames' roles swapped (Option B, Cerebro.cpp:1562-1565): world point of frame b at (int)uv_d
__global__ void __launch_bounds__(256)
collect_kernel(const float* __restrict__ kp1, const float* __restrict__ kp2, const int* __restrict__ off1, const int* __restrict__ off2,
               const int* __restrict__ train_idx, const unsigned char* __restrict__ mask, const float* __restrict__ img_a,
               const float* __restrict__ img_b, int H, int W, const double* __restrict__ Kinv, int mode, double* __restrict__ out_X,
               double* __restrict__ out_uv, double* __restrict__ out_uvd, double* __restrict__ out_Y, int* __restrict__ out_count) {
  __shared__ int s_warp[8];
  __shared__ int s_base;
  const int pair = blockIdx.x;
  const int q0 = off1[pair], nm = off1[pair + 1] - q0, t0 = off2[pair];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ia = img_a ? img_a + (size_t)pair * H * W * 3 : nullptr;
  const float* ib = img_b ? img_b + (size_t)pair * H * W * 3 : nullptr;
  const float* src3d = mode == 2 ? ib : ia;  // the frame whose 3-D image is looked up first
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < nm; i0 += 256) {
    const int i = i0 + tid;
    bool keep = false;
    float u = 0.f, v = 0.f, ud = 0.f, vd = 0.f;
    float3 pa = make_float3(0.f, 0.f, 0.f), pb = pa;
    if (i < nm && mask[q0 + i]) {
      const int t = train_idx[q0 + i];
      u = kp1[2 * (size_t)(q0 + i)], v = kp1[2 * (size_t)(q0 + i) + 1];
      ud = kp2[2 * (size_t)(t0 + t)], vd = kp2[2 * (size_t)(t0 + t) + 1];
      const int xa = (int)(mode == 2 ? ud : u), ya = (int)(mode == 2 ? vd : v);  // (int)uv(1,k), (int)uv(0,k): truncation, :124
      if (src3d && xa >= 0 && xa < W && ya >= 0 && ya < H) {
        const float* p = src3d + ((size_t)ya * W + xa) * 3;
        pa = make_float3(p[0], p[1], p[2]);
        keep = !(pa.z < 0.1f || pa.z > 25.f);
      }
      if (mode == 1 && keep) {
        const int xb = (int)ud, yb = (int)vd;
        keep = false;
        if (ib && xb >= 0 && xb < W && yb >= 0 && yb < H) {
          const float* p = ib + ((size_t)yb * W + xb) * 3;
          pb = make_float3(p[0], p[1], p[2]);
          keep = !(pb.z < 0.1f || pb.z > 25.f);
        }
      }
    }
    const unsigned bal = __ballot_sync(FULL, keep);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int w = 0; w < warp; ++w) before += s_warp[w];
    const int pos = before + __popc(bal & ((1u << lane) - 1u));
    if (keep) {
      const size_t o = (size_t)q0 + pos;  // outputs are laid out per pair at the pair's query offset
      out_X[3 * o] = (double)pa.x, out_X[3 * o + 1] = (double)pa.y, out_X[3 * o + 2] = (double)pa.z;
      if (mode != 1) {
        // K^-1 [u v 1]^T (stereogeom->get_K().inverse() * uv, :117-118): rows 0 and 1
        out_uv[2 * o] = Kinv[0] * (double)u + Kinv[1] * (double)v + Kinv[2];
        out_uv[2 * o + 1] = Kinv[3] * (double)u + Kinv[4] * (double)v + Kinv[5];
        out_uvd[2 * o] = Kinv[0] * (double)ud + Kinv[1] * (double)vd + Kinv[2];
        out_uvd[2 * o + 1] = Kinv[3] * (double)ud + Kinv[4] * (double)vd + Kinv[5];
      } else {
        out_Y[3 * o] = (double)pb.x, out_Y[3 * o + 1] = (double)pb.y, out_Y[3 * o + 2] = (double)pb.z;
      }
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 8; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) out_count[pair] = s_base;
}

// ---------------------------------------------------------------------------------------------
// Stereo block matching (cv::StereoBM::create(ndisp, wsz)->compute, src/utils/CameraGeometry.cpp:81, :410-418) and the
// reference's disparity -> 3-D image loop (:459-520).  Integer SAD work; the per-thread bodies live in stereo_core.h and
// are also run on the CPU by host/stereo_emul.cpp.
//   sbm_prefilter_kernel : thread per pixel (x-Sobel, clamp, row-pair borders)
//   sbm_hsad_kernel      : grid (column segments, rows, pairs), block = ndisp threads (one candidate each): sliding
//                          horizontal window sums -> hsad[row][column][candidate] (u16, candidate fastest: coalesced)
//   sbm_vsad_kernel      : grid (columns, row stripes, pairs), block = ndisp threads: sliding vertical window sums with
//                          clamped rows; per pixel the candidates' SADs meet in shared memory and thread 0 takes the
//                          decision (first minimum, texture / uniqueness tests, parabola fit, ROI)
//   sbm_to3d_kernel      : thread per pixel
// ---------------------------------------------------------------------------------------------
constexpr int kSbmSeg = 64;     // output columns per horizontal-pass thread (21 window columns are summed before the first one: 64 instead of 32 measured 517 -> 472 us per 8 pairs)
constexpr int kSbmStripe = 60;  // rows per vertical-pass block

__global__ void sbm_fill_kernel(int16_t* __restrict__ disp, size_t n, int16_t v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) disp[i] = v;
}

__global__ void sbm_prefilter_kernel(const uint8_t* __restrict__ img, int n_img, int h, int w, int cap, uint8_t* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)h * w;
  if (i >= per * n_img) return;
  const size_t im = i / per, r = i % per;
  out[i] = sbm_prefilter_px(img + im * per, h, w, (int)(r / w), (int)(r % w), cap);
}

__global__ void sbm_hsad_kernel(const uint8_t* __restrict__ PL, const uint8_t* __restrict__ PR, SbmGeom g, uint16_t* __restrict__ hsad,
                                int* __restrict__ htext) {
  const int pair = blockIdx.z, y = blockIdx.y, d = threadIdx.x;
  const int x0 = blockIdx.x * kSbmSeg, x1 = min(x0 + kSbmSeg, g.width1);
  const size_t img = (size_t)pair * g.h * g.w, hs = (size_t)pair * g.h * g.width1;
  sbm_hsad_thread(PL + img, PR + img, g, y, d, x0, x1, hsad + hs * g.ndisp, htext + hs);
}

__global__ void sbm_vsad_kernel(const uint16_t* __restrict__ hsad, const int* __restrict__ htext, SbmGeom g, int16_t* __restrict__ disp) {
  extern __shared__ int s_sad_raw[];  // [ndisp + 2]: one guard element on either side for the sub-pixel fit
  int* s_sad = s_sad_raw + 1;
  const int pair = blockIdx.z, x = blockIdx.x, d = threadIdx.x;
  const int y0 = blockIdx.y * kSbmStripe, y1 = min(y0 + kSbmStripe, g.h);
  const size_t hs = (size_t)pair * g.h * g.width1;
  const uint16_t* H = hsad + hs * g.ndisp;
  const int* T = htext + hs;
  int16_t* out = disp + (size_t)pair * g.h * g.w;
  int run = sbm_vsad_init(H, g, x, d, y0);
  int tsum = d == 0 ? sbm_vtext_init(T, g, x, y0) : 0;
  for (int y = y0; y < y1; ++y) {
    s_sad[d] = run;
    __syncthreads();
    if (d == 0) {
      const int v = sbm_decide(s_sad, g, tsum);
      out[(size_t)y * g.w + g.lofs + x] = (int16_t)(sbm_in_roi(g, y, g.lofs + x) ? v : g.filtered);
      tsum = sbm_vtext_step(T, g, x, y, tsum);
    }
    __syncthreads();
    run = sbm_vsad_step(H, g, x, d, y, run);
  }
}

// Vertical pass, second version (the default for numDisparities == 64, the reference's value; CB_SBM_V1=1 keeps the kernel
// above): ONE WARP per (output column, row stripe), a lane owns candidates 2*lane and 2*lane+1 -- one aligned 32-bit load
// fetches both u16 sums, a warp reads the 128 contiguous bytes of a pixel -- and the per-pixel decision is warp-parallel:
// first minimum = one REDUX.MIN over (sad << 8 | d), uniqueness test = one vote, the two parabola neighbours by shuffle.
// Same integers as sbm_decide (tests compare the two kernels and the oracle).  [The first version's decision was a serial
// scan by thread 0 of a 64-thread block between two block barriers: 6.4 ms for 8 pairs of 480 x 640.]
constexpr int kSbm2Warps = 8;     // adjacent output columns per block
constexpr int kSbm2Stripe = 120;  // rows per warp

__global__ void __launch_bounds__(32 * kSbm2Warps) sbm_vsad2_kernel(const uint16_t* __restrict__ hsad, const int* __restrict__ htext,
                                                                    SbmGeom g, int16_t* __restrict__ disp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = blockIdx.z, x = blockIdx.x * kSbm2Warps + warp;
  if (x >= g.width1) return;
  const int y0 = blockIdx.y * kSbm2Stripe, y1 = min(y0 + kSbm2Stripe, g.h);
  const size_t hs = (size_t)pair * g.h * g.width1;
  const uint32_t* H = reinterpret_cast<const uint32_t*>(hsad + hs * 64);  // [row][column][32 pairs of candidates]
  const int* T = htext + hs;
  int16_t* out = disp + (size_t)pair * g.h * g.w;
  const int d0 = 2 * lane, d1 = 2 * lane + 1;
  int r0 = 0, r1 = 0;
  for (int j = -g.wsz2; j <= g.wsz2; ++j) {
    const uint32_t v = __ldg(H + ((size_t)sbm_clampi(y0 + j, 0, g.h - 1) * g.width1 + x) * 32 + lane);
    r0 += (int)(v & 0xffffu);
    r1 += (int)(v >> 16);
  }
  int tsum = sbm_vtext_init(T, g, x, y0);  // every lane keeps it (broadcast loads)
  for (int y = y0; y < y1; ++y) {
    // next row's two loads first: they do not depend on the decision
    const uint32_t va = __ldg(H + ((size_t)sbm_clampi(y + 1 + g.wsz2, 0, g.h - 1) * g.width1 + x) * 32 + lane);
    const uint32_t vb = __ldg(H + ((size_t)sbm_clampi(y - g.wsz2, 0, g.h - 1) * g.width1 + x) * 32 + lane);
    const unsigned k0 = ((unsigned)r0 << 8) | (unsigned)d0, k1 = ((unsigned)r1 << 8) | (unsigned)d1;
    const unsigned kmin = __reduce_min_sync(0xffffffffu, k0 < k1 ? k0 : k1);
    const int minsad = (int)(kmin >> 8), mind = (int)(kmin & 255u);
    int v = g.filtered;
    if (tsum >= g.texture_threshold) {  // warp-uniform
      const int thresh = minsad + (minsad * g.uniqueness_ratio / 100);
      const bool bad = ((d0 < mind - 1 || d0 > mind + 1) && r0 <= thresh) || ((d1 < mind - 1 || d1 > mind + 1) && r1 <= thresh);
      if (!(g.uniqueness_ratio > 0 && __any_sync(0xffffffffu, bad))) {
        const int ip = mind + 1 == 64 ? 62 : mind + 1, in = mind - 1 < 0 ? 1 : mind - 1;  // the two guard elements
        const int pa = __shfl_sync(0xffffffffu, r0, ip >> 1), pb = __shfl_sync(0xffffffffu, r1, ip >> 1);
        const int na = __shfl_sync(0xffffffffu, r0, in >> 1), nb = __shfl_sync(0xffffffffu, r1, in >> 1);
        const int p = (ip & 1) ? pb : pa, n = (in & 1) ? nb : na;
        const int dd = p + n - 2 * minsad + (p > n ? p - n : n - p);
        v = (((64 - mind - 1 + g.mindisp) * 256 + (dd != 0 ? (p - n) * 256 / dd : 0) + 15) >> 4);
      }
    }
    if (lane == 0) out[(size_t)y * g.w + g.lofs + x] = (int16_t)(sbm_in_roi(g, y, g.lofs + x) ? v : g.filtered);
    r0 += (int)(va & 0xffffu) - (int)(vb & 0xffffu);
    r1 += (int)(va >> 16) - (int)(vb >> 16);
    tsum = sbm_vtext_step(T, g, x, y, tsum);
  }
}

__global__ void sbm_to3d_kernel(const int16_t* __restrict__ disp, int n, int h, int w, float Q03, float Q13, float Q23, float Q32,
                                float Q33, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = (size_t)h * w;
  if (i >= per * n) return;
  const size_t r = i % per;
  sbm_point3d(disp[i], (int)(r / w), (int)(r % w), Q03, Q13, Q23, Q32, Q33, out + i * 3);
}

int grow_dev(void** p, size_t* cur, size_t need) {
  if (*cur >= need) return CB_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cur = 0;
  cudaError_t e = cudaMalloc(p, need);
  if (e != cudaSuccess) return cb::fail(CB_ENOMEM, "cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
  *cur = need;
  return CB_OK;
}

}  // namespace

struct cb_frontend {
  int device = 0, max_pairs = 0, max_features = 0;
  cudaStream_t stream = nullptr;
  // device copies of one batch (sized for max_pairs * max_features at create time)
  uint32_t* d1 = nullptr;
  uint32_t* d2 = nullptr;
  float* kp1 = nullptr;
  float* kp2 = nullptr;
  int* off1 = nullptr;
  int* off2 = nullptr;
  unsigned long long* best = nullptr;
  // tensor-core matcher: descriptors expanded to one byte per bit + their popcounts (CB_MATCH_SIMT=1 keeps the popc kernel)
  bool match_simt = false;
  bool sbm_v1 = false;  // CB_SBM_V1=1: block-per-column vertical pass with the serial decision (cross-check of the default)
  uint8_t* e1 = nullptr;
  uint8_t* e2 = nullptr;
  int* pop1 = nullptr;
  int* pop2 = nullptr;
  CUtensorMap tmQ, tmT;
  int sm_count = 0;
  int* train_idx = nullptr;
  int* dist = nullptr;
  unsigned char* mask = nullptr;
  int* n_inl = nullptr;
  unsigned int* tickets = nullptr;
  double* X = nullptr;
  double* uv = nullptr;
  double* uvd = nullptr;
  double* Y = nullptr;
  double* Kinv = nullptr;
  int* count = nullptr;
  float* img_a = nullptr;
  float* img_b = nullptr;
  size_t img_bytes_a = 0, img_bytes_b = 0;
  // shape of the batch currently resident (set by cb_frontend_match_gms, used by cb_frontend_collect)
  int n_pairs = 0, total1 = 0, total2 = 0;
  bool have_matches = false;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  float last_match_ms = 0.f;
  // stereo block matching scratch (grown on demand)
  uint8_t* sb_img = nullptr;   // left | right of the current chunk
  uint8_t* sb_pre = nullptr;   // pre-filtered left | right
  uint16_t* sb_hsad = nullptr;
  int* sb_htext = nullptr;
  int16_t* sb_disp = nullptr;
  float* sb_3d = nullptr;
  size_t sb_img_bytes = 0, sb_pre_bytes = 0, sb_hsad_bytes = 0, sb_htext_bytes = 0, sb_disp_bytes = 0, sb_3d_bytes = 0;
  float last_stereo_ms = 0.f;
};

extern "C" {

int cb_frontend_create(cb_frontend** out, int max_pairs, int max_features, int device) {
  if (!out) return cb::fail(CB_EINVAL, "out is NULL");
  *out = nullptr;
  if (max_pairs < 1 || max_features < 1 || max_features > 16384)
    return cb::fail(CB_EINVAL, "max_pairs must be >= 1 and max_features in [1, 16384]");
  int sm_count = 0;
  int rc = cb::select_device(device, &sm_count);
  if (rc) return rc;
  cb::DeviceGuard g(device);
  cb_frontend* f = new cb_frontend();
  f->device = device;
  f->sm_count = sm_count;
  const char* env = getenv("CB_MATCH_SIMT");
  f->match_simt = env && env[0] == '1';
  if (const char* e2 = getenv("CB_SBM_V1")) f->sbm_v1 = e2[0] == '1';
  f->max_pairs = max_pairs;
  f->max_features = max_features;
  const size_t tot = (size_t)max_pairs * max_features;
  cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
  auto A = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  A((void**)&f->d1, tot * 32);
  A((void**)&f->d2, tot * 32);
  A((void**)&f->kp1, tot * 8);
  A((void**)&f->kp2, tot * 8);
  A((void**)&f->off1, (size_t)(max_pairs + 1) * 4);
  A((void**)&f->off2, (size_t)(max_pairs + 1) * 4);
  A((void**)&f->best, tot * 8);
  const size_t tot_tm = tot < (size_t)kHtN ? (size_t)kHtN : tot;  // the TMA boxes are up to 256 rows tall
  if (!f->match_simt) {
    A((void**)&f->e1, tot_tm * 256);
    A((void**)&f->e2, tot_tm * 256);
    A((void**)&f->pop1, tot_tm * 4);
    A((void**)&f->pop2, tot_tm * 4);
  }
  A((void**)&f->train_idx, tot * 4);
  A((void**)&f->dist, tot * 4);
  A((void**)&f->mask, tot);
  A((void**)&f->n_inl, (size_t)max_pairs * 4);
  A((void**)&f->tickets, (size_t)max_pairs * 4);
  if (e == cudaSuccess) e = cudaMemset(f->tickets, 0, (size_t)max_pairs * 4);
  A((void**)&f->X, tot * 24);
  A((void**)&f->uv, tot * 16);
  A((void**)&f->uvd, tot * 16);
  A((void**)&f->Y, tot * 24);
  A((void**)&f->Kinv, 9 * 8);
  A((void**)&f->count, (size_t)max_pairs * 4);
  if (e == cudaSuccess) e = cudaEventCreate(&f->ev[0]);
  if (e == cudaSuccess) e = cudaEventCreate(&f->ev[1]);
  if (e != cudaSuccess) {
    cb_frontend_destroy(f);
    return cb::fail(CB_ENOMEM, "front-end allocation failed: %s", cudaGetErrorString(e));
  }
  if (!f->match_simt) {
    rc = make_map_2d_u8(&f->tmQ, f->e1, (uint64_t)tot_tm, 256, kHtM);
    if (!rc) rc = make_map_2d_u8(&f->tmT, f->e2, (uint64_t)tot_tm, 256, kHtN);
    if (!rc && cudaFuncSetAttribute(hamming_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHtSmem) != cudaSuccess)
      rc = cb::fail(CB_ECUDA, "cannot reserve %d bytes of shared memory for hamming_tc_kernel", kHtSmem);
    if (rc) {
      cb_frontend_destroy(f);
      return rc;
    }
  }
  *out = f;
  return CB_OK;
}

int cb_frontend_destroy(cb_frontend* f) {
  if (!f) return CB_OK;
  cb::DeviceGuard g(f->device);
  if (f->stream) cb::sync_stream(f->stream);
  void* ps[] = {f->d1, f->d2, f->kp1, f->kp2, f->off1, f->off2, f->best, f->train_idx, f->dist, f->mask, f->n_inl,
                f->X,  f->uv, f->uvd, f->Y,   f->Kinv, f->count, f->img_a, f->img_b, f->e1, f->e2, f->pop1, f->pop2, f->tickets, f->sb_img, f->sb_pre, f->sb_hsad, f->sb_htext,
                f->sb_disp, f->sb_3d};
  for (void* p : ps)
    if (p) cudaFree(p);
  for (cudaEvent_t ev : f->ev)
    if (ev) cudaEventDestroy(ev);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
  return CB_OK;
}

int cb_frontend_match_gms(cb_frontend* f, int n_pairs, const int32_t* off1, const int32_t* off2, const uint8_t* desc1,
                          const uint8_t* desc2, const float* kp1, const float* kp2, int width1, int height1, int width2,
                          int height2, int32_t* train_idx, int32_t* distance, uint8_t* inlier_mask, int32_t* n_inliers) {
  if (!f || !off1 || !off2 || !desc1 || !desc2 || !kp1 || !kp2 || !train_idx || !inlier_mask || !n_inliers)
    return cb::fail(CB_EINVAL, "NULL argument to cb_frontend_match_gms");
  if (n_pairs < 1 || n_pairs > f->max_pairs) return cb::fail(CB_EINVAL, "n_pairs %d outside [1,%d]", n_pairs, f->max_pairs);
  if (width1 < 1 || height1 < 1 || width2 < 1 || height2 < 1) return cb::fail(CB_EINVAL, "bad image size");
  if (off1[0] != 0 || off2[0] != 0) return cb::fail(CB_EINVAL, "offsets must start at 0");
  int max1 = 0, max2 = 0;
  for (int p = 0; p < n_pairs; ++p) {
    const int a = off1[p + 1] - off1[p], b = off2[p + 1] - off2[p];
    if (a < 0 || b < 0 || a > f->max_features || b > f->max_features)
      return cb::fail(CB_EINVAL, "pair %d has %d / %d features, outside [0,%d]", p, a, b, f->max_features);
    max1 = a > max1 ? a : max1;
    max2 = b > max2 ? b : max2;
  }
  const int tot1 = off1[n_pairs], tot2 = off2[n_pairs];
  cb::DeviceGuard g(f->device);
  cudaStream_t st = f->stream;
  f->have_matches = false;
  CB_CUDA(cudaMemcpyAsync(f->off1, off1, (size_t)(n_pairs + 1) * 4, cudaMemcpyHostToDevice, st));
  CB_CUDA(cudaMemcpyAsync(f->off2, off2, (size_t)(n_pairs + 1) * 4, cudaMemcpyHostToDevice, st));
  if (tot1) CB_CUDA(cudaMemcpyAsync(f->d1, desc1, (size_t)tot1 * 32, cudaMemcpyHostToDevice, st));
  if (tot2) CB_CUDA(cudaMemcpyAsync(f->d2, desc2, (size_t)tot2 * 32, cudaMemcpyHostToDevice, st));
  if (tot1) CB_CUDA(cudaMemcpyAsync(f->kp1, kp1, (size_t)tot1 * 8, cudaMemcpyHostToDevice, st));
  if (tot2) CB_CUDA(cudaMemcpyAsync(f->kp2, kp2, (size_t)tot2 * 8, cudaMemcpyHostToDevice, st));
  if (tot1) {
    CB_CUDA(cudaEventRecord(f->ev[0], st));
    CB_CUDA(cudaMemsetAsync(f->best, 0xff, (size_t)tot1 * 8, st));
    if (max2 > 0 && !f->match_simt) {
      expand_bits_kernel<<<(unsigned)((tot1 * kDescWords + 255) / 256), 256, 0, st>>>(f->d1, tot1, f->e1, f->pop1);
      CB_LAUNCH_CHECK();
      expand_bits_kernel<<<(unsigned)((tot2 * kDescWords + 255) / 256), 256, 0, st>>>(f->d2, tot2, f->e2, f->pop2);
      CB_LAUNCH_CHECK();
      const int m_tiles = (max1 + kHtM - 1) / kHtM, n_tiles = (max2 + kHtN - 1) / kHtN;
      // split the train tiles over gridDim.y until the grid fills the machine about twice
      int splits = 1;
      while (splits < n_tiles && (long long)m_tiles * n_pairs * splits < 2LL * f->sm_count) ++splits;
      const int tiles_per_split = (n_tiles + splits - 1) / splits;
      dim3 grid((unsigned)m_tiles, (unsigned)((n_tiles + tiles_per_split - 1) / tiles_per_split), (unsigned)n_pairs);
      hamming_tc_kernel<<<grid, kHtThreads, kHtSmem, st>>>(f->tmQ, f->tmT, f->off1, f->off2, f->pop1, f->pop2, tiles_per_split, f->best);
      CB_LAUNCH_CHECK();
    } else if (max2 > 0) {
      dim3 grid((unsigned)((max1 + kMatchThreads - 1) / kMatchThreads), (unsigned)((max2 + kTrainChunk - 1) / kTrainChunk),
                (unsigned)n_pairs);
      hamming_match_kernel<<<grid, kMatchThreads, 0, st>>>(f->d1, f->d2, f->off1, f->off2, f->best);
      CB_LAUNCH_CHECK();
    }
    unpack_matches_kernel<<<(unsigned)((tot1 + 255) / 256), 256, 0, st>>>(f->best, tot1, f->train_idx, f->dist);
    CB_LAUNCH_CHECK();
  }
  {
    const size_t smem = (size_t)max1 * 6 + 8 + (size_t)(4 * kCells + 1) * 4;
    CB_CUDA(cudaFuncSetAttribute(gms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (tot1) CB_CUDA(cudaMemsetAsync(f->mask, 0, (size_t)tot1, st));
    gms_kernel<<<dim3((unsigned)n_pairs, 4), kGmsThreads, smem, st>>>(f->kp1, f->kp2, f->off1, f->off2, f->train_idx, width1, height1,
                                                                      width2, height2, max1, f->mask, f->n_inl, f->tickets);
    CB_LAUNCH_CHECK();
  }
  if (tot1) {
    CB_CUDA(cudaEventRecord(f->ev[1], st));
    CB_CUDA(cudaMemcpyAsync(train_idx, f->train_idx, (size_t)tot1 * 4, cudaMemcpyDeviceToHost, st));
    if (distance) CB_CUDA(cudaMemcpyAsync(distance, f->dist, (size_t)tot1 * 4, cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaMemcpyAsync(inlier_mask, f->mask, (size_t)tot1, cudaMemcpyDeviceToHost, st));
  }
  CB_CUDA(cudaMemcpyAsync(n_inliers, f->n_inl, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, st));
  CB_CUDA(cb::sync_stream(st));
  if (tot1) cudaEventElapsedTime(&f->last_match_ms, f->ev[0], f->ev[1]);
  f->n_pairs = n_pairs;
  f->total1 = tot1;
  f->total2 = tot2;
  f->have_matches = true;
  return CB_OK;
}

float cb_frontend_last_match_ms(const cb_frontend* f) { return f ? f->last_match_ms : -1.f; }

int cb_frontend_stereo_bm(cb_frontend* f, int n, const uint8_t* left, const uint8_t* right, int rows, int cols, int ndisp, int wsz,
                          int16_t* disparity) {
  if (!f || !left || !right || !disparity) return cb::fail(CB_EINVAL, "NULL argument to cb_frontend_stereo_bm");
  if (n < 1 || rows < 2 || cols < 2) return cb::fail(CB_EINVAL, "bad batch / image size");
  if (ndisp < 16 || ndisp > 256 || ndisp % 16) return cb::fail(CB_EINVAL, "numDisparities must be a multiple of 16 in [16, 256]");
  if (wsz < 5 || wsz > 255 || wsz % 2 == 0 || wsz >= rows || wsz >= cols)
    return cb::fail(CB_EINVAL, "SADWindowSize must be odd, within 5..255 and smaller than the image");
  cb::DeviceGuard gd(f->device);
  cudaStream_t st = f->stream;
  const SbmGeom g = sbm_make_geom(rows, cols, ndisp, wsz);
  const size_t per = (size_t)rows * cols;
  if (g.lofs >= cols || g.rofs >= cols || g.width1 < 1) {  // OpenCV: everything FILTERED
    for (size_t i = 0; i < per * n; ++i) disparity[i] = (int16_t)g.filtered;
    return CB_OK;
  }
  // pairs per chunk: keep the horizontal sums below ~1 GB
  const size_t hsad_per = (size_t)rows * g.width1 * ndisp * sizeof(uint16_t);
  int chunk = (int)((size_t)(1u << 30) / hsad_per);
  chunk = chunk < 1 ? 1 : (chunk > n ? n : chunk);
  int rc = grow_dev((void**)&f->sb_img, &f->sb_img_bytes, 2 * per * chunk);
  if (!rc) rc = grow_dev((void**)&f->sb_pre, &f->sb_pre_bytes, 2 * per * chunk);
  if (!rc) rc = grow_dev((void**)&f->sb_hsad, &f->sb_hsad_bytes, hsad_per * chunk);
  if (!rc) rc = grow_dev((void**)&f->sb_htext, &f->sb_htext_bytes, (size_t)rows * g.width1 * sizeof(int) * chunk);
  if (!rc) rc = grow_dev((void**)&f->sb_disp, &f->sb_disp_bytes, per * sizeof(int16_t) * chunk);
  if (rc) return rc;
  float total_ms = 0.f;
  for (int c0 = 0; c0 < n; c0 += chunk) {
    const int nc = n - c0 < chunk ? n - c0 : chunk;
    CB_CUDA(cudaMemcpyAsync(f->sb_img, left + (size_t)c0 * per, per * nc, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaMemcpyAsync(f->sb_img + per * nc, right + (size_t)c0 * per, per * nc, cudaMemcpyHostToDevice, st));
    CB_CUDA(cudaEventRecord(f->ev[0], st));
    sbm_fill_kernel<<<(unsigned)((per * nc + 255) / 256), 256, 0, st>>>(f->sb_disp, per * nc, (int16_t)g.filtered);
    CB_LAUNCH_CHECK();
    sbm_prefilter_kernel<<<(unsigned)((2 * per * nc + 255) / 256), 256, 0, st>>>(f->sb_img, 2 * nc, rows, cols, g.cap, f->sb_pre);
    CB_LAUNCH_CHECK();
    sbm_hsad_kernel<<<dim3((unsigned)((g.width1 + kSbmSeg - 1) / kSbmSeg), (unsigned)rows, (unsigned)nc), ndisp, 0, st>>>(
        f->sb_pre, f->sb_pre + per * nc, g, f->sb_hsad, f->sb_htext);
    CB_LAUNCH_CHECK();
    if (ndisp == 64 && !f->sbm_v1)
      sbm_vsad2_kernel<<<dim3((unsigned)((g.width1 + kSbm2Warps - 1) / kSbm2Warps), (unsigned)((rows + kSbm2Stripe - 1) / kSbm2Stripe),
                              (unsigned)nc), 32 * kSbm2Warps, 0, st>>>(f->sb_hsad, f->sb_htext, g, f->sb_disp);
    else
      sbm_vsad_kernel<<<dim3((unsigned)g.width1, (unsigned)((rows + kSbmStripe - 1) / kSbmStripe), (unsigned)nc), ndisp,
                        (size_t)(ndisp + 2) * sizeof(int), st>>>(f->sb_hsad, f->sb_htext, g, f->sb_disp);
    CB_LAUNCH_CHECK();
    CB_CUDA(cudaEventRecord(f->ev[1], st));
    CB_CUDA(cudaMemcpyAsync(disparity + (size_t)c0 * per, f->sb_disp, per * sizeof(int16_t) * nc, cudaMemcpyDeviceToHost, st));
    CB_CUDA(cb::sync_stream(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, f->ev[0], f->ev[1]);
    total_ms += ms;
  }
  f->last_stereo_ms = total_ms;
  return CB_OK;
}

float cb_frontend_last_stereo_ms(const cb_frontend* f) { return f ? f->last_stereo_ms : -1.f; }

int cb_frontend_disparity_to_3d(cb_frontend* f, int n, const int16_t* disparity, int rows, int cols, float Q03, float Q13, float Q23,
                                float Q32, float Q33, float* out3d) {
  if (!f || !disparity || !out3d) return cb::fail(CB_EINVAL, "NULL argument to cb_frontend_disparity_to_3d");
  if (n < 1 || rows < 1 || cols < 1) return cb::fail(CB_EINVAL, "bad batch / image size");
  cb::DeviceGuard gd(f->device);
  cudaStream_t st = f->stream;
  const size_t per = (size_t)rows * cols;
  int rc = grow_dev((void**)&f->sb_disp, &f->sb_disp_bytes, per * sizeof(int16_t) * n);
  if (!rc) rc = grow_dev((void**)&f->sb_3d, &f->sb_3d_bytes, per * 3 * sizeof(float) * n);
  if (rc) return rc;
  CB_CUDA(cudaMemcpyAsync(f->sb_disp, disparity, per * sizeof(int16_t) * n, cudaMemcpyHostToDevice, st));
  sbm_to3d_kernel<<<(unsigned)((per * n + 255) / 256), 256, 0, st>>>(f->sb_disp, n, rows, cols, Q03, Q13, Q23, Q32, Q33, f->sb_3d);
  CB_LAUNCH_CHECK();
  CB_CUDA(cudaMemcpyAsync(out3d, f->sb_3d, per * 3 * sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  CB_CUDA(cb::sync_stream(st));
  return CB_OK;
}

int cb_frontend_collect(cb_frontend* f, int mode, const float* img3d_a, const float* img3d_b, int rows, int cols,
                        const double* K_inverse, int32_t* counts, double* X, double* uv, double* uv_d, double* Y) {
  if (!f || !counts || !X) return cb::fail(CB_EINVAL, "NULL argument to cb_frontend_collect");
  if (!f->have_matches) return cb::fail(CB_EINVAL, "cb_frontend_collect needs a preceding cb_frontend_match_gms on this handle");
  if (mode < 0 || mode > 2) return cb::fail(CB_EINVAL, "mode must be 0 (3D-2D), 1 (3D-3D) or 2 (3D-2D, frames swapped)");
  if (mode != 2 && !img3d_a) return cb::fail(CB_EINVAL, "modes 0 and 1 need img3d_a");
  if (mode != 1 && (!K_inverse || !uv || !uv_d)) return cb::fail(CB_EINVAL, "3D-2D modes need K_inverse, uv and uv_d");
  if (mode == 1 && !Y) return cb::fail(CB_EINVAL, "3D-3D mode needs Y");
  if (mode != 0 && !img3d_b) return cb::fail(CB_EINVAL, "modes 1 and 2 need img3d_b");
  if (rows < 1 || cols < 1) return cb::fail(CB_EINVAL, "bad depth-image size");
  cb::DeviceGuard g(f->device);
  cudaStream_t st = f->stream;
  const size_t img_bytes = (size_t)f->n_pairs * rows * cols * 3 * sizeof(float);
  int rc = img3d_a ? grow_dev((void**)&f->img_a, &f->img_bytes_a, img_bytes) : CB_OK;
  if (!rc && img3d_b) rc = grow_dev((void**)&f->img_b, &f->img_bytes_b, img_bytes);
  if (rc) return rc;
  if (img3d_a) CB_CUDA(cudaMemcpyAsync(f->img_a, img3d_a, img_bytes, cudaMemcpyHostToDevice, st));
  if (img3d_b) CB_CUDA(cudaMemcpyAsync(f->img_b, img3d_b, img_bytes, cudaMemcpyHostToDevice, st));
  if (K_inverse) CB_CUDA(cudaMemcpyAsync(f->Kinv, K_inverse, 9 * sizeof(double), cudaMemcpyHostToDevice, st));
  collect_kernel<<<f->n_pairs, 256, 0, st>>>(f->kp1, f->kp2, f->off1, f->off2, f->train_idx, f->mask, img3d_a ? f->img_a : nullptr,
                                             img3d_b ? f->img_b : nullptr, rows, cols, f->Kinv, mode, f->X, f->uv, f->uvd, f->Y, f->count);
  CB_LAUNCH_CHECK();
  const size_t n = (size_t)f->total1;
  CB_CUDA(cudaMemcpyAsync(counts, f->count, (size_t)f->n_pairs * 4, cudaMemcpyDeviceToHost, st));
  if (n) {
    CB_CUDA(cudaMemcpyAsync(X, f->X, n * 24, cudaMemcpyDeviceToHost, st));
    if (mode != 1) {
      CB_CUDA(cudaMemcpyAsync(uv, f->uv, n * 16, cudaMemcpyDeviceToHost, st));
      CB_CUDA(cudaMemcpyAsync(uv_d, f->uvd, n * 16, cudaMemcpyDeviceToHost, st));
    } else {
      CB_CUDA(cudaMemcpyAsync(Y, f->Y, n * 24, cudaMemcpyDeviceToHost, st));
    }
  }
  CB_CUDA(cb::sync_stream(st));
  return CB_OK;
}

}  // extern "C"
