// Batched DLS-PnP RANSAC (sm_100a), warp-per-hypothesis.
//
// Replaces (reference): StaticTheiaPoseCompute::PNP (src/DlsPnpWithRansac.cpp:132-245) =
// theia::Ransac<DlsPnpWithRansac>: sample 15 -> theia::DlsPnp (src/DlsPnpWithRansac.h:61) ->
// accept only a unique solution (:62-71) -> L1 reprojection residuals (:75-99) -> MLE cost.
// Theia's source is not part of the reference tree; the solver implements the published
// DLS-PnP algorithm (Hesch & Roumeliotis, ICCV 2011) -- see oracle/dls_pnp.py for the maths and
// tools/gen_dls_tables.py for the derivation of the elimination template and a numpy prototype
// of exactly this control flow.
//
// Stages (one launch each per chunk of hypotheses; intermediates stay in HBM/L2):
//   setup      thread per hypothesis : sample 15 correspondences, build the quartic cost J'
//                                      (10x10 Q) and the 3x20 gradient coefficients, T (3x9)
//   eliminate  warp per hypothesis   : block-triangular Macaulay elimination (five dense
//                                      Gauss-Jordan solves 3,9,18,27,36 with 27 right-hand sides
//                                      in shared memory) -> 27x27 action matrix S
//   roots      warp per hypothesis   : Householder-Hessenberg + Francis double-shift QR
//                                      (eigenvalues), inverse iteration per real eigenvalue
//                                      -> (s1,s2,s3) -> R,t -> cheirality; model iff 1 solution
//   score      warp per hypothesis   : residuals over all n points, MLE cost, inlier count
//   select     warp per candidate    : replay of Theia's sequential adaptive termination
// fp64 throughout (the elimination is ill-conditioned in fp32).  Bound: FP64 pipe / shared-memory
// bandwidth, not HBM (inputs are 40 B per correspondence).
#include "common.cuh"
#include "dls_tables.h"

#include <math_constants.h>

namespace {

using cb::FULL;

constexpr int kSample = 15;       // DlsPnpWithRansac::SampleSize (DlsPnpWithRansac.h:45)
constexpr int kN = 27;            // quotient ring dimension / action matrix size
constexpr int kMaxSol = 27;

// ---------------------------------------------------------------------------------------------
// counter-based sampler, bit-identical to oracle/dls_pnp.py:sample_indices
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
constexpr unsigned long long kGold = 0x9E3779B97F4A7C15ULL;

__device__ inline void sample_indices(unsigned long long seed, int cand, int hyp, int n, int* out, int k = kSample) {
  unsigned long long key = mix64(seed + kGold * (unsigned long long)(cand + 1));
  key = mix64(key + kGold * (unsigned long long)(hyp + 1));
  int cnt = 0;
  unsigned long long ctr = 0;
  while (cnt < k) {
    ++ctr;
    const unsigned long long r = mix64(key + kGold * ctr);
    const int idx = (int)(((r >> 32) * (unsigned long long)n) >> 32);
    bool dup = false;
    for (int j = 0; j < cnt; ++j) dup |= (out[j] == idx);
    if (!dup) out[cnt++] = idx;
  }
}

// ---------------------------------------------------------------------------------------------
// stage 1: setup (thread per hypothesis)
// ---------------------------------------------------------------------------------------------
struct SetupArgs {
  const int* offsets;     // [n_cand+1]
  const double* X;        // [total][3]
  const double* uv;       // [total][2]
  const int* samples;     // optional [n_cand][H][15]
  unsigned long long seed;
  int H;                  // hypotheses per candidate
  long long g0;           // first global hypothesis of this chunk
  int count;              // hypotheses in this chunk
  int* idx_out;           // [count][15]  (global point indices)
  double* coef_out;       // [count][60]
  double* T_out;          // [count][27]  row-major 3x9
  int* status;            // [count] 0 ok, -1 candidate refused (<20 points) / degenerate
};

__device__ inline void sym3_inv(const double* a /*xx,xy,xz,yy,yz,zz*/, double* inv) {
  const double a00 = a[0], a01 = a[1], a02 = a[2], a11 = a[3], a12 = a[4], a22 = a[5];
  const double c00 = a11 * a22 - a12 * a12;
  const double c01 = a02 * a12 - a01 * a22;
  const double c02 = a01 * a12 - a02 * a11;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  const double id = 1.0 / det;
  inv[0] = c00 * id;
  inv[1] = c01 * id;
  inv[2] = c02 * id;
  inv[3] = (a00 * a22 - a02 * a02) * id;
  inv[4] = (a01 * a02 - a00 * a12) * id;
  inv[5] = (a00 * a11 - a01 * a01) * id;
}

__device__ __forceinline__ int sym_idx(int i, int j) {  // 3x3 symmetric packed xx,xy,xz,yy,yz,zz
  const int a = i < j ? i : j, b = i < j ? j : i;
  return a == 0 ? b : (a == 1 ? 2 + b : 5);
}

__global__ void __launch_bounds__(128) dls_setup_kernel(SetupArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.count) return;
  const long long g = a.g0 + t;
  const int cand = (int)(g / a.H), hyp = (int)(g % a.H);
  const int p0 = a.offsets[cand], n = a.offsets[cand + 1] - p0;
  if (n < 20) {  // DlsPnpWithRansac.cpp:136-139
    a.status[t] = -1;
    return;
  }
  int idx[kSample];
  if (a.samples) {
    const int* s = a.samples + ((size_t)cand * a.H + hyp) * kSample;
    for (int i = 0; i < kSample; ++i) idx[i] = s[i];
  } else {
    sample_indices(a.seed, cand, hyp, n, idx);
  }
  for (int i = 0; i < kSample; ++i) {
    if (idx[i] < 0 || idx[i] >= n) {
      a.status[t] = -1;
      return;
    }
    idx[i] += p0;
    a.idx_out[(size_t)t * kSample + i] = idx[i];
  }

  // sums over the 15 correspondences (paper eq. 10-17; oracle/dls_pnp.py:dls_setup)
  double SF[6] = {0, 0, 0, 0, 0, 0};   // sum f f^T
  double B[3][6];                      // B_a = sum r_a (f f^T - I)          (symmetric)
  double S[6][6];                      // S_ab = sum r_a r_b (I - f f^T)     (ab packed, symmetric)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 6; ++j) B[i][j] = 0.0;
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) S[i][j] = 0.0;
  for (int i = 0; i < kSample; ++i) {
    const double u = a.uv[2 * (size_t)idx[i]], v = a.uv[2 * (size_t)idx[i] + 1];
    const double inv = 1.0 / sqrt(u * u + v * v + 1.0);
    const double f[3] = {u * inv, v * inv, inv};
    const double r[3] = {a.X[3 * (size_t)idx[i]], a.X[3 * (size_t)idx[i] + 1], a.X[3 * (size_t)idx[i] + 2]};
    double F[6] = {f[0] * f[0], f[0] * f[1], f[0] * f[2], f[1] * f[1], f[1] * f[2], f[2] * f[2]};
    double FmI[6] = {F[0] - 1.0, F[1], F[2], F[3] - 1.0, F[4], F[5] - 1.0};
    for (int j = 0; j < 6; ++j) SF[j] += F[j];
    for (int q = 0; q < 3; ++q)
      for (int j = 0; j < 6; ++j) B[q][j] += r[q] * FmI[j];
    int ab = 0;
    for (int q = 0; q < 3; ++q)
      for (int w = q; w < 3; ++w, ++ab) {
        const double rr = r[q] * r[w];
        for (int j = 0; j < 6; ++j) S[ab][j] -= rr * FmI[j];
      }
  }
  double Hinv[6] = {kSample - SF[0], -SF[1], -SF[2], kSample - SF[3], -SF[4], kSample - SF[5]};
  double Hm[6];
  sym3_inv(Hinv, Hm);

  // T = [H B_0, H B_1, H B_2]  (3x9, row-major), G (9x9) block(a,b) = S_ab - B_a H B_b
  double T[3][9];
  for (int q = 0; q < 3; ++q)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += Hm[sym_idx(i, k)] * B[q][sym_idx(k, j)];
        T[i][3 * q + j] = s;
      }
  double G[9][9];
  for (int q = 0; q < 3; ++q)
    for (int w = 0; w < 3; ++w) {
      const int ab = q <= w ? (q == 0 ? w : (q == 1 ? 2 + w : 5)) : (w == 0 ? q : (w == 1 ? 2 + q : 5));
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double s = S[ab][sym_idx(i, j)];
          for (int k = 0; k < 3; ++k) s -= B[q][sym_idx(i, k)] * T[k][3 * w + j];
          G[3 * q + i][3 * w + j] = s;
        }
    }
  // Q = A^T G A  (10x10), A = kCayleyA (9x10, entries in {-2..2})
  double W[9][10];
  for (int p = 0; p < 9; ++p)
    for (int b = 0; b < 10; ++b) {
      double s = 0.0;
      for (int q = 0; q < 9; ++q) {
        const int c = dls::kCayleyA[q][b];
        if (c) s += G[p][q] * (double)c;
      }
      W[p][b] = s;
    }
  double Q[10][10];
  for (int aa = 0; aa < 10; ++aa)
    for (int b = 0; b < 10; ++b) {
      double s = 0.0;
      for (int p = 0; p < 9; ++p) {
        const int c = dls::kCayleyA[p][aa];
        if (c) s += W[p][b] * (double)c;
      }
      Q[aa][b] = s;
    }
  for (int aa = 0; aa < 10; ++aa)
    for (int b = aa + 1; b < 10; ++b) {
      const double m = 0.5 * (Q[aa][b] + Q[b][aa]);
      Q[aa][b] = m;
      Q[b][aa] = m;
    }
  double coef[60];
  for (int i = 0; i < 60; ++i) coef[i] = 0.0;
  for (int k = 0; k < 3; ++k)
    for (int aa = 0; aa < 10; ++aa) {
      const int c = dls::kGradCoef[k][aa];
      if (!c) continue;
      for (int b = 0; b < 10; ++b) coef[k * 20 + dls::kGradMono[k][aa][b]] += (double)c * Q[aa][b];
    }
  double* co = a.coef_out + (size_t)t * 60;
  for (int i = 0; i < 60; ++i) co[i] = coef[i];
  double* To = a.T_out + (size_t)t * 27;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 9; ++j) To[i * 9 + j] = T[i][j];
  bool fin = true;
  for (int i = 0; i < 60; ++i) fin = fin && isfinite(coef[i]);
  a.status[t] = fin ? 0 : -1;
}

// ---------------------------------------------------------------------------------------------
// stage 2: block-triangular elimination (one CTA of 4 warps per hypothesis)
// ---------------------------------------------------------------------------------------------
// shared memory per CTA: N[60][27] (normal forms of degrees 3..6 + the 3 border monomials of degree 7)
//                        | AUG[36][55] (work matrix) | coef[60] | part[4][3][27] | perm[36]
// Blocks of degree 3..6: Gauss-Jordan on [D | R] (n x (n+27)).  Block of degree 7: only three rows of
// D^-1 R are ever needed (the border monomials s_k*(2,2,2)), so solve D^T Y = E_J (36 x 3 right-hand
// sides) and contract Y with R row by row without storing R.
// Pivoting: partial, WITHOUT physical row swaps (a `used` mask + permutation), pivot-row scaling is
// deferred to the read-out; every warp owns the rows i == warp (mod 4); one __syncthreads per pivot.
constexpr int kElimWarps = 4;
constexpr int kElimThreads = 32 * kElimWarps;
constexpr int kAug = 55;  // odd stride (doubles) >= 27 + 27
constexpr int kNRows = 60;
constexpr int kElimSmemBytes = (kNRows * kN + 36 * kAug + 60 + kElimWarps * 3 * kN) * 8 + 36 * 4;

// Build one row of the (reduced | same-degree | lower-degree) split of f_i * x^mult.
//   direct scatter : reduced monomial cd -> dst_r[cd] -= c ; same-degree non-reduced -> (*put_d)(col, c)
//   lower degree   : dst_r[0..26] -= c * N[cd-27][.]
// Lanes 0..19 own the 20 terms; lanes 0..26 own the 27 right-hand-side columns.  Returns this lane's
// right-hand-side accumulator (lower-degree part only); the direct terms are scattered by the caller.
__device__ __forceinline__ double row_lower_part(const double* __restrict__ N, double c, int cd, int o0, int lane) {
  const bool lower = lane < 20 && cd >= kN && (cd - kN) < o0;
  unsigned m = __ballot_sync(FULL, lower);
  double acc = 0.0;
  while (m) {
    const int b = __ffs(m) - 1;
    m &= m - 1;
    const double cc = __shfl_sync(FULL, c, b);
    const int beta = __shfl_sync(FULL, cd, b) - kN;
    if (lane < kN) acc = fma(-cc, N[beta * kN + lane], acc);
  }
  return acc;
}

// Gauss-Jordan elimination of the first n columns of AUG (n rows, ncol columns), partial pivoting by
// permutation.  On return row perm[k] holds unknown k, still scaled by its pivot AUG[perm[k]][k].
__device__ __forceinline__ void block_gauss_jordan(double* AUG, int* perm, int n, int ncol, int warp, int lane) {
  unsigned long long used = 0ull;  // identical in every thread (deterministic arg-max)
  for (int k = 0; k < n; ++k) {
    // arg-max |AUG[i][k]| over unused rows (every warp computes it redundantly: no barrier needed)
    double best = -1.0;
    int bi = 0x7fffffff;
    for (int i = lane; i < n; i += 32) {
      if (!((used >> i) & 1ull)) {
        const double v = fabs(AUG[i * kAug + k]);
        if (v > best) {
          best = v;
          bi = i;
        }
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, off);
      const int oi = __shfl_xor_sync(FULL, bi, off);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    int p = bi;
    if (p >= n) {  // only NaNs left in the column: take the first unused row (the solve is lost anyway)
      p = 0;
      while ((used >> p) & 1ull) ++p;
    }
    used |= 1ull << p;
    if (warp == 0 && lane == 0) perm[k] = p;
    const double inv = 1.0 / AUG[p * kAug + k];
    // live columns k+1 .. ncol-1 ; lane owns two of them, pivot-row values stay in registers
    const int ca = k + 1 + lane, cb_ = k + 33 + lane;
    const bool va = ca < ncol, vb = cb_ < ncol;
    const double pa = va ? AUG[p * kAug + ca] : 0.0;
    const double pb = vb ? AUG[p * kAug + cb_] : 0.0;
#pragma unroll 3
    for (int i = warp; i < n; i += kElimWarps) {
      if (i == p) continue;
      const double f = AUG[i * kAug + k] * inv;
      if (va) AUG[i * kAug + ca] = fma(-f, pa, AUG[i * kAug + ca]);
      if (vb) AUG[i * kAug + cb_] = fma(-f, pb, AUG[i * kAug + cb_]);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kElimThreads) dls_eliminate_kernel(const double* __restrict__ coef_in,
                                                                    const int* __restrict__ status, int count,
                                                                    double* __restrict__ S_out) {
  extern __shared__ double sm[];
  double* N = sm;                       // [60][27]
  double* AUG = N + kNRows * kN;        // [36][55]
  double* coef = AUG + 36 * kAug;       // [60]
  double* part = coef + 60;             // [4][3][27]
  int* perm = reinterpret_cast<int*>(part + kElimWarps * 3 * kN);  // [36]
  const int t = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (t >= count || status[t] != 0) return;  // CTA-uniform
  if (tid < 60) coef[tid] = coef_in[(size_t)t * 60 + tid];
  __syncthreads();

  // ---------------- blocks of degree 3..6 ----------------
  for (int blk = 0; blk < 4; ++blk) {
    const int o0 = dls::kBlockOff[blk], n = dls::kBlockOff[blk + 1] - o0;
    const int ncol = n + kN;
    for (int r = warp; r < n; r += kElimWarps) {
      const int row = o0 + r;
      const int pi = dls::kRowPoly[row];
      double* arow = AUG + r * kAug;
      arow[lane] = 0.0;
      if (lane + 32 < kAug) arow[lane + 32] = 0.0;
      double c = 0.0;
      int cd = 0;
      if (lane < 20) {
        c = coef[pi * 20 + lane];
        cd = dls::kRowTerms[row][lane];
      }
      const double acc = row_lower_part(N, c, cd, o0, lane);
      __syncwarp();
      if (lane < 20) {
        if (cd < kN) arow[n + cd] = -c;                       // reduced monomial -> right-hand side
        else if (cd - kN >= o0) arow[cd - kN - o0] = c;       // same-degree non-reduced -> D
      }
      __syncwarp();
      if (lane < kN) arow[n + lane] += acc;
    }
    __syncthreads();
    block_gauss_jordan(AUG, perm, n, ncol, warp, lane);
    // normal forms: N[o0 + k][b] = R[perm[k]][b] / pivot_k
    for (int k = warp; k < n; k += kElimWarps) {
      const int p = perm[k];
      const double inv = 1.0 / AUG[p * kAug + k];
      if (lane < kN) N[(o0 + k) * kN + lane] = AUG[p * kAug + n + lane] * inv;
    }
    __syncthreads();
  }

  // ---------------- degree 7: three rows of D^-1 R via D^T Y = E_J ----------------
  {
    const int o0 = dls::kBlockOff[4], n = 36, ncol = 39;
    for (int i = tid; i < 36 * kAug; i += kElimThreads) AUG[i] = 0.0;
    __syncthreads();
    for (int r = warp; r < n; r += kElimWarps) {  // row r of D -> column r of AUG
      const int row = o0 + r;
      const int pi = dls::kRowPoly[row];
      if (lane < 20) {
        const int cd = dls::kRowTerms[row][lane];
        if (cd >= kN && cd - kN >= o0) AUG[(cd - kN - o0) * kAug + r] = coef[pi * 20 + lane];
      }
    }
    if (tid < 3) AUG[dls::kBorder7[tid] * kAug + 36 + tid] = 1.0;
    __syncthreads();
    block_gauss_jordan(AUG, perm, n, ncol, warp, lane);
    // contract: N7[j][b] = sum_r Y[r][j] * R[r][b],  Y[r][j] = AUG[perm[r]][36+j] / AUG[perm[r]][r]
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int r = warp; r < n; r += kElimWarps) {
      const int row = o0 + r;
      const int pi = dls::kRowPoly[row];
      double c = 0.0;
      int cd = 0;
      if (lane < 20) {
        c = coef[pi * 20 + lane];
        cd = dls::kRowTerms[row][lane];
      }
      double rr = row_lower_part(N, c, cd, o0, lane);
      // reduced terms of this row: lane b picks up -c of the term whose code is b
      const unsigned red = __ballot_sync(FULL, lane < 20 && cd < kN);
      unsigned m = red;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const double cc = __shfl_sync(FULL, c, b);
        const int code = __shfl_sync(FULL, cd, b);
        if (lane == code) rr -= cc;
      }
      const int p = perm[r];
      const double inv = 1.0 / AUG[p * kAug + r];
      const double y0 = AUG[p * kAug + 36] * inv, y1 = AUG[p * kAug + 37] * inv, y2 = AUG[p * kAug + 38] * inv;
      a0 = fma(y0, rr, a0);
      a1 = fma(y1, rr, a1);
      a2 = fma(y2, rr, a2);
    }
    if (lane < kN) {
      part[(warp * 3 + 0) * kN + lane] = a0;
      part[(warp * 3 + 1) * kN + lane] = a1;
      part[(warp * 3 + 2) * kN + lane] = a2;
    }
    __syncthreads();
    if (warp < 3 && lane < kN) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kElimWarps; ++w) v += part[(w * 3 + warp) * kN + lane];
      N[(57 + warp) * kN + lane] = v;
    }
    __syncthreads();
  }

  // ---------------- action matrix of f0: S[b][:] = sum_k F0[k] * (e_cd  or  N[cd-27][:]) ----------------
  double* So = S_out + (size_t)t * kN * kN;
  for (int b = warp; b < kN; b += kElimWarps) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cd = dls::kF0Terms[b][k];
      if (cd < kN) {
        if (lane == cd) v += dls::kF0[k];
      } else if (lane < kN) {
        v += dls::kF0[k] * N[(cd - kN) * kN + lane];
      }
    }
    if (lane < kN) So[b * kN + lane] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// stage 2, second version (the default; CB_PNP_ELIM_V1=1 keeps the kernel above): ONE WARP per hypothesis and the
// Gauss-Jordan solves run on REGISTERS.
// The first version was issue-bound on shared-memory traffic and redundancy (ncu, profiles/r2b: 126 k warp instructions
// per hypothesis -- 3 LDS/STS per FMA, the pivot search repeated by all four warps, a CTA barrier per pivot).  Here a lane
// owns one matrix ROW in registers (blocks of 3..27 rows; the 36-row block gives lanes 0..3 a second row), so a pivot
// step is: arg-max over the lanes with two REDUX + one ballot (lowest row wins ties, as before), the pivot lane
// publishes its row through a 448-byte shared buffer (STS.128 / broadcast LDS.128), and every lane updates its own row
// with (columns - k) independent DFMAs from registers.  Same pivots, same operations in the same order per element: the
// action matrix is bit-identical to the first version's (tests/test_pnp_gpu.py compares them).
// Shared memory per warp: N[60][27] | AUG (rows are still assembled lane = column, then read back lane = row) | the
// pivot-row double buffer (later Y of the degree-7 block) | coef[60]  = 26 224 bytes -> 8 warps per SM.
// ---------------------------------------------------------------------------------------------
constexpr int kE2Warps = 4;
constexpr int kE2Stride7 = 41;                       // AUG row stride of the 36 x 39 block (odd)
constexpr int kE2AugDoubles = 1486;                  // >= 27 * kAug = 1485 and 36 * 41 = 1476, even
constexpr int kE2RowBuf = 56;                        // one published pivot row (<= 54 columns), 16-byte aligned
constexpr int kE2PerWarp = kNRows * kN + kE2AugDoubles + 2 * kE2RowBuf + 60;  // doubles
constexpr int kE2SmemBytes = kE2Warps * kE2PerWarp * 8;
static_assert(kE2PerWarp % 2 == 0 && (kNRows * kN) % 2 == 0, "16-byte alignment of the per-warp regions");

// IEEE reciprocal, deliberately NOT inlined: the inline expansion is ~45 instructions per use, and straight-line code size
// -- not arithmetic -- was the limit of this kernel's first register-resident version (45 k SASS instructions executed once
// per hypothesis: ncu showed 32 % of the warp stalls on instruction fetch)
__device__ __noinline__ double dls_recip(double x) { return 1.0 / x; }

// Gauss-Jordan on the first NR columns of an NR x NC system, lane = row (lanes >= NR hold zeros).  The pivot loop is a
// real loop: after step k column k is dead, so every step shifts the row one column to the left while it eliminates
// (a[c] = a[c+1] - f * pivot_row[c+1]) and the pivot column is always a[0] -- register indices stay compile-time constants
// without unrolling the NR steps.  On return the right-hand sides sit in a[0 .. NC-NR) and the lane with myk == k holds
// unknown k, still scaled by its pivot mypiv.  Same pivots and the same operations per live element as the first version.
// (No __restrict__ on the shared-memory pointers of this kernel: they carry data BETWEEN lanes, and a restrict-qualified
// pointer lets the compiler keep a value it loaded before a __syncwarp() -- the pivot row of two steps earlier.)
template <int NR, int NC>
__device__ __forceinline__ void gauss_jordan_lanes(double (&a)[NC], double* rowbuf, int lane, int& myk, double& mypiv) {
  static_assert(NC % 2 == 0 || NC + 1 <= kE2RowBuf, "row buffer too small");
  unsigned used = 0u;
  myk = -1;
  mypiv = 1.0;
#pragma unroll 1
  for (int k = 0; k < NR; ++k) {
    const double x = a[0];
    const bool cand = lane < NR && !((used >> lane) & 1u) && (x == x);
    const int hi = cand ? (__double2hiint(x) & 0x7fffffff) : -1;
    const int mhi = __reduce_max_sync(FULL, hi);
    int p;
    if (mhi < 0) {  // only NaNs left in the column: first unused row (the solve is lost anyway)
      p = __ffs(~used & ((NR < 32) ? ((1u << NR) - 1u) : 0xffffffffu)) - 1;
    } else {
      const bool top = hi == mhi;
      const unsigned lo = top ? (unsigned)__double2loint(x) : 0u;
      const unsigned mlo = __reduce_max_sync(FULL, lo);
      p = __ffs(__ballot_sync(FULL, top && lo == mlo)) - 1;
    }
    used |= 1u << p;
    double* pr = rowbuf + (k & 1) * kE2RowBuf;
    if (lane == p) {
      myk = k;
      mypiv = x;
#pragma unroll
      for (int c = 0; c < NC; c += 2) {
        if (c + 1 < NC)
          *reinterpret_cast<double2*>(pr + c) = make_double2(a[c], a[c + 1]);
        else
          pr[c] = a[c];
      }
    }
    __syncwarp();
    const double inv = dls_recip(pr[0]);
    const double f = (lane == p || lane >= NR) ? 0.0 : x * inv;
#pragma unroll
    for (int c = 0; c < NC; c += 2) {  // aligned pairs of the pivot row; column c+1 lands in column c
      if (c + 1 < NC) {
        const double2 pv = *reinterpret_cast<const double2*>(pr + c);
        if (c >= 2) a[c - 1] = fma(-f, pv.x, a[c]);
        a[c] = fma(-f, pv.y, a[c + 1]);
      } else {
        a[c - 1] = fma(-f, pr[c], a[c]);
      }
    }
    a[NC - 1] = 0.0;
  }
}

constexpr int kOffC[6] = {0, 3, 12, 30, 57, 93};  // dls::kBlockOff as compile-time constants

// Row descriptions for lane = row access: per template row 24 bytes = its 20 term codes, the polynomial index, 3 pad.
// (The __constant__ tables of dls_tables.h serialise when every lane asks for a different row; this copy lives in global
// memory / L1 and is filled once per handle by dls_pack_rows_kernel.)
__device__ unsigned int g_dls_rows[dls::kNonRed * 6];

__global__ void dls_pack_rows_kernel() {
  const int r = threadIdx.x;
  if (r >= dls::kNonRed) return;
  unsigned char b[24];
  for (int t = 0; t < 20; ++t) b[t] = dls::kRowTerms[r][t];
  b[20] = dls::kRowPoly[r];
  b[21] = b[22] = b[23] = 0;
  for (int i = 0; i < 6; ++i)
    g_dls_rows[r * 6 + i] = (unsigned)b[4 * i] | ((unsigned)b[4 * i + 1] << 8) | ((unsigned)b[4 * i + 2] << 16) | ((unsigned)b[4 * i + 3] << 24);
}

// One template row per LANE: the lane walks the 20 terms of row `row`; terms on a same-degree non-reduced monomial or on
// a reduced monomial are scattered into the lane's own shared-memory row `arow` ([D (n) | R (27)], zeroed here), the
// lower-degree terms are folded into 27 register accumulators: racc[b] -= c * N[cd - 27][b], in ascending term order (the
// order of the first version's ballot loop, so the sums round identically).
template <int O0, int NCOLS_D>
__device__ __forceinline__ void lane_row_assemble(bool active, int row, const double* coef, const double* N, double* arow,
                                                  double (&racc)[kN]) {
  const unsigned char* rb = reinterpret_cast<const unsigned char*>(g_dls_rows) + (active ? row : 0) * 24;
  const int pi = (int)__ldg(rb + 20);
#pragma unroll
  for (int b = 0; b < kN; ++b) racc[b] = 0.0;
  if (active) {
#pragma unroll
    for (int c = 0; c < NCOLS_D + kN; ++c) arow[c] = 0.0;
  }
#pragma unroll 1
  for (int t = 0; t < 20; ++t) {
    const int cd = (int)__ldg(rb + t);
    const double c = coef[pi * 20 + t];
    const bool lower = active && cd >= kN && (cd - kN) < O0;
    if (active) {
      if (cd < kN) arow[NCOLS_D + cd] = -c;
      else if (cd - kN >= O0) arow[cd - kN - O0] = c;
    }
    if (O0 > 0 && __any_sync(FULL, lower)) {  // warp-uniform
      const double* Nb = N + (lower ? (cd - kN) * kN : 0);
#pragma unroll
      for (int b = 0; b < kN; ++b)
        if (lower) racc[b] = fma(-c, Nb[b], racc[b]);
    }
  }
}

template <int BLK>
__device__ __forceinline__ void elim2_block(double* N, double* AUG, double* rowbuf, const double* coef, int lane) {
  constexpr int o0 = kOffC[BLK], n = kOffC[BLK + 1] - kOffC[BLK], NC = n + kN;
  // every lane assembles ITS row [D | R]: scatter through its own shared-memory row (a term's column is a run-time index),
  // lower-degree terms through registers; nothing crosses lanes, so no barrier
  const bool active = lane < n;
  double* arow = AUG + lane * kAug;
  double a[NC];
  {
    double racc[kN];
    lane_row_assemble<o0, n>(active, o0 + lane, coef, N, arow, racc);
#pragma unroll
    for (int c = 0; c < NC; ++c) a[c] = active ? arow[c] : 0.0;
#pragma unroll
    for (int b = 0; b < kN; ++b) a[n + b] += racc[b];  // (direct term) + (lower-degree part), as the first version adds them
  }
  int myk;
  double mypiv;
  gauss_jordan_lanes<n, NC>(a, rowbuf, lane, myk, mypiv);
  if (lane < n) {  // normal forms: N[o0 + k][b] = R[row of unknown k][b] / pivot_k
    const double inv = 1.0 / mypiv;
    double* Nr = N + (o0 + myk) * kN;
#pragma unroll
    for (int b = 0; b < kN; ++b) Nr[b] = a[b] * inv;  // the solve shifted the right-hand sides down to a[0 .. 26]
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32 * kE2Warps, 2) dls_eliminate2_kernel(const double* __restrict__ coef_in,
                                                                         const int* __restrict__ status, int count,
                                                                         double* __restrict__ S_out) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * kE2Warps + warp;
  if (t >= count || status[t] != 0) return;  // warp-uniform, and the kernel has no CTA-wide barrier
  double* N = sm + (size_t)warp * kE2PerWarp;  // [60][27]
  double* AUG = N + kNRows * kN;
  double* rowbuf = AUG + kE2AugDoubles;
  double* coef = rowbuf + 2 * kE2RowBuf;
  coef[lane] = coef_in[(size_t)t * 60 + lane];
  if (lane + 32 < 60) coef[lane + 32] = coef_in[(size_t)t * 60 + lane + 32];
  __syncwarp();

  elim2_block<0>(N, AUG, rowbuf, coef, lane);
  elim2_block<1>(N, AUG, rowbuf, coef, lane);
  elim2_block<2>(N, AUG, rowbuf, coef, lane);
  elim2_block<3>(N, AUG, rowbuf, coef, lane);

  // ---------------- degree 7: three rows of D^-1 R via D^T Y = E_J (36 x 39; lanes 0..3 own rows 32..35 as well) ----
  {
    constexpr int o0 = kOffC[4], n = 36, NC = 39, ST = kE2Stride7;
    for (int i = lane; i < n * ST; i += 32) AUG[i] = 0.0;
    __syncwarp();
    for (int r = 0; r < n; ++r) {  // row r of D -> column r of AUG
      const int row = o0 + r;
      const int pi = dls::kRowPoly[row];
      if (lane < 20) {
        const int cd = dls::kRowTerms[row][lane];
        if (cd >= kN && cd - kN >= o0) AUG[(cd - kN - o0) * ST + r] = coef[pi * 20 + lane];
      }
    }
    if (lane < 3) AUG[dls::kBorder7[lane] * ST + 36 + lane] = 1.0;
    __syncwarp();
    double aA[NC], aB[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      aA[c] = AUG[lane * ST + c];
      aB[c] = lane < 4 ? AUG[(32 + lane) * ST + c] : 0.0;
    }
    unsigned usedA = 0u, usedB = 0u;
    int mykA = -1, mykB = -1;
    double mypivA = 1.0, mypivB = 1.0;
#pragma unroll 1
    for (int k = 0; k < n; ++k) {  // same shifting loop as gauss_jordan_lanes, two register rows per lane
      const double xa = aA[0], xb = aB[0];
      const bool candA = !((usedA >> lane) & 1u) && (xa == xa);
      const bool candB = lane < 4 && !((usedB >> lane) & 1u) && (xb == xb);
      const int hiA = candA ? (__double2hiint(xa) & 0x7fffffff) : -1;
      const int hiB = candB ? (__double2hiint(xb) & 0x7fffffff) : -1;
      const int mhi = __reduce_max_sync(FULL, hiA > hiB ? hiA : hiB);
      int p;  // row 0..35
      if (mhi < 0) {
        p = (~usedA) ? __ffs(~usedA) - 1 : 32 + __ffs(~usedB & 0xfu) - 1;
      } else {
        const bool topA = hiA == mhi, topB = hiB == mhi;
        const unsigned loA = topA ? (unsigned)__double2loint(xa) : 0u;
        const unsigned loB = topB ? (unsigned)__double2loint(xb) : 0u;
        const unsigned mlo = __reduce_max_sync(FULL, loA > loB ? loA : loB);
        const unsigned bA = __ballot_sync(FULL, topA && loA == mlo);
        const unsigned bB = __ballot_sync(FULL, topB && loB == mlo);
        p = bA ? __ffs(bA) - 1 : 32 + __ffs(bB) - 1;  // lowest row index wins ties
      }
      if (p < 32) usedA |= 1u << p;
      else usedB |= 1u << (p - 32);
      double* pr = rowbuf + (k & 1) * kE2RowBuf;
      if (lane == p) {
        mykA = k;
        mypivA = xa;
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
          if (c + 1 < NC) *reinterpret_cast<double2*>(pr + c) = make_double2(aA[c], aA[c + 1]);
          else pr[c] = aA[c];
        }
      }
      if (lane + 32 == p) {
        mykB = k;
        mypivB = xb;
#pragma unroll
        for (int c = 0; c < NC; c += 2) {
          if (c + 1 < NC) *reinterpret_cast<double2*>(pr + c) = make_double2(aB[c], aB[c + 1]);
          else pr[c] = aB[c];
        }
      }
      __syncwarp();
      const double inv = dls_recip(pr[0]);
      const double fA = (lane == p) ? 0.0 : xa * inv;
      const double fB = (lane >= 4 || lane + 32 == p) ? 0.0 : xb * inv;
#pragma unroll
      for (int c = 0; c < NC; c += 2) {
        if (c + 1 < NC) {
          const double2 pv = *reinterpret_cast<const double2*>(pr + c);
          if (c >= 2) {
            aA[c - 1] = fma(-fA, pv.x, aA[c]);
            aB[c - 1] = fma(-fB, pv.x, aB[c]);
          }
          aA[c] = fma(-fA, pv.y, aA[c + 1]);
          aB[c] = fma(-fB, pv.y, aB[c + 1]);
        } else {
          aA[c - 1] = fma(-fA, pr[c], aA[c]);
          aB[c - 1] = fma(-fB, pr[c], aB[c]);
        }
      }
      aA[NC - 1] = 0.0;
      aB[NC - 1] = 0.0;
    }
    // Y[r][j] = (row of unknown r)[36 + j] / pivot_r  -> shared (aliases the pivot-row buffers: wait for their readers)
    __syncwarp();
    double* Y = rowbuf;  // [36][3]
    {
      const double inv = 1.0 / mypivA;
      Y[mykA * 3 + 0] = aA[0] * inv;  // the solve shifted the three right-hand sides down to columns 0..2
      Y[mykA * 3 + 1] = aA[1] * inv;
      Y[mykA * 3 + 2] = aA[2] * inv;
    }
    if (lane < 4) {
      const double inv = 1.0 / mypivB;
      Y[mykB * 3 + 0] = aB[0] * inv;
      Y[mykB * 3 + 1] = aB[1] * inv;
      Y[mykB * 3 + 2] = aB[2] * inv;
    }
    __syncwarp();
    // right-hand side rows of the 36 degree-7 template rows, one per lane (two passes: rows 0..31, then 32..35), into
    // shared memory (AUG is free again); then lane = column b contracts N7[j][b] = sum_r Y[r][j] * R[r][b], four partial
    // sums over r mod 4 added in the first version's order
    double* Rsm = AUG;  // [36][27]
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int r = pass * 32 + lane;
      const bool active = r < n;
      double racc[kN];
      double* rrow = Rsm + (active ? r : 0) * kN;
      {
        const unsigned char* rb = reinterpret_cast<const unsigned char*>(g_dls_rows) + (o0 + (active ? r : 0)) * 24;
        const int pi = (int)__ldg(rb + 20);
#pragma unroll
        for (int b = 0; b < kN; ++b) racc[b] = 0.0;
#pragma unroll 1
        for (int t = 0; t < 20; ++t) {
          const int cd = (int)__ldg(rb + t);
          const double c = coef[pi * 20 + t];
          const bool lower = active && cd >= kN && (cd - kN) < o0;
          if (__any_sync(FULL, lower)) {
            const double* Nb = N + (lower ? (cd - kN) * kN : 0);
#pragma unroll
            for (int b = 0; b < kN; ++b)
              if (lower) racc[b] = fma(-c, Nb[b], racc[b]);
          }
        }
        if (active) {
#pragma unroll
          for (int b = 0; b < kN; ++b) rrow[b] = racc[b];
#pragma unroll 1
          for (int t = 0; t < 20; ++t) {
            const int cd = (int)__ldg(rb + t);
            if (cd < kN) rrow[cd] -= coef[pi * 20 + t];  // a reduced monomial appears once per row
          }
        }
      }
    }
    __syncwarp();
    double acc[4][3];
#pragma unroll
    for (int w = 0; w < 4; ++w) acc[w][0] = acc[w][1] = acc[w][2] = 0.0;
    for (int r4 = 0; r4 < n; r4 += 4) {
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const int r = r4 + w;
        const double rr = lane < kN ? Rsm[r * kN + lane] : 0.0;
        acc[w][0] = fma(Y[r * 3 + 0], rr, acc[w][0]);
        acc[w][1] = fma(Y[r * 3 + 1], rr, acc[w][1]);
        acc[w][2] = fma(Y[r * 3 + 2], rr, acc[w][2]);
      }
    }
    if (lane < kN) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < 4; ++w) v += acc[w][j];
        N[(57 + j) * kN + lane] = v;
      }
    }
    __syncwarp();
  }

  // ---------------- action matrix of f0 ----------------
  double* So = S_out + (size_t)t * kN * kN;
  for (int b = 0; b < kN; ++b) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cd = dls::kF0Terms[b][k];
      if (cd < kN) {
        if (lane == cd) v += dls::kF0[k];
      } else if (lane < kN) {
        v += dls::kF0[k] * N[(cd - kN) * kN + lane];
      }
    }
    if (lane < kN) So[b * kN + lane] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// stage 3: eigenvalues + roots + cheirality (warp per hypothesis)
// ---------------------------------------------------------------------------------------------
constexpr int kRootsWarps = 4;
constexpr int kRootsSmemPerWarp = kN * kN + 64;  // one 27x27 work matrix (H, then LU) | wr/wi/v scratch (doubles)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL, v, off);
  return v;
}

// Householder reduction to upper Hessenberg form, in place (no accumulation).
__device__ void warp_hessenberg(double* H, double* vs, int lane) {
  const int n = kN;
  for (int k = 0; k < n - 2; ++k) {
    const bool act = lane > k && lane < n;
    double xi = act ? H[lane * n + k] : 0.0;
    double alpha = sqrt(warp_sum(xi * xi));
    if (alpha == 0.0) continue;
    const double x0 = H[(k + 1) * n + k];
    if (x0 > 0.0) alpha = -alpha;
    double vi = xi;
    if (lane == k + 1) vi -= alpha;
    const double vn2 = warp_sum(vi * vi);
    if (vn2 == 0.0) continue;
    const double beta = 2.0 / vn2;
    vs[lane] = vi;
    __syncwarp();
    // left: H[k+1:, j] -= beta * v * (v^T H[k+1:, j])   for j >= k ; lane = column j
    if (lane >= k && lane < n) {
      double dot = 0.0;
      for (int i = k + 1; i < n; ++i) dot += vs[i] * H[i * n + lane];
      dot *= beta;
      for (int i = k + 1; i < n; ++i) H[i * n + lane] -= vs[i] * dot;
    }
    __syncwarp();
    // right: H[:, k+1:] -= beta * (H[:, k+1:] v) v^T ; lane = row i
    if (lane < n) {
      double dot = 0.0;
      for (int c = k + 1; c < n; ++c) dot += H[lane * n + c] * vs[c];
      dot *= beta;
      for (int c = k + 1; c < n; ++c) H[lane * n + c] -= dot * vs[c];
    }
    __syncwarp();
    if (lane > k + 1 && lane < n) H[lane * n + k] = 0.0;
    __syncwarp();
  }
}

// Francis double-shift QR, eigenvalues only (same control flow as tools/gen_dls_tables.py:hqr_eigenvalues).
// Returns false if an eigenvalue failed to converge.
__device__ bool warp_hqr(double* a, double* wr, double* wi, int lane) {
  const int n = kN;
  const double eps = 2.220446049250313e-16;
  double anorm_p = 0.0;
  if (lane < n) {
    const int j0 = lane > 0 ? lane - 1 : 0;
    for (int j = j0; j < n; ++j) anorm_p += fabs(a[lane * n + j]);
  }
  const double anorm = warp_sum(anorm_p);
  int nn = n - 1;
  double t = 0.0;
  while (nn >= 0) {
    int its = 0;
    while (true) {
      // ---- look for a single small sub-diagonal element: highest l in [1, nn] that satisfies
      bool small = false;
      if (lane >= 1 && lane <= nn) {
        double s = fabs(a[(lane - 1) * n + lane - 1]) + fabs(a[lane * n + lane]);
        if (s == 0.0) s = anorm;
        small = fabs(a[lane * n + lane - 1]) <= eps * s;
      }
      const unsigned sm = __ballot_sync(FULL, small);
      const int l = sm ? 31 - __clz(sm) : 0;
      if (l >= 1 && lane == 0) a[l * n + l - 1] = 0.0;
      __syncwarp();
      double x = a[nn * n + nn];
      if (l == nn) {  // one root found
        if (lane == 0) {
          wr[nn] = x + t;
          wi[nn] = 0.0;
        }
        nn -= 1;
        break;
      }
      double y = a[(nn - 1) * n + nn - 1];
      double w = a[nn * n + nn - 1] * a[(nn - 1) * n + nn];
      if (l == nn - 1) {  // two roots found
        const double p = 0.5 * (y - x);
        const double q = p * p + w;
        double z = sqrt(fabs(q));
        x += t;
        if (lane == 0) {
          if (q >= 0.0) {
            z = p + (p >= 0.0 ? z : -z);
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else {
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn - 1] = z;
            wi[nn] = -z;
          }
        }
        nn -= 2;
        break;
      }
      if (its >= 60) return false;
      if (its == 10 || its == 20) {  // exceptional shift
        t += x;
        if (lane <= nn) a[lane * n + lane] -= x;
        __syncwarp();
        const double s = fabs(a[nn * n + nn - 1]) + fabs(a[(nn - 1) * n + nn - 2]);
        y = x = 0.75 * s;
        w = -0.4375 * s * s;
      }
      ++its;
      // ---- find m: highest m in [l, nn-2] with two consecutive small sub-diagonals (or l)
      double p = 0.0, q = 0.0, r = 0.0;
      bool ok = false;
      if (lane >= l && lane <= nn - 2) {
        const int m = lane;
        const double z = a[m * n + m];
        double rr = x - z, ss = y - z;
        p = (rr * ss - w) / a[(m + 1) * n + m] + a[m * n + m + 1];
        q = a[(m + 1) * n + m + 1] - z - rr - ss;
        r = a[(m + 2) * n + m + 1];
        // the deflation test below is invariant to a common scale of (p,q,r): normalise after selection
        if (m == l) {
          ok = true;
        } else {
          const double u = fabs(a[m * n + m - 1]) * (fabs(q) + fabs(r));
          const double v = fabs(p) * (fabs(a[(m - 1) * n + m - 1]) + fabs(z) + fabs(a[(m + 1) * n + m + 1]));
          ok = u <= eps * v;
        }
      }
      const unsigned om = __ballot_sync(FULL, ok);
      const int m = 31 - __clz(om);  // lane l always sets its bit
      p = __shfl_sync(FULL, p, m);
      q = __shfl_sync(FULL, q, m);
      r = __shfl_sync(FULL, r, m);
      {
        const double inv = 1.0 / (fabs(p) + fabs(q) + fabs(r));
        p *= inv;
        q *= inv;
        r *= inv;
      }
      // clear the entries below the sub-diagonal in the active window
      if (lane >= m + 2 && lane <= nn) {
        a[lane * n + lane - 2] = 0.0;
        if (lane != m + 2) a[lane * n + lane - 3] = 0.0;
      }
      __syncwarp();
      // ---- double QR step on rows l..nn and columns m..nn
      for (int k = m; k <= nn - 1; ++k) {
        if (k != m) {
          p = a[k * n + k - 1];
          q = a[(k + 1) * n + k - 1];
          r = (k != nn - 1) ? a[(k + 2) * n + k - 1] : 0.0;
          // (no pre-scaling by |p|+|q|+|r|: the entries of a balanced 27x27 action matrix are far from the
          //  overflow/underflow range, and the Householder vector is scale invariant)
        }
        const double ss2 = p * p + q * q + r * r;
        if (ss2 != 0.0) {
          // one reciprocal square root gives both the norm and its inverse
          double is = rsqrt(ss2);
          double s = ss2 * is;
          if (p < 0.0) {
            s = -s;
            is = -is;
          }
          __syncwarp();
          if (lane == 0) {
            if (k == m) {
              if (l != m) a[k * n + k - 1] = -a[k * n + k - 1];
            } else {
              a[k * n + k - 1] = -s;
            }
          }
          p += s;
          const double ip = 1.0 / p;
          x = p * is;
          y = q * is;
          const double z = r * is;
          q *= ip;
          r *= ip;
          const bool three = (k != nn - 1);
          // row modification: lane = column j in [k, nn]
          if (lane >= k && lane <= nn) {
            double pp = a[k * n + lane] + q * a[(k + 1) * n + lane];
            if (three) {
              pp += r * a[(k + 2) * n + lane];
              a[(k + 2) * n + lane] -= pp * z;
            }
            a[(k + 1) * n + lane] -= pp * y;
            a[k * n + lane] -= pp * x;
          }
          __syncwarp();
          // column modification: lane = row i in [l, min(nn, k+3)]
          const int mmin = nn < k + 3 ? nn : k + 3;
          if (lane >= l && lane <= mmin) {
            double pp = x * a[lane * n + k] + y * a[lane * n + k + 1];
            if (three) {
              pp += z * a[lane * n + k + 2];
              a[lane * n + k + 2] -= pp * r;
            }
            a[lane * n + k + 1] -= pp * q;
            a[lane * n + k] -= pp;
          }
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();
  return true;
}

// One step of inverse iteration: LU (partial pivoting) of B = A - lam I in `B`, tiny pivots
// replaced, then U x = ones.  Result: lane i holds x[i] scaled to max |x| = 1.
__device__ double warp_null_vector(const double* A, double* B, double lam, int lane) {
  const int n = kN;
  double rowsum = 0.0;
  if (lane < n) {
    for (int c = 0; c < n; ++c) {
      const double v = A[lane * n + c];
      rowsum += fabs(v);
      B[lane * n + c] = (c == lane) ? v - lam : v;
    }
  }
  double nrm = rowsum;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) nrm = fmax(nrm, __shfl_xor_sync(FULL, nrm, off));
  const double tiny = 2.220446049250313e-16 * nrm;
  __syncwarp();
  for (int k = 0; k < n; ++k) {
    double best = (lane >= k && lane < n) ? fabs(B[lane * n + k]) : -1.0;
    int bi = lane;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, off);
      const int oi = __shfl_xor_sync(FULL, bi, off);
      if (ob > best || (ob == best && oi < bi)) {
        best = ob;
        bi = oi;
      }
    }
    if (bi != k) {
      if (lane >= k && lane < n) {
        const double x = B[k * n + lane], y = B[bi * n + lane];
        B[k * n + lane] = y;
        B[bi * n + lane] = x;
      }
      __syncwarp();
    }
    double piv = B[k * n + k];
    if (fabs(piv) < tiny) {
      piv = tiny;
      __syncwarp();
      if (lane == 0) B[k * n + k] = tiny;
    }
    __syncwarp();
    // lane = row i > k
    const double ipiv = 1.0 / piv;
    if (lane > k && lane < n) {
      const double f = B[lane * n + k] * ipiv;
      for (int c = k + 1; c < n; ++c) B[lane * n + c] = fma(-f, B[k * n + c], B[lane * n + c]);
    }
    __syncwarp();
  }
  // back substitution, column oriented: lane i holds rhs_i
  double rhs = 1.0, xk_mine = 0.0;
  for (int k = n - 1; k >= 0; --k) {
    const double xk = __shfl_sync(FULL, rhs, k) / B[k * n + k];
    if (lane == k) xk_mine = xk;
    if (lane < k) rhs = fma(-B[lane * n + k], xk, rhs);
  }
  double mx = lane < n ? fabs(xk_mine) : 0.0;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, off));
  __syncwarp();
  return lane < n ? xk_mine / mx : 0.0;
}

struct RootsArgs {
  const double* S;      // [count][27*27]
  const double* T;      // [count][27]
  const int* idx;       // [count][15]
  const double* X;      // [total][3]
  int* status;          // in: 0 ok ; out: number of solutions (>=0) or -1
  double* model;        // [count][max_sol][12]  R row-major (9) + t (3)
  int count;
  int max_sol;          // solutions stored per hypothesis (1 in the RANSAC path)
};

__global__ void __launch_bounds__(32 * kRootsWarps) dls_roots_kernel(RootsArgs a) {
  extern __shared__ double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * kRootsWarps + warp;
  if (t >= a.count) return;
  if (a.status[t] != 0) return;
  double* H = sm + (size_t)warp * kRootsSmemPerWarp;  // Hessenberg/QR work matrix, reused for the LU
  double* LU = H;
  double* wr = H + kN * kN;  // [27] (+ wi [27] share the 64-double scratch: wr at 0, wi at 32)
  double* wi = wr + 32;
  const double* Sg = a.S + (size_t)t * kN * kN;
  const double* A = Sg;  // the action matrix itself stays in global memory (L2-resident)
  for (int i = lane; i < kN * kN; i += 32) H[i] = Sg[i];
  __syncwarp();
  warp_hessenberg(H, wr /*scratch for v*/, lane);
  const bool conv = warp_hqr(H, wr, wi, lane);
  __syncwarp();
  if (!conv) {
    if (lane == 0) a.status[t] = 0;  // no usable model (counts as 0 solutions)
    return;
  }
  // the 15 sample points: lane i < 15 keeps X_i
  double px = 0.0, py = 0.0, pz = 0.0;
  if (lane < kSample) {
    const size_t id = (size_t)a.idx[(size_t)t * kSample + lane];
    px = a.X[3 * id];
    py = a.X[3 * id + 1];
    pz = a.X[3 * id + 2];
  }
  const double* Tg = a.T + (size_t)t * 27;
  int nsol = 0;
  for (int e = 0; e < kN; ++e) {
    if (wi[e] != 0.0) continue;  // uniform (shared memory broadcast)
    const double lam = wr[e];
    const double xv = warp_null_vector(A, LU, lam, lane);
    const double v0 = __shfl_sync(FULL, xv, dls::kIdxOne);
    const double s1 = __shfl_sync(FULL, xv, dls::kIdxS1) / v0;
    const double s2 = __shfl_sync(FULL, xv, dls::kIdxS2) / v0;
    const double s3 = __shfl_sync(FULL, xv, dls::kIdxS3) / v0;
    if (!(isfinite(s1) && isfinite(s2) && isfinite(s3))) continue;
    // Cayley -> rotation (oracle/dls_pnp.py:cayley_to_rotation)
    const double ss = s1 * s1 + s2 * s2 + s3 * s3;
    const double k = 1.0 / (1.0 + ss);
    double R[9];
    R[0] = (1.0 + s1 * s1 - s2 * s2 - s3 * s3) * k;
    R[1] = (2.0 * s1 * s2 - 2.0 * s3) * k;
    R[2] = (2.0 * s1 * s3 + 2.0 * s2) * k;
    R[3] = (2.0 * s1 * s2 + 2.0 * s3) * k;
    R[4] = (1.0 - s1 * s1 + s2 * s2 - s3 * s3) * k;
    R[5] = (2.0 * s2 * s3 - 2.0 * s1) * k;
    R[6] = (2.0 * s1 * s3 - 2.0 * s2) * k;
    R[7] = (2.0 * s2 * s3 + 2.0 * s1) * k;
    R[8] = (1.0 - s1 * s1 - s2 * s2 + s3 * s3) * k;
    // t = T vec(C), vec column-major: index 3*col + row
    double tv[3];
    for (int i = 0; i < 3; ++i) {
      double s = 0.0;
      for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) s += Tg[i * 9 + 3 * c + r] * R[r * 3 + c];
      tv[i] = s;
    }
    // cheirality on the sample (theia::DlsPnp drops solutions with a point behind the camera)
    const double z = R[6] * px + R[7] * py + R[8] * pz + tv[2];
    const bool bad = lane < kSample && !(z >= 0.0);
    if (__ballot_sync(FULL, bad)) continue;
    if (nsol < a.max_sol && lane == 0) {
      double* mo = a.model + ((size_t)t * a.max_sol + nsol) * 12;
      for (int i = 0; i < 9; ++i) mo[i] = R[i];
      mo[9] = tv[0];
      mo[10] = tv[1];
      mo[11] = tv[2];
    }
    ++nsol;
  }
  if (lane == 0) a.status[t] = nsol;
}

// ---------------------------------------------------------------------------------------------
// stage 4: scoring (warp per hypothesis): DlsPnpWithRansac::Error over all points + MLE cost
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pnp_score_kernel(const int* __restrict__ offsets,
                                                       const double* __restrict__ X,
                                                       const double* __restrict__ uv, int H, long long g0,
                                                       int count, const int* __restrict__ status,
                                                       const double* __restrict__ model, double thresh,
                                                       double* __restrict__ cost, int* __restrict__ ninl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= count) return;
  const long long g = g0 + t;
  if (status[t] != 1) {  // model only when DlsPnp returned exactly one solution (DlsPnpWithRansac.h:62)
    if (lane == 0) {
      cost[g] = CUDART_INF;
      ninl[g] = -1;
    }
    return;
  }
  const int cand = (int)(g / H);
  const int p0 = offsets[cand], n = offsets[cand + 1] - p0;
  const double* m = model + (size_t)t * 12;
  const double r0 = m[0], r1 = m[1], r2 = m[2], r3 = m[3], r4 = m[4], r5 = m[5], r6 = m[6], r7 = m[7], r8 = m[8];
  const double t0 = m[9], t1 = m[10], t2 = m[11];
  double c = 0.0;
  int ni = 0;
  for (int i = lane; i < n; i += 32) {
    const size_t id = (size_t)(p0 + i);
    const double x = X[3 * id], y = X[3 * id + 1], z = X[3 * id + 2];
    const double bx = r0 * x + r1 * y + r2 * z + t0;
    const double by = r3 * x + r4 * y + r5 * z + t1;
    const double bz = r6 * x + r7 * y + r8 * z + t2;
    const double e = fabs(bx / bz - uv[2 * id]) + fabs(by / bz - uv[2 * id + 1]);
    if (e < thresh) {  // theia MLEQualityMeasurement: inlier iff residual < error_thresh
      c += e;
      ++ni;
    } else {
      c += thresh;
    }
  }
  c = warp_sum(c);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) ni += __shfl_xor_sync(FULL, ni, off);
  if (lane == 0) {
    cost[g] = c;
    ninl[g] = ni;
  }
}

// ---------------------------------------------------------------------------------------------
// stage 5: per-candidate selection = replay of theia::Ransac::Estimate over the hypotheses
// ---------------------------------------------------------------------------------------------
struct SelectArgs {
  const int* offsets;
  int n_cand, H;
  const double* cost;   // [n_cand*H]
  const int* ninl;      // [n_cand*H]  (-1: hypothesis produced no model)
  const double* model;  // [n_cand*H][12]
  cb_ransac_params p;
  int sample_size;      // 15 (DLS-PnP) or 10 (Umeyama)
  double* c_T_w;        // [n_cand][16]
  float* confidence;
  int* num_iterations;
  int* n_inliers;
  int* best_hyp;
};

__device__ inline int max_iterations_for(double sample_size, double inlier_ratio, double log_fail,
                                         const cb_ransac_params& p) {
  if (inlier_ratio == 1.0) return p.min_iterations;
  const double log_prob = log(1.0 - pow(inlier_ratio, sample_size)) - 2.220446049250313e-16;
  const double num = log_fail / log_prob;
  return (int)fmax((double)p.min_iterations, fmin(num, (double)p.max_iterations));
}

__global__ void __launch_bounds__(32) pnp_select_kernel(SelectArgs a) {
  const int cand = blockIdx.x, lane = threadIdx.x;
  if (cand >= a.n_cand) return;
  const int n = a.offsets[cand + 1] - a.offsets[cand];
  double* To = a.c_T_w + (size_t)cand * 16;
  if (lane < 16) To[lane] = (lane % 5 == 0) ? 1.0 : 0.0;
  if (n < 20) {  // DlsPnpWithRansac.cpp:136-139: return -1
    if (lane == 0) {
      a.confidence[cand] = -1.0f;
      if (a.num_iterations) a.num_iterations[cand] = 0;
      if (a.n_inliers) a.n_inliers[cand] = 0;
      if (a.best_hyp) a.best_hyp[cand] = -1;
    }
    return;
  }
  const double* cost = a.cost + (size_t)cand * a.H;
  const int* ninl = a.ninl + (size_t)cand * a.H;
  int best = -1, iters = 0;
  double best_cost = CUDART_INF;
  if (!a.p.adaptive) {
    // every hypothesis counts: arg-min of cost, first occurrence wins (strict '<' in the reference)
    double bc = CUDART_INF;
    int bh = 0x7fffffff;
    for (int h = lane; h < a.H; h += 32) {
      const double c = cost[h];
      if (ninl[h] >= 0 && (c < bc)) {
        bc = c;
        bh = h;
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const double oc = __shfl_xor_sync(FULL, bc, off);
      const int oh = __shfl_xor_sync(FULL, bh, off);
      if (oc < bc || (oc == bc && oh < bh)) {
        bc = oc;
        bh = oh;
      }
    }
    if (bh != 0x7fffffff) {
      best = bh;
      best_cost = bc;
    }
    iters = a.H;
  } else {
    const double log_fail = log(a.p.failure_probability);
    int max_it = a.p.max_iterations;
    if (a.p.min_inlier_ratio > 0.0) {
      const int m = max_iterations_for((double)a.sample_size, a.p.min_inlier_ratio, log_fail, a.p);
      max_it = m < max_it ? m : max_it;
    }
    int it = 0;
    while (it < max_it) {
      if (it < a.H && ninl[it] >= 0) {
        const double c = cost[it];
        if (c < best_cost) {
          best = it;
          best_cost = c;
          const double ratio = (double)ninl[it] / (double)n;
          if (ratio >= (double)a.sample_size / (double)n) {
            const int m = max_iterations_for((double)a.sample_size, ratio, log_fail, a.p);
            max_it = m < max_it ? m : max_it;
          }
        }
      }
      ++it;
    }
    iters = it;
  }
  if (lane == 0) {
    if (a.num_iterations) a.num_iterations[cand] = iters;
    if (a.best_hyp) a.best_hyp[cand] = best;
    if (best < 0) {
      a.confidence[cand] = 0.0f;
      if (a.n_inliers) a.n_inliers[cand] = 0;
    } else {
      const double* m = a.model + ((size_t)cand * a.H + best) * 12;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) To[i * 4 + j] = m[i * 3 + j];
        To[i * 4 + 3] = m[9 + i];
      }
      const double ratio = (double)ninl[best] / (double)n;
      const double conf = 1.0 - pow(1.0 - pow(ratio, (double)a.sample_size), (double)iters);
      a.confidence[cand] = (float)conf;
      if (a.n_inliers) a.n_inliers[cand] = ninl[best];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Option C: AlignPointCloudsUmeyamaWithRansac (src/DlsPnpWithRansac.h:117-166), thread per hypothesis.
// Sample 10 pairs -> similarity b ~ s R a + t (Umeyama 1991).  R comes from Horn's quaternion form of the
// same optimum (largest eigenvector of the 4x4 symmetric N built from the cross-covariance, cyclic Jacobi),
// which yields a proper rotation in the reflection case exactly like Umeyama's S = diag(1,1,-1);
// s = trace(R^T Sigma) / var(a); the model keeps R and t and is accepted iff min(s, 1/s) > 0.9 (:139).
// ---------------------------------------------------------------------------------------------
constexpr int kIcpSample = 10;

struct IcpArgs {
  const int* offsets;
  const double* A;   // [total][3] points in frame a
  const double* B;   // [total][3] the same points in frame b
  const int* samples;  // optional [n_cand][H][10]
  unsigned long long seed;
  int H;
  long long g0;
  int count;
  int* status;       // out: 1 model, 0 rejected, -1 refused
  double* model;     // [count][12]
};

__device__ inline void jacobi4(double N[4][4], double V[4][4]) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 4; ++i)
      for (int j = i + 1; j < 4; ++j) off += N[i][j] * N[i][j];
    if (off < 1e-300) break;
    for (int pi = 0; pi < 3; ++pi)
      for (int qi = pi + 1; qi < 4; ++qi) {
        const double apq = N[pi][qi];
        if (apq == 0.0) continue;
        const double theta = (N[qi][qi] - N[pi][pi]) / (2.0 * apq);
        const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(tt * tt + 1.0), sn = tt * c;
        for (int k = 0; k < 4; ++k) {  // N <- N J
          const double nkp = N[k][pi], nkq = N[k][qi];
          N[k][pi] = c * nkp - sn * nkq;
          N[k][qi] = sn * nkp + c * nkq;
        }
        for (int k = 0; k < 4; ++k) {  // N <- J^T N
          const double npk = N[pi][k], nqk = N[qi][k];
          N[pi][k] = c * npk - sn * nqk;
          N[qi][k] = sn * npk + c * nqk;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][pi], vkq = V[k][qi];
          V[k][pi] = c * vkp - sn * vkq;
          V[k][qi] = sn * vkp + c * vkq;
        }
      }
  }
}

__global__ void __launch_bounds__(128) icp_model_kernel(IcpArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.count) return;
  const long long g = a.g0 + t;
  const int cand = (int)(g / a.H), hyp = (int)(g % a.H);
  const int p0 = a.offsets[cand], n = a.offsets[cand + 1] - p0;
  if (n < 20) {  // DlsPnpWithRansac.cpp:18-21
    a.status[t] = -1;
    return;
  }
  int idx[kIcpSample];
  if (a.samples) {
    const int* sp = a.samples + ((size_t)cand * a.H + hyp) * kIcpSample;
    for (int i = 0; i < kIcpSample; ++i) idx[i] = sp[i];
  } else {
    sample_indices(a.seed, cand, hyp, n, idx, kIcpSample);
  }
  double ma[3] = {0, 0, 0}, mb[3] = {0, 0, 0};
  for (int i = 0; i < kIcpSample; ++i) {
    if (idx[i] < 0 || idx[i] >= n) {
      a.status[t] = -1;
      return;
    }
    idx[i] += p0;
    for (int c = 0; c < 3; ++c) {
      ma[c] += a.A[3 * (size_t)idx[i] + c];
      mb[c] += a.B[3 * (size_t)idx[i] + c];
    }
  }
  for (int c = 0; c < 3; ++c) {
    ma[c] /= kIcpSample;
    mb[c] /= kIcpSample;
  }
  double Sg[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // Sigma = (1/n) sum (b - mb)(a - ma)^T
  double var_a = 0.0;
  for (int i = 0; i < kIcpSample; ++i) {
    double da[3], db[3];
    for (int c = 0; c < 3; ++c) {
      da[c] = a.A[3 * (size_t)idx[i] + c] - ma[c];
      db[c] = a.B[3 * (size_t)idx[i] + c] - mb[c];
      var_a += da[c] * da[c];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) Sg[r][c] += db[r] * da[c];
  }
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Sg[r][c] /= kIcpSample;
  var_a /= kIcpSample;
  // Horn: S_xy = sum a_x b_y = Sigma^T entries
  const double Sxx = Sg[0][0], Sxy = Sg[1][0], Sxz = Sg[2][0];
  const double Syx = Sg[0][1], Syy = Sg[1][1], Syz = Sg[2][1];
  const double Szx = Sg[0][2], Szy = Sg[1][2], Szz = Sg[2][2];
  double N[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                    {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                    {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                    {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
  double V[4][4];
  jacobi4(N, V);
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (N[i][i] > N[best][best]) best = i;
  double qw = V[0][best], qx = V[1][best], qy = V[2][best], qz = V[3][best];
  const double qn = 1.0 / sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  qw *= qn, qx *= qn, qy *= qn, qz *= qn;
  double R[9];
  R[0] = 1 - 2 * (qy * qy + qz * qz);
  R[1] = 2 * (qx * qy - qw * qz);
  R[2] = 2 * (qx * qz + qw * qy);
  R[3] = 2 * (qx * qy + qw * qz);
  R[4] = 1 - 2 * (qx * qx + qz * qz);
  R[5] = 2 * (qy * qz - qw * qx);
  R[6] = 2 * (qx * qz - qw * qy);
  R[7] = 2 * (qy * qz + qw * qx);
  R[8] = 1 - 2 * (qx * qx + qy * qy);
  double tr = 0.0;  // trace(R^T Sigma)
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) tr += R[r * 3 + c] * Sg[r][c];
  const double sc = tr / var_a;
  double* mo = a.model + (size_t)t * 12;
  for (int i = 0; i < 9; ++i) mo[i] = R[i];
  for (int r = 0; r < 3; ++r) mo[9 + r] = mb[r] - sc * (R[r * 3] * ma[0] + R[r * 3 + 1] * ma[1] + R[r * 3 + 2] * ma[2]);
  const bool okm = isfinite(sc) && sc > 0.0 && fmin(sc, 1.0 / sc) > 0.9;  // DlsPnpWithRansac.h:139
  a.status[t] = okm ? 1 : 0;
}

// residual || R a + t - b || over all pairs (DlsPnpWithRansac.h:152-164), MLE cost, inliers
__global__ void __launch_bounds__(128) icp_score_kernel(const int* __restrict__ offsets, const double* __restrict__ A,
                                                       const double* __restrict__ B, int H, long long g0, int count,
                                                       const int* __restrict__ status, const double* __restrict__ model,
                                                       double thresh, double* __restrict__ cost, int* __restrict__ ninl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 4 + warp;
  if (t >= count) return;
  const long long g = g0 + t;
  if (status[t] != 1) {
    if (lane == 0) {
      cost[g] = CUDART_INF;
      ninl[g] = -1;
    }
    return;
  }
  const int cand = (int)(g / H);
  const int p0 = offsets[cand], n = offsets[cand + 1] - p0;
  const double* m = model + (size_t)t * 12;
  double c = 0.0;
  int ni = 0;
  for (int i = lane; i < n; i += 32) {
    const size_t id = (size_t)(p0 + i);
    const double x = A[3 * id], y = A[3 * id + 1], z = A[3 * id + 2];
    const double ex = m[0] * x + m[1] * y + m[2] * z + m[9] - B[3 * id];
    const double ey = m[3] * x + m[4] * y + m[5] * z + m[10] - B[3 * id + 1];
    const double ez = m[6] * x + m[7] * y + m[8] * z + m[11] - B[3 * id + 2];
    const double e = sqrt(ex * ex + ey * ey + ez * ez);
    if (e < thresh) {
      c += e;
      ++ni;
    } else {
      c += thresh;
    }
  }
  c = warp_sum(c);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) ni += __shfl_xor_sync(FULL, ni, off);
  if (lane == 0) {
    cost[g] = c;
    ninl[g] = ni;
  }
}

__global__ void copy_models_kernel(const double* __restrict__ chunk_model, const int* __restrict__ status,
                                   long long g0, int count, double* __restrict__ all_model) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count * 12) return;
  const int t = i / 12;
  if (status[t] == 1) all_model[(size_t)(g0 + t) * 12 + (i % 12)] = chunk_model[i];
}

}  // namespace

// =============================================================================================
// host side
// =============================================================================================
struct cb_pnp {
  int device = 0, sm_count = 0;
  int max_cand = 0, max_points = 0, max_hyp = 0;
  int chunk = 16384;
  bool elim_v1 = false;  // CB_PNP_ELIM_V1=1: the CTA-per-hypothesis shared-memory elimination (cross-check of the default)
  cudaStream_t stream = nullptr;
  // inputs (host API staging)
  int* offsets = nullptr;
  double* X = nullptr;
  double* uv = nullptr;
  int* samples = nullptr;
  size_t samples_elems = 0;
  // per-hypothesis (all)
  double* cost = nullptr;
  int* ninl = nullptr;
  double* model = nullptr;
  // per-chunk scratch
  int* idx = nullptr;
  int* status = nullptr;
  double* coef = nullptr;
  double* T = nullptr;
  double* S = nullptr;
  double* cmodel = nullptr;  // [chunk][max_sol][12]
  // outputs (host API staging)
  double* out_T = nullptr;
  float* out_conf = nullptr;
  int* out_i = nullptr;  // 3 * max_cand
};

namespace {

int launch_eliminate(cb_pnp* p, int count, cudaStream_t st) {
  if (p->elim_v1) {
    CB_CUDA(cudaFuncSetAttribute(dls_eliminate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kElimSmemBytes));
    dls_eliminate_kernel<<<count, kElimThreads, kElimSmemBytes, st>>>(p->coef, p->status, count, p->S);
  } else {
    CB_CUDA(cudaFuncSetAttribute(dls_eliminate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kE2SmemBytes));
    dls_eliminate2_kernel<<<(count + kE2Warps - 1) / kE2Warps, 32 * kE2Warps, kE2SmemBytes, st>>>(p->coef, p->status, count,
                                                                                                 p->S);
  }
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int run_chunks(cb_pnp* p, int n_cand, const int* offsets_dev, const double* X, const double* uv, int H,
               const cb_ransac_params& prm, const int* samples_dev, cudaStream_t st) {
  const long long total = (long long)n_cand * H;
  const size_t roots_smem = (size_t)kRootsWarps * kRootsSmemPerWarp * sizeof(double);
  CB_CUDA(cudaFuncSetAttribute(dls_roots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roots_smem));
  for (long long g0 = 0; g0 < total; g0 += p->chunk) {
    const int count = (int)((total - g0) < p->chunk ? (total - g0) : p->chunk);
    SetupArgs sa;
    sa.offsets = offsets_dev;
    sa.X = X;
    sa.uv = uv;
    sa.samples = samples_dev;
    sa.seed = prm.seed;
    sa.H = H;
    sa.g0 = g0;
    sa.count = count;
    sa.idx_out = p->idx;
    sa.coef_out = p->coef;
    sa.T_out = p->T;
    sa.status = p->status;
    dls_setup_kernel<<<(count + 127) / 128, 128, 0, st>>>(sa);
    CB_LAUNCH_CHECK();
    if (int rc = launch_eliminate(p, count, st)) return rc;
    RootsArgs ra;
    ra.S = p->S;
    ra.T = p->T;
    ra.idx = p->idx;
    ra.X = X;
    ra.status = p->status;
    ra.model = p->cmodel;
    ra.count = count;
    ra.max_sol = 1;
    dls_roots_kernel<<<(count + kRootsWarps - 1) / kRootsWarps, 32 * kRootsWarps, roots_smem, st>>>(ra);
    CB_LAUNCH_CHECK();
    pnp_score_kernel<<<(count + 3) / 4, 128, 0, st>>>(offsets_dev, X, uv, H, g0, count, p->status, p->cmodel,
                                                      prm.error_thresh, p->cost, p->ninl);
    CB_LAUNCH_CHECK();
    copy_models_kernel<<<(count * 12 + 255) / 256, 256, 0, st>>>(p->cmodel, p->status, g0, count, p->model);
    CB_LAUNCH_CHECK();
  }
  return CB_OK;
}

}  // namespace

extern "C" {

int cb_pnp_create(cb_pnp** out, int max_candidates, int max_points_total, int max_hypotheses, int device) {
  if (!out) return cb::fail(CB_EINVAL, "out is NULL");
  *out = nullptr;
  if (max_candidates < 1 || max_points_total < 1 || max_hypotheses < 1)
    return cb::fail(CB_EINVAL, "bad sizes for cb_pnp_create");
  int sm = 0;
  int rc = cb::select_device(device, &sm);
  if (rc) return rc;
  cb::DeviceGuard g(device);
  cb_pnp* p = new cb_pnp();
  p->device = device;
  p->sm_count = sm;
  p->max_cand = max_candidates;
  p->max_points = max_points_total;
  p->max_hyp = max_hypotheses;
  if (const char* env = getenv("CB_PNP_ELIM_V1")) p->elim_v1 = env[0] == '1';
  dls_pack_rows_kernel<<<1, 96>>>();  // per-lane copy of the template rows (idempotent; device calls may arrive on any stream)
  if (cudaDeviceSynchronize() != cudaSuccess) {
    const cudaError_t e0 = cudaGetLastError();
    delete p;
    return cb::fail(CB_ECUDA, "dls_pack_rows_kernel failed: %s", cudaGetErrorString(e0));
  }
  const size_t all = (size_t)max_candidates * max_hypotheses;
  if ((size_t)p->chunk > all) p->chunk = (int)all;
  const size_t ch = (size_t)p->chunk;
  cudaError_t e = cudaSuccess;
#define CB_ALLOC(ptr, bytes)                                   \
  if (e == cudaSuccess) e = cudaMalloc((void**)&(ptr), (bytes))
  CB_ALLOC(p->offsets, (size_t)(max_candidates + 1) * sizeof(int));
  CB_ALLOC(p->X, (size_t)max_points_total * 3 * sizeof(double));
  CB_ALLOC(p->uv, (size_t)max_points_total * 3 * sizeof(double));  // 2 per point for PnP, 3 for the 3D-3D variant
  CB_ALLOC(p->cost, all * sizeof(double));
  CB_ALLOC(p->ninl, all * sizeof(int));
  CB_ALLOC(p->model, all * 12 * sizeof(double));
  CB_ALLOC(p->idx, ch * kSample * sizeof(int));
  CB_ALLOC(p->status, ch * sizeof(int));
  CB_ALLOC(p->coef, ch * 60 * sizeof(double));
  CB_ALLOC(p->T, ch * 27 * sizeof(double));
  CB_ALLOC(p->S, ch * kN * kN * sizeof(double));
  CB_ALLOC(p->cmodel, ch * kMaxSol * 12 * sizeof(double));
  CB_ALLOC(p->out_T, (size_t)max_candidates * 16 * sizeof(double));
  CB_ALLOC(p->out_conf, (size_t)max_candidates * sizeof(float));
  CB_ALLOC(p->out_i, (size_t)max_candidates * 3 * sizeof(int));
#undef CB_ALLOC
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    cb_pnp_destroy(p);
    return cb::fail(CB_ENOMEM, "cb_pnp_create allocation failed: %s", cudaGetErrorString(e));
  }
  *out = p;
  return CB_OK;
}

int cb_pnp_destroy(cb_pnp* p) {
  if (!p) return CB_OK;
  cb::DeviceGuard g(p->device);
  if (p->stream) cb::sync_stream(p->stream);
  void* ptrs[] = {p->offsets, p->X, p->uv, p->samples, p->cost, p->ninl, p->model, p->idx, p->status,
                  p->coef, p->T, p->S, p->cmodel, p->out_T, p->out_conf, p->out_i};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return CB_OK;
}

int cb_pnp_solve_batch_device(cb_pnp* p, int n_cand, const int32_t* offsets_dev, int total_points,
                              const double* X_dev, const double* uv_dev, const cb_ransac_params* params,
                              const int32_t* samples_dev, double* c_T_w_dev, float* confidence_dev,
                              int32_t* num_iterations_dev, int32_t* n_inliers_dev, int32_t* best_hyp_dev,
                              void* stream) {
  if (!p || !offsets_dev || !X_dev || !uv_dev || !params || !c_T_w_dev || !confidence_dev)
    return cb::fail(CB_EINVAL, "NULL argument to cb_pnp_solve_batch_device");
  if (n_cand < 1 || n_cand > p->max_cand) return cb::fail(CB_EINVAL, "n_cand %d outside [1,%d]", n_cand, p->max_cand);
  const int H = params->max_iterations;
  if (H < 1 || H > p->max_hyp) return cb::fail(CB_EINVAL, "max_iterations %d outside [1,%d]", H, p->max_hyp);
  if (!params->use_mle) return cb::fail(CB_EINVAL, "only use_mle=1 (the reference's setting) is implemented");
  (void)total_points;
  cb::DeviceGuard g(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = run_chunks(p, n_cand, offsets_dev, X_dev, uv_dev, H, *params, samples_dev, st);
  if (rc) return rc;
  SelectArgs sa;
  sa.offsets = offsets_dev;
  sa.n_cand = n_cand;
  sa.H = H;
  sa.cost = p->cost;
  sa.ninl = p->ninl;
  sa.model = p->model;
  sa.p = *params;
  sa.sample_size = kSample;
  sa.c_T_w = c_T_w_dev;
  sa.confidence = confidence_dev;
  sa.num_iterations = num_iterations_dev;
  sa.n_inliers = n_inliers_dev;
  sa.best_hyp = best_hyp_dev;
  pnp_select_kernel<<<n_cand, 32, 0, st>>>(sa);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int cb_pnp_solve_batch(cb_pnp* p, int n_cand, const int32_t* offsets, const double* X, const double* uv,
                       const cb_ransac_params* params, const int32_t* samples, double* c_T_w, float* confidence,
                       int32_t* num_iterations, int32_t* n_inliers, int32_t* best_hyp) {
  if (!p || !offsets || !X || !uv || !params || !c_T_w || !confidence)
    return cb::fail(CB_EINVAL, "NULL argument to cb_pnp_solve_batch");
  if (n_cand < 1 || n_cand > p->max_cand) return cb::fail(CB_EINVAL, "n_cand %d outside [1,%d]", n_cand, p->max_cand);
  const int total = offsets[n_cand];
  if (total < 0 || total > p->max_points) return cb::fail(CB_EINVAL, "total points %d outside [0,%d]", total, p->max_points);
  cb::DeviceGuard g(p->device);
  cudaStream_t st = p->stream;
  CB_CUDA(cudaMemcpyAsync(p->offsets, offsets, (size_t)(n_cand + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  CB_CUDA(cudaMemcpyAsync(p->X, X, (size_t)total * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  CB_CUDA(cudaMemcpyAsync(p->uv, uv, (size_t)total * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  const int* samples_dev = nullptr;
  if (samples) {
    const size_t ne = (size_t)n_cand * params->max_iterations * kSample;
    if (p->samples_elems < ne) {
      if (p->samples) cudaFree(p->samples);
      p->samples = nullptr;
      p->samples_elems = 0;
      CB_CUDA(cudaMalloc((void**)&p->samples, ne * sizeof(int)));
      p->samples_elems = ne;
    }
    CB_CUDA(cudaMemcpyAsync(p->samples, samples, ne * sizeof(int), cudaMemcpyHostToDevice, st));
    samples_dev = p->samples;
  }
  int* oi = p->out_i;
  int rc = cb_pnp_solve_batch_device(p, n_cand, p->offsets, total, p->X, p->uv, params, samples_dev, p->out_T,
                                     p->out_conf, oi, oi + p->max_cand, oi + 2 * p->max_cand, st);
  if (rc) return rc;
  CB_CUDA(cudaMemcpyAsync(c_T_w, p->out_T, (size_t)n_cand * 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaMemcpyAsync(confidence, p->out_conf, (size_t)n_cand * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (num_iterations)
    CB_CUDA(cudaMemcpyAsync(num_iterations, oi, (size_t)n_cand * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (n_inliers)
    CB_CUDA(cudaMemcpyAsync(n_inliers, oi + p->max_cand, (size_t)n_cand * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (best_hyp)
    CB_CUDA(cudaMemcpyAsync(best_hyp, oi + 2 * p->max_cand, (size_t)n_cand * sizeof(int), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cb::sync_stream(st));
  return CB_OK;
}

int cb_pnp_icp_batch_device(cb_pnp* p, int n_cand, const int32_t* offsets_dev, const double* A_dev, const double* B_dev,
                            const cb_ransac_params* params, const int32_t* samples_dev, double* b_T_a_dev,
                            float* confidence_dev, int32_t* num_iterations_dev, int32_t* n_inliers_dev,
                            int32_t* best_hyp_dev, void* stream) {
  if (!p || !offsets_dev || !A_dev || !B_dev || !params || !b_T_a_dev || !confidence_dev)
    return cb::fail(CB_EINVAL, "NULL argument to cb_pnp_icp_batch_device");
  if (n_cand < 1 || n_cand > p->max_cand) return cb::fail(CB_EINVAL, "n_cand %d outside [1,%d]", n_cand, p->max_cand);
  const int H = params->max_iterations;
  if (H < 1 || H > p->max_hyp) return cb::fail(CB_EINVAL, "max_iterations %d outside [1,%d]", H, p->max_hyp);
  if (!params->use_mle) return cb::fail(CB_EINVAL, "only use_mle=1 (the reference's setting) is implemented");
  cb::DeviceGuard g(p->device);
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)n_cand * H;
  for (long long g0 = 0; g0 < total; g0 += p->chunk) {
    const int count = (int)((total - g0) < p->chunk ? (total - g0) : p->chunk);
    IcpArgs ia;
    ia.offsets = offsets_dev;
    ia.A = A_dev;
    ia.B = B_dev;
    ia.samples = samples_dev;
    ia.seed = params->seed;
    ia.H = H;
    ia.g0 = g0;
    ia.count = count;
    ia.status = p->status;
    ia.model = p->cmodel;
    icp_model_kernel<<<(count + 127) / 128, 128, 0, st>>>(ia);
    CB_LAUNCH_CHECK();
    icp_score_kernel<<<(count + 3) / 4, 128, 0, st>>>(offsets_dev, A_dev, B_dev, H, g0, count, p->status, p->cmodel,
                                                      params->error_thresh, p->cost, p->ninl);
    CB_LAUNCH_CHECK();
    copy_models_kernel<<<(count * 12 + 255) / 256, 256, 0, st>>>(p->cmodel, p->status, g0, count, p->model);
    CB_LAUNCH_CHECK();
  }
  SelectArgs sa;
  sa.offsets = offsets_dev;
  sa.n_cand = n_cand;
  sa.H = H;
  sa.cost = p->cost;
  sa.ninl = p->ninl;
  sa.model = p->model;
  sa.p = *params;
  sa.sample_size = kIcpSample;
  sa.c_T_w = b_T_a_dev;
  sa.confidence = confidence_dev;
  sa.num_iterations = num_iterations_dev;
  sa.n_inliers = n_inliers_dev;
  sa.best_hyp = best_hyp_dev;
  pnp_select_kernel<<<n_cand, 32, 0, st>>>(sa);
  CB_LAUNCH_CHECK();
  return CB_OK;
}

int cb_pnp_icp_batch(cb_pnp* p, int n_cand, const int32_t* offsets, const double* A, const double* B,
                     const cb_ransac_params* params, const int32_t* samples, double* b_T_a, float* confidence,
                     int32_t* num_iterations, int32_t* n_inliers, int32_t* best_hyp) {
  if (!p || !offsets || !A || !B || !params || !b_T_a || !confidence) return cb::fail(CB_EINVAL, "NULL argument to cb_pnp_icp_batch");
  if (n_cand < 1 || n_cand > p->max_cand) return cb::fail(CB_EINVAL, "n_cand %d outside [1,%d]", n_cand, p->max_cand);
  const int total = offsets[n_cand];
  if (total < 0 || total > p->max_points) return cb::fail(CB_EINVAL, "total points %d outside [0,%d]", total, p->max_points);
  cb::DeviceGuard g(p->device);
  cudaStream_t st = p->stream;
  CB_CUDA(cudaMemcpyAsync(p->offsets, offsets, (size_t)(n_cand + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  CB_CUDA(cudaMemcpyAsync(p->X, A, (size_t)total * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  CB_CUDA(cudaMemcpyAsync(p->uv, B, (size_t)total * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
  const int* samples_dev = nullptr;
  if (samples) {
    const size_t ne = (size_t)n_cand * params->max_iterations * kIcpSample;
    if (p->samples_elems < ne) {
      if (p->samples) cudaFree(p->samples);
      p->samples = nullptr;
      p->samples_elems = 0;
      CB_CUDA(cudaMalloc((void**)&p->samples, ne * sizeof(int)));
      p->samples_elems = ne;
    }
    CB_CUDA(cudaMemcpyAsync(p->samples, samples, ne * sizeof(int), cudaMemcpyHostToDevice, st));
    samples_dev = p->samples;
  }
  int* oi = p->out_i;
  int rc = cb_pnp_icp_batch_device(p, n_cand, p->offsets, p->X, p->uv, params, samples_dev, p->out_T, p->out_conf, oi,
                                   oi + p->max_cand, oi + 2 * p->max_cand, st);
  if (rc) return rc;
  CB_CUDA(cudaMemcpyAsync(b_T_a, p->out_T, (size_t)n_cand * 16 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cudaMemcpyAsync(confidence, p->out_conf, (size_t)n_cand * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (num_iterations) CB_CUDA(cudaMemcpyAsync(num_iterations, oi, (size_t)n_cand * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (n_inliers) CB_CUDA(cudaMemcpyAsync(n_inliers, oi + p->max_cand, (size_t)n_cand * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (best_hyp) CB_CUDA(cudaMemcpyAsync(best_hyp, oi + 2 * p->max_cand, (size_t)n_cand * sizeof(int), cudaMemcpyDeviceToHost, st));
  CB_CUDA(cb::sync_stream(st));
  return CB_OK;
}

int cb_pnp_dls_minimal(cb_pnp* p, int n_sets, int m, const double* X, const double* uv, int32_t* n_solutions,
                       double* R, double* t) {
  if (!p || !X || !uv || !n_solutions || !R || !t) return cb::fail(CB_EINVAL, "NULL argument to cb_pnp_dls_minimal");
  if (m != kSample) return cb::fail(CB_EINVAL, "the minimal solver is built for exactly %d points per set, got %d", kSample, m);
  if (n_sets < 1 || n_sets > p->chunk || (long long)n_sets * m > p->max_points)
    return cb::fail(CB_EINVAL, "n_sets %d too large for this handle", n_sets);
  cb::DeviceGuard g(p->device);
  cudaStream_t st = p->stream;
  // every set is its own "candidate" of 15 points... but candidates need >= 20 points, so drive the
  // stages directly with an identity sample table
  const int total = n_sets * m;
  int* h_off = new int[2];
  h_off[0] = 0;
  h_off[1] = total;  // one pseudo-candidate holding all points (n >= 20 as soon as n_sets >= 2)
  int* h_samples = new int[(size_t)n_sets * m];
  for (int i = 0; i < n_sets * m; ++i) h_samples[i] = i;
  if (p->samples_elems < (size_t)n_sets * m) {
    if (p->samples) cudaFree(p->samples);
    p->samples = nullptr;
    p->samples_elems = 0;
    cudaError_t e = cudaMalloc((void**)&p->samples, (size_t)n_sets * m * sizeof(int));
    if (e != cudaSuccess) {
      delete[] h_off;
      delete[] h_samples;
      return cb::fail(CB_ENOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    p->samples_elems = (size_t)n_sets * m;
  }
  cudaMemcpyAsync(p->offsets, h_off, 2 * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(p->samples, h_samples, (size_t)n_sets * m * sizeof(int), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(p->X, X, (size_t)total * 3 * sizeof(double), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(p->uv, uv, (size_t)total * 2 * sizeof(double), cudaMemcpyHostToDevice, st);
  cb::sync_stream(st);
  delete[] h_off;
  delete[] h_samples;
  if (total < 20) return cb::fail(CB_EINVAL, "need at least 2 sets");
  const size_t roots_smem = (size_t)kRootsWarps * kRootsSmemPerWarp * sizeof(double);
  CB_CUDA(cudaFuncSetAttribute(dls_roots_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)roots_smem));
  SetupArgs sa;
  sa.offsets = p->offsets;
  sa.X = p->X;
  sa.uv = p->uv;
  sa.samples = p->samples;
  sa.seed = 0;
  sa.H = n_sets;
  sa.g0 = 0;
  sa.count = n_sets;
  sa.idx_out = p->idx;
  sa.coef_out = p->coef;
  sa.T_out = p->T;
  sa.status = p->status;
  dls_setup_kernel<<<(n_sets + 127) / 128, 128, 0, st>>>(sa);
  CB_LAUNCH_CHECK();
  if (int rc = launch_eliminate(p, n_sets, st)) return rc;
  RootsArgs ra;
  ra.S = p->S;
  ra.T = p->T;
  ra.idx = p->idx;
  ra.X = p->X;
  ra.status = p->status;
  ra.model = p->cmodel;
  ra.count = n_sets;
  ra.max_sol = kMaxSol;
  dls_roots_kernel<<<(n_sets + kRootsWarps - 1) / kRootsWarps, 32 * kRootsWarps, roots_smem, st>>>(ra);
  CB_LAUNCH_CHECK();
  double* hm = new double[(size_t)n_sets * kMaxSol * 12];
  cudaError_t e1 = cudaMemcpyAsync(hm, p->cmodel, (size_t)n_sets * kMaxSol * 12 * sizeof(double), cudaMemcpyDeviceToHost, st);
  cudaError_t e2 = cudaMemcpyAsync(n_solutions, p->status, (size_t)n_sets * sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaError_t e3 = cb::sync_stream(st);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    delete[] hm;
    return cb::fail(CB_ECUDA, "dls_minimal failed: %s", cudaGetErrorString(e3 != cudaSuccess ? e3 : (e1 != cudaSuccess ? e1 : e2)));
  }
  for (int s = 0; s < n_sets; ++s)
    for (int j = 0; j < kMaxSol; ++j) {
      const double* m12 = hm + ((size_t)s * kMaxSol + j) * 12;
      for (int i = 0; i < 9; ++i) R[((size_t)s * kMaxSol + j) * 9 + i] = m12[i];
      for (int i = 0; i < 3; ++i) t[((size_t)s * kMaxSol + j) * 3 + i] = m12[9 + i];
    }
  delete[] hm;
  return CB_OK;
}

int64_t cb_pnp_debug_read(cb_pnp* p, int what, int n_sets, double* out, int64_t max_doubles) {
  if (!p || !out || n_sets < 1 || n_sets > p->chunk || what < 0 || what > 1) return cb::fail(CB_EINVAL, "bad arguments");
  cb::DeviceGuard g(p->device);
  const int64_t per = what == 0 ? kN * kN : 60;
  const int64_t n = per * n_sets;
  if (n > max_doubles) return cb::fail(CB_EINVAL, "buffer too small: %lld doubles needed", (long long)n);
  CB_CUDA(cb::sync_stream(p->stream));
  CB_CUDA(cudaMemcpy(out, what == 0 ? p->S : p->coef, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  return n;
}

}  // extern "C"
