// Part A -- multi-bandwidth Gaussian MMD on sm_100a (reference: code/MMD.py:3-74).
//
//   K1  prep_colsum / prep_center : Z = [X; Y] centred, rounded to TF32 (hi [+ lo]), transposed copy,
//                                   row norms r_i, block weights a_i, sum r  -> analytic bandwidth
//   K2  mmd_fwd_kernel            : persistent, warp-specialised: TMA -> smem ring -> tcgen05.mma (kind::tf32)
//                                   -> TMEM (2 accumulator stages) -> fused distance / exp-sum / block-reduce
//                                   epilogue.  Only upper-triangular 128x128 tiles are computed; the n x n
//                                   kernel matrix never leaves the SM.
//   K3  mmd_bwd_kernel            : one CTA per (128-row panel, 256-column slice of d).  Re-computes the Gram
//                                   tiles, forms G in shared memory as a TF32 A-operand and accumulates
//                                   G.Z_J in TMEM with a second tcgen05.mma; nothing n x n is stored.
//
// Math: L_ij = max(0, r_i + r_j - 2 z_i.z_j); sigma_k = sigma_0 mul^k;
//   M = sum_ij a_i a_j sum_k exp(-L_ij / sigma_k),  a_i = 1/n_s (source rows) or -1/n_t (target rows);
//   Q_ij = sum_k exp(-L_ij / sigma_k) / mul^k;  D = sum_ij a_i a_j L_ij Q_ij / sigma_0^2;
//   G_ij = (-a_i a_j Q_ij / sigma_0 + c) [L_raw >= 0],  c = D / ((n^2 - n) mul^(num/2));
//   dZ_i = g sign(M) 4 (rowsum(G)_i z_i - (G Z)_i).
// Centring Z (distances are translation invariant) keeps TF32 rounding relative to the spread of the data
// and makes sum(L) = 2 n sum_i |z_i|^2 exactly the reference's bandwidth statistic (code/MMD.py:31).
#include <math.h>
#include <stdlib.h>

#include "../../include/edrl_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace edrl {
namespace mmd {

using namespace ptx;

constexpr int BM = 128;          // tile rows (I)
constexpr int BN = 128;          // tile cols (J)
constexpr int BK = 32;           // K chunk in floats = one 128-byte swizzle atom
constexpr int UMMA_K = 8;        // tf32: 32 bytes per MMA K step
constexpr int TILE_BYTES = BM * BK * 4;   // 16 KiB: one operand chunk
constexpr int MAX_KERNELS = 16;  // generic (non mul==2) path: at most this many bandwidths
constexpr float LOG2E = 1.4426950408889634f;

// ----------------------------------------------------------------------------- workspace
struct Layout {
  int n, n_pad, d_pad;
  bool split3, h16, s16;
  size_t off_acc, off_colsum, off_colsum_hi, off_colmax, off_r, off_a, off_fscale, off_zhi, off_zthi, off_zlo, off_ztlo,
      off_zt16, off_z16, off_rowsum, total;
  size_t zero_bytes;  // [off_acc, off_acc + zero_bytes) must be cleared before prep
};

static Layout make_layout(int n_s, int n_t, int d, int flags) {
  Layout L;
  L.n = n_s + n_t;
  L.n_pad = (int)align_up((size_t)L.n, 256);   // 256: the pair forward works on 256 x 256 tiles
  L.split3 = (flags & EDRL_MMD_3XTF32) != 0;
  L.s16 = !L.split3 && (flags & EDRL_MMD_F16S) != 0;
  L.h16 = !L.split3 && (flags & (EDRL_MMD_TF32H | EDRL_MMD_F16S)) != 0;
  // 64: the pair kernels stage two 32-column chunks at a time (F16S: two 64-column binary16 chunks)
  L.d_pad = (int)align_up((size_t)d, L.s16 ? 128 : 64);
  size_t o = 0;
  L.off_acc = o;      o += 256;                                   // 8 doubles + ticket counter
  L.off_colsum = o;   o += align_up((size_t)L.d_pad * 8, 256);    // double[d_pad]: column sums of [X; Y]
  L.off_colsum_hi = o; o += align_up((size_t)L.d_pad * 8, 256);   // double[d_pad]: column sums of the centred TF32 copy
  L.off_colmax = o;   o += align_up((size_t)L.d_pad * 4, 256);    // uint32[d_pad]: bits of max |x| per column
  L.off_r = o;        o += align_up((size_t)L.n_pad * 8, 256);    // double[n_pad]
  L.zero_bytes = o - L.off_acc;
  L.off_a = o;        o += align_up((size_t)L.n_pad * 4, 256);    // float[n_pad]
  L.off_fscale = o;   o += align_up((size_t)(L.d_pad + 1) * 4, 256);   // int[d_pad]: binary16 scale exponent per column; [d_pad]: of Z16
  o = align_up(o, 1024);
  size_t zbytes = align_up((size_t)L.n_pad * L.d_pad * 4, 1024);
  L.off_zhi = o;      o += zbytes;
  L.off_zthi = o;     o += zbytes;
  L.off_zlo = o;      if (L.split3) o += zbytes;
  L.off_ztlo = o;     if (L.split3) o += zbytes;
  L.off_zt16 = o;     if (L.h16) o += align_up((size_t)L.n_pad * L.d_pad * 2, 1024);   // binary16 Z^T [d_pad, n_pad]
  L.off_z16 = o;      if (L.s16) o += align_up((size_t)L.n_pad * L.d_pad * 2, 1024);   // binary16 Z [n_pad, d_pad]
  // fused sweep: row sums of G' per (512-column feature pass, column slab), float[passes][8][n_pad]
  L.off_rowsum = o;   if (!L.split3) o += align_up((size_t)((L.d_pad + 511) / 512) * 8 * L.n_pad * 4, 1024);
  L.total = o;
  return L;
}

// ----------------------------------------------------------------------------- K1: prep
// column sums of Z = [X; Y] in double (for the mean)
__global__ void __launch_bounds__(128) prep_colsum_kernel(const float *__restrict__ X, const float *__restrict__ Y,
                                                          int n_s, int n, int d, double *__restrict__ colsum,
                                                          unsigned *__restrict__ colmax) {
  const int col = blockIdx.x * 128 + threadIdx.x;
  const int r0 = blockIdx.y * 64;
  if (col >= d) return;
  float acc = 0.f, mx = 0.f;
  const int r1 = min(r0 + 64, n);
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const float *src = (r < n_s) ? (X + (size_t)r * d) : (Y + (size_t)(r - n_s) * d);
    const float v = __ldg(src + col);
    acc += v;
    mx = fmaxf(mx, fabsf(v));
  }
  atomicAdd(colsum + col, (double)acc);
  if (colmax) atomicMax(colmax + col, __float_as_uint(mx));       // non-negative floats order like their bit patterns
}

// the same with 128-bit loads (d % 4 == 0, 16-byte aligned inputs): a thread owns 4 columns and 16 rows (8 loads in
// flight), a block of 128 x 4 threads 512 columns x 64 rows, reduced through shared memory to one atomic per column --
// the scalar version keeps 14 KiB in flight per SM and runs at 2 TB/s
__global__ void __launch_bounds__(512) prep_colsum_vec4_kernel(const float *__restrict__ X, const float *__restrict__ Y,
                                                               int n_s, int n, int d, double *__restrict__ colsum,
                                                               unsigned *__restrict__ colmax) {
  __shared__ float4 s_acc[3][128], s_max[3][128];
  const int col = (blockIdx.x * 128 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * 64 + threadIdx.y * 16;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), mx = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < d) {
    const int r1 = min(r0 + 16, n);
#pragma unroll 8
    for (int r = r0; r < r1; ++r) {
      const float *src = (r < n_s) ? (X + (size_t)r * d) : (Y + (size_t)(r - n_s) * d);
      const float4 v = __ldg(reinterpret_cast<const float4 *>(src + col));
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      mx.x = fmaxf(mx.x, fabsf(v.x)); mx.y = fmaxf(mx.y, fabsf(v.y));
      mx.z = fmaxf(mx.z, fabsf(v.z)); mx.w = fmaxf(mx.w, fabsf(v.w));
    }
  }
  if (threadIdx.y > 0) {
    s_acc[threadIdx.y - 1][threadIdx.x] = acc;
    s_max[threadIdx.y - 1][threadIdx.x] = mx;
  }
  __syncthreads();
  if (threadIdx.y == 0 && col < d) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 a2 = s_acc[k][threadIdx.x], m2 = s_max[k][threadIdx.x];
      acc.x += a2.x; acc.y += a2.y; acc.z += a2.z; acc.w += a2.w;
      mx.x = fmaxf(mx.x, m2.x); mx.y = fmaxf(mx.y, m2.y); mx.z = fmaxf(mx.z, m2.z); mx.w = fmaxf(mx.w, m2.w);
    }
    atomicAdd(colsum + col + 0, (double)acc.x);
    atomicAdd(colsum + col + 1, (double)acc.y);
    atomicAdd(colsum + col + 2, (double)acc.z);
    atomicAdd(colsum + col + 3, (double)acc.w);
    if (colmax) {
      atomicMax(colmax + col + 0, __float_as_uint(mx.x));
      atomicMax(colmax + col + 1, __float_as_uint(mx.y));
      atomicMax(colmax + col + 2, __float_as_uint(mx.z));
      atomicMax(colmax + col + 3, __float_as_uint(mx.w));
    }
  }
}

// centre, round to tf32 (hi, optionally lo), write Z [n_pad, d_pad] and Z^T [d_pad, n_pad], row norms, weights
template <bool SPLIT3, bool H16 = false>
__global__ void __launch_bounds__(256)
prep_center_kernel(const float *__restrict__ X, const float *__restrict__ Y, int n_s, int n_t, int d, int n_pad,
                   int d_pad, const double *__restrict__ colsum, float *__restrict__ zhi, float *__restrict__ zthi,
                   float *__restrict__ zlo, float *__restrict__ ztlo, double *__restrict__ racc,
                   float *__restrict__ a, double *__restrict__ acc, double *__restrict__ colsum_hi,
                   const unsigned *__restrict__ colmax = nullptr, int *__restrict__ fscale = nullptr,
                   __half *__restrict__ zt16 = nullptr, __half *__restrict__ z16 = nullptr) {
  __shared__ float tile_hi[32][33];
  __shared__ float s_scale[32];
  __shared__ float s_gmax[8];
  __shared__ float tile_lo[SPLIT3 ? 32 : 1][33];
  __shared__ float blk_sum[8];
  const int n = n_s + n_t;
  const int lane = threadIdx.x, wy = threadIdx.y;
  const int row0 = blockIdx.x * 32;
  const double inv_n = 1.0 / (double)n;
  float rs[4] = {0.f, 0.f, 0.f, 0.f};
  float gscale = 1.f;
  if (H16 && z16) {
    // F16S: the Gram operand Z16 = Z 2^e with ONE exponent for the whole matrix (a per-row or per-column scale would
    // not factor out of z_i . z_j): |z| <= max_c (max |x_c| + |mean_c|) < 2^ex, e = 15 - ex
    float b = 0.f;
    for (int c = wy * 32 + lane; c < d; c += 256)
      b = fmaxf(b, __uint_as_float(colmax[c]) + fabsf((float)(colsum[c] * inv_n)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (lane == 0) s_gmax[wy] = b;
    __syncthreads();
    b = s_gmax[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) b = fmaxf(b, s_gmax[k]);
    int ex = 0;
    if (b > 0.f) frexpf(b, &ex);
    const int e = (b > 0.f) ? 15 - ex : 0;
    gscale = ldexpf(1.f, e);
    if (blockIdx.x == 0 && blockIdx.y == 0 && wy == 0 && lane == 0) fscale[d_pad] = e;
  }
  // the raw values of the next column tile are fetched while the current one is processed (two block barriers per
  // tile would otherwise leave 4 loads in flight per thread)
  auto load_tile = [&](int ct, float (&out)[4]) {
    const int col = ct * 32 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int row = row0 + wy * 4 + k;
      out[k] = 0.f;
      if (row < n && col < d) {
        const float *src = (row < n_s) ? (X + (size_t)row * d) : (Y + (size_t)(row - n_s) * d);
        out[k] = __ldg(src + col);
      }
    }
  };
  float raw[4], nxt[4] = {0.f, 0.f, 0.f, 0.f};
  if ((int)blockIdx.y < d_pad / 32) load_tile(blockIdx.y, raw);
  for (int ct = blockIdx.y; ct < d_pad / 32; ct += gridDim.y) {
    if (ct + (int)gridDim.y < d_pad / 32) load_tile(ct + gridDim.y, nxt);
    const int col = ct * 32 + lane;
    const float mean = (col < d) ? (float)(colsum[col] * inv_n) : 0.f;
    if (H16 && wy == 0) {
      // binary16 copy of the column: scale by 2^e so that |z| < 2^14.  max |x - mean| <= max |x| + |mean|.
      const float bnd = (col < d) ? (__uint_as_float(colmax[col]) + fabsf(mean)) : 0.f;
      int ex = 0;
      if (bnd > 0.f) frexpf(bnd, &ex);                 // bnd < 2^ex
      const int e = (bnd > 0.f) ? 14 - ex : 0;
      s_scale[lane] = ldexpf(1.f, e);
      if (blockIdx.x == 0) fscale[col] = e;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int rr = wy * 4 + k;
      const int row = row0 + rr;
      const float v = (row < n && col < d) ? raw[k] - mean : 0.f;
      const float hi = to_tf32(v);
      zhi[(size_t)row * d_pad + col] = hi;
      if (H16 && z16) z16[(size_t)row * d_pad + col] = __float2half_rn(hi * gscale);   // exact unless it underflows
      tile_hi[rr][lane] = hi;
      if (SPLIT3) {
        const float lo = to_tf32(v - hi);
        zlo[(size_t)row * d_pad + col] = lo;
        tile_lo[rr][lane] = lo;
        rs[k] = fmaf(v, v, rs[k]);
      } else {
        rs[k] = fmaf(hi, hi, rs[k]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cc = wy * 4 + k;
      if (zthi) zthi[(size_t)(ct * 32 + cc) * n_pad + row0 + lane] = tile_hi[lane][cc];
      if (SPLIT3) ztlo[(size_t)(ct * 32 + cc) * n_pad + row0 + lane] = tile_lo[lane][cc];
      // the TF32 value has a 10-bit significand already: its scaled binary16 copy is exact (short of underflow)
      if (H16) zt16[(size_t)(ct * 32 + cc) * n_pad + row0 + lane] = __float2half_rn(tile_hi[lane][cc] * s_scale[cc]);
    }
    if (wy == 0) {
      // column sums of the rounded centred values (the closed-form bandwidth term of the fused gradient needs
      // sum_j z_j of exactly the operand the tensor core sees, not of the unrounded data)
      float cs = 0.f;
#pragma unroll 8
      for (int rr = 0; rr < 32; ++rr) cs += tile_hi[rr][lane] + (SPLIT3 ? tile_lo[rr][lane] : 0.f);
      if (cs != 0.f) atomicAdd(colsum_hi + ct * 32 + lane, (double)cs);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) raw[k] = nxt[k];
  }
  float wsum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float v = rs[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int row = row0 + wy * 4 + k;
    if (lane == 0) {
      if (v != 0.f) atomicAdd(racc + row, (double)v);
      if (blockIdx.y == 0) a[row] = (row < n_s) ? (1.0f / (float)n_s) : (row < n ? (-1.0f / (float)n_t) : 0.f);
    }
    wsum += v;
  }
  if (lane == 0) blk_sum[wy] = wsum;
  __syncthreads();
  if (wy == 0 && lane == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += blk_sum[k];
    if (s != 0.f) atomicAdd(acc + 2, (double)s);
  }
}

// ----------------------------------------------------------------------------- shared device helpers
struct KernelCoefs {          // per-launch bandwidth constants, built by every thread that needs them
  float sigma0;
  float inv_sigma0;
  float negc_last;            // -log2(e) / sigma_{num-1}      (fast path)
};

__device__ __forceinline__ double bandwidth_sigma0(double sum_r, int n, float mul, int num) {
  // code/MMD.py:31-34 with sum(L) = 2 n sum_i |z_i - mean|^2
  const double nn = (double)n;
  double s = 2.0 * nn * sum_r / (nn * nn - nn);
  for (int k = 0; k < num / 2; ++k) s /= (double)mul;
  return s;
}

// sum_k exp(-L/sigma_k) and Q = sum_k exp(-L/sigma_k)/mul^k.
// FAST: mul == 2, num == 5: one ex2 for the widest bandwidth, four squarings for the rest.
template <bool FAST>
__device__ __forceinline__ void kernel_terms(float L, float negc_last, const float *__restrict__ s_negc,
                                             const float *__restrict__ s_w, int num, float &K, float &Q) {
  if (FAST) {
    const float e4 = ex2_approx(L * negc_last);
    const float e3 = e4 * e4;
    const float e2 = e3 * e3;
    const float e1 = e2 * e2;
    const float e0 = e1 * e1;
    K = ((e0 + e1) + (e2 + e3)) + e4;
    Q = fmaf(fmaf(fmaf(fmaf(e4, 0.5f, e3), 0.5f, e2), 0.5f, e1), 0.5f, e0);
  } else {
    K = 0.f;
    Q = 0.f;
    for (int k = 0; k < num; ++k) {
      const float e = ex2_approx(L * s_negc[k]);
      K += e;
      Q = fmaf(e, s_w[k], Q);
    }
  }
}

__device__ __forceinline__ void fill_generic_coefs(float *s_negc, float *s_w, float sigma0, float mul, int num) {
  float sk = sigma0, w = 1.f;
  for (int k = 0; k < num; ++k) {
    s_negc[k] = -LOG2E / sk;
    s_w[k] = w;
    sk *= mul;
    w /= mul;
  }
}

__device__ __forceinline__ void decode_tile(long long t, int nb, int &I, int &J) {
  // upper-triangular tiles in row-major order: row I holds nb - I tiles, offset(I) = I nb - I (I-1) / 2.
  // fp32 estimate + exact integer correction (no fp64 in the per-tile path: sixteen epilogue warps decoding
  // with a double-precision sqrt showed up as FP64-pipe stalls in the profile).
  const float b = 2.0f * (float)nb + 1.0f;
  const float disc = fmaxf(b * b - 8.0f * (float)t, 0.0f);
  int i = (int)((b - sqrtf(disc)) * 0.5f);
  if (i < 0) i = 0;
  if (i > nb - 1) i = nb - 1;
  while ((long long)i * nb - (long long)i * (i - 1) / 2 > t) --i;
  while (i < nb - 1 && (long long)(i + 1) * nb - (long long)(i + 1) * i / 2 <= t) ++i;
  I = i;
  J = i + (int)(t - ((long long)i * nb - (long long)i * (i - 1) / 2));
}

__device__ __forceinline__ void write_final_stats(double M, double Dsum, double sum_r, int n, float mul, int num,
                                                  float *loss, float *stats) {
  const double sigma0 = bandwidth_sigma0(sum_r, n, mul, num);
  const double nn = (double)n;
  double half = 1.0;
  for (int k = 0; k < num / 2; ++k) half *= (double)mul;
  const double D = (sigma0 > 0.0) ? Dsum / (sigma0 * sigma0) : 0.0;
  const double c = D / ((nn * nn - nn) * half);
  *loss = (float)fabs(M);
  stats[EDRL_MMD_STAT_M] = (float)M;
  stats[EDRL_MMD_STAT_SIGMA0] = (float)sigma0;
  stats[EDRL_MMD_STAT_D] = (float)D;
  stats[EDRL_MMD_STAT_C] = (float)c;
  stats[EDRL_MMD_STAT_SUMR] = (float)sum_r;
  stats[5] = 0.f;
  stats[6] = 0.f;
  stats[7] = 0.f;
}

// ----------------------------------------------------------------------------- K2: forward
enum { MODE_LOSS = 0, MODE_KMAT = 1, MODE_GRAM = 2 };

struct FwdParams {
  int n, n_s, n_t, n_pad, d_pad, nb, kchunks, num;
  float mul;
  long long tiles_total;     // nb (nb + 1) / 2
  int tile_rank, tile_world;
  const double *racc;        // double[n_pad] row norms
  const float *a;            // float[n_pad] block weights
  double *acc;               // [0] M, [1] sum a a L Q, [2] sum r
  unsigned *ticket;
  float *loss, *stats;
  double *partial;           // sharded evaluation: partial sums out
  float *out;                // MODE_KMAT / MODE_GRAM: [n, n]
};

constexpr int FWD_THREADS = 320;          // warp 0 TMA, warp 1 MMA + TMEM alloc, warps 2-9 epilogue
constexpr int FWD_EPI_THREADS = 256;

template <bool SPLIT3>
struct FwdCfg {
  static constexpr int STAGE_BYTES = (SPLIT3 ? 4 : 2) * TILE_BYTES;
  static constexpr int STAGES = SPLIT3 ? 3 : 6;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4096 + 1024;   // + control block + alignment slack
};

struct FwdCtrl {                 // lives after the operand ring
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
  float2 colinfo[2][BN];         // (r_j, a_j) of the current J tile, per accumulator stage
  float negc[MAX_KERNELS];
  float w[MAX_KERNELS];
  double red[8][2];
};

template <bool SPLIT3, int MODE, bool FAST>
__global__ void __launch_bounds__(FWD_THREADS, 1)
mmd_fwd_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const FwdParams p) {
  using Cfg = FwdCfg<SPLIT3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  FwdCtrl *ctl = reinterpret_cast<FwdCtrl *>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->tmem_full[s], 1);
      mbar_init(&ctl->tmem_empty[s], FWD_EPI_THREADS / 32);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 256);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    if (SPLIT3) tma_prefetch_desc(&tm_lo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  // this CTA's tiles: local index q = blockIdx.x + i * gridDim.x, global tile t = tile_rank + tile_world * q
  const long long my_first = blockIdx.x;
  const long long q_total = (p.tiles_total - p.tile_rank + p.tile_world - 1) / p.tile_world;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (long long q = my_first; q < q_total; q += gridDim.x) {
        int I, J;
        decode_tile(p.tile_rank + p.tile_world * q, p.nb, I, J);
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t *st = smem + s * Cfg::STAGE_BYTES;
          mbar_expect_tx(&ctl->full[s], Cfg::STAGE_BYTES);
          tma_load_2d(st, &tm_hi, &ctl->full[s], kc * BK, I * BM);
          tma_load_2d(st + TILE_BYTES, &tm_hi, &ctl->full[s], kc * BK, J * BN);
          if (SPLIT3) {
            tma_load_2d(st + 2 * TILE_BYTES, &tm_lo, &ctl->full[s], kc * BK, I * BM);
            tma_load_2d(st + 3 * TILE_BYTES, &tm_lo, &ctl->full[s], kc * BK, J * BN);
          }
          if (++s == Cfg::STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (long long q = my_first; q < q_total; q += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t u = (uint32_t)(it >> 1);
        mbar_wait(&ctl->tmem_empty[as], (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
          const uint64_t a_hi = make_kmajor_sw128_desc(sa);
          const uint64_t b_hi = make_kmajor_sw128_desc(sa + TILE_BYTES);
          const uint64_t a_lo = make_kmajor_sw128_desc(sa + 2 * TILE_BYTES);
          const uint64_t b_lo = make_kmajor_sw128_desc(sa + 3 * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
            const uint32_t first = (kc > 0 || k > 0) ? 1u : 0u;
            if (SPLIT3) {
              // small cross terms first, the dominant hi.hi term last
              mma_tf32_ss(d_tmem, a_lo + adv, b_hi + adv, idesc, first);
              mma_tf32_ss(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
              mma_tf32_ss(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            } else {
              mma_tf32_ss(d_tmem, a_hi + adv, b_hi + adv, idesc, first);
            }
          }
          mma_commit(&ctl->empty[s]);
          if (++s == Cfg::STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        mma_commit(&ctl->tmem_full[as]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: 8 warps, thread = one row, warpgroup = 64 columns =====================
    const int ew = warp - 2;                 // 0..7
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int ch = ew >> 2;                  // column half
    const int et = ew * 32 + lane;           // 0..255
    const int row = lg * 32 + lane;

    const double sum_r = p.acc[2];
    const float sigma0 = (float)bandwidth_sigma0(sum_r, p.n, p.mul, p.num);
    float sig_last = sigma0;
    for (int k = 0; k < p.num - 1; ++k) sig_last *= p.mul;
    const float negc_last = -LOG2E / sig_last;
    if (!FAST && et == 0) fill_generic_coefs(ctl->negc, ctl->w, sigma0, p.mul, p.num);
    // (the named barrier inside the tile loop orders these writes before their first use)

    double accM = 0.0, accD = 0.0;
    int it = 0;
    for (long long q = my_first; q < q_total; q += gridDim.x, ++it) {
      int I, J;
      decode_tile(p.tile_rank + p.tile_world * q, p.nb, I, J);
      const int as = it & 1;
      const uint32_t u = (uint32_t)(it >> 1);
      if (et < BN) {
        const int gj = J * BN + et;
        ctl->colinfo[as][et] = make_float2((float)p.racc[gj], p.a[gj]);
      }
      const int gi = I * BM + row;
      const float ri = (float)p.racc[gi];
      const float ai = p.a[gi];
      named_barrier_sync(1, FWD_EPI_THREADS);
      mbar_wait(&ctl->tmem_full[as], u & 1);
      tc_fence_after();
      float tM = 0.f, tD = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = ch * 64 + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + col0), v);
        tmem_ld_wait();
        if (MODE == MODE_LOSS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float2 ci = ctl->colinfo[as][col0 + j];
            float L = fmaf(-2.f, __uint_as_float(v[j]), ri + ci.x);
            L = fmaxf(L, 0.f);
            float K, Q;
            kernel_terms<FAST>(L, negc_last, ctl->negc, ctl->w, p.num, K, Q);
            tM = fmaf(ci.y, K, tM);
            tD = fmaf(ci.y * L, Q, tD);
          }
        } else {
          for (int j = 0; j < 32; ++j) {
            const int gj = J * BN + col0 + j;
            float val;
            if (MODE == MODE_GRAM) {
              val = __uint_as_float(v[j]);
            } else {
              const float2 ci = ctl->colinfo[as][col0 + j];
              float L = fmaxf(fmaf(-2.f, __uint_as_float(v[j]), ri + ci.x), 0.f);
              float Q;
              kernel_terms<FAST>(L, negc_last, ctl->negc, ctl->w, p.num, val, Q);
            }
            if (gi < p.n && gj < p.n) {
              p.out[(size_t)gi * p.n + gj] = val;
              if (I != J) p.out[(size_t)gj * p.n + gi] = val;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->tmem_empty[as]);
      if (MODE == MODE_LOSS) {
        const float wgt = (I == J) ? ai : 2.f * ai;
        accM += (double)(wgt * tM);
        accD += (double)(wgt * tD);
      }
    }
    if (MODE == MODE_LOSS) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        accM += __shfl_xor_sync(0xffffffffu, accM, o);
        accD += __shfl_xor_sync(0xffffffffu, accD, o);
      }
      if (lane == 0) {
        ctl->red[ew][0] = accM;
        ctl->red[ew][1] = accD;
      }
      named_barrier_sync(1, FWD_EPI_THREADS);
      if (et == 0) {
        double m = 0.0, dd = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          m += ctl->red[k][0];
          dd += ctl->red[k][1];
        }
        atomicAdd(p.acc + 0, m);
        atomicAdd(p.acc + 1, dd);
        __threadfence();
        const unsigned t = atomicAdd(p.ticket, 1u);
        if (t == gridDim.x - 1) {
          __threadfence();
          const double M = atomicAdd(p.acc + 0, 0.0);
          const double Ds = atomicAdd(p.acc + 1, 0.0);
          if (p.partial) {
            p.partial[0] = M;
            p.partial[1] = Ds;
          }
          if (p.tile_world == 1) write_final_stats(M, Ds, sum_r, p.n, p.mul, p.num, p.loss, p.stats);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}


// ----------------------------------------------------------------------------- K2p: CTA-pair forward (TF32)
// Persistent 2-CTA clusters; each pair owns 256 x 256 upper-triangular tiles of the Gram matrix:
// tcgen05 cta_group::2, M = 256 (128 rows of Z_I per CTA), N = 256 (128 rows of Z_J per CTA), K = d.
// Per 128 x 256 half-tile a CTA ingests 32 KiB per 32 columns of d (its rows of Z_I + half of Z_J) -- half of what
// the 128 x 128 single-CTA kernel moves per output element; both are bound by the L2 -> SM ingest rate.
constexpr int F2_TILE = 256;
constexpr int F2_STAGE = 2 * TILE_BYTES;          // 128 rows of Z_I + 128 rows of Z_J, 32 columns each
constexpr int F2_STAGES = 6;
constexpr int F2_CTRL_BYTES = 6144;
constexpr int F2_EPI_WARPS = 16;              // 4 per TMEM lane group, 64 accumulator columns each
constexpr int F2_EPI_THREADS = F2_EPI_WARPS * 32;
constexpr int F2_THREADS = 64 + F2_EPI_THREADS;
constexpr int F2_SMEM_BYTES = F2_STAGES * F2_STAGE + F2_CTRL_BYTES;

struct Fwd2Ctrl {
  uint64_t full[8];              // leader CTA only
  uint64_t empty[8];             // per CTA (multicast commit)
  uint64_t tmem_full[2];         // per CTA (multicast commit)
  uint64_t tmem_empty[2];        // leader, 16 arrivals
  uint32_t tmem_base;
  uint32_t pad;
  float2 colinfo[2][F2_TILE];    // (r_j, a_j) of the current J block, per accumulator stage
  float negc[MAX_KERNELS];
  float w[MAX_KERNELS];
  double red[F2_EPI_WARPS][2];
};
static_assert(sizeof(Fwd2Ctrl) <= F2_CTRL_BYTES, "Fwd2Ctrl does not fit its smem slot");
static_assert(F2_SMEM_BYTES <= 232448, "smem budget");

template <bool FAST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F2_THREADS, 1)
mmd_fwd_pair_kernel(const __grid_constant__ CUtensorMap tm_z, const FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Fwd2Ctrl *ctl = reinterpret_cast<Fwd2Ctrl *>(smem + F2_STAGES * F2_STAGE);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int nb2 = p.n_pad / F2_TILE;
  const long long tiles_total = (long long)nb2 * (nb2 + 1) / 2;
  const long long q_total = (tiles_total - p.tile_rank + p.tile_world - 1) / p.tile_world;
  const int kchunks = p.kchunks;

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (threadIdx.x == 0) {
    for (int s = 0; s < F2_STAGES; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->tmem_full[s], 1);
      mbar_init(&ctl->tmem_empty[s], 2 * F2_EPI_WARPS);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc_pair(&ctl->tmem_base, 512);
    tmem_relinquish_pair();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_z);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, warp-converged issue) =====================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t full0 = mapa_u32(smem_u32(&ctl->full[0]), 0);
    for (long long q = pair; q < q_total; q += npairs) {
      int I, J;
      decode_tile(p.tile_rank + p.tile_world * q, nb2, I, J);
      const int irow = I * F2_TILE + (int)rank * 128;
      const int jrow = J * F2_TILE + (int)rank * 128;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&ctl->empty[s], ph ^ 1);
        mbar_expect_tx_elect(&ctl->full[s], 2 * F2_STAGE, leader ? 1u : 0u);
        uint8_t *st = smem + s * F2_STAGE;
        const uint32_t bar = full0 + 8u * (uint32_t)s;
        tma_load_2d_pair_elect(st, &tm_z, bar, kc * BK, irow);
        tma_load_2d_pair_elect(st + TILE_BYTES, &tm_z, bar, kc * BK, jrow);
        if (++s == F2_STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, warp-converged issue) =====================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_tf32(256, F2_TILE);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      const uint32_t ring_addr = smem_u32(smem);
      for (long long q = pair; q < q_total; q += npairs, ++it) {
        const int as = it & 1;
        const uint32_t u = (uint32_t)(it >> 1);
        mbar_wait_cluster(&ctl->tmem_empty[as], (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * F2_TILE;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          const uint32_t sa = ring_addr + s * F2_STAGE;
          const uint64_t a_d = make_kmajor_sw128_desc(sa);
          const uint64_t b_d = make_kmajor_sw128_desc(sa + TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
            mma_tf32_ss_pair_elect(d_tmem, a_d + adv, b_d + adv, idesc, (kc > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit_pair_elect(&ctl->empty[s]);
          if (++s == F2_STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
        mma_commit_pair_elect(&ctl->tmem_full[as]);
      }
    }
  } else {
    // ===================== epilogue (both CTAs): thread = one row, 4 warps per lane group x 64 columns ==========
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int ch = ew >> 2;                  // column quarter (64 columns)
    const int et = ew * 32 + lane;
    const int row = lg * 32 + lane;

    const double sum_r = p.acc[2];
    const float sigma0 = (float)bandwidth_sigma0(sum_r, p.n, p.mul, p.num);
    float sig_last = sigma0;
    for (int k = 0; k < p.num - 1; ++k) sig_last *= p.mul;
    const float negc_last = -LOG2E / sig_last;
    if (!FAST && et == 0) fill_generic_coefs(ctl->negc, ctl->w, sigma0, p.mul, p.num);
    const uint32_t tmem_empty0 = mapa_u32(smem_u32(&ctl->tmem_empty[0]), 0);

    double accM = 0.0, accD = 0.0;
    int it = 0;
    for (long long q = pair; q < q_total; q += npairs, ++it) {
      int I, J;
      decode_tile(p.tile_rank + p.tile_world * q, nb2, I, J);
      const int as = it & 1;
      const uint32_t u = (uint32_t)(it >> 1);
      if (et < F2_TILE) {
        const int gj = J * F2_TILE + et;
        ctl->colinfo[as][et] = make_float2((float)p.racc[gj], p.a[gj]);
      }
      const int gi = I * F2_TILE + (int)rank * 128 + row;
      const float ri = (float)p.racc[gi];
      const float ai = p.a[gi];
      named_barrier_sync(1, F2_EPI_THREADS);
      mbar_wait(&ctl->tmem_full[as], u & 1);
      tc_fence_after();
      float tM = 0.f, tD = 0.f;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = ch * 64 + c * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * F2_TILE + col0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float2 ci = ctl->colinfo[as][col0 + j];
          float L = fmaf(-2.f, __uint_as_float(v[j]), ri + ci.x);
          L = fmaxf(L, 0.f);
          float K, Q;
          kernel_terms<FAST>(L, negc_last, ctl->negc, ctl->w, p.num, K, Q);
          tM = fmaf(ci.y, K, tM);
          tD = fmaf(ci.y * L, Q, tD);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tmem_empty0 + 8u * (uint32_t)as);
      const float wgt = (I == J) ? ai : 2.f * ai;
      accM += (double)(wgt * tM);
      accD += (double)(wgt * tD);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accM += __shfl_xor_sync(0xffffffffu, accM, o);
      accD += __shfl_xor_sync(0xffffffffu, accD, o);
    }
    if (lane == 0) {
      ctl->red[ew][0] = accM;
      ctl->red[ew][1] = accD;
    }
    named_barrier_sync(1, F2_EPI_THREADS);
    if (et == 0) {
      double m = 0.0, dd = 0.0;
#pragma unroll
      for (int k = 0; k < F2_EPI_WARPS; ++k) {
        m += ctl->red[k][0];
        dd += ctl->red[k][1];
      }
      atomicAdd(p.acc + 0, m);
      atomicAdd(p.acc + 1, dd);
      __threadfence();
      const unsigned t = atomicAdd(p.ticket, 1u);
      if (t == gridDim.x - 1) {
        __threadfence();
        const double M = atomicAdd(p.acc + 0, 0.0);
        const double Ds = atomicAdd(p.acc + 1, 0.0);
        if (p.partial) {
          p.partial[0] = M;
          p.partial[1] = Ds;
        }
        if (p.tile_world == 1) write_final_stats(M, Ds, sum_r, p.n, p.mul, p.num, p.loss, p.stats);
      }
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

__global__ void mmd_finalize_kernel(const double *partial, const double *acc, int n, float mul, int num, float *loss,
                                    float *stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0) write_final_stats(partial[0], partial[1], acc[2], n, mul, num, loss, stats);
}

// ----------------------------------------------------------------------------- K3: backward
constexpr int DC = 256;                       // columns of dZ one CTA accumulates in TMEM
constexpr int BWD_THREADS = 320;
constexpr int BWD_EPI_THREADS = 256;
constexpr int BWD_STAGE_BYTES = 2 * TILE_BYTES;   // A chunk + B chunk, or one 256 x 32 chunk of Z^T
constexpr int G_BYTES = BM * BN * 4;          // 64 KiB, four 128-byte-swizzle K atoms

template <bool SPLIT3>
struct BwdCfg {
  // SPLIT3 keeps G as hi + lo (2 x 64 KiB) and therefore a shorter operand ring (each step of
  // the 3xTF32 product needs a hi stage and a lo stage resident together).
  static constexpr int STAGES = SPLIT3 ? 2 : 4;
  static constexpr int G_TOTAL = SPLIT3 ? 2 * G_BYTES : G_BYTES;
  static constexpr int CTRL_BYTES = 6144;
  static constexpr int SMEM_BYTES = STAGES * BWD_STAGE_BYTES + G_TOTAL + CTRL_BYTES + 1024;
};

struct BwdCtrl {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t s_full[2];
  uint64_t s_empty[2];
  uint64_t g_full;
  uint64_t g_empty;
  uint64_t dz_full;
  uint32_t tmem_base;
  uint32_t pad;
  float4 colinfo[2][BN];          // (r_j, a_j, c_j, -) per S stage
  float negc[MAX_KERNELS];
  float w[MAX_KERNELS];
  float rowsum[2][BM];
};

static_assert(sizeof(BwdCtrl) <= 6144, "BwdCtrl does not fit its smem slot");
static_assert(sizeof(FwdCtrl) <= 4096, "FwdCtrl does not fit its smem slot");

struct BwdParams {
  int n, n_s, n_pad, d, d_pad, nb, kchunks, num;
  float mul;
  int row_begin, row_count;
  const double *racc;
  const float *a;
  const float *zhi, *zlo;          // [n_pad, d_pad]
  const float *stats;
  const float *grad_out;
  float *dz;                       // [row_count, d]
  // fused forward + gradient pass (mmd_sweep256_kernel / mmd_sweep_quad_kernel)
  double *acc;                     // [0] M, [1] sum a a L Q, [2] sum r (from prep)
  unsigned *ticket;
  double *partial;                 // optional: partial sums out (sharded evaluation)
  float *loss, *stats_out;         // written by the last CTA when finalize != 0
  int n_t, finalize;
  int row_begin2, row_count2;      // optional second row range (a rank's target rows); output rows follow range 1
  const int *fscale;               // TF32H: binary16 scale exponent per feature column
  // mmd_sweep256_kernel work list (make_plan): virtual panel = (feature pass, row panel); the first `full_items` virtual
  // panels sweep all column groups, every later one is split into `split` column slabs with one partial output each
  int panels, full_items, split, items;
  float *rowsum;                   // [feature pass][SW_MAX_SPLIT][n_pad]: rowsum(G')_i per column slab, for apply_grad
};

// ring order (producer and MMA issuer walk the same sequence):
//   S(0) chunks | for J: S(J+1) chunks (if any), Z^T(J) chunks
// SPLIT3 chunks: S -> (A_hi,B_hi), (A_lo,B_lo) per K chunk; Z^T -> hi chunk, lo chunk per 32 columns of J.
template <bool SPLIT3, bool FAST>
__global__ void __launch_bounds__(BWD_THREADS, 1)
mmd_bwd_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
               const __grid_constant__ CUtensorMap tm_thi, const __grid_constant__ CUtensorMap tm_tlo,
               const BwdParams p) {
  using Cfg = BwdCfg<SPLIT3>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *g_smem = smem + Cfg::STAGES * BWD_STAGE_BYTES;
  BwdCtrl *ctl = reinterpret_cast<BwdCtrl *>(g_smem + Cfg::G_TOTAL);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row_base = p.row_begin + blockIdx.x * BM;      // first global row of this panel
  const int f0 = blockIdx.y * DC;                           // first feature column of this slice
  const int nJ = p.nb;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->s_full[s], 1);
      mbar_init(&ctl->s_empty[s], BWD_EPI_THREADS / 32);
    }
    mbar_init(&ctl->g_full, BWD_EPI_THREADS);
    mbar_init(&ctl->g_empty, 1);
    mbar_init(&ctl->dz_full, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc(&ctl->tmem_base, 512);
    tmem_relinquish();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_thi);
    if (SPLIT3) {
      tma_prefetch_desc(&tm_lo);
      tma_prefetch_desc(&tm_tlo);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t tmem_dz = tmem_base;              // columns [0, 256)
  const uint32_t tmem_s = tmem_base + DC;          // two S stages of 128 columns

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      auto next = [&]() {
        if (++s == Cfg::STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      auto load_S = [&](int J) {
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t *st = smem + s * BWD_STAGE_BYTES;
          mbar_expect_tx(&ctl->full[s], BWD_STAGE_BYTES);
          tma_load_2d(st, &tm_hi, &ctl->full[s], kc * BK, row_base);
          tma_load_2d(st + TILE_BYTES, &tm_hi, &ctl->full[s], kc * BK, J * BN);
          next();
          if (SPLIT3) {
            mbar_wait(&ctl->empty[s], ph ^ 1);
            st = smem + s * BWD_STAGE_BYTES;
            mbar_expect_tx(&ctl->full[s], BWD_STAGE_BYTES);
            tma_load_2d(st, &tm_lo, &ctl->full[s], kc * BK, row_base);
            tma_load_2d(st + TILE_BYTES, &tm_lo, &ctl->full[s], kc * BK, J * BN);
            next();
          }
        }
      };
      auto load_Zt = [&](int J) {
        for (int a4 = 0; a4 < BN / BK; ++a4) {
          mbar_wait(&ctl->empty[s], ph ^ 1);
          uint8_t *st = smem + s * BWD_STAGE_BYTES;
          mbar_expect_tx(&ctl->full[s], BWD_STAGE_BYTES);
          tma_load_2d(st, &tm_thi, &ctl->full[s], J * BN + a4 * BK, f0);
          next();
          if (SPLIT3) {
            mbar_wait(&ctl->empty[s], ph ^ 1);
            st = smem + s * BWD_STAGE_BYTES;
            mbar_expect_tx(&ctl->full[s], BWD_STAGE_BYTES);
            tma_load_2d(st, &tm_tlo, &ctl->full[s], J * BN + a4 * BK, f0);
            next();
          }
        }
      };
      load_S(0);
      for (int J = 0; J < nJ; ++J) {
        if (J + 1 < nJ) load_S(J + 1);
        load_Zt(J);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_tf32(BM, BN);
      constexpr uint32_t idesc_p = make_idesc_tf32(BM, DC);
      int s = 0;
      uint32_t ph = 0;
      auto next = [&]() {
        if (++s == Cfg::STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      auto issue_S = [&](int J) {
        const int b = J & 1;
        const uint32_t u = (uint32_t)(J >> 1);
        mbar_wait(&ctl->s_empty[b], (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_s + b * BN;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * BWD_STAGE_BYTES);
          const uint64_t a_hi = make_kmajor_sw128_desc(sa);
          const uint64_t b_hi = make_kmajor_sw128_desc(sa + TILE_BYTES);
          if (!SPLIT3) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
              mma_tf32_ss(d_tmem, a_hi + adv, b_hi + adv, idesc_s, (kc > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(&ctl->empty[s]);
            next();
          } else {
            // stage s: (A_hi, B_hi); stage s+1: (A_lo, B_lo)
            const int s_hi = s;
            next();
            mbar_wait(&ctl->full[s], ph);
            tc_fence_after();
            const uint32_t sl = smem_u32(smem + s * BWD_STAGE_BYTES);
            const uint64_t a_lo = make_kmajor_sw128_desc(sl);
            const uint64_t b_lo = make_kmajor_sw128_desc(sl + TILE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
              mma_tf32_ss(d_tmem, a_lo + adv, b_hi + adv, idesc_s, (kc > 0 || k > 0) ? 1u : 0u);
              mma_tf32_ss(d_tmem, a_hi + adv, b_lo + adv, idesc_s, 1u);
              mma_tf32_ss(d_tmem, a_hi + adv, b_hi + adv, idesc_s, 1u);
            }
            mma_commit(&ctl->empty[s_hi]);
            mma_commit(&ctl->empty[s]);
            next();
          }
        }
        mma_commit(&ctl->s_full[b]);
      };
      auto issue_P = [&](int J) {
        mbar_wait(&ctl->g_full, (uint32_t)(J & 1));
        tc_fence_after();
        const uint32_t g_hi = smem_u32(g_smem);
        const uint32_t g_lo = g_hi + G_BYTES;
        for (int a4 = 0; a4 < BN / BK; ++a4) {
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          const uint64_t a_hi = make_kmajor_sw128_desc(g_hi + a4 * TILE_BYTES);
          const uint64_t b_hi = make_kmajor_sw128_desc(smem_u32(smem + s * BWD_STAGE_BYTES));
          if (!SPLIT3) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
              mma_tf32_ss(tmem_dz, a_hi + adv, b_hi + adv, idesc_p, (J > 0 || a4 > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(&ctl->empty[s]);
            next();
          } else {
            const int s_hi = s;
            next();
            mbar_wait(&ctl->full[s], ph);
            tc_fence_after();
            const uint64_t a_lo = make_kmajor_sw128_desc(g_lo + a4 * TILE_BYTES);
            const uint64_t b_lo = make_kmajor_sw128_desc(smem_u32(smem + s * BWD_STAGE_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
              mma_tf32_ss(tmem_dz, a_lo + adv, b_hi + adv, idesc_p, (J > 0 || a4 > 0 || k > 0) ? 1u : 0u);
              mma_tf32_ss(tmem_dz, a_hi + adv, b_lo + adv, idesc_p, 1u);
              mma_tf32_ss(tmem_dz, a_hi + adv, b_hi + adv, idesc_p, 1u);
            }
            mma_commit(&ctl->empty[s_hi]);
            mma_commit(&ctl->empty[s]);
            next();
          }
        }
        mma_commit(&ctl->g_empty);
      };
      issue_S(0);
      for (int J = 0; J < nJ; ++J) {
        if (J + 1 < nJ) issue_S(J + 1);
        issue_P(J);
      }
      mma_commit(&ctl->dz_full);
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int ch = ew >> 2;
    const int et = ew * 32 + lane;
    const int row = lg * 32 + lane;
    const int gi = row_base + row;
    const bool row_ok = (row < p.row_count - blockIdx.x * BM) && gi < p.n;

    const float M = p.stats[EDRL_MMD_STAT_M];
    const float sigma0 = p.stats[EDRL_MMD_STAT_SIGMA0];
    const float cval = p.stats[EDRL_MMD_STAT_C];
    float sig_last = sigma0;
    for (int k = 0; k < p.num - 1; ++k) sig_last *= p.mul;
    const float negc_last = -LOG2E / sig_last;
    if (!FAST && et == 0) fill_generic_coefs(ctl->negc, ctl->w, sigma0, p.mul, p.num);

    const float ri = (gi < p.n_pad) ? (float)p.racc[gi] : 0.f;
    const float ai = (gi < p.n_pad) ? p.a[gi] : 0.f;
    const float nai_sig = -ai / sigma0;
    float rowsum = 0.f;

    for (int J = 0; J < nJ; ++J) {
      const int b = J & 1;
      const uint32_t u = (uint32_t)(J >> 1);
      if (et < BN) {
        const int gj = J * BN + et;
        ctl->colinfo[b][et] = make_float4((float)p.racc[gj], p.a[gj], (gj < p.n) ? cval : 0.f, 0.f);
      }
      named_barrier_sync(1, BWD_EPI_THREADS);
      mbar_wait(&ctl->s_full[b], u & 1);
      tc_fence_after();
      // G buffer must have been consumed by the P-MMA of tile J-1
      mbar_wait(&ctl->g_empty, (uint32_t)((J & 1) ^ 1));
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = ch * 64 + c * 32;          // 32 columns == one swizzle atom of G
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * BN + col0), v);
        tmem_ld_wait();
        float g[32];
        float glo[SPLIT3 ? 32 : 1];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 ci = ctl->colinfo[b][col0 + j];
          const float Lraw = fmaf(-2.f, __uint_as_float(v[j]), ri + ci.x);
          const float L = fmaxf(Lraw, 0.f);
          float K, Q;
          kernel_terms<FAST>(L, negc_last, ctl->negc, ctl->w, p.num, K, Q);
          float gv = fmaf(ci.y * Q, nai_sig, ci.z);
          gv = (Lraw >= 0.f) ? gv : 0.f;
          const float gh = to_tf32(gv);
          g[j] = gh;
          if (SPLIT3) {
            const float gl = to_tf32(gv - gh);
            glo[j] = gl;
            rowsum += gh + gl;
          } else {
            rowsum += gh;
          }
        }
        // store this thread's 32 values of row `row` into K-atom (col0 / 32), 128-byte swizzle
        uint8_t *atom = g_smem + (col0 >> 5) * TILE_BYTES + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 val = make_float4(g[q4 * 4 + 0], g[q4 * 4 + 1], g[q4 * 4 + 2], g[q4 * 4 + 3]);
          *reinterpret_cast<float4 *>(atom + ((q4 ^ (row & 7)) << 4)) = val;
          if (SPLIT3) {
            float4 vl = make_float4(glo[q4 * 4 + 0], glo[q4 * 4 + 1], glo[q4 * 4 + 2], glo[q4 * 4 + 3]);
            *reinterpret_cast<float4 *>(atom + G_BYTES + ((q4 ^ (row & 7)) << 4)) = vl;
          }
        }
      }
      // S stage may be overwritten; G is visible to the tensor core (async proxy)
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&ctl->g_full);
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->s_empty[b]);
    }

    // ---- final: dZ = coef * (rowsum * z_i - P) ----
    ctl->rowsum[ch][row] = rowsum;
    named_barrier_sync(1, BWD_EPI_THREADS);
    const float rs_total = ctl->rowsum[0][row] + ctl->rowsum[1][row];
    const float sgn = (M > 0.f) ? 1.f : ((M < 0.f) ? -1.f : 0.f);
    const float coef = 4.f * sgn * p.grad_out[0];
    mbar_wait(&ctl->dz_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      const int col0 = ch * 128 + c * 32;
      if (f0 + col0 >= p.d) break;              // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(tmem_dz + ((uint32_t)(lg * 32) << 16) + (uint32_t)col0, v);
      tmem_ld_wait();
      if (row_ok) {
        const float *zr = p.zhi + (size_t)gi * p.d_pad + f0 + col0;
        const float *zl = SPLIT3 ? (p.zlo + (size_t)gi * p.d_pad + f0 + col0) : nullptr;
        float *out = p.dz + (size_t)(gi - p.row_begin) * p.d + f0 + col0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (f0 + col0 + j < p.d) {
            float zv = zr[j];
            if (SPLIT3) zv += zl[j];
            out[j] = coef * fmaf(rs_total, zv, -__uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ----------------------------------------------------------------------------- K3p: CTA-pair backward (TF32)
// One 2-CTA cluster (an SM pair, tcgen05 cta_group::2) owns a 128-row panel I of Z and 512 feature columns:
//   S phase   S_IJ = Z_I Z_J^T          M=128 (64 rows per CTA), N=128 (64 rows of J per CTA), K = d
//   epilogue  each CTA turns its 64 x 128 slice of S into G (TF32) in its own shared memory
//   P phase   dZ^T[f, i] += Zt[f, j] G[i, j]   M=256 features (128 per CTA), N=128 rows i (the two CTAs' G halves
//             are the two halves of the B operand -- no exchange), K = 128 columns j
// so the Gram tile is recomputed once per (I, J) -- not once per 256-column slice of d like mmd_bwd_kernel --
// and, for d <= 512, the Z_I rows stay resident in shared memory for the whole J loop.  Per J tile each CTA
// ingests 256 KiB (64 rows of Z_J + its 256 x 128 block of Z^T) instead of 640 KiB: both kernels are bound by
// the ~11 TB/s the L2 delivers to the SMs (profiles/), so bytes per tile is what sets the time.
constexpr int P2_STAGE = 16384;                 // ring stage per CTA
constexpr int P2_CHUNK = 8192;                  // 64 rows x 32 floats, 128-byte swizzle
constexpr int P2_ZI_BYTES = 16 * P2_CHUNK;      // 64 rows x 512 floats
constexpr int P2_G_BYTES = 4 * P2_CHUNK;        // 64 rows x 128 columns j
constexpr int P2_CTRL_BYTES = 3072;
constexpr int P2_FEATS = 512;                   // feature columns per pair and pass (2 M-tiles of 256)

// RES = number of 32-column chunks of Z_I kept resident (0: none, d > 512; 8: half of a 512-wide panel -- the
// other half is re-streamed so that 8 ring stages (128 KiB in flight per SM) still fit; 16 left only 4 stages and
// was TMA-latency bound).
template <int RES>
struct Bwd2Cfg {
  static constexpr int ZI_BYTES = RES * P2_CHUNK;
  static constexpr int STAGES = (232448 - ZI_BYTES - P2_G_BYTES - P2_CTRL_BYTES) / P2_STAGE > 11
                                    ? 11
                                    : (232448 - ZI_BYTES - P2_G_BYTES - P2_CTRL_BYTES) / P2_STAGE;
  static constexpr int SMEM_BYTES = ZI_BYTES + P2_G_BYTES + STAGES * P2_STAGE + P2_CTRL_BYTES;
};

struct Bwd2Ctrl {
  uint64_t full[12];              // used in the leader CTA only (both CTAs' TMA bytes land here)
  uint64_t empty[12];             // per CTA, arrived on by the multicast tcgen05.commit
  uint64_t zi_full;               // leader
  uint64_t s_full[2];             // per CTA (multicast commit)
  uint64_t s_empty[2];            // leader, 16 arrivals: 8 epilogue warps x 2 CTAs
  uint64_t g_full;                // leader, 16 arrivals
  uint64_t g_empty;               // per CTA (multicast commit)
  uint64_t dz_full;               // per CTA (multicast commit)
  uint32_t tmem_base;
  uint32_t pad;
  float2 colinfo[2][BN];          // (r_j, a_j) per S stage; re-used for the row-sum exchange after the J loop
  float negc[MAX_KERNELS];
  float w[MAX_KERNELS];
};
static_assert(sizeof(Bwd2Ctrl) <= P2_CTRL_BYTES, "Bwd2Ctrl does not fit its smem slot");
static_assert(Bwd2Cfg<0>::SMEM_BYTES <= 232448 && Bwd2Cfg<8>::SMEM_BYTES <= 232448, "smem budget");

// (The fused forward + gradient training pass lives in mmd_sweep256_kernel / mmd_sweep_quad_kernel below; this kernel is
// the separate backward of edrl_mmd_backward and the independent cross-check of the sweep in the tests.)
template <bool FAST, int RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(BWD_THREADS, 1)
mmd_bwd_pair_kernel(const __grid_constant__ CUtensorMap tm_z64, const __grid_constant__ CUtensorMap tm_zt,
                    const BwdParams p) {
  using Cfg = Bwd2Cfg<RES>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *zi_smem = smem;
  uint8_t *g_smem = smem + Cfg::ZI_BYTES;
  uint8_t *ring = g_smem + P2_G_BYTES;
  Bwd2Ctrl *ctl = reinterpret_cast<Bwd2Ctrl *>(ring + Cfg::STAGES * P2_STAGE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int panel = blockIdx.x >> 1;
  // the panels of row range 1 come first, then those of the optional range 2 (sharded: source rows, target rows)
  const int np1 = (p.row_count + BM - 1) / BM;
  const bool second = panel >= np1;
  const int lpanel = second ? panel - np1 : panel;
  const int rng_begin = second ? p.row_begin2 : p.row_begin;
  const int rng_count = second ? p.row_count2 : p.row_count;
  const int out_row0 = (second ? p.row_count : 0) + lpanel * BM;   // row of dz this panel starts at
  const int row_base = rng_begin + lpanel * BM;           // first global row of the pair's 128-row panel
  const int f0 = blockIdx.y * P2_FEATS;                   // first feature column of this pass
  // gridDim.z splits the column (J) range: slab z sweeps J tiles [J0, J0 + nJ) and writes its own partial output
  // slab (summed by edrl_mmd_apply_grad).  128 panels on 74 SM pairs are two waves; four slabs make it 1.75.
  const int J0 = (int)(((long long)blockIdx.z * p.nb) / gridDim.z);
  const int nJ = (int)(((long long)(blockIdx.z + 1) * p.nb) / gridDim.z) - J0;
  const int kchunks = p.kchunks;                          // even (d_pad is a multiple of 64)
  const int ntile = (p.d_pad - f0 > 256) ? 2 : 1;         // M-tiles of 256 features that hold real columns
  const int nres = (kchunks < RES) ? kchunks : RES;       // resident chunks of Z_I (even)

  if ((smem_u32(smem) & 1023u) != 0u) __trap();           // the swizzled tiles need a 1024-byte aligned base
  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    mbar_init(&ctl->zi_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->s_full[s], 1);
      mbar_init(&ctl->s_empty[s], 2 * BWD_EPI_THREADS / 32);
    }
    mbar_init(&ctl->g_full, 2 * BWD_EPI_THREADS / 32);
    mbar_init(&ctl->g_empty, 1);
    mbar_init(&ctl->dz_full, 1);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc_pair(&ctl->tmem_base, 512);
    tmem_relinquish_pair();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_z64);
    tma_prefetch_desc(&tm_zt);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                     // the peer's barriers exist before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t tmem_dz = tmem_base;                     // columns [0, 256): two M-tiles of dZ^T
  const uint32_t tmem_s = tmem_base + 256;                // two S stages of 64 columns (128 x 64 = 64 rows x 128)

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are signalled on the leader's barriers) ==========
    // The whole warp runs this code converged; one elected lane issues (see ptx.cuh, *_elect).
    {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t full0 = mapa_u32(smem_u32(&ctl->full[0]), 0);        // leader's full[0]; full[s] = + 8 s
      auto acquire = [&]() -> uint8_t * {
        mbar_wait(&ctl->empty[s], ph ^ 1);
        mbar_expect_tx_elect(&ctl->full[s], 2 * P2_STAGE, leader ? 1u : 0u);
        return ring + s * P2_STAGE;
      };
      auto next = [&]() {
        if (++s == Cfg::STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      const int irow = row_base + (int)rank * 64;
      if (RES > 0) {
        mbar_expect_tx_elect(&ctl->zi_full, 2u * (uint32_t)nres * P2_CHUNK, leader ? 1u : 0u);
        const uint32_t bar = mapa_u32(smem_u32(&ctl->zi_full), 0);
        for (int kc = 0; kc < nres; ++kc)
          tma_load_2d_pair_elect(zi_smem + kc * P2_CHUNK, &tm_z64, bar, kc * BK, irow);
      }
      auto load_S = [&](int J) {
        const int jrow = (J0 + J) * BN + (int)rank * 64;
        for (int kc = 0; kc < nres; kc += 2) {           // Z_I resident: a stage carries two chunks of Z_J
          uint8_t *st = acquire();
          const uint32_t bar = full0 + 8u * (uint32_t)s;
          tma_load_2d_pair_elect(st, &tm_z64, bar, kc * BK, jrow);
          tma_load_2d_pair_elect(st + P2_CHUNK, &tm_z64, bar, (kc + 1) * BK, jrow);
          next();
        }
        for (int kc = nres; kc < kchunks; ++kc) {        // streamed: a stage carries one chunk of Z_I and of Z_J
          uint8_t *st = acquire();
          const uint32_t bar = full0 + 8u * (uint32_t)s;
          tma_load_2d_pair_elect(st, &tm_z64, bar, kc * BK, irow);
          tma_load_2d_pair_elect(st + P2_CHUNK, &tm_z64, bar, kc * BK, jrow);
          next();
        }
      };
      auto load_P = [&](int J) {
        for (int t = 0; t < ntile; ++t)
          for (int a4 = 0; a4 < BN / BK; ++a4) {
            uint8_t *st = acquire();
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            tma_load_2d_pair_elect(st, &tm_zt, bar, (J0 + J) * BN + a4 * BK, f0 + t * 256 + (int)rank * 128);
            next();
          }
      };
      load_S(0);
      for (int J = 0; J < nJ; ++J) {
        if (J + 1 < nJ) load_S(J + 1);
        load_P(J);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the leader CTA's warp 1 drives both tensor cores =====================
    if (leader) {
      constexpr uint32_t idesc_s = make_idesc_tf32(128, BN);      // 64 rows per CTA
      constexpr uint32_t idesc_p = make_idesc_tf32(256, BN);      // 128 feature rows per CTA
      int s = 0;
      uint32_t ph = 0;
      auto next = [&]() {
        if (++s == Cfg::STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      const uint32_t ring_addr = smem_u32(ring);
      const uint32_t zi_addr = smem_u32(zi_smem);
      const uint32_t g_addr = smem_u32(g_smem);
      if (RES > 0) {
        mbar_wait(&ctl->zi_full, 0);
        tc_fence_after();
      }
      auto issue_S = [&](int J) {
        const int b = J & 1;
        const uint32_t u = (uint32_t)(J >> 1);
        mbar_wait_cluster(&ctl->s_empty[b], (u & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_s + b * 64;
        for (int kc = 0; kc < nres; kc += 2) {           // resident Z_I, two Z_J chunks per stage
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          const uint32_t st = ring_addr + s * P2_STAGE;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint64_t a_d = make_kmajor_sw128_desc(zi_addr + (kc + h) * P2_CHUNK);
            const uint64_t b_d = make_kmajor_sw128_desc(st + h * P2_CHUNK);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
              mma_tf32_ss_pair_elect(d_tmem, a_d + adv, b_d + adv, idesc_s, (kc + h > 0 || k > 0) ? 1u : 0u);
            }
          }
          mma_commit_pair_elect(&ctl->empty[s]);
          next();
        }
        for (int kc = nres; kc < kchunks; ++kc) {        // streamed Z_I chunk + Z_J chunk
          mbar_wait(&ctl->full[s], ph);
          tc_fence_after();
          const uint32_t st = ring_addr + s * P2_STAGE;
          const uint64_t a_d = make_kmajor_sw128_desc(st);
          const uint64_t b_d = make_kmajor_sw128_desc(st + P2_CHUNK);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
            mma_tf32_ss_pair_elect(d_tmem, a_d + adv, b_d + adv, idesc_s, (kc > 0 || k > 0) ? 1u : 0u);
          }
          mma_commit_pair_elect(&ctl->empty[s]);
          next();
        }
        mma_commit_pair_elect(&ctl->s_full[b]);
      };
      auto issue_P = [&](int J) {
        mbar_wait_cluster(&ctl->g_full, (uint32_t)(J & 1));
        tc_fence_after();
        for (int t = 0; t < ntile; ++t)
          for (int a4 = 0; a4 < BN / BK; ++a4) {
            mbar_wait(&ctl->full[s], ph);
            tc_fence_after();
            const uint64_t a_d = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
            const uint64_t b_d = make_kmajor_sw128_desc(g_addr + a4 * P2_CHUNK);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);
              mma_tf32_ss_pair_elect(tmem_dz + t * BN, a_d + adv, b_d + adv, idesc_p,
                                     (J > 0 || a4 > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit_pair_elect(&ctl->empty[s]);
            next();
          }
        mma_commit_pair_elect(&ctl->g_empty);
      };
      issue_S(0);
      for (int J = 0; J < nJ; ++J) {
        if (J + 1 < nJ) issue_S(J + 1);
        issue_P(J);
      }
      mma_commit_pair_elect(&ctl->dz_full);
    }
  } else {
    // ===================== epilogue (both CTAs): S -> G for this CTA's 64 rows =====================
    const int ew = warp - 2;
    const int lg = warp & 3;                 // TMEM lane group of this warp
    const int ch = ew >> 2;                  // which 32 of the S stage's 64 TMEM columns
    const int et = ew * 32 + lane;
    const int tl = lg * 32 + lane;           // TMEM lane
    const int r = tl & 63;                   // row of this CTA's 64-row slice
    const int jh = tl >> 6;                  // lanes 64..127 hold columns 64..127 of the same rows (2x2 layout)
    const int j0 = jh * 64 + ch * 32;        // first of this thread's 32 columns inside the J tile
    const int gi = row_base + (int)rank * 64 + r;

    const float sigma0 = p.stats[EDRL_MMD_STAT_SIGMA0];
    const float cval = p.stats[EDRL_MMD_STAT_C];
    float sig_last = sigma0;
    for (int k = 0; k < p.num - 1; ++k) sig_last *= p.mul;
    const float negc_last = -LOG2E / sig_last;
    if (!FAST && et == 0) fill_generic_coefs(ctl->negc, ctl->w, sigma0, p.mul, p.num);

    const float ri = (gi < p.n_pad) ? (float)p.racc[gi] : 0.f;
    const float ai = (gi < p.n_pad) ? p.a[gi] : 0.f;
    const float nai_sig = -ai / sigma0;
    const uint32_t g_full_leader = mapa_u32(smem_u32(&ctl->g_full), 0);
    float rowsum = 0.f;

    for (int J = 0; J < nJ; ++J) {
      const int b = J & 1;
      const uint32_t u = (uint32_t)(J >> 1);
      if (et < BN) {
        const int gj = (J0 + J) * BN + et;
        ctl->colinfo[b][et] = make_float2((float)p.racc[gj], p.a[gj]);
      }
      named_barrier_sync(1, BWD_EPI_THREADS);
      mbar_wait(&ctl->s_full[b], u & 1);
      tc_fence_after();
      mbar_wait(&ctl->g_empty, (uint32_t)((J & 1) ^ 1));     // P(J-1) has consumed the G buffer
      uint32_t v[32];
      tmem_ld_32x32(tmem_s + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * 64 + ch * 32), v);
      tmem_ld_wait();
      float g[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float2 ci = ctl->colinfo[b][j0 + j];
        const float Lraw = fmaf(-2.f, __uint_as_float(v[j]), ri + ci.x);
        const float L = fmaxf(Lraw, 0.f);
        float K, Q;
        kernel_terms<FAST>(L, negc_last, ctl->negc, ctl->w, p.num, K, Q);
        float gv = fmaf(ci.y * Q, nai_sig, (ci.y != 0.f) ? cval : 0.f);   // a_j == 0 <=> padded column
        gv = (Lraw >= 0.f) ? gv : 0.f;
        const float gh = to_tf32(gv);
        g[j] = gh;
        rowsum += gh;
      }
      // this thread's 32 values of row r go to K-atom (j0 / 32) of the B operand, 128-byte swizzle
      uint8_t *atom = g_smem + (j0 >> 5) * P2_CHUNK + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4)
        *reinterpret_cast<float4 *>(atom + ((q4 ^ (r & 7)) << 4)) =
            make_float4(g[q4 * 4 + 0], g[q4 * 4 + 1], g[q4 * 4 + 2], g[q4 * 4 + 3]);
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(g_full_leader);
        mbar_arrive_cluster(mapa_u32(smem_u32(&ctl->s_empty[b]), 0));
      }
    }
    // ---- row sums of G: 4 partials per row (2 column halves x 2 lane halves) -> all 128 rows in both CTAs ----
    named_barrier_sync(1, BWD_EPI_THREADS);                  // everyone is done with colinfo
    float *part = reinterpret_cast<float *>(&ctl->colinfo[0][0]);   // [4][64]
    float *rs_all = reinterpret_cast<float *>(&ctl->colinfo[1][0]); // [128]
    part[(jh * 2 + ch) * 64 + r] = rowsum;
    named_barrier_sync(1, BWD_EPI_THREADS);
    if (et < 64) {
      const float tot = (part[et] + part[64 + et]) + (part[128 + et] + part[192 + et]);
      rs_all[rank * 64 + et] = tot;
      st_cluster_f32(mapa_u32(smem_u32(&rs_all[rank * 64 + et]), rank ^ 1u), tot);
    }
  }
  __syncwarp();
  cluster_sync_all();                                        // row sums exchanged (DSMEM writes visible)

  if (warp >= 2) {
    // ===================== write-out: dZ[i, f] = coef (rowsum_i z_i[f] - dZ^T[f, i]) =====================
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int ch = ew >> 2;
    const int tl = lg * 32 + lane;                           // feature lane of the M-tile
    const float *rs_all = reinterpret_cast<const float *>(&ctl->colinfo[1][0]);
    const float M = p.stats[EDRL_MMD_STAT_M];
    const float sgn = (M > 0.f) ? 1.f : ((M < 0.f) ? -1.f : 0.f);
    const float coef = 4.f * sgn * p.grad_out[0];
    int rows_here = rng_count - lpanel * BM;
    if (rows_here > BM) rows_here = BM;
    if (p.n - row_base < rows_here) rows_here = p.n - row_base;
    mbar_wait(&ctl->dz_full, 0);
    tc_fence_after();
    for (int t = 0; t < ntile; ++t) {
      const int f = f0 + t * 256 + (int)rank * 128 + tl;
      const bool f_ok = f < p.d;
#pragma unroll 1
      for (int c2 = 0; c2 < 2; ++c2) {
        const int i0 = ch * 64 + c2 * 32;
        if (i0 >= rows_here) break;                          // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tmem_dz + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * BN + i0), v);
        tmem_ld_wait();
        if (f_ok) {
          const float *zc = p.zhi + (size_t)(row_base + i0) * p.d_pad + f;
          float *oc = p.dz + ((size_t)blockIdx.z * (size_t)(p.row_count + p.row_count2) + out_row0 + i0) * p.d + f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (i0 + j < rows_here)
              oc[(size_t)j * p.d] = coef * fmaf(rs_all[i0 + j], zc[(size_t)j * p.d_pad], -__uint_as_float(v[j]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                        // nobody exits while the peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ----------------------------------------------------------------------------- host side
// need_zt: also write the fp32 transposed copy Z^T (the separate backward and the TF32 sweep read it; the binary16
// sweeps read Z^T16 instead and skip these n d 4 bytes)
static int run_prep(const float *X, const float *Y, int n_s, int n_t, int d, const Layout &L, uint8_t *ws,
                    cudaStream_t st, bool need_zt = true) {
  EDRL_CUDA_OK(cudaMemsetAsync(ws + L.off_acc, 0, L.zero_bytes, st));
  const int n = L.n;
  double *acc = reinterpret_cast<double *>(ws + L.off_acc);
  double *colsum = reinterpret_cast<double *>(ws + L.off_colsum);
  double *racc = reinterpret_cast<double *>(ws + L.off_r);
  float *a = reinterpret_cast<float *>(ws + L.off_a);
  float *zhi = reinterpret_cast<float *>(ws + L.off_zhi);
  float *zthi = need_zt ? reinterpret_cast<float *>(ws + L.off_zthi) : nullptr;
  float *zlo = reinterpret_cast<float *>(ws + L.off_zlo);
  float *ztlo = reinterpret_cast<float *>(ws + L.off_ztlo);
  unsigned *colmax = L.h16 ? reinterpret_cast<unsigned *>(ws + L.off_colmax) : nullptr;
  if (d % 4 == 0 && (((uintptr_t)X | (uintptr_t)Y) & 15) == 0) {
    dim3 g1((d / 4 + 127) / 128, (n + 63) / 64);
    prep_colsum_vec4_kernel<<<g1, dim3(128, 4), 0, st>>>(X, Y, n_s, n, d, colsum, colmax);
  } else {
    dim3 g1((d + 127) / 128, (n + 63) / 64);
    prep_colsum_kernel<<<g1, 128, 0, st>>>(X, Y, n_s, n, d, colsum, colmax);
  }
  EDRL_LAUNCHED();
  const int rb = L.n_pad / 32;
  int nsplit = (2368 + rb - 1) / rb;                  // >= 16 blocks of 256 threads per SM when the matrix allows it
  if (nsplit < 1) nsplit = 1;
  if (nsplit > L.d_pad / 32) nsplit = L.d_pad / 32;
  dim3 g2(rb, nsplit), b2(32, 8);
  if (L.split3)
    prep_center_kernel<true><<<g2, b2, 0, st>>>(X, Y, n_s, n_t, d, L.n_pad, L.d_pad, colsum, zhi, zthi, zlo, ztlo,
                                                 racc, a, acc, reinterpret_cast<double *>(ws + L.off_colsum_hi));
  else if (L.h16)
    prep_center_kernel<false, true><<<g2, b2, 0, st>>>(
        X, Y, n_s, n_t, d, L.n_pad, L.d_pad, colsum, zhi, zthi, zlo, ztlo, racc, a, acc,
        reinterpret_cast<double *>(ws + L.off_colsum_hi), reinterpret_cast<const unsigned *>(ws + L.off_colmax),
        reinterpret_cast<int *>(ws + L.off_fscale), reinterpret_cast<__half *>(ws + L.off_zt16),
        L.s16 ? reinterpret_cast<__half *>(ws + L.off_z16) : nullptr);
  else
    prep_center_kernel<false><<<g2, b2, 0, st>>>(X, Y, n_s, n_t, d, L.n_pad, L.d_pad, colsum, zhi, zthi, zlo, ztlo,
                                                  racc, a, acc, reinterpret_cast<double *>(ws + L.off_colsum_hi));
  EDRL_LAUNCHED();
  return 0;
}

static int check_common(int n_s, int n_t, int d, float mul, int num, const Layout &L, const void *ws,
                        size_t ws_bytes) {
  EDRL_CHECK_ARG(n_s > 0 && n_t > 0 && d > 0, "MK_MMD: empty input (n_s=%d n_t=%d d=%d)", n_s, n_t, d);
  EDRL_CHECK_ARG((long long)n_s + n_t >= 2, "MK_MMD: needs at least two samples");
  EDRL_CHECK_ARG(num >= 1 && num <= MAX_KERNELS, "MK_MMD: kernel_num must be in [1, %d], got %d", MAX_KERNELS, num);
  EDRL_CHECK_ARG(mul > 0.f, "MK_MMD: kernel_mul must be positive");
  EDRL_CHECK_ARG(ws != nullptr && ws_bytes >= L.total, "MK_MMD: workspace too small (%zu < %zu)", ws_bytes, L.total);
  EDRL_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 1023) == 0, "MK_MMD: workspace must be 1024-byte aligned");
  return 0;
}

template <bool SPLIT3, int MODE, bool FAST>
static int launch_fwd_t(const CUtensorMap &tm_hi, const CUtensorMap &tm_lo, const FwdParams &p, int grid,
                        cudaStream_t st) {
  using Cfg = FwdCfg<SPLIT3>;
  auto kern = mmd_fwd_kernel<SPLIT3, MODE, FAST>;
  // per launch (cheap): the attribute is per device, a process may drive several
  EDRL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  kern<<<grid, FWD_THREADS, Cfg::SMEM_BYTES, st>>>(tm_hi, tm_lo, p);
  EDRL_LAUNCHED();
  return 0;
}

template <int MODE>
static int launch_fwd(bool split3, bool fast, const CUtensorMap &tm_hi, const CUtensorMap &tm_lo, const FwdParams &p,
                      int grid, cudaStream_t st) {
  if (split3) {
    if (fast) return launch_fwd_t<true, MODE, true>(tm_hi, tm_lo, p, grid, st);
    return launch_fwd_t<true, MODE, false>(tm_hi, tm_lo, p, grid, st);
  }
  if (fast) return launch_fwd_t<false, MODE, true>(tm_hi, tm_lo, p, grid, st);
  return launch_fwd_t<false, MODE, false>(tm_hi, tm_lo, p, grid, st);
}

static int forward_impl(int mode, const float *X, const float *Y, int n_s, int n_t, int d, float mul, int num,
                        int flags, int tile_rank, int tile_world, float *loss, float *stats, double *partial,
                        float *out, void *workspace, size_t ws_bytes, void *stream) {
  Layout L = make_layout(n_s, n_t, d, flags);
  if (int rc = check_common(n_s, n_t, d, mul, num, L, workspace, ws_bytes)) return rc;
  EDRL_CHECK_ARG(tile_world >= 1 && tile_rank >= 0 && tile_rank < tile_world, "MK_MMD: bad tile shard %d/%d",
                 tile_rank, tile_world);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  if (int rc = run_prep(X, Y, n_s, n_t, d, L, ws, st)) return rc;

  CUtensorMap tm_hi, tm_lo;
  if (int rc = make_tmap_2d_f32(&tm_hi, ws + L.off_zhi, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, BM, BK)) return rc;
  if (L.split3) {
    if (int rc = make_tmap_2d_f32(&tm_lo, ws + L.off_zlo, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, BM, BK)) return rc;
  } else {
    tm_lo = tm_hi;
  }
  FwdParams p;
  p.n = L.n; p.n_s = n_s; p.n_t = n_t; p.n_pad = L.n_pad; p.d_pad = L.d_pad;
  p.nb = L.n_pad / BM; p.kchunks = L.d_pad / BK; p.num = num; p.mul = mul;
  p.tiles_total = (long long)p.nb * (p.nb + 1) / 2;
  p.tile_rank = tile_rank; p.tile_world = tile_world;
  p.racc = reinterpret_cast<const double *>(ws + L.off_r);
  p.a = reinterpret_cast<const float *>(ws + L.off_a);
  p.acc = reinterpret_cast<double *>(ws + L.off_acc);
  p.ticket = reinterpret_cast<unsigned *>(ws + L.off_acc + 128);
  p.loss = loss; p.stats = stats; p.partial = partial; p.out = out;
  const long long q_total = (p.tiles_total - tile_rank + tile_world - 1) / tile_world;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  int grid = (int)((q_total < sms) ? (q_total > 0 ? q_total : 1) : sms);
  const bool fast = (mul == 2.0f && num == 5);
  static const bool legacy = (getenv("EDRL_MMD_FWD_LEGACY") != nullptr);   // A/B switch for profiling
  if (mode == MODE_LOSS && !L.split3 && !legacy) {
    // TF32 loss path: persistent CTA pairs over 256 x 256 tiles
    const int nb2 = L.n_pad / F2_TILE;
    const long long tiles2 = (long long)nb2 * (nb2 + 1) / 2;
    const long long q2 = (tiles2 - tile_rank + tile_world - 1) / tile_world;
    int pairs = sms / 2;
    if (q2 < pairs) pairs = (int)(q2 > 0 ? q2 : 1);
    auto kern = fast ? mmd_fwd_pair_kernel<true> : mmd_fwd_pair_kernel<false>;
    // per launch (cheap): the attribute is per device, a process may drive several
  EDRL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, F2_SMEM_BYTES));
    kern<<<2 * pairs, F2_THREADS, F2_SMEM_BYTES, st>>>(tm_hi, p);
    EDRL_LAUNCHED();
    return 0;
  }
  if (mode == MODE_LOSS) return launch_fwd<MODE_LOSS>(L.split3, fast, tm_hi, tm_lo, p, grid, st);
  if (mode == MODE_KMAT) return launch_fwd<MODE_KMAT>(L.split3, fast, tm_hi, tm_lo, p, grid, st);
  return launch_fwd<MODE_GRAM>(L.split3, fast, tm_hi, tm_lo, p, grid, st);
}

template <bool SPLIT3, bool FAST>
static int launch_bwd_t(const CUtensorMap &a, const CUtensorMap &b, const CUtensorMap &c, const CUtensorMap &d4,
                        const BwdParams &p, dim3 grid, cudaStream_t st) {
  using Cfg = BwdCfg<SPLIT3>;
  auto kern = mmd_bwd_kernel<SPLIT3, FAST>;
  // per launch (cheap): the attribute is per device, a process may drive several
  EDRL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  kern<<<grid, BWD_THREADS, Cfg::SMEM_BYTES, st>>>(a, b, c, d4, p);
  EDRL_LAUNCHED();
  return 0;
}



// ----------------------------------------------------------------------------- K3q: CTA-pair sweep, 256-column S tiles
// Same decomposition as mmd_bwd_pair_kernel (S phase -> G -> transposed P phase over a 2-CTA cluster), but the S
// phase works on TWO column tiles at once: tcgen05 M = 128 (64 rows of the panel per CTA), N = 256 (128 rows of Z_J
// per CTA).  With 64 A rows per CTA the N = 128 S phase re-read 4 KiB of shared memory per 32-clk MMA (the whole
// 128 B/clk port); N = 256 reads 6 KiB per 64 clk, the Z_I chunk is fetched once per 256 columns instead of per 128,
// and the MMAs are twice as long (half the issue slots).  Z_I is streamed (two 32-column chunks per ring stage), the
// ring has 9 stages of 16 KiB, G is 64 rows x 256 columns (64 KiB) per CTA.
constexpr int Q_GROUP = 256;                   // columns per S group
constexpr int Q_G_BYTES = 8 * P2_CHUNK;        // 64 rows x 256 columns j
constexpr int Q_CTRL_BYTES = 8192;
constexpr int SW_EPI_WARPS = 16;               // 4 per TMEM lane group: one 32-column chunk of the S stage each
constexpr int SW_EPI_THREADS = SW_EPI_WARPS * 32;
constexpr int SW_THREADS = 64 + SW_EPI_THREADS;
constexpr int SW_MAX_SPLIT = 8;                // column slabs of a split virtual panel (make_plan)

// MODE 0: TF32 everywhere.  1 (EDRL_MMD_TF32H): binary16 P phase.  2 (EDRL_MMD_F16S): the S phase too reads scaled
// binary16 operands (Z16, kind::f16): a ring stage then holds 64 feature columns instead of 32.
// The binary16 modes keep TWO G buffers (32 KiB each), so the epilogue of group g+1 overlaps the P phase of group g.
template <int MODE>
struct SweepCfg {
  static constexpr bool H16 = MODE >= 1;
  static constexpr bool S16 = MODE == 2;
  static constexpr int S_COLS = S16 ? 64 : BK;                         // feature columns per 128-byte row of an S operand
  static constexpr int G_BYTES = H16 ? Q_G_BYTES / 2 : Q_G_BYTES;      // one G buffer: 64 rows x 256 columns
  static constexpr int G_BUFS = H16 ? 2 : 1;
  static constexpr int STAGES = 9;
  static constexpr int SMEM_BYTES = G_BUFS * G_BYTES + STAGES * P2_STAGE + Q_CTRL_BYTES;
  static constexpr int P_ATOMS = H16 ? Q_GROUP / 64 : Q_GROUP / BK;     // K atoms (128-byte rows) per column group
  static constexpr int P_ATOM_COLS = H16 ? 64 : BK;
};

struct SweepCtrl {
  uint64_t full[12];              // leader CTA only
  uint64_t empty[12];             // per CTA (multicast commit)
  uint64_t s_full[2];             // per CTA (multicast commit)
  uint64_t s_empty[2];            // leader, one arrival per epilogue warp of the pair
  uint64_t g_full[2];             // leader, one arrival per epilogue warp of the pair
  uint64_t g_empty[2];            // per CTA (multicast commit)
  uint64_t dz_full;               // per CTA (multicast commit): the item's dZ^T accumulators are complete
  uint64_t dz_empty;              // leader, one arrival per epilogue warp of the pair: ... and have been read out
  uint64_t g_ready[2];            // quad kernel, per CTA: the other pair's CTA for the same rows has written its G tile
  uint64_t g_copied[2];           // quad kernel, per CTA: the other pair's CTA has copied our G tile out
  uint32_t tmem_base;
  uint32_t pad;
  alignas(16) float col_r[2][Q_GROUP];   // r_j per S stage (read as float4)
  alignas(16) float col_a[2][Q_GROUP];   // a_j per S stage
  float negc[MAX_KERNELS];
  float w[MAX_KERNELS];
  double red[SW_EPI_WARPS][2];
  float part[8][64];              // row-sum partials of an item (2 lane halves x 4 column chunks per row)
};
static_assert(sizeof(SweepCtrl) <= Q_CTRL_BYTES, "SweepCtrl does not fit its smem slot");
static_assert(SweepCfg<0>::SMEM_BYTES <= 232448 && SweepCfg<1>::SMEM_BYTES <= 232448, "smem budget");

// One work item of the sweep (make_plan): a 128-row panel x 512 feature columns, over the column groups
// [g_begin, g_end) of 256 columns each; split panels write one partial output per slab.
struct SweepItem {
  int ypass, slab, g_begin, ng, row_base, out_row0, rng_begin, rng_count, rows_here, f0, ntile;
};
__device__ __forceinline__ SweepItem sweep_item(const BwdParams &p, int item, int quad_pair = -1) {
  SweepItem it;
  const int nG_all = p.nb / 2;                            // n_pad is a multiple of 256
  int vp, g_end;
  if (item < p.full_items) {
    vp = item; it.slab = 0; it.g_begin = 0; g_end = nG_all;
  } else {
    const int q = item - p.full_items;
    vp = p.full_items + q / p.split;
    it.slab = q % p.split;
    it.g_begin = (int)((long long)it.slab * nG_all / p.split);
    g_end = (int)((long long)(it.slab + 1) * nG_all / p.split);
  }
  it.ng = g_end - it.g_begin;
  it.ypass = vp / p.panels;
  const int panel = vp - it.ypass * p.panels;
  const int np1 = (p.row_count + BM - 1) / BM;
  const bool second = panel >= np1;
  const int lpanel = second ? panel - np1 : panel;
  it.rng_begin = second ? p.row_begin2 : p.row_begin;
  it.rng_count = second ? p.row_count2 : p.row_count;
  it.out_row0 = (second ? p.row_count : 0) + lpanel * BM;
  it.row_base = it.rng_begin + lpanel * BM;
  // pair kernel: a feature pass is 512 columns; quad kernel: 1024, of which pair quad_pair takes one half
  it.f0 = (quad_pair < 0) ? it.ypass * P2_FEATS : (2 * it.ypass + quad_pair) * P2_FEATS;
  const int left = p.d_pad - it.f0;
  it.ntile = left > 256 ? 2 : (left > 0 ? 1 : 0);
  int rows_here = it.rng_count - lpanel * BM;
  if (rows_here > BM) rows_here = BM;
  if (p.n - it.row_base < rows_here) rows_here = p.n - it.row_base;
  it.rows_here = rows_here;
  return it;
}

// Persistent: CTA pair c walks the work items c, c + pairs, c + 2 pairs, ... ; the three roles (TMA producer, MMA
// issuer, epilogue) each loop over the same item sequence, so the loads and the S phase of the next item run while the
// epilogue warps still write the previous item out.
template <bool FAST, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(SW_THREADS, 1)
mmd_sweep256_kernel(const __grid_constant__ CUtensorMap tm_z64, const __grid_constant__ CUtensorMap tm_z128,
                    const __grid_constant__ CUtensorMap tm_zt, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using Cfg = SweepCfg<MODE>;
  constexpr bool H16 = Cfg::H16;
  constexpr bool S16 = Cfg::S16;
  constexpr int Q_STAGES = Cfg::STAGES;
  constexpr int GB = Cfg::G_BUFS;
  uint8_t *g_smem = smem;
  uint8_t *ring = g_smem + GB * Cfg::G_BYTES;
  SweepCtrl *ctl = reinterpret_cast<SweepCtrl *>(ring + Q_STAGES * P2_STAGE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int kchunks = S16 ? p.d_pad / 64 : p.kchunks;     // 128-byte K chunks of an S operand row; even

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (threadIdx.x == 0) {
    for (int s = 0; s < Q_STAGES; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->s_full[s], 1);
      mbar_init(&ctl->s_empty[s], 2 * SW_EPI_WARPS);
      mbar_init(&ctl->g_full[s], 2 * SW_EPI_WARPS);
      mbar_init(&ctl->g_empty[s], 1);
    }
    mbar_init(&ctl->dz_full, 1);
    mbar_init(&ctl->dz_empty, 2 * SW_EPI_WARPS);
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc_pair(&ctl->tmem_base, 512);
    tmem_relinquish_pair();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_z64);
    tma_prefetch_desc(&tm_z128);
    tma_prefetch_desc(&tm_zt);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t tmem_dz = tmem_base;                     // columns [0, 256): two M-tiles of dZ^T
  const uint32_t tmem_s = tmem_base + 256;                // two S stages of 128 columns (64 rows x 256)

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, warp-converged issue) =====================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t full0 = mapa_u32(smem_u32(&ctl->full[0]), 0);
    auto acquire = [&]() -> uint8_t * {
      mbar_wait(&ctl->empty[s], ph ^ 1);
      mbar_expect_tx_elect(&ctl->full[s], 2 * P2_STAGE, leader ? 1u : 0u);
      return ring + s * P2_STAGE;
    };
    auto next = [&]() {
      if (++s == Q_STAGES) {
        s = 0;
        ph ^= 1;
      }
    };
    for (int item = pair; item < p.items; item += npairs) {
      const SweepItem it = sweep_item(p, item);
      const int irow = it.row_base + (int)rank * 64;
      auto load_S = [&](int g) {
        const int jrow = (it.g_begin + g) * Q_GROUP + (int)rank * 128;
        for (int kc = 0; kc < kchunks; kc += 2) {
          {                                                   // two chunks of this CTA's 64 panel rows
            uint8_t *st = acquire();
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            tma_load_2d_pair_elect(st, &tm_z64, bar, kc * Cfg::S_COLS, irow);
            tma_load_2d_pair_elect(st + P2_CHUNK, &tm_z64, bar, (kc + 1) * Cfg::S_COLS, irow);
            next();
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {                       // one chunk of this CTA's 128 rows of the column group each
            uint8_t *st = acquire();
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            tma_load_2d_pair_elect(st, &tm_z128, bar, (kc + h) * Cfg::S_COLS, jrow);
            next();
          }
        }
      };
      auto load_P = [&](int g) {
        for (int t = 0; t < it.ntile; ++t)
          for (int a8 = 0; a8 < Cfg::P_ATOMS; ++a8) {
            uint8_t *st = acquire();
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            tma_load_2d_pair_elect(st, &tm_zt, bar, (it.g_begin + g) * Q_GROUP + a8 * Cfg::P_ATOM_COLS,
                                   it.f0 + t * 256 + (int)rank * 128);
            next();
          }
      };
      load_S(0);
      for (int g = 0; g < it.ng; ++g) {
        if (g + 1 < it.ng) load_S(g + 1);
        load_P(g);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, warp-converged issue) =====================
    if (leader) {
      constexpr uint32_t idesc_s = S16 ? make_idesc_f16(128, Q_GROUP) : make_idesc_tf32(128, Q_GROUP);   // 64 panel rows / 128 column rows per CTA
      constexpr uint32_t idesc_p = H16 ? make_idesc_f16(256, BN) : make_idesc_tf32(256, BN);   // 128 features / 64 rows per CTA
      int s = 0;
      uint32_t ph = 0;
      auto next = [&]() {
        if (++s == Q_STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      const uint32_t ring_addr = smem_u32(ring);
      const uint32_t g_addr = smem_u32(g_smem);
      int gc = 0;                                          // running group counter over all items of this pair
      int itn = 0;                                         // running item counter
      for (int item = pair; item < p.items; item += npairs, ++itn) {
        const SweepItem it = sweep_item(p, item);
        auto issue_S = [&](int c) {                        // c: running index of the group
          const int b = c & 1;
          const uint32_t u = (uint32_t)(c >> 1);
          mbar_wait_cluster(&ctl->s_empty[b], (u & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_s + b * 128;
          for (int kc = 0; kc < kchunks; kc += 2) {
            mbar_wait(&ctl->full[s], ph);                    // the Z_I stage (two chunks)
            tc_fence_after();
            const int sa = s;
            const uint32_t a_st = ring_addr + s * P2_STAGE;
            next();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mbar_wait(&ctl->full[s], ph);                  // the Z_J chunk
              tc_fence_after();
              const uint64_t a_d = make_kmajor_sw128_desc(a_st + h * P2_CHUNK);
              const uint64_t b_d = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
#pragma unroll
              for (int k = 0; k < 4; ++k) {                 // 32-byte K steps: 8 TF32 or 16 binary16 values
                const uint64_t adv = (uint64_t)(k * 2);
                if (S16)
                  mma_f16_ss_pair_elect(d_tmem, a_d + adv, b_d + adv, idesc_s, (kc + h > 0 || k > 0) ? 1u : 0u);
                else
                  mma_tf32_ss_pair_elect(d_tmem, a_d + adv, b_d + adv, idesc_s, (kc + h > 0 || k > 0) ? 1u : 0u);
              }
              if (h == 1) mma_commit_pair_elect(&ctl->empty[sa]);
              mma_commit_pair_elect(&ctl->empty[s]);
              next();
            }
          }
          mma_commit_pair_elect(&ctl->s_full[b]);
        };
        auto issue_P = [&](int g, int c) {                 // g: group inside the item, c: running index
          const int gb = c % GB;
          const uint32_t gu = (uint32_t)(c / GB);
          mbar_wait_cluster(&ctl->g_full[gb], gu & 1);
          tc_fence_after();
          for (int t = 0; t < it.ntile; ++t)
            for (int a8 = 0; a8 < Cfg::P_ATOMS; ++a8) {
              mbar_wait(&ctl->full[s], ph);
              tc_fence_after();
              const uint64_t a_d = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
              const uint64_t b_d = make_kmajor_sw128_desc(g_addr + gb * Cfg::G_BYTES + a8 * P2_CHUNK);
#pragma unroll
              for (int k = 0; k < 4; ++k) {                 // 32-byte K steps: 8 TF32 or 16 binary16 values
                const uint64_t adv = (uint64_t)(k * 2);
                if (H16)
                  mma_f16_ss_pair_elect(tmem_dz + t * BN, a_d + adv, b_d + adv, idesc_p,
                                        (g > 0 || a8 > 0 || k > 0) ? 1u : 0u);
                else
                  mma_tf32_ss_pair_elect(tmem_dz + t * BN, a_d + adv, b_d + adv, idesc_p,
                                         (g > 0 || a8 > 0 || k > 0) ? 1u : 0u);
              }
              mma_commit_pair_elect(&ctl->empty[s]);
              next();
            }
          mma_commit_pair_elect(&ctl->g_empty[gb]);
        };
        issue_S(gc);
        for (int g = 0; g < it.ng; ++g) {
          if (g + 1 < it.ng) issue_S(gc + g + 1);
          if (g == 0 && itn > 0) {                         // the previous item's dZ^T has been read out of TMEM
            mbar_wait_cluster(&ctl->dz_empty, (uint32_t)((itn - 1) & 1));
            tc_fence_after();
          }
          issue_P(g, gc + g);
        }
        gc += it.ng;
        mma_commit_pair_elect(&ctl->dz_full);
      }
    }
  } else {
    // ===================== epilogue (both CTAs): S -> G for this CTA's 64 rows x 256 columns =====================
    // 16 warps: warp % 4 fixes the TMEM lane group, cq = which 32 of the S stage's 128 TMEM columns.  A thread owns one
    // row and 32 columns per group; the element math runs on packed fp32 pairs (FFMA2 / FMUL2 / FADD2).
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int cq = ew >> 2;
    const int et = ew * 32 + lane;
    const int tl = lg * 32 + lane;           // TMEM lane
    const int r = tl & 63;                   // row of this CTA's 64-row slice
    const int jh = tl >> 6;                  // lanes 64..127 hold columns 128..255 of the same rows (2x2 layout)
    const int j0 = jh * 128 + cq * 32;       // first of this thread's 32 columns inside a group

    const double sum_r = p.acc[2];
    const float sigma0 = (float)bandwidth_sigma0(sum_r, p.n, p.mul, p.num);
    float sig_last = sigma0;
    for (int k = 0; k < p.num - 1; ++k) sig_last *= p.mul;
    const float negc_last = -LOG2E / sig_last;
    if (!FAST && et == 0) fill_generic_coefs(ctl->negc, ctl->w, sigma0, p.mul, p.num);
    const uint32_t s_empty_leader0 = mapa_u32(smem_u32(&ctl->s_empty[0]), 0);
    const uint32_t g_full_leader0 = mapa_u32(smem_u32(&ctl->g_full[0]), 0);
    const uint32_t dz_empty_leader = mapa_u32(smem_u32(&ctl->dz_empty), 0);
    // H16: |G'| <= (sum_k mul^-k) / (sigma_0 min(n_s, n_t)^2); scale by 2^eg so that it stays below 2^14
    float gs = 1.f, gs_inv = 1.f;
    if (H16) {
      float qmax = 0.f, wk = 1.f;
      for (int k = 0; k < p.num; ++k) {
        qmax += wk;
        wk /= p.mul;
      }
      const float nmin = (float)min(p.n_s, p.n_t);
      int ex = 0;
      frexpf(qmax / (sigma0 * nmin * nmin), &ex);
      gs = ldexpf(1.f, 14 - ex);
      gs_inv = ldexpf(1.f, ex - 14);
    }
    // S16: the tensor core saw Z 2^e on both sides: S = 2^(2e) z_i . z_j
    const float m2s = S16 ? -ldexpf(2.f, -2 * p.fscale[p.d_pad]) : -2.f;
    double accM = 0.0, accD = 0.0;
    int gc = 0, itn = 0;

    for (int item = pair; item < p.items; item += npairs, ++itn) {
      const SweepItem it = sweep_item(p, item);
      const int gi = it.row_base + (int)rank * 64 + r;
      const float ri = (gi < p.n_pad) ? (float)p.racc[gi] : 0.f;
      const float ai = (gi < p.n_pad) ? p.a[gi] : 0.f;
      const float rc = (-ai / sigma0) * gs;                   // G'_ij 2^eg = (a_j Q_ij) rc
      const bool count_row = it.ypass == 0 && (gi - it.rng_begin) < it.rng_count && gi < p.n;
      const float ai_m = count_row ? ai : 0.f;
      float rowsum = 0.f;                                     // of the rounded G values, in units of 2^-eg
      float2 tM2 = make_float2(0.f, 0.f), tD2 = make_float2(0.f, 0.f);   // this row's forward sums over the item
      float tMs = 0.f, tDs = 0.f;                             // (generic kernel_mul / kernel_num path)

      // (r_j, a_j) of the next group: fetched one group ahead by the first 256 epilogue threads and parked in
      // registers unconverted, so that nothing waits for the load before the next group starts
      double nxt_r = 0.0;
      float nxt_a = 0.f;
      if (et < Q_GROUP) {
        nxt_r = p.racc[it.g_begin * Q_GROUP + et];
        nxt_a = p.a[it.g_begin * Q_GROUP + et];
      }

      for (int g = 0; g < it.ng; ++g, ++gc) {
        const int b = gc & 1;
        const uint32_t u = (uint32_t)(gc >> 1);
        const int gb = gc % GB;
        const uint32_t gu = (uint32_t)(gc / GB);
        if (et < Q_GROUP) {
          ctl->col_r[b][et] = (float)nxt_r;
          ctl->col_a[b][et] = nxt_a;
          if (g + 1 < it.ng) {
            nxt_r = p.racc[(it.g_begin + g + 1) * Q_GROUP + et];
            nxt_a = p.a[(it.g_begin + g + 1) * Q_GROUP + et];
          }
        }
        named_barrier_sync(1, SW_EPI_THREADS);
        mbar_wait(&ctl->s_full[b], u & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * 128 + cq * 32), v);
        tmem_ld_wait();
        // the S stage is free as soon as its values sit in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(s_empty_leader0 + 8u * (uint32_t)b);
        uint32_t gp[H16 ? 16 : 32];                            // packed binary16 pairs / TF32 words of this row's G
        const float4 *cr4 = reinterpret_cast<const float4 *>(&ctl->col_r[b][j0]);
        const float4 *ca4 = reinterpret_cast<const float4 *>(&ctl->col_a[b][j0]);
        if (FAST) {
          const float2 ri2 = make_float2(ri, ri), m2s2 = make_float2(m2s, m2s), nc2 = make_float2(negc_last, negc_last);
          const float2 half2c = make_float2(0.5f, 0.5f), rc2 = make_float2(rc, rc);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 rj = cr4[q], aj = ca4[q];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int j = q * 4 + hh * 2;
              const float2 rj2 = hh ? make_float2(rj.z, rj.w) : make_float2(rj.x, rj.y);
              const float2 aj2 = hh ? make_float2(aj.z, aj.w) : make_float2(aj.x, aj.y);
              const float2 s2 = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              const float2 lraw = fma2(m2s2, s2, add2(ri2, rj2));
              const float2 L = make_float2(fmaxf(lraw.x, 0.f), fmaxf(lraw.y, 0.f));
              const float2 t = mul2(L, nc2);
              const float2 e4 = make_float2(ex2_approx(t.x), ex2_approx(t.y));
              const float2 e3 = mul2(e4, e4);
              const float2 e2 = mul2(e3, e3);
              const float2 e1 = mul2(e2, e2);
              const float2 e0 = mul2(e1, e1);
              const float2 Q = fma2(fma2(fma2(fma2(e4, half2c, e3), half2c, e2), half2c, e1), half2c, e0);
              const float2 aQ = mul2(aj2, Q);
              const float2 K = add2(add2(add2(e0, e1), add2(e2, e3)), e4);
              tM2 = fma2(aj2, K, tM2);
              tD2 = fma2(aQ, L, tD2);
              // the clamp mask [L_raw >= 0] is not applied to G: a pair with L_raw < 0 is a numerical duplicate
              // (z_i = z_j up to rounding), whose term G_ij (z_i - z_j) vanishes whatever G_ij is
              const float2 gv = mul2(aQ, rc2);
              if (H16) {
                const uint32_t pk = pack_half2(gv.x, gv.y);
                gp[j >> 1] = pk;
                rowsum = add_half2_f32(rowsum, pk);
              } else {
                const float g0 = to_tf32(gv.x), g1 = to_tf32(gv.y);
                gp[j] = __float_as_uint(g0);
                gp[j + 1] = __float_as_uint(g1);
                rowsum += g0 + g1;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float rj = ctl->col_r[b][j0 + j], aj = ctl->col_a[b][j0 + j];
            const float Lraw = fmaf(m2s, __uint_as_float(v[j]), ri + rj);
            const float L = fmaxf(Lraw, 0.f);
            float K, Q;
            kernel_terms<false>(L, negc_last, ctl->negc, ctl->w, p.num, K, Q);
            tMs = fmaf(aj, K, tMs);
            tDs = fmaf(aj * L, Q, tDs);
            const float gv = (aj * Q) * rc;
            if (H16) {
              const __half hv = __float2half_rn(gv);
              const uint32_t hb = (uint32_t)__half_as_ushort(hv);
              if (j & 1) gp[j >> 1] |= hb << 16; else gp[j >> 1] = hb;
              rowsum += __half2float(hv);
            } else {
              const float g0 = to_tf32(gv);
              gp[j] = __float_as_uint(g0);
              rowsum += g0;
            }
          }
        }
        // ---- G row segment -> shared memory (K-major, 128-byte swizzle), once P(g - GB) has consumed the buffer ----
        mbar_wait(&ctl->g_empty[gb], (gu & 1) ^ 1);
        uint8_t *gbuf = g_smem + gb * Cfg::G_BYTES;
        if (H16) {
          // 32 halfs = 64 bytes = four 16-byte chunks of row r in K-atom (j0 / 64)
          uint8_t *atom = gbuf + (j0 >> 6) * P2_CHUNK + (r >> 3) * 1024 + (r & 7) * 128;
          const int cb = (j0 & 63) >> 3;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4 *>(atom + (((cb + q4) ^ (r & 7)) << 4)) =
                make_uint4(gp[q4 * 4 + 0], gp[q4 * 4 + 1], gp[q4 * 4 + 2], gp[q4 * 4 + 3]);
        } else {
          uint8_t *atom = gbuf + (j0 >> 5) * P2_CHUNK + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            *reinterpret_cast<uint4 *>(atom + ((q4 ^ (r & 7)) << 4)) =
                make_uint4(gp[q4 * 4 + 0], gp[q4 * 4 + 1], gp[q4 * 4 + 2], gp[q4 * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(g_full_leader0 + 8u * (uint32_t)gb);
      }
      // ---- end of the item: forward sums, row sums of G, write-out ----
      accM += (double)(ai_m * ((tM2.x + tM2.y) + tMs));
      accD += (double)(ai_m * ((tD2.x + tD2.y) + tDs));
      ctl->part[jh * 4 + cq][r] = rowsum * gs_inv;
      named_barrier_sync(1, SW_EPI_THREADS);
      if (et < 64) {
        // 8 partials per row (2 lane halves x 4 column chunks) -> rowsum(G')_i of this item's columns, for apply_grad
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += ctl->part[k][et];
        // (only this panel's own rows: rows past the range end may belong to another panel with another split)
        if ((int)rank * 64 + et < it.rows_here)
          p.rowsum[(size_t)(it.ypass * SW_MAX_SPLIT + it.slab) * p.n_pad + it.row_base + (int)rank * 64 + et] = tot;
      }
      // U[slab][i, f] = -(G' Z)_i[f] of this item's columns (rowsum_i z_i is added by edrl_mmd_apply_grad)
      mbar_wait(&ctl->dz_full, (uint32_t)(itn & 1));
      tc_fence_after();
      const int i0 = cq * 32;
      for (int t = 0; t < it.ntile; ++t) {
        const int f = it.f0 + t * 256 + (int)rank * 128 + tl;
        const bool f_ok = f < p.d;
        const float unscale = (H16 && f_ok) ? -ldexpf(gs_inv, -p.fscale[f]) : -1.f;     // also of column f of Z^T
        uint32_t v[32];
        tmem_ld_32x32(tmem_dz + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * BN + i0), v);
        tmem_ld_wait();
        if (t == it.ntile - 1) {                              // the accumulators may be overwritten by the next item
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(dz_empty_leader);
        }
        if (f_ok && i0 < it.rows_here) {
          float *oc = p.dz + ((size_t)it.slab * (p.row_count + p.row_count2) + it.out_row0 + i0) * p.d + f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (i0 + j < it.rows_here) oc[(size_t)j * p.d] = __uint_as_float(v[j]) * unscale;
          }
        }
      }
    }
    // ---- forward sums of this CTA -> global accumulators; the last CTA finalises ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accM += __shfl_xor_sync(0xffffffffu, accM, o);
      accD += __shfl_xor_sync(0xffffffffu, accD, o);
    }
    if (lane == 0) {
      ctl->red[ew][0] = accM;
      ctl->red[ew][1] = accD;
    }
    named_barrier_sync(1, SW_EPI_THREADS);
    if (et == 0) {
      double m = 0.0, dd = 0.0;
#pragma unroll
      for (int k = 0; k < SW_EPI_WARPS; ++k) {
        m += ctl->red[k][0];
        dd += ctl->red[k][1];
      }
      atomicAdd(p.acc + 0, m);
      atomicAdd(p.acc + 1, dd);
      __threadfence();
      const unsigned t = atomicAdd(p.ticket, 1u);
      if (t == gridDim.x - 1) {
        __threadfence();
        const double Mv = atomicAdd(p.acc + 0, 0.0);
        const double Ds = atomicAdd(p.acc + 1, 0.0);
        if (p.partial) {
          p.partial[0] = Mv;
          p.partial[1] = Ds;
        }
        if (p.finalize) write_final_stats(Mv, Ds, sum_r, p.n, p.mul, p.num, p.loss, p.stats_out);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// The sweep for d_pad > 512: a cluster of FOUR CTAs = two MMA pairs on the same 128-row panel and column range.  TMEM holds
// the dZ^T accumulators of 512 feature columns per pair next to the S stages, so one pair alone has to sweep the Gram
// once per 512-column feature pass.  Here pair p accumulates feature columns [1024 q + 512 p, + 512) and the two pairs
// SHARE the S phase: pair p computes S and G only for the column groups g = p (mod 2); the other pair's epilogue warps
// (idle for that group) copy the finished G tile out of the owner's shared memory (ld.shared::cluster after the owner's
// warps arrived on an mbarrier of the copying CTA) into their own, from where their tensor cores read it as usual.  Per
// two groups a pair then issues one S phase and two P phases instead of two and two, and streams the matching operands.
template <bool FAST, int MODE>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(SW_THREADS, 1)
mmd_sweep_quad_kernel(const __grid_constant__ CUtensorMap tm_z64, const __grid_constant__ CUtensorMap tm_z128,
                    const __grid_constant__ CUtensorMap tm_zt, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using Cfg = SweepCfg<MODE>;
  constexpr bool H16 = Cfg::H16;
  constexpr bool S16 = Cfg::S16;
  constexpr int Q_STAGES = Cfg::STAGES;
  constexpr int GB = Cfg::G_BUFS;
  uint8_t *g_smem = smem;
  uint8_t *ring = g_smem + GB * Cfg::G_BYTES;
  SweepCtrl *ctl = reinterpret_cast<SweepCtrl *>(ring + Q_STAGES * P2_STAGE);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cr = cluster_ctarank();                  // 0..3
  const int pairidx = (int)(cr >> 1);                     // which MMA pair of the cluster
  const uint32_t rank = cr & 1u;                          // rank inside the pair
  const bool leader = (rank == 0);
  const uint32_t lead_cr = cr & 2u;                       // cluster rank of this pair's leader CTA
  const uint32_t other_cr = cr ^ 2u;                      // the CTA of the other pair that holds the same 64 panel rows
  const uint16_t pmask = (uint16_t)(3u << (pairidx * 2)); // commit multicast: the two CTAs of this pair
  const int pair = blockIdx.x >> 2;                       // (cluster index: the unit that walks the work list)
  const int npairs = gridDim.x >> 2;
  auto owns = [&](int g) { return (g & 1) == pairidx; };  // which pair computes S / G of column group g of an item
  const int kchunks = S16 ? p.d_pad / 64 : p.kchunks;     // 128-byte K chunks of an S operand row; even

  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (threadIdx.x == 0) {
    for (int s = 0; s < Q_STAGES; ++s) {
      mbar_init(&ctl->full[s], 1);
      mbar_init(&ctl->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->s_full[s], 1);
      mbar_init(&ctl->s_empty[s], 2 * SW_EPI_WARPS);
      mbar_init(&ctl->g_full[s], 2 * SW_EPI_WARPS);
      mbar_init(&ctl->g_empty[s], 1);
    }
    mbar_init(&ctl->dz_full, 1);
    mbar_init(&ctl->dz_empty, 2 * SW_EPI_WARPS);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ctl->g_ready[s], SW_EPI_WARPS);          // the owner's CTA has written G (one arrival per epilogue warp)
      mbar_init(&ctl->g_copied[s], SW_EPI_WARPS);         // the other pair's CTA has copied it out
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp == 1) {
    tmem_alloc_pair(&ctl->tmem_base, 512);
    tmem_relinquish_pair();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_z64);
    tma_prefetch_desc(&tm_z128);
    tma_prefetch_desc(&tm_zt);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t tmem_dz = tmem_base;                     // columns [0, 256): two M-tiles of dZ^T
  const uint32_t tmem_s = tmem_base + 256;                // two S stages of 128 columns (64 rows x 256)

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, warp-converged issue) =====================
    int s = 0;
    uint32_t ph = 0;
    const uint32_t full0 = mapa_u32(smem_u32(&ctl->full[0]), lead_cr);
    auto acquire = [&]() -> uint8_t * {
      mbar_wait(&ctl->empty[s], ph ^ 1);
      mbar_expect_tx_elect(&ctl->full[s], 2 * P2_STAGE, leader ? 1u : 0u);
      return ring + s * P2_STAGE;
    };
    auto next = [&]() {
      if (++s == Q_STAGES) {
        s = 0;
        ph ^= 1;
      }
    };
    for (int item = pair; item < p.items; item += npairs) {
      const SweepItem it = sweep_item(p, item, pairidx);
      const int irow = it.row_base + (int)rank * 64;
      auto load_S = [&](int g) {
        const int jrow = (it.g_begin + g) * Q_GROUP + (int)rank * 128;
        for (int kc = 0; kc < kchunks; kc += 2) {
          {                                                   // two chunks of this CTA's 64 panel rows
            uint8_t *st = acquire();
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            tma_load_2d_pair_elect(st, &tm_z64, bar, kc * Cfg::S_COLS, irow);
            tma_load_2d_pair_elect(st + P2_CHUNK, &tm_z64, bar, (kc + 1) * Cfg::S_COLS, irow);
            next();
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {                       // one chunk of this CTA's 128 rows of the column group each
            uint8_t *st = acquire();
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            tma_load_2d_pair_elect(st, &tm_z128, bar, (kc + h) * Cfg::S_COLS, jrow);
            next();
          }
        }
      };
      auto load_P = [&](int g) {
        for (int t = 0; t < it.ntile; ++t)
          for (int a8 = 0; a8 < Cfg::P_ATOMS; ++a8) {
            uint8_t *st = acquire();
            const uint32_t bar = full0 + 8u * (uint32_t)s;
            tma_load_2d_pair_elect(st, &tm_zt, bar, (it.g_begin + g) * Q_GROUP + a8 * Cfg::P_ATOM_COLS,
                                   it.f0 + t * 256 + (int)rank * 128);
            next();
          }
      };
      if (owns(0)) load_S(0);
      for (int g = 0; g < it.ng; ++g) {
        if (g + 1 < it.ng && owns(g + 1)) load_S(g + 1);
        load_P(g);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, warp-converged issue) =====================
    if (leader) {
      constexpr uint32_t idesc_s = S16 ? make_idesc_f16(128, Q_GROUP) : make_idesc_tf32(128, Q_GROUP);   // 64 panel rows / 128 column rows per CTA
      constexpr uint32_t idesc_p = H16 ? make_idesc_f16(256, BN) : make_idesc_tf32(256, BN);   // 128 features / 64 rows per CTA
      int s = 0;
      uint32_t ph = 0;
      auto next = [&]() {
        if (++s == Q_STAGES) {
          s = 0;
          ph ^= 1;
        }
      };
      const uint32_t ring_addr = smem_u32(ring);
      const uint32_t g_addr = smem_u32(g_smem);
      int gc = 0;                                          // running group counter over all items of this cluster
      int sc = 0;                                          // running counter of the S phases of THIS pair
      int itn = 0;                                         // running item counter
      for (int item = pair; item < p.items; item += npairs, ++itn) {
        const SweepItem it = sweep_item(p, item, pairidx);
        auto issue_S = [&](int c) {                        // c: running index of the group
          const int b = c & 1;
          const uint32_t u = (uint32_t)(c >> 1);
          mbar_wait_cluster(&ctl->s_empty[b], (u & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_s + b * 128;
          for (int kc = 0; kc < kchunks; kc += 2) {
            mbar_wait(&ctl->full[s], ph);                    // the Z_I stage (two chunks)
            tc_fence_after();
            const int sa = s;
            const uint32_t a_st = ring_addr + s * P2_STAGE;
            next();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              mbar_wait(&ctl->full[s], ph);                  // the Z_J chunk
              tc_fence_after();
              const uint64_t a_d = make_kmajor_sw128_desc(a_st + h * P2_CHUNK);
              const uint64_t b_d = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
#pragma unroll
              for (int k = 0; k < 4; ++k) {                 // 32-byte K steps: 8 TF32 or 16 binary16 values
                const uint64_t adv = (uint64_t)(k * 2);
                if (S16)
                  mma_f16_ss_pair_elect(d_tmem, a_d + adv, b_d + adv, idesc_s, (kc + h > 0 || k > 0) ? 1u : 0u);
                else
                  mma_tf32_ss_pair_elect(d_tmem, a_d + adv, b_d + adv, idesc_s, (kc + h > 0 || k > 0) ? 1u : 0u);
              }
              if (h == 1) mma_commit_mask_elect(&ctl->empty[sa], pmask);
              mma_commit_mask_elect(&ctl->empty[s], pmask);
              next();
            }
          }
          mma_commit_mask_elect(&ctl->s_full[b], pmask);
        };
        auto issue_P = [&](int g, int c) {                 // g: group inside the item, c: running index
          const int gb = c % GB;
          const uint32_t gu = (uint32_t)(c / GB);
          mbar_wait_cluster(&ctl->g_full[gb], gu & 1);
          tc_fence_after();
          for (int t = 0; t < it.ntile; ++t)
            for (int a8 = 0; a8 < Cfg::P_ATOMS; ++a8) {
              mbar_wait(&ctl->full[s], ph);
              tc_fence_after();
              const uint64_t a_d = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
              const uint64_t b_d = make_kmajor_sw128_desc(g_addr + gb * Cfg::G_BYTES + a8 * P2_CHUNK);
#pragma unroll
              for (int k = 0; k < 4; ++k) {                 // 32-byte K steps: 8 TF32 or 16 binary16 values
                const uint64_t adv = (uint64_t)(k * 2);
                if (H16)
                  mma_f16_ss_pair_elect(tmem_dz + t * BN, a_d + adv, b_d + adv, idesc_p,
                                        (g > 0 || a8 > 0 || k > 0) ? 1u : 0u);
                else
                  mma_tf32_ss_pair_elect(tmem_dz + t * BN, a_d + adv, b_d + adv, idesc_p,
                                         (g > 0 || a8 > 0 || k > 0) ? 1u : 0u);
              }
              mma_commit_mask_elect(&ctl->empty[s], pmask);
              next();
            }
          mma_commit_mask_elect(&ctl->g_empty[gb], pmask);
        };
        if (owns(0)) issue_S(sc++);
        for (int g = 0; g < it.ng; ++g) {
          if (g + 1 < it.ng && owns(g + 1)) issue_S(sc++);
          if (g == 0 && itn > 0) {                         // the previous item's dZ^T has been read out of TMEM
            mbar_wait_cluster(&ctl->dz_empty, (uint32_t)((itn - 1) & 1));
            tc_fence_after();
          }
          issue_P(g, gc + g);
        }
        gc += it.ng;
        mma_commit_mask_elect(&ctl->dz_full, pmask);
      }
    }
  } else {
    // ===================== epilogue (both CTAs): S -> G for this CTA's 64 rows x 256 columns =====================
    // 16 warps: warp % 4 fixes the TMEM lane group, cq = which 32 of the S stage's 128 TMEM columns.  A thread owns one
    // row and 32 columns per group; the element math runs on packed fp32 pairs (FFMA2 / FMUL2 / FADD2).
    const int ew = warp - 2;
    const int lg = warp & 3;
    const int cq = ew >> 2;
    const int et = ew * 32 + lane;
    const int tl = lg * 32 + lane;           // TMEM lane
    const int r = tl & 63;                   // row of this CTA's 64-row slice
    const int jh = tl >> 6;                  // lanes 64..127 hold columns 128..255 of the same rows (2x2 layout)
    const int j0 = jh * 128 + cq * 32;       // first of this thread's 32 columns inside a group

    const double sum_r = p.acc[2];
    const float sigma0 = (float)bandwidth_sigma0(sum_r, p.n, p.mul, p.num);
    float sig_last = sigma0;
    for (int k = 0; k < p.num - 1; ++k) sig_last *= p.mul;
    const float negc_last = -LOG2E / sig_last;
    if (!FAST && et == 0) fill_generic_coefs(ctl->negc, ctl->w, sigma0, p.mul, p.num);
    const uint32_t s_empty_leader0 = mapa_u32(smem_u32(&ctl->s_empty[0]), lead_cr);
    const uint32_t g_full_leader0 = mapa_u32(smem_u32(&ctl->g_full[0]), lead_cr);
    const uint32_t dz_empty_leader = mapa_u32(smem_u32(&ctl->dz_empty), lead_cr);
    const uint32_t g_ready_other0 = mapa_u32(smem_u32(&ctl->g_ready[0]), other_cr);
    const uint32_t g_copied_other0 = mapa_u32(smem_u32(&ctl->g_copied[0]), other_cr);
    const uint32_t g_smem_other = mapa_u32(smem_u32(g_smem), other_cr);
    int n_own[2] = {0, 0}, n_copy[2] = {0, 0};             // per G buffer: tiles produced here / copied in so far
    int sc = 0;                                            // running counter of the S phases of this pair
    // H16: |G'| <= (sum_k mul^-k) / (sigma_0 min(n_s, n_t)^2); scale by 2^eg so that it stays below 2^14
    float gs = 1.f, gs_inv = 1.f;
    if (H16) {
      float qmax = 0.f, wk = 1.f;
      for (int k = 0; k < p.num; ++k) {
        qmax += wk;
        wk /= p.mul;
      }
      const float nmin = (float)min(p.n_s, p.n_t);
      int ex = 0;
      frexpf(qmax / (sigma0 * nmin * nmin), &ex);
      gs = ldexpf(1.f, 14 - ex);
      gs_inv = ldexpf(1.f, ex - 14);
    }
    // S16: the tensor core saw Z 2^e on both sides: S = 2^(2e) z_i . z_j
    const float m2s = S16 ? -ldexpf(2.f, -2 * p.fscale[p.d_pad]) : -2.f;
    double accM = 0.0, accD = 0.0;
    int gc = 0, itn = 0;

    for (int item = pair; item < p.items; item += npairs, ++itn) {
      const SweepItem it = sweep_item(p, item, pairidx);
      const int gi = it.row_base + (int)rank * 64 + r;
      const float ri = (gi < p.n_pad) ? (float)p.racc[gi] : 0.f;
      const float ai = (gi < p.n_pad) ? p.a[gi] : 0.f;
      const float rc = (-ai / sigma0) * gs;                   // G'_ij 2^eg = (a_j Q_ij) rc
      const bool count_row = it.ypass == 0 && (gi - it.rng_begin) < it.rng_count && gi < p.n;
      const float ai_m = count_row ? ai : 0.f;
      float rowsum = 0.f;                                     // of the rounded G values, in units of 2^-eg
      float2 tM2 = make_float2(0.f, 0.f), tD2 = make_float2(0.f, 0.f);   // this row's forward sums over the item
      float tMs = 0.f, tDs = 0.f;                             // (generic kernel_mul / kernel_num path)

      // (r_j, a_j) of the next group: fetched one group ahead by the first 256 epilogue threads and parked in
      // registers unconverted, so that nothing waits for the load before the next group starts
      double nxt_r = 0.0;
      float nxt_a = 0.f;
      if (et < Q_GROUP && pairidx < it.ng) {               // this pair's first group is g = pairidx
        nxt_r = p.racc[(it.g_begin + pairidx) * Q_GROUP + et];
        nxt_a = p.a[(it.g_begin + pairidx) * Q_GROUP + et];
      }

      for (int g = 0; g < it.ng; ++g, ++gc) {
        const int gb = gc % GB;
        const uint32_t gu = (uint32_t)(gc / GB);
        if (!owns(g)) {
          // ---- the other pair computes this group's G: copy its tile for the same 64 rows into our buffer ----
          mbar_wait_cluster(&ctl->g_ready[gb], (uint32_t)(n_copy[gb] & 1));
          mbar_wait(&ctl->g_empty[gb], (gu & 1) ^ 1);       // our P phase has consumed the buffer's previous tile
          {
            const uint32_t src = g_smem_other + (uint32_t)(gb * Cfg::G_BYTES);
            uint8_t *dst = g_smem + gb * Cfg::G_BYTES;
#pragma unroll
            for (int c16 = 0; c16 < Cfg::G_BYTES / 16 / SW_EPI_THREADS; ++c16) {
              const int o16 = c16 * SW_EPI_THREADS + et;
              *reinterpret_cast<uint4 *>(dst + o16 * 16) = ld_cluster_v4(src + (uint32_t)o16 * 16u);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive_cluster(g_full_leader0 + 8u * (uint32_t)gb);
            mbar_arrive_cluster(g_copied_other0 + 8u * (uint32_t)gb);      // (our copy's loads have completed: their
                                                                           //  values were stored above)
          }
          ++n_copy[gb];
          continue;
        }
        const int b = sc & 1;
        const uint32_t u = (uint32_t)(sc >> 1);
        ++sc;
        if (et < Q_GROUP) {
          ctl->col_r[b][et] = (float)nxt_r;
          ctl->col_a[b][et] = nxt_a;
          if (g + 2 < it.ng) {                              // this pair's next group
            nxt_r = p.racc[(it.g_begin + g + 2) * Q_GROUP + et];
            nxt_a = p.a[(it.g_begin + g + 2) * Q_GROUP + et];
          }
        }
        named_barrier_sync(1, SW_EPI_THREADS);
        mbar_wait(&ctl->s_full[b], u & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_s + ((uint32_t)(lg * 32) << 16) + (uint32_t)(b * 128 + cq * 32), v);
        tmem_ld_wait();
        // the S stage is free as soon as its values sit in registers
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(s_empty_leader0 + 8u * (uint32_t)b);
        uint32_t gp[H16 ? 16 : 32];                            // packed binary16 pairs / TF32 words of this row's G
        const float4 *cr4 = reinterpret_cast<const float4 *>(&ctl->col_r[b][j0]);
        const float4 *ca4 = reinterpret_cast<const float4 *>(&ctl->col_a[b][j0]);
        if (FAST) {
          const float2 ri2 = make_float2(ri, ri), m2s2 = make_float2(m2s, m2s), nc2 = make_float2(negc_last, negc_last);
          const float2 half2c = make_float2(0.5f, 0.5f), rc2 = make_float2(rc, rc);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 rj = cr4[q], aj = ca4[q];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int j = q * 4 + hh * 2;
              const float2 rj2 = hh ? make_float2(rj.z, rj.w) : make_float2(rj.x, rj.y);
              const float2 aj2 = hh ? make_float2(aj.z, aj.w) : make_float2(aj.x, aj.y);
              const float2 s2 = make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
              const float2 lraw = fma2(m2s2, s2, add2(ri2, rj2));
              const float2 L = make_float2(fmaxf(lraw.x, 0.f), fmaxf(lraw.y, 0.f));
              const float2 t = mul2(L, nc2);
              const float2 e4 = make_float2(ex2_approx(t.x), ex2_approx(t.y));
              const float2 e3 = mul2(e4, e4);
              const float2 e2 = mul2(e3, e3);
              const float2 e1 = mul2(e2, e2);
              const float2 e0 = mul2(e1, e1);
              const float2 Q = fma2(fma2(fma2(fma2(e4, half2c, e3), half2c, e2), half2c, e1), half2c, e0);
              const float2 aQ = mul2(aj2, Q);
              const float2 K = add2(add2(add2(e0, e1), add2(e2, e3)), e4);
              tM2 = fma2(aj2, K, tM2);
              tD2 = fma2(aQ, L, tD2);
              // the clamp mask [L_raw >= 0] is not applied to G: a pair with L_raw < 0 is a numerical duplicate
              // (z_i = z_j up to rounding), whose term G_ij (z_i - z_j) vanishes whatever G_ij is
              const float2 gv = mul2(aQ, rc2);
              if (H16) {
                const uint32_t pk = pack_half2(gv.x, gv.y);
                gp[j >> 1] = pk;
                rowsum = add_half2_f32(rowsum, pk);
              } else {
                const float g0 = to_tf32(gv.x), g1 = to_tf32(gv.y);
                gp[j] = __float_as_uint(g0);
                gp[j + 1] = __float_as_uint(g1);
                rowsum += g0 + g1;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float rj = ctl->col_r[b][j0 + j], aj = ctl->col_a[b][j0 + j];
            const float Lraw = fmaf(m2s, __uint_as_float(v[j]), ri + rj);
            const float L = fmaxf(Lraw, 0.f);
            float K, Q;
            kernel_terms<false>(L, negc_last, ctl->negc, ctl->w, p.num, K, Q);
            tMs = fmaf(aj, K, tMs);
            tDs = fmaf(aj * L, Q, tDs);
            const float gv = (aj * Q) * rc;
            if (H16) {
              const __half hv = __float2half_rn(gv);
              const uint32_t hb = (uint32_t)__half_as_ushort(hv);
              if (j & 1) gp[j >> 1] |= hb << 16; else gp[j >> 1] = hb;
              rowsum += __half2float(hv);
            } else {
              const float g0 = to_tf32(gv);
              gp[j] = __float_as_uint(g0);
              rowsum += g0;
            }
          }
        }
        // ---- G row segment -> shared memory (K-major, 128-byte swizzle), once P(g - GB) has consumed the buffer ----
        mbar_wait(&ctl->g_empty[gb], (gu & 1) ^ 1);
        mbar_wait_cluster(&ctl->g_copied[gb], (uint32_t)((n_own[gb] & 1) ^ 1));   // ... and the other pair has copied it
        ++n_own[gb];
        uint8_t *gbuf = g_smem + gb * Cfg::G_BYTES;
        if (H16) {
          // 32 halfs = 64 bytes = four 16-byte chunks of row r in K-atom (j0 / 64)
          uint8_t *atom = gbuf + (j0 >> 6) * P2_CHUNK + (r >> 3) * 1024 + (r & 7) * 128;
          const int cb = (j0 & 63) >> 3;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            *reinterpret_cast<uint4 *>(atom + (((cb + q4) ^ (r & 7)) << 4)) =
                make_uint4(gp[q4 * 4 + 0], gp[q4 * 4 + 1], gp[q4 * 4 + 2], gp[q4 * 4 + 3]);
        } else {
          uint8_t *atom = gbuf + (j0 >> 5) * P2_CHUNK + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4)
            *reinterpret_cast<uint4 *>(atom + ((q4 ^ (r & 7)) << 4)) =
                make_uint4(gp[q4 * 4 + 0], gp[q4 * 4 + 1], gp[q4 * 4 + 2], gp[q4 * 4 + 3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(g_full_leader0 + 8u * (uint32_t)gb);
          // the tile sits in this SM's shared memory (a single point of coherence for local and DSMEM readers) before
          // the arrive is issued; a cluster-scope release (MEMBAR.ALL.GPU + ERRBAR, microseconds under TMA load)
          // made the hand-off the bottleneck
          mbar_arrive_cluster(g_ready_other0 + 8u * (uint32_t)gb);
        }
      }
      // ---- end of the item: forward sums, row sums of G, write-out ----
      accM += (double)(ai_m * ((tM2.x + tM2.y) + tMs));
      accD += (double)(ai_m * ((tD2.x + tD2.y) + tDs));
      ctl->part[jh * 4 + cq][r] = rowsum * gs_inv;
      named_barrier_sync(1, SW_EPI_THREADS);
      if (et < 64) {
        // 8 partials per row (2 lane halves x 4 column chunks) -> rowsum(G')_i of this item's columns, for apply_grad
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += ctl->part[k][et];
        // (only this panel's own rows: rows past the range end may belong to another panel with another split)
        // (each pair saw half of the column groups: the two partial sums are added in the slot zeroed by prep; two
        //  addends commute, so the result does not depend on the order)
        if ((int)rank * 64 + et < it.rows_here)
          atomicAdd(&p.rowsum[(size_t)(it.ypass * SW_MAX_SPLIT + it.slab) * p.n_pad + it.row_base + (int)rank * 64 + et],
                    tot);
      }
      // U[slab][i, f] = -(G' Z)_i[f] of this item's columns (rowsum_i z_i is added by edrl_mmd_apply_grad)
      mbar_wait(&ctl->dz_full, (uint32_t)(itn & 1));
      tc_fence_after();
      const int i0 = cq * 32;
      if (it.ntile == 0) {                                    // this pair holds no feature columns of the pass
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(dz_empty_leader);
      }
      for (int t = 0; t < it.ntile; ++t) {
        const int f = it.f0 + t * 256 + (int)rank * 128 + tl;
        const bool f_ok = f < p.d;
        const float unscale = (H16 && f_ok) ? -ldexpf(gs_inv, -p.fscale[f]) : -1.f;     // also of column f of Z^T
        uint32_t v[32];
        tmem_ld_32x32(tmem_dz + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * BN + i0), v);
        tmem_ld_wait();
        if (t == it.ntile - 1) {                              // the accumulators may be overwritten by the next item
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(dz_empty_leader);
        }
        if (f_ok && i0 < it.rows_here) {
          float *oc = p.dz + ((size_t)it.slab * (p.row_count + p.row_count2) + it.out_row0 + i0) * p.d + f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (i0 + j < it.rows_here) oc[(size_t)j * p.d] = __uint_as_float(v[j]) * unscale;
          }
        }
      }
    }
    // ---- forward sums of this CTA -> global accumulators; the last CTA finalises ----
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      accM += __shfl_xor_sync(0xffffffffu, accM, o);
      accD += __shfl_xor_sync(0xffffffffu, accD, o);
    }
    if (lane == 0) {
      ctl->red[ew][0] = accM;
      ctl->red[ew][1] = accD;
    }
    named_barrier_sync(1, SW_EPI_THREADS);
    if (et == 0) {
      double m = 0.0, dd = 0.0;
#pragma unroll
      for (int k = 0; k < SW_EPI_WARPS; ++k) {
        m += ctl->red[k][0];
        dd += ctl->red[k][1];
      }
      atomicAdd(p.acc + 0, m);
      atomicAdd(p.acc + 1, dd);
      __threadfence();
      const unsigned t = atomicAdd(p.ticket, 1u);
      if (t == gridDim.x - 1) {
        __threadfence();
        const double Mv = atomicAdd(p.acc + 0, 0.0);
        const double Ds = atomicAdd(p.acc + 1, 0.0);
        if (p.partial) {
          p.partial[0] = Mv;
          p.partial[1] = Ds;
        }
        if (p.finalize) write_final_stats(Mv, Ds, sum_r, p.n, p.mul, p.num, p.loss, p.stats_out);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// dZ[i, f] = g sign(M) 4 (U[i, f] + c (n z_i[f] - sum_j z_j[f])) -- the closed-form bandwidth term on top of the
// fused pass, on the same rounded centred operand the sweep used (sum_j z_j is its column sum, ~0 but not 0);
// one block row per output row (no per-element division), 128-bit accesses when d % 4 == 0
template <bool VEC4>
__global__ void __launch_bounds__(128)
mmd_apply_grad_kernel(const float *__restrict__ U, const float *__restrict__ zhi, const double *__restrict__ colsum_hi,
                      const float *__restrict__ stats, const float *__restrict__ grad_out, int row_begin, int row_count,
                      int row_begin2, int row_count2, int d, int d_pad, int n, int n_pad, int panels, int full_items,
                      int split, int pass_feats, const float *__restrict__ rowsum, float *__restrict__ dz) {
  const float M = stats[EDRL_MMD_STAT_M];
  const float sgn = (M > 0.f) ? 1.f : ((M < 0.f) ? -1.f : 0.f);
  const float coef = 4.f * sgn * grad_out[0];
  const float cv = stats[EDRL_MMD_STAT_C];
  const float fn = (float)n;
  const int r = blockIdx.x;                                   // output row
  const size_t gr = (r < row_count) ? (size_t)row_begin + r : (size_t)row_begin2 + (r - row_count);
  const size_t slab = (size_t)(row_count + row_count2) * d;
  const float *zr = zhi + gr * d_pad;
  const float *ur = U + (size_t)r * d;
  float *orow = dz + (size_t)r * d;
  // the sweep's work list (make_plan): virtual panel (feature pass, row panel) >= full_items was swept in `split` slabs
  const int panel = (r < row_count) ? r / BM : (row_count + BM - 1) / BM + (r - row_count) / BM;
  if (VEC4) {
    for (int f = (blockIdx.y * 128 + threadIdx.x) * 4; f < d; f += gridDim.y * 512) {
      const int yp = f / pass_feats;
      const int nslab = (yp * panels + panel < full_items) ? 1 : split;
      float4 u = *reinterpret_cast<const float4 *>(ur + f);
      float rs = rowsum[(size_t)(yp * 8) * n_pad + gr];
      for (int sl = 1; sl < nslab; ++sl) {
        const float4 w = *reinterpret_cast<const float4 *>(ur + sl * slab + f);
        u.x += w.x; u.y += w.y; u.z += w.z; u.w += w.w;
        rs += rowsum[(size_t)(yp * 8 + sl) * n_pad + gr];
      }
      const float zc = fmaf(cv, fn, rs);                      // (rowsum(G')_i + c n) z_i
      const float4 z = __ldg(reinterpret_cast<const float4 *>(zr + f));
      float4 o;
      o.x = coef * (fmaf(zc, z.x, -cv * (float)colsum_hi[f + 0]) + u.x);
      o.y = coef * (fmaf(zc, z.y, -cv * (float)colsum_hi[f + 1]) + u.y);
      o.z = coef * (fmaf(zc, z.z, -cv * (float)colsum_hi[f + 2]) + u.z);
      o.w = coef * (fmaf(zc, z.w, -cv * (float)colsum_hi[f + 3]) + u.w);
      *reinterpret_cast<float4 *>(orow + f) = o;
    }
  } else {
    for (int f = blockIdx.y * 128 + threadIdx.x; f < d; f += gridDim.y * 128) {
      const int yp = f / pass_feats;
      const int nslab = (yp * panels + panel < full_items) ? 1 : split;
      float u = ur[f];
      float rs = rowsum[(size_t)(yp * 8) * n_pad + gr];
      for (int sl = 1; sl < nslab; ++sl) {
        u += ur[sl * slab + f];
        rs += rowsum[(size_t)(yp * 8 + sl) * n_pad + gr];
      }
      dz[(size_t)r * d + f] = coef * (fmaf(fmaf(cv, fn, rs), __ldg(zr + f), -cv * (float)colsum_hi[f]) + u);
    }
  }
}

template <bool FAST, int RES>
static int launch_bwd_pair_t(const CUtensorMap &tm_z64, const CUtensorMap &tm_zt, const BwdParams &p, dim3 grid,
                             cudaStream_t st) {
  using Cfg = Bwd2Cfg<RES>;
  auto kern = mmd_bwd_pair_kernel<FAST, RES>;
  // per launch (cheap): the attribute is per device, a process may drive several
  EDRL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
  kern<<<grid, BWD_THREADS, Cfg::SMEM_BYTES, st>>>(tm_z64, tm_zt, p);
  EDRL_LAUNCHED();
  return 0;
}

template <bool FAST, int MODE = 0>
static int launch_sweep_quad_t(const CUtensorMap &tm_z64, const CUtensorMap &tm_z128, const CUtensorMap &tm_zt,
                               const BwdParams &p, dim3 grid, cudaStream_t st) {
  auto kern = mmd_sweep_quad_kernel<FAST, MODE>;
  constexpr int SMEM = SweepCfg<MODE>::SMEM_BYTES;
  EDRL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  kern<<<grid, SW_THREADS, SMEM, st>>>(tm_z64, tm_z128, tm_zt, p);
  EDRL_LAUNCHED();
  return 0;
}

template <bool FAST, int MODE = 0>
static int launch_sweep256_t(const CUtensorMap &tm_z64, const CUtensorMap &tm_z128, const CUtensorMap &tm_zt,
                             const BwdParams &p, dim3 grid, cudaStream_t st) {
  auto kern = mmd_sweep256_kernel<FAST, MODE>;
  constexpr int SMEM = SweepCfg<MODE>::SMEM_BYTES;
  EDRL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  kern<<<grid, SW_THREADS, SMEM, st>>>(tm_z64, tm_z128, tm_zt, p);
  EDRL_LAUNCHED();
  return 0;
}

// Work list of the fused sweep.  A virtual panel is (feature pass of 512 columns, 128-row panel); one CTA pair sweeps it.
// P virtual panels on C = SMs / 2 pairs run in ceil(P / C) waves, e.g. 128 panels on 74 pairs take 2 waves for 1.73 waves
// of work.  The first floor(P / C) C panels are swept whole; each of the rest is split into `split` column slabs (one
// partial output per slab, summed by edrl_mmd_apply_grad) so that the last wave is (nearly) full: 128 -> 74 whole panels
// + 54 x 4 quarter sweeps = 1.75 waves.  Small problems (P < C) are split the same way to fill the machine.
struct SweepPlan {
  int panels;       // row panels
  int vpanels;      // panels x feature passes
  int full_items;   // virtual panels swept whole
  int split;        // column slabs of every later virtual panel (1, 2, 4 or 8)
  int items;        // work items (whole or slab sweeps of a virtual panel)
  int pairs;        // persistent clusters to launch: cluster c takes items c, c + pairs, ...
  int quad;         // d_pad > 512: clusters of 4 CTAs (two MMA pairs sharing the S phase), feature passes of 1024 columns
  int pass_feats;   // feature columns per pass: 512 (pair kernel) or 1024 (quad kernel)
};

static int quad_clusters_resident() {
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = mmd_sweep_quad_kernel<true, 0>;
  constexpr int SMEM = SweepCfg<0>::SMEM_BYTES;
  int n = 0;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4 * 64);
    cfg.blockDim = dim3(SW_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 4;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
  }
  (void)cudaGetLastError();
  cached = n > 0 ? n : 0;
  return cached;
}

static SweepPlan make_plan(const Layout &L, int row_count, int row_count2, int sms_override = 0) {
  SweepPlan pl;
  pl.panels = (row_count + BM - 1) / BM + (row_count2 + BM - 1) / BM;
  static const bool no_quad = (getenv("EDRL_MMD_QUAD") != nullptr && atoi(getenv("EDRL_MMD_QUAD")) == 0);   // A/B runs
  // (up to 768 columns the second pair would hold one 256-column tile or less: two pair passes are faster)
  pl.quad = (L.d_pad > P2_FEATS + P2_FEATS / 2 && !no_quad) ? 1 : 0;
  pl.pass_feats = pl.quad ? 2 * P2_FEATS : P2_FEATS;
  const int ny = (L.d_pad + pl.pass_feats - 1) / pl.pass_feats;
  pl.vpanels = pl.panels * ny;
  int sms = sms_override > 0 ? sms_override : device_sm_count();
  if (sms <= 0) sms = 148;
  const int csz = pl.quad ? 4 : 2;
  int C = sms / csz > 0 ? sms / csz : 1;
  if (pl.quad && sms_override <= 0) {
    // 4-CTA clusters must sit inside one GPC: fewer than SMs / 4 of them are co-resident, and a persistent cluster
    // that only starts when another has finished doubles the run time -- launch exactly as many as fit
    const int fit = quad_clusters_resident();
    if (fit > 0 && fit < C) C = fit;
  }
  const int nG = L.n_pad / Q_GROUP;
  pl.full_items = (pl.vpanels / C) * C;
  const int R = pl.vpanels - pl.full_items;
  pl.split = 1;
  static const char *env = getenv("EDRL_MMD_SLABS");      // =1: never split (A/B runs)
  const int kmax = env ? atoi(env) : 8;
  if (R > 0) {
    double best = 1.0;
    for (int k = 2; k <= 8 && k <= kmax && k <= nG; k *= 2) {
      const double t = (double)((R * k + C - 1) / C) / k;
      if (t < best - 1e-9) {
        best = t;
        pl.split = k;
      }
    }
  }
  if (pl.split == 1) pl.full_items = pl.vpanels;
  pl.items = pl.full_items + (pl.vpanels - pl.full_items) * pl.split;
  pl.pairs = pl.items < C ? pl.items : C;
  return pl;
}


}  // namespace mmd
}  // namespace edrl

using namespace edrl;
using namespace edrl::mmd;

extern "C" {

size_t edrl_mmd_workspace_bytes(int n_s, int n_t, int d, int flags) {
  if (n_s <= 0 || n_t <= 0 || d <= 0) return 0;
  return make_layout(n_s, n_t, d, flags).total;
}

int edrl_mmd_forward(const float *X, const float *Y, int n_s, int n_t, int d, float kernel_mul, int kernel_num,
                     int flags, int tile_rank, int tile_world, float *loss, float *stats, double *partial,
                     void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_CHECK_ARG(X && Y, "MK_MMD: null input");
  EDRL_CHECK_ARG(tile_world > 1 ? partial != nullptr : (loss && stats), "MK_MMD: null output");
  return forward_impl(MODE_LOSS, X, Y, n_s, n_t, d, kernel_mul, kernel_num, flags, tile_rank, tile_world, loss, stats,
                      partial, nullptr, workspace, workspace_bytes, stream);
}

int edrl_mmd_finalize(const double *partial, int n_s, int n_t, float kernel_mul, int kernel_num, float *loss,
                      float *stats, void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_CHECK_ARG(partial && loss && stats && workspace, "edrl_mmd_finalize: null argument");
  EDRL_CHECK_ARG(workspace_bytes >= 256, "edrl_mmd_finalize: workspace too small");
  mmd_finalize_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, reinterpret_cast<const double *>(workspace), n_s + n_t, kernel_mul, kernel_num, loss, stats);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_mmd_kernel_matrix(const float *X, const float *Y, int n_s, int n_t, int d, float kernel_mul, int kernel_num,
                           int flags, float *K, void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_CHECK_ARG(X && Y && K, "gaussian_kernel: null argument");
  const int mode = (flags & 0x100) ? MODE_GRAM : MODE_KMAT;   // 0x100: debug, raw centred Gram
  return forward_impl(mode, X, Y, n_s, n_t, d, kernel_mul, kernel_num, flags & 0xff, 0, 1, nullptr, nullptr, nullptr,
                      K, workspace, workspace_bytes, stream);
}

int edrl_mmd_backward(int n_s, int n_t, int d, float kernel_mul, int kernel_num, int flags, const float *stats,
                      const float *grad_out, int row_begin, int row_count, float *dZ, void *workspace,
                      size_t workspace_bytes, void *stream) {
  Layout L = make_layout(n_s, n_t, d, flags);
  if (int rc = check_common(n_s, n_t, d, kernel_mul, kernel_num, L, workspace, workspace_bytes)) return rc;
  EDRL_CHECK_ARG(stats && grad_out && dZ, "MK_MMD backward: null argument");
  EDRL_CHECK_ARG(row_begin >= 0 && row_count > 0 && row_begin + row_count <= L.n,
                 "MK_MMD backward: row range [%d, %d) outside [0, %d)", row_begin, row_begin + row_count, L.n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  CUtensorMap tm_hi, tm_lo, tm_thi, tm_tlo;
  if (int rc = make_tmap_2d_f32(&tm_hi, ws + L.off_zhi, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, BM, BK)) return rc;
  if (int rc = make_tmap_2d_f32(&tm_thi, ws + L.off_zthi, L.d_pad, L.n_pad, (uint64_t)L.n_pad * 4, DC, BK)) return rc;
  if (L.split3) {
    if (int rc = make_tmap_2d_f32(&tm_lo, ws + L.off_zlo, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, BM, BK)) return rc;
    if (int rc = make_tmap_2d_f32(&tm_tlo, ws + L.off_ztlo, L.d_pad, L.n_pad, (uint64_t)L.n_pad * 4, DC, BK))
      return rc;
  } else {
    tm_lo = tm_hi;
    tm_tlo = tm_thi;
  }
  BwdParams p;
  p.n = L.n; p.n_s = n_s; p.n_pad = L.n_pad; p.d = d; p.d_pad = L.d_pad;
  p.nb = L.n_pad / BN; p.kchunks = L.d_pad / BK; p.num = kernel_num; p.mul = kernel_mul;
  p.row_begin = row_begin; p.row_count = row_count;
  p.racc = reinterpret_cast<const double *>(ws + L.off_r);
  p.a = reinterpret_cast<const float *>(ws + L.off_a);
  p.zhi = reinterpret_cast<const float *>(ws + L.off_zhi);
  p.zlo = reinterpret_cast<const float *>(ws + L.off_zlo);
  p.stats = stats; p.grad_out = grad_out; p.dz = dZ;
  p.acc = nullptr; p.ticket = nullptr; p.partial = nullptr; p.loss = nullptr; p.stats_out = nullptr;
  p.n_t = n_t; p.finalize = 0; p.row_begin2 = 0; p.row_count2 = 0; p.fscale = nullptr;
  const bool fast = (kernel_mul == 2.0f && kernel_num == 5);
  static const bool legacy = (getenv("EDRL_MMD_BWD_LEGACY") != nullptr);   // A/B switch for profiling
  if (!L.split3 && !legacy) {
    // TF32: CTA-pair kernel (one Gram recompute per tile, Z_I resident for d <= 512)
    CUtensorMap tm_z64, tm_zt;
    if (int rc = make_tmap_2d_f32(&tm_z64, ws + L.off_zhi, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, 64, BK)) return rc;
    if (int rc = make_tmap_2d_f32(&tm_zt, ws + L.off_zthi, L.d_pad, L.n_pad, (uint64_t)L.n_pad * 4, 128, BK)) return rc;
    dim3 grid2(2 * ((row_count + BM - 1) / BM), (L.d_pad + P2_FEATS - 1) / P2_FEATS);
    // resident Z_I chunks: 8 of 16 for d <= 512 (measured: 16 -> 1.22 ms, 12 -> 1.10 ms, 8 -> 1.05 ms, 0 -> 1.12 ms
    // at N=8192, d=512: more residency leaves too few ring stages), none beyond
    const int res = (L.d_pad <= 512) ? 8 : 0;
    if (res == 8) {
      if (fast) return launch_bwd_pair_t<true, 8>(tm_z64, tm_zt, p, grid2, st);
      return launch_bwd_pair_t<false, 8>(tm_z64, tm_zt, p, grid2, st);
    }
    if (fast) return launch_bwd_pair_t<true, 0>(tm_z64, tm_zt, p, grid2, st);
    return launch_bwd_pair_t<false, 0>(tm_z64, tm_zt, p, grid2, st);
  }
  dim3 grid((row_count + BM - 1) / BM, (d + DC - 1) / DC);
  if (L.split3) {
    if (fast) return launch_bwd_t<true, true>(tm_hi, tm_lo, tm_thi, tm_tlo, p, grid, st);
    return launch_bwd_t<true, false>(tm_hi, tm_lo, tm_thi, tm_tlo, p, grid, st);
  }
  if (fast) return launch_bwd_t<false, true>(tm_hi, tm_lo, tm_thi, tm_tlo, p, grid, st);
  return launch_bwd_t<false, false>(tm_hi, tm_lo, tm_thi, tm_tlo, p, grid, st);
}

int edrl_mmd_grad_slabs(int n_s, int n_t, int d, int flags, int row_count, int row_count2) {
  if (n_s <= 0 || n_t <= 0 || d <= 0 || row_count <= 0 || row_count2 < 0) return 1;
  return make_plan(make_layout(n_s, n_t, d, flags), row_count, row_count2).split;
}

int edrl_mmd_sweep_plan(int n_s, int n_t, int d, int flags, int row_count, int row_count2, int sms, int *plan) {
  EDRL_CHECK_ARG(plan && n_s > 0 && n_t > 0 && d > 0 && row_count > 0 && row_count2 >= 0, "MK_MMD sweep_plan: bad argument");
  const Layout L = make_layout(n_s, n_t, d, flags);
  const SweepPlan pl = make_plan(L, row_count, row_count2, sms);
  plan[0] = pl.panels; plan[1] = pl.vpanels; plan[2] = pl.full_items; plan[3] = pl.split; plan[4] = pl.items;
  plan[5] = pl.pairs; plan[6] = L.n_pad / Q_GROUP; plan[7] = L.d_pad; plan[8] = pl.quad; plan[9] = pl.pass_feats;
  return 0;
}

int edrl_mmd_forward_grad(const float *X, const float *Y, int n_s, int n_t, int d, float kernel_mul, int kernel_num,
                          int flags, int row_begin, int row_count, int row_begin2, int row_count2, int finalize,
                          float *loss, float *stats, double *partial, float *U, void *workspace,
                          size_t workspace_bytes, void *stream) {
  EDRL_CHECK_ARG(X && Y && U, "MK_MMD forward_grad: null argument");
  EDRL_CHECK_ARG((flags & EDRL_MMD_3XTF32) == 0, "MK_MMD forward_grad: the fused pass is TF32 / TF32H only");
  Layout L = make_layout(n_s, n_t, d, flags);
  if (int rc = check_common(n_s, n_t, d, kernel_mul, kernel_num, L, workspace, workspace_bytes)) return rc;
  EDRL_CHECK_ARG(row_begin >= 0 && row_count > 0 && row_begin + row_count <= L.n,
                 "MK_MMD forward_grad: row range [%d, %d) outside [0, %d)", row_begin, row_begin + row_count, L.n);
  EDRL_CHECK_ARG(row_count2 == 0 || (row_begin2 >= row_begin + row_count && row_begin2 + row_count2 <= L.n),
                 "MK_MMD forward_grad: second row range [%d, %d) must follow the first and end inside [0, %d)",
                 row_begin2, row_begin2 + row_count2, L.n);
  EDRL_CHECK_ARG(finalize ? (loss && stats) : (partial != nullptr), "MK_MMD forward_grad: null output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  if (int rc = run_prep(X, Y, n_s, n_t, d, L, ws, st, !L.h16)) return rc;
  CUtensorMap tm_z64, tm_zt;
  if (int rc = make_tmap_2d_f32(&tm_z64, ws + L.off_zhi, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, 64, BK)) return rc;
  if (int rc = make_tmap_2d_f32(&tm_zt, ws + L.off_zthi, L.d_pad, L.n_pad, (uint64_t)L.n_pad * 4, 128, BK)) return rc;
  BwdParams p;
  p.n = L.n; p.n_s = n_s; p.n_pad = L.n_pad; p.d = d; p.d_pad = L.d_pad;
  p.nb = L.n_pad / BN; p.kchunks = L.d_pad / BK; p.num = kernel_num; p.mul = kernel_mul;
  p.row_begin = row_begin; p.row_count = row_count;
  p.racc = reinterpret_cast<const double *>(ws + L.off_r);
  p.a = reinterpret_cast<const float *>(ws + L.off_a);
  p.zhi = reinterpret_cast<const float *>(ws + L.off_zhi);
  p.zlo = nullptr;
  p.stats = nullptr; p.grad_out = nullptr; p.dz = U;
  p.acc = reinterpret_cast<double *>(ws + L.off_acc);
  p.ticket = reinterpret_cast<unsigned *>(ws + L.off_acc + 128);
  p.partial = partial; p.loss = loss; p.stats_out = stats;
  p.n_t = n_t; p.finalize = finalize; p.row_begin2 = row_begin2; p.row_count2 = row_count2;
  const SweepPlan pl = make_plan(L, row_count, row_count2);
  p.panels = pl.panels; p.full_items = pl.full_items; p.split = pl.split; p.items = pl.items;
  p.rowsum = reinterpret_cast<float *>(ws + L.off_rowsum);
  dim3 grid2((pl.quad ? 4 : 2) * pl.pairs, 1, 1);
  if (pl.quad)      // the two pairs of a cluster ADD their partial row sums
    EDRL_CUDA_OK(cudaMemsetAsync(ws + L.off_rowsum, 0,
                                 (size_t)((L.d_pad + 511) / 512) * SW_MAX_SPLIT * L.n_pad * sizeof(float), st));
  const bool fast = (kernel_mul == 2.0f && kernel_num == 5);
  p.fscale = reinterpret_cast<const int *>(ws + L.off_fscale);
  if (L.h16) {
    // binary16 (scaled) operands for G.Z; the Gram on TF32 (TF32H) or on the binary16 copy Z16 (F16S)
    CUtensorMap tm_z128, tm_zt16;
    if (int rc = make_tmap_2d_f16(&tm_zt16, ws + L.off_zt16, L.d_pad, L.n_pad, (uint64_t)L.n_pad * 2, 128, 64)) return rc;
    if (L.s16) {
      CUtensorMap tm_z64h;
      if (int rc = make_tmap_2d_f16(&tm_z64h, ws + L.off_z16, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 2, 64, 64)) return rc;
      if (int rc = make_tmap_2d_f16(&tm_z128, ws + L.off_z16, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 2, 128, 64)) return rc;
      if (pl.quad) {
        if (fast) return launch_sweep_quad_t<true, 2>(tm_z64h, tm_z128, tm_zt16, p, grid2, st);
        return launch_sweep_quad_t<false, 2>(tm_z64h, tm_z128, tm_zt16, p, grid2, st);
      }
      if (fast) return launch_sweep256_t<true, 2>(tm_z64h, tm_z128, tm_zt16, p, grid2, st);
      return launch_sweep256_t<false, 2>(tm_z64h, tm_z128, tm_zt16, p, grid2, st);
    }
    if (int rc = make_tmap_2d_f32(&tm_z128, ws + L.off_zhi, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, 128, BK)) return rc;
    if (pl.quad) {
      if (fast) return launch_sweep_quad_t<true, 1>(tm_z64, tm_z128, tm_zt16, p, grid2, st);
      return launch_sweep_quad_t<false, 1>(tm_z64, tm_z128, tm_zt16, p, grid2, st);
    }
    if (fast) return launch_sweep256_t<true, 1>(tm_z64, tm_z128, tm_zt16, p, grid2, st);
    return launch_sweep256_t<false, 1>(tm_z64, tm_z128, tm_zt16, p, grid2, st);
  }
  {
    CUtensorMap tm_z128;
    if (int rc = make_tmap_2d_f32(&tm_z128, ws + L.off_zhi, L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, 128, BK)) return rc;
    if (pl.quad) {
      if (fast) return launch_sweep_quad_t<true>(tm_z64, tm_z128, tm_zt, p, grid2, st);
      return launch_sweep_quad_t<false>(tm_z64, tm_z128, tm_zt, p, grid2, st);
    }
    if (fast) return launch_sweep256_t<true>(tm_z64, tm_z128, tm_zt, p, grid2, st);
    return launch_sweep256_t<false>(tm_z64, tm_z128, tm_zt, p, grid2, st);
  }
}

int edrl_mmd_apply_grad(int n_s, int n_t, int d, int flags, const float *stats, const float *grad_out,
                        const float *U, int row_begin, int row_count, int row_begin2, int row_count2, float *dZ,
                        void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_CHECK_ARG(stats && grad_out && U && dZ && workspace, "MK_MMD apply_grad: null argument");
  Layout L = make_layout(n_s, n_t, d, flags);
  EDRL_CHECK_ARG(workspace_bytes >= L.total, "MK_MMD apply_grad: workspace too small");
  EDRL_CHECK_ARG(row_begin >= 0 && row_count > 0 && row_begin + row_count <= L.n, "MK_MMD apply_grad: bad row range");
  const uint8_t *ws = reinterpret_cast<const uint8_t *>(workspace);
  const int rows = row_count + row_count2;
  const SweepPlan pl = make_plan(L, row_count, row_count2);
  const float *zhi = reinterpret_cast<const float *>(ws + L.off_zhi);
  const double *cs = reinterpret_cast<const double *>(ws + L.off_colsum_hi);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool v4 = (d % 4 == 0) && ((((uintptr_t)U | (uintptr_t)dZ) & 15) == 0);
  dim3 grid(rows, v4 ? (d + 2047) / 2048 : (d + 511) / 512);
  if (v4)
    mmd_apply_grad_kernel<true><<<grid, 128, 0, st>>>(U, zhi, cs, stats, grad_out, row_begin, row_count, row_begin2,
                                                      row_count2, d, L.d_pad, L.n, L.n_pad, pl.panels, pl.full_items, pl.split, pl.pass_feats,
        reinterpret_cast<const float *>(ws + L.off_rowsum), dZ);
  else
    mmd_apply_grad_kernel<false><<<grid, 128, 0, st>>>(U, zhi, cs, stats, grad_out, row_begin, row_count, row_begin2,
                                                       row_count2, d, L.d_pad, L.n, L.n_pad, pl.panels, pl.full_items, pl.split, pl.pass_feats,
        reinterpret_cast<const float *>(ws + L.off_rowsum), dZ);
  EDRL_LAUNCHED();
  return 0;
}

}  // extern "C"
