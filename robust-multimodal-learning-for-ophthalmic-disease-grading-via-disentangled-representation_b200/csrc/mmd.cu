// Part A -- multi-bandwidth Gaussian MMD on sm_100a (reference: code/MMD.py:3-74).
//
//   mmd_prep.cuh   K1   prep_colsum / prep_center: Z = [X; Y] centred, rounded to TF32 (hi [+ lo]), transposed copy,
//                       binary16 copies (tf32h / f16s), row norms r_i, block weights a_i, sum r -> analytic bandwidth
//   mmd_fwd.cuh    K2   loss-only forward: persistent, warp-specialised TMA -> smem ring -> tcgen05.mma (kind::tf32) ->
//                       TMEM -> fused distance / exp-sum / block-reduce epilogue over upper-triangular tiles
//                       (mmd_fwd_pair_kernel: 256 x 256 tiles, cta_group::2; mmd_fwd_kernel: 128 x 128, 3xTF32, K matrix)
//   mmd_bwd.cuh         BwdParams, shared-memory units (edrl_mmd_backward runs the sweep + apply_grad in place)
//   mmd_sweep.cuh  K3q  the training path: fused forward sums + gradient in one persistent sweep over the Gram tiles
//                       (mmd_sweep256_kernel: CTA pairs; mmd_sweep_quad_kernel: 4-CTA clusters for d > 768) and
//                       mmd_apply_grad_kernel
//   this file           workspace layout, shared device helpers, work list (make_plan), launchers, the C-ABI
// The kernel matrix never leaves the SM; nothing n x n is stored.
//
// Math: L_ij = max(0, r_i + r_j - 2 z_i.z_j); sigma_k = sigma_0 mul^k;
//   M = sum_ij a_i a_j sum_k exp(-L_ij / sigma_k),  a_i = 1/n_s (source rows) or -1/n_t (target rows);
//   Q_ij = sum_k exp(-L_ij / sigma_k) / mul^k;  D = sum_ij a_i a_j L_ij Q_ij / sigma_0^2;
//   G_ij = (-a_i a_j Q_ij / sigma_0 + c) [L_raw >= 0],  c = D / ((n^2 - n) mul^(num/2));
//   dZ_i = g sign(M) 4 (rowsum(G)_i z_i - (G Z)_i).
// Centring Z (distances are translation invariant) keeps TF32 rounding relative to the spread of the data
// and makes sum(L) = 2 n sum_i |z_i|^2 exactly the reference's bandwidth statistic (code/MMD.py:31).
#include <mutex>
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
  // (3xTF32: Z_lo directly behind Z_hi and Z_lo^T directly behind Z_hi^T -- the fused sweep addresses each pair through
  //  ONE tensor map of twice the rows)
  L.off_zhi = o;      o += zbytes;
  L.off_zlo = o;      if (L.split3) o += zbytes;
  L.off_zthi = o;     o += zbytes;
  L.off_ztlo = o;     if (L.split3) o += zbytes;
  L.off_zt16 = o;     if (L.h16) o += align_up((size_t)L.n_pad * L.d_pad * 2, 1024);   // binary16 Z^T [d_pad, n_pad]
  L.off_z16 = o;      if (L.s16) o += align_up((size_t)L.n_pad * L.d_pad * 2, 1024);   // binary16 Z [n_pad, d_pad]
  // fused sweep: row sums of G' per (512-column feature pass, column slab), float[passes][8][n_pad]
  L.off_rowsum = o;   o += align_up((size_t)((L.d_pad + 511) / 512) * 8 * L.n_pad * 4, 1024);
  L.total = o;
  return L;
}

#include "mmd_prep.cuh"

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

#include "mmd_fwd.cuh"
#include "mmd_bwd.cuh"

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

#include "mmd_sweep.cuh"

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

static bool quad_wanted(const Layout &L) {
  static const bool no_quad = (getenv("EDRL_MMD_QUAD") != nullptr && atoi(getenv("EDRL_MMD_QUAD")) == 0);   // A/B runs
  // (up to 768 columns the second pair would hold one 256-column tile or less: two pair passes are faster)
  // (3xTF32 has no quad variant: the pair kernel sweeps the Gram once per 512-column pass)
  return L.d_pad > P2_FEATS + P2_FEATS / 2 && !no_quad && !L.split3;
}

// force_kind: -1 = by width, 0 = pair kernel, 1 = quad kernel; clusters_override: persistent clusters available (hybrid)
static SweepPlan make_plan_panels(const Layout &L, int panels, int sms_override = 0, int force_kind = -1,
                                  int clusters_override = 0, int max_split = SW_MAX_SPLIT) {
  SweepPlan pl;
  pl.panels = panels;
  pl.quad = (force_kind < 0 ? quad_wanted(L) : force_kind != 0) ? 1 : 0;
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
  if (clusters_override > 0) C = clusters_override;
  const int nG = L.n_pad / Q_GROUP;
  pl.full_items = (pl.vpanels / C) * C;
  const int R = pl.vpanels - pl.full_items;
  pl.split = 1;
  static const char *env = getenv("EDRL_MMD_SLABS");      // =1: never split (A/B runs)
  const int kmax = (env && atoi(env) < max_split) ? atoi(env) : max_split;
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

static SweepPlan make_plan(const Layout &L, int row_count, int row_count2, int sms_override = 0) {
  return make_plan_panels(L, (row_count + BM - 1) / BM + (row_count2 + BM - 1) / BM, sms_override);
}

// A second stream for one call: `side` runs concurrently with the caller's stream between fork() and join().  One side
// stream and two events per device, created on first use (event record / wait only: legal under stream capture, where
// they fork and join the captured graph).
struct ForkJoin {
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // the side stream and its two events are shared by every call on the device: host threads take turns from fork() to
  // join() (enqueue calls only), or one thread's wait could pick up the other's record
  std::unique_lock<std::mutex> turn;
  int fork(cudaStream_t main) {
    static std::mutex s_mutex[16];
    static cudaStream_t s_side[16] = {};
    static cudaEvent_t s_ev[16][2] = {};
    int dev = 0;
    EDRL_CUDA_OK(cudaGetDevice(&dev));
    EDRL_CHECK_ARG(dev >= 0 && dev < 16, "MK_MMD: device index %d not supported by the hybrid launch", dev);
    turn = std::unique_lock<std::mutex>(s_mutex[dev]);
    if (s_side[dev] == nullptr) {
      EDRL_CUDA_OK(cudaStreamCreateWithFlags(&s_side[dev], cudaStreamNonBlocking));
      EDRL_CUDA_OK(cudaEventCreateWithFlags(&s_ev[dev][0], cudaEventDisableTiming));
      EDRL_CUDA_OK(cudaEventCreateWithFlags(&s_ev[dev][1], cudaEventDisableTiming));
    }
    side = s_side[dev];
    ev_fork = s_ev[dev][0];
    ev_join = s_ev[dev][1];
    EDRL_CUDA_OK(cudaEventRecord(ev_fork, main));
    EDRL_CUDA_OK(cudaStreamWaitEvent(side, ev_fork, 0));
    return 0;
  }
  int join(cudaStream_t main) {
    EDRL_CUDA_OK(cudaEventRecord(ev_join, side));
    EDRL_CUDA_OK(cudaStreamWaitEvent(main, ev_join, 0));
    return 0;
  }
};

// Hybrid launch for d > 768: 4-CTA clusters must sit inside one GPC, so only 33 of them are resident on 148 SMs and 16
// SMs would idle through the whole sweep.  The quad kernel takes the first panels, a PAIR kernel (two 512-column feature
// passes per panel, concurrently on a forked stream) the last ones, in the ratio of their measured speeds: a pair needs
// ny (d_pad + 512) per panel where a cluster needs ny_q (d_pad / 2 + 512) (S + P columns streamed; x 0.88 measured at
// d = 1024: 11.1 ms against 4.2 ms per panel at n = 131072).
struct HybridPlan {
  bool on;
  SweepPlan q, pr;     // quad part: panels [0, q.panels); pair part: panels [q.panels, q.panels + pr.panels)
};
static HybridPlan make_hybrid(const Layout &L, int row_count, int row_count2) {
  HybridPlan h;
  h.on = false;
  h.q = make_plan(L, row_count, row_count2);
  h.pr = h.q;
  static const int mode = getenv("EDRL_MMD_HYBRID") ? atoi(getenv("EDRL_MMD_HYBRID")) : 1;   // 0 off, 1 auto, 2 force (tests)
  if (!h.q.quad || mode == 0 || L.h16 || L.split3) return h;     // (the TF32 sweep only: tf32h / f16s keep one launch)
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  const int cq = quad_clusters_resident();
  const int cp = (sms - 4 * cq) / 2;
  if (cq <= 0 || cp <= 0) return h;
  const int P = h.q.panels;
  const int ny_p = (L.d_pad + P2_FEATS - 1) / P2_FEATS, ny_q = (L.d_pad + 2 * P2_FEATS - 1) / (2 * P2_FEATS);
  static const double kr = getenv("EDRL_MMD_HYBRID_RATIO") ? atof(getenv("EDRL_MMD_HYBRID_RATIO")) : 0.88;   // A/B runs
  const double ratio = kr * (double)ny_p * (L.d_pad + 512.0) / ((double)ny_q * (L.d_pad / 2.0 + 512.0));
  int pp = (int)((double)P * (cp / ratio) / (cq + cp / ratio));
  if (mode == 2 && pp < 1 && P >= 2) pp = 1;
  if (pp < (mode == 2 ? 1 : cp) || pp >= P) return h;        // too few panels to keep the idle SMs busy for the whole sweep
  h.on = true;
  h.q = make_plan_panels(L, P - pp, 0, 1);
  h.pr = make_plan_panels(L, pp, 0, 0, cp);
  return h;
}


static HybridPlan single_plan(const Layout &L, int row_count, int row_count2) {
  HybridPlan h;
  h.on = false;
  h.q = make_plan_panels(L, (row_count + BM - 1) / BM + (row_count2 + BM - 1) / BM, 0, -1, 0, 1);
  h.pr = h.q;
  return h;
}

// The sweep of a prepared workspace over the row ranges [row_begin, +row_count) and [row_begin2, +row_count2): forward
// block sums (into `partial`, or finalised into loss / stats) and U.  single: one launch, no column slabs (the plan
// edrl_mmd_backward's in-place use needs: one U slab).
static int launch_sweep(const Layout &L, uint8_t *ws, int n_s, int n_t, int d, float kernel_mul, int kernel_num,
                        int row_begin, int row_count, int row_begin2, int row_count2, int finalize, float *loss,
                        float *stats, double *partial, float *U, cudaStream_t st, bool single) {
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
  const HybridPlan hy = single ? single_plan(L, row_count, row_count2) : make_hybrid(L, row_count, row_count2);
  const SweepPlan pl = hy.q;
  p.panels = pl.panels; p.full_items = pl.full_items; p.split = pl.split; p.items = pl.items;
  p.panel0 = 0;
  {
    static const int order = getenv("EDRL_MMD_QUAD_ORDER") ? atoi(getenv("EDRL_MMD_QUAD_ORDER")) : 1;   // A/B runs
    p.s_ahead = order;
  }
  p.ticket_total = hy.on ? 4 * hy.q.pairs + 2 * hy.pr.pairs : 0;
  p.rowsum = reinterpret_cast<float *>(ws + L.off_rowsum);
  dim3 grid2((pl.quad ? 4 : 2) * pl.pairs, 1, 1);
  if (pl.quad)      // the two pairs of a cluster ADD their partial row sums
    EDRL_CUDA_OK(cudaMemsetAsync(ws + L.off_rowsum, 0,
                                 (size_t)((L.d_pad + 511) / 512) * SW_MAX_SPLIT * L.n_pad * sizeof(float), st));
  const bool fast = (kernel_mul == 2.0f && kernel_num == 5);
  p.fscale = reinterpret_cast<const int *>(ws + L.off_fscale);
  if (L.split3) {
    // 3xTF32: hi and lo parts behind one another -- rows [0, n_pad) of the S maps are Z_hi, [n_pad, 2 n_pad) Z_lo; rows
    // [0, d_pad) of the P map are Z_hi^T, [d_pad, 2 d_pad) Z_lo^T
    CUtensorMap tm3_z64, tm3_z128, tm3_zt;
    if (int rc = make_tmap_2d_f32(&tm3_z64, ws + L.off_zhi, 2ull * L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, 64, BK)) return rc;
    if (int rc = make_tmap_2d_f32(&tm3_z128, ws + L.off_zhi, 2ull * L.n_pad, L.d_pad, (uint64_t)L.d_pad * 4, 128, BK)) return rc;
    if (int rc = make_tmap_2d_f32(&tm3_zt, ws + L.off_zthi, 2ull * L.d_pad, L.n_pad, (uint64_t)L.n_pad * 4, 128, BK)) return rc;
    if (fast) return launch_sweep256_t<true, 3>(tm3_z64, tm3_z128, tm3_zt, p, grid2, st);
    return launch_sweep256_t<false, 3>(tm3_z64, tm3_z128, tm3_zt, p, grid2, st);
  }
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
      if (hy.on) {
        // the quad clusters first (they need whole groups of 4 SMs inside a GPC), then the pair part on the forked
        // stream: its few clusters land on the SMs the quad grid left over
        ForkJoin fj;
        if (int rc = fj.fork(st)) return rc;
        BwdParams pp = p;
        pp.panels = hy.pr.panels; pp.full_items = hy.pr.full_items; pp.split = hy.pr.split; pp.items = hy.pr.items;
        pp.panel0 = hy.q.panels;
        const dim3 gridp(2 * hy.pr.pairs, 1, 1);
        const int rc2 = fast ? launch_sweep_quad_t<true>(tm_z64, tm_z128, tm_zt, p, grid2, st)
                             : launch_sweep_quad_t<false>(tm_z64, tm_z128, tm_zt, p, grid2, st);
        const int rc1 = fast ? launch_sweep256_t<true>(tm_z64, tm_z128, tm_zt, pp, gridp, fj.side)
                             : launch_sweep256_t<false>(tm_z64, tm_z128, tm_zt, pp, gridp, fj.side);
        const int rc3 = fj.join(st);
        return rc1 ? rc1 : (rc2 ? rc2 : rc3);
      }
      if (fast) return launch_sweep_quad_t<true>(tm_z64, tm_z128, tm_zt, p, grid2, st);
      return launch_sweep_quad_t<false>(tm_z64, tm_z128, tm_zt, p, grid2, st);
    }
    if (fast) return launch_sweep256_t<true>(tm_z64, tm_z128, tm_zt, p, grid2, st);
    return launch_sweep256_t<false>(tm_z64, tm_z128, tm_zt, p, grid2, st);
  }
}

static int launch_apply(const Layout &L, const uint8_t *ws, int d, const float *stats, const float *grad_out, const float *U,
                        int row_begin, int row_count, int row_begin2, int row_count2, float *dZ, const HybridPlan &hy,
                        cudaStream_t st) {
  const int rows = row_count + row_count2;
  const ApplyPlan pa{hy.q.panels, hy.q.full_items, hy.q.split, hy.q.pass_feats};
  const ApplyPlan pb{hy.pr.panels, hy.pr.full_items, hy.pr.split, hy.pr.pass_feats};
  const float *zhi = reinterpret_cast<const float *>(ws + L.off_zhi);
  const float *zlo = L.split3 ? reinterpret_cast<const float *>(ws + L.off_zlo) : nullptr;
  const double *cs = reinterpret_cast<const double *>(ws + L.off_colsum_hi);
  const bool v4 = (d % 4 == 0) && ((((uintptr_t)U | (uintptr_t)dZ) & 15) == 0);
  dim3 grid(rows, v4 ? (d + 2047) / 2048 : (d + 511) / 512);
  if (v4)
    mmd_apply_grad_kernel<true><<<grid, 128, 0, st>>>(U, zhi, zlo, cs, stats, grad_out, row_begin, row_count, row_begin2,
                                                      row_count2, d, L.d_pad, L.n, L.n_pad, pa, pb,
        reinterpret_cast<const float *>(ws + L.off_rowsum), dZ);
  else
    mmd_apply_grad_kernel<false><<<grid, 128, 0, st>>>(U, zhi, zlo, cs, stats, grad_out, row_begin, row_count, row_begin2,
                                                       row_count2, d, L.d_pad, L.n, L.n_pad, pa, pb,
        reinterpret_cast<const float *>(ws + L.off_rowsum), dZ);
  EDRL_LAUNCHED();
  return 0;
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
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(X && Y, "MK_MMD: null input");
  EDRL_CHECK_ARG(tile_world > 1 ? partial != nullptr : (loss && stats), "MK_MMD: null output");
  return forward_impl(MODE_LOSS, X, Y, n_s, n_t, d, kernel_mul, kernel_num, flags, tile_rank, tile_world, loss, stats,
                      partial, nullptr, workspace, workspace_bytes, stream);
}

int edrl_mmd_finalize(const double *partial, int n_s, int n_t, float kernel_mul, int kernel_num, float *loss,
                      float *stats, void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(partial && loss && stats && workspace, "edrl_mmd_finalize: null argument");
  EDRL_CHECK_ARG(workspace_bytes >= 256, "edrl_mmd_finalize: workspace too small");
  mmd_finalize_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      partial, reinterpret_cast<const double *>(workspace), n_s + n_t, kernel_mul, kernel_num, loss, stats);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_mmd_kernel_matrix(const float *X, const float *Y, int n_s, int n_t, int d, float kernel_mul, int kernel_num,
                           int flags, float *K, void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(X && Y && K, "gaussian_kernel: null argument");
  const int mode = (flags & 0x100) ? MODE_GRAM : MODE_KMAT;   // 0x100: debug, raw centred Gram
  return forward_impl(mode, X, Y, n_s, n_t, d, kernel_mul, kernel_num, flags & 0xff, 0, 1, nullptr, nullptr, nullptr,
                      K, workspace, workspace_bytes, stream);
}

int edrl_mmd_backward(int n_s, int n_t, int d, float kernel_mul, int kernel_num, int flags, const float *stats,
                      const float *grad_out, int row_begin, int row_count, float *dZ, void *workspace,
                      size_t workspace_bytes, void *stream) {
  EDRL_DEVICE_GUARD();
  Layout L = make_layout(n_s, n_t, d, flags);
  if (int rc = check_common(n_s, n_t, d, kernel_mul, kernel_num, L, workspace, workspace_bytes)) return rc;
  EDRL_CHECK_ARG(stats && grad_out && dZ, "MK_MMD backward: null argument");
  EDRL_CHECK_ARG(row_begin >= 0 && row_count > 0 && row_begin + row_count <= L.n,
                 "MK_MMD backward: row range [%d, %d) outside [0, %d)", row_begin, row_begin + row_count, L.n);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace);
  // every precision mode: the fused sweep (its forward sums are a by-product nobody reads here) writes U into dZ,
  // apply_grad finishes it in place -- one U slab, hence the single-launch plan
  if (int rc = launch_sweep(L, ws, n_s, n_t, d, kernel_mul, kernel_num, row_begin, row_count, 0, 0, 0, nullptr, nullptr,
                            nullptr, dZ, st, true))
    return rc;
  return launch_apply(L, ws, d, stats, grad_out, dZ, row_begin, row_count, 0, 0, dZ, single_plan(L, row_count, 0), st);
}

int edrl_mmd_grad_slabs(int n_s, int n_t, int d, int flags, int row_count, int row_count2) {
  if (n_s <= 0 || n_t <= 0 || d <= 0 || row_count <= 0 || row_count2 < 0) return 1;
  const HybridPlan h = make_hybrid(make_layout(n_s, n_t, d, flags), row_count, row_count2);
  return (h.on && h.pr.split > h.q.split) ? h.pr.split : h.q.split;
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
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(X && Y && U, "MK_MMD forward_grad: null argument");
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
  return launch_sweep(L, ws, n_s, n_t, d, kernel_mul, kernel_num, row_begin, row_count, row_begin2, row_count2, finalize,
                      loss, stats, partial, U, st, false);
}

int edrl_mmd_apply_grad(int n_s, int n_t, int d, int flags, const float *stats, const float *grad_out,
                        const float *U, int row_begin, int row_count, int row_begin2, int row_count2, float *dZ,
                        void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(stats && grad_out && U && dZ && workspace, "MK_MMD apply_grad: null argument");
  Layout L = make_layout(n_s, n_t, d, flags);
  EDRL_CHECK_ARG(workspace_bytes >= L.total, "MK_MMD apply_grad: workspace too small");
  EDRL_CHECK_ARG(row_begin >= 0 && row_count > 0 && row_begin + row_count <= L.n, "MK_MMD apply_grad: bad row range");
  const uint8_t *ws = reinterpret_cast<const uint8_t *>(workspace);
  return launch_apply(L, ws, d, stats, grad_out, U, row_begin, row_count, row_begin2, row_count2, dZ,
                      make_hybrid(L, row_count, row_count2), reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
