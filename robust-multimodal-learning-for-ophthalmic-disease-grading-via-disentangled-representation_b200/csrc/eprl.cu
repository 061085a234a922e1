// Part B -- Essence-Point scoring and top-k selection on sm_100a
// (reference: class EPRL, code/fusion_net.py:133-255).  These are HBM / latency bound streaming kernels:
// coalesced (vectorised where the shape allows) loads, warp-shuffle reductions, no tensor cores.
//
//   K4  token_stats_{fwd,bwd}, token_featmean : F.normalize(z, dim=1) + mean over tokens, hoisted (B3/B4)
//   K5  proxy_normalize_{fwd,bwd}             : z_p = mu + sigma eps, normalised over the sample dim (B2/B3)
//   K6  score_{fwd,bwd}                       : att = zbar . z_pn^T  (small fp32 GEMMs, B4)
//   K7  topk_rows / select_topk_fwd           : exact radix select + bitonic sort, label-addressed rows (B5/B6)
//   K8  proxy_loss_fwd / select_loss_bwd      : loss and its scatter backward (B6/B7)
//   K9  gather_rows_{fwd,bwd}                 : feature gather by index (north-star extension)
#include <math.h>
#include <type_traits>
#include <stdlib.h>

#include "../../include/edrl_b200.h"
#include "common.cuh"

namespace edrl {
namespace eprl {

constexpr float NORM_EPS = 1e-12f;   // torch.nn.functional.normalize default eps

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ----------------------------------------------------------------------------- K4 token statistics
// grid (ceil(F/32), B), block (32, 8): lane -> feature column, y -> token slice
__global__ void __launch_bounds__(256)
token_stats_fwd_kernel(const float *__restrict__ z, int T, int F, float *__restrict__ zbar,
                       float *__restrict__ colsum, float *__restrict__ colnorm) {
  __shared__ float s_sum[8][33], s_sq[8][33];
  const int f = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y;
  float sum = 0.f, sq = 0.f;
  if (f < F) {
    const float *zp = z + (size_t)b * T * F + f;
#pragma unroll 4
    for (int t = threadIdx.y; t < T; t += 8) {
      const float v = __ldg(zp + (size_t)t * F);
      sum += v;
      sq = fmaf(v, v, sq);
    }
  }
  s_sum[threadIdx.y][threadIdx.x] = sum;
  s_sq[threadIdx.y][threadIdx.x] = sq;
  __syncthreads();
  if (threadIdx.y == 0 && f < F) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      sum += s_sum[k][threadIdx.x];
      sq += s_sq[k][threadIdx.x];
    }
    const float nrm = sqrtf(sq);
    const size_t o = (size_t)b * F + f;
    colsum[o] = sum;
    colnorm[o] = nrm;
    zbar[o] = sum / ((float)T * fmaxf(nrm, NORM_EPS));
  }
}

// dz = alpha[b,f] + beta[b,f] * z ; alpha = dzbar/(T m), beta = -dzbar S /(T m^2 nrm) [nrm > eps]
__global__ void __launch_bounds__(256)
token_stats_bwd_kernel(const float *__restrict__ z, const float *__restrict__ colsum,
                       const float *__restrict__ colnorm, const float *__restrict__ dzbar, int T, int F,
                       size_t total, float *__restrict__ dz) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int f = (int)(i % F);
  const size_t bt = i / F;
  const size_t b = bt / T;
  const size_t o = b * F + f;
  const float nrm = colnorm[o];
  const float m = fmaxf(nrm, NORM_EPS);
  const float g = dzbar[o];
  const float alpha = g / ((float)T * m);
  const float beta = (nrm > NORM_EPS) ? (-g * colsum[o] / ((float)T * m * m * nrm)) : 0.f;
  dz[i] = fmaf(beta, z[i], alpha);
}


// 128-bit variants (F % 4 == 0, 16-byte aligned z): the streams over z are the HBM traffic of Part B.
// forward: grid (ceil(F/128), B), block (32, 8): lane -> one float4 column group, y -> token slice
__global__ void __launch_bounds__(256)
token_stats_fwd_v4_kernel(const float4 *__restrict__ z, int T, int F4, float4 *__restrict__ zbar,
                          float4 *__restrict__ colsum, float4 *__restrict__ colnorm) {
  __shared__ float4 s_sum[8][33], s_sq[8][33];
  const int f4 = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y;
  float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), sq = sum;
  if (f4 < F4) {
    const float4 *zp = z + (size_t)b * T * F4 + f4;
#pragma unroll 4
    for (int t = threadIdx.y; t < T; t += 8) {
      const float4 v = __ldg(zp + (size_t)t * F4);
      sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      sq.x = fmaf(v.x, v.x, sq.x); sq.y = fmaf(v.y, v.y, sq.y); sq.z = fmaf(v.z, v.z, sq.z); sq.w = fmaf(v.w, v.w, sq.w);
    }
  }
  s_sum[threadIdx.y][threadIdx.x] = sum;
  s_sq[threadIdx.y][threadIdx.x] = sq;
  __syncthreads();
  if (threadIdx.y == 0 && f4 < F4) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 a = s_sum[k][threadIdx.x], q = s_sq[k][threadIdx.x];
      sum.x += a.x; sum.y += a.y; sum.z += a.z; sum.w += a.w;
      sq.x += q.x; sq.y += q.y; sq.z += q.z; sq.w += q.w;
    }
    const float4 nrm = make_float4(sqrtf(sq.x), sqrtf(sq.y), sqrtf(sq.z), sqrtf(sq.w));
    const float ft = (float)T;
    const size_t o = (size_t)b * F4 + f4;
    colsum[o] = sum;
    colnorm[o] = nrm;
    zbar[o] = make_float4(sum.x / (ft * fmaxf(nrm.x, NORM_EPS)), sum.y / (ft * fmaxf(nrm.y, NORM_EPS)),
                          sum.z / (ft * fmaxf(nrm.z, NORM_EPS)), sum.w / (ft * fmaxf(nrm.w, NORM_EPS)));
  }
}

__device__ __forceinline__ void token_bwd_coefs(float nrm, float g, float s, float ft, float &alpha, float &beta) {
  const float m = fmaxf(nrm, NORM_EPS);
  alpha = g / (ft * m);
  beta = (nrm > NORM_EPS) ? (-g * s / (ft * m * m * nrm)) : 0.f;
}

// backward: grid (ceil(T/TCH), B), block 256 = (F4 columns) x (256/F4 token rows); coefficients once per thread
constexpr int TSB_TCH = 32;     // tokens per block
__global__ void __launch_bounds__(256)
token_stats_bwd_v4_kernel(const float4 *__restrict__ z, const float4 *__restrict__ colsum,
                          const float4 *__restrict__ colnorm, const float4 *__restrict__ dzbar, int T, int F4,
                          float4 *__restrict__ dz) {
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * TSB_TCH;
  const int t1 = min(t0 + TSB_TCH, T);
  const int rows_per_iter = 256 / F4;            // F4 divides 256 (host checks)
  const int f4 = threadIdx.x % F4;
  const int ty = threadIdx.x / F4;
  const size_t o = (size_t)b * F4 + f4;
  const float4 nrm = __ldg(colnorm + o), g = __ldg(dzbar + o), cs = __ldg(colsum + o);
  const float ft = (float)T;
  float4 al, be;
  token_bwd_coefs(nrm.x, g.x, cs.x, ft, al.x, be.x);
  token_bwd_coefs(nrm.y, g.y, cs.y, ft, al.y, be.y);
  token_bwd_coefs(nrm.z, g.z, cs.z, ft, al.z, be.z);
  token_bwd_coefs(nrm.w, g.w, cs.w, ft, al.w, be.w);
  const float4 *zp = z + (size_t)b * T * F4 + f4;
  float4 *dp = dz + (size_t)b * T * F4 + f4;
#pragma unroll 4
  for (int t = t0 + ty; t < t1; t += rows_per_iter) {
    const float4 v = __ldg(zp + (size_t)t * F4);
    dp[(size_t)t * F4] = make_float4(fmaf(be.x, v.x, al.x), fmaf(be.y, v.y, al.y), fmaf(be.z, v.z, al.z),
                                     fmaf(be.w, v.w, al.w));
  }
}

// zmean[b,t] = mean_f z[b,t,f] / max(colnorm[b,f], eps); one warp per (b,t)
__global__ void __launch_bounds__(256)
token_featmean_kernel(const float *__restrict__ z, const float *__restrict__ colnorm, int B, int T, int F,
                      float *__restrict__ zmean) {
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= B * T) return;
  const int b = w / T;
  const float *zp = z + (size_t)w * F;
  const float *np_ = colnorm + (size_t)b * F;
  float acc = 0.f;
  for (int f = lane; f < F; f += 32) acc += zp[f] / fmaxf(np_[f], NORM_EPS);
  acc = warp_sum(acc);
  if (lane == 0) zmean[w] = acc / (float)F;
}

// ----------------------------------------------------------------------------- K5 proxies
// softplus exactly as torch.nn.functional.softplus (beta = 1, threshold = 20), code/fusion_net.py:118
__device__ __forceinline__ float softplus_torch(float x) { return (x > 20.f) ? x : log1pf(expf(x)); }

// grid (ceil(F/32), C), block (32, 32): lane -> feature, y -> sample slice.
// mu / sigma rows are `ldp` floats apart; raw != 0: `sigma` holds the raw proxy half and softplus is applied here
// (lets the fused entry point read the [C, 2F] proxies parameter directly).
__global__ void __launch_bounds__(1024)
proxy_normalize_fwd_kernel(const float *__restrict__ mu, const float *__restrict__ sigma,
                           const float *__restrict__ eps, int S, int F, int ldp, int raw, float *__restrict__ z_pn,
                           float *__restrict__ pnorm) {
  __shared__ float s_sq[32][33];
  const int f = blockIdx.x * 32 + threadIdx.x;
  const int c = blockIdx.y;
  float m = 0.f, sg = 0.f, sq = 0.f;
  if (f < F) {
    m = mu[(size_t)c * ldp + f];
    sg = sigma[(size_t)c * ldp + f];
    if (raw) sg = softplus_torch(sg);
    const float *ep = eps + (size_t)c * S * F + f;
    for (int s = threadIdx.y; s < S; s += 32) {
      const float v = fmaf(sg, __ldg(ep + (size_t)s * F), m);
      sq = fmaf(v, v, sq);
    }
  }
  s_sq[threadIdx.y][threadIdx.x] = sq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll 8
  for (int k = 0; k < 32; ++k) tot += s_sq[k][threadIdx.x];
  if (f < F) {
    const float nrm = sqrtf(tot);
    if (threadIdx.y == 0) pnorm[(size_t)c * F + f] = nrm;
    const float inv = 1.f / fmaxf(nrm, NORM_EPS);
    const float *ep = eps + (size_t)c * S * F + f;
    float *op = z_pn + (size_t)c * S * F + f;
    for (int s = threadIdx.y; s < S; s += 32) op[(size_t)s * F] = fmaf(sg, __ldg(ep + (size_t)s * F), m) * inv;
  }
}

// five column sums over s: g, g*eps, g*z_p, z_p, z_p*eps  ->  dmu, dsigma
__global__ void __launch_bounds__(1024)
proxy_normalize_bwd_kernel(const float *__restrict__ mu, const float *__restrict__ sigma,
                           const float *__restrict__ eps, const float *__restrict__ pnorm,
                           const float *__restrict__ dz_pn, int S, int F, int ldp, int raw, int ldo,
                           float *__restrict__ dmu, float *__restrict__ dsigma) {
  __shared__ float red[5][32][33];
  const int f = blockIdx.x * 32 + threadIdx.x;
  const int c = blockIdx.y;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
  float m = 0.f, sg = 0.f, chain = 1.f;
  if (f < F) {
    m = mu[(size_t)c * ldp + f];
    sg = sigma[(size_t)c * ldp + f];
    if (raw) {
      chain = 1.f / (1.f + expf(-sg));        // d softplus / d raw = sigmoid(raw) (1 beyond the threshold, to fp32)
      sg = softplus_torch(sg);
    }
    const float *ep = eps + (size_t)c * S * F + f;
    const float *gp = dz_pn + (size_t)c * S * F + f;
    for (int s = threadIdx.y; s < S; s += 32) {
      const float e = __ldg(ep + (size_t)s * F);
      const float g = __ldg(gp + (size_t)s * F);
      const float zp = fmaf(sg, e, m);
      a0 += g;
      a1 = fmaf(g, e, a1);
      a2 = fmaf(g, zp, a2);
      a3 += zp;
      a4 = fmaf(zp, e, a4);
    }
  }
  red[0][threadIdx.y][threadIdx.x] = a0;
  red[1][threadIdx.y][threadIdx.x] = a1;
  red[2][threadIdx.y][threadIdx.x] = a2;
  red[3][threadIdx.y][threadIdx.x] = a3;
  red[4][threadIdx.y][threadIdx.x] = a4;
  __syncthreads();
  if (threadIdx.y == 0 && f < F) {
    float t[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      float v = 0.f;
      for (int k = 0; k < 32; ++k) v += red[q][k][threadIdx.x];
      t[q] = v;
    }
    const float nrm = pnorm[(size_t)c * F + f];
    const float q = fmaxf(nrm, NORM_EPS);
    const float kk = (nrm > NORM_EPS) ? (t[2] / (q * q * nrm)) : 0.f;
    dmu[(size_t)c * ldo + f] = t[0] / q - t[3] * kk;
    dsigma[(size_t)c * ldo + f] = (t[1] / q - t[4] * kk) * chain;
  }
}

// ----------------------------------------------------------------------------- K6 small fp32 GEMM
// C[m,n] = sum_k A(m,k) B(k,n), arbitrary strides; 64x64 tile, 16x16 threads, 4x4 per thread
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float *__restrict__ A, long long sam, long long sak, const float *__restrict__ Bm,
                     long long sbk, long long sbn, float *__restrict__ Cm, int M, int N, int K) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      int kk, mm;
      if (sak == 1) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < K) ? __ldg(A + gm * sam + gk * sak) : 0.f;
    }
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      int kk, nn;
      if (sbk == 1) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
      const int gn = n0 + nn, gk = k0 + kk;
      Bs[kk][nn] = (gn < N && gk < K) ? __ldg(Bm + gk * sbk + gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn < N) Cm[(size_t)gm * N + gn] = acc[i][j];
    }
  }
}

static int sgemm(const float *A, long long sam, long long sak, const float *B, long long sbk, long long sbn, float *C,
                 int M, int N, int K, cudaStream_t st) {
  dim3 grid((N + 63) / 64, (M + 63) / 64);
  sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, M, N, K);
  EDRL_LAUNCHED();
  return 0;
}


// ----------------------------------------------------------------------------- K6 score kernels (skinny shapes)
// The score contraction is [B, F] x [F, R] with B ~ 4-64, F = 256, R = C S = 1600: far too skinny for 64 x 64 tiles
// (25 blocks forward; the dzbar product ran in 4 blocks and took 409 us).  One warp per proxy row r instead.
// att[b, r] = sum_f zbar[b, f] z_pn[r, f]
__global__ void __launch_bounds__(256)
score_fwd_kernel(const float *__restrict__ zbar, const float *__restrict__ z_pn, int B, int R, int F,
                 float *__restrict__ att) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  const float *pr = z_pn + (size_t)r * F;
  for (int b = 0; b < B; ++b) {
    const float *zb = zbar + (size_t)b * F;
    float acc = 0.f;
    for (int f = lane; f < F; f += 32) acc = fmaf(__ldg(zb + f), __ldg(pr + f), acc);
    acc = warp_sum(acc);
    if (lane == 0) att[(size_t)b * R + r] = acc;
  }
}
// dzbar[b, f] = sum_r datt[b, r] z_pn[r, f]; grid (ceil(F/32), B), block (32, 8)
__global__ void __launch_bounds__(256)
score_bwd_dzbar_kernel(const float *__restrict__ datt, const float *__restrict__ z_pn, int R, int F,
                       float *__restrict__ dzbar) {
  __shared__ float s_acc[8][33];
  const int f = blockIdx.x * 32 + threadIdx.x;
  const int b = blockIdx.y;
  float acc = 0.f;
  if (f < F) {
    const float *dr = datt + (size_t)b * R;
    for (int r = threadIdx.y; r < R; r += 8) {
      const float dv = __ldg(dr + r);
      if (dv != 0.f) acc = fmaf(dv, __ldg(z_pn + (size_t)r * F + f), acc);   // datt is zero off the selection
    }
  }
  s_acc[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && f < F) {
#pragma unroll
    for (int k = 1; k < 8; ++k) acc += s_acc[k][threadIdx.x];
    dzbar[(size_t)b * F + f] = acc;
  }
}
// dz_pn[r, f] = sum_b datt[b, r] zbar[b, f]; one warp per r
__global__ void __launch_bounds__(256)
score_bwd_dzpn_kernel(const float *__restrict__ datt, const float *__restrict__ zbar, int B, int R, int F,
                      float *__restrict__ dz_pn) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= R) return;
  for (int f0 = 0; f0 < F; f0 += 256) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int b = 0; b < B; ++b) {
      const float dv = __ldg(datt + (size_t)b * R + r);
      if (dv != 0.f) {
        const float *zb = zbar + (size_t)b * F + f0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int f = i * 32 + lane;
          if (f0 + f < F) acc[i] = fmaf(dv, __ldg(zb + f), acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int f = f0 + i * 32 + lane;
      if (f < F) dz_pn[(size_t)r * F + f] = acc[i];
    }
  }
}

// ----------------------------------------------------------------------------- K7 top-k
// Order-preserving key: larger float <=> larger uint32; every NaN, whatever its sign bit, gets the largest key
// (torch.topk treats NaN as greater than anything).
__device__ __forceinline__ uint32_t f2key(float x) {
  if (x != x) return 0xffffffffu;
  const uint32_t u = __float_as_uint(x + 0.f);               // -0 -> +0: the two compare equal, so they must tie
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Row accessors -------------------------------------------------------------
struct PlainRows {          // x [R, W], row stride ld
  const float *x;
  int W, ld;
  struct Cursor {
    const float *base;
    __device__ __forceinline__ const float *at(int j) const { return base + j; }
  };
  __device__ __forceinline__ Cursor cursor(int r) const { return Cursor{x + (size_t)r * ld}; }
  __device__ __forceinline__ int width(int) const { return W; }
  __device__ __forceinline__ const float *seg(int r, int j, int &run) const {
    run = W - j;
    return x + (size_t)r * ld + j;
  }
  // rows can be read as float4 (topk_vec_kernel)
  bool vec4_ok() const { return W % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0; }
};
struct EssenceRows {        // virtual rows over att [B,C,S]: v < B positives (class y_b), else negatives
  const float *att;
  const long long *y;
  int B, C, S;
  // negatives are the classes != y_b in class-major order: virtual element j is physical element j below the
  // label's block and j + S from it on -- one compare per element instead of a division
  struct Cursor {
    const float *base;
    int skip_from, skip;
    __device__ __forceinline__ const float *at(int j) const { return base + j + ((j >= skip_from) ? skip : 0); }
  };
  __device__ __forceinline__ Cursor cursor(int v) const {
    const int b = (v < B) ? v : v - B;
    const int yb = min(max((int)y[b], 0), C - 1);
    const float *row = att + (size_t)b * C * S;
    if (v < B) return Cursor{row + (size_t)yb * S, 0x7fffffff, 0};
    return Cursor{row, yb * S, S};
  }
  __device__ __forceinline__ int width(int v) const { return (v < B) ? S : (C - 1) * S; }
  // pointer to element j of virtual row v and the number of contiguous elements that follow it
  __device__ __forceinline__ const float *seg(int v, int j, int &run) const {
    const int b = (v < B) ? v : v - B;
    const int yb = min(max((int)y[b], 0), C - 1);   // the host rejects labels outside [0, C); never read out of bounds
    int cls, s;
    if (v < B) {
      cls = yb;
      s = j;
    } else {
      const int q = j / S;
      s = j - q * S;
      cls = q + (q >= yb ? 1 : 0);
    }
    run = S - s;
    return att + ((size_t)b * C + cls) * S + s;
  }
  bool vec4_ok() const { return S % 4 == 0 && (reinterpret_cast<uintptr_t>(att) & 15) == 0; }
};

// bitonic sort (descending) of KP (power of two, <= 1024) 64-bit composites in shared memory by NT threads
template <int NT>
__device__ __forceinline__ void bitonic_sort_desc(unsigned long long *buf, int KP, int tid) {
  for (int size = 2; size <= KP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < KP / 2; t += NT) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = buf[lo], b = buf[hi];
        if ((a < b) == desc) {
          buf[lo] = b;
          buf[hi] = a;
        }
      }
      if (NT == 32) __syncwarp(); else __syncthreads();
    }
  }
}

// One warp per row, the whole row in registers (W <= 32 E).  Exact k-th key by bitwise search with
// early exit, ordered tie handling (lowest index first), then a shared-memory bitonic sort of the winners.
template <int E, class Rows>
__global__ void __launch_bounds__(128)
topk_warp_kernel(Rows rows, int R, int k, int KP, float *__restrict__ vals, int *__restrict__ idx) {
  extern __shared__ unsigned long long sbuf[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + wib;
  if (r >= R) return;
  unsigned long long *buf = sbuf + (size_t)wib * KP;
  const int W = rows.width(r);
  uint32_t key[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int j = e * 32 + lane;
    uint32_t kv = 0u;                      // below every real key (keys of finite floats are >= 0x00800000)
    if (j < W) {
      int run;
      kv = f2key(__ldg(rows.seg(r, j, run)));
      if (kv == 0u) kv = 1u;               // keep 0 reserved for padding
    }
    key[e] = kv;
  }
  // largest T with count(key >= T) >= k
  uint32_t T = 0u;
  int cnt_ge = 0;
  bool exact = false;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c += (key[e] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) {
      T = cand;
      cnt_ge = c;
      if (c == k) { exact = true; break; }   // the winner set is already determined
    }
  }
  if (!exact && cnt_ge == 0) cnt_ge = W;     // T == 0: every element qualifies
  // winners: key > T all; key == T lowest indices first.  With `exact`, every key >= T wins.
  for (int t = lane; t < KP; t += 32) buf[t] = 0ull;
  __syncwarp();
  int base = 0;
  if (exact) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool win = key[e] >= T;
      const uint32_t m = __ballot_sync(0xffffffffu, win);
      if (win) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        buf[slot] = ((unsigned long long)key[e] << 32) | (uint32_t)(0xffffffffu - (uint32_t)(e * 32 + lane));
      }
      base += __popc(m);
    }
  } else {
    int c_gt = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) c_gt += (key[e] > T) ? 1 : 0;
    c_gt = __reduce_add_sync(0xffffffffu, c_gt);
    int need_eq = k - c_gt;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool gt = key[e] > T;
      const bool eq = key[e] == T;
      const uint32_t m_eq = __ballot_sync(0xffffffffu, eq);
      const int rank_eq = __popc(m_eq & ((1u << lane) - 1u));
      const bool win = gt || (eq && rank_eq < need_eq);
      const uint32_t m = __ballot_sync(0xffffffffu, win);
      if (win) {
        const int slot = base + __popc(m & ((1u << lane) - 1u));
        buf[slot] = ((unsigned long long)key[e] << 32) | (uint32_t)(0xffffffffu - (uint32_t)(e * 32 + lane));
      }
      base += __popc(m);
      need_eq -= min(need_eq, __popc(m_eq));
    }
  }
  __syncwarp();
  bitonic_sort_desc<32>(buf, KP, lane);
  for (int t = lane; t < k; t += 32) {
    const unsigned long long c = buf[t];
    vals[(size_t)r * k + t] = key2f((uint32_t)(c >> 32));
    idx[(size_t)r * k + t] = (int)(0xffffffffu - (uint32_t)(c & 0xffffffffu));
  }
}


// One warp per row, the row in registers as order-preserving keys (W <= 32 E), k <= 128.
// Select: MSD radix select, 8 bits per pass, on a warp-private 256-bin shared-memory histogram (at most 4 passes,
// usually 3; stops as soon as the boundary bucket is taken whole).  That is ~12 instructions per element instead
// of the ~2 x 20 of the bit-by-bit search in topk_warp_kernel.  Winners are compacted in ascending index order
// (ties: lowest index first); SORTED additionally runs a 128-element bitonic network in registers (4 per lane,
// 64-bit key|~index composites, shuffles for the cross-lane stages) to give torch.topk's descending order.
template <int SIZE, int STRIDE>
__device__ __forceinline__ void bitonic_step_reg(unsigned long long (&c)[4], int lane) {
  if (STRIDE >= 4) {
    const int lx = STRIDE >> 2;
    const bool lower = (lane & lx) == 0;
    const bool desc = (SIZE >= 128) ? true : ((lane & (SIZE >> 2)) == 0);
    const bool keep_max = (lower == desc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, c[i], lx);
      c[i] = keep_max ? (c[i] > o ? c[i] : o) : (c[i] < o ? c[i] : o);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((i & STRIDE) == 0) {
        const int p = i | STRIDE;
        bool desc;
        if (SIZE == 2) desc = ((i & 2) == 0);
        else if (SIZE == 4) desc = ((lane & 1) == 0);
        else if (SIZE >= 128) desc = true;
        else desc = ((lane & (SIZE >> 2)) == 0);
        const unsigned long long a = c[i], b = c[p];
        const bool swap = desc ? (a < b) : (a > b);
        c[i] = swap ? b : a;
        c[p] = swap ? a : b;
      }
    }
  }
}

// the same network step on 32-bit composites (one shuffle and one min/max per element across lanes)
template <int SIZE, int STRIDE>
__device__ __forceinline__ void bitonic_step_reg32(uint32_t (&c)[4], int lane) {
  if (STRIDE >= 4) {
    const int lx = STRIDE >> 2;
    const bool lower = (lane & lx) == 0;
    const bool desc = (SIZE >= 128) ? true : ((lane & (SIZE >> 2)) == 0);
    const bool keep_max = (lower == desc);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t o = __shfl_xor_sync(0xffffffffu, c[i], lx);
      c[i] = keep_max ? max(c[i], o) : min(c[i], o);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((i & STRIDE) == 0) {
        const int p = i | STRIDE;
        bool desc;
        if (SIZE == 2) desc = ((i & 2) == 0);
        else if (SIZE == 4) desc = ((lane & 1) == 0);
        else if (SIZE >= 128) desc = true;
        else desc = ((lane & (SIZE >> 2)) == 0);
        const uint32_t a = c[i], b = c[p];
        const uint32_t hi = max(a, b), lo = min(a, b);
        c[i] = desc ? hi : lo;
        c[p] = desc ? lo : hi;
      }
    }
  }
}
// the whole descending network for 32 NPL composites, NPL (1 or 2) per lane at positions lane * NPL + i: small k
template <int NPL>
__device__ __forceinline__ void bitonic_sort_desc32(uint32_t (&c)[NPL], int lane) {
  constexpr int TOT = 32 * NPL;
#pragma unroll
  for (int size = 2; size <= TOT; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= NPL) {
        const int lx = stride / NPL;
        const bool lower = (lane & lx) == 0;
        const bool desc = (size >= TOT) ? true : (((lane * NPL) & size) == 0);
        const bool keep_max = (lower == desc);
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          const uint32_t o = __shfl_xor_sync(0xffffffffu, c[i], lx);
          c[i] = keep_max ? max(c[i], o) : min(c[i], o);
        }
      } else {
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
          if ((i & stride) == 0) {
            const int p2 = i | stride;
            const bool desc = (size >= TOT) ? true : (((lane * NPL + i) & size) == 0);
            const uint32_t a = c[i], b = c[p2];
            const uint32_t hi = max(a, b), lo = min(a, b);
            c[i] = desc ? hi : lo;
            c[p2] = desc ? lo : hi;
          }
        }
      }
    }
  }
}

// order-preserving key with two integer ops: flip all bits of negatives, only the sign bit of the rest
__device__ __forceinline__ uint32_t f2key_fast(float x) {
  const uint32_t u = __float_as_uint(x + 0.f);               // -0 -> +0: the two compare equal, so they must tie
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
// hist[dg] += 1 if (key >= lo && dg < nbins), as ONE predicated shared-memory reduction (no branch)
__device__ __forceinline__ void hist_inc_if(uint32_t hist_addr, uint32_t key, uint32_t lo, uint32_t dg,
                                            uint32_t nbins) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ge.u32 q, %1, %2;\n\t"
      "setp.lt.and.u32 q, %3, %4, q;\n\t"
      "@q red.shared.add.u32 [%0], 1;\n\t"
      "}\n" ::"r"(hist_addr + (dg << 2)),
      "r"(key), "r"(lo), "r"(dg), "r"(nbins)
      : "memory");
}

// buf[slot] = (~index, key) if p, as one predicated store (no branch)
__device__ __forceinline__ void st_shared_v2_if(uint32_t addr, uint32_t x, uint32_t y, bool p) {
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "setp.ne.b32 q, %3, 0;\n\t"
      "@q st.shared.v2.u32 [%0], {%1, %2};\n\t"
      "}\n" ::"r"(addr),
      "r"(x), "r"(y), "r"((uint32_t)p)
      : "memory");
}

// FULL: W == 32 E exactly (no bounds checks, no padding keys).
// One row by one warp; `hist` = 256 words of warp-private shared memory.  INLINE = false: the rare-path fallback of
// topk_vec_kernel (re-reads the row from global memory, kept out of line so that it does not cost the fast path
// registers).
template <int E, bool FULL, bool SORTED, class Rows>
__device__ __forceinline__ void radix_select_row(const Rows &rows, int r, int k, unsigned int *hist, int lane,
                                                 float *__restrict__ vals, int *__restrict__ idx) {
  uint2 *buf = reinterpret_cast<uint2 *>(hist);              // .x = ~index, .y = key  (little-endian key|~index)
  const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(hist);
  const int W = FULL ? E * 32 : rows.width(r);
  const typename Rows::Cursor cur = rows.cursor(r);
  uint32_t key[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int j = e * 32 + lane;
    if (FULL || j < W)
      key[e] = max(f2key(__ldg(cur.at(j))), 1u);             // 0 stays reserved for padding (NaNs of either sign: top key)
    else
      key[e] = 0u;
  }
  // ---- range-adaptive radix select: after the loop every key > T wins and `need` keys == T (or, if `exact`,
  //      every key >= T) win.  Digits are taken from (key - lo) over the row's own [lo, hi] key range, 8 bits per
  //      pass from the top of that range, so the row spreads over the 256 bins whatever its exponent range is.
  uint32_t kmin = 0xffffffffu, kmax = 1u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    kmin = min(kmin, (FULL || key[e] != 0u) ? key[e] : 0xffffffffu);
    kmax = max(kmax, key[e]);
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  int shift = 32 - __clz((kmax - kmin) | 1u) - 8;            // ((kmax - kmin) >> shift) < 256
  if (shift < 0) shift = 0;
  uint32_t lo = kmin;                                        // candidates: keys in [lo, lo + (nbins << shift))
  uint32_t nbins = 256u;
  int need = k;
  bool exact = false;
#pragma unroll 1
  for (;;) {
    reinterpret_cast<uint4 *>(hist)[lane * 2] = make_uint4(0u, 0u, 0u, 0u);
    reinterpret_cast<uint4 *>(hist)[lane * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
#pragma unroll
    for (int e = 0; e < E; ++e) hist_inc_if(hist_addr, key[e], lo, (key[e] - lo) >> shift, nbins);
    __syncwarp();
    // lane l owns bins 8 (31 - l) .. 8 (31 - l) + 7, walked from the top: lane 0 holds the 8 largest digits
    const int base = (31 - lane) * 8;
    const uint4 hlo = reinterpret_cast<const uint4 *>(hist)[(31 - lane) * 2];
    const uint4 hhi = reinterpret_cast<const uint4 *>(hist)[(31 - lane) * 2 + 1];
    const int c[8] = {(int)hhi.w, (int)hhi.z, (int)hhi.y, (int)hhi.x, (int)hlo.w, (int)hlo.z, (int)hlo.y, (int)hlo.x};
    const int t = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
    int incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const uint32_t hit = __ballot_sync(0xffffffffu, incl >= need);
    const int L = __ffs(hit) - 1;                            // candidates number >= need: a lane always hits
    int packed = 0;                                          // bin | cntb << 8 | rem << 20
    if (lane == L) {
      int rem = need - (incl - t);
      int bin = 0, cntb = 0;
      bool found = false;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (!found) {
          if (c[i] >= rem) {
            bin = base + 7 - i;
            cntb = c[i];
            found = true;
          } else {
            rem -= c[i];
          }
        }
      }
      packed = bin | (cntb << 8) | (rem << 20);              // counts <= 2048 fit 12 bits
    }
    packed = __shfl_sync(0xffffffffu, packed, L);
    const int bin = packed & 255, cntb = (packed >> 8) & 4095;
    need = packed >> 20;
    lo += (uint32_t)bin << shift;                            // the boundary bucket is [lo, lo + (1 << shift))
    if (cntb == need) {                                      // whole boundary bucket taken: every key >= lo wins
      exact = true;
      break;
    }
    if (shift == 0) break;                                   // bucket == one key value: `need` of its duplicates win
    nbins = (shift >= 8) ? 256u : (1u << shift);
    shift = (shift > 8) ? shift - 8 : 0;                     // split the bucket into (up to) 256 sub-buckets
  }
  const uint32_t T = lo;
  __syncwarp();
  // ---- compaction in ascending index order (e-major, then lane): slot = number of winners before this element
  {
    uint4 *b4 = reinterpret_cast<uint4 *>(hist);
    b4[lane * 2] = make_uint4(0u, 0u, 0u, 0u);
    b4[lane * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncwarp();
  const uint32_t lt_mask = (1u << lane) - 1u;
  int basec = 0;
  if (exact) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const bool win = key[e] >= T && (FULL || key[e] != 0u);
      const uint32_t m = __ballot_sync(0xffffffffu, win);
      st_shared_v2_if(hist_addr + 8u * (uint32_t)(basec + __popc(m & lt_mask)), ~(uint32_t)(e * 32 + lane), key[e], win);
      basec += __popc(m);
    }
  } else {
    int need_eq = need;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t kv = key[e];
      const bool eq = (kv == T);
      const uint32_t m_eq = __ballot_sync(0xffffffffu, eq);
      const bool win = (kv > T) || (eq && __popc(m_eq & lt_mask) < need_eq);
      const uint32_t m = __ballot_sync(0xffffffffu, win);
      if (win) buf[basec + __popc(m & lt_mask)] = make_uint2(~(uint32_t)(e * 32 + lane), kv);
      basec += __popc(m);
      need_eq -= min(need_eq, __popc(m_eq));
    }
  }
  __syncwarp();
  if (SORTED) {
    unsigned long long c4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) c4[i] = reinterpret_cast<const unsigned long long *>(hist)[lane * 4 + i];
    bitonic_step_reg<2, 1>(c4, lane);
    bitonic_step_reg<4, 2>(c4, lane);   bitonic_step_reg<4, 1>(c4, lane);
    bitonic_step_reg<8, 4>(c4, lane);   bitonic_step_reg<8, 2>(c4, lane);   bitonic_step_reg<8, 1>(c4, lane);
    bitonic_step_reg<16, 8>(c4, lane);  bitonic_step_reg<16, 4>(c4, lane);  bitonic_step_reg<16, 2>(c4, lane);
    bitonic_step_reg<16, 1>(c4, lane);
    bitonic_step_reg<32, 16>(c4, lane); bitonic_step_reg<32, 8>(c4, lane);  bitonic_step_reg<32, 4>(c4, lane);
    bitonic_step_reg<32, 2>(c4, lane);  bitonic_step_reg<32, 1>(c4, lane);
    bitonic_step_reg<64, 32>(c4, lane); bitonic_step_reg<64, 16>(c4, lane); bitonic_step_reg<64, 8>(c4, lane);
    bitonic_step_reg<64, 4>(c4, lane);  bitonic_step_reg<64, 2>(c4, lane);  bitonic_step_reg<64, 1>(c4, lane);
    bitonic_step_reg<128, 64>(c4, lane); bitonic_step_reg<128, 32>(c4, lane); bitonic_step_reg<128, 16>(c4, lane);
    bitonic_step_reg<128, 8>(c4, lane);  bitonic_step_reg<128, 4>(c4, lane);  bitonic_step_reg<128, 2>(c4, lane);
    bitonic_step_reg<128, 1>(c4, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pos = lane * 4 + i;
      if (pos < k) {
        vals[(size_t)r * k + pos] = key2f((uint32_t)(c4[i] >> 32));
        idx[(size_t)r * k + pos] = (int)~(uint32_t)(c4[i] & 0xffffffffu);
      }
    }
  } else {
    for (int t2 = lane; t2 < k; t2 += 32) {
      const uint2 cc = buf[t2];
      vals[(size_t)r * k + t2] = key2f(cc.y);
      idx[(size_t)r * k + t2] = (int)~cc.x;
    }
  }
}

template <int E, bool FULL, bool SORTED, class Rows>
__global__ void __launch_bounds__(128)
topk_warp_radix_kernel(Rows rows, int R, int k, float *__restrict__ vals, int *__restrict__ idx) {
  __shared__ __align__(16) unsigned int s_hist[4][256];      // per warp: histogram, later 128 64-bit winners
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + wib;
  if (r >= R) return;
  radix_select_row<E, FULL, SORTED, Rows>(rows, r, k, s_hist[wib], lane, vals, idx);
}
template <int E, bool SORTED, class Rows>
__device__ __noinline__ void radix_select_row_slow(const Rows &rows, int r, int k, unsigned int *hist, int lane,
                                                   float *__restrict__ vals, int *__restrict__ idx) {
  radix_select_row<E, false, SORTED, Rows>(rows, r, k, hist, lane, vals, idx);
}

// The winners of one row -- 64-bit (~index, value bits) pairs in buf[0, k), zeros behind them when SORTED -- to global
// memory, by one warp: SORTED in torch.topk's order (descending value, ties by ascending index), else as they are.
template <bool SORTED>
__device__ __forceinline__ void emit_winners(uint2 *buf, int k, int lane, uint32_t tkey_in, uint32_t xkey_in,
                                             float *__restrict__ vrow, int *__restrict__ irow) {
  if (SORTED) {
    // Fast path: 32-bit composites (key - key(T)) << 7 | (127 - slot) when the winners' keys span < 2^25 (about four
    // binades above T): one shuffle and one min/max per element and step instead of a 64-bit compare-select.  Slots
    // are in lane order, so equal keys would come out in the wrong order: any duplicate key among the winners (and a
    // wider key range) falls through to the 64-bit key | ~index network below.
    const uint32_t tkey = tkey_in, xkey = xkey_in;
    bool sorted_done = false;
    auto small_sort = [&](auto npl_tag) {                    // k <= 32 NPL: NPL composites per lane
      constexpr int NPL = decltype(npl_tag)::value;
      uint32_t c[NPL];
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int sl = lane * NPL + i;
        const uint32_t kk = f2key_fast(__uint_as_float(buf[sl].y)) - tkey;
        c[i] = (sl < k) ? ((kk << 7) | (uint32_t)(127 - sl)) : 0u;
      }
      bitonic_sort_desc32<NPL>(c, lane);
      const uint32_t nxt0 = __shfl_down_sync(0xffffffffu, c[0], 1);
      bool dup = false;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const uint32_t nx = (i < NPL - 1) ? c[(i + 1) % NPL] : nxt0;
        const bool last = (i == NPL - 1) && (lane == 31);
        dup = dup || (!last && lane * NPL + i + 1 < k && (c[i] >> 7) == (nx >> 7));
      }
      if (__any_sync(0xffffffffu, dup)) return false;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int p2 = lane * NPL + i;
        if (p2 < k) {
          const uint2 w = buf[127 - (int)(c[i] & 127u)];
          vrow[p2] = __uint_as_float(w.y);
          irow[p2] = (int)~w.x;
        }
      }
      return true;
    };
    const bool narrow = xkey - tkey < (1u << 25);
    if (narrow && k <= 32) {
      sorted_done = small_sort(std::integral_constant<int, 1>{});
    } else if (narrow && k <= 64) {
      sorted_done = small_sort(std::integral_constant<int, 2>{});
    } else if (narrow) {
      uint32_t c[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int sl = lane * 4 + i;
        const uint32_t kk = f2key_fast(__uint_as_float(buf[sl].y)) - tkey;
        c[i] = (sl < k) ? ((kk << 7) | (uint32_t)(127 - sl)) : 0u;
      }
      bitonic_step_reg32<2, 1>(c, lane);
      bitonic_step_reg32<4, 2>(c, lane);   bitonic_step_reg32<4, 1>(c, lane);
      bitonic_step_reg32<8, 4>(c, lane);   bitonic_step_reg32<8, 2>(c, lane);   bitonic_step_reg32<8, 1>(c, lane);
      bitonic_step_reg32<16, 8>(c, lane);  bitonic_step_reg32<16, 4>(c, lane);  bitonic_step_reg32<16, 2>(c, lane);
      bitonic_step_reg32<16, 1>(c, lane);
      bitonic_step_reg32<32, 16>(c, lane); bitonic_step_reg32<32, 8>(c, lane);  bitonic_step_reg32<32, 4>(c, lane);
      bitonic_step_reg32<32, 2>(c, lane);  bitonic_step_reg32<32, 1>(c, lane);
      bitonic_step_reg32<64, 32>(c, lane); bitonic_step_reg32<64, 16>(c, lane); bitonic_step_reg32<64, 8>(c, lane);
      bitonic_step_reg32<64, 4>(c, lane);  bitonic_step_reg32<64, 2>(c, lane);  bitonic_step_reg32<64, 1>(c, lane);
      bitonic_step_reg32<128, 64>(c, lane); bitonic_step_reg32<128, 32>(c, lane); bitonic_step_reg32<128, 16>(c, lane);
      bitonic_step_reg32<128, 8>(c, lane);  bitonic_step_reg32<128, 4>(c, lane);  bitonic_step_reg32<128, 2>(c, lane);
      bitonic_step_reg32<128, 1>(c, lane);
      // duplicates: position p and p + 1 carry the same key (p < k - 1)
      const uint32_t nxt0 = __shfl_down_sync(0xffffffffu, c[0], 1);
      bool dup = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t nx = (i < 3) ? c[i + 1] : nxt0;
        const bool last = (i == 3) && (lane == 31);
        dup = dup || (!last && lane * 4 + i + 1 < k && (c[i] >> 7) == (nx >> 7));
      }
      if (!__any_sync(0xffffffffu, dup)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int p2 = lane * 4 + i;
          if (p2 < k) {
            const uint2 w = buf[127 - (int)(c[i] & 127u)];
            vrow[p2] = __uint_as_float(w.y);
            irow[p2] = (int)~w.x;
          }
        }
        sorted_done = true;
      }
    }
    if (!sorted_done) {
    // composites key | ~index (value bits -> order-preserving key; empty slots stay 0 = below every key; -0 and +0
    // get the same key so that they order by index)
    unsigned long long c4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint2 w = buf[lane * 4 + i];
      const uint32_t kk = (lane * 4 + i < k) ? f2key_fast(__uint_as_float(w.y)) : 0u;
      c4[i] = ((unsigned long long)kk << 32) | w.x;
    }
    bitonic_step_reg<2, 1>(c4, lane);
    bitonic_step_reg<4, 2>(c4, lane);   bitonic_step_reg<4, 1>(c4, lane);
    bitonic_step_reg<8, 4>(c4, lane);   bitonic_step_reg<8, 2>(c4, lane);   bitonic_step_reg<8, 1>(c4, lane);
    bitonic_step_reg<16, 8>(c4, lane);  bitonic_step_reg<16, 4>(c4, lane);  bitonic_step_reg<16, 2>(c4, lane);
    bitonic_step_reg<16, 1>(c4, lane);
    bitonic_step_reg<32, 16>(c4, lane); bitonic_step_reg<32, 8>(c4, lane);  bitonic_step_reg<32, 4>(c4, lane);
    bitonic_step_reg<32, 2>(c4, lane);  bitonic_step_reg<32, 1>(c4, lane);
    bitonic_step_reg<64, 32>(c4, lane); bitonic_step_reg<64, 16>(c4, lane); bitonic_step_reg<64, 8>(c4, lane);
    bitonic_step_reg<64, 4>(c4, lane);  bitonic_step_reg<64, 2>(c4, lane);  bitonic_step_reg<64, 1>(c4, lane);
    bitonic_step_reg<128, 64>(c4, lane); bitonic_step_reg<128, 32>(c4, lane); bitonic_step_reg<128, 16>(c4, lane);
    bitonic_step_reg<128, 8>(c4, lane);  bitonic_step_reg<128, 4>(c4, lane);  bitonic_step_reg<128, 2>(c4, lane);
    bitonic_step_reg<128, 1>(c4, lane);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int p2 = lane * 4 + i;
      if (p2 < k) {
        const int id = (int)~(uint32_t)(c4[i] & 0xffffffffu);
        vrow[p2] = key2f((uint32_t)(c4[i] >> 32));
        irow[p2] = id;
      }
    }
    }
  } else {
    // (row pointers advanced by the lane ONCE and made opaque: otherwise the compiler rebuilds both 64-bit addresses
    //  inside each of the four predicated stores, 16 instructions a piece)
    float *vp = vrow + lane;
    int *ip = irow + lane;
    asm volatile("" : "+l"(vp), "+l"(ip));
    const uint2 *bp = buf + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (lane + 32 * i < k) {
        const uint2 cc = bp[32 * i];
        vp[32 * i] = __uint_as_float(cc.y);
        ip[32 * i] = (int)~cc.x;
      }
    }
  }
}

#include "topk_sift.cuh"

// ----------------------------------------------------------------------------- K7b vectorised warp select
// One warp per row, uniform row width W (a multiple of 4, rows 16-byte aligned), k <= 128.  The select is bound by
// instruction issue and the integer pipe, not by HBM, so it is built around the instruction count:
//   * 128-bit loads: lane l holds float4 number i * 32 + l of the row for i < FI (+ one partial iteration);
//   * ONE histogram pass over 256 bins that are uniform in VALUE between the row's min and max: bin = the low bits of
//     fma(x - min, scale, 2^23), i.e. FADD + FFMA (fma pipe) + mask + address + red per element.  (Bins uniform in the
//     order-preserving integer key are one binade wide around 1.0: 110 of N(0,1)'s 800 values share the threshold's bin
//     and two more passes are needed; value bins leave 4 or 5.)
//   * the threshold bin's values are collected (lane-local lists, one warp prefix) and ranked against each other with
//     shuffles: T = the need-th largest of them;
//   * winners (x >= T) are compacted lane-locally (count, one warp prefix, predicated stores) instead of one ballot and
//     two popcounts per element -- the unsorted output order is by lane, which edrl_topk_rows(sorted = 0) leaves
//     unspecified; ties at T (fewer wanted than present) take a ballot path that keeps the lowest indices;
//   * SORTED: the winners go through the register bitonic network as key | ~index composites.
// Rows holding NaN or infinities, constant rows, a threshold bin with more than 32 values or a lane with more than 4
// of them take radix_select_row (the integer-key select above) instead.  Comparisons are on float values: -0 and +0
// tie (lowest index first), as in torch.topk.
// FI: full iterations (W / 128), PARTIAL: one more with lanes < (W / 4) % 32.
constexpr int TV_LCAP = 8;                    // threshold-bin values a lane may hold before the row takes the slow path
constexpr int TV_HIST = 264;                  // winners' buffer: 128 x 8 bytes; also bins + overflow / dummy bins
constexpr int TV_CSTRIDE = 33;                // words per lane of the candidate staging area (a bin taken here holds <= 32; odd)
template <int FI, bool PARTIAL, bool SORTED, class Rows>
__global__ void __launch_bounds__(128, (FI <= 6 && !SORTED) ? 8 : 4)
topk_vec_kernel(Rows rows, int R, int W, int k, float *__restrict__ vals, int *__restrict__ idx) {
  constexpr int NI = FI + (PARTIAL ? 1 : 0);
  constexpr int E = NI * 4;
  // value bins of the histogram pass: about one threshold-bin candidate per 50 row elements keeps the ranked list short
  // (<= 32), while fewer bins mean fewer distinct (bin, bank) pairs per warp increment -- the select is bound by
  // shared-memory wavefronts
  constexpr int BINS = (NI <= 7) ? 64 : ((NI <= 14) ? 128 : 256);
  __shared__ __align__(16) unsigned int s_hist[4][TV_HIST];  // per warp: histogram, later 128 64-bit winners
  __shared__ unsigned int s_cand[4][32 * TV_CSTRIDE];        // per warp: lane-local candidate lists, then the compact list
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + wib;
  if (r >= R) return;
  unsigned int *hist = s_hist[wib];
  uint2 *buf = reinterpret_cast<uint2 *>(hist);              // .x = ~index, .y = value bits
  const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(hist);
  const typename Rows::Cursor cur = rows.cursor(r);
  const bool pvalid = PARTIAL && lane < ((W >> 2) & 31);     // lanes of the partial iteration that hold data
  // element e of lane l is row element ((e >> 2) * 32 + l) * 4 + (e & 3); slots of the partial iteration count only
  // where pvalid
  float x[E];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < FI || pvalid) v = __ldg(reinterpret_cast<const float4 *>(cur.at((i * 32 + lane) * 4)));
    x[4 * i + 0] = v.x;
    x[4 * i + 1] = v.y;
    x[4 * i + 2] = v.z;
    x[4 * i + 3] = v.w;
  }
  // min, max and a NaN / infinity detector (x * 0 is NaN for either)
  float xmin = x[0], xmax = x[0], nf = 0.f;
  if (FI == 0) {
    xmin = pvalid ? x[0] : __int_as_float(0x7f800000);
    xmax = pvalid ? x[0] : __int_as_float(0xff800000);
  }
#pragma unroll
  for (int e = 0; e < FI * 4; ++e) {
    xmin = fminf(xmin, x[e]);
    xmax = fmaxf(xmax, x[e]);
    nf = fmaf(x[e], 0.f, nf);
  }
  if (PARTIAL && pvalid) {
#pragma unroll
    for (int e = FI * 4; e < E; ++e) {
      xmin = fminf(xmin, x[e]);
      xmax = fmaxf(xmax, x[e]);
      nf = fmaf(x[e], 0.f, nf);
    }
  }
  {
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, f2key_fast(xmin));
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, f2key_fast(xmax));
    xmin = key2f(kmin);
    xmax = key2f(kmax);
  }
  const float range = xmax - xmin;
  const bool irregular = __any_sync(0xffffffffu, !(nf == 0.f)) || !(range > 0.f) || !(range <= 3.0e38f);
  bool done = false, exact = false;
  int need = k;
  float T = 0.f;
  if (!irregular) {
    // ---- one histogram pass over value bins: bin(x) = round((x - xmin) * 255 / range) in [0, 255] ----
    // y = (x - xmin) scale + 2^23 in [2^23, 2^23 + 255.5): bin = y's low bits.  (x - xmin >= 0 exactly; folding xmin into
    // the addend would round it to an integer and push the smallest values below 2^23.)
    const float scale = (float)(BINS - 1) / range;
    const float off = 8388608.f;
    for (int i = lane; i < (BINS + 8) / 4; i += 32) reinterpret_cast<uint4 *>(hist)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    const uint32_t dummy = hist_addr + (uint32_t)(BINS + 4) * 4u;
    // y's bits are 0x4b000000 + bin (bin <= BINS: rounding may reach one past the last, counted apart), so the bin's
    // address is one shift-add away: 4 y + (hist - 4 * 0x4b000000) in 32-bit wrap-around arithmetic
    const uint32_t hbase = hist_addr - (0x4b000000u << 2);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t yb = __float_as_uint(fmaf(x[e] - xmin, scale, off));
      uint32_t addr = (yb << 2) + hbase;
      if (e >= FI * 4) addr = pvalid ? addr : dummy;
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
    }
    __syncwarp();
    // lane l owns bins BPL (31 - l) .. BPL (31 - l) + BPL - 1, walked from the top: lane 0 holds the largest
    constexpr int BPL = BINS / 32;
    const int base = (31 - lane) * BPL;
    int c[BPL];
    int t = 0;
#pragma unroll
    for (int i = 0; i < BPL; ++i) {
      c[i] = (int)hist[base + BPL - 1 - i];
      t += c[i];
    }
    const int over = (int)hist[BINS];                     // bin BINS (fp rounding at the very top): above the last
    int incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    incl += over;
    const uint32_t hit = __ballot_sync(0xffffffffu, incl >= need);
    const int L = __ffs(hit) - 1;
    int packed = 0;                                          // bin | cntb << 9 | rem << 21
    if (lane == L) {
      int rem = need - (incl - t);
      int bn = 0, cb = 0;
      bool found = false;
#pragma unroll
      for (int i = 0; i < BPL; ++i) {
        if (!found) {
          if (c[i] >= rem) {
            bn = base + BPL - 1 - i;
            cb = c[i];
            found = true;
          } else {
            rem -= c[i];
          }
        }
      }
      packed = bn | (cb << 9) | (rem << 21);
    }
    // (need <= over cannot happen for k <= W unless bin 256 alone holds k values: then rem <= 0 and cb = 0 below)
    packed = __shfl_sync(0xffffffffu, packed, L < 0 ? 0 : L);
    const int bin = packed & 511, cntb = (packed >> 9) & 4095;
    const int rem = packed >> 21;
    if (L >= 0 && rem >= 1 && cntb >= rem && cntb <= 32) {
      // ---- the threshold bin's values, collected and ranked: T = its rem-th largest ----
      const uint32_t ytarget = 0x4b000000u + (uint32_t)bin;
      unsigned int *mine = s_cand[wib] + lane * TV_CSTRIDE;
      uint32_t lp = (uint32_t)__cvta_generic_to_shared(mine);
      const uint32_t lp0 = lp;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        // (compare, predicated store and predicated pointer bump in one asm block: passing a bool in costs a select and
        //  a second compare per element)
        const uint32_t yb = __float_as_uint(fmaf(x[e] - xmin, scale, off));
        const uint32_t yt = (e < FI * 4 || pvalid) ? ytarget : 0u;         // never equal for lanes without data
        asm volatile(
            "{\n\t"
            ".reg .pred q;\n\t"
            "setp.eq.u32 q, %2, %3;\n\t"
            "@q st.shared.b32 [%0], %1;\n\t"
            "@q add.u32 %0, %0, 4;\n\t"
            "}\n"
            : "+r"(lp)
            : "r"(__float_as_uint(x[e])), "r"(yb), "r"(yt)
            : "memory");
      }
      const int mycnt = (int)((lp - lp0) >> 2);
      int cincl = mycnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, cincl, o);
        if (lane >= o) cincl += v;
      }
      const uint32_t more = __ballot_sync(0xffffffffu, mycnt > TV_LCAP);
      if (more == 0u) {
        __syncwarp();
        uint32_t mk[TV_LCAP];
#pragma unroll
        for (int c2 = 0; c2 < TV_LCAP; ++c2) mk[c2] = (c2 < mycnt) ? mine[c2] : 0u;
        __syncwarp();
        unsigned int *clist = s_cand[wib];                   // compact list (overwrites the lane lists: read first)
        const int coff = cincl - mycnt;
#pragma unroll
        for (int c2 = 0; c2 < TV_LCAP; ++c2)
          if (c2 < mycnt) clist[coff + c2] = mk[c2];
        __syncwarp();
        const float ci = (lane < cntb) ? __uint_as_float(clist[lane]) : 0.f;
        int gt = 0, eqb = 0, eqt = 0;
        for (int j = 0; j < cntb; ++j) {
          const float cj = __shfl_sync(0xffffffffu, ci, j);
          gt += (cj > ci) ? 1 : 0;
          eqt += (cj == ci) ? 1 : 0;
          eqb += (cj == ci && j < lane) ? 1 : 0;
        }
        // the candidate of rank rem - 1 (ties broken by list position) carries the threshold
        const uint32_t sel = __ballot_sync(0xffffffffu, lane < cntb && gt + eqb == rem - 1);
        const int src = __ffs(sel) - 1;
        T = __shfl_sync(0xffffffffu, ci, src);
        const int gt_t = __shfl_sync(0xffffffffu, gt, src);
        const int eq_t = __shfl_sync(0xffffffffu, eqt, src);
        need = rem - gt_t;                                   // wanted among the values == T
        exact = (eq_t == need);
        done = true;
      }
    }
  }
  if (!done) {
    // NaN / infinity, constant row, or a crowded threshold bin: the integer-key radix select, from global memory
    radix_select_row_slow<E, SORTED, Rows>(rows, r, k, hist, lane, vals, idx);
    return;
  }
  __syncwarp();
  if (SORTED) {                                              // the sort reads all 128 slots: empty ones must be 0
    uint4 *b4 = reinterpret_cast<uint4 *>(hist);
    b4[lane * 2] = make_uint4(0u, 0u, 0u, 0u);
    b4[lane * 2 + 1] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
  }
  const uint32_t lane4 = (uint32_t)lane << 2;
  if (exact) {
    // ---- lane-local compaction: every x >= T wins ----
    // (the win flags are kept in an explicit bit mask: the compiler otherwise rebuilds one with three instructions
    //  per element to share the compares between the counting and the storing sweep)
    uint32_t wm[(E + 31) / 32];
#pragma unroll
    for (int w2 = 0; w2 < (E + 31) / 32; ++w2) wm[w2] = 0u;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      // lanes without data in the partial iteration compare against +inf (rows here are finite: never true)
      const float te = (e < FI * 4 || pvalid) ? T : __int_as_float(0x7f800000);
      asm("{\n\t"
          ".reg .pred q;\n\t"
          "setp.ge.f32 q, %1, %2;\n\t"
          "@q or.b32 %0, %0, %3;\n\t"
          "}\n"
          : "+r"(wm[e >> 5])
          : "f"(x[e]), "f"(te), "r"(1u << (e & 31)));
    }
    int cnt = 0;
#pragma unroll
    for (int w2 = 0; w2 < (E + 31) / 32; ++w2) cnt += __popc(wm[w2]);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    uint32_t pos = hist_addr + 8u * (uint32_t)(incl - cnt);
    const uint32_t nlane4 = ~lane4;                          // ~(lane4 + c) = ~lane4 - c
#pragma unroll
    for (int e = 0; e < E; ++e) {
      asm volatile(
          "{\n\t"
          ".reg .pred q;\n\t"
          ".reg .b32 t;\n\t"
          "and.b32 t, %3, %4;\n\t"
          "setp.ne.b32 q, t, 0;\n\t"
          "@q st.shared.v2.u32 [%0], {%1, %2};\n\t"
          "@q add.u32 %0, %0, 8;\n\t"
          "}\n"
          : "+r"(pos)
          : "r"(nlane4 - (uint32_t)((e >> 2) * 128 + (e & 3))), "r"(__float_as_uint(x[e])), "r"(wm[e >> 5]),
            "r"(1u << (e & 31))
          : "memory");
    }
  } else {
    // ---- ties at T: every x > T wins, and the `need` lowest-index elements with x == T ----
    const uint32_t lt_mask = (1u << lane) - 1u;
    int basec = 0, need_eq = need;
#pragma unroll 1
    for (int i = 0; i < NI; ++i) {
      const bool ok = (i < FI) || pvalid;
      uint32_t m_eq[4], m_gt[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float xv = 0.f;
#pragma unroll
        for (int ii = 0; ii < NI; ++ii)
          if (ii == i) xv = x[4 * ii + cc];
        m_eq[cc] = __ballot_sync(0xffffffffu, ok && xv == T);
        m_gt[cc] = __ballot_sync(0xffffffffu, ok && xv > T);
      }
      // index order inside the iteration: lane major, then component
      const int eq_before_lane = __popc(m_eq[0] & lt_mask) + __popc(m_eq[1] & lt_mask) + __popc(m_eq[2] & lt_mask) +
                                 __popc(m_eq[3] & lt_mask);
      const int eq_total = __popc(m_eq[0]) + __popc(m_eq[1]) + __popc(m_eq[2]) + __popc(m_eq[3]);
      int my_eq = 0, my_win[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const bool eq = (m_eq[cc] >> lane) & 1u, gt = (m_gt[cc] >> lane) & 1u;
        my_win[cc] = (gt || (eq && eq_before_lane + my_eq < need_eq)) ? 1 : 0;
        my_eq += eq ? 1 : 0;
      }
      const int my_cnt = my_win[0] + my_win[1] + my_win[2] + my_win[3];
      int incl = my_cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      int slot = basec + incl - my_cnt;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float xv = 0.f;
#pragma unroll
        for (int ii = 0; ii < NI; ++ii)
          if (ii == i) xv = x[4 * ii + cc];
        if (my_win[cc]) buf[slot++] = make_uint2(~(lane4 + (uint32_t)(i * 128 + cc)), __float_as_uint(xv));
      }
      basec += __shfl_sync(0xffffffffu, incl, 31);
      need_eq -= min(need_eq, eq_total);
    }
  }
  __syncwarp();
  emit_winners<SORTED>(buf, k, lane, f2key_fast(T), f2key_fast(xmax), vals + (size_t)r * k, idx + (size_t)r * k);
}

// ----------------------------------------------------------------------------- K7c vectorised block select (wide rows)
// One 256-thread block per row for 2048 < W <= 8192 (uniform width, W % 4 == 0, 16-byte aligned rows, k <= 128): the
// same plan as topk_vec_kernel with the row spread over 8 warps -- thread t holds float4 number i * 256 + t (i < NI4) in
// registers; one histogram pass over 256 value bins in shared memory, the threshold bin's values collected and ranked
// by warp 0, winners compacted through a per-thread win mask and a block prefix.  Rows with NaN / infinities, constant
// rows and crowded threshold bins run the multi-pass radix select on order-preserving integer keys instead (same
// registers, further histogram passes); ties at the threshold keep the lowest indices (an index-ordered block scan).
constexpr int VB_THREADS = 256;
struct VbCtl {
  unsigned int cand_n, bin, cntb, rem, exact, tkey, need, pad;
  unsigned int kmin[8], kmax[8], bad[8], wsum[8];
};
// exclusive prefix of v over the 256 threads of the block (and the block total); two barriers
__device__ __forceinline__ int vb_excl_scan(int v, int &total, unsigned int *wsum, int warp, int lane) {
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) wsum[warp] = (unsigned int)incl;
  __syncthreads();
  int before = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    const int s = (int)wsum[w];
    before += (w < warp) ? s : 0;
    total += s;
  }
  __syncthreads();
  return before + incl - v;
}
// warp 0: walk the 256-bin histogram from the top; the bin that holds the need-th largest, its count, and how many of
// its values are wanted
__device__ __forceinline__ void vb_scan_hist(const unsigned int *hist, int need, int lane, VbCtl *ctl) {
  const int base = (31 - lane) * 8;
  const uint4 hlo = reinterpret_cast<const uint4 *>(hist)[(31 - lane) * 2];
  const uint4 hhi = reinterpret_cast<const uint4 *>(hist)[(31 - lane) * 2 + 1];
  const int c[8] = {(int)hhi.w, (int)hhi.z, (int)hhi.y, (int)hhi.x, (int)hlo.w, (int)hlo.z, (int)hlo.y, (int)hlo.x};
  const int over = (int)hist[256];
  const int t = ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
  int incl = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  incl += over;
  const uint32_t hit = __ballot_sync(0xffffffffu, incl >= need);
  const int L = __ffs(hit) - 1;
  if (L < 0) {                                               // (cannot happen for k <= W; leave a value that fails the checks)
    if (lane == 0) { ctl->bin = 0; ctl->cntb = 0; ctl->rem = 0; }
    return;
  }
  if (lane == L) {
    int rem = need - (incl - t);
    int bn = 0, cb = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (!found) {
        if (c[i] >= rem) {
          bn = base + 7 - i;
          cb = c[i];
          found = true;
        } else {
          rem -= c[i];
        }
      }
    }
    ctl->bin = (unsigned int)bn;
    ctl->cntb = (unsigned int)cb;
    ctl->rem = (unsigned int)(rem > 0 ? rem : 0);
  }
}

template <int NI4, bool SORTED, class Rows>
__device__ __forceinline__ void vecblock_select_row(const Rows &rows, int r, int W, int k, float *__restrict__ vals,
                                                    int *__restrict__ idx) {
  constexpr int E = NI4 * 4;
  static_assert(E <= 32, "one 32-bit win mask per thread");
  __shared__ __align__(16) unsigned int s_hist[264];
  __shared__ __align__(16) uint2 s_buf[128];
  __shared__ unsigned int s_cand[64];
  __shared__ VbCtl s_ctl;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  VbCtl *ctl = &s_ctl;
  const typename Rows::Cursor cur = rows.cursor(r);
  const int W4 = W >> 2;
  float x[E];
  uint32_t vmask = 0u;                                       // bit e: slot e holds data
#pragma unroll
  for (int i = 0; i < NI4; ++i) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i * VB_THREADS + tid < W4) {
      v = __ldg(reinterpret_cast<const float4 *>(cur.at((i * VB_THREADS + tid) * 4)));
      vmask |= 0xfu << (4 * i);
    }
    x[4 * i + 0] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
  }
  // ---- row statistics: a NaN / infinity detector, then the key range ----
  float nf = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e)
    if ((vmask >> e) & 1u) nf = fmaf(x[e], 0.f, nf);
  const bool wbad = __any_sync(0xffffffffu, !(nf == 0.f));
  if (wbad) {
    // NaNs of either sign become the canonical positive NaN: the largest key, as torch.topk orders them
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (x[e] != x[e]) x[e] = __int_as_float(0x7fc00000);
  }
  uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if ((vmask >> e) & 1u) {
      const uint32_t kk = f2key_fast(x[e]);
      kmin = min(kmin, kk);
      kmax = max(kmax, kk);
    }
  }
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) {
    ctl->kmin[warp] = kmin;
    ctl->kmax[warp] = kmax;
    ctl->bad[warp] = wbad ? 1u : 0u;
  }
  if (tid < 128) s_buf[tid] = make_uint2(0u, 0u);
  if (tid == 0) ctl->cand_n = 0u;
  __syncthreads();
  bool bad = false;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    kmin = min(kmin, ctl->kmin[w]);
    kmax = max(kmax, ctl->kmax[w]);
    bad = bad || (ctl->bad[w] != 0u);
  }
  const float xmin = key2f(kmin), xmax = key2f(kmax);
  const float range = xmax - xmin;
  const bool irregular = bad || !(range > 0.f) || !(range <= 3.0e38f);
  const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(s_hist);
  const uint32_t dummy = hist_addr + 260u * 4u;
  int need = k;
  bool done = false, exact = false;
  uint32_t tkey = 0u;

  if (!irregular) {
    // ---- one histogram pass over 256 value bins ----
    const float scale = 255.f / range;
    const float off = 8388608.f;
    if (tid < 66) reinterpret_cast<uint4 *>(s_hist)[tid] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    const uint32_t hbase = hist_addr - (0x4b000000u << 2);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const uint32_t yb = __float_as_uint(fmaf(x[e] - xmin, scale, off));
      const uint32_t addr = ((vmask >> e) & 1u) ? (yb << 2) + hbase : dummy;
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
    }
    __syncthreads();
    if (warp == 0) vb_scan_hist(s_hist, need, lane, ctl);
    __syncthreads();
    const int bin = (int)ctl->bin, cntb = (int)ctl->cntb, rem = (int)ctl->rem;
    if (rem >= 1 && cntb >= rem && cntb <= 32) {
      const uint32_t ytarget = 0x4b000000u + (uint32_t)bin;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        if (((vmask >> e) & 1u) && __float_as_uint(fmaf(x[e] - xmin, scale, off)) == ytarget) {
          const unsigned int pos = atomicAdd(&ctl->cand_n, 1u);
          if (pos < 64u) s_cand[pos] = __float_as_uint(x[e]);
        }
      }
      __syncthreads();
      if (warp == 0) {
        const float ci = (lane < cntb) ? __uint_as_float(s_cand[lane]) : 0.f;
        int gt = 0, eqb = 0, eqt = 0;
        for (int j = 0; j < cntb; ++j) {
          const float cj = __shfl_sync(0xffffffffu, ci, j);
          gt += (cj > ci) ? 1 : 0;
          eqt += (cj == ci) ? 1 : 0;
          eqb += (cj == ci && j < lane) ? 1 : 0;
        }
        const uint32_t sel = __ballot_sync(0xffffffffu, lane < cntb && gt + eqb == rem - 1);
        const int src = __ffs(sel) - 1;
        const float T = __shfl_sync(0xffffffffu, ci, src);
        const int gt_t = __shfl_sync(0xffffffffu, gt, src);
        const int eq_t = __shfl_sync(0xffffffffu, eqt, src);
        if (lane == 0) {
          ctl->tkey = f2key_fast(T);
          ctl->need = (unsigned int)(rem - gt_t);
          ctl->exact = (eq_t == rem - gt_t) ? 1u : 0u;
        }
      }
      __syncthreads();
      tkey = ctl->tkey;
      need = (int)ctl->need;
      exact = ctl->exact != 0u;
      done = true;
    }
  }
  if (!done) {
    // ---- multi-pass radix select on integer keys (256 buckets of the candidates' key range per pass) ----
    int shift = 32 - __clz((kmax - kmin) | 1u) - 8;
    if (shift < 0) shift = 0;
    uint32_t lo = kmin, span_m1 = 0xffffffffu;
    need = k;
#pragma unroll 1
    for (;;) {
      __syncthreads();
      if (tid < 66) reinterpret_cast<uint4 *>(s_hist)[tid] = make_uint4(0u, 0u, 0u, 0u);
      __syncthreads();
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const uint32_t dk = f2key_fast(x[e]) - lo;
        const bool in = ((vmask >> e) & 1u) && dk <= span_m1;
        const uint32_t addr = in ? hist_addr + ((dk >> shift) << 2) : dummy;
        asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
      }
      __syncthreads();
      if (warp == 0) vb_scan_hist(s_hist, need, lane, ctl);
      __syncthreads();
      const int bin = (int)ctl->bin, cntb = (int)ctl->cntb;
      need = (int)ctl->rem;
      lo += (uint32_t)bin << shift;
      if (cntb == need) {
        exact = true;
        break;
      }
      if (shift == 0) break;
      span_m1 = (1u << shift) - 1u;
      shift = (shift > 8) ? shift - 8 : 0;
    }
    tkey = lo;
  }
  // ---- winners: key >= T, or key > T plus the `need` lowest-index elements with key == T ----
  uint32_t wmask = 0u;
  if (exact) {
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (((vmask >> e) & 1u) && f2key_fast(x[e]) >= tkey) wmask |= 1u << e;
  } else {
    int need_eq = need;
#pragma unroll 1
    for (int i = 0; i < NI4; ++i) {
      uint32_t eqm = 0u;
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        float xv = 0.f;
#pragma unroll
        for (int ii = 0; ii < NI4; ++ii)
          if (ii == i) xv = x[4 * ii + cc];
        const bool ok = (vmask >> (4 * i + cc)) & 1u;
        const uint32_t kk = f2key_fast(xv);
        if (ok && kk > tkey) wmask |= 1u << (4 * i + cc);
        if (ok && kk == tkey) eqm |= 1u << cc;
      }
      int total;
      int before = vb_excl_scan(__popc(eqm), total, ctl->wsum, warp, lane);   // index order: thread, then component
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        if ((eqm >> cc) & 1u) {
          if (before < need_eq) wmask |= 1u << (4 * i + cc);
          ++before;
        }
      }
      need_eq -= min(need_eq, total);
    }
  }
  {
    int total;
    int pos = vb_excl_scan(__popc(wmask), total, ctl->wsum, warp, lane);
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if ((wmask >> e) & 1u) {
        const uint32_t id = (uint32_t)(((e >> 2) * VB_THREADS + tid) * 4 + (e & 3));
        if (pos < 128) s_buf[pos] = make_uint2(~id, __float_as_uint(x[e]));
        ++pos;
      }
    }
  }
  __syncthreads();
  if (warp == 0)
    emit_winners<SORTED>(s_buf, k, lane, tkey, kmax, vals + (size_t)r * k, idx + (size_t)r * k);
}

// marked = 0: block b selects row b.  marked = 1: the blocks walk all rows and redo those the streaming sift kernel
// marked with idx[r k] = -1 (rows it could not finish on its fast path).
template <int NI4, bool SORTED, class Rows>
__global__ void __launch_bounds__(VB_THREADS, 4)
topk_vecblock_kernel(Rows rows, int R, int W, int k, float *__restrict__ vals, int *__restrict__ idx, int marked) {
  if (!marked) {
    if ((int)blockIdx.x < R) vecblock_select_row<NI4, SORTED, Rows>(rows, blockIdx.x, W, k, vals, idx);
    return;
  }
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    if (idx[(size_t)r * k] == -1) {                          // (block-uniform)
      __syncthreads();                                       // the previous row's shared state is done with
      vecblock_select_row<NI4, SORTED, Rows>(rows, r, W, k, vals, idx);
    }
  }
}

// One 256-thread block per row, any width: three radix passes (11 + 11 + 10 bits) with shared-memory
// histograms; pass 1 streams the row from global memory, the candidates of the winning bucket are
// compacted to shared memory for the later passes (falls back to re-streaming if they do not fit).
constexpr int BLK_T = 256;
constexpr int CAND_MAX = 4096;
template <class Rows>
__global__ void __launch_bounds__(BLK_T)
topk_block_kernel(Rows rows, int R, int k, int KP, float *__restrict__ vals, int *__restrict__ idx) {
  extern __shared__ unsigned long long sbuf[];          // KP composites (output staging + sort)
  __shared__ unsigned int hist[2048];
  __shared__ uint32_t cand_key[CAND_MAX];
  __shared__ int cand_idx[CAND_MAX];
  __shared__ unsigned int s_cnt, s_out;
  __shared__ uint32_t s_prefix;
  __shared__ int s_need, s_eq_taken;
  const int r = blockIdx.x;
  const int tid = threadIdx.x;
  const int W = rows.width(r);

  // ---- pass 1: top 11 bits over the whole row
  for (int t = tid; t < 2048; t += BLK_T) hist[t] = 0u;
  if (tid == 0) { s_cnt = 0u; s_out = 0u; s_eq_taken = 0; }
  __syncthreads();
  for (int j = tid; j < W; j += BLK_T) {
    int run;
    const uint32_t kv = f2key(__ldg(rows.seg(r, j, run)));
    atomicAdd(&hist[kv >> 21], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    int need = k, b = 2047;
    for (; b > 0; --b) {
      if ((int)hist[b] >= need) break;
      need -= (int)hist[b];
    }
    s_prefix = (uint32_t)b << 21;
    s_need = need;                 // how many of bucket b are still wanted
  }
  __syncthreads();
  const uint32_t bucket1 = s_prefix >> 21;
  const int need1 = s_need;
  const bool fits = (int)hist[bucket1] <= CAND_MAX;
  // winners above the bucket go straight to the output staging; bucket members become candidates
  for (int j0 = 0; j0 < W; j0 += BLK_T) {
    const int j = j0 + tid;
    if (j < W) {
      int run;
      const uint32_t kv = f2key(__ldg(rows.seg(r, j, run)));
      const uint32_t d = kv >> 21;
      if (d > bucket1) {
        const unsigned int slot = atomicAdd(&s_out, 1u);
        sbuf[slot] = ((unsigned long long)kv << 32) | (uint32_t)(0xffffffffu - (uint32_t)j);
      } else if (d == bucket1 && fits) {
        const unsigned int slot = atomicAdd(&s_cnt, 1u);
        cand_key[slot] = kv;
        cand_idx[slot] = j;
      }
    }
  }
  __syncthreads();
  // ---- exact threshold inside the bucket: bitwise search over the low 21 bits
  uint32_t T = s_prefix;
  const int ncand = fits ? (int)s_cnt : 0;
  for (int bit = 20; bit >= 0; --bit) {
    const uint32_t cnd = T | (1u << bit);
    if (tid == 0) s_cnt = 0u;
    __syncthreads();
    int c = 0;
    if (fits) {
      for (int t = tid; t < ncand; t += BLK_T) c += (cand_key[t] >= cnd) ? 1 : 0;
    } else {
      for (int j = tid; j < W; j += BLK_T) {
        int run;
        const uint32_t kv = f2key(__ldg(rows.seg(r, j, run)));
        c += ((kv >> 21) == bucket1 && kv >= cnd) ? 1 : 0;
      }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, (unsigned)c);
    __syncthreads();
    if ((int)s_cnt >= need1) T = cnd;
    __syncthreads();
  }
  // ---- winners inside the bucket: key > T all, key == T lowest index first (ordered, single thread scan
  //      over the few equal keys keeps it exact and simple)
  if (tid == 0) s_cnt = 0u;
  __syncthreads();
  if (fits) {
    for (int t = tid; t < ncand; t += BLK_T) {
      if (cand_key[t] > T) {
        const unsigned int slot = atomicAdd(&s_out, 1u);
        sbuf[slot] = ((unsigned long long)cand_key[t] << 32) | (uint32_t)(0xffffffffu - (uint32_t)cand_idx[t]);
      }
    }
  } else {
    for (int j = tid; j < W; j += BLK_T) {
      int run;
      const uint32_t kv = f2key(__ldg(rows.seg(r, j, run)));
      if ((kv >> 21) == bucket1 && kv > T) {
        const unsigned int slot = atomicAdd(&s_out, 1u);
        sbuf[slot] = ((unsigned long long)kv << 32) | (uint32_t)(0xffffffffu - (uint32_t)j);
      }
    }
  }
  __syncthreads();
  // equal keys: ascending index order.  Candidates were appended in arbitrary order, so pick the smallest
  // remaining index repeatedly (need_eq is tiny unless the row is full of duplicates).
  {
    const int need_eq = k - (int)s_out;
    for (int it = 0; it < need_eq; ++it) {
      if (tid == 0) s_need = 0x7fffffff;
      __syncthreads();
      const int last = s_eq_taken;        // indices <= last - 1 already taken (last = next lower bound)
      int best = 0x7fffffff;
      if (fits) {
        for (int t = tid; t < ncand; t += BLK_T)
          if (cand_key[t] == T && cand_idx[t] >= last) best = min(best, cand_idx[t]);
      } else {
        for (int j = tid; j < W; j += BLK_T) {
          int run;
          const uint32_t kv = f2key(__ldg(rows.seg(r, j, run)));
          if (kv == T && j >= last) best = min(best, j);
        }
      }
      best = __reduce_min_sync(0xffffffffu, best);
      if ((tid & 31) == 0) atomicMin(&s_need, best);
      __syncthreads();
      if (tid == 0) {
        const int j = s_need;
        sbuf[s_out] = ((unsigned long long)T << 32) | (uint32_t)(0xffffffffu - (uint32_t)j);
        s_out = s_out + 1;
        s_eq_taken = j + 1;
      }
      __syncthreads();
    }
  }
  for (int t = k + tid; t < KP; t += BLK_T) sbuf[t] = 0ull;
  __syncthreads();
  bitonic_sort_desc<BLK_T>(sbuf, KP, tid);
  for (int t = tid; t < k; t += BLK_T) {
    const unsigned long long c = sbuf[t];
    vals[(size_t)r * k + t] = key2f((uint32_t)(c >> 32));
    idx[(size_t)r * k + t] = (int)(0xffffffffu - (uint32_t)(c & 0xffffffffu));
  }
}

static int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

template <int E, class Rows>
static void launch_radix(const Rows &rows, int R, int k, bool sorted, bool full, float *vals, int *idx,
                         cudaStream_t st) {
  const int grid = (R + 3) / 4;
  if (full) {
    if (sorted)
      topk_warp_radix_kernel<E, true, true, Rows><<<grid, 128, 0, st>>>(rows, R, k, vals, idx);
    else
      topk_warp_radix_kernel<E, true, false, Rows><<<grid, 128, 0, st>>>(rows, R, k, vals, idx);
  } else {
    if (sorted)
      topk_warp_radix_kernel<E, false, true, Rows><<<grid, 128, 0, st>>>(rows, R, k, vals, idx);
    else
      topk_warp_radix_kernel<E, false, false, Rows><<<grid, 128, 0, st>>>(rows, R, k, vals, idx);
  }
}

// Sampling plan of the sift select for rows of width W, k winners, a sample of m elements and a survivors' list of `cap`
// entries.  The count of row elements above the sample's q-quantile scatters around q W with
// sigma = W sqrt(q (1 - q) / m): aim the pivot at mu = k + z sigma survivors (z = 3.3: a row in a thousand falls short and
// takes the slow path) and require that mu + z sigma, plus the pivot's own histogram bin, still fits the list.
struct SiftPlan {
  bool ok;
  int jtarget;
};
static SiftPlan sift_plan(int W, int k, int m_samples, int cap) {
  const double m = (double)m_samples, z = 3.3;
  double mu = k, sigma = 0.0;
  for (int it = 0; it < 30; ++it) {
    const double q = mu / W;
    if (q >= 0.6) return SiftPlan{false, 0};
    sigma = W * sqrt(q * (1.0 - q) / m);
    mu = k + z * sigma;
  }
  const double q = mu / W;
  const double upper = mu + z * sigma + 0.125 * mu;          // (+ the pivot's own bin: 61 bins over ~5.6 sigma of a bell-shaped sample)
  SiftPlan p;
  p.ok = upper <= cap && q < 0.5;
  p.jtarget = (int)ceil(q * m);
  if (p.jtarget < 1) p.jtarget = 1;
  if (p.jtarget > (int)m) p.ok = false;
  return p;
}

// Wmax: widest row; uniform: every row has exactly Wmax elements
template <class Rows>
static int launch_topk(const Rows &rows, int R, int Wmax, int k, float *vals, int *idx, cudaStream_t st,
                       bool sorted = true, bool uniform = true) {
  const int KP = next_pow2(k < 32 ? 32 : k);
  EDRL_CHECK_ARG(KP <= 1024, "topk: k = %d is larger than the supported 1024", k);
  static const bool legacy = (getenv("EDRL_TOPK_LEGACY") != nullptr);    // A/B switch for profiling
  static const bool novec = (getenv("EDRL_TOPK_VEC") != nullptr && atoi(getenv("EDRL_TOPK_VEC")) == 0);
  static const bool nosift = (getenv("EDRL_TOPK_SIFT") != nullptr && atoi(getenv("EDRL_TOPK_SIFT")) == 0);
  // rows from this width on take the streaming sift kernel instead of the resident one (A/B runs; with the resident
  // kernel compiled for 6-8 blocks per SM it is the faster one up to its limit of 2048 columns)
  static const int stream_minw = getenv("EDRL_TOPK_STREAM_MINW") ? atoi(getenv("EDRL_TOPK_STREAM_MINW")) : 2049;
  if (uniform && Wmax >= 512 && Wmax <= 2048 && Wmax < stream_minw && k <= 128 && !legacy && !novec && !nosift && rows.vec4_ok()) {
    // sift select (topk_sift.cuh): sample pivot -> survivors -> exact select, when the sampling plan fits the list
    const int W4 = Wmax >> 2;
    bool done = false;
#define EDRL_SIFT_CASE(FI_, P_, SSTR_, SL_, BLK_)                                                                    \
  if (!done && W4 / 32 == FI_ && ((W4 % 32) != 0) == P_) {                                                           \
    const SiftPlan sp = sift_plan(Wmax, k, 32 * ((FI_ * 4 + SSTR_ - 1) / SSTR_), 32 * SL_);                         \
    if (sp.ok) {                                                                                                      \
      const int grid = (R + 3) / 4;                                                                                   \
      if (sorted)                                                                                                     \
        topk_sift_kernel<FI_, P_, SSTR_, SL_, true, Rows, BLK_><<<grid, 128, 0, st>>>(rows, R, Wmax, k, sp.jtarget, vals, idx); \
      else                                                                                                            \
        topk_sift_kernel<FI_, P_, SSTR_, SL_, false, Rows, BLK_><<<grid, 128, 0, st>>>(rows, R, Wmax, k, sp.jtarget, vals, idx); \
      done = true;                                                                                                    \
    }                                                                                                                 \
  }
    // BLK_: resident blocks per SM the kernel is compiled for (launch bounds -> registers per thread).  The kernel is bound
    // by the latency of its per-row scan / histogram-walk chains (38 % of the stall samples: short scoreboard), so it
    // wants warps, not registers -- measured, unsorted / sorted % of HBM: W = 800 with 8 | 10 | 12 blocks 65.7 / 43.2 |
    // 69.7 / 44.1 | 64.2 / 42.0 (40 registers spill); W = 1024 with 4 | 7 | 9 blocks 62.4 / 44.4 | 71.8 / 49.1 | 77.5 /
    // 51.0; W = 1600 with 4 | 6 | 8 blocks 64.5 / 50.3 | 77.1 / 58.1 | 83.5 / 59.2; W = 2048 with 4 | 6 blocks 71.1 / 56.2 |
    // 84.7 / 64.0; W = 512 with 8 | 12 | 16 blocks 51.7 | 53.9 | 48.7
    EDRL_SIFT_CASE(4, false, 1, 8, 12)      // W = 512
    EDRL_SIFT_CASE(6, true, 3, 8, 10)       // W = 800 (the reference's S): 8 sample values per lane
    EDRL_SIFT_CASE(8, false, 2, 8, 9)       // W = 1024
    EDRL_SIFT_CASE(12, true, 2, 8, 8)       // W = 1600 (C = 3 negatives)
    EDRL_SIFT_CASE(16, false, 2, 8, 6)      // W = 2048
#undef EDRL_SIFT_CASE
    if (done) {
      EDRL_LAUNCHED();
      return 0;
    }
  }
  if (uniform && Wmax <= 2048 && k <= 128 && !legacy && !novec && rows.vec4_ok()) {
    // vectorised warp select for the common uniform widths (W / 4 = 32 FI + partial lanes)
    const int W4 = Wmax >> 2, fi = W4 >> 5;
    const bool part = (W4 & 31) != 0;
    const int grid = (R + 3) / 4;
    bool done = true;
#define EDRL_VEC_CASE(FI_, P_)                                                                                   \
  if (fi == FI_ && part == P_) {                                                                                 \
    if (sorted) topk_vec_kernel<FI_, P_, true, Rows><<<grid, 128, 0, st>>>(rows, R, Wmax, k, vals, idx);        \
    else topk_vec_kernel<FI_, P_, false, Rows><<<grid, 128, 0, st>>>(rows, R, Wmax, k, vals, idx);              \
  } else
    EDRL_VEC_CASE(0, true) EDRL_VEC_CASE(1, true) EDRL_VEC_CASE(2, false) EDRL_VEC_CASE(4, false)
    EDRL_VEC_CASE(6, true) EDRL_VEC_CASE(8, false) EDRL_VEC_CASE(12, true) EDRL_VEC_CASE(16, false) { done = false; }
#undef EDRL_VEC_CASE
    if (done) {
      EDRL_LAUNCHED();
      return 0;
    }
  }
  if (uniform && ((Wmax > 2048 && Wmax <= 8192) || (Wmax >= 1024 && Wmax >= stream_minw && Wmax <= 2048 && !nosift)) && k <= 128 &&
      !legacy && !novec && rows.vec4_ok()) {
    // wide rows.  Streaming sift select, one warp per row, when the sampling plan fits (1024 sample elements, 512
    // list entries); the rows it marks -- or, without it, all rows -- go to one 256-thread block per row with the row
    // in registers (4 or 8 float4 per thread)
    int marked = 0;
    if (!nosift) {
      const SiftPlan sp = sift_plan(Wmax, k, 1024, 512);
      if (sp.ok) {
        const int grid = (R + 3) / 4;
        if (sorted) topk_sift_stream_kernel<true, Rows><<<grid, 128, 0, st>>>(rows, R, Wmax, k, sp.jtarget, vals, idx);
        else topk_sift_stream_kernel<false, Rows><<<grid, 128, 0, st>>>(rows, R, Wmax, k, sp.jtarget, vals, idx);
        EDRL_LAUNCHED();
        marked = 1;
      }
    }
    int sms = device_sm_count();
    if (sms <= 0) sms = 148;
    const int grid = marked ? (R < 4 * sms ? R : 4 * sms) : R;
    if (!marked && Wmax <= 2048) {
      // (no sampling plan for this (W, k): rows this narrow go on to the warp-per-row kernels below)
    } else if (Wmax <= 4096) {
      if (sorted) topk_vecblock_kernel<4, true, Rows><<<grid, VB_THREADS, 0, st>>>(rows, R, Wmax, k, vals, idx, marked);
      else topk_vecblock_kernel<4, false, Rows><<<grid, VB_THREADS, 0, st>>>(rows, R, Wmax, k, vals, idx, marked);
    } else {
      if (sorted) topk_vecblock_kernel<8, true, Rows><<<grid, VB_THREADS, 0, st>>>(rows, R, Wmax, k, vals, idx, marked);
      else topk_vecblock_kernel<8, false, Rows><<<grid, VB_THREADS, 0, st>>>(rows, R, Wmax, k, vals, idx, marked);
    }
    if (marked || Wmax > 2048) {
      EDRL_LAUNCHED();
      return 0;
    }
  }
  if (Wmax <= 2048 && k <= 128 && !legacy) {
    const bool full = uniform && (Wmax % 32 == 0);
    if (Wmax <= 256) launch_radix<8>(rows, R, k, sorted, full && Wmax == 256, vals, idx, st);
    else if (Wmax <= 512) launch_radix<16>(rows, R, k, sorted, full && Wmax == 512, vals, idx, st);
    else if (Wmax <= 800) launch_radix<25>(rows, R, k, sorted, full && Wmax == 800, vals, idx, st);
    else if (Wmax <= 1024) launch_radix<32>(rows, R, k, sorted, full && Wmax == 1024, vals, idx, st);
    else if (Wmax <= 1600) launch_radix<50>(rows, R, k, sorted, full && Wmax == 1600, vals, idx, st);
    else launch_radix<64>(rows, R, k, sorted, full && Wmax == 2048, vals, idx, st);
    EDRL_LAUNCHED();
    return 0;
  }
  if (Wmax <= 2048 && KP <= 256) {
    const size_t smem = (size_t)4 * KP * sizeof(unsigned long long);
    const int grid = (R + 3) / 4;
    if (Wmax <= 256)
      topk_warp_kernel<8, Rows><<<grid, 128, smem, st>>>(rows, R, k, KP, vals, idx);
    else if (Wmax <= 512)
      topk_warp_kernel<16, Rows><<<grid, 128, smem, st>>>(rows, R, k, KP, vals, idx);
    else if (Wmax <= 1024)
      topk_warp_kernel<32, Rows><<<grid, 128, smem, st>>>(rows, R, k, KP, vals, idx);
    else
      topk_warp_kernel<64, Rows><<<grid, 128, smem, st>>>(rows, R, k, KP, vals, idx);
  } else {
    const size_t smem = (size_t)KP * sizeof(unsigned long long);
    // ~40 KiB of static shared memory + up to 8 KiB of sort staging: opt in above the 48 KiB default
    // per launch (cheap): the attribute is per device, a process may drive several
  EDRL_CUDA_OK(cudaFuncSetAttribute(topk_block_kernel<Rows>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        1024 * (int)sizeof(unsigned long long)));
    topk_block_kernel<Rows><<<R, BLK_T, smem, st>>>(rows, R, k, KP, vals, idx);
  }
  EDRL_LAUNCHED();
  return 0;
}

// ----------------------------------------------------------------------------- K8 loss + scatter backward
__global__ void __launch_bounds__(256)
proxy_loss_fwd_kernel(const float *__restrict__ pos_val, const float *__restrict__ neg_val, int B, int k,
                      float *__restrict__ loss, float *__restrict__ rowexp) {
  __shared__ float s_part[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float part = 0.f;
  for (int b = warp; b < B; b += 8) {
    float sp = 0.f, sn = 0.f;
    for (int j = lane; j < k; j += 32) {
      sp += pos_val[(size_t)b * k + j];
      sn += neg_val[(size_t)b * k + j];
    }
    sp = warp_sum(sp);
    sn = warp_sum(sn);
    const float e = expf((sn - sp) / (float)k);
    if (lane == 0) rowexp[b] = e;
    part += e;
  }
  if (lane == 0) s_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_part[w];
    loss[0] = t / (float)B;
  }
}

// one block per batch row: zero datt[b, :, :] then scatter the two coefficient sets
__global__ void __launch_bounds__(256)
select_loss_bwd_kernel(const float *__restrict__ rowexp, const int *__restrict__ pos_idx,
                       const int *__restrict__ neg_idx, const long long *__restrict__ y,
                       const float *__restrict__ grad_out, int B, int C, int S, int k, float *__restrict__ datt) {
  const int b = blockIdx.x;
  float *row = datt + (size_t)b * C * S;
  for (int i = threadIdx.x; i < C * S; i += blockDim.x) row[i] = 0.f;
  __syncthreads();
  const int yb = min(max((int)y[b], 0), C - 1);
  const float coef = grad_out[0] * rowexp[b] / ((float)B * (float)k);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    row[(size_t)yb * S + pos_idx[(size_t)b * k + j]] = -coef;
    const int q = neg_idx[(size_t)b * k + j];
    const int cq = q / S, s = q - cq * S;
    const int cls = cq + (cq >= yb ? 1 : 0);
    row[(size_t)cls * S + s] = coef;
  }
}

// ----------------------------------------------------------------------------- K9 gather
// one warp per gathered row; 128-bit accesses when D % 4 == 0 and pointers are 16-byte aligned
__global__ void __launch_bounds__(256)
gather_rows_fwd_kernel(const float *__restrict__ feat, const int *__restrict__ idx, int B, int T, int D, int k,
                       int vec_ok, float *__restrict__ out) {
  const long long w = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (w >= (long long)B * k) return;
  const int b = (int)(w / k);
  const int t = idx[w];
  const float *src = feat + ((size_t)b * T + t) * D;
  float *dst = out + (size_t)w * D;
  if (vec_ok) {
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    const int n4 = D >> 2;
    int i = lane;
    for (; i + 96 < n4; i += 128) {           // 4 independent 16-byte loads in flight per lane
      const float4 v0 = __ldg(s4 + i), v1 = __ldg(s4 + i + 32), v2 = __ldg(s4 + i + 64), v3 = __ldg(s4 + i + 96);
      d4[i] = v0; d4[i + 32] = v1; d4[i + 64] = v2; d4[i + 96] = v3;
    }
    for (; i < n4; i += 32) d4[i] = __ldg(s4 + i);
  } else {
    for (int i = lane; i < D; i += 32) dst[i] = __ldg(src + i);
  }
}

// one block per batch element: inverse map in shared memory, every output row written exactly once
__global__ void __launch_bounds__(256)
gather_rows_bwd_kernel(const float *__restrict__ dout, const int *__restrict__ idx, int T, int D, int k, int vec_ok,
                       float *__restrict__ dfeat) {
  extern __shared__ int inv[];
  const int b = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += blockDim.x) inv[t] = -1;
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) inv[idx[(size_t)b * k + j]] = j;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int t = warp; t < T; t += 8) {
    const int j = inv[t];
    float *dst = dfeat + ((size_t)b * T + t) * D;
    const float *src = (j >= 0) ? dout + ((size_t)b * k + j) * D : nullptr;
    if (vec_ok) {
      float4 *d4 = reinterpret_cast<float4 *>(dst);
      const float4 *s4 = reinterpret_cast<const float4 *>(src);
      for (int i = lane; i < (D >> 2); i += 32) d4[i] = src ? __ldg(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int i = lane; i < D; i += 32) dst[i] = src ? __ldg(src + i) : 0.f;
    }
  }
}

}  // namespace eprl
}  // namespace edrl

using namespace edrl;
using namespace edrl::eprl;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" {

int edrl_token_stats_fwd(const float *z, int B, int T, int F, float *zbar, float *colsum, float *colnorm,
                         void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(z && zbar && colsum && colnorm, "token_stats_fwd: null argument");
  EDRL_CHECK_ARG(B > 0 && T > 0 && F > 0 && B <= 65535, "token_stats_fwd: bad shape B=%d T=%d F=%d", B, T, F);
  const bool v4 = (F % 4 == 0) && ((((uintptr_t)z | (uintptr_t)zbar | (uintptr_t)colsum | (uintptr_t)colnorm) & 15) == 0);
  if (v4) {
    const int F4 = F / 4;
    dim3 grid((F4 + 31) / 32, B), block(32, 8);
    token_stats_fwd_v4_kernel<<<grid, block, 0, ST(stream)>>>(reinterpret_cast<const float4 *>(z), T, F4,
                                                             reinterpret_cast<float4 *>(zbar),
                                                             reinterpret_cast<float4 *>(colsum),
                                                             reinterpret_cast<float4 *>(colnorm));
  } else {
    dim3 grid((F + 31) / 32, B), block(32, 8);
    token_stats_fwd_kernel<<<grid, block, 0, ST(stream)>>>(z, T, F, zbar, colsum, colnorm);
  }
  EDRL_LAUNCHED();
  return 0;
}

int edrl_token_stats_bwd(const float *z, const float *colsum, const float *colnorm, const float *dzbar, int B, int T,
                         int F, float *dz, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(z && colsum && colnorm && dzbar && dz, "token_stats_bwd: null argument");
  EDRL_CHECK_ARG(B > 0 && T > 0 && F > 0, "token_stats_bwd: bad shape");
  const bool v4 = (F % 4 == 0) && (256 % (F / 4) == 0) && B <= 65535 &&
                  ((((uintptr_t)z | (uintptr_t)dz | (uintptr_t)colsum | (uintptr_t)colnorm | (uintptr_t)dzbar) & 15) == 0);
  if (v4) {
    dim3 grid((T + TSB_TCH - 1) / TSB_TCH, B);
    token_stats_bwd_v4_kernel<<<grid, 256, 0, ST(stream)>>>(
        reinterpret_cast<const float4 *>(z), reinterpret_cast<const float4 *>(colsum),
        reinterpret_cast<const float4 *>(colnorm), reinterpret_cast<const float4 *>(dzbar), T, F / 4,
        reinterpret_cast<float4 *>(dz));
  } else {
    const size_t total = (size_t)B * T * F;
    token_stats_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ST(stream)>>>(z, colsum, colnorm, dzbar, T, F,
                                                                                    total, dz);
  }
  EDRL_LAUNCHED();
  return 0;
}

int edrl_token_featmean(const float *z, const float *colnorm, int B, int T, int F, float *zmean, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(z && colnorm && zmean, "token_featmean: null argument");
  EDRL_CHECK_ARG(B > 0 && T > 0 && F > 0, "token_featmean: bad shape");
  token_featmean_kernel<<<(B * T + 7) / 8, 256, 0, ST(stream)>>>(z, colnorm, B, T, F, zmean);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_proxy_normalize_fwd(const float *mu, const float *sigma, const float *eps, int C, int S, int F, float *z_pn,
                             float *pnorm, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(mu && sigma && eps && z_pn && pnorm, "proxy_normalize_fwd: null argument");
  EDRL_CHECK_ARG(C > 0 && S > 0 && F > 0, "proxy_normalize_fwd: bad shape");
  dim3 grid((F + 31) / 32, C), block(32, 32);
  proxy_normalize_fwd_kernel<<<grid, block, 0, ST(stream)>>>(mu, sigma, eps, S, F, F, 0, z_pn, pnorm);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_proxy_normalize_bwd(const float *mu, const float *sigma, const float *eps, const float *pnorm,
                             const float *dz_pn, int C, int S, int F, float *dmu, float *dsigma, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(mu && sigma && eps && pnorm && dz_pn && dmu && dsigma, "proxy_normalize_bwd: null argument");
  EDRL_CHECK_ARG(C > 0 && S > 0 && F > 0, "proxy_normalize_bwd: bad shape");
  dim3 grid((F + 31) / 32, C), block(32, 32);
  proxy_normalize_bwd_kernel<<<grid, block, 0, ST(stream)>>>(mu, sigma, eps, pnorm, dz_pn, S, F, F, 0, F, dmu, dsigma);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_score_fwd(const float *zbar, const float *z_pn, int B, int R, int F, float *att, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(zbar && z_pn && att, "score_fwd: null argument");
  EDRL_CHECK_ARG(B > 0 && R > 0 && F > 0, "score_fwd: bad shape");
  // att[b, r] = sum_f zbar[b, f] z_pn[r, f]
  if (B > 256) return sgemm(zbar, F, 1, z_pn, 1, F, att, B, R, F, ST(stream));     // wide batches: tiled GEMM
  score_fwd_kernel<<<(R + 7) / 8, 256, 0, ST(stream)>>>(zbar, z_pn, B, R, F, att);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_score_bwd(const float *datt, const float *zbar, const float *z_pn, int B, int R, int F, float *dzbar,
                   float *dz_pn, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(datt && zbar && z_pn, "score_bwd: null argument");
  EDRL_CHECK_ARG(B > 0 && R > 0 && F > 0, "score_bwd: bad shape");
  EDRL_CHECK_ARG(B <= 65535, "score_bwd: batch too large");
  if (dzbar) {   // dzbar[b, f] = sum_r datt[b, r] z_pn[r, f]
    dim3 grid((F + 31) / 32, B), block(32, 8);
    score_bwd_dzbar_kernel<<<grid, block, 0, ST(stream)>>>(datt, z_pn, R, F, dzbar);
    EDRL_LAUNCHED();
  }
  if (dz_pn) {   // dz_pn[r, f] = sum_b datt[b, r] zbar[b, f]
    if (B > 256) {
      if (int rc = sgemm(datt, 1, R, zbar, F, 1, dz_pn, R, F, B, ST(stream))) return rc;
    } else {
      score_bwd_dzpn_kernel<<<(R + 7) / 8, 256, 0, ST(stream)>>>(datt, zbar, B, R, F, dz_pn);
      EDRL_LAUNCHED();
    }
  }
  return 0;
}

int edrl_topk_rows(const float *x, int R, int W, int ld, int k, int sorted, float *vals, int32_t *idx,
                   void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(x && vals && idx, "topk: null argument");
  EDRL_CHECK_ARG(R > 0 && W > 0 && ld >= W, "topk: bad shape R=%d W=%d ld=%d", R, W, ld);
  EDRL_CHECK_ARG(k >= 1 && k <= W, "selected index k out of range (k=%d, row width %d)", k, W);
  PlainRows rows{x, W, ld};
  return launch_topk(rows, R, W, k, vals, idx, ST(stream), sorted != 0);
}

int edrl_select_topk_fwd(const float *att, const int64_t *y, int B, int C, int S, int k, int sorted, float *pos_val,
                         int32_t *pos_idx, float *neg_val, int32_t *neg_idx, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(att && y && pos_val && pos_idx && neg_val && neg_idx, "select_topk_fwd: null argument");
  EDRL_CHECK_ARG(B > 0 && C >= 2 && S > 0, "select_topk_fwd: bad shape B=%d C=%d S=%d", B, C, S);
  EDRL_CHECK_ARG(k >= 1 && k <= S, "selected index k out of range (k=%d, row width %d)", k, S);
  EDRL_CHECK_ARG(pos_val + (size_t)B * k == neg_val && pos_idx + (size_t)B * k == neg_idx,
                 "select_topk_fwd: pos/neg outputs must be the two halves of one [2B, k] buffer");
  EssenceRows rows{att, reinterpret_cast<const long long *>(y), B, C, S};
  return launch_topk(rows, 2 * B, (C - 1) * S > S ? (C - 1) * S : S, k, pos_val, pos_idx, ST(stream), sorted != 0,
                     /*uniform=*/C == 2);
}

int edrl_proxy_loss_fwd(const float *pos_val, const float *neg_val, int B, int k, float *loss, float *rowexp,
                        void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(pos_val && neg_val && loss && rowexp, "proxy_loss_fwd: null argument");
  EDRL_CHECK_ARG(B > 0 && k > 0, "proxy_loss_fwd: bad shape");
  proxy_loss_fwd_kernel<<<1, 256, 0, ST(stream)>>>(pos_val, neg_val, B, k, loss, rowexp);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_select_loss_bwd(const float *rowexp, const int32_t *pos_idx, const int32_t *neg_idx, const int64_t *y,
                         const float *grad_out, int B, int C, int S, int k, float *datt, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(rowexp && pos_idx && neg_idx && y && grad_out && datt, "select_loss_bwd: null argument");
  EDRL_CHECK_ARG(B > 0 && C >= 2 && S > 0 && k > 0, "select_loss_bwd: bad shape");
  select_loss_bwd_kernel<<<B, 256, 0, ST(stream)>>>(rowexp, pos_idx, neg_idx,
                                                    reinterpret_cast<const long long *>(y), grad_out, B, C, S, k, datt);
  EDRL_LAUNCHED();
  return 0;
}

static int vec4_ok(const void *a, const void *b, int D) {
  return (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

int edrl_gather_rows_fwd(const float *features, const int32_t *idx, int B, int T, int D, int k, float *out,
                         void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(features && idx && out, "gather_rows_fwd: null argument");
  EDRL_CHECK_ARG(B > 0 && T > 0 && D > 0 && k > 0, "gather_rows_fwd: bad shape");
  const long long rows = (long long)B * k;
  gather_rows_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, ST(stream)>>>(features, idx, B, T, D, k,
                                                                             vec4_ok(features, out, D), out);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_gather_rows_bwd(const float *dout, const int32_t *idx, int B, int T, int D, int k, float *dfeatures,
                         void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(dout && idx && dfeatures, "gather_rows_bwd: null argument");
  EDRL_CHECK_ARG(B > 0 && T > 0 && D > 0 && k > 0, "gather_rows_bwd: bad shape");
  EDRL_CHECK_ARG((size_t)T * sizeof(int) <= 48 * 1024, "gather_rows_bwd: T = %d too large", T);
  gather_rows_bwd_kernel<<<B, 256, (size_t)T * sizeof(int), ST(stream)>>>(dout, idx, T, D, k,
                                                                          vec4_ok(dout, dfeatures, D), dfeatures);
  EDRL_LAUNCHED();
  return 0;
}

/* ---- fused train path: one call forward, one call backward (the separate entry points above stay for the eval
 *      branch and the tests).  `saved` is carved by ess_layout(); every segment starts 16-byte aligned. ---- */
namespace {
struct EssLayout {
  size_t zbar, colsum, colnorm, z_pn, pnorm, att, vals, idx, rowexp, total;      // saved (floats)
  size_t datt, dzbar, dz_pn, scratch_total;                                     // scratch (floats)
};
inline size_t up4(size_t v) { return (v + 3) / 4 * 4; }
EssLayout ess_layout(int B, int T, int F, int C, int S, int k) {
  (void)T;
  EssLayout L;
  size_t o = 0;
  L.zbar = o;    o += up4((size_t)B * F);
  L.colsum = o;  o += up4((size_t)B * F);
  L.colnorm = o; o += up4((size_t)B * F);
  L.z_pn = o;    o += up4((size_t)C * S * F);
  L.pnorm = o;   o += up4((size_t)C * F);
  L.att = o;     o += up4((size_t)B * C * S);
  L.vals = o;    o += up4((size_t)2 * B * k);
  L.idx = o;     o += up4((size_t)2 * B * k);
  L.rowexp = o;  o += up4((size_t)B);
  L.total = o;
  o = 0;
  L.datt = o;    o += up4((size_t)B * C * S);
  L.dzbar = o;   o += up4((size_t)B * F);
  L.dz_pn = o;   o += up4((size_t)C * S * F);
  L.scratch_total = o;
  return L;
}
}  // namespace

size_t edrl_essence_saved_floats(int B, int T, int F, int C, int S, int k) {
  if (B <= 0 || T <= 0 || F <= 0 || C < 2 || S <= 0 || k <= 0) return 0;
  return ess_layout(B, T, F, C, S, k).total;
}
size_t edrl_essence_scratch_floats(int B, int T, int F, int C, int S, int k) {
  if (B <= 0 || T <= 0 || F <= 0 || C < 2 || S <= 0 || k <= 0) return 0;
  return ess_layout(B, T, F, C, S, k).scratch_total;
}

int edrl_essence_train_fwd(const float *z, const float *proxies, const float *eps, const int64_t *y, int B, int T,
                           int F, int C, int S, int k, float *loss, float *saved, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(z && proxies && eps && y && loss && saved, "essence_train_fwd: null argument");
  EDRL_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && F > 0 && C >= 2 && S > 0, "essence_train_fwd: bad shape");
  EDRL_CHECK_ARG(k >= 1 && k <= S, "selected index k out of range (k=%d, row width %d)", k, S);
  const EssLayout L = ess_layout(B, T, F, C, S, k);
  float *vals = saved + L.vals;
  int32_t *idx = reinterpret_cast<int32_t *>(saved + L.idx);
  if (int rc = edrl_token_stats_fwd(z, B, T, F, saved + L.zbar, saved + L.colsum, saved + L.colnorm, stream)) return rc;
  {
    dim3 grid((F + 31) / 32, C), block(32, 32);
    proxy_normalize_fwd_kernel<<<grid, block, 0, ST(stream)>>>(proxies, proxies + F, eps, S, F, 2 * F, 1,
                                                               saved + L.z_pn, saved + L.pnorm);
    EDRL_LAUNCHED();
  }
  if (int rc = edrl_score_fwd(saved + L.zbar, saved + L.z_pn, B, C * S, F, saved + L.att, stream)) return rc;
  if (int rc = edrl_select_topk_fwd(saved + L.att, y, B, C, S, k, /*sorted=*/0, vals, idx, vals + (size_t)B * k,
                                    idx + (size_t)B * k, stream))
    return rc;
  return edrl_proxy_loss_fwd(vals, vals + (size_t)B * k, B, k, loss, saved + L.rowexp, stream);
}

int edrl_essence_train_bwd(const float *z, const float *proxies, const float *eps, const int64_t *y, int B, int T,
                           int F, int C, int S, int k, const float *saved, const float *grad_out, float *scratch,
                           float *dz, float *dproxies, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(z && proxies && eps && y && saved && grad_out && scratch, "essence_train_bwd: null argument");
  EDRL_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && F > 0 && C >= 2 && S > 0 && k >= 1 && k <= S,
                 "essence_train_bwd: bad shape");
  const EssLayout L = ess_layout(B, T, F, C, S, k);
  const int32_t *idx = reinterpret_cast<const int32_t *>(saved + L.idx);
  if (int rc = edrl_select_loss_bwd(saved + L.rowexp, idx, idx + (size_t)B * k, y, grad_out, B, C, S, k,
                                    scratch + L.datt, stream))
    return rc;
  if (int rc = edrl_score_bwd(scratch + L.datt, saved + L.zbar, saved + L.z_pn, B, C * S, F, dz ? scratch + L.dzbar : nullptr,
                              dproxies ? scratch + L.dz_pn : nullptr, stream))
    return rc;
  if (dz) {
    if (int rc = edrl_token_stats_bwd(z, saved + L.colsum, saved + L.colnorm, scratch + L.dzbar, B, T, F, dz, stream))
      return rc;
  }
  if (dproxies) {
    dim3 grid((F + 31) / 32, C), block(32, 32);
    proxy_normalize_bwd_kernel<<<grid, block, 0, ST(stream)>>>(proxies, proxies + F, eps, saved + L.pnorm,
                                                               scratch + L.dz_pn, S, F, 2 * F, 1, 2 * F, dproxies,
                                                               dproxies + F);
    EDRL_LAUNCHED();
  }
  return 0;
}

}  // extern "C"
