// DILR Barlow-Twins cross-correlation loss (reference: code/fusion_net.py:656-677 + the two
// BatchNorm1d(2048, affine=False) of :653-654) -- the producer of Part A's input, same family as Part A: a contraction
// whose D x D result is reduced to a few scalars in the epilogue and never reaches HBM.
//
//   zh1 = bn1(z1), zh2 = bn2(z2)                         [B, D] -> normalised, kept TRANSPOSED [D, Bp] for the contraction
//   c   = zh1^T zh2 / (4 batch_size)                     [D, D]; only the two diagonal blocks matter:
//   common block c[:dc, :dc]:  on = sum_i (c_ii - 1)^2,  off = sum_{i != j} c_ij^2
//   unique block c[dc:, dc:]:  on = sum_i c_ii^2,        off = sum_{i != j} c_ij^2,        loss = on + 0.0051 off
//
// The contraction length is the BATCH (32-64 rows in the reference's training): 2 x 1024^2 x 64 MACs, a few microseconds
// of fp32 FFMA -- the reference spends ~30 launches and a 16 MB matrix on it.  Four kernels here: batch statistics +
// normalise + transpose; 64 x 64 correlation tiles with the masked-square epilogue (fp64 block sums, last-CTA finalise);
// backward: tiles recomputed, W = dL/dc formed in shared memory, dzh = W zh (the S -> W -> P structure of the MMD sweep);
// BatchNorm backward.  fp32 CUDA cores: at K = batch the tensor pipe would idle on operand staging, and fp32 products
// keep the result at the reference's own precision.
#include <stdint.h>

#include "../../include/edrl_b200.h"
#include "common.cuh"

namespace edrl {
namespace dilr {

constexpr int TILE = 64;      // correlation tile edge
constexpr int KC = 16;        // batch rows per shared-memory step
constexpr float OFF_W = 0.0051f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- K1: per-feature batch statistics, normalise, transpose.  grid (ceil(D / 32), 2), 256 threads: lane = feature, the
// 8 warps split the batch rows.  stats[which][0] = mean, [1] = 1 / sqrt(var + eps)  (D each).
__global__ void __launch_bounds__(256)
bn_fwd_kernel(const float *__restrict__ z1, const float *__restrict__ z2, int B, int D, int Bp, float eps, int training,
              float momentum, float *__restrict__ rm1, float *__restrict__ rv1, float *__restrict__ rm2,
              float *__restrict__ rv2, float *__restrict__ zt1, float *__restrict__ zt2, float *__restrict__ stats,
              double *__restrict__ acc, unsigned *__restrict__ ticket) {
  __shared__ float s_sum[8][32], s_sq[8][32];
  const int which = blockIdx.y;
  const float *z = which ? z2 : z1;
  float *zt = which ? zt2 : zt1;
  float *rm = which ? rm2 : rm1, *rv = which ? rv2 : rv1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane;
  if (blockIdx.x == 0 && which == 0 && threadIdx.x < 4) acc[threadIdx.x] = 0.0;     // the loss kernel's block sums
  if (blockIdx.x == 0 && which == 0 && threadIdx.x == 4) *ticket = 0u;
  float mean = 0.f, inv = 0.f;
  if (training) {
    float s = 0.f;
    if (j < D)
      for (int b = warp; b < B; b += 8) s += z[(size_t)b * D + j];
    s_sum[warp][lane] = s;
    __syncthreads();
    s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += s_sum[w][lane];
    mean = s / (float)B;
    float q = 0.f;                                           // second pass: centred squares (no cancellation)
    if (j < D)
      for (int b = warp; b < B; b += 8) {
        const float dv = z[(size_t)b * D + j] - mean;
        q = fmaf(dv, dv, q);
      }
    s_sq[warp][lane] = q;
    __syncthreads();
    q = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) q += s_sq[w][lane];
    const float var = q / (float)B;                          // biased: what normalises in train mode
    inv = 1.0f / sqrtf(var + eps);
    if (warp == 0 && j < D && rm != nullptr) {               // running statistics like nn.BatchNorm1d (unbiased variance)
      const float unb = (B > 1) ? q / (float)(B - 1) : var;
      rm[j] = (1.f - momentum) * rm[j] + momentum * mean;
      rv[j] = (1.f - momentum) * rv[j] + momentum * unb;
    }
  } else if (j < D) {
    mean = rm[j];
    inv = 1.0f / sqrtf(rv[j] + eps);
  }
  if (j < D) {
    if (warp == 0) {
      stats[(size_t)which * 2 * D + j] = mean;
      stats[(size_t)which * 2 * D + D + j] = inv;
    }
    // (the transposed store is strided by Bp: B is a few dozen rows, 2048 x B floats in all)
    for (int b = warp; b < Bp; b += 8) zt[(size_t)j * Bp + b] = (b < B) ? (z[(size_t)b * D + j] - mean) * inv : 0.f;
  }
}

// One 64 x 64 tile of c = A B^T / scale over the batch: A = zt_a[i0.., :], B = zt_b[j0.., :] ([feature][Bp] layouts, zero
// padded to Bp % KC == 0).  256 threads as 16 x 16, a 4 x 4 micro-tile each: acc[u][v] = c[i0 + 4 ty + u][j0 + 4 tx + v].
__device__ __forceinline__ void corr_tile(const float *__restrict__ za, const float *__restrict__ zb, int i0, int j0,
                                          int na, int nb, int Bp, float (*As)[TILE + 4], float (*Bs)[TILE + 4],
                                          float (&acc)[4][4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int lf = tid >> 2, lk = (tid & 3) * 4;               // loader: feature lf of the tile, batch rows lk .. lk + 3
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
  for (int k0 = 0; k0 < Bp; k0 += KC) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (i0 + lf < na) a = *reinterpret_cast<const float4 *>(za + (size_t)(i0 + lf) * Bp + k0 + lk);
    if (j0 + lf < nb) b = *reinterpret_cast<const float4 *>(zb + (size_t)(j0 + lf) * Bp + k0 + lk);
    __syncthreads();
    As[lk + 0][lf] = a.x; As[lk + 1][lf] = a.y; As[lk + 2][lf] = a.z; As[lk + 3][lf] = a.w;
    Bs[lk + 0][lf] = b.x; Bs[lk + 1][lf] = b.y; Bs[lk + 2][lf] = b.z; Bs[lk + 3][lf] = b.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 av = *reinterpret_cast<const float4 *>(&As[kk][4 * ty]);
      const float4 bv = *reinterpret_cast<const float4 *>(&Bs[kk][4 * tx]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a4[u], b4[v], acc[u][v]);
    }
  }
}

// ---- K2: loss.  grid (tiles_j, tiles_i, 2 blocks): block z = 0 common [0, dc), 1 unique [dc, D).
__global__ void __launch_bounds__(256)
corr_loss_kernel(const float *__restrict__ zt1, const float *__restrict__ zt2, int D, int dc, int Bp, float scale,
                 double *__restrict__ acc, unsigned *__restrict__ ticket, float *__restrict__ out6) {
  __shared__ __align__(16) float As[KC][TILE + 4], Bs[KC][TILE + 4];
  __shared__ double s_red[8][2];
  const int blk = blockIdx.z;
  const int f0 = blk ? dc : 0, nf = blk ? D - dc : dc;
  const int i0 = blockIdx.y * TILE, j0 = blockIdx.x * TILE;
  double on = 0.0, off = 0.0;
  if (i0 < nf && j0 < nf) {                                  // (block-uniform)
    float c[4][4];
    corr_tile(zt1 + (size_t)f0 * Bp, zt2 + (size_t)f0 * Bp, i0, j0, nf, nf, Bp, As, Bs, c);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const float target = blk ? 0.f : 1.f;
    float fon = 0.f, foff = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = i0 + 4 * ty + u, j = j0 + 4 * tx + v;
        if (i < nf && j < nf) {
          const float cv = c[u][v] * scale;
          if (i == j) fon = fmaf(cv - target, cv - target, fon);
          else foff = fmaf(cv, cv, foff);
        }
      }
    on = fon;
    off = foff;
  }
  on = warp_sum(on);
  off = warp_sum(off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_red[warp][0] = on;
    s_red[warp][1] = off;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += s_red[w][0];
      b += s_red[w][1];
    }
    atomicAdd(&acc[2 * blk + 0], a);
    atomicAdd(&acc[2 * blk + 1], b);
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(ticket, 1u) == total - 1) {                // the last block writes the six outputs
      __threadfence();
      const volatile double *va = acc;                       // (written by atomics in L2: not through this SM's L1)
      const double on_c = va[0], off_c = va[1], on_u = va[2], off_u = va[3];
      out6[0] = (float)(on_c + (double)OFF_W * off_c);
      out6[1] = (float)on_c;
      out6[2] = (float)off_c;
      out6[3] = (float)(on_u + (double)OFF_W * off_u);
      out6[4] = (float)on_u;
      out6[5] = (float)off_u;
    }
  }
}

// ---- K3: backward through c.  grid (panels of 64 features over both blocks, 2 sides, JS column splits x batch chunks):
// side 0 computes dzh1^T for a panel of z1 features (W zh2 over the block's column tiles), side 1 dzh2^T for a panel of z2
// features (W^T zh1).  A block takes every JS-th column tile and one chunk of BCH batch rows: per tile the correlation tile
// is recomputed, W = dL/dc goes to shared memory, then dzh^T[i, b] += sum_j W[i][j] zh_other^T[j, b] with the chunk's BCH / 4
// batch columns per thread in registers.  Split js writes its own partial output (summed, in a fixed order, by K4).
constexpr int BCH = 64;
constexpr int JS = 4;
__global__ void __launch_bounds__(256)
corr_bwd_kernel(const float *__restrict__ zt1, const float *__restrict__ zt2, int D, int dc, int Bp, float scale,
                const float *__restrict__ gout6, float *__restrict__ dzt1, float *__restrict__ dzt2, size_t part_stride) {
  __shared__ __align__(16) float As[KC][TILE + 4], Bs[KC][TILE + 4];
  __shared__ float Ws[TILE][TILE + 1];
  __shared__ __align__(16) float Zs[TILE][BCH + 4];
  const int side = blockIdx.y;
  const int js = blockIdx.z % JS, b0 = (blockIdx.z / JS) * BCH;
  const int pc = (dc + TILE - 1) / TILE;                     // panels of the common block come first
  const int blk = ((int)blockIdx.x >= pc) ? 1 : 0;
  const int f0 = blk ? dc : 0, nf = blk ? D - dc : dc;
  const int i0 = (blk ? (int)blockIdx.x - pc : (int)blockIdx.x) * TILE;
  if (i0 >= nf) return;
  const float *za = (side ? zt2 : zt1) + (size_t)f0 * Bp;    // this side's features (rows of the panel)
  const float *zb = (side ? zt1 : zt2) + (size_t)f0 * Bp;    // the other side's features (columns)
  float *dz = (side ? dzt2 : dzt1) + (size_t)js * part_stride + (size_t)f0 * Bp;
  const float g_loss = gout6[blk * 3 + 0], g_on = gout6[blk * 3 + 1], g_off = gout6[blk * 3 + 2];
  const float w_on = 2.f * (g_loss + g_on) * scale, w_off = 2.f * (OFF_W * g_loss + g_off) * scale;
  const float target = blk ? 0.f : 1.f;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int pi = tid >> 2, pb = (tid & 3) * (BCH / 4);       // P phase: row pi of the panel, batch columns pb .. pb + 15
  float accp[BCH / 4];
#pragma unroll
  for (int q = 0; q < BCH / 4; ++q) accp[q] = 0.f;
  for (int j0 = js * TILE; j0 < nf; j0 += JS * TILE) {
    float c[4][4];
    corr_tile(za, zb, i0, j0, nf, nf, Bp, As, Bs, c);        // c[u][v] = <za[i], zb[j]> (c is symmetric in the roles)
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int i = i0 + 4 * ty + u, j = j0 + 4 * tx + v;
        const float cv = c[u][v] * scale;
        float w = 0.f;
        if (i < nf && j < nf) w = (i == j) ? w_on * (cv - target) : w_off * cv;
        Ws[4 * ty + u][4 * tx + v] = w;
      }
    // the other side's normalised rows for this column tile and batch chunk
    for (int t = tid; t < TILE * (BCH / 4); t += 256) {
      const int jj = t / (BCH / 4), b4 = (t % (BCH / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + jj < nf && b0 + b4 < Bp) v = *reinterpret_cast<const float4 *>(zb + (size_t)(j0 + jj) * Bp + b0 + b4);
      *reinterpret_cast<float4 *>(&Zs[jj][b4]) = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < TILE; ++jj) {
      const float w = Ws[pi][jj];
#pragma unroll
      for (int q = 0; q < BCH / 4; q += 4) {
        const float4 zv = *reinterpret_cast<const float4 *>(&Zs[jj][pb + q]);
        accp[q + 0] = fmaf(w, zv.x, accp[q + 0]);
        accp[q + 1] = fmaf(w, zv.y, accp[q + 1]);
        accp[q + 2] = fmaf(w, zv.z, accp[q + 2]);
        accp[q + 3] = fmaf(w, zv.w, accp[q + 3]);
      }
    }
    __syncthreads();
  }
  if (i0 + pi < nf) {
#pragma unroll
    for (int q = 0; q < BCH / 4; q += 4) {
      if (b0 + pb + q < Bp)
        *reinterpret_cast<float4 *>(dz + (size_t)(i0 + pi) * Bp + b0 + pb + q) =
            make_float4(accp[q], accp[q + 1], accp[q + 2], accp[q + 3]);
    }
  }
}

// d zh^T element: the JS column-split partials of K3, added in a fixed order
__device__ __forceinline__ float dzh(const float *__restrict__ dzt, size_t part_stride, size_t o) {
  float g = dzt[o];
#pragma unroll
  for (int p = 1; p < JS; ++p) g += dzt[(size_t)p * part_stride + o];
  return g;
}

// ---- K4: BatchNorm backward (train: dz = inv (dzh - mean_b dzh - zh mean_b(dzh zh)); eval: dz = inv dzh), back to [B, D].
// grid (ceil(D / 32), 2), 256 threads: lane = feature, the 8 warps split the batch rows.
__global__ void __launch_bounds__(256)
bn_bwd_kernel(const float *__restrict__ zt1, const float *__restrict__ zt2, const float *__restrict__ dzt1,
              const float *__restrict__ dzt2, size_t part_stride, const float *__restrict__ stats, int B, int D, int Bp,
              int training, float *__restrict__ dz1, float *__restrict__ dz2) {
  __shared__ float s_a[8][32], s_b[8][32];
  const int which = blockIdx.y;
  const float *zt = which ? zt2 : zt1, *dzt = which ? dzt2 : dzt1;
  float *dz = which ? dz2 : dz1;
  if (dz == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane;
  float m1 = 0.f, m2 = 0.f;
  if (training) {
    float a = 0.f, b = 0.f;
    if (j < D)
      for (int r = warp; r < B; r += 8) {
        const float g = dzh(dzt, part_stride, (size_t)j * Bp + r);
        a += g;
        b = fmaf(g, zt[(size_t)j * Bp + r], b);
      }
    s_a[warp][lane] = a;
    s_b[warp][lane] = b;
    __syncthreads();
    a = 0.f;
    b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      a += s_a[w][lane];
      b += s_b[w][lane];
    }
    m1 = a / (float)B;
    m2 = b / (float)B;
  }
  if (j < D) {
    const float inv = stats[(size_t)which * 2 * D + D + j];
    for (int r = warp; r < B; r += 8)
      dz[(size_t)r * D + j] = inv * (dzh(dzt, part_stride, (size_t)j * Bp + r) - m1 - zt[(size_t)j * Bp + r] * m2);
  }
}

}  // namespace dilr
}  // namespace edrl

using namespace edrl;
using namespace edrl::dilr;

extern "C" {

size_t edrl_dilr_workspace_bytes(int B, int D) {
  if (B <= 0 || D <= 0) return 0;
  const size_t Bp = align_up((size_t)B, 64);
  // [acc f64[4] | ticket | pad -> 256 B | stats f32[4 D] | zt1, zt2 f32[D, Bp] | dzt1, dzt2 f32[JS][D, Bp]]
  return 256 + align_up((size_t)4 * D * 4, 256) + (size_t)(2 + 2 * dilr::JS) * align_up((size_t)D * Bp * 4, 256);
}

struct DilrWs {
  double *acc;
  unsigned *ticket;
  float *stats, *zt1, *zt2, *dzt1, *dzt2;
  size_t part_stride;        // floats between the JS partial copies of a dzt array
  int Bp;
};
static DilrWs dilr_ws(void *ws, int B, int D) {
  DilrWs w;
  uint8_t *p = reinterpret_cast<uint8_t *>(ws);
  w.Bp = (int)align_up((size_t)B, 64);
  w.acc = reinterpret_cast<double *>(p);
  w.ticket = reinterpret_cast<unsigned *>(p + 64);
  p += 256;
  w.stats = reinterpret_cast<float *>(p);
  p += align_up((size_t)4 * D * 4, 256);
  const size_t m = align_up((size_t)D * w.Bp * 4, 256);
  w.zt1 = reinterpret_cast<float *>(p);
  w.zt2 = reinterpret_cast<float *>(p + m);
  w.dzt1 = reinterpret_cast<float *>(p + 2 * m);
  w.dzt2 = reinterpret_cast<float *>(p + (2 + JS) * m);
  w.part_stride = m / 4;
  return w;
}

int edrl_dilr_bt_loss_fwd(const float *z1, const float *z2, int B, int D, int common_dim, int batch_size, float eps,
                          int training, float momentum, float *run_mean1, float *run_var1, float *run_mean2,
                          float *run_var2, float *out6, void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(z1 && z2 && out6 && workspace, "bt_loss_cross: null argument");
  EDRL_CHECK_ARG(B > 0 && D > 0 && common_dim >= 0 && common_dim <= D && batch_size > 0,
                 "bt_loss_cross: bad shape B=%d D=%d common_dim=%d batch_size=%d", B, D, common_dim, batch_size);
  EDRL_CHECK_ARG(training || (run_mean1 && run_var1 && run_mean2 && run_var2),
                 "bt_loss_cross: eval mode needs the running statistics");
  EDRL_CHECK_ARG(workspace_bytes >= edrl_dilr_workspace_bytes(B, D), "bt_loss_cross: workspace too small");
  EDRL_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "bt_loss_cross: workspace must be 256-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DilrWs w = dilr_ws(workspace, B, D);
  bn_fwd_kernel<<<dim3((D + 31) / 32, 2), 256, 0, st>>>(z1, z2, B, D, w.Bp, eps, training, momentum, run_mean1, run_var1,
                                                        run_mean2, run_var2, w.zt1, w.zt2, w.stats, w.acc, w.ticket);
  EDRL_LAUNCHED();
  const int nmax = common_dim > D - common_dim ? common_dim : D - common_dim;
  const int nt = (nmax + TILE - 1) / TILE;
  corr_loss_kernel<<<dim3(nt, nt, 2), 256, 0, st>>>(w.zt1, w.zt2, D, common_dim, w.Bp, 1.0f / (4.0f * (float)batch_size),
                                                    w.acc, w.ticket, out6);
  EDRL_LAUNCHED();
  return 0;
}

int edrl_dilr_bt_loss_bwd(int B, int D, int common_dim, int batch_size, int training, const float *grad_out6, float *dz1,
                          float *dz2, void *workspace, size_t workspace_bytes, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(grad_out6 && workspace && (dz1 || dz2), "bt_loss_cross backward: null argument");
  EDRL_CHECK_ARG(B > 0 && D > 0 && common_dim >= 0 && common_dim <= D && batch_size > 0, "bt_loss_cross backward: bad shape");
  EDRL_CHECK_ARG(workspace_bytes >= edrl_dilr_workspace_bytes(B, D), "bt_loss_cross backward: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const DilrWs w = dilr_ws(workspace, B, D);
  const int panels = (common_dim + TILE - 1) / TILE + (D - common_dim + TILE - 1) / TILE;
  corr_bwd_kernel<<<dim3(panels, 2, JS * (w.Bp / BCH)), 256, 0, st>>>(w.zt1, w.zt2, D, common_dim, w.Bp,
                                                                      1.0f / (4.0f * (float)batch_size), grad_out6, w.dzt1,
                                                                      w.dzt2, w.part_stride);
  EDRL_LAUNCHED();
  bn_bwd_kernel<<<dim3((D + 31) / 32, 2), 256, 0, st>>>(w.zt1, w.zt2, w.dzt1, w.dzt2, w.part_stride, w.stats, B, D, w.Bp,
                                                        training, dz1, dz2);
  EDRL_LAUNCHED();
  return 0;
}

}  // extern "C"
