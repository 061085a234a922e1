// mmd_prep.cuh -- Part A, K1: column statistics, centring, TF32 / binary16 operand copies (see mmd.cu)
// (textually included by mmd.cu inside namespace edrl::mmd; not a stand-alone header)
#pragma once

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

