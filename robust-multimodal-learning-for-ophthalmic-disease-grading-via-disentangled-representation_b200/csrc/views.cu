// Device-side noise views (reference: code/data_harvard.py:698-783 under --condition noise
// --condition_name Gaussian): per sample the reference's loader reseeds numpy (np.random.seed(seed_idx), :698), draws a
// zero-variance field for the clean view (clip(x + N(0, 0), 0, 1), :722-731) and a sigma = 0.5 field for the noisy one
// (clip(x + N(0, 0.5), 0, 1), :769-783) over 96^3 + 3 x 384^2 values on a CPU worker -- at batch 64 that, not the model,
// sets the step time (bench.py `reference_drivers`: 200 ms per step around a 15 ms model forward).  Here one kernel reads
// the preprocessed volume once (fp32 in [0, 1], or uint8 with the / 255 of :694-695 fused) and writes both views.
//
// Noise: Philox4x32-10 keyed by the seed, counter = element index / 4 (+ the item index unless `shared_field`: the
// reference reseeds with the SAME seed for every item, so all items of a batch carry the same field; shared_field = 1
// reproduces that, 0 gives independent fields), Box-Muller on the four words.  A counter-based generator cannot replay
// numpy's MT19937 / ziggurat stream, so parity has two legs (tests/test_gpu_views.py): with an INJECTED noise tensor the
// views equal the reference formula to the last bit of fp32; with the device generator the field is N(0, sigma) by its
// moments, a pure function of (seed, index), and the clipping statistics match.
// HBM-bound: 4 (or 1) bytes read + 8 written per element.
#include <stdint.h>

#include "../../include/edrl_b200.h"
#include "common.cuh"

namespace edrl {
namespace views {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t lo0 = 0xD2511F53u * c0, hi0 = __umulhi(0xD2511F53u, c0);
    const uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = __umulhi(0xCD9E8D57u, c2);
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// four N(0, 1) values from four 32-bit words (two Box-Muller pairs; u in (0, 1])
__device__ __forceinline__ void normals4(const uint32_t (&w)[4], float (&n)[4]) {
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const float u1 = ((float)w[2 * p] + 1.0f) * 2.3283064365386963e-10f;        // (0, 1]
    const float u2 = (float)w[2 * p + 1] * 2.3283064365386963e-10f;             // [0, 1)
    // (MUFU lg2 / sin / cos: the kernel is otherwise bound by the transcendental sequences, not by HBM; a noise field
    //  does not need the last two bits of a logarithm)
    const float rad = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    n[2 * p] = rad * c;
    n[2 * p + 1] = rad * s;
  }
}

template <bool U8>
__global__ void __launch_bounds__(256)
noise_views_kernel(const void *__restrict__ xin, long long per_item, int items, float sigma, unsigned long long seed,
                   int shared_field, const float *__restrict__ noise, float *__restrict__ low,
                   float *__restrict__ high) {
  const long long total4 = (per_item * items + 3) / 4;
  const long long total = per_item * items;
#pragma unroll 2                                             // (two independent quads in flight per thread)
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total4;
       q += (long long)gridDim.x * blockDim.x) {
    const long long e0 = q * 4;
    float x[4], nz[4];
    if (U8) {
      const uint8_t *xb = reinterpret_cast<const uint8_t *>(xin);
      if (e0 + 3 < total) {
        const uchar4 v = *reinterpret_cast<const uchar4 *>(xb + e0);
        x[0] = v.x / 255.0f; x[1] = v.y / 255.0f; x[2] = v.z / 255.0f; x[3] = v.w / 255.0f;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = (e0 + i < total) ? xb[e0 + i] / 255.0f : 0.f;
      }
    } else {
      const float *xf = reinterpret_cast<const float *>(xin);
      if (e0 + 3 < total) {
        const float4 v = *reinterpret_cast<const float4 *>(xf + e0);
        x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = (e0 + i < total) ? xf[e0 + i] : 0.f;
      }
    }
    if (noise != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) nz[i] = (e0 + i < total) ? noise[e0 + i] : 0.f;
    } else {
      // the field's coordinates: element index inside the item (per_item % 4 == 0 is not required: the counter is the
      // global quad when fields are independent, the in-item element otherwise, one Philox call per element then)
      uint32_t w[4];
      if (!shared_field) {
        philox4x32_10((uint32_t)q, (uint32_t)(q >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
        normals4(w, nz);
      } else if (per_item % 4 == 0) {
        const long long qi = (e0 % per_item) / 4;
        philox4x32_10((uint32_t)qi, (uint32_t)(qi >> 32), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
        normals4(w, nz);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long long ei = (e0 + i) % per_item;
          float t[4];
          philox4x32_10((uint32_t)(ei >> 2), (uint32_t)(ei >> 34), 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
          normals4(w, t);
          nz[i] = t[ei & 3];
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) nz[i] *= sigma;
    }
    float lo[4], hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      lo[i] = fminf(fmaxf(x[i], 0.f), 1.f);                  // clip(x + N(0, 0)): the clean view
      hi[i] = fminf(fmaxf(x[i] + nz[i], 0.f), 1.f);
    }
    if (e0 + 3 < total) {
      *reinterpret_cast<float4 *>(low + e0) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<float4 *>(high + e0) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (e0 + i < total) {
          low[e0 + i] = lo[i];
          high[e0 + i] = hi[i];
        }
    }
  }
}

}  // namespace views
}  // namespace edrl

using namespace edrl;

extern "C" {

int edrl_noise_views(const void *x, int is_u8, long long per_item, int items, float sigma, unsigned long long seed,
                     int shared_field, const float *noise, float *low, float *high, void *stream) {
  EDRL_DEVICE_GUARD();
  EDRL_CHECK_ARG(x && low && high, "noise_views: null argument");
  EDRL_CHECK_ARG(per_item > 0 && items > 0, "noise_views: bad shape per_item=%lld items=%d", per_item, items);
  EDRL_CHECK_ARG(((reinterpret_cast<uintptr_t>(low) | reinterpret_cast<uintptr_t>(high)) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(x) & (is_u8 ? 3 : 15)) == 0,
                 "noise_views: buffers must be 16-byte aligned (4 for uint8 input)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total4 = (per_item * items + 3) / 4;
  int sms = device_sm_count();
  if (sms <= 0) sms = 148;
  long long blocks = (total4 + 255) / 256;
  const long long cap = (long long)sms * 32;
  if (blocks > cap) blocks = cap;
  if (is_u8)
    views::noise_views_kernel<true><<<(int)blocks, 256, 0, st>>>(x, per_item, items, sigma, seed, shared_field, noise, low, high);
  else
    views::noise_views_kernel<false><<<(int)blocks, 256, 0, st>>>(x, per_item, items, sigma, seed, shared_field, noise, low, high);
  EDRL_LAUNCHED();
  return 0;
}

}  // extern "C"
