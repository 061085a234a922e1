// K7d "sift" select: top-k of a row by one warp (G = 32 lanes; the helpers are written for a group of G), built around the
// instruction count -- the select is bound by instruction issue and by shared-memory wavefronts, not by HBM (ncu on
// topk_vec_kernel: 87 % issue-active, 41 % DRAM).  Included by eprl.cu after the row accessors, emit_winners() and
// radix_select_row_slow().
//
// topk_vec_kernel histograms EVERY element of the row (an FADD, an FFMA, an address and a shared-memory reduction each),
// collects the threshold bin with a second pass over every element and compacts the winners with a third, and pays
// about 700 instructions per row for the warp-wide scans, histogram walks and the emit around those passes.  Here
//   * only the elements that can still matter go through the histogram machinery:
//       1. a strided SAMPLE of the row (NS elements per lane) is histogrammed over 64 value bins between the sample's
//          min and max; walking that histogram from the top gives a pivot p with about `jtarget / (G NS)` of the row
//          above it -- the host chooses jtarget so that the row's count above p is k + 3.3 sigma of the sampling error;
//       2. ONE counting pass (compare + predicated add per element) and ONE storing pass (compare + predicated 64-bit
//          store + predicated pointer bump) move the SURVIVORS (x >= p: about 1.6 k of them, at most CAP) to a compact
//          list in shared memory;
//       3. the exact select runs on the survivors only, SL per lane from coalesced shared-memory loads: a 64-bin value
//          histogram between p and the row maximum, the threshold bin's (<= 32) values ranked against each other with
//          shuffles by (value descending, index ascending) -- which IS the tie rule, so there is no separate tie path --
//          winners above the bin compacted, the ranked candidates appended behind them;
// Measured and dropped: a half-warp per row (every shuffle, vote and scan step issued once for two rows; with 16
// survivor slots and 4 histogram bins per lane the per-row instruction count did not drop: 46 % of HBM against 57 %), and
// persistent warps that issue the next row's loads under the exact select of the current one (same time as one row per
// warp at equal occupancy, 0.332 against 0.329 ms on 2^18 x 800 -- and the registers it needs cost two resident blocks
// per SM, which costs more: the kernel lives on occupancy, 0.281 ms at 8 blocks per SM).
// Every bin index is a monotone function of the value (one FFMA: y = fma(x, scale, off) with off = 2^23 + 1 - lo * scale
// rounded once; the bin is y's low mantissa bits), which is all the select needs: bin(a) > bin(b) implies a > b.
// Anything irregular -- NaN or +inf in the row (NaN-propagating 3-input max), a constant or non-finite sample, a pivot
// that leaves fewer than k or more than CAP survivors, a threshold bin with more than 32 values, a range so narrow
// that `lo * scale` loses integer precision -- takes radix_select_row_slow (the integer-key radix select on the whole
// warp, out of line), so the result is exact for every input; only the speed depends on the row looking like a sample of
// itself.
#pragma once

namespace sift {

constexpr int SBINS = 64;                     // bins of both value histograms
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float max3_nan(float a, float b, float c) {
  float d;
  asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float d;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float max2_nan(float a, float b) {
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// reductions over the group of G lanes that holds this lane (G = 32: one redux; G = 16: a butterfly both halves share)
template <int G>
__device__ __forceinline__ float group_max_nan(float v) {
  if (G == 32) {
    float d;
    asm volatile("redux.sync.max.NaN.f32 %0, %1, 0xffffffff;" : "=f"(d) : "f"(v));
    return d;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = max2_nan(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ float group_max(float v) {
  if (G == 32) {
    float d;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(d) : "f"(v));
    return d;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
template <int G>
__device__ __forceinline__ float group_min(float v) {
  if (G == 32) {
    float d;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(d) : "f"(v));
    return d;
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}
// inclusive prefix sum over the group: SHFL.UP sets the in-range predicate itself, one predicated add per step
template <int G>
__device__ __forceinline__ int group_incl_scan(int v) {
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .b32 t;\n\t"
        "shfl.sync.up.b32 t|q, %0, %1, %2, 0xffffffff;\n\t"
        "@q add.s32 %0, %0, t;\n\t"
        "}\n"
        : "+r"(v)
        : "r"(o), "r"((32 - G) << 8));
  }
  return v;
}
// this group's bits of a warp ballot, in bits [0, G)
template <int G>
__device__ __forceinline__ uint32_t group_ballot(bool pred, int lane) {
  const uint32_t b = __ballot_sync(FULL, pred);
  return (G == 32) ? b : ((b >> (lane & 16)) & 0xffffu);
}

// Counting x >= p without predicates (ptxas parks every predicate of an unrolled compare chain in a bit mask, five
// instructions per element instead of two): the sign bit of x - p is 0 exactly when x >= p (IEEE subtraction has the sign
// of the true difference, denormals included; a -inf pad gives -inf), and one funnel shift appends it to an accumulator.
// After n elements: n - popc(acc) of them were >= p.
__device__ __forceinline__ void push_sign(uint32_t &acc, float x, float p) {
  acc = __funnelshift_l(__float_as_uint(x - p), acc, 1);
}
// The affine map of a value histogram over [lo, hi]: bin(x) = low bits of fma(x, scale, off), in [0, SBINS - 1] for
// lo <= x <= hi (the range maps onto SBINS - 3 bins, + 1 of offset, +- 1/2 for each of the two roundings).  ok = false
// when the range is empty / not finite or too narrow for `lo * scale` to keep integer precision (|lo| scale >= 2^22).
struct BinMap {
  float scale, off;
  bool ok;
};
__device__ __forceinline__ BinMap make_binmap(float lo, float hi) {
  BinMap m;
  const float range = hi - lo;
  m.scale = __fdividef((float)(SBINS - 3), range);
  m.off = fmaf(-lo, m.scale, 8388609.f);                     // 2^23 + 1 - lo * scale, one rounding (to an integer)
  m.ok = (range > 0.f) && (range <= 3.0e38f) && (fabsf(lo) * m.scale < 4.0e6f) && (m.scale <= 3.0e38f);
  return m;
}

// Walk a SBINS-bin histogram from the top (lane 0 of the group holds the largest bins): the bin in which the
// cumulative count reaches `need`, its own count, and how many of its elements are wanted.  false (for this group):
// fewer than `need` elements in all.
template <int G>
__device__ __forceinline__ bool walk_from_top(const unsigned int *hist, int need, int lane, int &bin, int &cntb,
                                              int &rem) {
  constexpr int BPL = SBINS / G;                             // 2 or 4 bins per lane
  const int gl = lane & (G - 1);
  int c[BPL];                                                // c[0] = this lane's highest bin
  if (BPL == 2) {
    const uint2 h = reinterpret_cast<const uint2 *>(hist)[G - 1 - gl];
    c[0] = (int)h.y;
    c[1] = (int)h.x;
  } else {
    const uint4 h = reinterpret_cast<const uint4 *>(hist)[G - 1 - gl];
    c[0] = (int)h.w;
    c[1] = (int)h.z;
    c[2] = (int)h.y;
    c[3] = (int)h.x;
  }
  int t = 0;
#pragma unroll
  for (int i = 0; i < BPL; ++i) t += c[i];
  const int incl = group_incl_scan<G>(t);
  if (BPL == 2) {
    // branch-free: every lane works out the answer of its own two bins, the one lane whose range [excl, incl) holds
    // position need - 1 publishes it
    const int excl = incl - t;
    const bool mine = excl < need && incl >= need;
    const uint32_t hit = group_ballot<G>(mine, lane);
    const int want = need - excl;                            // wanted from this lane's two bins
    const bool first = c[0] >= want;
    const int bn = 2 * (G - 1 - gl) + (first ? 1 : 0);
    const int cb = first ? c[0] : c[1];
    const int rm = first ? want : want - c[0];
    const int pk = bn | (cb << 8) | (rm << 20);
    const int got = __shfl_sync(FULL, pk, (__ffs(hit | 0x80000000u) - 1 + (lane & ~(G - 1))) & 31);
    bin = got & 255;
    cntb = (got >> 8) & 4095;
    rem = got >> 20;
    return hit != 0u;
  }
  const uint32_t hit = group_ballot<G>(incl >= need, lane);
  const int L = __ffs(hit) - 1;                              // -1: no lane reaches `need`
  int packed = 0;                                            // bin | cntb << 8 | rem << 20
  if (gl == L) {
    int r0 = need - (incl - t);                              // wanted from this lane's bins
    int bn = 0, cb = 0, rm = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < BPL; ++i) {
      if (!found) {
        if (c[i] >= r0) {
          bn = BPL * (G - 1 - gl) + (BPL - 1 - i);
          cb = c[i];
          rm = r0;
          found = true;
        } else {
          r0 -= c[i];
        }
      }
    }
    packed = bn | (cb << 8) | (rm << 20);
  }
  packed = __shfl_sync(FULL, packed, (L < 0 ? 0 : L) + (lane & ~(G - 1)));
  bin = packed & 255;
  cntb = (packed >> 8) & 4095;
  rem = packed >> 20;
  return L >= 0;
}

// Stage 3: exact top-k of the C survivors in list[0, C) (uint2: .x = ~index, .y = value bits; all values in [p, rowmax],
// k <= C <= G SL) by the group.  On success the winners are in list[0, k) (list[k, 128) zero when SORTED) and T is the
// k-th largest value; `list` is reused as the winners' buffer once its entries are in registers.  false: crowded
// threshold bin or a degenerate value range -- the caller takes the slow path for this row.
// Every lane of the warp must call this (full-mask shuffles); `active` = this group has a row to select.
template <int G, int SL, bool SORTED>
__device__ __forceinline__ bool select_survivors(bool active, uint2 *list, uint2 *cand, unsigned int *hist, int C,
                                                 float p, float rowmax, int k, int lane, float &T) {
  constexpr int CPL = 32 / G;                                // candidates per lane (the threshold bin holds <= 32)
  const int gl = lane & (G - 1);
  const BinMap bm = make_binmap(p, rowmax);
  bool ok = active && bm.ok;
  const uint32_t hist_addr = (uint32_t)__cvta_generic_to_shared(hist);
  const uint32_t hbase = hist_addr - (0x4b000000u << 2);
  const int nv = ok ? (C - gl + G - 1) / G : 0;              // slots i < nv of this lane hold survivors
  uint2 s[SL];
#pragma unroll
  for (int i = 0; i < SL; ++i) s[i] = list[i * G + gl];
  if (gl < SBINS / 4) reinterpret_cast<uint4 *>(hist)[gl] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();                                              // (also: every lane has read its list entries)
  uint32_t yb[SL];                                           // bin code of slot i, 0 for an empty slot (below every bin)
#pragma unroll
  for (int i = 0; i < SL; ++i) {
    // (an empty slot counts into the spare word behind the bins: a predicated reduction becomes a branch region around
    //  the warp-aggregated ATOMS, four instructions instead of two)
    const uint32_t y = __float_as_uint(fmaf(__uint_as_float(s[i].y), bm.scale, bm.off));
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .b32 a;\n\t"
        "setp.lt.s32 q, %2, %3;\n\t"
        "selp.b32 a, %1, %5, q;\n\t"
        "red.shared.add.u32 [a], 1;\n\t"
        "selp.b32 %0, %4, 0, q;\n\t"
        "}\n"
        : "=r"(yb[i])
        : "r"((y << 2) + hbase), "r"(i), "r"(nv), "r"(y), "r"(hist_addr + 4u * SBINS)
        : "memory");
  }
  __syncwarp();
  int bin, cntb, rem;
  ok = walk_from_top<G>(hist, k, lane, bin, cntb, rem) && ok;
  ok = ok && cntb <= 32 && rem >= 1 && rem <= cntb;
  const uint32_t yt = ok ? 0x4b000000u + (uint32_t)bin : 0x7fffffffu;     // not ok: nothing is above or in the bin
  // winners above the threshold bin and the bin's candidates: lane-local counts, one group prefix for both.  Sign bits
  // again (bin codes are < 2^31): yt - y < 0 above the bin, yt - y - 1 < 0 in or above it.
  uint32_t above = 0u, notbelow = 0u;
#pragma unroll
  for (int i = 0; i < SL; ++i) {
    const uint32_t dw = yt - yb[i];
    above = __funnelshift_l(dw, above, 1);
    notbelow = __funnelshift_l(dw - 1u, notbelow, 1);
  }
  const int cw = __popc(above), cc = __popc(notbelow) - cw;
  const int mine = cw | (cc << 16);
  const int incl = group_incl_scan<G>(mine);
  const int tot = __shfl_sync(FULL, incl, (lane & ~(G - 1)) + G - 1);
  const int nwin = tot & 0xffff;                             // == k - rem
  const int excl = incl - mine;
  if (SORTED) {                                              // the sort networks read all 128 slots: empty ones are 0
    for (int t2 = gl; t2 < 64; t2 += G) reinterpret_cast<uint4 *>(list)[t2] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
  }
  uint32_t pw = (uint32_t)__cvta_generic_to_shared(list) + 8u * (uint32_t)(excl & 0xffff);
  uint32_t pc = (uint32_t)__cvta_generic_to_shared(cand) + 8u * (uint32_t)(excl >> 16);
#pragma unroll
  for (int i = 0; i < SL; ++i) {
    asm volatile(
        "{\n\t"
        ".reg .pred q, r;\n\t"
        "setp.gt.u32 q, %4, %5;\n\t"
        "setp.eq.u32 r, %4, %5;\n\t"
        "@q st.shared.v2.u32 [%0], {%2, %3};\n\t"
        "@q add.u32 %0, %0, 8;\n\t"
        "@r st.shared.v2.u32 [%1], {%2, %3};\n\t"
        "@r add.u32 %1, %1, 8;\n\t"
        "}\n"
        : "+r"(pw), "+r"(pc)
        : "r"(s[i].x), "r"(s[i].y), "r"(yb[i]), "r"(yt)
        : "memory");
  }
  __syncwarp();
  // rank the candidates by (value descending, index ascending): the first `rem` of that order win.  Every lane walks
  // the candidate list in shared memory (broadcast loads, independent of one another: they pipeline, where a shuffle
  // per candidate and operand did not) and counts the entries that come before its own.
  uint2 me[CPL];
  int rank[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const bool have = ok && (q * G + gl) < cntb;
    me[q] = have ? cand[q * G + gl] : make_uint2(0u, 0xff800000u);       // (-inf, index 2^32 - 1: behind everything)
    rank[q] = 0;
  }
  const int cn = ok ? cntb : 0;
#pragma unroll 4
  for (int j = 0; j < cn; ++j) {
    const uint2 o = cand[j];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
      // rank += (ov > cv) || (ov == cv && o.x > me.x)   (.x = ~index: larger = lower index = earlier), without branches
      asm("{\n\t"
          ".reg .pred a, b;\n\t"
          "setp.eq.f32 b, %1, %2;\n\t"
          "setp.gt.and.u32 b, %3, %4, b;\n\t"
          "setp.gt.or.f32 a, %1, %2, b;\n\t"
          "@a add.s32 %0, %0, 1;\n\t"
          "}\n"
          : "+r"(rank[q])
          : "f"(__uint_as_float(o.y)), "f"(__uint_as_float(me[q].y)), "r"(o.x), "r"(me[q].x));
    }
  }
  float tv = __int_as_float(0xff800000);
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const bool have = ok && (q * G + gl) < cntb;
    if (have && rank[q] < rem) list[nwin + rank[q]] = me[q];
    if (have && rank[q] == rem - 1) tv = __uint_as_float(me[q].y);
  }
  T = group_max<G>(tv);                                      // exactly one candidate has rank rem - 1
  __syncwarp();
  return ok;
}

// The pivot from a sample held in registers (NS values per lane): a value with at least `jtarget` sample elements at
// or above it (about jtarget of the G NS, plus the pivot's bin).  false: degenerate / non-finite sample.
template <int G, int NS>
__device__ __forceinline__ bool sample_pivot(bool ok, const float (&sv)[NS], unsigned int *hist, int jtarget, int lane,
                                             float &p) {
  const int gl = lane & (G - 1);
  float lo = sv[0], hi = sv[0];
#pragma unroll
  for (int i = 1; i + 1 < NS; i += 2) {
    lo = min3(lo, sv[i], sv[i + 1]);
    hi = max3(hi, sv[i], sv[i + 1]);
  }
  if ((NS & 1) == 0) {
    lo = fminf(lo, sv[NS - 1]);
    hi = fmaxf(hi, sv[NS - 1]);
  }
  lo = group_min<G>(lo);
  hi = group_max<G>(hi);
  const BinMap bm = make_binmap(lo, hi);
  ok = ok && bm.ok;
  const uint32_t hbase = (uint32_t)__cvta_generic_to_shared(hist) - (0x4b000000u << 2);
  if (gl < SBINS / 4) reinterpret_cast<uint4 *>(hist)[gl] = make_uint4(0u, 0u, 0u, 0u);
  __syncwarp();
  if (ok) {                                                  // (a NaN in the row or a degenerate map: no wild addresses)
#pragma unroll
    for (int i = 0; i < NS; ++i) {
      const uint32_t y = __float_as_uint(fmaf(sv[i], bm.scale, bm.off));
      asm volatile("red.shared.add.u32 [%0], 1;" ::"r"((y << 2) + hbase) : "memory");
    }
  }
  __syncwarp();
  int bin, cntb, rem;
  const bool found = walk_from_top<G>(hist, jtarget, lane, bin, cntb, rem);
  // the lower edge of `bin`: fma(x, scale, off) - 2^23 = bin - 1/2  <=>  x = lo + (bin - 3/2) / scale
  p = fmaf((float)bin - 1.5f, __fdividef(hi - lo, (float)(SBINS - 3)), lo);
  __syncwarp();                                              // the histogram is free again
  return found && ok;
}

}  // namespace sift

// One warp per row: uniform width W = 4 (32 FI + partial lanes) <= 2048, the row in registers (LDG.128), k <= 128.
// SSTR: sampling stride over the lane's full-iteration elements; SL: survivor slots per lane (capacity 32 SL >= 128).
template <int FI, bool PARTIAL, int SSTR, int SL, bool SORTED, class Rows, int BLK>
__global__ void __launch_bounds__(128, BLK)
topk_sift_kernel(Rows rows, int R, int W, int k, int jtarget, float *__restrict__ vals, int *__restrict__ idx) {
  constexpr int G = 32;
  constexpr int NI = FI + (PARTIAL ? 1 : 0);
  constexpr int E = NI * 4;
  constexpr int NS = (FI * 4 + SSTR - 1) / SSTR;
  constexpr int CAP = G * SL;
  static_assert(FI >= 2 && NS >= 4 && CAP >= 128, "sift select: row too narrow / list too small for the winners' buffer");
  __shared__ __align__(16) uint2 s_list[4][CAP];             // per warp: survivors, later the winners (first 128)
  __shared__ __align__(16) uint2 s_cand[4][32];
  __shared__ __align__(16) unsigned int s_hist[4][sift::SBINS + 4];        // + a spare word for empty slots
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + wib;
  if (r >= R) return;
  uint2 *list = s_list[wib];
  const typename Rows::Cursor cur = rows.cursor(r);
  const bool pvalid = PARTIAL && lane < ((W >> 2) & 31);
  float x[E];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const float ninf = __int_as_float(0xff800000);           // lanes without data: below every pivot
    float4 v = make_float4(ninf, ninf, ninf, ninf);
    if (i < FI || pvalid) v = __ldg(reinterpret_cast<const float4 *>(cur.at((i * G + lane) * 4)));
    x[4 * i + 0] = v.x;
    x[4 * i + 1] = v.y;
    x[4 * i + 2] = v.z;
    x[4 * i + 3] = v.w;
  }
  // row maximum, NaN-propagating: NaN or +inf anywhere sends the row to the slow path
  float mx = x[0];
#pragma unroll
  for (int e = 1; e + 1 < E; e += 2) mx = sift::max3_nan(mx, x[e], x[e + 1]);
  mx = sift::max2_nan(mx, x[E - 1]);
  mx = sift::group_max_nan<G>(mx);
  bool ok = mx < __int_as_float(0x7f800000);
  float sv[NS];
#pragma unroll
  for (int i = 0; i < NS; ++i) sv[i] = x[i * SSTR];
  float p;
  ok = sift::sample_pivot<G, NS>(ok, sv, s_hist[wib], jtarget, lane, p);
  uint32_t neg[4] = {0u, 0u, 0u, 0u};                        // four chains: E / 4 <= 16 sign bits each
#pragma unroll
  for (int e = 0; e < E; ++e) sift::push_sign(neg[e & 3], x[e], p);
  const int cnt = E - (__popc(neg[0]) + __popc(neg[1])) - (__popc(neg[2]) + __popc(neg[3]));
  const int incl = sift::group_incl_scan<G>(cnt);
  const int C = __shfl_sync(sift::FULL, incl, 31);
  ok = ok && C >= k && C <= CAP;
  if (ok) {                                                  // (warp-uniform)
    // (the pivot again, through a shuffle ptxas cannot see through: with the same register in both passes it keeps the
    //  counting pass's 28 predicates in a bit mask for this one -- two LOP3 per element there, an unpack here)
    const float ps = __shfl_sync(sift::FULL, p, 0);
    uint32_t pos = (uint32_t)__cvta_generic_to_shared(list) + 8u * (uint32_t)(incl - cnt);
    const uint32_t nlane4 = ~((uint32_t)lane << 2);          // ~(4 lane + c) = ~(4 lane) - c
#pragma unroll
    for (int e = 0; e < E; ++e) {
      asm volatile(
          "{\n\t"
          ".reg .pred q;\n\t"
          "setp.ge.f32 q, %2, %3;\n\t"
          "@q st.shared.v2.u32 [%0], {%1, %4};\n\t"
          "@q add.u32 %0, %0, 8;\n\t"
          "}\n"
          : "+r"(pos)
          : "r"(nlane4 - (uint32_t)((e >> 2) * (4 * G) + (e & 3))), "f"(x[e]), "f"(ps), "r"(__float_as_uint(x[e]))
          : "memory");
    }
  }
  __syncwarp();
  float T;
  ok = sift::select_survivors<G, SL, SORTED>(ok, list, s_cand[wib], s_hist[wib], C, p, mx, k, lane, T);
  if (ok) {
    emit_winners<SORTED>(list, k, lane, f2key_fast(T), f2key_fast(mx), vals + (size_t)r * k, idx + (size_t)r * k);
  } else {
    // (radix_select_row_slow wants 256 words of histogram that later hold 128 64-bit winners: the survivors' list)
    radix_select_row_slow<E, SORTED, Rows>(rows, r, k, reinterpret_cast<unsigned int *>(list), lane, vals, idx);
  }
}

// One warp per row for 2048 < W (uniform width, W % 4 == 0, 16-byte aligned rows, k <= 128): the same sift with the
// row STREAMED through registers instead of held in them -- 32 sample values per lane (8 float4 spread over the row),
// then chunks of 8 float4 per lane: running maximum, count, group prefix, store.  No block-wide barrier anywhere (the
// block-per-row kernel spends its time in them: 35 % of HBM at 2^15 x 8192).  A row that falls out of the fast path
// (NaN / +inf, degenerate sample, survivors not in [k, 512], crowded threshold bin) is marked with idx[r k] = -1 and
// redone by topk_vecblock_kernel's marked-row pass, launched right behind this kernel.
template <bool SORTED, class Rows>
__global__ void __launch_bounds__(128, 8)      // (64 registers; 6 blocks / 80 registers: 77.8 / 67.5 % against 79.1 / 69.1 %)
topk_sift_stream_kernel(Rows rows, int R, int W, int k, int jtarget, float *__restrict__ vals, int *__restrict__ idx) {
  constexpr int SL = 16, CAP = 32 * SL, CH = 8;              // CH float4 per lane and chunk
  __shared__ __align__(16) uint2 s_list[4][CAP];
  __shared__ __align__(16) uint2 s_cand[4][32];
  __shared__ __align__(16) unsigned int s_hist[4][sift::SBINS + 4];        // + a spare word for empty slots
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + wib;
  if (r >= R) return;
  uint2 *list = s_list[wib];
  const typename Rows::Cursor cur = rows.cursor(r);
  const int W4 = W >> 2;
  const int NT = (W4 + 31) >> 5;                             // float4 iterations of the warp over the row
  const int NTF = W4 >> 5;                                   // ... of which full
  // ---- sample: iterations 0, st, 2 st, ... (all full: st * 7 < NTF)
  const int st = NTF / CH;
  float sv[CH * 4];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const float4 v = __ldg(reinterpret_cast<const float4 *>(cur.at(((i * st) * 32 + lane) * 4)));
    sv[4 * i + 0] = v.x;
    sv[4 * i + 1] = v.y;
    sv[4 * i + 2] = v.z;
    sv[4 * i + 3] = v.w;
  }
  float smx = sv[0];
#pragma unroll
  for (int e = 1; e + 1 < CH * 4; e += 2) smx = sift::max3_nan(smx, sv[e], sv[e + 1]);
  smx = sift::group_max_nan<32>(sift::max2_nan(smx, sv[CH * 4 - 1]));
  float p;
  bool ok = sift::sample_pivot<32, CH * 4>(smx < __int_as_float(0x7f800000), sv, s_hist[wib], jtarget, lane, p);
  // ---- stream the row: maximum (NaN-propagating), survivors to the list
  float mx = __int_as_float(0xff800000);
  int C = 0;
  const float ninf = __int_as_float(0xff800000);
  for (int t0 = 0; t0 < NT && ok; t0 += CH) {
    float x[CH * 4];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      float4 v = make_float4(ninf, ninf, ninf, ninf);
      if ((t0 + i) * 32 + lane < W4) v = __ldg(reinterpret_cast<const float4 *>(cur.at(((t0 + i) * 32 + lane) * 4)));
      x[4 * i + 0] = v.x;
      x[4 * i + 1] = v.y;
      x[4 * i + 2] = v.z;
      x[4 * i + 3] = v.w;
    }
    int cnt = 0;
#pragma unroll
    for (int e = 0; e + 1 < CH * 4; e += 2) mx = sift::max3_nan(mx, x[e], x[e + 1]);
    uint32_t neg[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int e = 0; e < CH * 4; ++e) sift::push_sign(neg[e & 3], x[e], p);
    cnt = CH * 4 - (__popc(neg[0]) + __popc(neg[1])) - (__popc(neg[2]) + __popc(neg[3]));
    const int incl = sift::group_incl_scan<32>(cnt);
    const int tot = __shfl_sync(sift::FULL, incl, 31);
    if (C + tot > CAP) {                                     // (uniform) too many survivors: the marked-row pass
      ok = false;
      break;
    }
    const float ps = __shfl_sync(sift::FULL, p, 0);          // (opaque copy: see topk_sift_kernel)
    uint32_t pos = (uint32_t)__cvta_generic_to_shared(list) + 8u * (uint32_t)(C + incl - cnt);
    const uint32_t nbase = ~((uint32_t)(t0 * 32 + lane) << 2);           // ~(4 (32 t0 + lane) + c) = nbase - c
#pragma unroll
    for (int e = 0; e < CH * 4; ++e) {
      asm volatile(
          "{\n\t"
          ".reg .pred q;\n\t"
          "setp.ge.f32 q, %2, %3;\n\t"
          "@q st.shared.v2.u32 [%0], {%1, %4};\n\t"
          "@q add.u32 %0, %0, 8;\n\t"
          "}\n"
          : "+r"(pos)
          : "r"(nbase - (uint32_t)((e >> 2) * 128 + (e & 3))), "f"(x[e]), "f"(ps), "r"(__float_as_uint(x[e]))
          : "memory");
    }
    C += tot;
  }
  mx = sift::group_max_nan<32>(mx);
  ok = ok && (mx < __int_as_float(0x7f800000)) && C >= k;
  __syncwarp();
  float T;
  ok = sift::select_survivors<32, SL, SORTED>(ok, list, s_cand[wib], s_hist[wib], C, p, mx, k, lane, T) && ok;
  if (ok) {
    emit_winners<SORTED>(list, k, lane, f2key_fast(T), f2key_fast(mx), vals + (size_t)r * k, idx + (size_t)r * k);
  } else if (lane == 0) {
    idx[(size_t)r * k] = -1;                                 // marked: redone by the block-per-row kernel
  }
}
