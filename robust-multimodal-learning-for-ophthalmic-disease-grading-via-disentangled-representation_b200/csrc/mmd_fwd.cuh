// mmd_fwd.cuh -- Part A, K2: loss-only forward kernels (single-CTA tiles and the cta_group::2 pair kernel)
// (textually included by mmd.cu inside namespace edrl::mmd; not a stand-alone header)
#pragma once

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

