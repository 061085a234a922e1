// mmd_bwd.cuh -- Part A, K3: separate tile-recomputing backward kernels (single CTA and cta_group::2 pair)
// (textually included by mmd.cu inside namespace edrl::mmd; not a stand-alone header)
#pragma once

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
  // a launch may own only the row panels [panel0, panel0 + panels) of the call's row ranges (the hybrid quad + pair launch
  // for d > 768 gives the last panels to a pair kernel on the SMs the 4-CTA clusters cannot use); ticket_total = CTAs of
  // all launches that share the accumulators (0: this launch alone)
  int panel0, ticket_total;
  int s_ahead;                 // quad sweep: issue order of the S phases (mmd_sweep_quad_kernel)
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


// ----------------------------------------------------------------------------- shared-memory units of the pair sweeps
// (mmd_sweep.cuh; the separate CTA-pair backward that first used them -- 128-column S tiles, 0.32 of the TF32 roofline --
//  is gone: edrl_mmd_backward runs the fused sweep and apply_grad in place)
constexpr int P2_STAGE = 16384;                 // ring stage per CTA
constexpr int P2_CHUNK = 8192;                  // 64 rows x 32 floats, 128-byte swizzle
constexpr int P2_FEATS = 512;                   // feature columns per pair and pass (2 M-tiles of 256)
