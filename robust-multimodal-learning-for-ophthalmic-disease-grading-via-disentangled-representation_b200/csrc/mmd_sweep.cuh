// mmd_sweep.cuh -- Part A, K3q: the fused training sweep (persistent pair kernel, 4-CTA quad kernel) and apply_grad
// (textually included by mmd.cu inside namespace edrl::mmd; not a stand-alone header)
#pragma once

// ----------------------------------------------------------------------------- K3q: CTA-pair sweep, 256-column S tiles
// One 2-CTA cluster (an SM pair, tcgen05 cta_group::2) owns a 128-row panel I of Z and 512 feature columns:
//   S phase   S_IJ = Z_I Z_J^T                 M = 128 (64 rows of the panel per CTA), N = 256 (128 rows of Z_J per CTA), K = d
//   epilogue  each CTA turns its 64 x 256 slice of S into G in its own shared memory
//   P phase   dZ^T[f, i] += Zt[f, j] G[i, j]   M = 256 features (128 per CTA), N = 128 rows i (the two CTAs' G halves are the
//             two halves of the B operand -- no exchange), K = 256 columns j
// The S phase works on TWO 128-column tiles at once: with 64 A rows per CTA an N = 128 S phase (the first version of this
// decomposition, the deleted separate backward kernel) re-read 4 KiB of shared memory per 32-clk MMA (the whole
// 128 B/clk port); N = 256 reads 6 KiB per 64 clk, the Z_I chunk is fetched once per 256 columns instead of per 128,
// and the MMAs are twice as long (half the issue slots).  Z_I is streamed (two 32-column chunks per ring stage), the
// ring has 9 stages of 16 KiB, G is 64 rows x 256 columns (64 KiB) per CTA.
constexpr int Q_GROUP = 256;                   // columns per S group
constexpr int Q_G_BYTES = 8 * P2_CHUNK;        // 64 rows x 256 columns j
constexpr int Q_CTRL_BYTES = 8192;
constexpr int Q_CTRL_SHORT_BYTES = 3072;       // 3xTF32: the control block up to and including its first column buffers
constexpr int SW_EPI_WARPS = 16;               // 4 per TMEM lane group: one 32-column chunk of the S stage each
constexpr int SW_EPI_THREADS = SW_EPI_WARPS * 32;
constexpr int SW_THREADS = 64 + SW_EPI_THREADS;
constexpr int SW_MAX_SPLIT = 8;                // column slabs of a split virtual panel (make_plan)

// MODE 0: TF32 everywhere.  1 (EDRL_MMD_TF32H): binary16 P phase.  2 (EDRL_MMD_F16S): the S phase too reads scaled
// binary16 operands (Z16, kind::f16): a ring stage then holds 64 feature columns instead of 32.
// The binary16 modes keep TWO G buffers (32 KiB each), so the epilogue of group g+1 overlaps the P phase of group g.
// MODE 3 (EDRL_MMD_3XTF32): hi / lo split operands, three TF32 MMAs per product, fp32-level accuracy, in the same
// accumulators.  S = Z_hi Z_hi^T + Z_lo Z_hi^T + Z_hi Z_lo^T: per 32-column K chunk the producer fetches [Z_I hi | Z_I lo],
// Z_J hi and Z_J lo (the lo rows of the operand maps sit below the hi rows) and the issuer forms the three products, so
// every hi chunk is fetched ONCE for its two products (2/3 of the bytes of three independent passes -- the sweep is bound
// by L2 -> SM ingest).  P the same over the column group: the Z_hi^T stage meets G_hi and G_lo, the Z_lo^T stage G_hi; G
// is written as its TF32 hi and lo parts (2 x 64 KiB), which leaves 5 ring stages instead of 9.
template <int MODE>
struct SweepCfg {
  static constexpr bool H16 = MODE == 1 || MODE == 2;
  static constexpr bool S16 = MODE == 2;
  static constexpr bool X3 = MODE == 3;
  static constexpr int S_COLS = S16 ? 64 : BK;                         // feature columns per 128-byte row of an S operand
  static constexpr int G_BYTES = H16 ? Q_G_BYTES / 2 : (X3 ? 2 * Q_G_BYTES : Q_G_BYTES);   // one G buffer: 64 rows x 256 columns
  static constexpr int G_BUFS = H16 ? 2 : 1;
  // 3xTF32: G as hi + lo takes 128 KiB; a SIXTH ring stage fits only with the control block cut to its first 3 KiB (one
  // (r_j, a_j) buffer instead of two, the row-sum partials on top of it: SweepCtrl)
  // (TF32 with the short block and a TENTH stage -- 64 KiB of G + 160 KiB + 3 KiB = 232448 B as well -- measured no
  //  faster: 0.854 against 0.850 ms at N=8192, d=512, 1.86 against 1.88 ms at d=1024: that mode is bound by throughput,
  //  and the extra barriers cost what the stage buys)
  static constexpr bool SHORT_CTRL = X3;
  static constexpr int STAGES = X3 ? 6 : 9;
  static constexpr int CTRL_BYTES = SHORT_CTRL ? Q_CTRL_SHORT_BYTES : Q_CTRL_BYTES;
  static constexpr int SMEM_BYTES = G_BUFS * G_BYTES + STAGES * P2_STAGE + CTRL_BYTES;
  static constexpr int P_ATOMS = H16 ? Q_GROUP / 64 : (X3 ? 2 * (Q_GROUP / BK) : Q_GROUP / BK);   // ring stages per column group and dZ^T tile
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
  float negc[MAX_KERNELS];
  float w[MAX_KERNELS];
  double red[SW_EPI_WARPS][2];
  // (r_j, a_j) of a column group, read as float4.  Buffer 0 first: the 3xTF32 mode has room for the block only up to
  // here (Q_CTRL_SHORT_BYTES) -- it uses buffer 0 for every group (one more barrier per group) and lays the row-sum
  // partials of an item over it (2 x 256 floats = 8 x 64)
  alignas(16) float col_r0[Q_GROUP];
  alignas(16) float col_a0[Q_GROUP];
  alignas(16) float col_r1[Q_GROUP];     // buffer 1: the S stage's other parity
  alignas(16) float col_a1[Q_GROUP];
  float part[8][64];              // row-sum partials of an item (2 lane halves x 4 column chunks per row)
};
static_assert(sizeof(SweepCtrl) <= Q_CTRL_BYTES, "SweepCtrl does not fit its smem slot");
static_assert(offsetof(SweepCtrl, col_r1) <= Q_CTRL_SHORT_BYTES && offsetof(SweepCtrl, col_a0) == offsetof(SweepCtrl, col_r0) + Q_GROUP * 4,
              "the short control block ends behind its first column buffers");
static_assert(SweepCfg<0>::SMEM_BYTES <= 232448 && SweepCfg<1>::SMEM_BYTES <= 232448 && SweepCfg<3>::SMEM_BYTES <= 232448,
              "smem budget");

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
  const int panel = p.panel0 + (vp - it.ypass * p.panels);   // index in the call's panel list (range 1, then range 2)
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
  constexpr bool X3 = Cfg::X3;
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
  const int kch1 = S16 ? p.d_pad / 64 : p.kchunks;        // 128-byte K chunks of an S operand row; even
  const int kchunks = kch1;                               // (3xTF32 walks kch1 single chunks, three ring stages each)

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
        if (X3) {
          // 3xTF32, per 32-column K chunk three stages: [Z_I hi | Z_I lo] (64 rows each), Z_J hi, Z_J lo (128 rows each;
          // the lo rows sit n_pad rows below the hi rows in both maps).  The issuer forms hi x hi, lo x hi and hi x lo
          // from them: each hi chunk is fetched once for its two products (48 KiB per chunk instead of 72).
          for (int kk = 0; kk < kch1; ++kk) {
            {
              uint8_t *st = acquire();
              const uint32_t bar = full0 + 8u * (uint32_t)s;
              tma_load_2d_pair_elect(st, &tm_z64, bar, kk * Cfg::S_COLS, irow);
              tma_load_2d_pair_elect(st + P2_CHUNK, &tm_z64, bar, kk * Cfg::S_COLS, irow + p.n_pad);
              next();
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint8_t *st = acquire();
              const uint32_t bar = full0 + 8u * (uint32_t)s;
              tma_load_2d_pair_elect(st, &tm_z128, bar, kk * Cfg::S_COLS, jrow + h * p.n_pad);
              next();
            }
          }
          return;
        }
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
            // 3xTF32: stage 2a holds K atom a of Z_hi^T (for G_hi and G_lo), stage 2a + 1 of Z_lo^T (d_pad rows below, for G_hi)
            const int ac = X3 ? (a8 >> 1) : a8;
            const int fr = it.f0 + t * 256 + (int)rank * 128 + ((X3 && (a8 & 1)) ? p.d_pad : 0);
            tma_load_2d_pair_elect(st, &tm_zt, bar, (it.g_begin + g) * Q_GROUP + ac * Cfg::P_ATOM_COLS, fr);
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
          if (X3) {
            for (int kk = 0; kk < kch1; ++kk) {
              mbar_wait(&ctl->full[s], ph);                  // [Z_I hi | Z_I lo] of this chunk
              tc_fence_after();
              const int sa = s;
              const uint64_t a_hi = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
              const uint64_t a_lo = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE + P2_CHUNK);
              next();
              mbar_wait(&ctl->full[s], ph);                  // Z_J hi: hi x hi, lo x hi
              tc_fence_after();
              const uint64_t b_hi = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_tf32_ss_pair_elect(d_tmem, a_hi + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc_s,
                                       (kk > 0 || k > 0) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_tf32_ss_pair_elect(d_tmem, a_lo + (uint64_t)(k * 2), b_hi + (uint64_t)(k * 2), idesc_s, 1u);
              mma_commit_pair_elect(&ctl->empty[s]);
              next();
              mbar_wait(&ctl->full[s], ph);                  // Z_J lo: hi x lo
              tc_fence_after();
              const uint64_t b_lo = make_kmajor_sw128_desc(ring_addr + s * P2_STAGE);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_tf32_ss_pair_elect(d_tmem, a_hi + (uint64_t)(k * 2), b_lo + (uint64_t)(k * 2), idesc_s, 1u);
              mma_commit_pair_elect(&ctl->empty[sa]);
              mma_commit_pair_elect(&ctl->empty[s]);
              next();
            }
          }
          for (int kc = 0; kc < (X3 ? 0 : kchunks); kc += 2) {
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
              // (3xTF32: G atoms 0-7 hold G_hi, 8-15 G_lo; the Z_hi^T stage 2a meets both, the Z_lo^T stage 2a + 1 G_hi)
              const uint64_t b_d = make_kmajor_sw128_desc(g_addr + gb * Cfg::G_BYTES + (X3 ? (a8 >> 1) : a8) * P2_CHUNK);
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
              if (X3 && !(a8 & 1)) {
                const uint64_t b_lo = make_kmajor_sw128_desc(g_addr + gb * Cfg::G_BYTES + (8 + (a8 >> 1)) * P2_CHUNK);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  mma_tf32_ss_pair_elect(tmem_dz + t * BN, a_d + (uint64_t)(k * 2), b_lo + (uint64_t)(k * 2), idesc_p, 1u);
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
        const int cbuf = Cfg::SHORT_CTRL ? 0 : b;           // (short control block: one column buffer, see SweepCtrl)
        if (Cfg::SHORT_CTRL && g > 0) named_barrier_sync(1, SW_EPI_THREADS);   // ... which every warp has finished reading
        if (et < Q_GROUP) {
          (cbuf ? ctl->col_r1 : ctl->col_r0)[et] = (float)nxt_r;
          (cbuf ? ctl->col_a1 : ctl->col_a0)[et] = nxt_a;
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
        const float4 *cr4 = reinterpret_cast<const float4 *>(&(cbuf ? ctl->col_r1 : ctl->col_r0)[j0]);
        const float4 *ca4 = reinterpret_cast<const float4 *>(&(cbuf ? ctl->col_a1 : ctl->col_a0)[j0]);
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
              } else if (X3) {
                gp[j] = __float_as_uint(gv.x);               // fp32: split into TF32 hi + lo when it is written out
                gp[j + 1] = __float_as_uint(gv.y);
                rowsum += gv.x + gv.y;
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
            const float rj = (cbuf ? ctl->col_r1 : ctl->col_r0)[j0 + j], aj = (cbuf ? ctl->col_a1 : ctl->col_a0)[j0 + j];
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
            } else if (X3) {
              gp[j] = __float_as_uint(gv);
              rowsum += gv;
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
        } else if (X3) {
          // G = G_hi + G_lo (two TF32 numbers): the same atom layout twice, G_lo 64 KiB behind G_hi
          uint8_t *atom = gbuf + (j0 >> 5) * P2_CHUNK + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            float hi4[4], lo4[4];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float gvv = __uint_as_float(gp[q4 * 4 + c4]);
              hi4[c4] = to_tf32(gvv);
              lo4[c4] = to_tf32(gvv - hi4[c4]);
            }
            *reinterpret_cast<uint4 *>(atom + ((q4 ^ (r & 7)) << 4)) =
                make_uint4(__float_as_uint(hi4[0]), __float_as_uint(hi4[1]), __float_as_uint(hi4[2]), __float_as_uint(hi4[3]));
            *reinterpret_cast<uint4 *>(atom + Q_G_BYTES + ((q4 ^ (r & 7)) << 4)) =
                make_uint4(__float_as_uint(lo4[0]), __float_as_uint(lo4[1]), __float_as_uint(lo4[2]), __float_as_uint(lo4[3]));
          }
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
      // (3xTF32: the partials lie over the column buffers, which the item's last group has finished with only after a
      //  barrier -- and which the next item must not refill before they have been summed)
      float (*part)[64] = Cfg::SHORT_CTRL ? reinterpret_cast<float (*)[64]>(ctl->col_r0) : ctl->part;
      if (Cfg::SHORT_CTRL) named_barrier_sync(1, SW_EPI_THREADS);
      part[jh * 4 + cq][r] = rowsum * gs_inv;
      named_barrier_sync(1, SW_EPI_THREADS);
      if (et < 64) {
        // 8 partials per row (2 lane halves x 4 column chunks) -> rowsum(G')_i of this item's columns, for apply_grad
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += part[k][et];
        // (only this panel's own rows: rows past the range end may belong to another panel with another split)
        if ((int)rank * 64 + et < it.rows_here)
          p.rowsum[(size_t)(it.ypass * SW_MAX_SPLIT + it.slab) * p.n_pad + it.row_base + (int)rank * 64 + et] = tot;
      }
      if (Cfg::SHORT_CTRL) named_barrier_sync(1, SW_EPI_THREADS);
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
      if (t == (p.ticket_total > 0 ? (unsigned)p.ticket_total : gridDim.x) - 1) {
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
  const int khalf = (kchunks / 4) * 2;                    // an S phase is issued as the K halves [0, khalf), [khalf, kchunks)

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
      auto load_S = [&](int g, int k_begin, int k_end) {
        const int jrow = (it.g_begin + g) * Q_GROUP + (int)rank * 128;
        for (int kc = k_begin; kc < k_end; kc += 2) {
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
      // the issue order of the MMA warp (see there): S of this pair's first two groups, then per group P(g) followed by
      // one half of an S phase two own groups ahead
      if (p.s_ahead) {
        if (pairidx < it.ng) load_S(pairidx, 0, kchunks);
        if (pairidx + 2 < it.ng) load_S(pairidx + 2, 0, kchunks);
        for (int g = 0; g < it.ng; ++g) {
          load_P(g);
          if (g + 4 < it.ng && owns(g + 4)) load_S(g + 4, 0, khalf);
          if (g >= 1 && g + 3 < it.ng && owns(g + 3)) load_S(g + 3, khalf, kchunks);
        }
      } else {
        if (owns(0)) load_S(0, 0, kchunks);
        for (int g = 0; g < it.ng; ++g) {
          if (g + 1 < it.ng && owns(g + 1)) load_S(g + 1, 0, kchunks);
          load_P(g);
        }
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
        // K chunks [k_begin, k_end) of the S phase number c of this pair (c: running index of its own groups)
        auto issue_S = [&](int c, int k_begin, int k_end) {
          const int b = c & 1;
          const uint32_t u = (uint32_t)(c >> 1);
          if (k_begin == 0) {
            mbar_wait_cluster(&ctl->s_empty[b], (u & 1) ^ 1);
            tc_fence_after();
          }
          const uint32_t d_tmem = tmem_s + b * 128;
          for (int kc = k_begin; kc < k_end; kc += 2) {
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
          if (k_end == kchunks) mma_commit_mask_elect(&ctl->s_full[b], pmask);
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
        // Order on the (in-order) tensor pipe: S of this pair's first two groups, then per group P(g) followed by HALF
        // of the S phase of an own group two own groups ahead.  Every refill of the single G buffer -- the epilogue's
        // write of G(g + 1) or the copy of the other pair's tile, both of which have to wait for P(g) to finish
        // reading -- then happens under half an S phase instead of stalling the pipe, and the epilogue math of a group
        // has two P phases and an S phase of slack.
        if (p.s_ahead) {
          if (pairidx < it.ng) issue_S(sc++, 0, kchunks);
          if (pairidx + 2 < it.ng) issue_S(sc++, 0, kchunks);
        } else if (owns(0)) {
          issue_S(sc++, 0, kchunks);
        }
        for (int g = 0; g < it.ng; ++g) {
          if (!p.s_ahead && g + 1 < it.ng && owns(g + 1)) issue_S(sc++, 0, kchunks);
          if (g == 0 && itn > 0) {                         // the previous item's dZ^T has been read out of TMEM
            mbar_wait_cluster(&ctl->dz_empty, (uint32_t)((itn - 1) & 1));
            tc_fence_after();
          }
          issue_P(g, gc + g);
          if (p.s_ahead) {
            if (g + 4 < it.ng && owns(g + 4)) issue_S(sc, 0, khalf);
            if (g >= 1 && g + 3 < it.ng && owns(g + 3)) issue_S(sc++, khalf, kchunks);
          }
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

      bool own_seen = false;                                // (short control block: one column buffer per CTA)
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
        const int cbuf = Cfg::SHORT_CTRL ? 0 : b;
        if (Cfg::SHORT_CTRL && own_seen) named_barrier_sync(1, SW_EPI_THREADS);   // (every warp has finished reading the buffer)
        own_seen = true;
        if (et < Q_GROUP) {
          (cbuf ? ctl->col_r1 : ctl->col_r0)[et] = (float)nxt_r;
          (cbuf ? ctl->col_a1 : ctl->col_a0)[et] = nxt_a;
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
        const float4 *cr4 = reinterpret_cast<const float4 *>(&(cbuf ? ctl->col_r1 : ctl->col_r0)[j0]);
        const float4 *ca4 = reinterpret_cast<const float4 *>(&(cbuf ? ctl->col_a1 : ctl->col_a0)[j0]);
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
            const float rj = (cbuf ? ctl->col_r1 : ctl->col_r0)[j0 + j], aj = (cbuf ? ctl->col_a1 : ctl->col_a0)[j0 + j];
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
      float (*part)[64] = Cfg::SHORT_CTRL ? reinterpret_cast<float (*)[64]>(ctl->col_r0) : ctl->part;
      if (Cfg::SHORT_CTRL) named_barrier_sync(1, SW_EPI_THREADS);
      part[jh * 4 + cq][r] = rowsum * gs_inv;
      named_barrier_sync(1, SW_EPI_THREADS);
      if (et < 64) {
        // 8 partials per row (2 lane halves x 4 column chunks) -> rowsum(G')_i of this item's columns, for apply_grad
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += part[k][et];
        // (only this panel's own rows: rows past the range end may belong to another panel with another split)
        // (each pair saw half of the column groups: the two partial sums are added in the slot zeroed by prep; two
        //  addends commute, so the result does not depend on the order)
        if ((int)rank * 64 + et < it.rows_here)
          atomicAdd(&p.rowsum[(size_t)(it.ypass * SW_MAX_SPLIT + it.slab) * p.n_pad + it.row_base + (int)rank * 64 + et],
                    tot);
      }
      if (Cfg::SHORT_CTRL) named_barrier_sync(1, SW_EPI_THREADS);
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
      if (t == (p.ticket_total > 0 ? (unsigned)p.ticket_total : gridDim.x) - 1) {
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
struct ApplyPlan {                // the part of a SweepPlan edrl_mmd_apply_grad needs to find a row's partial outputs
  int panels, full_items, split, pass_feats;
};
// one block row per output row (no per-element division), 128-bit accesses when d % 4 == 0
template <bool VEC4>
__global__ void __launch_bounds__(128)
mmd_apply_grad_kernel(const float *U, const float *__restrict__ zhi, const float *__restrict__ zlo,
                      const double *__restrict__ colsum_hi,
                      const float *__restrict__ stats, const float *__restrict__ grad_out, int row_begin, int row_count,
                      int row_begin2, int row_count2, int d, int d_pad, int n, int n_pad, ApplyPlan pa, ApplyPlan pb,
                      const float *__restrict__ rowsum, float *dz) {     // (U may be dz: edrl_mmd_backward)
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
  const int gpanel = (r < row_count) ? r / BM : (row_count + BM - 1) / BM + (r - row_count) / BM;
  // which launch swept this row's panel: the first pa.panels panels plan A (the quad kernel in a hybrid launch), the rest B
  const bool in_b = gpanel >= pa.panels;
  const int panel = in_b ? gpanel - pa.panels : gpanel;
  const int panels = in_b ? pb.panels : pa.panels, full_items = in_b ? pb.full_items : pa.full_items;
  const int split = in_b ? pb.split : pa.split, pass_feats = in_b ? pb.pass_feats : pa.pass_feats;
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
      float4 z = __ldg(reinterpret_cast<const float4 *>(zr + f));
      if (zlo != nullptr) {                                   // 3xTF32: the operand is hi + lo
        const float4 zl = __ldg(reinterpret_cast<const float4 *>(zlo + gr * d_pad + f));
        z.x += zl.x; z.y += zl.y; z.z += zl.z; z.w += zl.w;
      }
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
      const float zv = __ldg(zr + f) + (zlo != nullptr ? __ldg(zlo + gr * d_pad + f) : 0.f);
      dz[(size_t)r * d + f] = coef * (fmaf(fmaf(cv, fn, rs), zv, -cv * (float)colsum_hi[f]) + u);
    }
  }
}

