// mmd_bwd.cuh -- Part A: parameters of the gradient sweeps (textually included by mmd.cu inside namespace edrl::mmd)
// (The first-generation tile-recomputing backward kernels that lived here -- single-CTA 128 x 128 tiles, then a CTA-pair
//  kernel with 128-column S tiles -- are gone: every edrl_mmd_backward runs the fused sweep of mmd_sweep.cuh and
//  mmd_apply_grad_kernel in place.  DESIGN.md section 3 keeps their measurements.)
#pragma once

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

// ----------------------------------------------------------------------------- shared-memory units of the pair sweeps
// (mmd_sweep.cuh; the separate CTA-pair backward that first used them -- 128-column S tiles, 0.32 of the TF32 roofline --
//  is gone: edrl_mmd_backward runs the fused sweep and apply_grad in place)
constexpr int P2_STAGE = 16384;                 // ring stage per CTA
constexpr int P2_CHUNK = 8192;                  // 64 rows x 32 floats, 128-byte swizzle
constexpr int P2_FEATS = 512;                   // feature columns per pair and pass (2 M-tiles of 256)
