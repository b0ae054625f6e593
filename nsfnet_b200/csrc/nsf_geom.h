// Geometry of the packed parameter image and of the per-CTA gradient rows.
// Plain C++ (host + device + the host SIMT emulation used by tests/emu).
#pragma once
#include <stdint.h>

#ifndef NSF_HD
#if defined(__CUDACC__)
#define NSF_HD __host__ __device__ __forceinline__
#else
#define NSF_HD inline
#endif
#endif

#define NSF_MAX_HIDDEN 128
#define NSF_MAX_LAYERS 16
#define NSF_LOSS_SLOTS_I 16

// One FCNet (net.py:22-54): 2 -> H x L -> n_out.  HP = H rounded up to a multiple of 4 (the
// register tile of the FFMA kernels is 4 neurons x 4 points); padded neurons have zero weights
// and biases, so they carry exact zeros through every stream and every adjoint.
//
// Packed image `pk` (rebuilt from the flat state_dict-order buffer at the start of each call):
//   w0x[HP] w0y[HP] b0[HP]                                   layer 0 (K = 2: columns of W0)
//   for l = 1..L-1:  Wt_l[HP k][HP j]  W_l[HP j][HP k]  b_l[HP]   (forward / dgrad operand)
//   WL[4][HP] (rows >= n_out are zero)  bL[4]
// Gradient row `gs` (one per CTA, summed by the finalize kernel):
//   gw0x[HP] gw0y[HP] gb0[HP]  (gW_l[HP j][HP k] gb_l[HP])*  gWL[4][HP] gbL[4]  loss[16]
struct NsfNetGeom {
  int L, H, HP, n_out;
  int n_params;  // flat (unpadded) parameter count
  NSF_HD int pk_w0x() const { return 0; }
  NSF_HD int pk_w0y() const { return HP; }
  NSF_HD int pk_b0() const { return 2 * HP; }
  NSF_HD int pk_wt(int l) const { return 3 * HP + (l - 1) * (2 * HP * HP + HP); }
  NSF_HD int pk_w(int l) const { return pk_wt(l) + HP * HP; }
  NSF_HD int pk_b(int l) const { return pk_wt(l) + 2 * HP * HP; }
  NSF_HD int pk_wl() const { return 3 * HP + (L - 1) * (2 * HP * HP + HP); }
  NSF_HD int pk_bl() const { return pk_wl() + 4 * HP; }
  NSF_HD int pk_size() const { return pk_bl() + 4; }
  NSF_HD int gs_w0x() const { return 0; }
  NSF_HD int gs_w0y() const { return HP; }
  NSF_HD int gs_b0() const { return 2 * HP; }
  NSF_HD int gs_w(int l) const { return 3 * HP + (l - 1) * (HP * HP + HP); }
  NSF_HD int gs_b(int l) const { return gs_w(l) + HP * HP; }
  NSF_HD int gs_wl() const { return 3 * HP + (L - 1) * (HP * HP + HP); }
  NSF_HD int gs_bl() const { return gs_wl() + 4 * HP; }
  NSF_HD int gs_loss() const { return gs_bl() + 4; }
  NSF_HD int gs_row() const { return (gs_loss() + NSF_LOSS_SLOTS_I + 3) & ~3; }
};

static inline NsfNetGeom nsf_make_geom(int n_out, int L, int H) {
  NsfNetGeom g;
  g.L = L; g.H = H; g.HP = (H + 3) & ~3; g.n_out = n_out;
  g.n_params = 2 * H + H + (L - 1) * (H * H + H) + n_out * H + n_out;
  return g;
}

// flat (state_dict order) index -> index inside the gradient row / packed image without Wt.
// Returns the gs_* offset of flat parameter i.
static inline int nsf_flat_to_gs(const NsfNetGeom& g, int i) {
  const int H = g.H, HP = g.HP;
  if (i < 2 * H) { int j = i / 2, c = i % 2; return (c ? g.gs_w0y() : g.gs_w0x()) + j; }
  i -= 2 * H;
  if (i < H) return g.gs_b0() + i;
  i -= H;
  for (int l = 1; l < g.L; ++l) {
    if (i < H * H) { int j = i / H, k = i % H; return g.gs_w(l) + j * HP + k; }
    i -= H * H;
    if (i < H) return g.gs_b(l) + i;
    i -= H;
  }
  if (i < g.n_out * H) { int o = i / H, j = i % H; return g.gs_wl() + o * HP + j; }
  i -= g.n_out * H;
  return g.gs_bl() + i;
}

// modes of the step kernels
enum {
  NSF_MODE_FWD = 0,        // value forward only, write out[n][n_out]
  NSF_MODE_JET_RESID = 1,  // jet forward + residuals, no reverse
  NSF_MODE_JET_STEP = 2,   // jet forward + residuals + reverse
  NSF_MODE_MSE_STEP = 3,   // value forward + MSE-to-targets + reverse (boundary / supervised)
  NSF_MODE_BAR_STEP = 4    // value forward + given output adjoint + reverse (EVM net, ev :501-511)
};

struct NsfKernelArgs {
  NsfNetGeom g;
  const float* pk;  // packed parameter image
  const float* x;
  const float* y;
  long long n;
  int mode;
  int accumulate;   // 0: the first tile of every CTA overwrites its gradient row, 1: add to it
  // jet extras
  const float* e_in;    // [n] EVM output (has_evm) or null
  const float* vtm_in;  // [n] or null
  float* vtm_out;       // [n] or null
  const float* w;       // [n] or null
  float inv_Re, vis_t0, alpha_evm, cs1, cs2, k4, c_eq;
  int has_evm;
  float* resid_out;  // [4][n] or null
  float* vis_t_out;  // [n] or null
  float* ebar_out;   // [n] or null: d(loss)/d(e) = -g4
  // mse block / given bar
  const float* tu; const float* tv; const float* tp;
  float cu, cv, cp;
  int loss_slot;
  const float* bar_in;  // [n]
  // outputs / workspace
  float* out;              // mode FWD: [n][n_out]
  float* stash;            // per-CTA activation stash [grid][stash_stride]
  long long stash_stride;
  float* scratch;          // per-CTA gradient rows [grid][gs_row]
  int n_tiles;
};
