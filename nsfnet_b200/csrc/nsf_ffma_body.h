// FP32 (FFMA) jet kernels: one CTA works on tiles of PT collocation points, a thread owns a
// register tile of 4 neurons x 4 points x NS jet streams (NS = 4: value, d/dx, d/dy, laplacian;
// NS = 1: value only).  The body is written as barrier-separated *phases* over per-thread state
// so that the very same source runs (a) as the CUDA kernel in nsf_ffma.cu and (b) under the host
// SIMT emulation of tests/emu (NSF_EMU), which is how the index logic is checked without a GPU.
//
// Math (reference ev-NSFnet/pinn_solver.py:290-342,372-428 + loss.backward()):
//   streams s in {0, x, y, lap}; the reference needs u_xx and u_yy only as their sum
//   (pinn_solver.py:337-338), so the two second-derivative streams are carried as one laplacian
//   stream:  a_lap = d2 (zx^2 + zy^2) + d1 z_lap.
//   tanh:  t = tanh z0, d1 = 1 - t^2, d2 = -2 t d1, d3 = -2 d1 (1 - 3 t^2)
//   reverse:  zb_lap = ab_lap d1;  zb_x = ab_x d1 + 2 ab_lap d2 zx;  zb_y likewise;
//             zb_0 = ab_0 d1 + ab_x d2 zx + ab_y d2 zy + ab_lap (d3 (zx^2+zy^2) + d2 z_lap)
#pragma once
#include "nsf_geom.h"

#ifdef NSF_EMU
#include <cmath>
#include <cstring>
#define NSF_DEV inline
struct nsf_f4 { float x, y, z, w; };
static inline nsf_f4 nsf_ld4(const float* p) { nsf_f4 v; std::memcpy(&v, p, 16); return v; }
static inline nsf_f4 nsf_ldg4(const float* p) { return nsf_ld4(p); }
static inline nsf_f4 nsf_ldcs4(const float* p) { return nsf_ld4(p); }
static inline void nsf_st4(float* p, nsf_f4 v) { std::memcpy(p, &v, 16); }
static inline void nsf_stcs4(float* p, nsf_f4 v) { nsf_st4(p, v); }
static inline float nsf_ldg(const float* p) { return *p; }
static inline void nsf_cp16(float* dst_smem, const float* src) { std::memcpy(dst_smem, src, 16); }
static inline void nsf_cp_wait() {}
static inline float nsf_tanh(float x) { return std::tanh(x); }
static inline float nsf_fma(float a, float b, float c) { return std::fma(a, b, c); }
#else
#include "nsf_math.cuh"
#define NSF_DEV __device__ __forceinline__
typedef float4 nsf_f4;
NSF_DEV nsf_f4 nsf_ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
NSF_DEV nsf_f4 nsf_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
NSF_DEV nsf_f4 nsf_ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
NSF_DEV void nsf_st4(float* p, nsf_f4 v) { *reinterpret_cast<float4*>(p) = v; }
NSF_DEV void nsf_stcs4(float* p, nsf_f4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
NSF_DEV float nsf_ldg(const float* p) { return __ldg(p); }
// 16 bytes global -> shared without a register round trip (LDGSTS); nsf_cp_wait: this thread's copies have landed
NSF_DEV void nsf_cp16(float* dst_smem, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
NSF_DEV void nsf_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
NSF_DEV float nsf_tanh(float x) { return nsf_tanh_fast(x); }
NSF_DEV float nsf_fma(float a, float b, float c) { return fmaf(a, b, c); }
#endif

NSF_DEV float& nsf_f4at(nsf_f4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
NSF_DEV float nsf_f4get(const nsf_f4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

template <int NS>
struct NsfRegs {
  float acc[NS][4][4];  // [stream][neuron ji][point pi]
  float loss[6];        // per-thread partial sums over all tiles of this CTA (threads < PT)
};

// Uniform (per CTA) context of one tile.
struct NsfTile {
  const NsfKernelArgs* a;
  float* act;    // smem [NS][HP][PT]  layer input streams (swizzled float4 columns)
  float* zb;     // smem [NS][HP][PT]  pre-activation adjoints
  float* xs;     // smem [PT]
  float* ys;     // smem [PT]
  float* ov;     // smem [NS][4][PT]  network outputs, then their adjoints
  float* red;    // smem [PT][6]
  float* wsm;    // smem [HP][HP]  (NS = 1) the weights of the next contraction, staged by nsf_ph_wstage
  float* stash;  // global, this CTA: [L][NS][HP][PT]
  float* grow;   // global, this CTA's gradient row
  long long p0;  // first point of the tile
  int nvalid;    // valid points in the tile (<= PT)
  int first;     // first tile of this CTA and !accumulate: overwrite the gradient row
  int NT;        // threads per CTA = (HP/4) * (PT/4)
};

// float offset of the float4 holding points 4*c4..4*c4+3 of row `row`, stream s.  The float4
// column is XOR-swizzled with the row's 4-neuron group so that the weight-gradient phase (lanes on
// different rows, same point group) is bank-conflict free.
template <int PT>
NSF_DEV int nsf_aoff(int HP, int s, int row, int c4) {
  constexpr int NC = PT / 4;
  return (s * HP + row) * PT + ((c4 ^ ((row >> 2) & (NC - 1))) << 2);
}

NSF_DEV void nsf_gadd(const NsfTile& c, int idx, float v) {
  c.grow[idx] = c.first ? v : c.grow[idx] + v;
}

// ---- phase: load the tile's points ------------------------------------------------------
template <int NS, int PT>
NSF_DEV void nsf_ph_load(const NsfTile& c, int tid) {
  for (int p = tid; p < PT; p += c.NT) {
    const bool ok = p < c.nvalid;
    c.xs[p] = ok ? nsf_ldg(c.a->x + c.p0 + p) : 0.f;
    c.ys[p] = ok ? nsf_ldg(c.a->y + c.p0 + p) : 0.f;
  }
}

// ---- phase: layer 0 (K = 2) into the accumulators ------------------------------------------
template <int NS, int PT>
NSF_DEV void nsf_ph_layer0(const NsfTile& c, NsfRegs<NS>& r, int tid) {
  constexpr int NPG = PT / 4;
  const int jg = tid / NPG, pg = tid % NPG;
  const NsfNetGeom& g = c.a->g;
  const float* pk = c.a->pk;
  float xv[4], yv[4];
#pragma unroll
  for (int pi = 0; pi < 4; ++pi) { xv[pi] = c.xs[pg * 4 + pi]; yv[pi] = c.ys[pg * 4 + pi]; }
#pragma unroll
  for (int ji = 0; ji < 4; ++ji) {
    const int j = jg * 4 + ji;
    const float wx = nsf_ldg(pk + g.pk_w0x() + j), wy = nsf_ldg(pk + g.pk_w0y() + j), b = nsf_ldg(pk + g.pk_b0() + j);
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      r.acc[0][ji][pi] = nsf_fma(wx, xv[pi], nsf_fma(wy, yv[pi], b));
      if constexpr (NS == 4) { r.acc[1][ji][pi] = wx; r.acc[2][ji][pi] = wy; r.acc[3][ji][pi] = 0.f; }
    }
  }
}

// ---- phase: tanh jet on the accumulators -> act (in place in smem), optional stash --------------
template <int NS, int PT>
NSF_DEV void nsf_ph_act(const NsfTile& c, NsfRegs<NS>& r, int tid, int l, bool do_stash) {
  constexpr int NPG = PT / 4;
  const int jg = tid / NPG, pg = tid % NPG;
  const int HP = c.a->g.HP;
#pragma unroll
  for (int ji = 0; ji < 4; ++ji) {
    const int j = jg * 4 + ji;
    nsf_f4 T, ZX, ZY, ZL, AX, AY, AL;
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      const float t = nsf_tanh(r.acc[0][ji][pi]);
      nsf_f4at(T, pi) = t;
      if constexpr (NS == 4) {
        const float d1 = nsf_fma(-t, t, 1.f);
        const float d2 = -2.f * t * d1;
        const float zx = r.acc[1][ji][pi], zy = r.acc[2][ji][pi], zl = r.acc[3][ji][pi];
        nsf_f4at(ZX, pi) = zx; nsf_f4at(ZY, pi) = zy; nsf_f4at(ZL, pi) = zl;
        nsf_f4at(AX, pi) = d1 * zx;
        nsf_f4at(AY, pi) = d1 * zy;
        nsf_f4at(AL, pi) = nsf_fma(d2, nsf_fma(zx, zx, zy * zy), d1 * zl);
      }
    }
    nsf_st4(c.act + nsf_aoff<PT>(HP, 0, j, pg), T);
    if constexpr (NS == 4) {
      nsf_st4(c.act + nsf_aoff<PT>(HP, 1, j, pg), AX);
      nsf_st4(c.act + nsf_aoff<PT>(HP, 2, j, pg), AY);
      nsf_st4(c.act + nsf_aoff<PT>(HP, 3, j, pg), AL);
    }
    if (do_stash) {
      float* sp = c.stash + ((long long)(l * NS) * HP + j) * PT + pg * 4;
      nsf_stcs4(sp, T);
      if constexpr (NS == 4) {
        nsf_stcs4(sp + (long long)1 * HP * PT, ZX);
        nsf_stcs4(sp + (long long)2 * HP * PT, ZY);
        nsf_stcs4(sp + (long long)3 * HP * PT, ZL);
      }
    }
  }
}

// ---- phase (NS = 1): the HP x HP weights of the next contraction, global -> shared memory ---------
// The value-stream launches are small (a boundary block is 2052 points: ONE tile per CTA), so a CTA reads every weight exactly
// once and the contraction loop below, fed through L1, ran at one L2 round trip per two k (170 us for the block, 11 % of a
// 120 000-point training iteration).  Here all of a layer's loads are in flight at once (asynchronous copies, no registers),
// beside the other work of the phase; nsf_cp_wait() before the phase's barrier.
template <int NS, int PT>
NSF_DEV void nsf_ph_wstage(const NsfTile& c, int tid, const float* W) {
  const int HP = c.a->g.HP, n4 = HP * HP / 4;
  for (int i = tid; i < n4; i += c.NT) nsf_cp16(c.wsm + 4 * i, W + 4 * i);
}

// ---- phase: acc[s][ji][pi] = sum_k src[s][k][p] * W[k][j]  (+ bias on stream 0) -----------------
// W is row-major [k][HP] (Wt_l for the forward pass, W_l for dgrad): NS = 4 reads it through L1, NS = 1 from the staged copy.
template <int NS, int PT>
NSF_DEV void nsf_ph_gemm(const NsfTile& c, NsfRegs<NS>& r, int tid, const float* src, const float* W, const float* bias) {
  constexpr int NPG = PT / 4;
  const int jg = tid / NPG, pg = tid % NPG;
  const int HP = c.a->g.HP;
#pragma unroll
  for (int ji = 0; ji < 4; ++ji) {
    const float b = bias ? nsf_ldg(bias + jg * 4 + ji) : 0.f;
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      r.acc[0][ji][pi] = b;
#pragma unroll
      for (int s = 1; s < NS; ++s) r.acc[s][ji][pi] = 0.f;
    }
  }
  const float* wp = (NS == 1 ? c.wsm : W) + jg * 4;
#pragma unroll 2
  for (int k = 0; k < HP; ++k) {
    const nsf_f4 w = NS == 1 ? nsf_ld4(wp + k * HP) : nsf_ldg4(wp + (long long)k * HP);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const nsf_f4 av = nsf_ld4(src + nsf_aoff<PT>(HP, s, k, pg));
      r.acc[s][0][0] = nsf_fma(w.x, av.x, r.acc[s][0][0]); r.acc[s][0][1] = nsf_fma(w.x, av.y, r.acc[s][0][1]);
      r.acc[s][0][2] = nsf_fma(w.x, av.z, r.acc[s][0][2]); r.acc[s][0][3] = nsf_fma(w.x, av.w, r.acc[s][0][3]);
      r.acc[s][1][0] = nsf_fma(w.y, av.x, r.acc[s][1][0]); r.acc[s][1][1] = nsf_fma(w.y, av.y, r.acc[s][1][1]);
      r.acc[s][1][2] = nsf_fma(w.y, av.z, r.acc[s][1][2]); r.acc[s][1][3] = nsf_fma(w.y, av.w, r.acc[s][1][3]);
      r.acc[s][2][0] = nsf_fma(w.z, av.x, r.acc[s][2][0]); r.acc[s][2][1] = nsf_fma(w.z, av.y, r.acc[s][2][1]);
      r.acc[s][2][2] = nsf_fma(w.z, av.z, r.acc[s][2][2]); r.acc[s][2][3] = nsf_fma(w.z, av.w, r.acc[s][2][3]);
      r.acc[s][3][0] = nsf_fma(w.w, av.x, r.acc[s][3][0]); r.acc[s][3][1] = nsf_fma(w.w, av.y, r.acc[s][3][1]);
      r.acc[s][3][2] = nsf_fma(w.w, av.z, r.acc[s][3][2]); r.acc[s][3][3] = nsf_fma(w.w, av.w, r.acc[s][3][3]);
    }
  }
}

// ---- phase: output layer  ov[s][o][p] = sum_j act[s][j][p] WL[o][j] (+ bL on stream 0) -----------
template <int NS, int PT>
NSF_DEV void nsf_ph_out(const NsfTile& c, int tid) {
  const NsfNetGeom& g = c.a->g;
  const int HP = g.HP;
  const float* WL = c.a->pk + g.pk_wl();
  const float* bL = c.a->pk + g.pk_bl();
  for (int i = tid; i < PT * NS; i += c.NT) {
    const int p = i % PT, s = i / PT;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    for (int j = 0; j < HP; ++j) {
      const float a = c.act[nsf_aoff<PT>(HP, s, j, p >> 2) + (p & 3)];
      o0 = nsf_fma(a, nsf_ldg(WL + j), o0);
      o1 = nsf_fma(a, nsf_ldg(WL + HP + j), o1);
      o2 = nsf_fma(a, nsf_ldg(WL + 2 * HP + j), o2);
    }
    if (s == 0) { o0 += nsf_ldg(bL + 0); o1 += nsf_ldg(bL + 1); o2 += nsf_ldg(bL + 2); }
    c.ov[(s * 4 + 0) * PT + p] = o0;
    c.ov[(s * 4 + 1) * PT + p] = o1;
    c.ov[(s * 4 + 2) * PT + p] = o2;
    c.ov[(s * 4 + 3) * PT + p] = 0.f;
  }
}

// ---- phase: value outputs to global (mode FWD) -----------------------------------------------
template <int NS, int PT>
NSF_DEV void nsf_ph_store_out(const NsfTile& c, int tid) {
  const int no = c.a->g.n_out;
  for (int i = tid; i < PT * no; i += c.NT) {
    const int p = i / no, o = i % no;
    if (p < c.nvalid) c.a->out[(c.p0 + p) * no + o] = c.ov[o * PT + p];
  }
}

// ---- phase: NS residuals, loss partial sums, output adjoints (ev :311-342,387-397) ---------------
template <int NS, int PT>
NSF_DEV void nsf_ph_resid(const NsfTile& c, NsfRegs<NS>& r, int tid, bool want_bar) {
  const NsfKernelArgs& a = *c.a;
  for (int p = tid; p < PT; p += c.NT) {
    const bool ok = p < c.nvalid;
    const long long gp = c.p0 + p;
    float* o = c.ov + p;
#define NSF_OV(s, k) o[((s) * 4 + (k)) * PT]
    const float u = NSF_OV(0, 0), v = NSF_OV(0, 1);
    const float ux = a.cs1 * NSF_OV(1, 0), vx = a.cs1 * NSF_OV(1, 1), px = a.cs1 * NSF_OV(1, 2);
    const float uy = a.cs1 * NSF_OV(2, 0), vy = a.cs1 * NSF_OV(2, 1), py = a.cs1 * NSF_OV(2, 2);
    const float ul = a.cs2 * NSF_OV(3, 0), vl = a.cs2 * NSF_OV(3, 1);
    float e = 0.f, vis = 0.f;
    if (a.has_evm) {
      e = ok ? nsf_ldg(a.e_in + gp) : 0.f;
      vis = a.vis_t0;
      if (a.vtm_in && ok) vis = fminf(a.vis_t0, nsf_ldg(a.vtm_in + gp));
    }
    const float nu = a.inv_Re + vis;
    const float eq1 = (u * ux + v * uy) + px - nu * ul;
    const float eq2 = (u * vx + v * vy) + py - nu * vl;
    const float eq3 = ux + vy;
    const float eq4 = a.has_evm ? (eq1 * (u - 0.5f) + eq2 * (v - 0.5f)) - e : 0.f;
    const float w = (a.w && ok) ? nsf_ldg(a.w + gp) : 1.f;
    if (ok) {
      r.loss[0] += w * eq1 * eq1; r.loss[1] += w * eq2 * eq2; r.loss[2] += w * eq3 * eq3; r.loss[3] += w * eq4 * eq4;
      r.loss[4] += vis; r.loss[5] += 1.f;
      if (a.resid_out) {
        a.resid_out[gp] = eq1; a.resid_out[a.n + gp] = eq2; a.resid_out[2 * a.n + gp] = eq3; a.resid_out[3 * a.n + gp] = eq4;
      }
      if (a.vis_t_out) a.vis_t_out[gp] = vis;
      if (a.has_evm && a.vtm_out) a.vtm_out[gp] = a.alpha_evm * fabsf(e);
    }
    if (want_bar) {
      const float cw = ok ? a.c_eq * w : 0.f;
      const float g1 = cw * (2.f * eq1 + a.k4 * eq4 * (u - 0.5f));
      const float g2 = cw * (2.f * eq2 + a.k4 * eq4 * (v - 0.5f));
      const float g3 = 2.f * cw * eq3;
      const float g4 = a.k4 * cw * eq4;
      NSF_OV(0, 0) = g1 * ux + g2 * vx + g4 * eq1;
      NSF_OV(0, 1) = g1 * uy + g2 * vy + g4 * eq2;
      NSF_OV(0, 2) = 0.f;
      NSF_OV(1, 0) = a.cs1 * (g1 * u + g3); NSF_OV(1, 1) = a.cs1 * (g2 * u); NSF_OV(1, 2) = a.cs1 * g1;
      NSF_OV(2, 0) = a.cs1 * (g1 * v); NSF_OV(2, 1) = a.cs1 * (g2 * v + g3); NSF_OV(2, 2) = a.cs1 * g2;
      NSF_OV(3, 0) = -a.cs2 * nu * g1; NSF_OV(3, 1) = -a.cs2 * nu * g2; NSF_OV(3, 2) = 0.f;
      if (a.ebar_out && ok) a.ebar_out[gp] = -g4;
    }
#undef NSF_OV
  }
}

// ---- phase: value-stream MSE against targets (ev :374-379, :399-411) or a given output adjoint ---
template <int NS, int PT>
NSF_DEV void nsf_ph_mse(const NsfTile& c, NsfRegs<NS>& r, int tid) {
  const NsfKernelArgs& a = *c.a;
  for (int p = tid; p < PT; p += c.NT) {
    const bool ok = p < c.nvalid;
    const long long gp = c.p0 + p;
    float* o = c.ov + p;
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (ok) {
      if (a.mode == NSF_MODE_BAR_STEP) {
        b0 = nsf_ldg(a.bar_in + gp);
      } else {
        const float du = o[0] - nsf_ldg(a.tu + gp), dv = o[PT] - nsf_ldg(a.tv + gp);
        r.loss[0] += du * du; r.loss[1] += dv * dv;
        b0 = 2.f * a.cu * du; b1 = 2.f * a.cv * dv;
        if (a.tp) {
          const float tp = nsf_ldg(a.tp + gp);
          if (tp - tp == 0.f) {  // finite (NaN / Inf targets are masked out, ev :404-409)
            const float dp = o[2 * PT] - tp;
            r.loss[2] += dp * dp; r.loss[3] += 1.f;
            b2 = 2.f * a.cp * dp;
          }
        }
      }
    }
    o[0] = b0; o[PT] = b1; o[2 * PT] = b2; o[3 * PT] = 0.f;
  }
}

// ---- phase: reverse of the output layer -----------------------------------------------------
//  (a) acc[s][ji][pi] = sum_o ov[s][o][p] WL[o][j]      (adjoint of the last hidden activation)
//  (b) gWL[o][j] += sum_{s,p} ov[s][o][p] act[s][j][p],  gbL[o] += sum_p ov[0][o][p]
template <int NS, int PT>
NSF_DEV void nsf_ph_outbwd(const NsfTile& c, NsfRegs<NS>& r, int tid) {
  constexpr int NPG = PT / 4;
  const int jg = tid / NPG, pg = tid % NPG;
  const NsfNetGeom& g = c.a->g;
  const int HP = g.HP;
  const float* WL = c.a->pk + g.pk_wl();
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const nsf_f4 b0 = nsf_ld4(c.ov + (s * 4 + 0) * PT + pg * 4);
    const nsf_f4 b1 = nsf_ld4(c.ov + (s * 4 + 1) * PT + pg * 4);
    const nsf_f4 b2 = nsf_ld4(c.ov + (s * 4 + 2) * PT + pg * 4);
#pragma unroll
    for (int ji = 0; ji < 4; ++ji) {
      const int j = jg * 4 + ji;
      const float w0 = nsf_ldg(WL + j), w1 = nsf_ldg(WL + HP + j), w2 = nsf_ldg(WL + 2 * HP + j);
      r.acc[s][ji][0] = nsf_fma(b0.x, w0, nsf_fma(b1.x, w1, b2.x * w2));
      r.acc[s][ji][1] = nsf_fma(b0.y, w0, nsf_fma(b1.y, w1, b2.y * w2));
      r.acc[s][ji][2] = nsf_fma(b0.z, w0, nsf_fma(b1.z, w1, b2.z * w2));
      r.acc[s][ji][3] = nsf_fma(b0.w, w0, nsf_fma(b1.w, w1, b2.w * w2));
    }
  }
  for (int i = tid; i < 3 * HP; i += c.NT) {
    const int o = i / HP, j = i % HP;
    float sum = 0.f;
    for (int s = 0; s < NS; ++s)
      for (int c4 = 0; c4 < NPG; ++c4) {
        const nsf_f4 av = nsf_ld4(c.act + nsf_aoff<PT>(HP, s, j, c4));
        const nsf_f4 bv = nsf_ld4(c.ov + (s * 4 + o) * PT + c4 * 4);
        sum = nsf_fma(av.x, bv.x, nsf_fma(av.y, bv.y, nsf_fma(av.z, bv.z, nsf_fma(av.w, bv.w, sum))));
      }
    nsf_gadd(c, g.gs_wl() + o * HP + j, sum);
  }
  for (int o = tid; o < 4; o += c.NT) {
    float sum = 0.f;
    for (int p = 0; p < PT; ++p) sum += c.ov[o * PT + p];
    nsf_gadd(c, g.gs_bl() + o, sum);
  }
}

// ---- phase: adjoint through tanh of layer l (acc holds ab_s) -> zb; rebuild layer l's input
//      streams a^{l-1} from the stash of layer l-1 into act --------------------------------------
template <int NS, int PT>
NSF_DEV void nsf_ph_zbar(const NsfTile& c, NsfRegs<NS>& r, int tid, int l) {
  constexpr int NPG = PT / 4;
  const int jg = tid / NPG, pg = tid % NPG;
  const int HP = c.a->g.HP;
  const long long QS = (long long)HP * PT;
#pragma unroll
  for (int ji = 0; ji < 4; ++ji) {
    const int j = jg * 4 + ji;
    const float* sp = c.stash + ((long long)(l * NS) * HP + j) * PT + pg * 4;
    const nsf_f4 T = nsf_ldcs4(sp);
    nsf_f4 ZX, ZY, ZL, B0, BX, BY, BL;
    if constexpr (NS == 4) { ZX = nsf_ldcs4(sp + QS); ZY = nsf_ldcs4(sp + 2 * QS); ZL = nsf_ldcs4(sp + 3 * QS); }
#pragma unroll
    for (int pi = 0; pi < 4; ++pi) {
      const float t = nsf_f4get(T, pi);
      const float d1 = nsf_fma(-t, t, 1.f);
      if constexpr (NS == 4) {
        const float d2 = -2.f * t * d1;
        const float d3 = -2.f * d1 * nsf_fma(-3.f * t, t, 1.f);
        const float zx = nsf_f4get(ZX, pi), zy = nsf_f4get(ZY, pi), zl = nsf_f4get(ZL, pi);
        const float a0 = r.acc[0][ji][pi], ax = r.acc[1][ji][pi], ay = r.acc[2][ji][pi], al = r.acc[3][ji][pi];
        nsf_f4at(BL, pi) = al * d1;
        nsf_f4at(BX, pi) = nsf_fma(ax, d1, 2.f * al * d2 * zx);
        nsf_f4at(BY, pi) = nsf_fma(ay, d1, 2.f * al * d2 * zy);
        const float q = nsf_fma(zx, zx, zy * zy);
        nsf_f4at(B0, pi) = nsf_fma(a0, d1, nsf_fma(ax * d2, zx, nsf_fma(ay * d2, zy, al * nsf_fma(d3, q, d2 * zl))));
      } else {
        nsf_f4at(B0, pi) = r.acc[0][ji][pi] * d1;
      }
    }
    nsf_st4(c.zb + nsf_aoff<PT>(HP, 0, j, pg), B0);
    if constexpr (NS == 4) {
      nsf_st4(c.zb + nsf_aoff<PT>(HP, 1, j, pg), BX);
      nsf_st4(c.zb + nsf_aoff<PT>(HP, 2, j, pg), BY);
      nsf_st4(c.zb + nsf_aoff<PT>(HP, 3, j, pg), BL);
    }
    if (l >= 1) {  // a^{l-1} for the weight gradient of layer l
      const float* sq = c.stash + ((long long)((l - 1) * NS) * HP + j) * PT + pg * 4;
      const nsf_f4 T1 = nsf_ldcs4(sq);
      nsf_st4(c.act + nsf_aoff<PT>(HP, 0, j, pg), T1);
      if constexpr (NS == 4) {
        const nsf_f4 X1 = nsf_ldcs4(sq + QS), Y1 = nsf_ldcs4(sq + 2 * QS), L1 = nsf_ldcs4(sq + 3 * QS);
        nsf_f4 AX, AY, AL;
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) {
          const float t = nsf_f4get(T1, pi), zx = nsf_f4get(X1, pi), zy = nsf_f4get(Y1, pi), zl = nsf_f4get(L1, pi);
          const float d1 = nsf_fma(-t, t, 1.f);
          const float d2 = -2.f * t * d1;
          nsf_f4at(AX, pi) = d1 * zx;
          nsf_f4at(AY, pi) = d1 * zy;
          nsf_f4at(AL, pi) = nsf_fma(d2, nsf_fma(zx, zx, zy * zy), d1 * zl);
        }
        nsf_st4(c.act + nsf_aoff<PT>(HP, 1, j, pg), AX);
        nsf_st4(c.act + nsf_aoff<PT>(HP, 2, j, pg), AY);
        nsf_st4(c.act + nsf_aoff<PT>(HP, 3, j, pg), AL);
      }
    }
  }
}

// ---- phase: weight / bias gradient of hidden layer l from zb (adjoints) and act (inputs) --------
//   gW_l[j][k] += sum_{s,p} zb[s][j][p] act[s][k][p];  gb_l[j] += sum_p zb[0][j][p]
// 4x4 (j,k) register tiles; a warp covers a 4 (j-tiles) x 8 (k-tiles) block.
template <int NS, int PT>
NSF_DEV void nsf_ph_wgrad(const NsfTile& c, int tid, int l) {
  constexpr int NPG = PT / 4;
  const NsfNetGeom& g = c.a->g;
  const int HP = g.HP, NJG = HP / 4;
  const int TJB = (NJG + 3) / 4, TKB = (NJG + 7) / 8;
  const int gw = g.gs_w(l);
  for (int sidx = tid; sidx < TJB * TKB * 32; sidx += c.NT) {
    const int blk = sidx >> 5, lane = sidx & 31;
    const int tj = (blk / TKB) * 4 + (lane >> 3), tk = (blk % TKB) * 8 + (lane & 7);
    if (tj >= NJG || tk >= NJG) continue;
    float d[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) d[a][b] = 0.f;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll 2
      for (int c4 = 0; c4 < NPG; ++c4) {
        nsf_f4 zv[4], av[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) zv[a] = nsf_ld4(c.zb + nsf_aoff<PT>(HP, s, tj * 4 + a, c4));
#pragma unroll
        for (int b = 0; b < 4; ++b) av[b] = nsf_ld4(c.act + nsf_aoff<PT>(HP, s, tk * 4 + b, c4));
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b)
            d[a][b] = nsf_fma(zv[a].x, av[b].x, nsf_fma(zv[a].y, av[b].y, nsf_fma(zv[a].z, av[b].z, nsf_fma(zv[a].w, av[b].w, d[a][b]))));
      }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float* gp = c.grow + gw + (tj * 4 + a) * HP + tk * 4;
      nsf_f4 v;
      if (c.first) { v.x = d[a][0]; v.y = d[a][1]; v.z = d[a][2]; v.w = d[a][3]; }
      else { v = nsf_ld4(gp); v.x += d[a][0]; v.y += d[a][1]; v.z += d[a][2]; v.w += d[a][3]; }
      nsf_st4(gp, v);
    }
  }
  for (int j = tid; j < HP; j += c.NT) {
    float sum = 0.f;
    for (int c4 = 0; c4 < NPG; ++c4) {
      const nsf_f4 v = nsf_ld4(c.zb + nsf_aoff<PT>(HP, 0, j, c4));
      sum += (v.x + v.y) + (v.z + v.w);
    }
    nsf_gadd(c, g.gs_b(l) + j, sum);
  }
}

// ---- phase: gradient of layer 0 (inputs are x, y and the constant unit tangents) ----------------
template <int NS, int PT>
NSF_DEV void nsf_ph_l0bwd(const NsfTile& c, int tid) {
  const NsfNetGeom& g = c.a->g;
  const int HP = g.HP;
  for (int j = tid; j < HP; j += c.NT) {
    float sx = 0.f, sy = 0.f, sb = 0.f;
    for (int p = 0; p < PT; ++p) {
      const int o = (p & 3);
      const float z0 = c.zb[nsf_aoff<PT>(HP, 0, j, p >> 2) + o];
      sb += z0;
      sx = nsf_fma(z0, c.xs[p], sx);
      sy = nsf_fma(z0, c.ys[p], sy);
      if constexpr (NS == 4) {
        sx += c.zb[nsf_aoff<PT>(HP, 1, j, p >> 2) + o];
        sy += c.zb[nsf_aoff<PT>(HP, 2, j, p >> 2) + o];
      }
    }
    nsf_gadd(c, g.gs_w0x() + j, sx);
    nsf_gadd(c, g.gs_w0y() + j, sy);
    nsf_gadd(c, g.gs_b0() + j, sb);
  }
}

// ---- CTA epilogue: per-thread loss partials -> the gradient row's loss slots ---------------------
template <int NS, int PT>
NSF_DEV void nsf_ph_loss_a(const NsfTile& c, NsfRegs<NS>& r, int tid) {
  if (tid < PT)
    for (int k = 0; k < 6; ++k) c.red[tid * 6 + k] = r.loss[k];
}
template <int NS, int PT>
NSF_DEV void nsf_ph_loss_b(const NsfTile& c, int tid, bool accumulate) {
  const NsfKernelArgs& a = *c.a;
  const int base = a.g.gs_loss();
  if (tid < NSF_LOSS_SLOTS_I) {
    float v = 0.f;
    const int npt = PT < c.NT ? PT : c.NT;
    int src = -1;
    if (a.mode == NSF_MODE_JET_STEP || a.mode == NSF_MODE_JET_RESID) { if (tid < 6) src = tid; }
    else if (a.mode == NSF_MODE_MSE_STEP) { if (tid >= a.loss_slot && tid < a.loss_slot + 4) src = tid - a.loss_slot; }
    if (src >= 0)
      for (int p = 0; p < npt; ++p) v += c.red[p * 6 + src];
    c.grow[base + tid] = accumulate ? c.grow[base + tid] + v : v;
  }
}

// =================================================================================================
// The CTA program.  NSF_PHASE(stmt...) runs the statements for every thread of the CTA (variables
// `tid` and `r`) and ends with a CTA barrier.  On the GPU the per-thread state is a local object
// and the barrier is __syncthreads(); under NSF_EMU the threads are a host loop over an array of
// per-thread states.
// =================================================================================================
#ifdef NSF_EMU
#define NSF_REGS_PARAM(NS) , NsfRegs<NS>* regs_
#define NSF_REGS_DECL(NS)
extern int nsf_emu_reverse;  // run the threads of every phase in reverse order (race check)
#define NSF_PHASE(...)                                                  \
  for (int t_ = 0; t_ < nthreads; ++t_) {                               \
    const int tid = nsf_emu_reverse ? nthreads - 1 - t_ : t_;           \
    NsfRegs<NS>& r = regs_[tid];                                        \
    (void)r;                                                            \
    __VA_ARGS__;                                                        \
  }
#else
#define NSF_REGS_PARAM(NS)
#define NSF_REGS_DECL(NS) NsfRegs<NS> regs_;
#define NSF_PHASE(...)                                                  \
  {                                                                     \
    const int tid = threadIdx.x;                                        \
    NsfRegs<NS>& r = regs_;                                             \
    (void)r;                                                            \
    __VA_ARGS__;                                                        \
  }                                                                     \
  __syncthreads();
#endif

template <int NS, int PT>
NSF_DEV void nsf_cta_program(const NsfKernelArgs& a, float* smem, int bid, int nblocks, int nthreads NSF_REGS_PARAM(NS)) {
  NSF_REGS_DECL(NS)
  const NsfNetGeom& g = a.g;
  const int mode = a.mode;
  const bool train = mode >= NSF_MODE_JET_STEP;
  const bool jet = mode == NSF_MODE_JET_RESID || mode == NSF_MODE_JET_STEP;
  NsfTile c;
  c.a = &a;
  c.act = smem;
  c.zb = smem + NS * g.HP * PT;
  c.xs = smem + 2 * NS * g.HP * PT;
  c.ys = c.xs + PT;
  c.ov = c.ys + PT;
  c.red = c.ov + NS * 4 * PT;
  c.wsm = c.red + PT * 6;
  c.stash = a.stash ? a.stash + (long long)bid * a.stash_stride : (float*)0;
  c.grow = a.scratch ? a.scratch + (long long)bid * g.gs_row() : (float*)0;
  c.NT = nthreads;
  c.first = a.accumulate ? 0 : 1;
  c.p0 = 0;
  c.nvalid = 0;
  NSF_PHASE({ for (int k = 0; k < 6; ++k) r.loss[k] = 0.f; })
  for (int tile = bid; tile < a.n_tiles; tile += nblocks) {
    c.p0 = (long long)tile * PT;
    c.nvalid = (int)((a.n - c.p0) < PT ? (a.n - c.p0) : PT);
    NSF_PHASE(nsf_ph_load<NS, PT>(c, tid))
    // (NS = 1: the weights of a contraction are staged one phase ahead, beside work that does not touch the staging buffer)
    NSF_PHASE(if (NS == 1 && g.L > 1) nsf_ph_wstage<NS, PT>(c, tid, a.pk + g.pk_wt(1));
              nsf_ph_layer0<NS, PT>(c, r, tid); nsf_ph_act<NS, PT>(c, r, tid, 0, train); if (NS == 1) nsf_cp_wait())
    for (int l = 1; l < g.L; ++l) {
      NSF_PHASE(nsf_ph_gemm<NS, PT>(c, r, tid, c.act, a.pk + g.pk_wt(l), a.pk + g.pk_b(l)))
      NSF_PHASE(if (NS == 1 && l + 1 < g.L) nsf_ph_wstage<NS, PT>(c, tid, a.pk + g.pk_wt(l + 1));
                nsf_ph_act<NS, PT>(c, r, tid, l, train); if (NS == 1) nsf_cp_wait())
    }
    NSF_PHASE(nsf_ph_out<NS, PT>(c, tid))
    if (mode == NSF_MODE_FWD) {
      NSF_PHASE(nsf_ph_store_out<NS, PT>(c, tid))
      continue;
    }
    if (jet) {
      if constexpr (NS == 4) { NSF_PHASE(nsf_ph_resid<NS, PT>(c, r, tid, train)) }
    } else {
      NSF_PHASE(nsf_ph_mse<NS, PT>(c, r, tid))
    }
    if (!train) continue;
    NSF_PHASE(nsf_ph_outbwd<NS, PT>(c, r, tid))
    for (int l = g.L - 1; l >= 1; --l) {
      NSF_PHASE(if (NS == 1) nsf_ph_wstage<NS, PT>(c, tid, a.pk + g.pk_w(l));
                nsf_ph_zbar<NS, PT>(c, r, tid, l); if (NS == 1) nsf_cp_wait())
      NSF_PHASE(nsf_ph_wgrad<NS, PT>(c, tid, l); nsf_ph_gemm<NS, PT>(c, r, tid, c.zb, a.pk + g.pk_w(l), (const float*)0))
    }
    NSF_PHASE(nsf_ph_zbar<NS, PT>(c, r, tid, 0))
    NSF_PHASE(nsf_ph_l0bwd<NS, PT>(c, tid))
    c.first = 0;
  }
  if (mode != NSF_MODE_FWD && c.grow) {
    NSF_PHASE(nsf_ph_loss_a<NS, PT>(c, r, tid))
    NSF_PHASE(nsf_ph_loss_b<NS, PT>(c, tid, a.accumulate != 0))
  }
}

// floats of dynamic shared memory the program needs
static inline long long nsf_ffma_smem_floats(int NS, int PT, int HP) {
  return 2LL * NS * HP * PT + 2 * PT + NS * 4 * PT + PT * 6 + (NS == 1 ? (long long)HP * HP : 0);
}
