// tcgen05 jet kernel, second generation ("layer-major"): hidden = 80, 2..6 hidden layers.
//
// What round-1 profiling of the first kernel (nsf_umma_jet.cu) showed and what this one changes:
//   * with the weights as the shared-memory A operand every N = 32 MMA re-reads 4 KB of weights: ~40 cycles of
//     operand fetch for 16 cycles of math (ncu: tensor pipe 27 % busy, issue-side counters: 1300 cycles per
//     stage).  Here the weights of the current layer live in TENSOR MEMORY (A operand from TMEM, lane = neuron,
//     one column per input neuron, hi | lo = 160 columns, double buffered); an MMA only reads the 1.5 KB
//     activation image from shared memory.
//   * weights were streamed per tile pair (16 points).  Here a CTA walks a super-batch of NT tiles LAYER BY LAYER
//     (forward layers 1..L, then reverse layers L-1..1), so a layer's weights are fetched once per super-batch and
//     only ONE weight-gradient accumulator (80 TMEM columns) is live; it is flushed after every layer phase, which
//     also bounds the number of truncating tensor-core accumulations per value.
//   * the per-tile state between layer phases (stashed pre-activations, adjoints) goes through global memory
//     (L2 resident: the super-batch working set is ~0.4 MB per CTA); the same thread writes and re-reads it.
//   * no CTA-wide barrier in the steady state: epilogue threads hand operands to the issuer through an mbarrier
//     per slot (`ready`), the issuer hands results back through tcgen05.commit (`done`); the items (layer, tile)
//     form one seamless stream, so the MMAs of item k+1 overlap the epilogue of item k across phase boundaries.
//
// Orientation, operand images, 3xTF32 ordering and thread mapping are those of the first kernel: TMEM lane =
// neuron, column = 4*point + stream; epilogue warp 4*sub + q (q = 0..2) owns quadrant q and points 3*sub..3*sub+2
// of a 12-point tile; warp 3 issues.
#include "nsf_internal.h"
#include "nsf_tc.cuh"
#include "nsf_math.cuh"

using namespace nsftc;

namespace {

constexpr int KP = 80;
constexpr int P = 12;              // points per tile
constexpr int PPT = 3;             // points per epilogue thread
constexpr int NSUB = P / PPT;      // 4
constexpr int NCOL = 4 * P;        // 48
constexpr int NW = 80;
constexpr int MAXL = 6;
constexpr int NT_MAX = 8;          // tiles per super-batch (run-time choice <= NT_MAX)
constexpr int NWARPS = 15, NTHREADS = NWARPS * 32, NEPI = 12 * 32, ISSUER_WARP = 3;

// tensor memory columns
constexpr uint32_t TM_W0 = 0, TM_W1 = 160, TM_DW = 320, TM_D0 = 400, TM_D1 = 448;   // 496 of 512

// shared memory
constexpr uint32_t R_LBO = 144, R_SBO = (KP / 4) * R_LBO;      // 2880
constexpr uint32_t RB = (NCOL / 8) * R_SBO;                    // 17280
constexpr uint32_t C_SP = (KP / 8) * 128;                      // 1280
constexpr uint32_t CB = P * C_SP;                              // 15360
constexpr uint32_t SLOT = 2 * RB + 4 * CB;                     // 96000: R hi, R lo, ZC hi, ZC lo, AC hi, AC lo
constexpr uint32_t OFF_MISC = 2 * SLOT + 1024;                 // 1 KB of slack: the wgrad A operand over-reads 768 B
constexpr uint32_t MISC = 2560;
constexpr uint32_t SMEM_BYTES = OFF_MISC + MISC;               // 195584

constexpr int WIMG_FLOATS = KP * 160;                          // one stage image: rows j, hi[80] | lo[80]

struct U2Args {
  NsfNetGeom g;
  const float* pk;
  const float* wimg;     // (2L-1) stage images: forward l = 1..L, reverse l = L-1..1
  const float* x; const float* y; long long n;
  int train;
  const float* e_in; const float* vtm_in; float* vtm_out; const float* w;
  float inv_Re, vis_t0, alpha_evm, cs1, cs2, k4, c_eq;
  int has_evm;
  float* resid_out; float* vis_t_out; float* ebar_out;
  float* stash;          // [grid][NT][L][P][KP][4]
  float* abar;           // [grid][NT][P][KP][4]
  float* scratch;
  int nt;                // tiles per super-batch
  long long n_sb;        // super-batches
  long long* dbg;        // optional per-warp cycle counters [grid][16][16]
};

struct Misc2 {
  uint64_t ready[2];     // operands of the slot written, previous results of the slot consumed (NEPI arrivals)
  uint64_t done[2];      // the slot's MMAs have completed (tcgen05.commit)
  uint64_t dwfree;       // the weight-gradient accumulator has been flushed (NEPI arrivals)
  uint32_t tmem_base;
  uint32_t pad[5];
  float ov[2][P][16];
  float red[P][12];
};
static_assert(sizeof(Misc2) <= MISC, "misc region too small");

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }
__device__ __forceinline__ void st4(uint8_t* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void st1(uint8_t* p, float a) { *reinterpret_cast<float*>(p) = a; }

struct Epi {
  int j, sub, q, lane;
  bool active;
  uint32_t lane_addr;
  uint32_t r_base[PPT], c_base[PPT], s_base[PPT];
};

__device__ __forceinline__ void store_RC(uint8_t* rh, uint8_t* zh, const Epi& e, int pi, const float v[4], bool do_r, bool do_c) {
  float hi[4], lo[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) split_tf32_fast(v[s], hi[s], lo[s]);
  if (do_r) {
    const uint32_t base = e.r_base[pi];
#pragma unroll
    for (int s = 0; s < 4; ++s) { st1(rh + base + s * 16, hi[s]); st1(rh + RB + base + s * 16, lo[s]); }
  }
  if (do_c) {
    const uint32_t off = e.c_base[pi];
    st4(zh + off, hi[0], hi[1], hi[2], hi[3]);
    st4(zh + CB + off, lo[0], lo[1], lo[2], lo[3]);
  }
}

// activations of a layer from its stashed (t, zx, zy, z_lap)
__device__ __forceinline__ void act_from_stash(const float4 s, float a[4]) {
  const float d1 = fmaf(-s.x, s.x, 1.f), d2 = -2.f * s.x * d1;
  a[0] = s.x; a[1] = d1 * s.y; a[2] = d1 * s.z;
  a[3] = fmaf(d2, fmaf(s.y, s.y, s.z * s.z), d1 * s.w);
}
// adjoint through tanh: ab (adjoint of the activations), stash of the layer -> zb
__device__ __forceinline__ void zbar_from(const float4 st, const float ab[4], float zb[4]) {
  const float t = st.x, zx = st.y, zy = st.z, zl = st.w;
  const float d1 = fmaf(-t, t, 1.f), d2 = -2.f * t * d1, d3 = -2.f * d1 * fmaf(-3.f * t, t, 1.f);
  const float q = fmaf(zx, zx, zy * zy);
  zb[3] = ab[3] * d1;
  zb[1] = fmaf(ab[1], d1, 2.f * ab[3] * d2 * zx);
  zb[2] = fmaf(ab[2], d1, 2.f * ab[3] * d2 * zy);
  zb[0] = fmaf(ab[0], d1, fmaf(ab[1] * d2, zx, fmaf(ab[2] * d2, zy, ab[3] * fmaf(d3, q, d2 * zl))));
}

// this thread's quarter of a stage image row -> TMEM weight buffer (columns 40*sub .. 40*sub+39 of 160)
__device__ __forceinline__ void load_weights(const U2Args& a, const Epi& e, uint32_t tmem, int stage_img, int buf) {
  const float* src = a.wimg + (size_t)stage_img * WIMG_FLOATS + (size_t)(e.active ? e.j : 0) * 160 + e.sub * 40;
  const uint32_t dst = tmem + e.lane_addr + (buf ? TM_W1 : TM_W0) + (uint32_t)(e.sub * 40);
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src + c * 8));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + c * 8 + 4));
    const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    tmem_st8(dst + c * 8, v);
  }
  tmem_st_wait();
}

// ---- issuer: MMAs of one item ------------------------------------------------------------------------------------
__device__ __forceinline__ void issue_main(uint8_t* smem, uint32_t tmem, int slot, int wbuf, uint32_t leader) {
  const uint32_t wa = tmem + (wbuf ? TM_W1 : TM_W0);
  const uint32_t sb = smem_u32(smem) + (uint32_t)slot * SLOT;
  const uint32_t d_col = tmem + (slot ? TM_D1 : TM_D0);
  const uint32_t idesc = idesc_tf32(128, NCOL, 0, 0);
  constexpr uint32_t BHI = desc_hi(R_SBO);
  const uint32_t bh0 = desc_lo(sb, R_LBO), bl0 = desc_lo(sb + RB, R_LBO);
  // corrections first (small accumulator), hi*hi last: see nsf_umma_jet.cu
#pragma unroll
  for (int ks = 0; ks < KP / 8; ++ks) {
    const uint32_t db = ks * ((2 * R_LBO) >> 4);
    mma_tf32_ts_elect(d_col, wa + 80 + ks * 8, bh0 + db, BHI, idesc, ks > 0, leader);   // W_lo * act_hi
    mma_tf32_ts_elect(d_col, wa + ks * 8, bl0 + db, BHI, idesc, 1, leader);            // W_hi * act_lo
  }
#pragma unroll
  for (int ks = 0; ks < KP / 8; ++ks) {
    const uint32_t db = ks * ((2 * R_LBO) >> 4);
    mma_tf32_ts_elect(d_col, wa + ks * 8, bh0 + db, BHI, idesc, 1, leader);
  }
}
__device__ __forceinline__ void issue_wgrad(uint8_t* smem, uint32_t tmem, int slot, bool zero, uint32_t leader) {
  const uint32_t zh = smem_u32(smem) + (uint32_t)slot * SLOT + 2 * RB;
  const uint32_t idesc = idesc_tf32(128, NW, 0, 0);
  const uint32_t dw_col = tmem + TM_DW;
  constexpr uint32_t CHI = desc_hi(128);
  const uint32_t ah0 = desc_lo(zh, C_SP), al0 = desc_lo(zh + CB, C_SP);
  const uint32_t bh0 = desc_lo(zh + 2 * CB, C_SP), bl0 = desc_lo(zh + 3 * CB, C_SP);
#pragma unroll
  for (int ks = 0; ks < NCOL / 8; ++ks) {
    const uint32_t d = ks * ((2 * C_SP) >> 4);
    mma_tf32_elect2(dw_col, al0 + d, CHI, bh0 + d, CHI, idesc, !(zero && ks == 0), leader);
    mma_tf32_elect2(dw_col, ah0 + d, CHI, bl0 + d, CHI, idesc, 1, leader);
  }
#pragma unroll
  for (int ks = 0; ks < NCOL / 8; ++ks) {
    const uint32_t d = ks * ((2 * C_SP) >> 4);
    mma_tf32_elect2(dw_col, ah0 + d, CHI, bh0 + d, CHI, idesc, 1, leader);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) nsf_umma_jet2_kernel(const U2Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc2* misc = reinterpret_cast<Misc2*>(smem + OFF_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const NsfNetGeom& g = a.g;
  const int L = g.L, NT = a.nt;
  const int nstage = a.train ? 2 * L - 1 : L;                    // MMA stages per super-batch
  const long long my_sb = ((long long)blockIdx.x < a.n_sb) ? (a.n_sb - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const long long total_stages = my_sb * nstage;
  const long long total_items = total_stages * NT;

  if (warp == 0) tmem_alloc(&misc->tmem_base, 512);
  if (tid == 0) {
    mbar_init(&misc->ready[0], NEPI); mbar_init(&misc->ready[1], NEPI);
    mbar_init(&misc->done[0], 1); mbar_init(&misc->done[1], 1);
    mbar_init(&misc->dwfree, NEPI);
    mbar_fence_init();
  }
  for (uint32_t i = tid * 16; i < OFF_MISC; i += NTHREADS * 16) *reinterpret_cast<float4*>(smem + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_base;

  if (warp == ISSUER_WARP) {
    // =========================== issuer ===========================
    const uint32_t leader = elect_one();
    uint32_t rph[2] = {0, 0}, dwph = 0;
    long long k = 0, bwd_stages = 0;
    long long ic[4] = {0, 0, 0, 0}, t0 = 0, t1 = 0;    // ready wait, issue, dw wait, items
    for (long long gs = 0; gs < total_stages; ++gs) {
      const int sidx = (int)(gs % nstage);
      const bool bwd = sidx >= L;
      for (int i = 0; i < NT; ++i, ++k) {
        const int slot = (int)(k & 1);
        if (a.dbg) t0 = clock64();
        mbar_wait(&misc->ready[slot], rph[slot]); rph[slot] ^= 1;
        tc_fence_after();
        if (a.dbg) { t1 = clock64(); ic[0] += t1 - t0; t0 = t1; }
        issue_main(smem, tmem, slot, (int)(gs & 1), leader);
        if (a.dbg) { t1 = clock64(); ic[1] += t1 - t0; t0 = t1; }
        if (bwd) {
          if (i == 0 && bwd_stages > 0) { mbar_wait(&misc->dwfree, dwph); dwph ^= 1; tc_fence_after(); }
          if (a.dbg) { t1 = clock64(); ic[2] += t1 - t0; t0 = t1; }
          issue_wgrad(smem, tmem, slot, i == 0, leader);
        }
        mma_commit_elect(&misc->done[slot], leader);
        if (a.dbg) { t1 = clock64(); ic[1] += t1 - t0; ic[3] += 1; }
      }
      if (bwd) ++bwd_stages;
    }
    if (a.dbg && lane == 0)
      for (int q = 0; q < 4; ++q) a.dbg[((size_t)blockIdx.x * 16 + warp) * 16 + q] = ic[q];
  } else if ((warp & 3) != 3) {
    // =========================== epilogue warps ===========================
    Epi e;
    e.lane = lane; e.q = warp & 3; e.sub = warp >> 2;
    e.j = e.q * 32 + lane;
    e.active = e.j < KP;
    e.lane_addr = (uint32_t)(e.q * 32) << 16;
    const int jj = e.active ? e.j : 0;
    {
      const uint32_t r_off = (uint32_t)(jj >> 2) * R_LBO + (uint32_t)(jj & 3) * 4;
      const uint32_t c_off = (uint32_t)(jj >> 3) * 128 + (uint32_t)(jj & 7) * 16;
#pragma unroll
      for (int pi = 0; pi < PPT; ++pi) {
        const int p = e.sub * PPT + pi;
        e.r_base[pi] = (uint32_t)(p >> 1) * R_SBO + (uint32_t)((p & 1) * 4) * 16 + r_off;
        e.c_base[pi] = (uint32_t)p * C_SP + c_off;
        e.s_base[pi] = (uint32_t)((p * KP + jj) * 4);
      }
    }
    const float* pk = a.pk;
    const float w0x = __ldg(pk + g.pk_w0x() + jj), w0y = __ldg(pk + g.pk_w0y() + jj), b0 = __ldg(pk + g.pk_b0() + jj);
    const float wl0 = __ldg(pk + g.pk_wl() + jj), wl1 = __ldg(pk + g.pk_wl() + g.HP + jj), wl2 = __ldg(pk + g.pk_wl() + 2 * g.HP + jj);
    constexpr size_t LSTR = (size_t)P * KP * 4;                        // floats per (tile, layer) in the stash
    float* stash_cta = a.stash + (size_t)blockIdx.x * NT_MAX * MAXL * LSTR;
    float* abar_cta = a.abar + (size_t)blockIdx.x * NT_MAX * LSTR;
    float* grow = a.scratch ? a.scratch + (size_t)blockIdx.x * g.gs_row() : nullptr;
    float gw0x = 0.f, gw0y = 0.f, gwl[3] = {0.f, 0.f, 0.f};
    float gb[MAXL];
#pragma unroll
    for (int i = 0; i < MAXL; ++i) gb[i] = 0.f;
    float lossacc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, gbl[3] = {0.f, 0.f, 0.f};
    uint32_t dph[2] = {0, 0};
    bool first_flush = true;

    long long ec[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // pre, fence+arrive, post (incl. MMA wait), stage end, items, MMA wait
    // item k -> (super-batch, stage, tile)
    auto pre = [&](long long k) {
      const long long gs = k / NT;
      const int i = (int)(k % NT), sidx = (int)(gs % nstage), slot = (int)(k & 1);
      const long long sb = (long long)blockIdx.x + (gs / nstage) * gridDim.x;
      const long long p0 = (sb * NT + i) * P;
      const int nvalid = (int)((a.n - p0) < 0 ? 0 : ((a.n - p0) < P ? (a.n - p0) : P));
      uint8_t* sbp = smem + (size_t)slot * SLOT;
      float* st_tile = stash_cta + (size_t)i * MAXL * LSTR;
      if (!e.active) return;
      if (sidx < L) {
        const int l = sidx + 1;                       // forward layer whose MMA follows (L = output layer)
        if (l == 1) {                                 // a^0 straight from the points
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            const int p = e.sub * PPT + pi;
            const float xv = p < nvalid ? __ldg(a.x + p0 + p) : 0.f, yv = p < nvalid ? __ldg(a.y + p0 + p) : 0.f;
            const float t = nsf_tanh_fast(fmaf(w0x, xv, fmaf(w0y, yv, b0)));
            const float4 s0 = make_float4(t, w0x, w0y, 0.f);
            if (a.train) __stcs(reinterpret_cast<float4*>(st_tile + e.s_base[pi]), s0);
            float v[4];
            act_from_stash(s0, v);
            store_RC(sbp, nullptr, e, pi, v, true, false);
          }
        } else {
          float4 s[PPT];
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) s[pi] = __ldcs(reinterpret_cast<const float4*>(st_tile + (size_t)(l - 1) * LSTR + e.s_base[pi]));
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            float v[4];
            act_from_stash(s[pi], v);
            store_RC(sbp, nullptr, e, pi, v, true, false);
          }
        }
      } else {
        const int l = 2 * L - 1 - sidx;               // reverse layer L-1 .. 1
        float4 ab[PPT], sl[PPT], sm[PPT];
#pragma unroll
        for (int pi = 0; pi < PPT; ++pi) {
          ab[pi] = __ldcs(reinterpret_cast<const float4*>(abar_cta + (size_t)i * LSTR + e.s_base[pi]));
          sl[pi] = __ldcs(reinterpret_cast<const float4*>(st_tile + (size_t)l * LSTR + e.s_base[pi]));
          sm[pi] = __ldcs(reinterpret_cast<const float4*>(st_tile + (size_t)(l - 1) * LSTR + e.s_base[pi]));
        }
        float sb0 = 0.f;
#pragma unroll
        for (int pi = 0; pi < PPT; ++pi) {
          const float abv[4] = {ab[pi].x, ab[pi].y, ab[pi].z, ab[pi].w};
          float zb[4], av[4];
          zbar_from(sl[pi], abv, zb);
          sb0 += zb[0];
          store_RC(sbp, sbp + 2 * RB, e, pi, zb, true, true);
          act_from_stash(sm[pi], av);
          store_RC(nullptr, sbp + 2 * RB + 2 * CB, e, pi, av, false, true);
        }
#pragma unroll
        for (int q = 0; q < MAXL; ++q) if (q == l) gb[q] += sb0;
      }
    };

    long long ec5_dummy = 0; (void)ec5_dummy;
    auto post = [&](long long k) {
      const long long gs = k / NT;
      const int i = (int)(k % NT), sidx = (int)(gs % nstage), slot = (int)(k & 1);
      const long long sb = (long long)blockIdx.x + (gs / nstage) * gridDim.x;
      const long long p0 = (sb * NT + i) * P;
      const int nvalid = (int)((a.n - p0) < 0 ? 0 : ((a.n - p0) < P ? (a.n - p0) : P));
      float* st_tile = stash_cta + (size_t)i * MAXL * LSTR;
      const uint32_t d_addr = tmem + e.lane_addr + (slot ? TM_D1 : TM_D0) + (uint32_t)(e.sub * 4 * PPT);
      // operands from L2 first, then the (usually already satisfied) wait for the MMAs
      float bias_s = 0.f;
      float4 s0[PPT];
      float xv[PPT], yv[PPT];
      const bool is_out = sidx == L - 1, is_l1 = sidx == nstage - 1 && a.train;
      if (e.active && sidx < L - 1) bias_s = __ldg(pk + g.pk_b(sidx + 1) + e.j);
      if (e.active && a.train && (is_out || is_l1)) {
#pragma unroll
        for (int pi = 0; pi < PPT; ++pi)
          s0[pi] = __ldcs(reinterpret_cast<const float4*>(st_tile + (size_t)(is_out ? L - 1 : 0) * LSTR + e.s_base[pi]));
      }
      if (is_l1) {
#pragma unroll
        for (int pi = 0; pi < PPT; ++pi) {
          const int p = e.sub * PPT + pi;
          xv[pi] = p < nvalid ? __ldg(a.x + p0 + p) : 0.f;
          yv[pi] = p < nvalid ? __ldg(a.y + p0 + p) : 0.f;
        }
      }
      long long tw = 0;
      if (a.dbg) tw = clock64();
      mbar_wait(&misc->done[slot], dph[slot]); dph[slot] ^= 1;
      tc_fence_after();
      if (a.dbg) ec[5] += clock64() - tw;
      float d[PPT][4];
      tmem_ld8(d_addr, &d[0][0]);
      tmem_ld4(d_addr + 8, &d[2][0]);
      tmem_ld_wait();

      if (sidx < L - 1) {
        // ---- hidden layer l = sidx + 1 forward: z -> stash -----------------------------------------
        if (e.active) {
          const int l = sidx + 1;
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            const float t = nsf_tanh_fast(d[pi][0] + bias_s);
            __stcs(reinterpret_cast<float4*>(st_tile + (size_t)l * LSTR + e.s_base[pi]), make_float4(t, d[pi][1], d[pi][2], d[pi][3]));
          }
        }
      } else if (is_out) {
        // ---- output layer: gather, residuals, adjoint seeds ------------------------------------------
        if (e.q == 0 && lane < 3) {
          const float bo = __ldg(pk + g.pk_bl() + lane);
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            float* ov = misc->ov[slot][e.sub * PPT + pi];
            ov[0 * 4 + lane] = d[pi][0] + bo; ov[1 * 4 + lane] = d[pi][1]; ov[2 * 4 + lane] = d[pi][2]; ov[3 * 4 + lane] = d[pi][3];
          }
        }
        epi_bar();
        if (tid < P) {
          const int p = tid;
          const bool ok = p < nvalid;
          const long long gp = p0 + p;
          float* ov = misc->ov[slot][p];
          const float u = ov[0], v = ov[1];
          const float ux = a.cs1 * ov[4], vx = a.cs1 * ov[5], px = a.cs1 * ov[6];
          const float uy = a.cs1 * ov[8], vy = a.cs1 * ov[9], py = a.cs1 * ov[10];
          const float ul = a.cs2 * ov[12], vl = a.cs2 * ov[13];
          float ee = 0.f, vis = 0.f;
          if (a.has_evm) {
            ee = ok ? __ldg(a.e_in + gp) : 0.f;
            vis = a.vis_t0;
            if (a.vtm_in && ok) vis = fminf(a.vis_t0, __ldg(a.vtm_in + gp));
          }
          const float nu = a.inv_Re + vis;
          const float eq1 = (u * ux + v * uy) + px - nu * ul;
          const float eq2 = (u * vx + v * vy) + py - nu * vl;
          const float eq3 = ux + vy;
          const float eq4 = a.has_evm ? (eq1 * (u - 0.5f) + eq2 * (v - 0.5f)) - ee : 0.f;
          const float w = (a.w && ok) ? __ldg(a.w + gp) : 1.f;
          if (ok) {
            lossacc[0] += w * eq1 * eq1; lossacc[1] += w * eq2 * eq2; lossacc[2] += w * eq3 * eq3; lossacc[3] += w * eq4 * eq4;
            lossacc[4] += vis; lossacc[5] += 1.f;
            if (a.resid_out) { a.resid_out[gp] = eq1; a.resid_out[a.n + gp] = eq2; a.resid_out[2 * a.n + gp] = eq3; a.resid_out[3 * a.n + gp] = eq4; }
            if (a.vis_t_out) a.vis_t_out[gp] = vis;
            if (a.has_evm && a.vtm_out) a.vtm_out[gp] = a.alpha_evm * fabsf(ee);
          }
          if (a.train) {
            const float cw = ok ? a.c_eq * w : 0.f;
            const float g1 = cw * (2.f * eq1 + a.k4 * eq4 * (u - 0.5f));
            const float g2 = cw * (2.f * eq2 + a.k4 * eq4 * (v - 0.5f));
            const float g3 = 2.f * cw * eq3;
            const float g4 = a.k4 * cw * eq4;
            ov[0] = g1 * ux + g2 * vx + g4 * eq1; ov[1] = g1 * uy + g2 * vy + g4 * eq2; ov[2] = 0.f;
            ov[4] = a.cs1 * (g1 * u + g3); ov[5] = a.cs1 * (g2 * u); ov[6] = a.cs1 * g1;
            ov[8] = a.cs1 * (g1 * v); ov[9] = a.cs1 * (g2 * v + g3); ov[10] = a.cs1 * g2;
            ov[12] = -a.cs2 * nu * g1; ov[13] = -a.cs2 * nu * g2; ov[14] = 0.f;
            gbl[0] += ov[0]; gbl[1] += ov[1]; gbl[2] += ov[2];
            if (a.ebar_out && ok) a.ebar_out[gp] = -g4;
          }
        }
        if (a.train) {
          epi_bar();
          if (e.active) {
#pragma unroll
            for (int pi = 0; pi < PPT; ++pi) {
              const float* ov = misc->ov[slot][e.sub * PPT + pi];
              float ab[4], act[4];
#pragma unroll
              for (int st = 0; st < 4; ++st) ab[st] = fmaf(ov[st * 4 + 0], wl0, fmaf(ov[st * 4 + 1], wl1, ov[st * 4 + 2] * wl2));
              __stcs(reinterpret_cast<float4*>(abar_cta + (size_t)i * LSTR + e.s_base[pi]), make_float4(ab[0], ab[1], ab[2], ab[3]));
              act_from_stash(s0[pi], act);       // a^{L-1}: weight gradient of the output layer
#pragma unroll
              for (int st = 0; st < 4; ++st) {
                gwl[0] = fmaf(ov[st * 4 + 0], act[st], gwl[0]);
                gwl[1] = fmaf(ov[st * 4 + 1], act[st], gwl[1]);
                gwl[2] = fmaf(ov[st * 4 + 2], act[st], gwl[2]);
              }
            }
          }
          epi_bar();    // ov[slot] is rewritten two items later; keep the readers ahead of the next gather
        }
      } else {
        // ---- reverse layer l = 2L-1-sidx: D holds the adjoint of a^{l-1} ------------------------------------
        const int l = 2 * L - 1 - sidx;
        if (e.active) {
          if (l >= 2) {
#pragma unroll
            for (int pi = 0; pi < PPT; ++pi)
              __stcs(reinterpret_cast<float4*>(abar_cta + (size_t)i * LSTR + e.s_base[pi]), make_float4(d[pi][0], d[pi][1], d[pi][2], d[pi][3]));
          } else {
            // layer 0: inputs are x, y and the constant unit tangents
#pragma unroll
            for (int pi = 0; pi < PPT; ++pi) {
              float zb[4];
              zbar_from(s0[pi], d[pi], zb);
              gb[0] += zb[0];
              gw0x += fmaf(zb[0], xv[pi], zb[1]);
              gw0y += fmaf(zb[0], yv[pi], zb[2]);
            }
          }
        }
      }
    };

    long long t0 = 0, t1 = 0;
    if (total_items > 0) {
      load_weights(a, e, tmem, 0, 0);
      if (total_stages > 1) load_weights(a, e, tmem, 1 % nstage, 1);
      pre(0);
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(&misc->ready[0]);
      for (long long k = 0; k < total_items; ++k) {
        if (a.dbg) t0 = clock64();
        if (k + 1 < total_items) {
          pre(k + 1);
          if (a.dbg) { t1 = clock64(); ec[0] += t1 - t0; t0 = t1; }
          fence_async_smem();
          tc_fence_before();
          mbar_arrive(&misc->ready[(k + 1) & 1]);
          if (a.dbg) { t1 = clock64(); ec[1] += t1 - t0; t0 = t1; }
        }
        post(k);
        if (a.dbg) { t1 = clock64(); ec[2] += t1 - t0; ec[4] += 1; t0 = t1; }
        if ((k + 1) % NT == 0) {                       // item k closes stage gs
          const long long gs = k / NT;
          const int sidx = (int)(gs % nstage);
          if (sidx >= L && grow) {                     // reverse stage: flush the weight-gradient accumulator of layer l
            const int l = 2 * L - 1 - sidx;
            float v[20];
            const uint32_t src = tmem + e.lane_addr + TM_DW + (uint32_t)(e.sub * 20);
            tmem_ld16(src, v);
            tmem_ld4(src + 16, v + 16);
            tmem_ld_wait();
            if (e.active) {
              float4* dst = reinterpret_cast<float4*>(grow + g.gs_w(l) + (size_t)e.j * g.HP + e.sub * 20);
#pragma unroll
              for (int c = 0; c < 5; ++c) {
                float4 o = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                if (!first_flush || gs >= nstage) { const float4 pv = dst[c]; o.x += pv.x; o.y += pv.y; o.z += pv.z; o.w += pv.w; }
                dst[c] = o;
              }
            }
            tc_fence_before();
            mbar_arrive(&misc->dwfree);
          }
          if (gs + 2 < total_stages) {                 // the buffer of stage gs is free: fetch the weights of stage gs + 2
            load_weights(a, e, tmem, (int)((gs + 2) % nstage), (int)(gs & 1));
            tc_fence_before();
          }
          if (a.dbg) { t1 = clock64(); ec[3] += t1 - t0; }
        }
      }
    }
    (void)first_flush;
    if (a.dbg && lane == 0)
      for (int q = 0; q < 8; ++q) a.dbg[((size_t)blockIdx.x * 16 + warp) * 16 + q] = ec[q];

    // ---- CTA epilogue: thread-local partials -> this CTA's row ---------------------------------------------
    if (grow) {
      float* redf = reinterpret_cast<float*>(smem);    // [NSUB-1][KP][16]; every MMA has completed
      epi_bar();
      if (e.sub >= 1 && e.active) {
        float* r = redf + ((e.sub - 1) * KP + e.j) * 16;
        r[0] = gw0x; r[1] = gw0y; r[2] = gwl[0]; r[3] = gwl[1]; r[4] = gwl[2];
#pragma unroll
        for (int i = 0; i < MAXL; ++i) r[5 + i] = gb[i];
      }
      if (tid < P) {
#pragma unroll
        for (int q = 0; q < 6; ++q) misc->red[tid][q] = lossacc[q];
#pragma unroll
        for (int q = 0; q < 3; ++q) misc->red[tid][6 + q] = gbl[q];
      }
      epi_bar();
      if (a.train && e.sub == 0 && e.active) {
        float r[5 + MAXL];
#pragma unroll
        for (int i = 0; i < 5 + MAXL; ++i) {
          r[i] = 0.f;
#pragma unroll
          for (int q = 0; q < NSUB - 1; ++q) r[i] += redf[(q * KP + e.j) * 16 + i];
        }
        const int j = e.j;
        grow[g.gs_w0x() + j] = gw0x + r[0];
        grow[g.gs_w0y() + j] = gw0y + r[1];
        grow[g.gs_b0() + j] = gb[0] + r[5];
        grow[g.gs_wl() + j] = gwl[0] + r[2];
        grow[g.gs_wl() + g.HP + j] = gwl[1] + r[3];
        grow[g.gs_wl() + 2 * g.HP + j] = gwl[2] + r[4];
        grow[g.gs_wl() + 3 * g.HP + j] = 0.f;
#pragma unroll
        for (int l = 1; l < MAXL; ++l)
          if (l < L) grow[g.gs_b(l) + j] = gb[l] + r[5 + l];
      }
      if (tid < NSF_LOSS_SLOTS) {
        float v = 0.f;
        if (tid < 6) for (int p = 0; p < P; ++p) v += misc->red[p][tid];
        grow[g.gs_loss() + tid] = v;
      }
      if (a.train && tid < 4) {
        float v = 0.f;
        if (tid < 3) for (int p = 0; p < P; ++p) v += misc->red[p][6 + tid];
        grow[g.gs_bl() + tid] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// stage images: thread per (stage, row, col)
__global__ void nsf_umma2_pack_kernel(NsfNetGeom g, const float* __restrict__ flat, float* __restrict__ wimg) {
  const int L = g.L, H = g.H;
  const int n_img = 2 * L - 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_img * KP * KP) return;
  const int img = (int)(idx / (KP * KP)), r = (int)(idx % (KP * KP)) / KP, c = (int)(idx % KP);   // row r (TMEM lane), contraction index c
  float v = 0.f;
  if (img < L - 1) {                 // forward hidden layer l = img + 1: W_l[r][c]
    const int l = img + 1, fo = 3 * H + (l - 1) * (H * H + H);
    if (r < H && c < H) v = flat[fo + r * H + c];
  } else if (img == L - 1) {         // output layer rows o < n_out
    const int fo = 3 * H + (L - 1) * (H * H + H);
    if (r < g.n_out && c < H) v = flat[fo + r * H + c];
  } else {                           // reverse layer l = 2L - 1 - img: W_l^T[r][c] = W_l[c][r]
    const int l = 2 * L - 1 - img, fo = 3 * H + (l - 1) * (H * H + H);
    if (r < H && c < H) v = flat[fo + c * H + r];
  }
  float hi, lo;
  split_tf32(v, hi, lo);
  float* row = wimg + (size_t)img * WIMG_FLOATS + (size_t)r * 160;
  row[c] = hi;
  row[80 + c] = lo;
}

struct Umma2State {
  float* wimg = nullptr;
  float* stash = nullptr;
  float* abar = nullptr;
  long long* dbg = nullptr;
  int dbg_on = 0, last_grid = 0;
  int grid = 0;
};

}  // namespace

int nsf_umma2_init(NsfCtx* ctx) {
  if (ctx->umma2) return NSF_OK;
  if (!nsf_umma_supported(ctx->main.g)) { nsf_set_error("tcgen05 path covers hidden = 80, 2..6 hidden layers"); return NSF_E_SHAPE; }
  Umma2State* s = new Umma2State();
  const NsfNetGeom& g = ctx->main.g;
  s->grid = ctx->sms < ctx->main.rows ? ctx->sms : ctx->main.rows;
  const size_t lstr = (size_t)P * KP * 4;
  NSF_CUDA_OK(cudaMalloc((void**)&s->wimg, (size_t)(2 * g.L - 1) * WIMG_FLOATS * sizeof(float)));
  NSF_CUDA_OK(cudaMalloc((void**)&s->stash, (size_t)s->grid * NT_MAX * MAXL * lstr * sizeof(float)));
  NSF_CUDA_OK(cudaMalloc((void**)&s->abar, (size_t)s->grid * NT_MAX * lstr * sizeof(float)));
  NSF_CUDA_OK(cudaFuncSetAttribute(nsf_umma_jet2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  ctx->ws_bytes += (long long)((2 * g.L - 1) * WIMG_FLOATS + (size_t)s->grid * NT_MAX * (MAXL + 1) * lstr) * 4;
  ctx->umma2 = s;
  return NSF_OK;
}

void nsf_umma2_free(NsfCtx* ctx) {
  Umma2State* s = (Umma2State*)ctx->umma2;
  if (!s) return;
  cudaFree(s->wimg); cudaFree(s->stash); cudaFree(s->abar); if (s->dbg) cudaFree(s->dbg);
  delete s;
  ctx->umma2 = nullptr;
}

int nsf_umma2_grid(const NsfCtx* ctx, long long n, int nt) {
  const long long n_sb = (n + (long long)nt * P - 1) / ((long long)nt * P);
  const int cap = ctx->sms < ctx->main.rows ? ctx->sms : ctx->main.rows;
  return (int)(n_sb < cap ? n_sb : cap);
}

int nsf_umma2_launch(NsfCtx* ctx, const NsfKernelArgs& k, const float* flat_params, int* grid_out, nsf_stream_t st, int* launches) {
  Umma2State* s = (Umma2State*)ctx->umma2;
  const NsfNetGeom& g = ctx->main.g;
  const long long tot = (long long)(2 * g.L - 1) * KP * KP;
  nsf_umma2_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, flat_params, s->wimg);
  NSF_CUDA_OK(cudaGetLastError());
  ++*launches;
  U2Args a;
  a.g = g; a.pk = k.pk; a.wimg = s->wimg; a.x = k.x; a.y = k.y; a.n = k.n;
  a.train = k.mode == NSF_MODE_JET_STEP ? 1 : 0;
  a.e_in = k.e_in; a.vtm_in = k.vtm_in; a.vtm_out = k.vtm_out; a.w = k.w;
  a.inv_Re = k.inv_Re; a.vis_t0 = k.vis_t0; a.alpha_evm = k.alpha_evm; a.cs1 = k.cs1; a.cs2 = k.cs2; a.k4 = k.k4; a.c_eq = k.c_eq;
  a.has_evm = k.has_evm;
  a.resid_out = k.resid_out; a.vis_t_out = k.vis_t_out; a.ebar_out = k.ebar_out;
  a.stash = s->stash; a.abar = s->abar;
  a.scratch = a.train ? k.scratch : nullptr;
  a.nt = ctx->umma2_nt > 0 && ctx->umma2_nt <= NT_MAX ? ctx->umma2_nt : 4;
  a.n_sb = (k.n + (long long)a.nt * P - 1) / ((long long)a.nt * P);
  a.dbg = s->dbg_on ? s->dbg : nullptr;
  const int grid = nsf_umma2_grid(ctx, k.n, a.nt);
  if (grid <= 0) { *grid_out = 0; return NSF_OK; }
  nsf_umma_jet2_kernel<<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
  NSF_CUDA_OK(cudaGetLastError());
  ++*launches;
  s->last_grid = grid;
  *grid_out = grid;
  return NSF_OK;
}

int nsf_umma2_stage_cycles(NsfCtx* ctx, double* out) {
  int rc = nsf_umma2_init(ctx);
  if (rc != NSF_OK) return rc;
  Umma2State* s = (Umma2State*)ctx->umma2;
  if (!s->dbg) { NSF_CUDA_OK(cudaMalloc((void**)&s->dbg, (size_t)s->grid * 256 * sizeof(long long))); NSF_CUDA_OK(cudaMemset(s->dbg, 0, (size_t)s->grid * 256 * sizeof(long long))); }
  s->dbg_on = 1;
  if (!out) return NSF_OK;
  NSF_CUDA_OK(cudaDeviceSynchronize());
  const int n = s->last_grid > 0 ? s->last_grid : 1;
  long long* h = new long long[(size_t)n * 256];
  if (cudaMemcpy(h, s->dbg, (size_t)n * 256 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) { delete[] h; nsf_set_error("cudaMemcpy failed"); return NSF_E_CUDA; }
  for (int i = 0; i < 256; ++i) { double acc = 0; for (int c = 0; c < n; ++c) acc += (double)h[(size_t)c * 256 + i]; out[i] = acc / n; }
  delete[] h;
  return NSF_OK;
}
