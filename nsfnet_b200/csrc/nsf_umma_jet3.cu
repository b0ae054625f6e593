// tcgen05 jet kernel, third generation ("tile-major, weights in tensor memory"): hidden = 80, 2..6 hidden layers.
//
// Same tile-major structure, thread mapping, operand images and 3xTF32 ordering as nsf_umma_jet.cu (8-point tile slots in
// flight -- four here --, TMEM lane = neuron, column = 4*point + stream, mbarrier-only steady state), with ONE change of plan
// that round-1 profiling asked for (profiles/r1_umma_v10_1M_ncu_summary.txt: tensor pipe 43 % active, the N = 32 MMAs spend
// 43 cycles fetching a 4 KB weight operand from shared memory for 16 cycles of math):
//   * the weights of the current stage live in TENSOR MEMORY (A operand from TMEM: 17.6 cycles per N = 32 MMA instead of
//     43, profiles/r1_probe_mma_timing.txt).  A producer lane TMA-stages the stage image (rows hi[80] | lo[80], padded to 164
//     floats so that the LDS.128 of a column are conflict free) into ONE shared-memory buffer; the epilogue warps copy it into
//     the 160 TMEM columns (each thread its neuron's row quarter: 10 LDS.128 + 5 tcgen05.st) right after the last slot's MMAs
//     of the previous stage have completed, so one TMEM buffer is enough; the issuer waits for `wfull` before a stage's first MMA;
//   * the tensor-memory columns for it come from the weight-gradient accumulators: a layer's dW is accumulated over the
//     slots of ONE tile group only, then added to the CTA's gradient row in L2 by vector reductions (see flush_dw) -- two
//     rotating 80-column accumulators instead of five persistent ones.  That also bounds the number of truncating tensor-core
//     accumulations per value at 48.
// TMEM map: [0,160) weights hi | lo, [160,320) two rotating dW accumulators, [320,448) the four D slots.
#include "nsf_internal.h"
#include "nsf_tc.cuh"
#include "nsf_math.cuh"

using namespace nsftc;

namespace {

constexpr int KP = 80;            // hidden width handled by this kernel
constexpr int P = 8;              // points per tile slot
constexpr int NS = 4;             // tile slots in flight (shared memory has room for four once the weights live in tensor memory)
constexpr int NCOL = 4 * P;       // MMA N of the forward / dgrad contractions
constexpr int NW = 80;            // MMA N of the weight-gradient contraction (columns of dW_l)
constexpr int MAXL = 6;
constexpr int PPT = 2;            // points per epilogue thread and stage
constexpr int NSUB = P / PPT;     // epilogue warps per TMEM lane quadrant
constexpr int NWARPS = 4 * NSUB - 1;   // warps 4*sub + q, q = 0..2 epilogue; warp 3 issuer; warps 7, 11 idle
constexpr int NTHREADS = NWARPS * 32;  // 480
constexpr int NEPI = 3 * NSUB * 32;    // 384 epilogue threads
constexpr int ISSUER_WARP = 3;
constexpr int NDW = 2;            // rotating weight-gradient accumulators

constexpr uint32_t TM_W = 0, TM_DW = 160, TM_D = TM_DW + NDW * NW;   // 0, 160, 320
static_assert(TM_D + NS * NCOL <= 512, "tensor memory");
constexpr int WROW = 164;                       // floats per image row: hi[80] | lo[80] | 4 pad (quarter-warp LDS.128 of a column: 8 bank groups)
constexpr int WIMG_FLOATS = KP * WROW;          // one stage image in global memory
constexpr uint32_t WIMG_BYTES = WIMG_FLOATS * 4;   // 52480
constexpr int PRODUCER_WARP = 7;

constexpr uint32_t R_ATOM = 512;                // 4 neurons x 128 B (32 columns n) of an R image
constexpr uint32_t RB = (KP / 4) * R_ATOM;      // 10240: one R image
constexpr uint32_t SLOT = 4 * RB;                // 40960: R hi, R lo (z-bar or activations), RA hi, RA lo (activations of the layer below)
constexpr uint32_t OFF_SLOT = 0;
constexpr uint32_t OFF_WS = OFF_SLOT + NS * SLOT;    // 122880: two staging buffers for the stage images (TMA destination)
constexpr uint32_t OFF_MISC = OFF_WS + WIMG_BYTES;      // one staging buffer: a stage lasts several TMA latencies
constexpr uint32_t MISC = 4096;
constexpr uint32_t SMEM_BYTES = OFF_MISC + MISC;
static_assert(OFF_SLOT % 1024 == 0 && SLOT % 1024 == 0, "R images must keep the 512-byte swizzle phase");
static_assert(SMEM_BYTES <= 232448 && WIMG_BYTES % 16 == 0, "shared memory");

struct UArgs {
  NsfNetGeom g;
  const float* pk;       // FFMA packed image: layer 0, biases, output layer rows
  const float* wimg;     // (2L-1) stage images of WIMG_FLOATS: WF_1..WF_L, WB_{L-1}..WB_1 (rows j: hi[80] | lo[80])
  const float* x; const float* y; long long n;
  const float* e_in; const float* vtm_in; float* vtm_out; const float* w;
  float inv_Re, vis_t0, alpha_evm, cs1, cs2, k4, c_eq;
  int has_evm;
  float* resid_out; float* vis_t_out; float* ebar_out;
  float* stash;          // [grid][NS][L][P][KP][4]
  float* scratch;        // gradient rows [grid][gs_row]
  int n_pairs;
  long long* dbg;        // optional [grid][16 warps][16] cycle counters (nsf_get_stage_cycles)
};

struct Misc {
  uint64_t mbar[NS];     // "done": the slot's forward / dgrad MMAs have completed (tcgen05.commit)
  uint64_t ready[NS];    // the slot's operands are written and its previous results consumed (one arrival per epilogue warp)
  uint64_t wfull;        // the stage's weights are in tensor memory (one arrival per epilogue warp)
  uint64_t wsfull[2];    // staging buffer b holds a stage image (TMA complete_tx)
  uint64_t wsfree[2];    // every epilogue warp has copied staging buffer b into tensor memory
  uint64_t wdone[NS];    // the slot's weight-gradient MMAs (readers of its images) have completed (tcgen05.commit)
  uint32_t tmem_base;
  uint32_t pad[3];
  alignas(16) float ov[NS][P][16];   // outputs / output adjoints [slot][p][4*s + o]
  alignas(16) float red[NS * P][12]; // per-point-lane loss sums and output-bias gradient partials
};
static_assert(sizeof(Misc) <= MISC, "misc region too small");

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// shared-memory stores through 32-bit shared-window addresses (immediate offsets fold into the instruction)
__device__ __forceinline__ void sts4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__host__ __device__ constexpr uint32_t desc_hi_t(uint32_t sbo_bytes, uint32_t layout_type) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout_type << 29); }

// ---- issuer: the MMAs of stage s for one slot (executed by the whole issuer warp, see mma_tf32_elect2) ----
// Every descriptor is  sb4 + <compile-time constant>  with sb4 = (shared window base) >> 4 held in ONE uniform register
// (the slot loop is unrolled): the descriptor set-up between "operands ready" and the first MMA used to be ~40
// R2UR / uniform instructions (~170 cycles with the tensor pipe idle).  Shared addresses are < 2^18, so the 14-bit
// address field never carries into the LBO field.
__host__ __device__ constexpr uint32_t lbo_field(uint32_t lbo_bytes) { return ((lbo_bytes >> 4) & 0x3FFF) << 16; }

// forward (s <= L) or dgrad (s > L) MMAs of one slot: D[128, 4P] = W[128, 80] (tensor memory) * R[4P, 80]^T (shared memory).
// The tensor core truncates when it adds into the fp32 accumulator (measured: -2e-8 relative per MMA for same-sign
// sums): the 2^-11-sized correction products go first, while the accumulator is still small, the hi*hi products last.
__device__ __forceinline__ void issue_main(uint32_t sb4, uint32_t tmem, int slot, uint32_t leader, int half) {
  const uint32_t r4 = sb4 + (uint32_t)((OFF_SLOT + slot * SLOT) >> 4);
  const uint32_t d_col = tmem + TM_D + (uint32_t)(slot * NCOL);
  const uint32_t wa = tmem + TM_W;
  const uint32_t idesc = idesc_tf32(128, NCOL, 0, 1);
  constexpr uint32_t BHI = desc_hi_t(R_ATOM, 1);
  const uint32_t bh0 = r4 + lbo_field(1024), bl0 = bh0 + (RB >> 4);
  if (half == 0) {
#pragma unroll
    for (int ks = 0; ks < KP / 8; ++ks) {
      const uint32_t db = ks * ((2 * R_ATOM) >> 4);
      mma_tf32_ts_elect(d_col, wa + 80 + ks * 8, bh0 + db, BHI, idesc, ks > 0, leader);   // W_lo * a_hi
      mma_tf32_ts_elect(d_col, wa + ks * 8, bl0 + db, BHI, idesc, 1, leader);            // W_hi * a_lo
    }
  } else {
#pragma unroll
    for (int ks = 0; ks < KP / 8; ++ks) {
      const uint32_t db = ks * ((2 * R_ATOM) >> 4);
      mma_tf32_ts_elect(d_col, wa + ks * 8, bh0 + db, BHI, idesc, 1, leader);            // W_hi * a_hi
    }
  }
}
// wgrad: dW_l[128, 80] (+)= Zbar[128, 4P] * Act[80, 4P]^T, contraction over n = 4p + s.  Both operands are the row-per-neuron
// images the epilogue wrote for the MN-major reads, read K-MAJOR with the same descriptor layout type 1 (see nsf_umma_jet.cu).
// The accumulator of layer l is one of NDW rotating buffers; the first MMA of a tile group overwrites it.
__device__ __forceinline__ void issue_wgrad(uint32_t sb4, uint32_t tmem, int l, int slot, uint32_t leader) {
  const uint32_t r4 = sb4 + (uint32_t)((OFF_SLOT + slot * SLOT) >> 4);
  const uint32_t idesc = idesc_tf32(128, NW, 0, 0);
  const uint32_t dw_col = tmem + TM_DW + (uint32_t)((l % NDW) * NW);
  constexpr uint32_t RHI = desc_hi_t(R_ATOM, 1);
  const uint32_t ah0 = r4, al0 = r4 + (RB >> 4);
  const uint32_t bh0 = r4 + ((2 * RB) >> 4), bl0 = r4 + ((3 * RB) >> 4);
#pragma unroll
  for (int ks = 0; ks < NCOL / 8; ++ks) {
    const uint32_t d = ks * (32 >> 4);
    mma_tf32_elect2(dw_col, al0 + d, RHI, bh0 + d, RHI, idesc, !(slot == 0 && ks == 0), leader);
    mma_tf32_elect2(dw_col, ah0 + d, RHI, bl0 + d, RHI, idesc, 1, leader);
    mma_tf32_elect2(dw_col, ah0 + d, RHI, bh0 + d, RHI, idesc, 1, leader);
  }
}

// ---- epilogue helpers ---------------------------------------------------------------------------
struct Epi {
  int j, sub, q;
  bool active;           // j < KP
  uint32_t lane_addr;    // TMEM lane field of this warp's quadrant
  uint32_t r_off;        // byte offset of this thread's float4 (point pi = 0; pi = 1: + 16) in an R image
  uint32_t r_offA, r_offB; // first / second float4 store of a point pair (see store_R_pair)
  bool sw;               // this thread stores its point 1 first
};

__device__ __forceinline__ void split4(const float v[4], float hi[4], float lo[4]) {
#pragma unroll
  for (int s = 0; s < 4; ++s) split_tf32_fast(v[s], hi[s], lo[s]);
}
// the 4 streams of this thread's two points -> R image pair at `rimg` (hi; lo follows RB bytes later).  A point's float4
// lands in the 16-byte half (p & 1) of a 32-byte chunk whose position depends on (neuron & 3) only, so a plain
// "all lanes store point 0, then point 1" has lanes j and j+4 on the same banks (2-way conflict on every store).
// Lanes with (j >> 2) & 1 therefore store their point 1 first: the 8 lanes of a quarter-warp hit 8 different bank groups.
__device__ __forceinline__ void store_R_pair(uint32_t rimg, const Epi& e, const float v0[4], const float v1[4]) {
  float a[4], b[4], hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a[i] = e.sw ? v1[i] : v0[i]; b[i] = e.sw ? v0[i] : v1[i]; }
  split4(a, hi, lo);
  sts4(rimg + e.r_offA, hi[0], hi[1], hi[2], hi[3]);
  sts4(rimg + RB + e.r_offA, lo[0], lo[1], lo[2], lo[3]);
  split4(b, hi, lo);
  sts4(rimg + e.r_offB, hi[0], hi[1], hi[2], hi[3]);
  sts4(rimg + RB + e.r_offB, lo[0], lo[1], lo[2], lo[3]);
}

// tanh jet of one point: z -> activations
__device__ __forceinline__ void jet_fwd(const float z[4], float v[4]) {
  const float t = nsf_tanh_fast(z[0]);
  const float d1 = fmaf(-t, t, 1.f), d2 = -2.f * t * d1;
  v[0] = t; v[1] = d1 * z[1]; v[2] = d1 * z[2];
  v[3] = fmaf(d2, fmaf(z[1], z[1], z[2] * z[2]), d1 * z[3]);
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
  const float c = 2.f * ab[3] * d2;
  zb[3] = ab[3] * d1;
  zb[1] = fmaf(ab[1], d1, c * zx);
  zb[2] = fmaf(ab[2], d1, c * zy);
  zb[0] = fmaf(ab[0], d1, fmaf(ab[1] * d2, zx, fmaf(ab[2] * d2, zy, ab[3] * fmaf(d3, q, d2 * zl))));
}

// The weight-gradient accumulator of layer l (TMEM lane = j, columns TM_DW + (l % NDW)*80 + k) of the tile group just
// finished -> added to this CTA's gradient row (L2 resident).  A warp covers its 32 neurons for every fourth 8-column
// chunk (subs 0, 1: three chunks, subs 2, 3: two).  The additions are vector reductions (red.global.add.v4.f32): nothing is
// read back, and since every element of the row is only ever touched by ONE thread, in program order, the sums are exact
// and bit-reproducible (a read-modify-write cost 2400 cycles per flush, this costs 765).
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void flush_dw(const NsfNetGeom& g, float* grow, uint32_t tmem, const Epi& e, int l, bool first) {
  const uint32_t src = tmem + e.lane_addr + TM_DW + (uint32_t)((l % NDW) * NW) + (uint32_t)(e.sub * 8);
  float* dst = grow + g.gs_w(l) + (size_t)(e.active ? e.j : 0) * g.HP + e.sub * 8;
  const bool third = e.sub < 2;                      // chunk index sub + 8 exists only for sub < 2
  float v[3][8];
  tmem_ld8(src, v[0]);
  tmem_ld8(src + 32, v[1]);
  if (third) tmem_ld8(src + 64, v[2]);               // warp-uniform
  tmem_ld_wait();
  if (e.active) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (c < 2 || third) {
        float* d = dst + c * 32;
        if (first) {     // the row is this CTA's own: the first group overwrites, later groups add without reading (no L2 round trip)
          __stcg(reinterpret_cast<float4*>(d), make_float4(v[c][0], v[c][1], v[c][2], v[c][3]));
          __stcg(reinterpret_cast<float4*>(d) + 1, make_float4(v[c][4], v[c][5], v[c][6], v[c][7]));
        } else {
          red_add_v4(d, v[c][0], v[c][1], v[c][2], v[c][3]);
          red_add_v4(d + 4, v[c][4], v[c][5], v[c][6], v[c][7]);
        }
      }
    }
  }
}

// this thread's quarter of a stage image row (shared-memory staging buffer) -> the TMEM weight buffer
// (columns 40*sub .. 40*sub+39 of [hi | lo])
__device__ __forceinline__ void load_weights(const uint8_t* ws, const Epi& e, uint32_t tmem) {
  const float* src = reinterpret_cast<const float*>(ws) + (size_t)(e.active ? e.j : 0) * WROW + e.sub * 40;
  const uint32_t dst = tmem + e.lane_addr + TM_W + (uint32_t)(e.sub * 40);
  float4 v[10];
#pragma unroll
  for (int c = 0; c < 10; ++c) v[c] = *reinterpret_cast<const float4*>(src + c * 4);
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    const float w[8] = {v[2 * c].x, v[2 * c].y, v[2 * c].z, v[2 * c].w, v[2 * c + 1].x, v[2 * c + 1].y, v[2 * c + 1].z, v[2 * c + 1].w};
    tmem_st8(dst + c * 8, w);
  }
  tmem_st_wait();
}

template <int L, bool TRAIN, bool DBG>
__global__ void __launch_bounds__(NTHREADS, 1) nsf_umma_jet3_kernel(const UArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  Misc* misc = reinterpret_cast<Misc*>(smem + OFF_MISC);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const NsfNetGeom& g = a.g;
  constexpr int NSTAGE = TRAIN ? 2 * L : L + 1;     // stages per tile
  constexpr int NSTEPS = 2 * NSTAGE;                // steps per tile pair
  const uint32_t smem_base = smem_u32(smem);

  if (warp == 0) tmem_alloc(&misc->tmem_base, 512);
  if (tid == 0) {
    if (smem_base & 1023u) __trap();                // the swizzled R images assume a 1 KB aligned window
    for (int i = 0; i < NS; ++i) {
      mbar_init(&misc->mbar[i], 1); mbar_init(&misc->wdone[i], 1);
      mbar_init(&misc->ready[i], NEPI / 32);
    }
    mbar_init(&misc->wfull, NEPI / 32);
    for (int i = 0; i < 2; ++i) { mbar_init(&misc->wsfull[i], 1); mbar_init(&misc->wsfree[i], NEPI / 32); }
    mbar_fence_init();
  }
  // zero the operand slots once and the loss sums
  for (uint32_t i = tid * 16; i < NS * SLOT; i += NTHREADS * 16) *reinterpret_cast<float4*>(smem + OFF_SLOT + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < NS * P * 12) (&misc->red[0][0])[tid] = 0.f;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_base;

  const int my_pairs = ((int)blockIdx.x < a.n_pairs) ? (a.n_pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == ISSUER_WARP) {
    // =========================== issuer warp ===========================
    // No CTA-wide barrier in the steady state: the epilogue warps hand a slot's operands over through `ready` and the
    // stage's weights (tensor memory) through `wfull`; results come back through tcgen05.commit on `mbar` / `wdone`.
    // The whole warp walks the loop convergently with warp-uniform state; one elected lane issues the MMAs and commits.
    const uint32_t leader = elect_one();
    uint32_t rphases = 0;                  // bit = slot: parity to wait for
    uint32_t wphase = 0;
    bool pre_ok = false;                   // the next slot's `ready` phase was already seen complete
    long long icnt[5] = {0, 0, 0, 0, 0};   // weights wait, issue, operand wait, -, stage-slots
    for (int pr = 0; pr < my_pairs; ++pr) {
#pragma unroll 1
      for (int s = 1; s < NSTAGE; ++s) {
        long long t0 = 0, t1 = 0;
        if (DBG) t0 = clock64();
        mbar_wait(&misc->wfull, wphase);
        wphase ^= 1u;
        tc_fence_after();
        if (DBG) { t1 = clock64(); icnt[0] += t1 - t0; }
#pragma unroll
        for (int slot = 0; slot < NS; ++slot) {     // unrolled: the slot's operand descriptors are constants + smem_base
          if (DBG) t0 = clock64();
          const bool fenced = pre_ok;
          if (!pre_ok) mbar_wait(&misc->ready[slot], (rphases >> slot) & 1u);
          rphases ^= 1u << slot;
          if (!fenced) tc_fence_after();
          if (DBG) { t1 = clock64(); icnt[2] += t1 - t0; t0 = t1; }
          issue_main(smem_base >> 4, tmem, slot, leader, 0);
          {  // probe the next slot's operands while this slot's MMAs queue up
            const int nslot = (slot + 1) % NS;
            pre_ok = mbar_test_wait(&misc->ready[nslot], (rphases >> nslot) & 1u);
            if (pre_ok) tc_fence_after();
          }
          issue_main(smem_base >> 4, tmem, slot, leader, 1);
          mma_commit_elect(&misc->mbar[slot], leader);
          if (s > L) {
            issue_wgrad(smem_base >> 4, tmem, 2 * L - s, slot, leader);
            mma_commit_elect(&misc->wdone[slot], leader);
          }
          __syncwarp();
          if (DBG) { t1 = clock64(); icnt[1] += t1 - t0; icnt[4] += 1; }
        }
      }
    }
    if (DBG && lane == 0) {
#pragma unroll
      for (int i = 0; i < 5; ++i) a.dbg[((size_t)blockIdx.x * 16 + warp) * 16 + i] = icnt[i];
    }
  } else if (warp == PRODUCER_WARP) {
    // =========================== weight producer (one lane) ===========================
    // streams the per-stage weight images (rows hi | lo, 52 480 B) into the two staging buffers with one bulk TMA copy each
    if (lane == 0) {
      const long long total = (long long)my_pairs * (NSTAGE - 1);
      int img = 0;
      for (long long i = 0; i < total; ++i) {
        if (i >= 1) mbar_wait(&misc->wsfree[0], (uint32_t)((i - 1) & 1));
        mbar_expect_tx(&misc->wsfull[0], WIMG_BYTES);
        tma_bulk_g2s(smem + OFF_WS, a.wimg + (size_t)img * WIMG_FLOATS, WIMG_BYTES, &misc->wsfull[0]);
        if (++img == NSTAGE - 1) img = 0;
      }
    }
  } else if ((warp & 3) != 3) {
    // =========================== epilogue warps ===========================
    Epi e;
    e.q = warp & 3; e.sub = warp >> 2;
    e.j = e.q * 32 + lane;
    e.active = e.j < KP;
    e.lane_addr = (uint32_t)(e.q * 32) << 16;
    e.r_off = (uint32_t)(e.j >> 2) * R_ATOM + (uint32_t)(e.j & 3) * 128 + (uint32_t)((e.sub ^ (e.j & 3)) * 32);
    e.sw = ((e.j >> 2) & 1) != 0;
    e.r_offA = e.r_off + (e.sw ? 16u : 0u); e.r_offB = e.r_off + (e.sw ? 0u : 16u);
    const int jj = e.active ? e.j : 0;
    const float* pk = a.pk;
    const float w0x = __ldg(pk + g.pk_w0x() + jj), w0y = __ldg(pk + g.pk_w0y() + jj), b0 = __ldg(pk + g.pk_b0() + jj);
    const float wl0 = __ldg(pk + g.pk_wl() + jj), wl1 = __ldg(pk + g.pk_wl() + g.HP + jj), wl2 = __ldg(pk + g.pk_wl() + 2 * g.HP + jj);
    float bias[MAXL];      // b_l[j], l = 1..L-1
#pragma unroll
    for (int l = 1; l < MAXL; ++l) bias[l] = (l < L) ? __ldg(pk + g.pk_b(l) + jj) : 0.f;
    const float bo = (e.q == 0 && lane < 3) ? __ldg(pk + g.pk_bl() + lane) : 0.f;
    // this thread's float4 of (slot 0, layer 0, point 2*sub); + slot*L*P*KP + l*P*KP + pi*KP float4s
    float4* stash_thr = TRAIN ? reinterpret_cast<float4*>(a.stash) + (size_t)blockIdx.x * (NS * L * P * KP) + (size_t)(2 * e.sub) * KP + jj : nullptr;
    float* grow = TRAIN ? a.scratch + (size_t)blockIdx.x * g.gs_row() : nullptr;
    // per-thread gradient partials (this neuron, this thread's share of the points)
    float gw0x = 0.f, gw0y = 0.f, gwl[3] = {0.f, 0.f, 0.f};
    float gb[MAXL];
#pragma unroll
    for (int i = 0; i < MAXL; ++i) gb[i] = 0.f;
    uint32_t mphases = 0, wdphases = 0;                            // bit slot: parity to wait for on mbar[slot] / wdone[slot]
    const uint32_t d_base = tmem + e.lane_addr + TM_D + (uint32_t)(e.sub * (4 * PPT));
    bool more_groups = false, first_group = true;                  // set by the group loop below, read by stage_step
    uint32_t wstage = 0;                                           // MMA stages whose weights this warp has moved to tensor memory
    auto weights_to_tmem = [&]() {
      mbar_wait(&misc->wsfull[0], wstage & 1u);
      load_weights(smem + OFF_WS, e, tmem);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&misc->wfull); mbar_arrive(&misc->wsfree[0]); }
      ++wstage;
    };
    if (my_pairs > 0) weights_to_tmem();                           // weights of the first MMA stage
    const uint32_t slot0 = smem_base + OFF_SLOT;
    long long tcnt[10];      // [fwd | rev] x {MMA wait, work, fence, barrier, steps}
#pragma unroll
    for (int i = 0; i < 10; ++i) tcnt[i] = 0;

    // residuals, loss sums and adjoint seeds of point p of a slot's tile from its gathered outputs (one thread per point)
    auto residual_point = [&](const int slot, const int p, const long long p0, const int nvalid, const float pre_e, const float pre_vtm,
                              const float pre_w) {
      const bool ok = p < nvalid;
      const long long gp = p0 + p;
      float* ov = misc->ov[slot][p];
      float* red = misc->red[slot * P + p];
      const float u = ov[0], v = ov[1];
      const float ux = a.cs1 * ov[4], vx = a.cs1 * ov[5], px = a.cs1 * ov[6];
      const float uy = a.cs1 * ov[8], vy = a.cs1 * ov[9], py = a.cs1 * ov[10];
      const float ul = a.cs2 * ov[12], vl = a.cs2 * ov[13];
      float ee = 0.f, vis = 0.f;
      if (a.has_evm) {
        ee = pre_e;
        vis = ok ? fminf(a.vis_t0, pre_vtm) : a.vis_t0;
      }
      const float nu = a.inv_Re + vis;
      const float eq1 = (u * ux + v * uy) + px - nu * ul;
      const float eq2 = (u * vx + v * vy) + py - nu * vl;
      const float eq3 = ux + vy;
      const float eq4 = a.has_evm ? (eq1 * (u - 0.5f) + eq2 * (v - 0.5f)) - ee : 0.f;
      const float w = pre_w;
      if (ok) {
        red[0] += w * eq1 * eq1; red[1] += w * eq2 * eq2; red[2] += w * eq3 * eq3; red[3] += w * eq4 * eq4;
        red[4] += vis; red[5] += 1.f;
        if (a.resid_out) { a.resid_out[gp] = eq1; a.resid_out[a.n + gp] = eq2; a.resid_out[2 * a.n + gp] = eq3; a.resid_out[3 * a.n + gp] = eq4; }
        if (a.vis_t_out) a.vis_t_out[gp] = vis;
        if (a.has_evm && a.vtm_out) a.vtm_out[gp] = a.alpha_evm * fabsf(ee);
      }
      if (TRAIN) {
        const float cw = ok ? a.c_eq * w : 0.f;
        const float g1 = cw * (2.f * eq1 + a.k4 * eq4 * (u - 0.5f));
        const float g2 = cw * (2.f * eq2 + a.k4 * eq4 * (v - 0.5f));
        const float g3 = 2.f * cw * eq3;
        const float g4 = a.k4 * cw * eq4;
        ov[0] = g1 * ux + g2 * vx + g4 * eq1; ov[1] = g1 * uy + g2 * vy + g4 * eq2; ov[2] = 0.f;
        ov[4] = a.cs1 * (g1 * u + g3); ov[5] = a.cs1 * (g2 * u); ov[6] = a.cs1 * g1;
        ov[8] = a.cs1 * (g1 * v); ov[9] = a.cs1 * (g2 * v + g3); ov[10] = a.cs1 * g2;
        ov[12] = -a.cs2 * nu * g1; ov[13] = -a.cs2 * nu * g2; ov[14] = 0.f;
        red[6] += ov[0]; red[7] += ov[1]; red[8] += ov[2];
        if (a.ebar_out && ok) a.ebar_out[gp] = -g4;
      }
    };

    // one (stage, slot) step of this warp's share of a tile; `s` is a literal / unrolled constant at every call site
    // `part` splits the output stage (s = L): 1 = up to publishing the raw outputs, 2 = from the adjoint seeds on, 0 = all of it
    auto stage_step = [&](const int s, const int slot, const long long pA, const int part = 0) {
      const long long rem = a.n - pA - (long long)slot * P;
      const int nvalid = (int)(rem < 0 ? 0 : (rem < P ? rem : P));
      const long long p0 = pA + slot * P;
      const uint32_t sb = slot0 + (uint32_t)slot * SLOT;
      float4* st_slot = TRAIN ? stash_thr + (size_t)slot * (L * P * KP) : nullptr;
      const uint32_t d_addr = d_base + (uint32_t)(slot * NCOL);
      // reverse stages: fetch the stashed pre-activations (L2) BEFORE waiting for the MMAs of this stage
      const bool is_rev = TRAIN && s >= L;
      const int lrev = 2 * L - s - 1;                                   // layer whose tanh is differentiated (s = L: L-1)
      float4 st_l[PPT], st_lm1[PPT];
      if (is_rev && e.active) {
#pragma unroll
        for (int pi = 0; pi < PPT; ++pi) {
          if (part != 1) st_l[pi] = __ldcg(st_slot + lrev * (P * KP) + pi * KP);
          if (lrev >= 1 && part != 2) st_lm1[pi] = __ldcg(st_slot + (lrev - 1) * (P * KP) + pi * KP);
        }
      }
      float4 stv[PPT];                                                  // (t, zx, zy, z_lap) of a forward stage, stashed after the hand-over
      float xv[PPT], yv[PPT];
      if (s == 0 || (TRAIN && s == 2 * L - 1)) {
#pragma unroll
        for (int pi = 0; pi < PPT; ++pi) {
          const int p = e.sub * PPT + pi;
          xv[pi] = p < nvalid ? __ldg(a.x + p0 + p) : 0.f;
          yv[pi] = p < nvalid ? __ldg(a.y + p0 + p) : 0.f;
        }
      }
      // output stage: the residual threads' per-point inputs come from L2 as well; and the C image of a^{L-2} depends
      // on the stash only (its last readers finished with the previous tile), so it is written before the wait
      float pre_e = 0.f, pre_vtm = 0.f, pre_w = 1.f;
      if (s == L && part != 2) {
        if (part == 0 && tid < P && tid < nvalid) {
          const long long gp = p0 + tid;
          if (a.has_evm) { pre_e = __ldg(a.e_in + gp); pre_vtm = a.vtm_in ? __ldg(a.vtm_in + gp) : a.vis_t0; }
          if (a.w) pre_w = __ldg(a.w + gp);
        }
        if (TRAIN && L >= 2 && e.active) {
#pragma unroll
          float av[PPT][4];
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) act_from_stash(st_lm1[pi], av[pi]);
          store_R_pair(sb + 2 * RB, e, av[0], av[1]);
        }
      }
      const int cls = (s > L) ? 5 : 0;      // counter block: forward (incl. output stage) / reverse
      long long t0 = 0, t1 = 0;
      if (DBG) t0 = clock64();
      if (s >= 1 && !(s == L && part == 2)) {
        mbar_wait(&misc->mbar[slot], (mphases >> slot) & 1u);
        mphases ^= 1u << slot;
        tc_fence_after();
        // the last slot's MMAs of this stage have completed, and with them every reader of the weight buffer: bring in the
        // weights of the next MMA stage (stage 1 of the next tile group after the last stage)
        if (slot == NS - 1 && (s < NSTAGE - 1 || more_groups)) {
          long long tw = 0;
          if (DBG) tw = clock64();
          weights_to_tmem();
          if (DBG) tcnt[3] += clock64() - tw;
        }
      }
      if (DBG) { t1 = clock64(); tcnt[cls + 0] += t1 - t0; }

      if (s == 0) {
        // ---- layer 0 (K = 2) -------------------------------------------------------------
        if (e.active) {
#pragma unroll
          float v[PPT][4];
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            const float z[4] = {fmaf(w0x, xv[pi], fmaf(w0y, yv[pi], b0)), w0x, w0y, 0.f};
            jet_fwd(z, v[pi]);
            stv[pi] = make_float4(v[pi][0], z[1], z[2], z[3]);
          }
          store_R_pair(sb, e, v[0], v[1]);
        }
      } else if (s < L) {
        // ---- hidden layer s forward --------------------------------------------------------
        float z[PPT][4];
        tmem_ld8(d_addr, &z[0][0]);
        tmem_ld_wait();
        if (e.active) {
#pragma unroll
          float v[PPT][4];
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            z[pi][0] += bias[s];
            jet_fwd(z[pi], v[pi]);
            stv[pi] = make_float4(v[pi][0], z[pi][1], z[pi][2], z[pi][3]);
          }
          store_R_pair(sb, e, v[0], v[1]);
        }
      } else if (s == L) {
        // ---- output layer: gather, residuals, adjoint seeds ---------------------------------
        if (part != 2) {
        float o[PPT][4];
        tmem_ld8(d_addr, &o[0][0]);
        tmem_ld_wait();
        if (e.q == 0 && lane < 3) {
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            const int p = e.sub * PPT + pi;
            misc->ov[slot][p][0 * 4 + lane] = o[pi][0] + bo;
            misc->ov[slot][p][1 * 4 + lane] = o[pi][1];
            misc->ov[slot][p][2 * 4 + lane] = o[pi][2];
            misc->ov[slot][p][3 * 4 + lane] = o[pi][3];
          }
        }
        }
        if (part == 0) {
          epi_bar();
          if (tid < P) residual_point(slot, tid, p0, nvalid, pre_e, pre_vtm, pre_w);
          if (TRAIN) epi_bar();
        }
        if (TRAIN && part != 1) {
          if (e.active) {
            float sb0 = 0.f;
            float zb2[PPT][4];
#pragma unroll
            for (int pi = 0; pi < PPT; ++pi) {
              const float4* ov4 = reinterpret_cast<const float4*>(misc->ov[slot][e.sub * PPT + pi]);
              float ab[4], act[4];
              float* zb = zb2[pi];
              float4 ovs[4];
#pragma unroll
              for (int st = 0; st < 4; ++st) {
                ovs[st] = ov4[st];
                ab[st] = fmaf(ovs[st].x, wl0, fmaf(ovs[st].y, wl1, ovs[st].z * wl2));
              }
              zbar_from(st_l[pi], ab, zb);
              act_from_stash(st_l[pi], act);
#pragma unroll
              for (int st = 0; st < 4; ++st) {
                gwl[0] = fmaf(ovs[st].x, act[st], gwl[0]);
                gwl[1] = fmaf(ovs[st].y, act[st], gwl[1]);
                gwl[2] = fmaf(ovs[st].z, act[st], gwl[2]);
              }
              sb0 += zb[0];
            }
            if (L >= 2) store_R_pair(sb, e, zb2[0], zb2[1]);
            gb[L - 1] += sb0;
          }
        }
      } else {
        // ---- reverse: D holds the adjoint of layer l's activations, l = 2L - s - 1 ------------------
        const int l = (2 * L - s - 1) >= 0 ? (2 * L - s - 1) : 0;
        float ab[PPT][4];
        tmem_ld8(d_addr, &ab[0][0]);
        tmem_ld_wait();
        if (e.active) {
          float sb0 = 0.f;
          float zb2[PPT][4];
#pragma unroll
          for (int pi = 0; pi < PPT; ++pi) {
            zbar_from(st_l[pi], ab[pi], zb2[pi]);
            sb0 += zb2[pi][0];
            if (l == 0) { gw0x += fmaf(zb2[pi][0], xv[pi], zb2[pi][1]); gw0y += fmaf(zb2[pi][0], yv[pi], zb2[pi][2]); }
          }
          if (l >= 1) {
            float av[PPT][4];
#pragma unroll
            for (int pi = 0; pi < PPT; ++pi) act_from_stash(st_lm1[pi], av[pi]);
            // the weight-gradient MMAs of this stage still read both image pairs of the slot: wait for them
            mbar_wait(&misc->wdone[slot], (wdphases >> slot) & 1u);
            store_R_pair(sb, e, zb2[0], zb2[1]);
            store_R_pair(sb + 2 * RB, e, av[0], av[1]);
          }
          gb[l] += sb0;
        }
        if (l == 0 || !e.active) mbar_wait(&misc->wdone[slot], (wdphases >> slot) & 1u);   // keeps the parity; precedes flush_dw
        wdphases ^= 1u << slot;
      }
      if (DBG) { t0 = clock64(); tcnt[cls + 1] += t0 - t1; }
      if (s < NSTAGE - 1 && !(s == L && part == 1)) {
        // hand the slot to the issuer: operands visible to the async proxy, TMEM reads retired
        fence_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&misc->ready[slot]);
      }
      if (TRAIN && s > L && slot == NS - 1) {
        // the weight-gradient MMAs of all slots of this stage have completed (this warp waited for the last slot's above):
        // this group's dW_{2L-s} -> the CTA's gradient row, after the hand-over so that the next MMAs are not held up.  The
        // accumulator's next writer (layer l - 3, three stages on) is ordered behind this warp's later hand-overs.
        __syncwarp();
        tc_fence_after();
        long long tw = 0;
        if (DBG) tw = clock64();
        flush_dw(g, grow, tmem, e, 2 * L - s, first_group);
        if (DBG) tcnt[8] += clock64() - tw;
      }
      if (TRAIN && s < L && e.active) {
        // the stash stores go to L2 after the hand-over: the proxy fence above would wait for them
#pragma unroll
        for (int pi = 0; pi < PPT; ++pi) __stcg(st_slot + s * (P * KP) + pi * KP, stv[pi]);
      }
      if (DBG) { t1 = clock64(); tcnt[cls + 2] += t1 - t0; tcnt[cls + 4] += 1; }
    };

    const long long pair_stride = (long long)gridDim.x * NS * P;
    long long pA = (long long)blockIdx.x * NS * P;                        // first point of slot 0's tile; the other slots follow
    if (my_pairs > 0) {
#pragma unroll 1
      for (int slot = 0; slot < NS; ++slot) stage_step(0, slot, pA);
    }
    for (int pr = 0; pr < my_pairs; ++pr, pA += pair_stride) {
      more_groups = pr + 1 < my_pairs;
      first_group = pr == 0;
#pragma unroll
      for (int s = 1; s < NSTAGE - 1; ++s) {
        if (s == L) {
          // output stage of all slots at once: one residual thread per point of the group (NS x P), two CTA barriers per
          // group instead of two per slot; its epilogue is what the tensor pipe waits for in this part of the tile
          const int rs = tid / P, rp = tid % P;
          const long long rp0 = pA + (long long)rs * P;
          const long long rrem = a.n - rp0;
          const int rnv = (int)(rrem < 0 ? 0 : (rrem < P ? rrem : P));
          float pre_e = 0.f, pre_vtm = 0.f, pre_w = 1.f;
          if (tid < NS * P && rp < rnv) {
            if (a.has_evm) { pre_e = __ldg(a.e_in + rp0 + rp); pre_vtm = a.vtm_in ? __ldg(a.vtm_in + rp0 + rp) : a.vis_t0; }
            if (a.w) pre_w = __ldg(a.w + rp0 + rp);
          }
#pragma unroll 1
          for (int slot = 0; slot < NS; ++slot) stage_step(s, slot, pA, 1);
          epi_bar();
          if (tid < NS * P) residual_point(rs, rp, rp0, rnv, pre_e, pre_vtm, pre_w);
          epi_bar();
#pragma unroll 1
          for (int slot = 0; slot < NS; ++slot) stage_step(s, slot, pA, 2);
        } else {
#pragma unroll 1
          for (int slot = 0; slot < NS; ++slot) stage_step(s, slot, pA);
        }
      }
      // last stage of this pair fused with stage 0 of the next pair, slot by slot: the issuer gets slot A's first
      // operands of the next tile while slot B still finishes, so the tensor pipe does not drain between pairs
#pragma unroll 1
      for (int slot = 0; slot < NS; ++slot) {
        stage_step(NSTAGE - 1, slot, pA);
        if (pr + 1 < my_pairs) stage_step(0, slot, pA + pair_stride);
      }
    }

    if (DBG && lane == 0) {
#pragma unroll
      for (int i = 0; i < 10; ++i) a.dbg[((size_t)blockIdx.x * 16 + warp) * 16 + i] = tcnt[i];
    }
    // ---- CTA epilogue: thread-local gradient partials and loss sums -> this CTA's row -------------
    if (a.scratch) {
      float* growx = a.scratch + (size_t)blockIdx.x * g.gs_row();
      float* redf = reinterpret_cast<float*>(smem + OFF_SLOT);   // operand slots are free now: [NSUB-1][KP][16]
      if (TRAIN && e.sub >= 1 && e.active) {
        float* r = redf + ((e.sub - 1) * KP + e.j) * 16;
        r[0] = gw0x; r[1] = gw0y; r[2] = gwl[0]; r[3] = gwl[1]; r[4] = gwl[2];
#pragma unroll
        for (int i = 0; i < MAXL; ++i) r[5 + i] = gb[i];
      }
      epi_bar();
      if (TRAIN && e.sub == 0 && e.active) {
        float r[5 + MAXL];
#pragma unroll
        for (int i = 0; i < 5 + MAXL; ++i) {
          r[i] = 0.f;
#pragma unroll
          for (int q = 0; q < NSUB - 1; ++q) r[i] += redf[(q * KP + e.j) * 16 + i];
        }
        const int j = e.j;
        growx[g.gs_w0x() + j] = gw0x + r[0];
        growx[g.gs_w0y() + j] = gw0y + r[1];
        growx[g.gs_b0() + j] = gb[0] + r[5];
        growx[g.gs_wl() + j] = gwl[0] + r[2];
        growx[g.gs_wl() + g.HP + j] = gwl[1] + r[3];
        growx[g.gs_wl() + 2 * g.HP + j] = gwl[2] + r[4];
        growx[g.gs_wl() + 3 * g.HP + j] = 0.f;
#pragma unroll
        for (int l = 1; l < MAXL; ++l)
          if (l < L) growx[g.gs_b(l) + j] = gb[l] + r[5 + l];
      }
      if (tid < NSF_LOSS_SLOTS) {
        float v = 0.f;
        if (tid < 6) for (int p = 0; p < NS * P; ++p) v += misc->red[p][tid];
        growx[g.gs_loss() + tid] = v;
      }
      if (TRAIN && tid < 4) {
        float v = 0.f;
        if (tid < 3) for (int p = 0; p < NS * P; ++p) v += misc->red[p][6 + tid];
        growx[g.gs_bl() + tid] = v;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// stage images [hi(80) | lo(80)] per neuron row: thread per (image, row r, contraction index c)
__global__ void nsf_umma3_pack_kernel(NsfNetGeom g, const float* __restrict__ flat, float* __restrict__ wimg) {
  const int L = g.L, H = g.H;
  const int n_img = 2 * L - 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_img * KP * KP) return;
  const int img = (int)(idx / (KP * KP)), r = (int)(idx % (KP * KP)) / KP, c = (int)(idx % KP);
  float v = 0.f;
  if (img < L - 1) {                 // WF_l, l = img + 1: W_l[r][c]
    const int l = img + 1, fo = 3 * H + (l - 1) * (H * H + H);
    if (r < H && c < H) v = flat[fo + r * H + c];
  } else if (img == L - 1) {         // output layer rows o < n_out
    const int fo = 3 * H + (L - 1) * (H * H + H);
    if (r < g.n_out && c < H) v = flat[fo + r * H + c];
  } else {                           // WB_l, l = 2L - 1 - img: W_l^T[r][c] = W_l[c][r]
    const int l = 2 * L - 1 - img, fo = 3 * H + (l - 1) * (H * H + H);
    if (r < H && c < H) v = flat[fo + c * H + r];
  }
  float hi, lo;
  split_tf32(v, hi, lo);
  float* row = wimg + (size_t)img * WIMG_FLOATS + (size_t)r * WROW;
  row[c] = hi;
  row[80 + c] = lo;
}

struct Umma3State {
  float* wimg = nullptr;
  float* stash = nullptr;
  long long* dbg = nullptr;
  int dbg_on = 0, last_grid = 0;
  int grid = 0;
};

}  // namespace

int nsf_umma3_group_points() { return NS * P; }

typedef void (*Jet3Kernel)(const UArgs);
template <int L>
static Jet3Kernel jet3_kernel_of(bool train, bool dbg) {
  if (!train) return nsf_umma_jet3_kernel<L, false, false>;
  return dbg ? nsf_umma_jet3_kernel<L, true, true> : nsf_umma_jet3_kernel<L, true, false>;
}
static Jet3Kernel jet3_kernel(int L, bool train, bool dbg = false) {
  switch (L) {
    case 2: return jet3_kernel_of<2>(train, dbg);
    case 3: return jet3_kernel_of<3>(train, dbg);
    case 4: return jet3_kernel_of<4>(train, dbg);
    case 5: return jet3_kernel_of<5>(train, dbg);
    default: return jet3_kernel_of<6>(train, dbg);
  }
}

int nsf_umma3_init(NsfCtx* ctx) {
  if (ctx->umma3) return NSF_OK;
  if (!nsf_umma_supported(ctx->main.g)) { nsf_set_error("tcgen05 path covers hidden = 80, 2..6 hidden layers"); return NSF_E_SHAPE; }
  Umma3State* s = new Umma3State();
  const NsfNetGeom& g = ctx->main.g;
  s->grid = ctx->sms;
  if (s->grid > ctx->main.rows) s->grid = ctx->main.rows;
  NSF_CUDA_OK(cudaMalloc((void**)&s->wimg, (size_t)(2 * g.L - 1) * WIMG_FLOATS * sizeof(float)));
  NSF_CUDA_OK(cudaMalloc((void**)&s->stash, (size_t)s->grid * NS * g.L * P * KP * 4 * sizeof(float)));
  for (int train = 0; train < 2; ++train)
    for (int dbg = 0; dbg <= train; ++dbg)
      NSF_CUDA_OK(cudaFuncSetAttribute(jet3_kernel(g.L, train != 0, dbg != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  ctx->ws_bytes += (long long)(2 * g.L - 1) * WIMG_FLOATS * 4 + (long long)s->grid * NS * g.L * P * KP * 16;
  ctx->umma3 = s;
  return NSF_OK;
}

void nsf_umma3_free(NsfCtx* ctx) {
  Umma3State* s = (Umma3State*)ctx->umma3;
  if (!s) return;
  cudaFree(s->wimg); cudaFree(s->stash); if (s->dbg) cudaFree(s->dbg);
  delete s;
  ctx->umma3 = nullptr;
}

// Collocation jet step / residuals on the tcgen05 path with the weights in tensor memory.  Returns the grid (rows written)
// through *grid_out.
int nsf_umma3_launch(NsfCtx* ctx, const NsfKernelArgs& k, const float* flat_params, int* grid_out, nsf_stream_t st, int* launches) {
  Umma3State* s = (Umma3State*)ctx->umma3;
  const NsfNetGeom& g = ctx->main.g;
  const long long tot = (long long)(2 * g.L - 1) * KP * KP;
  nsf_umma3_pack_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, flat_params, s->wimg);
  NSF_CUDA_OK(cudaGetLastError());
  ++*launches;
  UArgs a;
  a.g = g; a.pk = k.pk; a.wimg = s->wimg; a.x = k.x; a.y = k.y; a.n = k.n;
  const bool train = k.mode == NSF_MODE_JET_STEP;
  a.e_in = k.e_in; a.vtm_in = k.vtm_in; a.vtm_out = k.vtm_out; a.w = k.w;
  a.inv_Re = k.inv_Re; a.vis_t0 = k.vis_t0; a.alpha_evm = k.alpha_evm; a.cs1 = k.cs1; a.cs2 = k.cs2; a.k4 = k.k4; a.c_eq = k.c_eq;
  a.has_evm = k.has_evm;
  a.resid_out = k.resid_out; a.vis_t_out = k.vis_t_out; a.ebar_out = k.ebar_out;
  a.stash = train ? s->stash : nullptr;
  a.scratch = train ? k.scratch : nullptr;
  a.n_pairs = (int)((k.n + NS * P - 1) / (NS * P));
  a.dbg = (s->dbg_on && train) ? s->dbg : nullptr;
  int grid = a.n_pairs < s->grid ? a.n_pairs : s->grid;
  if (grid <= 0) { *grid_out = 0; return NSF_OK; }
  jet3_kernel(g.L, train, a.dbg != nullptr)<<<grid, NTHREADS, SMEM_BYTES, st>>>(a);
  NSF_CUDA_OK(cudaGetLastError());
  ++*launches;
  s->last_grid = grid;
  *grid_out = grid;
  return NSF_OK;
}

// Diagnostics (nsf_get_stage_cycles): per-warp cycle counters of the last launch averaged over CTAs, out[w*16 + k]:
// epilogue warps {fwd: MMA wait (incl. weight load), work, fence, weight load, steps; rev: MMA wait, work, fence, dW flush, steps};
// issuer warp {weights wait, issue, operand wait, -, stage-slots}.
int nsf_umma3_stage_cycles(NsfCtx* ctx, double* out) {
  int rc = nsf_umma3_init(ctx);
  if (rc != NSF_OK) return rc;
  Umma3State* s = (Umma3State*)ctx->umma3;
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
