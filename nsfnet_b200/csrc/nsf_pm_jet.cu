// tcgen05 jet kernel, "points on M": the collocation step (jet forward + residuals + reverse pass) of a 2 -> H x L -> 3 tanh
// MLP with every hidden-layer contraction on the 5th-generation tensor cores, 3xTF32 split for fp32 fidelity.
//   hidden = 80  (ev-NSFnet main net, L <= 6):  tiles of 32 points, MMA M = 128
//   hidden = 120 (NSFnet, L <= 4)            :  tiles of 16 points, MMA M = 64
//
// Orientation (round 2; the round-1 kernel had the neurons on M and 8 points x 4 streams on N = 32, which left every MMA
// bound by the re-fetch of its 4 KB weight operand: 43 cycles for 16 cycles of math):
//   MMA row   m = 4 * point + stream      (streams value, d/dx, d/dy, laplacian; M = 128 rows = 32 points)
//   MMA col   n = output neuron           (N = H)
//   forward / dgrad :  D[m, n] = sum_k A[m, k] * Wt[n, k]      A = activations / adjoints of the tile, B = the layer's weights
//   weight gradient :  dW_l[j, k] += sum_m Zbar[m, j] * Act[m, k]   (contraction over the tile's rows; the accumulator lives for ONE tile and
//                                                                   layer and is then added to the CTA's gradient row in L2: numerics, below)
// One MMA now carries 32 points for 62 cycles (measured, scripts/probe_pm.cu) instead of 8 points for 43.
//
// Operand images of a tile (shared memory):  P (activations a_l going forward, adjoints zbar_l going back) and Q (a_{l-1}, the
// second operand of the weight gradient), each [hi | lo], rows of 128 bytes = 32 tile rows of ONE neuron, atoms of 4 neurons,
// 32-byte chunks XOR-swizzled by (neuron & 3): tcgen05 descriptor layout type 1.  The same bytes are read MN-major as the A
// operand of forward / dgrad and K-major as both operands of the weight gradient (round-1 finding, re-checked by the probes).
// A thread stores the 4 streams of a (point, neuron) with ONE st.shared.v4.
// Weights: per MMA stage [hi | lo] x NG blocks of 5 k-steps, K-major, streamed from L2 by TMA bulk copies (measured: > 50 B/clk/SM
// with every SM streaming) into NG "hi" slots (resident for the stage: read by the correction pass and by the main pass) and
// two rotating "lo" slots.
//
// Warps: 4*NSUB epilogue warps (a warp may only touch TMEM lanes of quadrant warp & 3: 32 rows = 8 points x 4 streams), one
// issuer warp (tcgen05.mma / commit, one elected lane), one weight producer (one lane).  One tile is in flight per CTA:
//   stage 0        : layer 0 (K = 2) on FFMA                                        -> P
//   stage 1..L-1   : hidden layer forward, MMA -> D, epilogue: tanh jet             -> P (in place), stash (t, zx, zy, z_lap) to L2
//   stage L        : output layer (N = 16), residuals, loss sums, adjoint seeds, output-layer dgrad / wgrad on FFMA -> P, Q
//   stage L+1..2L-1: layer l = 2L - s: dgrad MMA -> D, weight-gradient MMAs; epilogue: adjoint through tanh -> P, Q
// The epilogue thread reads its TMEM lane (= row = point, stream) for its worker's 20 adjacent columns (one x16 + one x4 load per
// accumulator) and a 4x4 shuffle transpose inside the quad of lanes hands every lane the 4 streams of ONE (point, neuron).  In the reverse stages the epilogue computes zbar while the
// weight-gradient MMAs still read P and Q, parks it in the (free) D columns of tensor memory, and stores it once the
// weight-gradient MMAs of ITS rows have completed (wdone[row group]).
#include "nsf_internal.h"
#include "nsf_tc.cuh"
#include "nsf_jet_math.cuh"
#include <cstdlib>
#include <type_traits>
#include <vector>

using namespace nsftc;

namespace {

// The tensor core rounds TOWARD ZERO every time it adds into the fp32 accumulator (oracle/tc_model.py reproduces raw tcgen05.mma
// results bit for bit: every term truncated at 2^(e_max - 25), the sum rounded toward zero).  The bias does not average out in a gradient that is a sum of cancelling terms: at
// trained weights a single K = 80 / 120 chain put the weight gradient 3 - 6x further from the fp64 truth than the reference's own
// fp32 path (profiles/r2_parity_trained_before_fix.txt).  Hence
//   * the full-magnitude hi * hi products of a forward contraction are spread over NACC_F accumulators (chains of 2 - 4 MMAs), summed
//     by the epilogue in fp32 round-to-nearest; the 2^-11-sized corrections go into the last one before its own hi * hi products;
//   * a weight-gradient accumulator lives for ONE tile and one layer: it is added to the CTA's gradient row (red.global.add, L2)
//     as soon as its MMAs have completed.
// Measured with the bit-exact model of the accumulation (scripts/emu_tc_numerics.py): the bias of the FORWARD chains is what
// moves the gradient; that of the dgrad chains does not (a relative 1e-7 on the adjoints).  Forward stages therefore use NACC_F
// accumulators (the weight-gradient accumulator is free then: all 512 columns), reverse stages NACC_R.  hidden = 120 at Re = 1000 needs
// chains of <= 4 MMAs (K = 120: 4 accumulators) to stay within 2x the reference's own fp32 error; hidden = 80 has margin with 3.
#ifndef NSF_PM_NACC_F128
#define NSF_PM_NACC_F128 3
#endif
#ifndef NSF_PM_NACC_R128
#define NSF_PM_NACC_R128 1
#endif
#ifndef NSF_PM_NACC_F64
#define NSF_PM_NACC_F64 4
#endif
#ifndef NSF_PM_NACC_R64
#define NSF_PM_NACC_R64 1
#endif
// Hand-over pattern: bit c set = the epilogue hands the operand image to the issuing warp after chunk c (the last chunk always).  Compile-time:
// as a kernel argument the per-chunk branches cost 2 % of the kernel (6.79 -> 6.68 ms at hidden = 80).
#ifndef NSF_PM_HO128
#define NSF_PM_HO128 0x19
#endif
#ifndef NSF_PM_HO64
#define NSF_PM_HO64 0x1b
#endif
template <int H_, int MT_>
struct Cfg {
  static constexpr int H = H_, MT = MT_;
  static constexpr int PTS = MT / 4;             // points per tile
  static constexpr int NQ = MT / 32;             // 32-row groups of an operand image
  static constexpr int KS = H / 8;               // k-steps of a layer contraction
  static constexpr int GK = 5;                   // k-steps per weight block
  static constexpr int NG = KS / GK;             // weight blocks per part (hi / lo) and stage
  static constexpr int NB = H;                   // N of the forward / dgrad MMAs
  static constexpr int DWN = ((H + 15) / 16) * 16;  // N of the weight-gradient MMAs (M = 128 needs N % 16 == 0)
  static constexpr int NACC_F = MT == 128 ? NSF_PM_NACC_F128 : NSF_PM_NACC_F64;   // accumulators of a forward contraction (columns [0, NACC_F * NB))
  static constexpr int NACC_R = MT == 128 ? NSF_PM_NACC_R128 : NSF_PM_NACC_R64;   // accumulators of a dgrad contraction (behind the weight gradient's)
  static constexpr int NSUB = MT == 128 ? 4 : 3; // epilogue warps per TMEM quadrant
  static constexpr int NEW = 4 * NSUB;           // epilogue warps
  static constexpr int NWORK = MT == 128 ? NSUB : 2 * NSUB;  // workers (warps resp. half-warps) per quadrant
  static constexpr int CW = 4 * NWORK;           // neurons per chunk: chunk c = neurons [CW c, CW (c+1)), 4 per worker
  static constexpr int NCH = H / CW;             // chunks per stage (= 4-neuron pieces per worker)
  static constexpr int KPC = CW / 8;             // k-steps of the next contraction that one chunk completes
  static constexpr int WCOLS = 4 * NCH;          // D columns of one worker: its neurons of every chunk side by side, so that one wide
                                                 // tcgen05.ld fetches them all.  Neuron n = CW c + 4 w + i  <->  column WCOLS w + 4 c + i
                                                 // (the row order of the weight images: nsf_pm_pack_kernel)
  __host__ __device__ static constexpr int dcol(int n) { return WCOLS * ((n % CW) / 4) + 4 * (n / CW) + (n % 4); }
  static constexpr int HO = MT == 128 ? NSF_PM_HO128 : NSF_PM_HO64;        // (measured over all patterns, scripts/build_variants.py: hidden 80 best with 0x19 = chunks 0, 3, 4; hidden 120 with 0x1b = 0, 1, 3, 4)
  static constexpr int NEPI = NEW * 32;
  static constexpr int NTHREADS = (NEW + 2) * 32;
  static constexpr uint32_t GRP = (H / 4) * 512; // one 32-row group of one image part
  static constexpr uint32_t PART = NQ * GRP;     // hi or lo
  static constexpr uint32_t IMG = 2 * PART;
  static constexpr uint32_t WSUB = NB * 32;      // one k-step of one weight part
  static constexpr uint32_t WBLK = GK * WSUB;
  static constexpr uint32_t WSUB_O = 16 * 32;    // output layer: 16 rows (3 real)
  static constexpr uint32_t WBLK_O = GK * WSUB_O;
  static constexpr uint32_t WSTAGE = 2 * NG * WBLK;   // one stage image in global memory: [hi blocks | lo blocks]
  static constexpr uint32_t OFF_P = 0, OFF_Q = IMG, OFF_WHI = 2 * IMG, OFF_WLO = OFF_WHI + NG * WBLK, OFF_MISC = OFF_WLO + 2 * WBLK;
  static constexpr bool GWL_SMEM = MT == 128;   // 18 warps leave 96 registers per thread: the output-layer gradient goes to shared memory
  static constexpr uint32_t MISC = MT == 128 ? 15360 : 12288;
  static constexpr uint32_t SMEM_BYTES = OFF_MISC + MISC;
  static_assert(H % 8 == 0 && KS % GK == 0 && H % CW == 0 && CW % 8 == 0 && NCH <= 5, "shape");
  static_assert((HO >> (NCH - 1)) & 1, "the last chunk is always handed over");
  // Accumulator of k-step ks' hi * hi product when na accumulators share a contraction: the LAST accumulator (the corrections') takes exactly
  // the last chunk's k-steps, so that its own hi * hi products follow every correction; the others share the rest evenly.
  __host__ __device__ static constexpr int acc_of(int ks, int na) { return (na == 1 || ks >= KS - KPC) ? na - 1 : ks * (na - 1) / (KS - KPC); }
  static_assert(SMEM_BYTES <= 232448, "shared memory");
  static_assert(OFF_Q % 1024 == 0 && OFF_WHI % 1024 == 0 && PART % 1024 == 0, "swizzle phase");
};

// Position of dW_l[j][k] inside the H * H block of a gradient row written by this kernel: the order in which the epilogue
// drains the accumulator (TMEM lane = j, 4 columns per instruction), so that the additions are coalesced.
__host__ __device__ inline int pm_dw_index(int H, int nsub, int j, int k) {
  const int cpw = H / nsub, q = j >> 5, lane = j & 31, nl = H - 32 * q < 32 ? H - 32 * q : 32;
  return q * 32 * H + ((k / cpw) * (cpw / 4) + ((k % cpw) >> 2)) * (nl * 4) + lane * 4 + (k & 3);
}

struct PArgs {
  NsfNetGeom g;
  const float* pk;       // FFMA packed image: layer 0, biases, output layer rows
  const uint8_t* wimg;   // (2L-1) stage images: WF_1..WF_{L-1}, W_out, WB_{L-1}..WB_1
  const float* x; const float* y; long long n;
  const float* e_in; const float* vtm_in; float* vtm_out; const float* w;
  float inv_Re, vis_t0, alpha_evm, cs1, cs2, k4, c_eq;
  int has_evm;
  float* resid_out; float* vis_t_out; float* ebar_out;
  float* stash;          // [grid][L][PTS][H] float4
  float* scratch;        // gradient rows [grid][gs_row]
  int n_tiles;
  int zero;              // 0 at run time (keeps the issuer's descriptors loop-variant AND warp-uniform, see the issuer)
  long long* dbg;        // optional [grid][32] cycle counters
  int dbg_wa, dbg_wb;    // the two epilogue warps whose counters are recorded
};

template <int H, int L, int NGWL>
struct Misc {
  uint64_t ready[5];     // chunk c of the next MMA stage's operands written, previous results consumed (one arrival per epilogue warp)
  uint64_t dfull;        // forward / dgrad MMAs of the stage complete
  uint64_t wdone[4];     // weight-gradient MMAs over row group q complete (P / Q rows of the group may be rewritten)
  uint64_t dwfree;       // the weight-gradient accumulator has been drained to the gradient row (one arrival per epilogue warp)
  uint64_t hi_full[3], hi_free[3], lo_full[2], lo_free[2];
  uint32_t tmem_base, pad;
  float loss[4][12];     // per TMEM quadrant: w eq1^2, w eq2^2, w eq3^2, w eq4^2, vis_t, count, gbL[0..2]
  float gb[4][L][H];     // bias gradients, one private copy per quadrant (single owner lane per entry: deterministic)
  float gw0x[4][H], gw0y[4][H];
  float gwl[NGWL ? NGWL : 1][3][NGWL ? H : 1];   // output-layer weight gradient per quadrant (NGWL = 4), or kept in registers (NGWL = 0)
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sts4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// The activation stash (36 MB) and the gradient rows (20 MB) are rewritten / re-read all the time and fit the 126 MB L2; under the
// default replacement policy their dirty lines are still written back as they age: 2.6 - 2.8 GB of DRAM writes per 1e6-point launch of
// the hidden = 80 kernel (ncu, profiles/r2_pm_ev_1M_ncu_summary.txt; 5 % of the HBM bandwidth, DRAM READS stay at 20 MB: nothing is
// fetched twice).  An L2 evict_last policy on these accesses removes the write-backs (33 MB per launch) but costs 5 % of the kernel time
// (8.13 against 7.75 ms; policy on the reductions alone: the same) -- measured, profiles/r2_pm_l2_policy.txt -- so it is off by default.
__device__ __forceinline__ uint64_t l2_keep_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#ifndef NSF_PM_KEEP
#define NSF_PM_KEEP 0     // 1: stash stores, 2: stash loads, 4: gradient-row reductions carry the policy
#endif
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d, uint64_t pol) {
  if (NSF_PM_KEEP & 4) asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d), "l"(pol) : "memory");
  else asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_keep(float4* addr, const float4 v, uint64_t pol) {
  if (NSF_PM_KEEP & 1) asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
  else __stcg(addr, v);
}
__device__ __forceinline__ float4 ld_keep(const float4* addr, uint64_t pol) {
  if (!(NSF_PM_KEEP & 2)) return __ldcg(addr);
  float4 v;
  asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(addr), "l"(pol) : "memory");
  return v;
}
__host__ __device__ constexpr uint32_t desc_hi_t(uint32_t sbo_bytes, uint32_t layout_type) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | (layout_type << 29); }
__host__ __device__ constexpr uint32_t lbo_field(uint32_t lbo_bytes) { return ((lbo_bytes >> 4) & 0x3FFF) << 16; }

// TMEM <-> registers: this thread's lane (row), 4 consecutive columns.  MT = 128: the warp's 32 lanes.  MT = 64: the tile's rows
// sit in lanes 0..15 of every quadrant; the two half-warps address the same 16 lanes at columns c and c + HALF.
template <int MT, int HALF>
__device__ __forceinline__ void ld_d4(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3;
  if (MT == 128) asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
  else asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr), "n"(HALF));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
template <int MT, int HALF>
__device__ __forceinline__ void st_d4(uint32_t taddr, const float* v) {
  if (MT == 128) asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])) : "memory");
  else asm volatile("tcgen05.st.sync.aligned.16x32bx2.x4.b32 [%0], %5, {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "n"(HALF) : "memory");
}

// this thread's 4 * NCH consecutive D columns (a worker's neurons are adjacent columns: Cfg::dcol): one x16 and one x4 load
template <int MT, int HALF, int NCH>
__device__ __forceinline__ void ld_dall(uint32_t taddr, float (&d)[NCH][4]) {
  static_assert(NCH == 5, "20 columns per worker");
  uint32_t r[20];
  if (MT == 128) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]) : "r"(taddr + 16));
  } else {
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16], %17;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr), "n"(HALF));
    asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x4.b32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]) : "r"(taddr + 16), "n"(HALF));
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) d[c][i] = __uint_as_float(r[4 * c + i]);
}

// lane s of a quad holds v[i] = (stream s, neuron i); afterwards lane i holds w[s] = (stream s, neuron i)   (scripts/probe_pm.cu)
__device__ __forceinline__ void quad_transpose(const float v[4], float w[4], int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  const float sa = b0 ? v[0] : v[1], sb = b0 ? v[2] : v[3];
  const float ra = __shfl_xor_sync(0xffffffffu, sa, 1), rb = __shfl_xor_sync(0xffffffffu, sb, 1);
  const float p0 = b0 ? ra : v[0], p1 = b0 ? v[1] : ra;
  const float q0 = b0 ? rb : v[2], q1 = b0 ? v[3] : rb;
  const float s0 = b1 ? p0 : q0, s1 = b1 ? p1 : q1;
  const float r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
  const float k0 = b1 ? q0 : p0, k1 = b1 ? q1 : p1;
  w[0] = b1 ? r0 : k0; w[1] = b1 ? r1 : k1; w[2] = b1 ? k0 : r0; w[3] = b1 ? k1 : r1;
}

// the 4 streams of one (point, neuron) -> [hi | lo] image at shared address `addr` (hi; lo is PART bytes further)
template <uint32_t PART>
__device__ __forceinline__ void store_jet(uint32_t addr, const float v[4]) {
  float hi[4], lo[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) split_tf32_fast(v[s], hi[s], lo[s]);
  sts4(addr, hi[0], hi[1], hi[2], hi[3]);
  sts4(addr + PART, lo[0], lo[1], lo[2], lo[3]);
}

template <int H, int L, int MT, bool TRAIN, bool DBG>
__global__ void __launch_bounds__(Cfg<H, MT>::NTHREADS, 1) nsf_pm_jet_kernel(const PArgs a) {
  using C = Cfg<H, MT>;
  using MiscT = Misc<H, L, C::GWL_SMEM ? 4 : 0>;
  static_assert(sizeof(MiscT) <= C::MISC, "misc region too small");
  static_assert(C::DWN + C::NACC_R * C::NB <= 512 && C::NACC_F * C::NB <= 512, "tensor memory columns");
  extern __shared__ __align__(1024) uint8_t smem[];
  MiscT* misc = reinterpret_cast<MiscT*>(smem + C::OFF_MISC);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform in the eyes of the compiler: the role branches below are uniform branches
  const NsfNetGeom& g = a.g;
  constexpr int NMS = TRAIN ? 2 * L - 1 : L;          // MMA stages per tile
  constexpr uint32_t DCOL_F = 0;
  constexpr uint32_t DCOL = C::DWN;                   // reverse stages: accumulators of NB columns behind the weight-gradient accumulator (columns [0, DWN))
  const uint32_t smem_base = smem_u32(smem);
  constexpr bool dbg = DBG;                           // cycle counters (nsf_get_stage_cycles): a separate instantiation, 6 % more instructions per stage
  constexpr int HO_MASK = C::HO;

  if (warp == C::NEW) tmem_alloc(&misc->tmem_base, 512);
  if (tid == 0) {
    if (smem_base & 1023u) __trap();
    for (int i = 0; i < 5; ++i) mbar_init(&misc->ready[i], C::NEW);
    mbar_init(&misc->dfull, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&misc->wdone[i], 1);
    mbar_init(&misc->dwfree, C::NEW);
    for (int i = 0; i < 3; ++i) { mbar_init(&misc->hi_full[i], 1); mbar_init(&misc->hi_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&misc->lo_full[i], 1); mbar_init(&misc->lo_free[i], 1); }
    mbar_fence_init();
  }
  {
    float* acc = &misc->loss[0][0];
    constexpr int NACC = (int)((sizeof(MiscT) - offsetof(MiscT, loss)) / 4);
    for (int i = tid; i < NACC; i += C::NTHREADS) acc[i] = 0.f;
  }
  // the weight-gradient MMAs (M = 128) read image rows past the H real neurons: keep them finite
  for (uint32_t i = tid * 16; i < 2 * C::IMG; i += C::NTHREADS * 16) *reinterpret_cast<float4*>(smem + i) = make_float4(0.f, 0.f, 0.f, 0.f);
  if (TRAIN && a.scratch) {   // the hidden-layer weight gradients are accumulated into the row tile by tile
    float* grow = a.scratch + (size_t)blockIdx.x * g.gs_row();
    for (int l = 1; l < L; ++l) {
      float4* w4 = reinterpret_cast<float4*>(grow + g.gs_w(l));
      for (int i = tid; i < H * g.HP / 4; i += C::NTHREADS) w4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = misc->tmem_base;
  const int my_tiles = ((int)blockIdx.x < a.n_tiles) ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  constexpr int W_ISSUE = C::NEW, W_PROD = C::NEW + 1;
  if (warp == W_ISSUE) {
    // =========================== issuer warp ===========================
    // (Dealing the MMA stages round-robin over 2 - 3 issuing warps on different schedulers was measured: the issuing warp's ~430 instructions per
    // stage do share a scheduler with quadrant 0's epilogue warps, but passing the turn costs more -- 7.67 against 7.31 ms.)
    const uint32_t leader = elect_one();
    const uint32_t sb4_0 = smem_base >> 4;
    uint32_t ready_ph = 0, stage_ctr = 0, lo_ctr0 = 0, lo_ctr1 = 0, wg_ctr = 0;
    long long c_wait = 0, c_issue = 0, c_wwait = 0;
    constexpr uint32_t AHI = desc_hi_t(512, 1), BHI = desc_hi(256);
    for (int t = 0; t < my_tiles; ++t) {
#pragma unroll 1
      for (int ms = 0; ms < NMS; ++ms) {
        const int s = ms + 1;
        const bool outst = (s == L);
        const uint32_t wsub = outst ? C::WSUB_O : C::WSUB;
        const uint32_t idesc = idesc_tf32(MT, outst ? 16 : C::NB, 1, 0);
        // Every descriptor below is a constant offset from the shared-memory base.  Left alone, the compiler hoists all ~150 of them out
        // of the stage loop and then SPILLS them: local-memory reloads between the MMAs of the one issuing thread cost 1.5 ms per 1e6
        // points (9.1 against 7.6 ms).  Making the base nominally depend on the stage index keeps them loop-variant: one integer add per operand.
        // (a.zero is a kernel argument equal to 0: the sum stays in the uniform datapath, a laundering asm would move it to a vector register
        // and cost an R2UR per operand and MMA on the scheduler that also serves quadrant 0's epilogue warps)
        const uint32_t sb4 = sb4_0 + (uint32_t)(a.zero * ms);
        const bool fwd_stage = s <= L;
        const uint32_t d_col = tmem + (fwd_stage ? DCOL_F : DCOL);
        long long t0 = 0, t1 = 0;
        const uint32_t a_hi = sb4 + (C::OFF_P >> 4) + lbo_field(C::GRP), a_lo = a_hi + (C::PART >> 4);
        // The epilogue hands the operand image over chunk by chunk (CW neurons = KPC k-steps of this contraction) and ALL three products
        // of a k-step are issued as soon as its neurons are written, so that only the last chunk's MMAs are left when the epilogue
        // finishes (a separate hi * hi pass after the hand-over left 10 - 15 MMAs, 600 - 900 cycles, exposed in every stage).
        // The tensor core TRUNCATES when it adds into the fp32 accumulator, so what shares an accumulator matters (oracle/tc_model.py):
        //   forward stages: the hi * hi products of k-step range a = ks NACC_F / KS go to accumulator a; the 2^-11-sized corrections
        //   (lo * hi, hi * lo) of every k-step go to the LAST accumulator, whose own hi * hi products (the last chunk's) are issued
        //   after all of them -- the small terms are in while that accumulator is small;
        //   reverse stages: NACC_R accumulators likewise (1 by default: the rounding bias of the dgrad chains does not move the gradient).
        const int na = fwd_stage ? C::NACC_F : C::NACC_R;
        const uint32_t d_corr = d_col + (uint32_t)((na - 1) * C::NB);
        int ready_upto = -1;      // chunks [0, ready_upto] have been handed over
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) {
          if (dbg) t0 = clock64();
          if (c > ready_upto) {
            int e = c;
            while (!((HO_MASK >> e) & 1)) ++e;
            mbar_wait_relaxed(&misc->ready[e], ready_ph, 32);
            tc_fence_after();
            ready_upto = e;
          }
          if (dbg) { t1 = clock64(); c_wait += t1 - t0; t0 = t1; }
#pragma unroll
          for (int kc = 0; kc < C::KPC; ++kc) {       // the chunk's corrections ...
            const int ks = c * C::KPC + kc, gi = ks / C::GK, kk = ks % C::GK;
            if (kk == 0) {
              if (dbg) t1 = clock64();
              mbar_wait(&misc->hi_full[gi], stage_ctr & 1u);
              if ((gi & 1) == 0) { mbar_wait(&misc->lo_full[0], lo_ctr0 & 1u); ++lo_ctr0; }
              else { mbar_wait(&misc->lo_full[1], lo_ctr1 & 1u); ++lo_ctr1; }
              if (dbg) c_wwait += clock64() - t1;
            }
            const uint32_t w_hi = sb4 + ((C::OFF_WHI + gi * C::WBLK) >> 4) + lbo_field(128);
            const uint32_t w_lo = sb4 + ((C::OFF_WLO + (gi & 1) * C::WBLK) >> 4) + lbo_field(128);
            const uint32_t da = (uint32_t)(ks * 1024) >> 4, dw = (kk * wsub) >> 4;
            mma_tf32_elect2(d_corr, a_lo + da, AHI, w_hi + dw, BHI, idesc, ks > 0, leader);
            mma_tf32_elect2(d_corr, a_hi + da, AHI, w_lo + dw, BHI, idesc, 1, leader);
            if (kk == C::GK - 1) mma_commit_elect(&misc->lo_free[gi & 1], leader);
          }
#pragma unroll
          for (int kc = 0; kc < C::KPC; ++kc) {       // ... then its hi * hi products (in the corrections' accumulator: after ALL corrections)
            const int ks = c * C::KPC + kc, gi = ks / C::GK, kk = ks % C::GK;
            const uint32_t w_hi = sb4 + ((C::OFF_WHI + gi * C::WBLK) >> 4) + lbo_field(128);
            const uint32_t da = (uint32_t)(ks * 1024) >> 4, dw = (kk * wsub) >> 4;
            // accumulator of this k-step's hi * hi product; it starts from zero unless it is the corrections' accumulator
            const int ac = fwd_stage ? C::acc_of(ks, C::NACC_F) : C::acc_of(ks, C::NACC_R);
            const int acp = ks == 0 ? -1 : (fwd_stage ? C::acc_of(ks - 1, C::NACC_F) : C::acc_of(ks - 1, C::NACC_R));
            const bool fresh = ac != acp && ac != na - 1;
            mma_tf32_elect2(d_col + (uint32_t)(ac * C::NB), a_hi + da, AHI, w_hi + dw, BHI, idesc, fresh ? 0 : 1, leader);
            if (kk == C::GK - 1) mma_commit_elect(&misc->hi_free[gi], leader);
          }
          if (dbg) c_issue += clock64() - t0;
        }
        if (dbg) t0 = clock64();
        ready_ph ^= 1u;
        mma_commit_elect(&misc->dfull, leader);
        ++stage_ctr;
        if (TRAIN && s > L) {
          // weight gradient of layer l = 2L - s: dW_l[128, DWN] += P[rows, j]^T Q[rows, k], both images read K-major (type 1)
          const uint32_t dw_col = tmem;
          if (wg_ctr >= 1) { mbar_wait_relaxed(&misc->dwfree, (wg_ctr - 1) & 1u, 32); tc_fence_after(); }   // previous contents drained
          ++wg_ctr;
          const uint32_t wdesc = idesc_tf32(128, C::DWN, 0, 0);
          const uint32_t p_hi = sb4 + (C::OFF_P >> 4), p_lo = p_hi + (C::PART >> 4);
          const uint32_t q_hi = sb4 + (C::OFF_Q >> 4), q_lo = q_hi + (C::PART >> 4);
#pragma unroll
          for (int qq = 0; qq < C::NQ; ++qq) {
#pragma unroll
            for (int kr = 0; kr < 4; ++kr) {
              const uint32_t o = (uint32_t)(qq * C::GRP + kr * 32) >> 4;
              mma_tf32_elect2(dw_col, p_lo + o, AHI, q_hi + o, AHI, wdesc, !(qq == 0 && kr == 0), leader);
              mma_tf32_elect2(dw_col, p_hi + o, AHI, q_lo + o, AHI, wdesc, 1, leader);
              mma_tf32_elect2(dw_col, p_hi + o, AHI, q_hi + o, AHI, wdesc, 1, leader);
            }
            mma_commit_elect(&misc->wdone[qq], leader);
          }
        }
        __syncwarp();
        if (dbg) c_issue += clock64() - t0;
      }
    }
    if (dbg && lane == 0) {
      long long* d = a.dbg + (size_t)blockIdx.x * 32;
      d[0] = c_wait; d[1] = c_issue; d[2] = c_wwait; d[3] = (long long)my_tiles * NMS;
    }
  } else if (warp == W_PROD) {
    // =========================== weight producer (one lane) ===========================
    if (lane == 0) {
      const long long total = (long long)my_tiles * NMS;
      uint32_t lo_use0 = 0, lo_use1 = 0;
      int img = 0;
      for (long long n = 0; n < total; ++n) {
        const uint8_t* src = a.wimg + (size_t)img * C::WSTAGE;
        const uint32_t blk = (img == L - 1) ? C::WBLK_O : C::WBLK;
#pragma unroll
        for (int gi = 0; gi < C::NG; ++gi) {
          if (n >= 1) mbar_wait_relaxed(&misc->hi_free[gi], (uint32_t)((n - 1) & 1), 200);
          mbar_expect_tx(&misc->hi_full[gi], blk);
          tma_bulk_g2s(smem + C::OFF_WHI + gi * C::WBLK, src + (size_t)gi * C::WBLK, blk, &misc->hi_full[gi]);
          if ((gi & 1) == 0) {
            if (lo_use0 >= 1) mbar_wait_relaxed(&misc->lo_free[0], (lo_use0 - 1) & 1u, 200);
            ++lo_use0;
            mbar_expect_tx(&misc->lo_full[0], blk);
            tma_bulk_g2s(smem + C::OFF_WLO, src + (size_t)(C::NG + gi) * C::WBLK, blk, &misc->lo_full[0]);
          } else {
            if (lo_use1 >= 1) mbar_wait_relaxed(&misc->lo_free[1], (lo_use1 - 1) & 1u, 200);
            ++lo_use1;
            mbar_expect_tx(&misc->lo_full[1], blk);
            tma_bulk_g2s(smem + C::OFF_WLO + C::WBLK, src + (size_t)(C::NG + gi) * C::WBLK, blk, &misc->lo_full[1]);
          }
        }
        if (++img == NMS) img = 0;
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    const int q = warp & 3, sub = warp >> 2;
    const int l16 = MT == 128 ? lane : (lane & 15);
    const int worker = MT == 128 ? sub : 2 * sub + (lane >> 4);
    const int kq = lane & 3;                              // neuron within a 4-neuron piece (after the transpose) == stream before it
    const int pt_loc = (MT == 128 ? 8 : 4) * q + (l16 >> 2);   // point of the tile this thread works for
    const int rg = MT == 128 ? q : (q >> 1);              // 32-row group of the images its rows are in
    const bool primary = worker == 0;                     // one worker per quadrant does the per-point bookkeeping
    const bool red_lane = l16 < 4;                        // lane that owns the warp-reduced sums of neuron kq
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // chunk c = neurons [CW c, CW (c + 1)); this worker's piece of it is neurons CW c + 4 worker + (0..3)
    // TMEM address of this worker's D columns in chunk 0 (MT = 64: the upper half-warp reads 4 columns further); chunk c: + CW c
    const uint32_t d_thr_0 = tmem + lane_addr + (uint32_t)(MT == 128 ? C::WCOLS * worker : 2 * C::WCOLS * sub);   // (MT = 64: the upper half-warp reads WCOLS columns further)
    const uint32_t d_thr_f = d_thr_0 + DCOL_F;                                                  // forward stages
    const uint32_t d_thr = d_thr_0 + DCOL;                                                      // reverse stages
    // byte offset of this thread's float4 (neuron 4 worker + kq of chunk 0) inside an image part; chunk c: + CSTR c
    const uint32_t img_thr = (uint32_t)(pt_loc >> 3) * C::GRP + (uint32_t)worker * 512u + (uint32_t)kq * 128u +
                             (uint32_t)((((pt_loc >> 1) & 3) ^ kq) * 32) + (uint32_t)(pt_loc & 1) * 16u;
    constexpr uint32_t CSTR = (C::CW / 4) * 512u;
    const uint32_t p_thr = smem_base + C::OFF_P + img_thr, q_thr = smem_base + C::OFF_Q + img_thr;
    const int k0 = 4 * worker + kq;                       // this thread's neuron in chunk 0 (chunk c: + CW c)
    const float* pk = a.pk;
    float4* stash_thr = TRAIN ? reinterpret_cast<float4*>(a.stash) + ((size_t)blockIdx.x * L * C::PTS + pt_loc) * H + k0 : nullptr;
    constexpr size_t STL = (size_t)C::PTS * H;            // stash stride between layers (float4)
    float gwl[3][C::NCH];
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
      for (int c = 0; c < C::NCH; ++c) gwl[o][c] = 0.f;
    uint32_t dfull_ph = 0, rs_ctr = 0;
    const uint64_t keep = l2_keep_policy();
    long long c_dwait = 0, c_wwait = 0, c_work = 0, c_s0 = 0, c_fwd = 0, c_out = 0, c_ra = 0, c_rb = 0, c_last = 0, c_dw_fwd = 0, c_dw_rev = 0;

    auto red_pts = [&](float v) {      // sum over the points of this worker's rows (fixed tree: deterministic)
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      if (MT == 128) v += __shfl_xor_sync(0xffffffffu, v, 16);
      return v;
    };
    auto chunk_done = [&](int c) {     // chunks up to c of the operands visible to the async proxy, TMEM accesses retired -> issuer
      if (!((HO_MASK >> c) & 1)) return;
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&misc->ready[c]);
    };

    // this worker's D cells of every chunk, the NA accumulators summed in fp32 (round to nearest), two accumulators per round trip
    auto load_d = [&](uint32_t base, auto na_tag, float (&d)[C::NCH][4]) {
      constexpr int NA = decltype(na_tag)::value;
      ld_dall<MT, C::WCOLS, C::NCH>(base, d);
      if (NA == 1) tmem_ld_wait();
#pragma unroll
      for (int ac = 1; ac < NA; ac += 2) {
        float e[C::NCH][4], f[C::NCH][4];
        ld_dall<MT, C::WCOLS, C::NCH>(base + (uint32_t)(ac * C::NB), e);
        if (ac + 1 < NA) ld_dall<MT, C::WCOLS, C::NCH>(base + (uint32_t)((ac + 1) * C::NB), f);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < C::NCH; ++c)
#pragma unroll
          for (int i = 0; i < 4; i += 2) {      // packed fp32x2 adds (sm_100): the same roundings, half the instructions
            float2 s2 = make_float2(e[c][i], e[c][i + 1]);
            if (ac + 1 < NA) s2 = __fadd2_rn(s2, make_float2(f[c][i], f[c][i + 1]));
            const float2 r2 = __fadd2_rn(make_float2(d[c][i], d[c][i + 1]), s2);
            d[c][i] = r2.x; d[c][i + 1] = r2.y;
          }
      }
    };
    // dW_l (TMEM lane = output neuron j, column = input neuron k) of the tile -> added to this CTA's gradient row.  Every element
    // of the row is only ever touched by one thread, in program order: the sums are bit-reproducible.  Then the accumulator is free.
    float* const grow = TRAIN ? a.scratch + (size_t)blockIdx.x * g.gs_row() : nullptr;
    auto flush_dw = [&](int l) {
      constexpr int CPW = H / C::NSUB;                  // columns per warp of a quadrant
      static_assert(H % C::NSUB == 0 && CPW % 20 == 0, "flush split");
      tc_fence_after();
      if (q * 32 < H) {                                 // (warp-uniform) this quadrant holds real rows
        // drain order (pm_dw_index): a warp instruction adds nl x 16 contiguous bytes
        const int nl = H - 32 * q < 32 ? H - 32 * q : 32;
        float* dst = grow + g.gs_w(l) + q * 32 * H + (sub * (CPW / 4)) * (nl * 4) + (lane < nl ? lane : 0) * 4;
#pragma unroll
        for (int h = 0; h < CPW; h += 20) {
          float v[20];
#pragma unroll
          for (int i = 0; i < 20; i += 4) tmem_ld4(tmem + lane_addr + (uint32_t)(sub * CPW + h + i), v + i);
          tmem_ld_wait();
          if (lane < nl) {
#pragma unroll
            for (int i = 0; i < 20; i += 4) red_add_v4(dst + ((h + i) / 4) * (nl * 4), v[i], v[i + 1], v[i + 2], v[i + 3], keep);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&misc->dwfree);
    };

    for (int t = 0; t < my_tiles; ++t) {
      const long long p_tile = ((long long)blockIdx.x + (long long)t * gridDim.x) * C::PTS;
      const long long gp = p_tile + pt_loc;
      const bool ok = gp < a.n;
      const float xv = ok ? __ldg(a.x + gp) : 0.f, yv = ok ? __ldg(a.y + gp) : 0.f;
      long long t0 = 0, t1 = 0;
      if (dbg) t0 = clock64();
      // ---------------- stage 0: layer 0 (K = 2) ----------------
      float4 st0[C::NCH];
#pragma unroll
      for (int c = 0; c < C::NCH; ++c) {
        const int k = k0 + C::CW * c;
        const float w0x = __ldg(pk + g.pk_w0x() + k), w0y = __ldg(pk + g.pk_w0y() + k), b0 = __ldg(pk + g.pk_b0() + k);
        const float z[4] = {fmaf(w0x, xv, fmaf(w0y, yv, b0)), w0x, w0y, 0.f};
        float v[4];
        nsf_jet_fwd(z, v);
        store_jet<C::PART>(p_thr + CSTR * c, v);
        st0[c] = make_float4(v[0], z[1], z[2], z[3]);
        chunk_done(c);
      }
      if (TRAIN) {      // the stash goes to L2 after the hand-overs: fence.proxy.async would wait for these stores
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) st_keep(stash_thr + C::CW * c, st0[c], keep);
      }
      if (dbg) { t1 = clock64(); c_work += t1 - t0; c_s0 += t1 - t0; }
      // ---------------- stages 1 .. L-1: hidden layers forward ----------------
#pragma unroll 1
      for (int s = 1; s < L; ++s) {
        float bias[C::NCH];
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) bias[c] = __ldg(pk + g.pk_b(s) + k0 + C::CW * c);
        if (dbg) t0 = clock64();
        mbar_wait(&misc->dfull, dfull_ph); dfull_ph ^= 1u;
        tc_fence_after();
        if (dbg) { t1 = clock64(); c_dwait += t1 - t0; c_dw_fwd += t1 - t0; }
        float d[C::NCH][4];
        load_d(d_thr_f, std::integral_constant<int, C::NACC_F>(), d);
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) {      // all transposes first: the hand-over fences below pin the shuffles in place
          float z[4];
          quad_transpose(d[c], z, lane);
          d[c][0] = z[0] + bias[c]; d[c][1] = z[1]; d[c][2] = z[2]; d[c][3] = z[3];
        }
#pragma unroll
        for (int c = 0; c < C::NCH; ++c) {
          float v[4];
          nsf_jet_fwd(d[c], v);
          store_jet<C::PART>(p_thr + CSTR * c, v);
          d[c][0] = v[0];
          chunk_done(c);
        }
        if (TRAIN) {    // (t, zx, zy, z_lap) to L2 after the hand-overs: fence.proxy.async would wait for these stores
#pragma unroll
          for (int c = 0; c < C::NCH; ++c) st_keep(stash_thr + s * STL + C::CW * c, make_float4(d[c][0], d[c][1], d[c][2], d[c][3]), keep);
        }
        if (dbg) { t0 = clock64(); c_work += t0 - t1; c_fwd += t0 - t1; }
      }
      // ---------------- stage L: output layer, residuals, adjoint seeds ----------------
      float ob[4][3];      // adjoint of the outputs: [stream][u, v, p]
      float res_mine = 0.f, res_vis = 0.f, res_vtm = 0.f, res_eb = 0.f;   // per-point results, written to global memory after the hand-overs
      float4 stl[C::NCH];  // stash of layer L-1 (this thread's neurons), for the output layer's backward
      float4 stc[C::NCH];  // stash of the layer below the one being differentiated: loaded once, used for the second operand of the
                           // weight gradient (a_{l-2}) and, one stage later, for the adjoint through its tanh
      {
        float pre_e = 0.f, pre_vtm = a.vis_t0, pre_w = 1.f;
        if (ok) {
          if (a.has_evm) { pre_e = __ldg(a.e_in + gp); if (a.vtm_in) pre_vtm = __ldg(a.vtm_in + gp); }
          if (a.w) pre_w = __ldg(a.w + gp);
        }
        const float bo0 = __ldg(pk + g.pk_bl() + 0), bo1 = __ldg(pk + g.pk_bl() + 1), bo2 = __ldg(pk + g.pk_bl() + 2);
        if (TRAIN) {
#pragma unroll
          for (int c = 0; c < C::NCH; ++c) stl[c] = ld_keep(stash_thr + (L - 1) * STL + C::CW * c, keep);
        }
        if (dbg) t0 = clock64();
        mbar_wait(&misc->dfull, dfull_ph); dfull_ph ^= 1u;
        tc_fence_after();
        if (dbg) { t1 = clock64(); c_dwait += t1 - t0; c_dw_fwd += t1 - t0; }
        float o[4];
        {                                                   // this row's (u, v, p, -) of stream kq, every accumulator
          float e[C::NACC_F][4];
#pragma unroll
          for (int ac = 0; ac < C::NACC_F; ++ac) ld_d4<MT, 4>(tmem + lane_addr + DCOL_F + (uint32_t)(ac * C::NB), e[ac]);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            o[i] = e[0][i];
#pragma unroll
            for (int ac = 1; ac < C::NACC_F; ++ac) o[i] += e[ac][i];
          }
        }
        if (MT == 64) {                                     // the upper half-warp read the neighbouring columns: take the lower one's
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = __shfl_sync(0xffffffffu, o[i], lane & 15);
        }
        if (kq == 0) { o[0] += bo0; o[1] += bo1; o[2] += bo2; }
        float U[4], V[4], Pp[4];
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
          const int src = (lane & ~3) | ss;
          U[ss] = __shfl_sync(0xffffffffu, o[0], src); V[ss] = __shfl_sync(0xffffffffu, o[1], src); Pp[ss] = __shfl_sync(0xffffffffu, o[2], src);
        }
        const float u = U[0], v = V[0];
        const float ux = a.cs1 * U[1], vx = a.cs1 * V[1], px = a.cs1 * Pp[1];
        const float uy = a.cs1 * U[2], vy = a.cs1 * V[2], py = a.cs1 * Pp[2];
        const float ul = a.cs2 * U[3], vl = a.cs2 * V[3];
        float ee = 0.f, vis = 0.f;
        if (a.has_evm) { ee = pre_e; vis = ok ? fminf(a.vis_t0, pre_vtm) : a.vis_t0; }
        const float nu = a.inv_Re + vis;
        const float eq1 = (u * ux + v * uy) + px - nu * ul;
        const float eq2 = (u * vx + v * vy) + py - nu * vl;
        const float eq3 = ux + vy;
        const float eq4 = a.has_evm ? (eq1 * (u - 0.5f) + eq2 * (v - 0.5f)) - ee : 0.f;
        const float cw = (TRAIN && ok) ? a.c_eq * pre_w : 0.f;
        const float g1 = cw * (2.f * eq1 + a.k4 * eq4 * (u - 0.5f));
        const float g2 = cw * (2.f * eq2 + a.k4 * eq4 * (v - 0.5f));
        const float g3 = 2.f * cw * eq3;
        const float g4 = a.k4 * cw * eq4;
        ob[0][0] = g1 * ux + g2 * vx + g4 * eq1; ob[0][1] = g1 * uy + g2 * vy + g4 * eq2; ob[0][2] = 0.f;
        ob[1][0] = a.cs1 * (g1 * u + g3); ob[1][1] = a.cs1 * (g2 * u); ob[1][2] = a.cs1 * g1;
        ob[2][0] = a.cs1 * (g1 * v); ob[2][1] = a.cs1 * (g2 * v + g3); ob[2][2] = a.cs1 * g2;
        ob[3][0] = -a.cs2 * nu * g1; ob[3][1] = -a.cs2 * nu * g2; ob[3][2] = 0.f;
        {
          const float e4[4] = {eq1, eq2, eq3, eq4};
          res_mine = e4[0];
#pragma unroll
          for (int i = 1; i < 4; ++i) res_mine = kq == i ? e4[i] : res_mine;
          res_vis = vis; res_vtm = a.alpha_evm * fabsf(ee); res_eb = -g4;
        }
        {
          // loss sums of this quadrant's points (lanes with kq == 0 carry one point each).  Every warp of the quadrant holds the same
          // per-point values: the nine sums are dealt out over them (one owner warp per sum: deterministic, and no warp carries them all)
          const bool own = (MT == 128 || lane < 16) && kq == 0, cnt = own && ok;
          const float r[9] = {cnt ? pre_w * eq1 * eq1 : 0.f, cnt ? pre_w * eq2 * eq2 : 0.f, cnt ? pre_w * eq3 * eq3 : 0.f, cnt ? pre_w * eq4 * eq4 : 0.f,
                              cnt ? vis : 0.f, cnt ? 1.f : 0.f, own ? ob[0][0] : 0.f, own ? ob[0][1] : 0.f, own ? ob[0][2] : 0.f};
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            if (i % C::NSUB == sub) {                       // warp-uniform
              const float vsum = red_pts(r[i]);
              if (lane == 0) misc->loss[q][i] += vsum;
            }
          }
        }
        if (!TRAIN) { if (dbg) c_work += clock64() - t1; }
      }
      auto write_point_results = [&]() {   // (MT = 64: the lower half-warp; the upper one works for the same rows)
        if (primary && ok) {
          if (a.resid_out) a.resid_out[(long long)kq * a.n + gp] = res_mine;
          if (kq == 0 && a.vis_t_out) a.vis_t_out[gp] = res_vis;
          if (kq == 1 && a.has_evm && a.vtm_out) a.vtm_out[gp] = res_vtm;
          if (TRAIN && kq == 2 && a.ebar_out) a.ebar_out[gp] = res_eb;
        }
      };
      if (!TRAIN) write_point_results();
      if (TRAIN) {
        {
          // output layer backward on FFMA (3 outputs): adjoint of a_{L-1}, weight gradient of the output layer.
          // a_{L-2} (second operand of the next weight gradient) comes from the stash alone: its loads fly meanwhile.
#pragma unroll
          for (int c = 0; c < C::NCH; ++c) stc[c] = ld_keep(stash_thr + (L - 2) * STL + C::CW * c, keep);
#pragma unroll
          for (int c = 0; c < C::NCH; ++c) {
            const int k = k0 + C::CW * c;
            const float wl0 = __ldg(pk + g.pk_wl() + k), wl1 = __ldg(pk + g.pk_wl() + g.HP + k), wl2 = __ldg(pk + g.pk_wl() + 2 * g.HP + k);
            float ab[4], act[4], zb[4];
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) ab[ss] = fmaf(ob[ss][0], wl0, fmaf(ob[ss][1], wl1, ob[ss][2] * wl2));
            nsf_act_from_stash(stl[c], act);
            float gl[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int ss = 0; ss < 4; ++ss) {
              gl[0] = fmaf(ob[ss][0], act[ss], gl[0]);
              gl[1] = fmaf(ob[ss][1], act[ss], gl[1]);
              gl[2] = fmaf(ob[ss][2], act[ss], gl[2]);
            }
            if (C::GWL_SMEM) {
#pragma unroll
              for (int o = 0; o < 3; ++o) {
                const float vsum = red_pts(gl[o]);
                if (red_lane) misc->gwl[q][o][k] += vsum;
              }
            } else {
#pragma unroll
              for (int o = 0; o < 3; ++o) gwl[o][c] += gl[o];
            }
            nsf_zbar_from(stl[c], ab, zb);
            const float sb0 = red_pts(zb[0]);
            if (red_lane) misc->gb[q][L - 1][k] += sb0;
            store_jet<C::PART>(p_thr + CSTR * c, zb);
            if (c == C::NCH - 1) {      // Q is read by the weight-gradient MMAs only, issued after the last chunk of P
#pragma unroll
              for (int cc = 0; cc < C::NCH; ++cc) {
                float am[4];
                nsf_act_from_stash(stc[cc], am);
                store_jet<C::PART>(q_thr + CSTR * cc, am);
              }
            }
            chunk_done(c);
          }
          write_point_results();
          if (dbg) { t0 = clock64(); c_work += t0 - t1; c_out += t0 - t1; }
        }
        // ---------------- stages L+1 .. 2L-1: reverse of hidden layer l = L-1 .. 1; D = adjoint of a_{l-1} ----------------
#pragma unroll 1
        for (int lm1 = L - 2; lm1 >= 0; --lm1) {       // lm1 = l - 1: the layer whose tanh is differentiated in this stage
          if (dbg) t0 = clock64();
          mbar_wait(&misc->dfull, dfull_ph); dfull_ph ^= 1u;
          tc_fence_after();
          if (dbg) { t1 = clock64(); c_dwait += t1 - t0; c_dw_rev += t1 - t0; }
          float d[C::NCH][4];
          load_d(d_thr, std::integral_constant<int, C::NACC_R>(), d);
          if (lm1 >= 1) {
#pragma unroll
            for (int c = 0; c < C::NCH; ++c) {
              float ab[4], zb[4];
              quad_transpose(d[c], ab, lane);
              nsf_zbar_from(stc[c], ab, zb);
              const float sb0 = red_pts(zb[0]);
              if (red_lane) misc->gb[q][lm1][k0 + C::CW * c] += sb0;
              st_d4<MT, C::WCOLS>(d_thr + 4 * c, zb);  // park zbar_{l-1} in this thread's own (now free) D cells
            }
            // a_{l-2} comes from the stash alone; the loads fly while this warp waits for the weight-gradient MMAs
#pragma unroll
            for (int c = 0; c < C::NCH; ++c) stc[c] = ld_keep(stash_thr + (lm1 - 1) * STL + C::CW * c, keep);
            tmem_st_wait();
            if (dbg) { t0 = clock64(); c_work += t0 - t1; c_ra += t0 - t1; }
            // the weight-gradient MMAs of this stage still read P and Q: wait for those over this thread's rows
            mbar_wait(&misc->wdone[rg], rs_ctr & 1u);
            ++rs_ctr;
            if (dbg) { t1 = clock64(); c_wwait += t1 - t0; }
            tc_fence_after();
            ld_dall<MT, C::WCOLS, C::NCH>(d_thr, d);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < C::NCH; ++c) {
              store_jet<C::PART>(p_thr + CSTR * c, d[c]);
              if (c == C::NCH - 1) {    // Q is read by the weight-gradient MMAs only, issued after the last chunk of P
#pragma unroll
                for (int cc = 0; cc < C::NCH; ++cc) {
                  float am[4];
                  nsf_act_from_stash(stc[cc], am);
                  store_jet<C::PART>(q_thr + CSTR * cc, am);
                }
              }
              chunk_done(c);
            }
            // every weight-gradient MMA of this stage (layer lm1 + 1), not only those over this thread's rows
            if (rg != C::NQ - 1) mbar_wait(&misc->wdone[C::NQ - 1], (rs_ctr - 1) & 1u);
            flush_dw(lm1 + 1);
            if (dbg) { t0 = clock64(); c_work += t0 - t1; c_rb += t0 - t1; }
          } else {
            // layer 0: its weight gradient (K = 2) and bias gradient on FFMA; nothing goes back to the tensor core
#pragma unroll
            for (int c = 0; c < C::NCH; ++c) {
              const int k = k0 + C::CW * c;
              float ab[4], zb[4];
              quad_transpose(d[c], ab, lane);
              nsf_zbar_from(stc[c], ab, zb);
              const float sb0 = red_pts(zb[0]);
              const float sx = red_pts(fmaf(zb[0], xv, zb[1])), sy = red_pts(fmaf(zb[0], yv, zb[2]));
              if (red_lane) { misc->gb[q][0][k] += sb0; misc->gw0x[q][k] += sx; misc->gw0y[q][k] += sy; }
            }
            if (dbg) { t0 = clock64(); c_work += t0 - t1; c_last += t0 - t1; }
            // every weight-gradient MMA of the tile: P and Q are rewritten by the next tile's stage 0 / output stage, and the flush follows
#pragma unroll
            for (int i = 0; i < C::NQ; ++i) mbar_wait(&misc->wdone[i], rs_ctr & 1u);
            ++rs_ctr;
            if (dbg) c_wwait += clock64() - t0;
            flush_dw(1);
          }
        }
      }
    }
    if (dbg && lane == 0) {
      long long* d = a.dbg + (size_t)blockIdx.x * 32 + 4;
      if (warp == a.dbg_wa) { d[0] = c_dwait; d[1] = c_wwait; d[2] = c_work; d[8] = c_s0; d[9] = c_fwd; d[10] = c_out; d[11] = c_ra; d[12] = c_rb; d[13] = c_last; d[14] = c_dw_fwd; d[15] = c_dw_rev; }
      if (warp == a.dbg_wb) { d[4] = c_dwait; d[5] = c_wwait; d[6] = c_work; d[16] = c_s0; d[17] = c_fwd; d[18] = c_out; d[19] = c_ra; d[20] = c_rb; d[21] = c_last; d[22] = c_dw_fwd; d[23] = c_dw_rev; }
    }
    // ---- CTA epilogue: per-quadrant accumulators and per-thread partials -> this CTA's gradient row ----
    if (a.scratch) {
      float* grow = a.scratch + (size_t)blockIdx.x * g.gs_row();
      float* gwls = C::GWL_SMEM ? &misc->gwl[0][0][0] : reinterpret_cast<float*>(smem + C::OFF_P);   // [4 quadrants][3][H] (the operand images are free now)
      if (TRAIN && !C::GWL_SMEM) {
#pragma unroll
        for (int c = 0; c < C::NCH; ++c)
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            const float vsum = red_pts(gwl[o][c]);
            if (red_lane) gwls[(q * 3 + o) * H + k0 + C::CW * c] = vsum;
          }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(C::NEPI) : "memory");
      if (TRAIN) {
        for (int k = tid; k < H; k += C::NEPI) {
          grow[g.gs_w0x() + k] = (misc->gw0x[0][k] + misc->gw0x[1][k]) + (misc->gw0x[2][k] + misc->gw0x[3][k]);
          grow[g.gs_w0y() + k] = (misc->gw0y[0][k] + misc->gw0y[1][k]) + (misc->gw0y[2][k] + misc->gw0y[3][k]);
#pragma unroll
          for (int l = 0; l < L; ++l) {
            const float v = (misc->gb[0][l][k] + misc->gb[1][l][k]) + (misc->gb[2][l][k] + misc->gb[3][l][k]);
            grow[(l == 0 ? g.gs_b0() : g.gs_b(l)) + k] = v;
          }
#pragma unroll
          for (int o = 0; o < 3; ++o)
            grow[g.gs_wl() + o * g.HP + k] = (gwls[(0 * 3 + o) * H + k] + gwls[(1 * 3 + o) * H + k]) + (gwls[(2 * 3 + o) * H + k] + gwls[(3 * 3 + o) * H + k]);
          grow[g.gs_wl() + 3 * g.HP + k] = 0.f;
        }
        if (tid < 4) grow[g.gs_bl() + tid] = tid < 3 ? (misc->loss[0][6 + tid] + misc->loss[1][6 + tid]) + (misc->loss[2][6 + tid] + misc->loss[3][6 + tid]) : 0.f;
      }
      if (tid < NSF_LOSS_SLOTS) grow[g.gs_loss() + tid] = tid < 6 ? (misc->loss[0][tid] + misc->loss[1][tid]) + (misc->loss[2][tid] + misc->loss[3][tid]) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == C::NEW) tmem_dealloc(tmem, 512);
}

// weight stage images: thread per (image, n, k).  Element (r = row of n, k) of part p lives at
//   p * NG * WBLK + (k / 40) * WBLK + ((k % 40) / 8) * (N_img * 32) + (n / 8) * 256 + ((k / 4) & 1) * 128 + (n % 8) * 16 + (k % 4) * 4
template <int H, int MT>
__global__ void nsf_pm_pack_kernel(NsfNetGeom g, const float* __restrict__ flat, uint8_t* __restrict__ wimg) {
  using C = Cfg<H, MT>;
  const int L = g.L;
  const int n_img = 2 * L - 1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n_img * H * H) return;
  const int img = (int)(idx / (H * H)), n = (int)(idx % (H * H)) / H, k = (int)(idx % H);
  float v = 0.f;
  uint32_t nrows = C::NB;
  if (img < L - 1) {                 // WF_l, l = img + 1: B[n = j_out][k = k_in] = W_l[j][k]
    const int l = img + 1, fo = 3 * H + (l - 1) * (H * H + H);
    v = flat[fo + n * H + k];
  } else if (img == L - 1) {         // output layer: rows o < n_out of 16
    if (n >= 16) return;
    nrows = 16;
    const int fo = 3 * H + (L - 1) * (H * H + H);
    if (n < g.n_out) v = flat[fo + n * H + k];
  } else {                           // WB_l, l = 2L - 1 - img: B[n = k_in][k = j_out] = W_l[j_out][k_in]
    const int l = 2 * L - 1 - img, fo = 3 * H + (l - 1) * (H * H + H);
    v = flat[fo + k * H + n];
  }
  float hi, lo;
  split_tf32(v, hi, lo);
  const int r = img == L - 1 ? n : C::dcol(n);     // image row = D column of the result (hidden layers: Cfg::dcol)
  const size_t off = (size_t)img * C::WSTAGE + (size_t)(k / (8 * C::GK)) * C::WBLK + (size_t)((k % (8 * C::GK)) / 8) * (nrows * 32) +
                     (size_t)(r >> 3) * 256 + (size_t)((k >> 2) & 1) * 128 + (size_t)(r & 7) * 16 + (size_t)(k & 3) * 4;
  *reinterpret_cast<float*>(wimg + off) = hi;
  *reinterpret_cast<float*>(wimg + off + (size_t)C::NG * C::WBLK) = lo;
}

struct PmState {
  uint8_t* wimg = nullptr;
  float* stash = nullptr;
  long long* dbg = nullptr;
  int* map = nullptr;          // flat parameter index -> offset inside one of THIS kernel's gradient rows
  int dbg_on = 0, last_grid = 0, grid = 0;
  size_t wstage = 0;
};

typedef void (*PmKernel)(const PArgs);
template <int H, int MT, int L>
PmKernel pm_kernel_of(bool train, bool dbg) {
  if (!train) return nsf_pm_jet_kernel<H, L, MT, false, false>;
  return dbg ? nsf_pm_jet_kernel<H, L, MT, true, true> : nsf_pm_jet_kernel<H, L, MT, true, false>;
}
PmKernel pm_kernel(int H, int L, bool train, bool dbg) {
  if (H == 80) {
    switch (L) {
      case 2: return pm_kernel_of<80, 128, 2>(train, dbg);
      case 3: return pm_kernel_of<80, 128, 3>(train, dbg);
      case 4: return pm_kernel_of<80, 128, 4>(train, dbg);
      case 5: return pm_kernel_of<80, 128, 5>(train, dbg);
      default: return pm_kernel_of<80, 128, 6>(train, dbg);
    }
  }
  switch (L) {
    case 2: return pm_kernel_of<120, 64, 2>(train, dbg);
    case 3: return pm_kernel_of<120, 64, 3>(train, dbg);
    default: return pm_kernel_of<120, 64, 4>(train, dbg);
  }
}

}  // namespace

int nsf_pm_supported(const NsfNetGeom& g) {
  if (g.n_out != 3 || g.L < 2) return 0;
  return (g.H == 80 && g.L <= 6) || (g.H == 120 && g.L <= 4);
}

// collocation points of one tile
int nsf_pm_tile_points(const NsfNetGeom& g) { return g.H == 80 ? Cfg<80, 128>::PTS : Cfg<120, 64>::PTS; }

int nsf_pm_init(NsfCtx* ctx) {
  if (ctx->pm) return NSF_OK;
  const NsfNetGeom& g = ctx->main.g;
  if (!nsf_pm_supported(g)) { nsf_set_error("tcgen05 path covers hidden = 80 (2..6 hidden layers) and hidden = 120 (2..4)"); return NSF_E_SHAPE; }
  PmState* s = new PmState();
  s->grid = ctx->sms < ctx->main.rows ? ctx->sms : ctx->main.rows;
  const size_t smem = g.H == 80 ? Cfg<80, 128>::SMEM_BYTES : Cfg<120, 64>::SMEM_BYTES;
  s->wstage = g.H == 80 ? Cfg<80, 128>::WSTAGE : Cfg<120, 64>::WSTAGE;
  const size_t wbytes = (size_t)(2 * g.L - 1) * s->wstage;
  const size_t sbytes = (size_t)s->grid * g.L * nsf_pm_tile_points(g) * g.H * 4 * sizeof(float);
  NSF_CUDA_OK(cudaMalloc((void**)&s->wimg, wbytes));
  NSF_CUDA_OK(cudaMemset(s->wimg, 0, wbytes));
  NSF_CUDA_OK(cudaMalloc((void**)&s->stash, sbytes));
  {
    std::vector<int> map(g.n_params);
    const int nsub = g.H == 80 ? Cfg<80, 128>::NSUB : Cfg<120, 64>::NSUB;
    for (int i = 0; i < g.n_params; ++i) {
      int o = nsf_flat_to_gs(g, i);
      for (int l = 1; l < g.L; ++l)
        if (o >= g.gs_w(l) && o < g.gs_w(l) + g.H * g.HP) { const int r = o - g.gs_w(l); o = g.gs_w(l) + pm_dw_index(g.H, nsub, r / g.HP, r % g.HP); break; }
      map[i] = o;
    }
    NSF_CUDA_OK(cudaMalloc((void**)&s->map, sizeof(int) * g.n_params));
    NSF_CUDA_OK(cudaMemcpy(s->map, map.data(), sizeof(int) * g.n_params, cudaMemcpyHostToDevice));
  }
  for (int v = 0; v < 3; ++v)
    NSF_CUDA_OK(cudaFuncSetAttribute(pm_kernel(g.H, g.L, v != 0, v == 2), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ctx->ws_bytes += (long long)(wbytes + sbytes);
  ctx->pm = s;
  return NSF_OK;
}

void nsf_pm_free(NsfCtx* ctx) {
  PmState* s = (PmState*)ctx->pm;
  if (!s) return;
  cudaFree(s->wimg); cudaFree(s->stash); cudaFree(s->map); if (s->dbg) cudaFree(s->dbg);
  delete s;
  ctx->pm = nullptr;
}

// gradient-row layout of the rows this kernel writes (nsf_finalize_launch)
const int* nsf_pm_map(NsfCtx* ctx) { return ((PmState*)ctx->pm)->map; }

// rows (CTAs) the launch for n points writes
int nsf_pm_grid(NsfCtx* ctx, long long n) {
  PmState* s = (PmState*)ctx->pm;
  const int pts = nsf_pm_tile_points(ctx->main.g);
  const long long tiles = (n + pts - 1) / pts;
  return (int)(tiles < s->grid ? tiles : s->grid);
}

int nsf_pm_launch(NsfCtx* ctx, const NsfKernelArgs& k, const float* flat_params, int* grid_out, nsf_stream_t st, int* launches) {
  PmState* s = (PmState*)ctx->pm;
  const NsfNetGeom& g = ctx->main.g;
  const long long tot = (long long)(2 * g.L - 1) * g.H * g.H;
  if (g.H == 80) nsf_pm_pack_kernel<80, 128><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, flat_params, s->wimg);
  else nsf_pm_pack_kernel<120, 64><<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(g, flat_params, s->wimg);
  NSF_CUDA_OK(cudaGetLastError());
  ++*launches;
  const bool train = k.mode == NSF_MODE_JET_STEP;
  PArgs a;
  a.g = g; a.pk = k.pk; a.wimg = s->wimg; a.x = k.x; a.y = k.y; a.n = k.n;
  a.e_in = k.e_in; a.vtm_in = k.vtm_in; a.vtm_out = k.vtm_out; a.w = k.w;
  a.inv_Re = k.inv_Re; a.vis_t0 = k.vis_t0; a.alpha_evm = k.alpha_evm; a.cs1 = k.cs1; a.cs2 = k.cs2; a.k4 = k.k4; a.c_eq = k.c_eq;
  a.has_evm = k.has_evm;
  a.resid_out = k.resid_out; a.vis_t_out = k.vis_t_out; a.ebar_out = k.ebar_out;
  a.stash = train ? s->stash : nullptr;
  a.scratch = train ? k.scratch : nullptr;
  const int pts = nsf_pm_tile_points(g);
  a.n_tiles = (int)((k.n + pts - 1) / pts);
  a.zero = 0;
  a.dbg = (s->dbg_on && train) ? s->dbg : nullptr;
  {
    static const int wa = [] { const char* v = getenv("NSF_PM_DBG_WA"); return v ? atoi(v) : 0; }();
    static const int wb = [] { const char* v = getenv("NSF_PM_DBG_WB"); return v ? atoi(v) : -1; }();
    a.dbg_wa = wa; a.dbg_wb = wb >= 0 ? wb : (g.H == 80 ? Cfg<80, 128>::NEW : Cfg<120, 64>::NEW) - 1;
  }
  const int grid = a.n_tiles < s->grid ? a.n_tiles : s->grid;
  if (grid <= 0) { *grid_out = 0; return NSF_OK; }
  const size_t smem = g.H == 80 ? Cfg<80, 128>::SMEM_BYTES : Cfg<120, 64>::SMEM_BYTES;
  const int nthreads = g.H == 80 ? Cfg<80, 128>::NTHREADS : Cfg<120, 64>::NTHREADS;
  pm_kernel(g.H, g.L, train, a.dbg != nullptr)<<<grid, nthreads, smem, st>>>(a);
  NSF_CUDA_OK(cudaGetLastError());
  ++*launches;
  s->last_grid = grid;
  *grid_out = grid;
  return NSF_OK;
}

// Diagnostics: enable (out == NULL) or read back the cycle counters of the last launch, averaged over CTAs:
//   out[0..3]  issuer: wait for operands, issue (incl. weight waits), wait for weights, MMA stages
//   out[4..6]  epilogue warp 0: wait for D, wait for the weight-gradient MMAs, work;  out[8..10] the same for the last epilogue warp
int nsf_pm_stage_cycles(NsfCtx* ctx, double* out) {
  int rc = nsf_pm_init(ctx);
  if (rc != NSF_OK) return rc;
  PmState* s = (PmState*)ctx->pm;
  if (!s->dbg) { NSF_CUDA_OK(cudaMalloc((void**)&s->dbg, (size_t)s->grid * 32 * sizeof(long long))); NSF_CUDA_OK(cudaMemset(s->dbg, 0, (size_t)s->grid * 32 * sizeof(long long))); }
  s->dbg_on = 1;
  if (!out) return NSF_OK;
  NSF_CUDA_OK(cudaDeviceSynchronize());
  const int n = s->last_grid > 0 ? s->last_grid : 1;
  long long* h = new long long[(size_t)n * 32];
  if (cudaMemcpy(h, s->dbg, (size_t)n * 32 * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) { delete[] h; nsf_set_error("cudaMemcpy failed"); return NSF_E_CUDA; }
  for (int i = 0; i < 256; ++i) out[i] = 0.0;
  for (int i = 0; i < 32; ++i) { double acc = 0; for (int c = 0; c < n; ++c) acc += (double)h[(size_t)c * 32 + i]; out[i] = acc / n; }
  delete[] h;
  return NSF_OK;
}
