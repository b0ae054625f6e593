// C ABI of libnsf_b200.so (include/nsf_b200.h): context management and the orchestration of one
// loss + gradient evaluation.  No torch types, no exceptions across the boundary.
#include "nsf_internal.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <vector>

static thread_local char g_err[512] = "";

void nsf_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

#ifdef NSF_EMU
#define NSF_RT_OK(call) do { if ((call) != 0) { nsf_set_error("%s failed", #call); return NSF_E_ALLOC; } } while (0)
#else
static inline int nsf_rt_malloc(void** p, size_t n) { return cudaMalloc(p, n ? n : 1) == cudaSuccess ? 0 : 1; }
static inline void nsf_rt_free(void* p) { if (p) cudaFree(p); }
static inline int nsf_rt_memset0(void* p, size_t n, nsf_stream_t st) { return cudaMemsetAsync(p, 0, n, st) == cudaSuccess ? 0 : 1; }
static inline int nsf_rt_upload(void* d, const void* h, size_t n, nsf_stream_t) {
  return cudaMemcpy(d, h, n, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : 1;
}
#define NSF_RT_OK(call) do { if ((call) != 0) { nsf_set_error("%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); return NSF_E_ALLOC; } } while (0)
#endif

#define NSF_TRY(call) do { int rc__ = (call); if (rc__ != NSF_OK) return rc__; } while (0)

extern "C" int nsf_abi_version(void) { return NSF_ABI_VERSION; }
extern "C" const char* nsf_last_error(void) { return g_err; }

static int check_desc(const NsfNetDesc* d, const char* what) {
  if (d->n_in != 2) { nsf_set_error("%s: n_in must be 2 (x, y), got %d", what, d->n_in); return NSF_E_SHAPE; }
  if (d->n_out < 1 || d->n_out > 3) { nsf_set_error("%s: n_out must be 1..3, got %d", what, d->n_out); return NSF_E_SHAPE; }
  if (d->n_hidden_layers < 1 || d->n_hidden_layers > NSF_MAX_LAYERS) { nsf_set_error("%s: n_hidden_layers must be 1..%d", what, NSF_MAX_LAYERS); return NSF_E_SHAPE; }
  if (d->hidden < 4 || d->hidden > NSF_MAX_HIDDEN) { nsf_set_error("%s: hidden must be 4..%d", what, NSF_MAX_HIDDEN); return NSF_E_SHAPE; }
  return NSF_OK;
}

static int init_net(NsfCtx* ctx, NsfNetState& s, const NsfNetDesc* d, bool jet) {
  s.desc = *d;
  s.g = nsf_make_geom(d->n_out, d->n_hidden_layers, d->hidden);
  const NsfNetGeom& g = s.g;
  int occ1 = nsf_ffma_occupancy(1, g.HP), occ4 = jet ? nsf_ffma_occupancy(4, g.HP) : 0;
  if (occ1 <= 0 || (jet && occ4 <= 0)) { nsf_set_error("FFMA kernel does not fit on the SM for hidden=%d", d->hidden); return NSF_E_SHAPE; }
  s.rows = ctx->sms * (occ1 > occ4 ? occ1 : occ4) + (jet ? ctx->sms : 0);   // + one row per SM: the data blocks' rows behind a persistent tcgen05 launch
  long long st1 = (long long)g.L * g.HP * nsf_ffma_pt(1, g.HP);
  long long st4 = jet ? (long long)g.L * 4 * g.HP * nsf_ffma_pt(4, g.HP) : 0;
  s.stash_stride = st1 > st4 ? st1 : st4;
  NSF_RT_OK(nsf_rt_malloc((void**)&s.pk, sizeof(float) * g.pk_size()));
  NSF_RT_OK(nsf_rt_malloc((void**)&s.scratch, sizeof(float) * (size_t)s.rows * g.gs_row()));
  NSF_RT_OK(nsf_rt_malloc((void**)&s.stash, sizeof(float) * (size_t)s.rows * s.stash_stride));
  NSF_RT_OK(nsf_rt_malloc((void**)&s.map, sizeof(int) * g.n_params));
  std::vector<int> map(g.n_params);
  for (int i = 0; i < g.n_params; ++i) map[i] = nsf_flat_to_gs(g, i);
  NSF_RT_OK(nsf_rt_upload(s.map, map.data(), sizeof(int) * g.n_params, nullptr));
  ctx->ws_bytes += sizeof(float) * ((long long)g.pk_size() + (long long)s.rows * g.gs_row() + (long long)s.rows * s.stash_stride) +
                   sizeof(int) * g.n_params;
  return NSF_OK;
}

static void free_net(NsfNetState& s) {
  nsf_rt_free(s.pk); nsf_rt_free(s.scratch); nsf_rt_free(s.stash); nsf_rt_free(s.map);
  s.pk = s.scratch = s.stash = nullptr; s.map = nullptr;
}

extern "C" int nsf_create(int device, const NsfNetDesc* main_net, const NsfNetDesc* evm, NsfCtx** out) {
  if (!main_net || !out) { nsf_set_error("nsf_create: null argument"); return NSF_E_ARG; }
  *out = nullptr;
  NSF_TRY(check_desc(main_net, "main net"));
  if (main_net->n_out != 3) { nsf_set_error("main net: n_out must be 3 (u, v, p)"); return NSF_E_SHAPE; }
  if (evm) { NSF_TRY(check_desc(evm, "evm net")); if (evm->n_out != 1) { nsf_set_error("evm net: n_out must be 1"); return NSF_E_SHAPE; } }
  NsfCtx* ctx = new (std::nothrow) NsfCtx();
  if (!ctx) { nsf_set_error("out of host memory"); return NSF_E_ALLOC; }
  ctx->device = device;
#ifdef NSF_EMU
  ctx->sms = 3;  // a few "SMs" so that multi-row reduction and tile striding are exercised
#else
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    nsf_set_error("cudaGetDeviceProperties(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
    delete ctx; return NSF_E_CUDA;
  }
  if (prop.major != 10) {
    nsf_set_error("device %d is sm_%d%d; libnsf_b200 is built for sm_100a only and has no fallback", device, prop.major, prop.minor);
    delete ctx; return NSF_E_ARCH;
  }
  if (cudaSetDevice(device) != cudaSuccess) { nsf_set_error("cudaSetDevice(%d) failed", device); delete ctx; return NSF_E_CUDA; }
  ctx->sms = prop.multiProcessorCount;
#endif
  ctx->has_evm = evm != nullptr;
  int rc = init_net(ctx, ctx->main, main_net, true);
  if (rc == NSF_OK && evm) rc = init_net(ctx, ctx->evm, evm, false);
  if (rc != NSF_OK) { free_net(ctx->main); free_net(ctx->evm); delete ctx; return rc; }
  *out = ctx;
  return NSF_OK;
}

extern "C" int nsf_destroy(NsfCtx* ctx) {
  if (!ctx) return NSF_OK;
  free_net(ctx->main); free_net(ctx->evm);
  nsf_rt_free(ctx->e_buf); nsf_rt_free(ctx->ebar_buf);
#ifndef NSF_EMU
  nsf_pm_free(ctx);
  if (ctx->side) cudaStreamDestroy((cudaStream_t)ctx->side);
  if (ctx->ev_fork) cudaEventDestroy((cudaEvent_t)ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy((cudaEvent_t)ctx->ev_join);
  if (ctx->ev0) cudaEventDestroy((cudaEvent_t)ctx->ev0);
  if (ctx->ev1) cudaEventDestroy((cudaEvent_t)ctx->ev1);
#endif
  delete ctx;
  return NSF_OK;
}

// kernel family the next collocation launch uses: 1 = FFMA, 3 = tcgen05 with the points on M (hidden = 80 and hidden = 120).
// Auto (0) prefers 3.  (2 was the round-1 tcgen05 kernel with the neurons on M: slower and further from fp64 at trained weights, removed.)
static int effective_path(const NsfCtx* ctx) {
#ifdef NSF_EMU
  return 1;
#else
  if (ctx->path == 1) return 1;
  return nsf_pm_supported(ctx->main.g) ? 3 : 1;
#endif
}

extern "C" int nsf_set_path(NsfCtx* ctx, int path) {
  if (!ctx || path < 0 || path > 3 || path == 2) { nsf_set_error("nsf_set_path: path must be 0 (auto), 1 (FFMA) or 3 (tcgen05)"); return NSF_E_ARG; }
#ifdef NSF_EMU
  if (path >= 2) { nsf_set_error("tcgen05 path does not exist in the host emulation"); return NSF_E_SHAPE; }
#else
  if (path == 3 && !nsf_pm_supported(ctx->main.g)) {
    nsf_set_error("tcgen05 path 3 covers hidden = 80 (2..6 hidden layers) and hidden = 120 (2..4); this net is %d x %d", ctx->main.g.L, ctx->main.g.H);
    return NSF_E_SHAPE;
  }
#endif
  ctx->path = path;
  return NSF_OK;
}

extern "C" int nsf_get_info(NsfCtx* ctx, int64_t info[4]) {
  if (!ctx || !info) { nsf_set_error("nsf_get_info: null argument"); return NSF_E_ARG; }
  info[0] = ctx->sms; info[1] = effective_path(ctx); info[2] = ctx->launches; info[3] = ctx->ws_bytes;
  return NSF_OK;
}

extern "C" int nsf_get_stage_cycles(NsfCtx* ctx, double* out) {
  if (!ctx) { nsf_set_error("nsf_get_stage_cycles: null context"); return NSF_E_ARG; }
#ifdef NSF_EMU
  (void)out; nsf_set_error("nsf_get_stage_cycles: tcgen05 path does not exist in the host emulation"); return NSF_E_SHAPE;
#else
  if (effective_path(ctx) == 3) return nsf_pm_stage_cycles(ctx, out);
  nsf_set_error("nsf_get_stage_cycles: the tcgen05 paths do not cover this net"); return NSF_E_SHAPE;
#endif
}

// the collocation jet launch (step or residuals) on whichever kernel family is selected
static int launch_jet(NsfCtx* ctx, NsfKernelArgs& a, const float* flat_main, int* grid, nsf_stream_t st) {
#ifndef NSF_EMU
  if (effective_path(ctx) == 3) {
    NSF_TRY(nsf_pm_init(ctx));
    return nsf_pm_launch(ctx, a, flat_main, grid, st, &ctx->launches);
  }
#endif
  (void)flat_main;
  NSF_TRY(nsf_ffma_launch(a, 4, *grid, st));
  ctx->launches++;
  return NSF_OK;
}

#ifdef NSF_EMU
static int time_mark(NsfCtx*, int, nsf_stream_t) { return NSF_OK; }
extern "C" int nsf_set_timing(NsfCtx* ctx, int enable) { if (!ctx) return NSF_E_ARG; ctx->timing = enable; return NSF_OK; }
extern "C" int nsf_last_kernel_ms(NsfCtx*, float* ms) { if (ms) *ms = 0.f; return NSF_OK; }
#else
static int time_mark(NsfCtx* ctx, int which, nsf_stream_t st) {
  if (!ctx->timing) return NSF_OK;
  NSF_CUDA_OK(cudaEventRecord((cudaEvent_t)(which ? ctx->ev1 : ctx->ev0), st));
  if (which) ctx->timed = 1;
  return NSF_OK;
}
extern "C" int nsf_set_timing(NsfCtx* ctx, int enable) {
  if (!ctx) { nsf_set_error("nsf_set_timing: null context"); return NSF_E_ARG; }
  if (enable && !ctx->ev0) {
    cudaEvent_t a, b;
    NSF_CUDA_OK(cudaEventCreate(&a));
    NSF_CUDA_OK(cudaEventCreate(&b));
    ctx->ev0 = a; ctx->ev1 = b;
  }
  ctx->timing = enable ? 1 : 0;
  ctx->timed = 0;
  return NSF_OK;
}
extern "C" int nsf_last_kernel_ms(NsfCtx* ctx, float* ms) {
  if (!ctx || !ms) { nsf_set_error("nsf_last_kernel_ms: null argument"); return NSF_E_ARG; }
  if (!ctx->timed) { nsf_set_error("nsf_last_kernel_ms: no timed kernel yet (nsf_set_timing + nsf_step first)"); return NSF_E_ARG; }
  NSF_CUDA_OK(cudaEventSynchronize((cudaEvent_t)ctx->ev1));
  NSF_CUDA_OK(cudaEventElapsedTime(ms, (cudaEvent_t)ctx->ev0, (cudaEvent_t)ctx->ev1));
  return NSF_OK;
}
#endif

static int ensure_cap(NsfCtx* ctx, long long n) {
  if (n <= ctx->cap) return NSF_OK;
  nsf_rt_free(ctx->e_buf); nsf_rt_free(ctx->ebar_buf);
  ctx->e_buf = ctx->ebar_buf = nullptr; ctx->cap = 0;
  NSF_RT_OK(nsf_rt_malloc((void**)&ctx->e_buf, sizeof(float) * n));
  NSF_RT_OK(nsf_rt_malloc((void**)&ctx->ebar_buf, sizeof(float) * n));
  ctx->cap = n;
  return NSF_OK;
}

static int grid_for(const NsfCtx* ctx, const NsfNetState& s, int ns, long long n) {
  const int pt = nsf_ffma_pt(ns, s.g.HP);
  const long long tiles = (n + pt - 1) / pt;
  long long cap = (long long)ctx->sms * nsf_ffma_occupancy(ns, s.g.HP);
  if (cap > s.rows) cap = s.rows;
  return (int)(tiles < cap ? tiles : cap);
}

static void base_args(NsfKernelArgs& a, const NsfNetState& s, const float* x, const float* y, long long n, int mode) {
  a = NsfKernelArgs();
  a.g = s.g; a.pk = s.pk; a.x = x; a.y = y; a.n = n; a.mode = mode;
  a.stash = s.stash; a.stash_stride = s.stash_stride; a.scratch = s.scratch;
}

static void phys_args(NsfKernelArgs& a, const NsfPhysics* ph, long long n) {
  const bool has_evm = (ph->flags & NSF_HAS_EVM) != 0;
  a.inv_Re = ph->inv_Re; a.vis_t0 = ph->vis_t0; a.alpha_evm = ph->alpha_evm;
  a.cs1 = ph->coord_scale; a.cs2 = ph->coord_scale * ph->coord_scale;
  a.k4 = has_evm ? 2.f * ph->eq4_weight : 0.f;
  const double nf = ph->n_f_norm > 0 ? ph->n_f_norm : (double)n;
  a.c_eq = (float)((double)ph->alpha_e / nf);
  a.has_evm = has_evm ? 1 : 0;
}

// e = net_1(x, y) at n points: the thread-per-point kernel when it covers the net, else the FFMA tile kernel (needs ctx->evm.pk packed)
static int evm_forward(NsfCtx* ctx, const float* params_evm, bool packed, const float* x, const float* y, long long n, float* out, nsf_stream_t st) {
#ifndef NSF_EMU
  if (nsf_value_fwd_supported(ctx->evm.g)) {
    NSF_TRY(nsf_value_fwd_launch(ctx->evm.g, ctx->sms, params_evm, x, y, n, out, st)); ctx->launches++;
    return NSF_OK;
  }
#endif
  if (!packed) { NSF_TRY(nsf_pack_launch(ctx->evm.g, params_evm, ctx->evm.pk, st)); ctx->launches++; }
  NsfKernelArgs f;
  f = NsfKernelArgs();
  f.g = ctx->evm.g; f.pk = ctx->evm.pk; f.x = x; f.y = y; f.n = n; f.mode = NSF_MODE_FWD;
  f.out = out; f.scratch = nullptr; f.stash = nullptr;
  const int pt = nsf_ffma_pt(1, ctx->evm.g.HP);
  const long long tiles = (n + pt - 1) / pt;
  long long cap = (long long)ctx->sms * nsf_ffma_occupancy(1, ctx->evm.g.HP);
  if (cap > ctx->evm.rows) cap = ctx->evm.rows;
  NSF_TRY(nsf_ffma_launch(f, 1, (int)(tiles < cap ? tiles : cap), st)); ctx->launches++;
  return NSF_OK;
}

// parameter / gradient / scalar buffers: 4-byte aligned (the solver passes offset views of one flat buffer);
// per-point arrays (coordinates, targets, weights, lag state, per-point outputs): 16-byte aligned, as include/nsf_b200.h requires
static int valid_ptr(const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }
#ifdef NSF_EMU
#define NSF_PTS_MASK 3u      // the host emulation (tests only) reads element by element: numpy views are fine
#else
#define NSF_PTS_MASK 15u
#endif
static int valid_pts(const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & NSF_PTS_MASK) == 0; }
static int valid_pts_opt(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & NSF_PTS_MASK) == 0; }

extern "C" int nsf_forward(NsfCtx* ctx, int32_t which, const float* params, const float* x, const float* y, int64_t n,
                           float* out, void* stream) {
  if (!ctx || !valid_ptr(params) || (n > 0 && (!valid_pts(x) || !valid_pts(y) || !valid_pts(out))) || n < 0 || which < 0 || which > 1) {
    nsf_set_error("nsf_forward: bad argument (null pointer, or a per-point array that is not 16-byte aligned)"); return NSF_E_ARG;
  }
  if (which == 1 && !ctx->has_evm) { nsf_set_error("nsf_forward: context has no EVM net"); return NSF_E_ARG; }
  nsf_stream_t st = (nsf_stream_t)stream;
  NsfNetState& s = which ? ctx->evm : ctx->main;
  ctx->launches = 0;
  if (which == 1) return n == 0 ? NSF_OK : evm_forward(ctx, params, false, x, y, n, out, st);
  NSF_TRY(nsf_pack_launch(s.g, params, s.pk, st)); ctx->launches++;
  if (n == 0) return NSF_OK;
  NsfKernelArgs a;
  base_args(a, s, x, y, n, NSF_MODE_FWD);
  a.out = out; a.scratch = nullptr; a.stash = nullptr;
  NSF_TRY(nsf_ffma_launch(a, 1, grid_for(ctx, s, 1, n), st)); ctx->launches++;
  return NSF_OK;
}

extern "C" int nsf_residuals(NsfCtx* ctx, const float* params_main, const float* params_evm, const float* x, const float* y,
                             const float* vtm_in, float* vtm_out, int64_t n, const NsfPhysics* ph, float* residuals_out,
                             float* e_out, float* vis_t_out, void* stream) {
  if (!ctx || !ph || !valid_ptr(params_main) || n < 0 || (n > 0 && (!valid_pts(x) || !valid_pts(y))) || !valid_pts_opt(vtm_in) ||
      !valid_pts_opt(vtm_out) || !valid_pts_opt(residuals_out) || !valid_pts_opt(e_out) || !valid_pts_opt(vis_t_out)) {
    nsf_set_error("nsf_residuals: bad argument (null pointer, or a per-point array that is not 16-byte aligned)"); return NSF_E_ARG;
  }
  const bool has_evm = (ph->flags & NSF_HAS_EVM) != 0;
  if (has_evm && (!ctx->has_evm || !valid_ptr(params_evm))) { nsf_set_error("nsf_residuals: NSF_HAS_EVM needs an EVM net and its parameters"); return NSF_E_ARG; }
  nsf_stream_t st = (nsf_stream_t)stream;
  ctx->launches = 0;
  NSF_TRY(nsf_pack_launch(ctx->main.g, params_main, ctx->main.pk, st)); ctx->launches++;
  if (n == 0) return NSF_OK;
  const float* e_ptr = nullptr;
  if (has_evm) {
    float* eb = e_out;
    if (!eb) { NSF_TRY(ensure_cap(ctx, n)); eb = ctx->e_buf; }
    NSF_TRY(evm_forward(ctx, params_evm, false, x, y, n, eb, st));
    e_ptr = eb;
  }
  NsfKernelArgs a;
  base_args(a, ctx->main, x, y, n, NSF_MODE_JET_RESID);
  phys_args(a, ph, n);
  a.e_in = e_ptr; a.vtm_in = vtm_in; a.vtm_out = vtm_out; a.resid_out = residuals_out; a.vis_t_out = vis_t_out;
  a.scratch = nullptr; a.stash = nullptr;
  NSF_TRY(time_mark(ctx, 0, st));
  int gj = grid_for(ctx, ctx->main, 4, n);
  NSF_TRY(launch_jet(ctx, a, params_main, &gj, st));
  NSF_TRY(time_mark(ctx, 1, st));
  return NSF_OK;
}

extern "C" int nsf_step(NsfCtx* ctx, const float* params_main, const float* params_evm, const float* x, const float* y,
                        const float* w, const float* vtm_in, float* vtm_out, int64_t n_f, const NsfDataBlock* blocks,
                        int32_t n_blocks, const NsfPhysics* ph, float* grad_main, float* grad_evm, float* loss_parts,
                        float* residuals_out, float* e_out, float* vis_t_out, void* stream) {
  if (!ctx || !ph || !valid_ptr(params_main) || !valid_ptr(grad_main) || !valid_ptr(loss_parts) || n_f < 0 ||
      (n_f > 0 && (!valid_pts(x) || !valid_pts(y))) || n_blocks < 0 || n_blocks > NSF_MAX_BLOCKS || (n_blocks > 0 && !blocks) ||
      !valid_pts_opt(w) || !valid_pts_opt(vtm_in) || !valid_pts_opt(vtm_out) || !valid_pts_opt(residuals_out) || !valid_pts_opt(e_out) ||
      !valid_pts_opt(vis_t_out)) {
    nsf_set_error("nsf_step: bad argument (null pointer, or a per-point array that is not 16-byte aligned)"); return NSF_E_ARG;
  }
  const bool has_evm = (ph->flags & NSF_HAS_EVM) != 0;
  const bool evm_train = has_evm && (ph->flags & NSF_EVM_TRAINABLE) != 0;
  if (has_evm && (!ctx->has_evm || !valid_ptr(params_evm))) { nsf_set_error("nsf_step: NSF_HAS_EVM needs an EVM net and its parameters"); return NSF_E_ARG; }
  if (evm_train && !valid_ptr(grad_evm)) { nsf_set_error("nsf_step: NSF_EVM_TRAINABLE needs grad_evm"); return NSF_E_ARG; }
  for (int b = 0; b < n_blocks; ++b) {
    const NsfDataBlock& k = blocks[b];
    if (k.n < 0 || (k.n > 0 && (!valid_pts(k.x) || !valid_pts(k.y) || !valid_pts(k.u) || !valid_pts(k.v) || !valid_pts_opt(k.p)))) {
      nsf_set_error("nsf_step: bad data block %d (null or not 16-byte aligned)", b); return NSF_E_ARG;
    }
  }
  nsf_stream_t st = (nsf_stream_t)stream;
  NsfNetState& M = ctx->main;
  ctx->launches = 0;
  NSF_TRY(nsf_pack_launch(M.g, params_main, M.pk, st)); ctx->launches++;


  // grids of the launches that accumulate into the main net's gradient rows
  int grids[1 + NSF_MAX_BLOCKS];
  grids[0] = n_f > 0 ? grid_for(ctx, M, 4, n_f) : 0;
#ifndef NSF_EMU
  if (n_f > 0 && effective_path(ctx) == 3) {   // one persistent CTA per SM, one tile of 32 / 16 points per iteration
    NSF_TRY(nsf_pm_init(ctx));
    grids[0] = nsf_pm_grid(ctx, n_f);
  }
#endif
  for (int b = 0; b < n_blocks; ++b) grids[1 + b] = blocks[b].n > 0 ? grid_for(ctx, M, 1, blocks[b].n) : 0;
#ifndef NSF_EMU
  if (n_f > 0 && effective_path(ctx) >= 2)     // the tcgen05 kernels keep their rows to themselves (own layout of the hidden-layer blocks)
    for (int b = 0; b < n_blocks; ++b) if (grids[1 + b] > M.rows - grids[0]) grids[1 + b] = M.rows - grids[0];
#endif
  // With the tcgen05 jet kernel (own activation stash, one gradient row per SM) the data blocks get the gradient rows
  // BEHIND the jet kernel's and run on a side stream beside the EVM forward and the jet kernel; otherwise every launch
  // accumulates into the same rows, in stream order.
  int blk_rows = 0, blk_first = -1;
  for (int b = 0; b < n_blocks; ++b) {
    if (grids[1 + b] > 0 && blk_first < 0) blk_first = b;
    if (grids[1 + b] > blk_rows) blk_rows = grids[1 + b];
  }
  bool side = false;
#ifndef NSF_EMU
  side = n_f > 0 && effective_path(ctx) >= 2 && blk_rows > 0 && grids[0] + blk_rows <= M.rows;
  if (side && !ctx->side) {
    cudaStream_t s2; cudaEvent_t e0, e1;
    NSF_CUDA_OK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    NSF_CUDA_OK(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
    NSF_CUDA_OK(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming));
    ctx->side = s2; ctx->ev_fork = e0; ctx->ev_join = e1;
  }
#endif
  int first = -1, rows_used = 0;
  for (int i = 0; i < 1 + n_blocks; ++i) {
    if (grids[i] > 0 && first < 0) first = i;
    if (grids[i] > rows_used) rows_used = grids[i];
  }
  if (side) rows_used = grids[0] + blk_rows;
  if (first < 0) {  // nothing to do: zero outputs
    NSF_RT_OK(nsf_rt_memset0(grad_main, sizeof(float) * M.g.n_params, st));
    NSF_RT_OK(nsf_rt_memset0(loss_parts, sizeof(float) * NSF_LOSS_SLOTS, st));
    if (evm_train) NSF_RT_OK(nsf_rt_memset0(grad_evm, sizeof(float) * ctx->evm.g.n_params, st));
    return NSF_OK;
  }
  if (!side && rows_used > grids[first]) {  // rows the first (overwriting) launch does not touch
    NSF_RT_OK(nsf_rt_memset0(M.scratch + (size_t)grids[first] * M.g.gs_row(), sizeof(float) * (size_t)(rows_used - grids[first]) * M.g.gs_row(), st));
    ctx->launches++;
  }

  // the data blocks (boundary / supervised MSE): value-stream forward + reverse of the main net on a few thousand points
  auto launch_blocks = [&](nsf_stream_t bst) -> int {
#ifndef NSF_EMU
    if (side) {
      NSF_CUDA_OK(cudaStreamWaitEvent(bst, (cudaEvent_t)ctx->ev_fork, 0));
      if (blk_rows > grids[1 + blk_first]) {   // rows of the block region that the first (overwriting) block does not touch
        NSF_RT_OK(nsf_rt_memset0(M.scratch + (size_t)(grids[0] + grids[1 + blk_first]) * M.g.gs_row(),
                                 sizeof(float) * (size_t)(blk_rows - grids[1 + blk_first]) * M.g.gs_row(), bst));
        ctx->launches++;
      }
    }
#endif
    for (int b = 0; b < n_blocks; ++b) {
      const NsfDataBlock& k = blocks[b];
      if (k.n <= 0) continue;
      NsfKernelArgs a;
      base_args(a, M, k.x, k.y, k.n, NSF_MODE_MSE_STEP);
      a.accumulate = side ? (b == blk_first ? 0 : 1) : ((first == 1 + b) ? 0 : 1);
      if (side) a.scratch = M.scratch + (size_t)grids[0] * M.g.gs_row();
      a.tu = k.u; a.tv = k.v; a.tp = k.p; a.cu = k.cu; a.cv = k.cv; a.cp = k.cp;
      a.loss_slot = 6 + 4 * b;
      NSF_TRY(nsf_ffma_launch(a, 1, grids[1 + b], bst)); ctx->launches++;
    }
#ifndef NSF_EMU
    if (side) NSF_CUDA_OK(cudaEventRecord((cudaEvent_t)ctx->ev_join, bst));
#endif
    return NSF_OK;
  };
#ifndef NSF_EMU
  if (side) {
    // fork: the blocks only need the packed image; they overlap the EVM forward and the start of the persistent jet kernel
    NSF_CUDA_OK(cudaEventRecord((cudaEvent_t)ctx->ev_fork, st));
    NSF_TRY(launch_blocks((cudaStream_t)ctx->side));
  }
#endif

  if (n_f > 0) {
    const float* e_ptr = nullptr;
    if (has_evm) {
      if (evm_train) { NSF_TRY(nsf_pack_launch(ctx->evm.g, params_evm, ctx->evm.pk, st)); ctx->launches++; }   // the reverse pass reads the packed image
      NSF_TRY(ensure_cap(ctx, n_f));
      float* eb = e_out ? e_out : ctx->e_buf;
      NSF_TRY(evm_forward(ctx, params_evm, evm_train, x, y, n_f, eb, st));
      e_ptr = eb;
      if (ph->flags & NSF_VTM_FROM_E) {   // init_vis_t fused into this evaluation: the lag state it would have left behind
        if (!valid_pts(vtm_out)) { nsf_set_error("nsf_step: NSF_VTM_FROM_E needs vis_t_minus_out"); return NSF_E_ARG; }
        NSF_TRY(nsf_scale_abs_launch(eb, vtm_out, n_f, ph->alpha_evm_init, st)); ctx->launches++;
        vtm_in = vtm_out;
      }
    }
    NsfKernelArgs a;
    base_args(a, M, x, y, n_f, NSF_MODE_JET_STEP);
    phys_args(a, ph, n_f);
    a.accumulate = 0;
    a.e_in = e_ptr; a.vtm_in = vtm_in; a.vtm_out = vtm_out; a.w = w;
    a.resid_out = residuals_out; a.vis_t_out = vis_t_out;
    a.ebar_out = evm_train ? ctx->ebar_buf : nullptr;
#ifndef NSF_EMU
    // The blocks are joined BEHIND the jet kernel (before the gradient-row reduction).  Block CTAs still running when the jet kernel
    // starts only hold back the jet CTAs with the highest indices -- the ones with one tile less whenever the tile count is not a
    // multiple of the grid -- and at 10^6 points the blocks have long finished under the EVM forward.  Measured (scripts/ab_join.py):
    // 1043 -> 1002 us per step at 120 000 points, 4064 -> 3974 at 500 000, no change at 10^6.  NSF_LATE_JOIN=0: join in front.
    const char* lj = getenv("NSF_LATE_JOIN");
    const bool late_join = !(lj && atoi(lj) == 0);
    if (side && !late_join) NSF_CUDA_OK(cudaStreamWaitEvent(st, (cudaEvent_t)ctx->ev_join, 0));
#endif
    NSF_TRY(time_mark(ctx, 0, st));
    NSF_TRY(launch_jet(ctx, a, params_main, &grids[0], st));
    NSF_TRY(time_mark(ctx, 1, st));
#ifndef NSF_EMU
    if (side && late_join) NSF_CUDA_OK(cudaStreamWaitEvent(st, (cudaEvent_t)ctx->ev_join, 0));
#endif
  }
  if (!side) { NSF_TRY(launch_blocks(st)); }
  {
    const int* map0 = nullptr;
#ifndef NSF_EMU
    if (n_f > 0 && effective_path(ctx) == 3) map0 = nsf_pm_map(ctx);
#endif
    NSF_TRY(nsf_finalize_launch(M.g, M.scratch, rows_used, M.map, grad_main, loss_parts, st, map0, grids[0])); ctx->launches++;
  }

  if (evm_train) {
    NsfNetState& E = ctx->evm;
    if (n_f > 0) {
      NsfKernelArgs a;
      base_args(a, E, x, y, n_f, NSF_MODE_BAR_STEP);
      a.accumulate = 0;
      a.bar_in = ctx->ebar_buf;
      const int ge = grid_for(ctx, E, 1, n_f);
      NSF_TRY(nsf_ffma_launch(a, 1, ge, st)); ctx->launches++;
      NSF_TRY(nsf_finalize_launch(E.g, E.scratch, ge, E.map, grad_evm, nullptr, st)); ctx->launches++;
    } else {
      NSF_RT_OK(nsf_rt_memset0(grad_evm, sizeof(float) * E.g.n_params, st));
    }
  }
  return NSF_OK;
}

extern "C" int nsf_adam(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                        float beta1, float beta2, float eps, int64_t step, float grad_scale, void* stream) {
  if (n < 0 || step < 1 || (n > 0 && (!valid_ptr(params) || !valid_ptr(grad) || !valid_ptr(exp_avg) || !valid_ptr(exp_avg_sq)))) {
    nsf_set_error("nsf_adam: bad argument"); return NSF_E_ARG;
  }
  const float bc1 = (float)(1.0 - std::pow((double)beta1, (double)step));
  const float bc2 = (float)(1.0 - std::pow((double)beta2, (double)step));
  return nsf_adam_launch(params, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, bc1, bc2, grad_scale, (nsf_stream_t)stream);
}

#ifdef NSF_EMU
// host twin of csrc/nsf_eval.cu (the emulation's "device" memory is host memory)
extern "C" int nsf_error_norms(const float* pred, const float* u, const float* v, const float* p, int64_t n, double* s, void*) {
  if (n < 0 || !s || (n > 0 && (!pred || !u || !v))) { nsf_set_error("nsf_error_norms: bad argument"); return NSF_E_ARG; }
  for (int k = 0; k < 8; ++k) s[k] = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    const double du = (double)u[i] - pred[3 * i], dv = (double)v[i] - pred[3 * i + 1];
    s[0] += du * du; s[1] += (double)u[i] * u[i]; s[2] += dv * dv; s[3] += (double)v[i] * v[i];
    if (p && p[i] == p[i]) { const double dp = (double)p[i] - pred[3 * i + 2]; s[4] += dp * dp; s[5] += (double)p[i] * p[i]; s[6] += 1.0; }
  }
  return NSF_OK;
}
extern "C" int nsf_selftest_umma(int, int32_t, const float*, const float*, float*, int32_t, int32_t, void*) {
  nsf_set_error("nsf_selftest_umma: not available in the host emulation");
  return NSF_E_ARCH;
}
#endif
