// Internal declarations shared by the translation units of libnsf_b200.so.
// Not part of the ABI (see include/nsf_b200.h for that).
//
// NSF_EMU builds the same host orchestration against a host "device" (malloc + loops over CTAs and
// threads, see nsf_ffma_body.h).  That build exists ONLY for tests/emu -- it checks the index logic
// of the kernels in a container without a GPU.  The product library is never built with it and the
// Python package never loads it.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "nsf_b200.h"
#include "nsf_geom.h"

#ifdef NSF_EMU
#include <cstdlib>
#include <cstring>
typedef void* nsf_stream_t;
static inline int nsf_rt_malloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 1); return *p ? 0 : 1; }
static inline void nsf_rt_free(void* p) { std::free(p); }
static inline int nsf_rt_memset0(void* p, size_t n, nsf_stream_t) { std::memset(p, 0, n); return 0; }
static inline int nsf_rt_upload(void* d, const void* h, size_t n, nsf_stream_t) { std::memcpy(d, h, n); return 0; }
#else
#include <cuda_runtime.h>
typedef cudaStream_t nsf_stream_t;
#endif

void nsf_set_error(const char* fmt, ...);

#ifndef NSF_EMU
#define NSF_CUDA_OK(call)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      nsf_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return NSF_E_CUDA;                                                                    \
    }                                                                                       \
  } while (0)
#endif

// One network's device-side state.
struct NsfNetState {
  NsfNetDesc desc;
  NsfNetGeom g;
  float* pk = nullptr;       // packed image [pk_size]
  float* scratch = nullptr;  // gradient rows [rows][gs_row]
  int* map = nullptr;        // flat index -> gs index [n_params]
  float* stash = nullptr;    // [rows][stash_stride]
  long long stash_stride = 0;
  int rows = 0;
};

struct NsfCtx {
  int device = 0;
  int sms = 0;
  int path = 0;  // requested: 0 auto, 1 ffma, 2 umma
  int launches = 0;
  long long ws_bytes = 0;
  bool has_evm = false;
  NsfNetState main, evm;
  float* e_buf = nullptr;     // [cap] EVM output at the collocation points
  float* ebar_buf = nullptr;  // [cap] d(loss)/d(e)
  long long cap = 0;
  void* pm = nullptr;    // tcgen05 path state, points-on-M kernel (nsf_pm_jet.cu)
  void* side = nullptr;     // side stream: the data blocks (boundary / supervised MSE) run beside the EVM forward + jet kernel
  void* ev_fork = nullptr;
  void* ev_join = nullptr;
  int timing = 0;        // bracket the dominant kernel with events (nsf_set_timing)
  void* ev0 = nullptr;
  void* ev1 = nullptr;
  int timed = 0;
};

// nsf_ffma.cu ------------------------------------------------------------------------------
// point-tile size the FFMA kernels use for (NS, HP)
int nsf_ffma_pt(int ns, int hp);
// resident CTAs per SM of that instantiation (0 on error)
int nsf_ffma_occupancy(int ns, int hp);
// launches mode a.mode; fills a.n_tiles; grid must be <= the rows behind a.scratch / a.stash
int nsf_ffma_launch(NsfKernelArgs& a, int ns, int grid, nsf_stream_t st);
int nsf_pack_launch(const NsfNetGeom& g, const float* flat, float* pk, nsf_stream_t st);
int nsf_finalize_launch(const NsfNetGeom& g, const float* scratch, int rows, const int* map, float* grad,
                        float* loss_parts, nsf_stream_t st, const int* map0 = nullptr, int split = 0);
int nsf_adam_launch(float* params, const float* grad, float* m, float* v, long long n, float lr, float b1, float b2,
                    float eps, float bc1, float bc2, float grad_scale, nsf_stream_t st);

// nsf_aux.cu: out = scale * |in|
int nsf_scale_abs_launch(const float* in, float* out, long long n, float scale, nsf_stream_t st);

// nsf_value_fwd.cu (CUDA build only): one-output value forward, thread per point ----------------------------------
int nsf_value_fwd_supported(const NsfNetGeom& g);
int nsf_value_fwd_launch(const NsfNetGeom& g, int sms, const float* flat, const float* x, const float* y, long long n, float* out,
                         nsf_stream_t st);

// nsf_pm_jet.cu (CUDA build only): points-on-M tcgen05 kernel, hidden = 80 and hidden = 120 ------------------------
int nsf_pm_supported(const NsfNetGeom& g);
int nsf_pm_tile_points(const NsfNetGeom& g);
int nsf_pm_init(NsfCtx* ctx);
void nsf_pm_free(NsfCtx* ctx);
int nsf_pm_grid(NsfCtx* ctx, long long n);
const int* nsf_pm_map(NsfCtx* ctx);
int nsf_pm_stage_cycles(NsfCtx* ctx, double* out);
int nsf_pm_launch(NsfCtx* ctx, const NsfKernelArgs& k, const float* flat_params, int* grid_out, nsf_stream_t st, int* launches);

// flat parameter i of the packed image (host + device)
NSF_HD float nsf_pack_value(const NsfNetGeom& g, const float* flat, int i) {
  const int H = g.H, HP = g.HP;
  if (i < HP) return i < H ? flat[i * 2 + 0] : 0.f;
  if (i < 2 * HP) { const int j = i - HP; return j < H ? flat[j * 2 + 1] : 0.f; }
  if (i < 3 * HP) { const int j = i - 2 * HP; return j < H ? flat[2 * H + j] : 0.f; }
  i -= 3 * HP;
  const int per = 2 * HP * HP + HP;
  if (i < (g.L - 1) * per) {
    const int l = 1 + i / per;
    int r = i % per;
    const int fo = 3 * H + (l - 1) * (H * H + H);
    if (r < HP * HP) { const int k = r / HP, j = r % HP; return (k < H && j < H) ? flat[fo + j * H + k] : 0.f; }
    r -= HP * HP;
    if (r < HP * HP) { const int j = r / HP, k = r % HP; return (k < H && j < H) ? flat[fo + j * H + k] : 0.f; }
    r -= HP * HP;
    return r < H ? flat[fo + H * H + r] : 0.f;
  }
  i -= (g.L - 1) * per;
  const int fo = 3 * H + (g.L - 1) * (H * H + H);
  if (i < 4 * HP) { const int o = i / HP, j = i % HP; return (o < g.n_out && j < H) ? flat[fo + o * H + j] : 0.f; }
  i -= 4 * HP;
  return i < g.n_out ? flat[fo + g.n_out * H + i] : 0.f;
}
