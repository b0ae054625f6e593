/*
 * nsf_b200.h -- C ABI of the B200-native NSFnet / ev-NSFnet training hot path.
 *
 * The reference (latteine1217/NSFnet) has no FFI: its seam is the Python method API of
 * `PysicsInformedNeuralNetwork` (ev-NSFnet/pinn_solver.py, NSFnet/pinn_solver.py).  This header
 * is the boundary a binding for that seam loads (ctypes / cffi / cgo / JNI all work: plain
 * pointers and sizes, no C++ or torch types).  Each entry point names the reference code it
 * replaces.
 *
 * Conventions
 *   - every function returns an int status (NSF_OK == 0, negative == error) and never throws;
 *     nsf_last_error() returns a thread-local message for the last failure on this thread;
 *   - every buffer is a caller-owned CUDA *device* pointer, fp32, contiguous; nothing is retained after
 *     the call returns except inside the opaque context.  PER-POINT arrays (coordinates, targets,
 *     weights, lagged viscosity, per-point outputs) must be 16-byte aligned -- what `tensor.data_ptr()`
 *     gives for a fresh torch CUDA tensor, NOT for an odd-offset slice such as x[1:] -- and a
 *     misaligned one is rejected with NSF_E_ARG; parameter, gradient and loss buffers need 4-byte
 *     alignment only (they may be offset views of one flat buffer);
 *   - all work is enqueued asynchronously on the caller-supplied `cudaStream_t` (passed as
 *     void*; NULL = legacy default stream); results are ready when the stream reaches that point;
 *   - flat parameter / gradient order is FCNet's state_dict order (net.py:38-46):
 *     W0 [out,in] row-major, b0, W1, b1, ...;
 *   - a context is used by one host thread at a time; contexts on different devices are
 *     independent (one process per GPU under torchrun, as ev-NSFnet/train.py:22-43);
 *   - there is NO CPU fallback: nsf_create fails with NSF_E_ARCH unless the device is sm_100.
 */
#ifndef NSF_B200_H
#define NSF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSF_ABI_VERSION 1

enum {
  NSF_OK = 0,
  NSF_E_ARG = -1,    /* null / misaligned pointer, bad size or flag                    */
  NSF_E_ARCH = -2,   /* device is not compute capability 10.x                          */
  NSF_E_CUDA = -3,   /* a CUDA runtime call failed (message in nsf_last_error)         */
  NSF_E_SHAPE = -4,  /* network shape outside what the kernels cover                   */
  NSF_E_ALLOC = -5   /* workspace allocation failed                                    */
};

/* FCNet(num_ins, num_outs, num_layers, hidden_size) -- net.py:23-30. */
typedef struct {
  int32_t n_in;            /* must be 2 (x, y)                       */
  int32_t n_out;           /* 3 for the u,v,p net, 1 for the EVM net */
  int32_t n_hidden_layers; /* num_layers  (tanh layers), 1..16       */
  int32_t hidden;          /* hidden_size, 4..128                    */
} NsfNetDesc;

/* flags of NsfPhysics */
#define NSF_HAS_EVM 1u       /* ev-NSFnet: second net, eq4, lagged entropy viscosity            */
#define NSF_EVM_TRAINABLE 2u /* net_1 unfrozen (ev :501-511): also produce its gradient         */
#define NSF_VTM_FROM_E 4u    /* nsf_step only: the lag state is initialised inside this call, vis_t_minus_in is ignored and
                                vis_t = min(vis_t0, alpha_evm_init*|e|) with the e of THIS evaluation -- `init_vis_t()` (ev :138-140)
                                followed by the first loss evaluation on unchanged net_1 weights, sharing one net_1 forward */

/* Scalars of one loss evaluation -- ev-NSFnet/pinn_solver.py:32-54,67,311-342,387-397,426. */
typedef struct {
  float inv_Re;      /* 1/Re                                                         */
  float vis_t0;      /* 20/Re, cap of the entropy viscosity (ev :67)                 */
  float alpha_evm;   /* alpha in vis_t_minus = alpha*|e| (ev :334)                   */
  float alpha_e;     /* eq_weight (ev :426)                                          */
  float coord_scale; /* 1.0 unless coordinate_transform (ev :311-324)                */
  float eq4_weight;  /* 0.1 (ev :397)                                                */
  uint32_t flags;    /* NSF_HAS_EVM | NSF_EVM_TRAINABLE | NSF_VTM_FROM_E             */
  float alpha_evm_init; /* alpha of `init_vis_t` (ev :140); read only with NSF_VTM_FROM_E */
  double n_f_norm;   /* denominator of the residual means; <=0 means n_f. Under data parallelism
                        pass the GLOBAL point count so that summing gradients over ranks gives
                        the gradient on the union of the shards. */
} NsfPhysics;

/* A "value-stream MSE against targets" block: the boundary term (ev :374-379) and the optional
 * supervised term (ev :399-411) are both instances.  loss contribution =
 *   cu*sum (u-u_hat)^2 + cv*sum (v-v_hat)^2 + cp*sum_{isfinite(p)} (p-p_hat)^2
 * so boundary = {cu = cv = alpha_b/N_b, cp = 0}. */
typedef struct {
  const float* x;
  const float* y;
  const float* u;  /* targets */
  const float* v;
  const float* p;  /* may be NULL; NaN entries are masked out (ev :404-409) */
  int64_t n;
  float cu, cv, cp;
  int32_t reserved;
} NsfDataBlock;

#define NSF_MAX_BLOCKS 2

/* loss_parts layout (device float[NSF_LOSS_SLOTS], UNNORMALISED sums so that they all-reduce):
 *   [0..3]  sum w*eq_k^2, k=1..4           [4] sum vis_t        [5] number of collocation points
 *   [6+4b+0..3]  block b: sum du^2, sum dv^2, sum_masked dp^2, number of finite p targets */
#define NSF_LOSS_SLOTS 16

typedef struct NsfCtx NsfCtx;

/* Library / build info. */
int nsf_abi_version(void);
const char* nsf_last_error(void);

/* Create / destroy the per-device context (workspace: activation stash, per-CTA gradient
 * partials, packed weight images).  evm may be NULL (plain NSFnet).  Replaces nothing in the
 * reference; it is the analogue of building the nets on `cuda:local_rank` (ev :61-63,97-100). */
int nsf_create(int device, const NsfNetDesc* main_net, const NsfNetDesc* evm_or_null, NsfCtx** out);
int nsf_destroy(NsfCtx* ctx);

/* Select the kernel family of the hidden-layer contractions: 0 = auto (the tcgen05 kernel that covers the shape, else FFMA),
 * 1 = force the FP32 FFMA kernels, 3 = tcgen05 3xTF32 kernel with the points on the MMA's M (hidden = 80 with 2..6 hidden layers:
 * the ev-NSFnet main net; hidden = 120 with 2..4: NSFnet); 3 returns NSF_E_SHAPE if the shape is not covered.  (2 was the round-1
 * tcgen05 kernel, removed: NSF_E_ARG.) */
int nsf_set_path(NsfCtx* ctx, int path);
/* info[0]=SM count, [1]=path that the next nsf_step will use (1 FFMA / 3 tcgen05),
 * [2]=kernel launches issued by the last nsf_* call, [3]=workspace bytes. */
int nsf_get_info(NsfCtx* ctx, int64_t info[4]);

/* Measurement hook (bench.py roofline): when enabled, nsf_step / nsf_residuals bracket their dominant
 * kernel (the collocation jet kernel) with CUDA events on the caller's stream; nsf_last_kernel_ms
 * synchronises on the closing event and returns the elapsed milliseconds of the last one. */
int nsf_set_timing(NsfCtx* ctx, int enable);
int nsf_last_kernel_ms(NsfCtx* ctx, float* ms);

/* Diagnostics of the tcgen05 kernel: out == NULL switches per-warp cycle counters on for the following launches;
 * out != NULL (double[256]) synchronises the device and returns the counters of the last launch averaged over CTAs
 * (layout documented at nsf_pm_stage_cycles in csrc/nsf_pm_jet.cu).  Adds clock reads to the kernel: not for
 * timed runs. */
int nsf_get_stage_cycles(NsfCtx* ctx, double* out);

/* One `fwd_computing_loss_2d()` + `loss.backward()`:
 *   ev-NSFnet/pinn_solver.py:372-428 + :468-469 (DDP all-reduce excluded), i.e.
 *   neural_net_u on the blocks (:280-288), neural_net_equations on the collocation points
 *   (:290-342, replacing the 7 autograd sweeps :301-309), the weighted MSE (:387-397) and the
 *   parameter gradient.  NSFnet/pinn_solver.py:197-226 + :252 is the same call without NSF_HAS_EVM.
 *
 *   params_main/params_evm : flat parameters (state_dict order)
 *   x,y [n_f]              : collocation points;  w [n_f] or NULL: SDF weights (ev :387-392)
 *   vis_t_minus_in [n_f]   : alpha*|e| of the previous evaluation, or NULL -> constant vis_t0
 *                            (ev :327-331);  vis_t_minus_out [n_f] (may alias _in): this step's
 *                            alpha_evm*|e| (ev :334).  Both ignored without NSF_HAS_EVM.
 *   blocks[n_blocks]       : boundary / supervised MSE terms (host array of device pointers)
 *   grad_main, grad_evm    : OVERWRITTEN with d(loss)/d(params); grad_evm untouched unless
 *                            NSF_EVM_TRAINABLE
 *   loss_parts             : device float[NSF_LOSS_SLOTS], overwritten
 *   residuals_out          : NULL or device float[4*n_f] = eq1|eq2|eq3|eq4 (eq4 = 0 without EVM)
 *   e_out, vis_t_out       : NULL or device float[n_f] (self.evm, self.vis_t)
 *   Asynchronous on `stream`.  With the tcgen05 kernel the data blocks are forked onto a library-owned non-blocking
 *   side stream (event fork / join, so it is CUDA-graph capturable) and joined on `stream` before the call's last
 *   kernels; everything the call writes is ordered on `stream` when it returns.
 */
int nsf_step(NsfCtx* ctx, const float* params_main, const float* params_evm,
             const float* x, const float* y, const float* w,
             const float* vis_t_minus_in, float* vis_t_minus_out, int64_t n_f,
             const NsfDataBlock* blocks, int32_t n_blocks, const NsfPhysics* phys,
             float* grad_main, float* grad_evm, float* loss_parts,
             float* residuals_out, float* e_out, float* vis_t_out, void* stream);

/* `neural_net_equations(x, y)` without gradient (ev :290-342; used by `divergence` :761-765):
 * forward jet only.  residuals_out = eq1|eq2|eq3|eq4 [4*n].  Updates the lag state exactly as
 * the reference does when vis_t_minus_out != NULL. */
int nsf_residuals(NsfCtx* ctx, const float* params_main, const float* params_evm,
                  const float* x, const float* y, const float* vis_t_minus_in,
                  float* vis_t_minus_out, int64_t n, const NsfPhysics* phys,
                  float* residuals_out, float* e_out, float* vis_t_out, void* stream);

/* `neural_net_u(x, y)` (ev :280-288; evaluate/test :669-740): value-only forward of one net.
 * which = 0 main net (out [n,3] row-major u,v,p), 1 EVM net (out [n,1]). */
int nsf_forward(NsfCtx* ctx, int32_t which, const float* params, const float* x, const float* y,
                int64_t n, float* out, void* stream);

/* Fused Adam on a flat buffer (torch.optim.Adam semantics, betas/eps/wd=0 as ev :126-129):
 * step is the 1-based step count after increment.  "Next" row f.1 of the scope table. */
int nsf_adam(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
             float lr, float beta1, float beta2, float eps, int64_t step, float grad_scale,
             void* stream);

/* Device-resident Adam state: hyper-parameters and the step counter live in device memory, so a training iteration
 * (nsf_step + nsf_adam_dev [+ nsf_adam_dev for the EVM net] + nsf_adam_tick) has no host-side scalar that changes from
 * step to step and ONE captured CUDA graph replays it ("next" row f.1: the reference's per-iteration Python of
 * ev :456-472).  `step` = number of updates already applied; the update uses t = step + 1 for the bias corrections.
 * A fresh `torch.optim.Adam` (what freeze_evm_net / defreeze_evm_net create, ev :489-511) = zeroed moments + step = 0;
 * a new stage learning rate (ev train :383-388) = a 4-byte write to `lr`. */
typedef struct {
  float lr, beta1, beta2, eps;
  float grad_scale; /* multiplies the gradient first (1.0; 1/W reproduces the reference's DDP quirk) */
  int32_t step;
  int32_t reserved[2];
} NsfAdamDev;

/* One Adam update of a flat buffer with the DEVICE-resident state (does not advance state->step). */
int nsf_adam_dev(float* params, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                 NsfAdamDev* state, void* stream);
/* state->step += 1 on the device, after every buffer of the iteration has been updated. */
int nsf_adam_tick(NsfAdamDev* state, void* stream);

/* On-device point layer ("next" row f.2; the reference's pure-Python LHSample + cKDTree, tools.py:30-57 and
 * cavity_data.py:96-130, take hours at 1e6 points).
 *
 * nsf_lhs_points: rows [first, first+count) of a Latin-hypercube design of n_total points in
 *   [x_min,x_max] x [y_min,y_max]: per dimension every 1/n_total stratum holds exactly one point, at a uniform
 *   position inside it, strata assigned by an independent pseudo-random permutation per dimension (a keyed Feistel
 *   network, so any row range of the SAME design can be produced by any rank without communication).
 * nsf_wall_distance: distance of every point to the nearest of the n_b discrete boundary points
 *   (cKDTree(pts_bc).query, cavity_data.py:118-121; also the sort key of tools.py:68-83).
 * nsf_sdf_weights: w = w_min + (1-w_min) exp(-decay d), UNNORMALISED (cavity_data.py:122-127 clamps included);
 *   when w_sum != NULL the sum of the weights is ADDED to *w_sum (device double, caller zeroes it), so the caller
 *   divides by the mean over the GLOBAL point set (all-reduce of one double under data parallelism). */
int nsf_lhs_points(int64_t n_total, int64_t first, int64_t count, uint32_t seed, float x_min, float x_max,
                   float y_min, float y_max, float* x_out, float* y_out, void* stream);
int nsf_wall_distance(const float* x, const float* y, int64_t n, const float* xb, const float* yb, int32_t n_b,
                      float* dist_out, void* stream);
int nsf_sdf_weights(const float* x, const float* y, int64_t n, const float* xb, const float* yb, int32_t n_b,
                    float min_weight, float decay, float* w_out, double* w_sum, void* stream);

/* Error norms of `evaluate` / `test` on the device ("next" row f.3; ev-NSFnet/pinn_solver.py:684-688 copies the predictions
 * to the host and calls numpy.linalg.norm).  uvp_pred is the [n][3] output of nsf_forward; u_ref, v_ref, p_ref the DNS fields
 * at the same points, p_ref may hold NaN (masked, ev :684) or be NULL.  sums8 (device, 8 doubles, zeroed by the call) receives
 *   [0] sum (u-u_pred)^2  [1] sum u^2  [2] sum (v-v_pred)^2  [3] sum v^2  [4] sum_mask (p-p_pred)^2  [5] sum_mask p^2  [6] #mask
 * so that error_u = 100 sqrt([0]/[1]) etc.  fp64 accumulation; the order of the atomic adds is not fixed. */
int nsf_error_norms(const float* uvp_pred, const float* u_ref, const float* v_ref, const float* p_ref_or_null, int64_t n,
                    double* sums8, void* stream);

/* Debug / validation: runs one tcgen05 TF32 GEMM D[128,n] = A[128,k] * B[n,k]^T with the same
 * shared-memory descriptors the jet kernel uses (variant selects operand majors / 3xTF32 split)
 * so tests can check the descriptor encodings against a CPU product. */
int nsf_selftest_umma(int device, int32_t variant, const float* a, const float* b, float* d,
                      int32_t n, int32_t k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSF_B200_H */
