// Internal MLP object (a Flux Chain of Dense layers + its Adam state).  Not part of the ABI.
#pragma once
#include "common.cuh"

#define CRUX_MAX_LAYERS 8

struct crux_mlp {
  crux_ctx *ctx = nullptr;
  int n_layers = 0;
  int dims[CRUX_MAX_LAYERS + 1] = {0};
  int acts[CRUX_MAX_LAYERS] = {0};
  int64_t n_params = 0;
  int64_t w_off[CRUX_MAX_LAYERS] = {0};  // offset of W_l ([in][out] row-major) in the flat vector; b_l follows
  float *params = nullptr;   // flat, Flux.params order
  float *grads = nullptr;    // same layout
  float *m = nullptr, *v = nullptr;  // Adam moments (float32 like Flux)
  double eta = (double)3e-4f, beta1 = 0.9, beta2 = 0.999, eps = 1e-8;  // training.jl:3: Adam(3f-4)
  int *step_dev = nullptr;   // number of Adam steps applied (device-resident: early stops never sync the host)
  // activation / gradient workspaces for batch `cap`
  int64_t cap = 0;
  float *act[CRUX_MAX_LAYERS + 1] = {nullptr};   // act[l] = output of layer l (l >= 1)
  float *dz[CRUX_MAX_LAYERS + 1] = {nullptr};    // dz[l] = gradient wrt pre-activation of layer l; dz[0] = input gradient
  float *partials = nullptr;                      // split-batch weight-gradient partials
  size_t partials_bytes = 0;
  double *norm_part = nullptr;                    // grad-norm partial sums (1024 doubles)
  // tensor-core PPO update (ppo_fused.cu): the weights in MMA B-fragment order + biases, rebuilt at the start of every
  // crux_ppo_update and kept in step with `params` by the fused Adam kernel for the rest of that update
  float *frag = nullptr;
  size_t frag_bytes = 0;
};

int mlp_ensure_workspace(crux_mlp *mlp, int64_t B);
// forward keeping activations (for a following backward). x: [B][dims[0]]. Result in mlp->act[n_layers].
int mlp_forward_keep(crux_mlp *mlp, const float *x, int64_t B, const int *skip_dev);
// forward into y (activations go to the workspace as well)
int mlp_forward_out(crux_mlp *mlp, const float *x, int64_t B, float *y);
// backward.  dY = gradient wrt the network OUTPUT [B][dims[L]] (overwritten when the last layer has an
// activation).  Writes (or accumulates into) mlp->grads.  If need_dx the input gradient is left in mlp->dz[0].
// params_grad == false: only the input gradient is propagated (a frozen critic under an actor loss).
int mlp_backward(crux_mlp *mlp, const float *x, int64_t B, float *dY, bool need_dx, bool accumulate,
                 bool params_grad, const int *skip_dev);

// Flux `train!` tail (training.jl:18-21) for up to 4 parameter segments sharing one optimiser:
// gnorm = ||g||_2 (NaN -> sticky CRUX_FLAG_NAN, no update), then the Adam step.
struct AdamSeg { float *p; float *g; float *m; float *v; int64_t n; };
struct AdamSegs { AdamSeg s[4]; int n; float clip = 0.f; };   // clip > 0: Flux.Optimiser(ClipValue(clip), Adam(...)) -- every gradient entry clamped to [-clip, clip] before Adam
// step_dev: device step counter (incremented here).  gnorm_out_dev nullable.  skip_dev nullable: *skip != 0 -> no-op.
int adam_step_segments(crux_ctx *ctx, const AdamSegs &segs, double eta, double beta1, double beta2, double eps,
                       int *step_dev, float *gnorm_out_dev, const int *skip_dev, double *norm_part);
int mlp_adam_step(crux_mlp *mlp, float *gnorm_out_dev, const int *skip_dev);
// sum gradients over ranks (no-op when world == 1)
int grads_allreduce(crux_ctx *ctx, float *g, int64_t n);

// ---- fused fast path (ppo_fused.cu): Chain(Dense(I,64,act), Dense(64,64,act), Dense(64,O)), I <= 32, O <= 8
// value(π, s) in one launch; *handled == 0 -> the caller runs the generic engine.
int mlp_forward_fused(crux_mlp *mlp, const float *x, int64_t B, float *y, int *handled);
int mlp_value_next_fused(crux_mlp *mlp, const float *sp, const float *s, const float *v_s, int64_t T, int64_t N, float *v_sp, int *handled);
