// Internal object behind crux_convq: the pixel-DQN trunk of examples/rl/atari.jl:8
//   Chain(x -> x ./ 255f0, Conv((k1,k1), C => c1, relu, stride = s1), Conv((k2,k2), c1 => c2, relu, stride = s2), flatten, Dense(F, hidden, relu), Dense(hidden, nA))
// as two implicit-GEMM convolution layers in front of a crux_mlp head.  Not part of the ABI.
#pragma once
#include "mlp.cuh"

struct ConvGeom {
  int C, H, W;     // input channels / height / width (memory order [b][c][h][w], w fastest = Flux WHCN)
  int K, S;        // square kernel, stride (no padding, dilation 1: the Flux defaults the example uses)
  int OH, OW, CO;  // output height / width / channels
};

struct crux_convq {
  crux_ctx *ctx = nullptr;
  ConvGeom g1, g2;
  int scale255 = 1;          // the leading `x -> x ./ 255f0` layer of the example (0: inputs are used as they are)
  int F = 0;                 // flatten width = c2 * OH2 * OW2 (NCHW order: Flux.flatten of a WHCN array)
  crux_mlp *head = nullptr;  // Dense(F, hidden, relu), Dense(hidden, nA); owned
  int64_t n_conv = 0;        // conv parameters: W1 [c1][C][k1][k1] | b1 [c1] | W2 [c2][c1][k2][k2] | b2 [c2]  (Flux.params order and memory)
  int64_t off_b1 = 0, off_w2 = 0, off_b2 = 0;
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
  float clip = 0.f;          // ClipValue(clip) in front of Adam (examples/rl/atari.jl:10); 0 = none
  // workspaces for batch `cap`
  int64_t cap = 0;
  float *y1 = nullptr;       // relu(conv1)  [cap][OH1*OW1][c1]   (NHWC)
  float *f = nullptr;        // relu(conv2)  [cap][F]             (NCHW = the flattened head input)
  float *dy1 = nullptr;      // gradient wrt conv1's pre-activation, same layout as y1
  float *partials = nullptr; // split-K weight-gradient partials
  size_t partials_bytes = 0;
};

// forward keeping activations; q values are left in net->head->act[L]
int convq_forward_keep(crux_convq *net, const void *s, int s_is_u8, int64_t B);
// backward from dq = head->dz[L] (gradient wrt the Q outputs): fills net->grads (conv) and net->head->grads
int convq_backward(crux_convq *net, const void *s, int s_is_u8, int64_t B, float *dq);
// ||g||, NaN check, [ClipValue] + Adam over conv and head parameters as ONE optimiser step (training.jl:18-21)
int convq_adam_step(crux_convq *net, float *gnorm_out_dev);
