"""Secondary measurements for BASELINE configs[2] (DQN + prioritized replay, 84x84x4 u8 obs, 1M buffer) and configs[3]
(SAC 376-obs/17-act, 256-256, batch 2048) on one B200.  Not the headline bench: numbers go to profiles/ as evidence for the
buffer / off-policy rows of SURVEY 8.   python scripts/bench_offpolicy.py [--rows 1000000] > gpurun_out/offpolicy.json"""
import ctypes as C
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import crux_b200 as crux
from crux_b200.device import ptr

F32 = np.float32
rows = 1_000_000
for i, a in enumerate(sys.argv):
    if a == "--rows":
        rows = int(sys.argv[i + 1])
ctx = crux.Context(0)
out = {}


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


# ------------------------------------------------------------------ C3: prioritized replay over u8 pixel observations
S = crux.ContinuousSpace((84, 84, 4), np.uint8)
A = crux.DiscreteSpace(4)
buf = crux.ExperienceBuffer(S, A, rows, prioritized=True, ctx=ctx)
chunk = 16384
g = torch.Generator(device=ctx.device).manual_seed(3)
pushed = 0
while pushed < rows:
    n = min(chunk, rows - pushed)
    d = {"s": torch.randint(0, 256, (n, 84, 84, 4), dtype=torch.uint8, device=ctx.device, generator=g),
         "sp": torch.randint(0, 256, (n, 84, 84, 4), dtype=torch.uint8, device=ctx.device, generator=g),
         "a": torch.nn.functional.one_hot(torch.randint(0, 4, (n,), device=ctx.device, generator=g), 4).float(),
         "r": torch.randn((n, 1), device=ctx.device, generator=g), "done": (torch.rand((n, 1), device=ctx.device, generator=g) < 0.01).to(torch.uint8)}
    buf.push_(d)
    pushed += n
assert len(buf) == rows
pr = (torch.randn(rows, device=ctx.device, generator=g).abs() + 1e-3)
idx_all = torch.arange(rows, dtype=torch.int32, device=ctx.device)
buf.update_priorities_(idx_all, pr)
B = 512
tgt = crux.buffer_like(buf, capacity=B)
row_bytes = 2 * 84 * 84 * 4 + 4 * 4 + 4 + 1 + 1 + 4
ctr = [0]


def sample():
    ctr[0] += 1
    crux.prioritized_sample_(tgt, buf, i=1, B=B, seed=3, ctr=ctr[0])


td = torch.rand(B, device=ctx.device)


def sample_after_update():  # the steady state of DQN+PER: every update invalidates the prefix sum (off_policy.jl:83)
    buf.update_priorities_(tgt.indices_dev(), td)
    sample()


ms_cached = timed(sample)
ms_full = timed(sample_after_update)
push_rows = 4096
dpush = {k: v[:push_rows].contiguous() for k, v in d.items()}
ms_push = timed(lambda: buf.push_(dpush), reps=10)
out["c3_dqn_per"] = {
    "buffer_rows": rows, "buffer_bytes": rows * row_bytes, "batch": B,
    "prioritized_sample_ms_prefix_cached": ms_cached, "prioritized_sample_ms_after_priority_update": ms_full,
    "gather_bytes_per_sample": 2 * B * row_bytes, "gather_GBps_prefix_cached": 2 * B * row_bytes / (ms_cached * 1e-3) / 1e9,
    "scan_bytes": 8 * rows, "push_4096_rows_ms": ms_push, "push_GBps": 2 * push_rows * row_bytes / (ms_push * 1e-3) / 1e9,
    "note": "u8 observations stay u8 in HBM (56 GB for 1M rows); sample = device prefix scan (when priorities changed) + stratified binary search "
            "+ IS weights + row gather of s, sp, a, r, done, weight"}
# the whole DQN + PER update on the pixel network of examples/rl/atari.jl:8 (value_training, off_policy.jl:66-111): prioritized sample ->
# Q-(sp) and dqn_target -> Q(s), td_error, update_priorities! -> weighted td_loss train! (ClipValue(1) + Adam(1e-3)); polyak once per call
rngc = np.random.default_rng(5)
chain = crux.Chain(crux.scale255, crux.Conv((8, 8), 4, 16, crux.relu, stride=4, rng=rngc), crux.Conv((4, 4), 16, 32, crux.relu, stride=2, rng=rngc), crux.flatten,
                   crux.Dense(2592, 256, crux.relu, rng=rngc), crux.Dense(256, 4, rng=rngc))
piq = crux.DiscreteNetwork(chain, [0, 1, 2, 3], ctx=ctx, input_dims=(84, 84, 4))
Sq = crux.DQN(piq, S, N=10, dN=1, c_opt=dict(batch_size=B, epochs=1, optimizer=crux.Adam(F32(1e-3), clip_value=1.0)), buffer=buf, prioritized=True,
              weighted_loss=True, buffer_init=0)
Dq = crux.buffer_like(buf, capacity=B)
ms_update = timed(lambda: Sq.value_training(Dq, F32(0.99)), reps=20)
sB = Dq.column("s")
ms_fwd = timed(lambda: piq.mlp.forward(sB), reps=20)
yB, aB, wB = torch.randn(B, device=ctx.device), Dq.column("a"), torch.rand(B, device=ctx.device)
ms_train = timed(lambda: piq.mlp.train_dqn(sB, aB, yB, wB, B), reps=20)
f_fwd = 2 * (400 * 16 * 256 + 81 * 32 * 256 + 2592 * 256 + 256 * 4)        # forward FLOP per sample
out["c3_dqn_per"].update({
    "pixel_dqn_update_ms": ms_update, "pixel_dqn_samples_per_s": B * 1e3 / ms_update,
    "conv_forward_ms": ms_fwd, "conv_forward_TFLOPs": B * f_fwd / (ms_fwd * 1e-3) / 1e12,
    "conv_train_step_ms": ms_train, "conv_train_step_TFLOPs": 3 * B * f_fwd / (ms_train * 1e-3) / 1e12,
    "conv_note": "implicit-GEMM fp32 FFMA convolutions (csrc/conv.cu), u8 observations dequantised inside the operand fetch; one update = PER sample + "
                 "target forward + priorities + train step + polyak; train step = forward + backward (2 x forward FLOP) + ClipValue/Adam"})
del buf, tgt, Sq, Dq
torch.cuda.empty_cache()

# ------------------------------------------------------------------ C4: SAC update, Humanoid-shaped
obs, act, hid, Bs = 376, 17, 256, 2048
rng = np.random.default_rng(4)
D = crux.Dense
Apol = crux.SquashedGaussianPolicy(crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, 2 * act, rng=rng)), ctx=ctx))
Q = lambda: crux.ContinuousNetwork(crux.Chain(D(obs + act, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, 1, rng=rng)), ctx=ctx)
pi = crux.ActorCritic(Apol, crux.DoubleNetwork(Q(), Q()))
S4 = crux.SAC(pi, crux.ContinuousSpace(obs), N=10, dN=1, c_opt=dict(batch_size=Bs, epochs=1), buffer_size=Bs, buffer_init=Bs)
s = torch.randn((Bs, obs), device=ctx.device, generator=g)
a = torch.tanh(torch.randn((Bs, act), device=ctx.device, generator=g))
sp = torch.randn((Bs, obs), device=ctx.device, generator=g)
r = torch.randn(Bs, device=ctx.device, generator=g)
dn = (torch.rand(Bs, device=ctx.device, generator=g) < 0.01).to(torch.uint8)
k = [0]


def sac_step():
    k[0] += 1
    ctx.check(ctx.lib.crux_sac_train(S4._sac, ptr(s), ptr(a), ptr(sp), ptr(r), ptr(dn), Bs, F32(0.99), None, None, None, 4, 3 * k[0], None, None))


ms_sac = timed(sac_step, reps=20)
fa = 2 * (obs * hid + hid * hid + hid * 2 * act)          # actor forward FLOP per sample
fq = 2 * ((obs + act) * hid + hid * hid + hid)            # critic forward FLOP per sample
flops = Bs * (5 * fa + 12 * fq)                           # ~5 actor-forward + ~12 critic-forward equivalents per update
out["c4_sac"] = {"batch": Bs, "ms_per_update": ms_sac, "updates_per_s": 1e3 / ms_sac, "samples_per_s": Bs * 1e3 / ms_sac,
                 "approx_TFLOPs": flops / (ms_sac * 1e-3) / 1e12,
                 "note": "generic layer-by-layer fp32 SGEMM engine (mlp.cu); one update = target + temperature + double-Q critics + actor + polyak"}
print(json.dumps(out, indent=1))
