"""Multi-GPU parity check, launched with torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu_check.py
Checks: NCCL + one-shot peer all-reduce, data-parallel PPO update == the oracle on the union batch, global whitening."""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import crux_b200 as crux  # noqa: E402
from crux_b200.device import ptr  # noqa: E402
from oracle import crux_oracle as o  # noqa: E402

F32 = np.float32


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = crux.Context(local)
    ctx.init_distributed(rank, world, peer_floats=0)   # NCCL only first; the peer buffers are mapped further down
    dev = lambda x, dt=None: ctx.to_device(x, dt)

    def allreduce_check(tag, iters):
        for it in range(iters):
            n = 11085 + 128
            x = np.random.default_rng(100 * it + rank).standard_normal(n).astype(F32)
            t = dev(x)
            ctx.check(ctx.lib.crux_nccl_allreduce_f32(ctx.h, ptr(t), n))
            want = sum(np.random.default_rng(100 * it + r).standard_normal(n).astype(np.float64) for r in range(world))
            got = t.cpu().numpy()
            assert np.allclose(got, want, rtol=1e-5, atol=1e-5), f"{tag} all-reduce mismatch at iteration {it}"
            gl = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(gl, t)
            assert all(torch.equal(gl[0], g) for g in gl), f"{tag}: ranks disagree bitwise"

    allreduce_check("nccl", 5)

    # ---- data-parallel PPO update vs the oracle on the union batch
    n_loc, ab, epochs = 512, 128, 2
    rng = np.random.default_rng(0)
    mu = o.MLP([17, 64, 64, 6], [1, 1, 0], rng)
    cr = o.MLP([17, 64, 64, 1], [1, 1, 0], rng)
    ls = np.full(6, -0.5, F32)
    pi = o.GaussianPolicy(mu, ls)
    n = n_loc * world
    s = rng.standard_normal((n, 17)).astype(F32)
    old = o.GaussianPolicy(o.MLP(mu.dims, mu.acts, Ws=[w.detach().numpy() * F32(0.97) for w in mu.W], bs=[b.detach().numpy() for b in mu.b]), ls - F32(0.05))
    a, lp = old.exploration(s, rng.standard_normal((n, 6)).astype(F32))
    D = {"s": s, "a": a.detach().numpy(), "logprob": lp.detach().numpy()[:, 0], "advantage": rng.standard_normal(n).astype(F32),
         "return": rng.standard_normal(n).astype(F32)}
    orders_a = [[rng.permutation(n_loc) for _ in range(epochs)] for _ in range(world)]
    orders_c = [[rng.permutation(n_loc) for _ in range(epochs)] for _ in range(world)]
    # oracle: global minibatch k of epoch e = union over ranks of rank-local rows order[r][e][k*ab:(k+1)*ab] (offset r*n_loc)
    P = {"eps": F32(0.2), "lp": F32(1.0), "le": F32(0.1)}
    opt_a, opt_c = o.Adam(F32(3e-4)), o.Adam(F32(3e-4))
    for e in range(epochs):
        for k in range(n_loc // ab):
            idx = np.concatenate([r * n_loc + orders_a[r][e][k * ab:(k + 1) * ab] for r in range(world)])
            mb = {kk: v[idx] for kk, v in D.items()}
            o.train_step(pi.params(), lambda inf, mb=mb: o.ppo_loss(pi, P, mb, inf), opt_a, {})
    for e in range(epochs):
        for k in range(n_loc // ab):
            idx = np.concatenate([r * n_loc + orders_c[r][e][k * ab:(k + 1) * ab] for r in range(world)])
            mb = {kk: v[idx] for kk, v in D.items()}
            o.train_step(cr.params(), lambda inf, mb=mb: o.value_mse_loss(cr, mb), opt_c, {})
    # device: every rank starts from the same parameters and sees only its shard
    def ppo_check(tag):
        rng0 = np.random.default_rng(0)
        mu0 = o.MLP([17, 64, 64, 6], [1, 1, 0], rng0); cr0 = o.MLP([17, 64, 64, 1], [1, 1, 0], rng0)

        def net(m):
            return crux.ContinuousNetwork(crux.Chain(*[crux.Dense(m.dims[l], m.dims[l + 1], m.acts[l], m.W[l].detach().numpy(), m.b[l].detach().numpy())
                                                       for l in range(3)]), ctx=ctx)
        pol = crux.ActorCritic(crux.GaussianPolicy(net(mu0), ls), net(cr0))
        pol.A.mu.mlp.set_adam(F32(3e-4)); pol.C.mlp.set_adam(F32(3e-4))
        sh = slice(rank * n_loc, (rank + 1) * n_loc)
        d = {k: dev(v[sh]) for k, v in D.items()}
        hp = crux._abi.PPOHp(eps_clip=0.2, lambda_p=1.0, lambda_e=0.1, target_kl=math.inf, a2c=0, actor_epochs=epochs, actor_batch=ab,
                             critic_epochs=epochs, critic_batch=ab, actor_max_batches=0, critic_max_batches=0)
        oa = dev(np.stack(orders_a[rank]).astype(np.int32), torch.int32)
        oc = dev(np.stack(orders_c[rank]).astype(np.int32), torch.int32)
        ia = np.zeros((epochs * (n_loc // ab), 8), F32); ic = np.zeros_like(ia)
        ctx.check(ctx.lib.crux_ppo_update(pol.A.h, pol.C.mlp.h, ptr(d["s"]), ptr(d["a"]), ptr(d["logprob"]), ptr(d["advantage"]), ptr(d["return"]),
                                          n_loc, C.byref(hp), ptr(oa), ptr(oc), 0, ptr(ia), ptr(ic)))
        assert ia[:, 7].all() and ic[:, 7].all(), f"{tag}: info records not all valid"
        for name, got, want in (("actor", pol.A.mu.mlp.get_flat(), mu.flat()), ("critic", pol.C.mlp.get_flat(), cr.flat())):
            err = np.abs(got - want)
            assert err.max() < 2 * 3e-4 * 8 + 2e-6, f"{tag} {name}: {err.max()}"
            assert (err > 2e-6 + 1e-5 * np.abs(want)).mean() < 2e-3, f"{tag} {name}: {(err > 2e-6 + 1e-5 * np.abs(want)).sum()} coordinates off"
            t = dev(got); gl = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(gl, t)
            assert all(torch.equal(gl[0], g) for g in gl), f"{tag} {name}: replicas diverged"
        # early stop must be taken identically on every rank (the KL comes from the all-reduced sums)
        hp2 = crux._abi.PPOHp(eps_clip=0.2, lambda_p=1.0, lambda_e=0.1, target_kl=1e-9, a2c=0, actor_epochs=epochs, actor_batch=ab,
                              critic_epochs=0, critic_batch=ab, actor_max_batches=0, critic_max_batches=0)
        ctx.check(ctx.lib.crux_ppo_update(pol.A.h, pol.C.mlp.h, ptr(d["s"]), ptr(d["a"]), ptr(d["logprob"]), ptr(d["advantage"]), ptr(d["return"]),
                                          n_loc, C.byref(hp2), ptr(oa), ptr(oc), 0, ptr(ia), ptr(ic)))
        v = torch.tensor(ia[:, 7].copy(), device=ctx.device); gl = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(gl, v)
        assert all(torch.equal(gl[0], g) for g in gl) and ia[0, 7] == 1 and ia[-1, 7] == 0, f"{tag}: early stop diverged {ia[:, 7]}"
        if rank == 0:
            print(f"ppo update parity OK ({tag})")

    ppo_check("nccl all-reduce")

    # ---- data-parallel SAC update (replicas with per-rank replay shards + gradient all-reduce): every rank trains on its shard
    # of a union minibatch with injected noise; parameters, targets and log α must equal the oracle's update on the union batch
    def sac_check(tag):
        import test_gpu_offpolicy as T
        from gpu_util import mlp_params, p as gp, dev as gdev
        sdim, A, H, B = 11, 3, 32, 64
        rng_s, nets, hs, pol = T._sac_setup(ctx, sdim, A, H, seed=77)
        actor, q1, q2, q1t, q2t = nets
        pis = o.SquashedGaussianPolicy(lambda s_: actor(s_)[:, :A], lambda s_: actor(s_)[:, A:], 1.0, actor.params())
        log_alpha0, h_target, tau, gamma = F32(math.log(0.2)), F32(-A), F32(0.005), F32(0.99)
        st = C.c_void_p()
        ctx.check(ctx.lib.crux_sac_create(pol, hs[1], hs[2], hs[3], hs[4], log_alpha0, h_target, float(F32(3e-4)), tau, C.byref(st)))
        log_alpha = torch.tensor([float(log_alpha0)], requires_grad=True)
        opt_t, opt_c, opt_a = o.Adam(F32(3e-4)), o.Adam(F32(3e-4)), o.Adam(F32(3e-4))
        n_u = B * world
        for step in range(2):
            Du = {"s": rng_s.standard_normal((n_u, sdim)).astype(F32), "a": np.tanh(rng_s.standard_normal((n_u, A))).astype(F32),
                  "sp": rng_s.standard_normal((n_u, sdim)).astype(F32), "r": rng_s.standard_normal(n_u).astype(F32),
                  "done": (rng_s.random(n_u) < 0.2).astype(np.uint8)}
            e1, e2, e3 = (rng_s.standard_normal((n_u, A)).astype(F32) for _ in range(3))
            y = o.sac_target(pis, q1t, q2t, Du, gamma, float(log_alpha.detach()[0]), e1)
            o.train_step([log_alpha], lambda inf: o.sac_temp_loss(pis, Du, log_alpha[0], h_target, e2), opt_t, {}, "temp_")
            sa = np.concatenate([Du["s"], Du["a"]], 1)
            o.train_step(q1.params() + q2.params(), lambda inf: 0.5 * (o.td_loss(q1(sa), y, None, inf, "Q1avg") + o.td_loss(q2(sa), y, None, inf, "Q2avg")),
                         opt_c, {}, "critic_")
            o.train_step(actor.params(), lambda inf: o.sac_actor_loss(pis, q1, q2, Du, float(log_alpha.detach()[0]), e3, inf), opt_a, {}, "actor_")
            o.polyak_average(q1t.params(), q1.params(), tau)
            o.polyak_average(q2t.params(), q2.params(), tau)
            shd = slice(rank * B, (rank + 1) * B)
            ctx.check(ctx.lib.crux_sac_train(st, gp(gdev(ctx, Du["s"][shd])), gp(gdev(ctx, Du["a"][shd])), gp(gdev(ctx, Du["sp"][shd])), gp(gdev(ctx, Du["r"][shd])),
                                             gp(gdev(ctx, Du["done"][shd])), B, gamma, gp(gdev(ctx, e1[shd])), gp(gdev(ctx, e2[shd])), gp(gdev(ctx, e3[shd])),
                                             0, 0, None, None))
            la = np.zeros(1, F32); ctx.check(ctx.lib.crux_sac_log_alpha(st, gp(la)))
            assert abs(la[0] - float(log_alpha.detach()[0])) < 1e-5, f"{tag}: log alpha {la[0]} vs {float(log_alpha.detach()[0])}"
            for name, h, m in (("q1", hs[1], q1), ("q2", hs[2], q2), ("actor", hs[0], actor), ("q1 target", hs[3], q1t)):
                got, want = mlp_params(ctx, h), m.flat()
                err = np.abs(got - want)
                assert err.max() < 2 * 3e-4 * (step + 1) + 2e-6, f"{tag} {name}: {err.max()}"
                assert (err > 2e-6 + 1e-5 * np.abs(want)).mean() < 5e-3, f"{tag} {name}: {(err > 2e-6 + 1e-5 * np.abs(want)).sum()} coordinates off"
                t_ = dev(got); gl_ = [torch.empty_like(t_) for _ in range(world)]
                dist.all_gather(gl_, t_)
                assert all(torch.equal(gl_[0], g) for g in gl_), f"{tag} {name}: replicas diverged"
        # device noise: ranks draw different streams, replicas stay identical
        ctx.check(ctx.lib.crux_sac_train(st, gp(gdev(ctx, Du["s"][shd])), gp(gdev(ctx, Du["a"][shd])), gp(gdev(ctx, Du["sp"][shd])), gp(gdev(ctx, Du["r"][shd])),
                                         gp(gdev(ctx, Du["done"][shd])), B, gamma, None, None, None, 5, 9, None, None))
        t_ = dev(mlp_params(ctx, hs[0])); gl_ = [torch.empty_like(t_) for _ in range(world)]
        dist.all_gather(gl_, t_)
        assert all(torch.equal(gl_[0], g) for g in gl_), f"{tag}: replicas diverged with device noise"
        ctx.lib.crux_sac_destroy(st)
        if rank == 0:
            print(f"sac update parity OK ({tag})")

    def ddpg_check(tag):
        import test_gpu_offpolicy as T
        from gpu_util import mlp_params, p as gp, dev as gdev
        sdim, A, H, B = 9, 2, 32, 64
        rng_d, actor, at, crit, ct, ha, hat, hc, hct = T._ddpg_setup(ctx, sdim, A, H, True, seed=31)
        st = C.c_void_p()
        ctx.check(ctx.lib.crux_ddpg_create(ha, hat, hc[0], hct[0], hc[1], hct[1], F32(0.005), C.byref(st)))
        n_u = B * world
        Du = [rng_d.standard_normal((n_u, sdim)).astype(F32), np.tanh(rng_d.standard_normal((n_u, A))).astype(F32), rng_d.standard_normal((n_u, sdim)).astype(F32),
              rng_d.standard_normal(n_u).astype(F32), (rng_d.random(n_u) < 0.2).astype(np.uint8)]
        es = rng_d.standard_normal((n_u, A)).astype(F32)

        def run(data, eps, Bx):
            ctx.check(ctx.lib.crux_ddpg_train(st, *[gp(gdev(ctx, x)) for x in data], Bx, F32(0.99), 1, F32(0.2), F32(-0.5), F32(0.5), None, 0, None, 0,
                                              gp(gdev(ctx, eps)), 0, 0, 1, 1, None, None))
        shd = slice(rank * B, (rank + 1) * B)
        run([x[shd] for x in Du], es[shd], B)
        sharded = [mlp_params(ctx, h).copy() for h in (ha, hc[0], hc[1], hat)]
        for t_ in sharded:
            tt = dev(t_); gl_ = [torch.empty_like(tt) for _ in range(world)]
            dist.all_gather(gl_, tt)
            assert all(torch.equal(gl_[0], g) for g in gl_), f"{tag}: replicas diverged"
        # the oracle's TD3 update (off_policy.jl:71-101 order) on the union batch
        y = o.ddpg_target(at, ct, {"sp": Du[2], "r": Du[3], "done": Du[4]}, F32(0.99),
                          dict(eps=es, sigma=F32(0.2), eps_min=-0.5, eps_max=0.5, a_min=-math.inf, a_max=math.inf))
        sa = np.concatenate([Du[0], Du[1]], 1)
        oc_, oa_ = o.Adam(F32(3e-4)), o.Adam(F32(3e-4))
        o.train_step(crit[0].params() + crit[1].params(), lambda inf: 0.5 * (o.td_loss(crit[0](sa), y, None, inf, "Q1avg") + o.td_loss(crit[1](sa), y, None, inf, "Q2avg")),
                     oc_, {}, "critic_")
        o.train_step(actor.params(), lambda inf: o.ddpg_actor_loss(actor, crit[0], {"s": Du[0]}), oa_, {}, "actor_")
        for name, got, m in (("actor", sharded[0], actor), ("q1", sharded[1], crit[0]), ("q2", sharded[2], crit[1])):
            err = np.abs(got - m.flat())
            assert err.max() < 2 * 3e-4 + 2e-6 and (err > 2e-6 + 1e-5 * np.abs(m.flat())).mean() < 5e-3, f"{tag} {name}: {err.max()}"
        ctx.lib.crux_ddpg_destroy(st)
        if rank == 0:
            print(f"td3 update parity OK ({tag})")

    # ---- data-parallel LagrangePPO: the PID cost estimate sum(cost)/sum(episode_end) is that of the UNION minibatch (summed over ranks in front
    # of the PID step), the three networks' gradients are all-reduced like PPO's; parameters, PID state and the info records == the oracle's
    def lagrange_check(tag):
        rngl = np.random.default_rng(31)
        mul = o.MLP([17, 64, 64, 6], [1, 1, 0], rngl); crl = o.MLP([17, 64, 64, 1], [1, 1, 0], rngl); vcl = o.MLP([17, 64, 64, 1], [1, 1, 0], rngl)
        flat0 = [m.flat().copy() for m in (mul, crl, vcl)]
        pil = o.GaussianPolicy(mul, ls.copy())
        Dl = dict(D)
        Dl["cost"] = (rngl.random(n) < 0.3).astype(F32) * rngl.random(n).astype(F32)
        Dl["cost_advantage"] = rngl.standard_normal(n).astype(F32)
        Dl["cost_return"] = rngl.standard_normal(n).astype(F32)
        Dl["episode_end"] = rngl.random(n) < 0.1
        ok_ = [[rngl.permutation(n_loc) for _ in range(epochs)] for _ in range(world)]
        Pl = o.lagrange_params(eps=0.2, lp=1.0, le=0.1, target_cost=0.025, Ki=0.05, Kp=1, Kd=0.5)
        recs = []

        def union(orders, e, k):
            idx = np.concatenate([r * n_loc + orders[r][e][k * ab:(k + 1) * ab] for r in range(world)])
            return {kk: v[idx] for kk, v in Dl.items()}
        opts = [o.Adam(F32(3e-4)) for _ in range(3)]
        for e in range(epochs):
            for k in range(n_loc // ab):
                inf = {}
                o.train_step(pil.params(), lambda i_, mb=union(orders_a, e, k): o.lagrange_ppo_loss(pil, Pl, mb, i_), opts[0], inf)
                recs.append(inf)
        for e in range(epochs):
            for k in range(n_loc // ab):
                o.train_step(crl.params(), lambda i_, mb=union(orders_c, e, k): o.value_mse_loss(crl, mb), opts[1], {})
        for e in range(epochs):
            for k in range(n_loc // ab):
                o.train_step(vcl.params(), lambda i_, mb=union(ok_, e, k): o.value_mse_loss(vcl, dict(mb, **{"return": mb["cost_return"]})), opts[2], {})

        def net(dims, acts, flat):
            ws, off = [], 0
            layers = []
            for l in range(3):
                nw = dims[l] * dims[l + 1]
                W = flat[off:off + nw].reshape(dims[l], dims[l + 1]).T.copy(); off += nw      # flat holds Julia memory order [in][out]
                b = flat[off:off + dims[l + 1]].copy(); off += dims[l + 1]
                layers.append(crux.Dense(dims[l], dims[l + 1], acts[l], W, b))
            return crux.ContinuousNetwork(crux.Chain(*layers), ctx=ctx)
        am, cm, km = net(mul.dims, mul.acts, flat0[0]), net(crl.dims, crl.acts, flat0[1]), net(vcl.dims, vcl.acts, flat0[2])
        assert np.array_equal(am.mlp.get_flat(), flat0[0]), "network construction order"
        pol = crux.ActorCritic(crux.GaussianPolicy(am, ls.copy()), cm)
        for m_ in (am, cm, km):
            m_.mlp.set_adam(F32(3e-4))
        sh = slice(rank * n_loc, (rank + 1) * n_loc)
        d = {k: dev(v[sh].astype(np.uint8) if v.dtype == bool else v[sh]) for k, v in Dl.items()}
        hp = crux._abi.PPOHp(eps_clip=0.2, lambda_p=1.0, lambda_e=0.1, target_kl=math.inf, a2c=0, actor_epochs=epochs, actor_batch=ab,
                             critic_epochs=epochs, critic_batch=ab, actor_max_batches=0, critic_max_batches=0)
        lhp = crux._abi.LagrangeHp(target_cost=0.025, penalty_max=math.inf, Ki_max=10.0, Ki=0.05, Kp=1.0, Kd=0.5, ema_alpha=0.95,
                                   cost_epochs=epochs, cost_batch=ab, cost_max_batches=0)
        oa = dev(np.stack(orders_a[rank]).astype(np.int32), torch.int32); oc = dev(np.stack(orders_c[rank]).astype(np.int32), torch.int32)
        okd = dev(np.stack(ok_[rank]).astype(np.int32), torch.int32)
        nm = epochs * (n_loc // ab)
        ia, il, ic, ik = (np.zeros((nm, 8), F32) for _ in range(4))
        state = dev(np.zeros(5, F32))
        ctx.check(ctx.lib.crux_lagrange_ppo_update(pol.A.h, cm.mlp.h, km.mlp.h, ptr(d["s"]), ptr(d["a"]), ptr(d["logprob"]), ptr(d["advantage"]), ptr(d["return"]),
                                                   ptr(d["cost"]), ptr(d["cost_advantage"]), ptr(d["cost_return"]), ptr(d["episode_end"]), n_loc, C.byref(hp),
                                                   C.byref(lhp), ptr(state), ptr(oa), ptr(oc), ptr(okd), 0, ptr(ia), ptr(ic), ptr(il), ptr(ik)))
        assert any(r_["penalty"] > 0 for r_ in recs), "the check must exercise a non-zero penalty"
        for k, r_ in enumerate(recs):
            want = np.array([r_["penalty"], r_["cur_cost"], r_["prop_term"], r_["deriv_term"], r_["integral term"], r_["p_loss"], r_["cost_loss"]], F32)
            assert il[k, 7] == 1.0 and np.allclose(il[k, :7], want, rtol=3e-4, atol=3e-6), f"{tag}: lagrange record {k}: {il[k, :7]} vs {want}"
        st = state.cpu().numpy()
        want_st = np.array([Pl["I"], Pl["smooth_D"], Pl["smooth_Jc"], Pl["Jc_prev"], recs[-1]["penalty"]], F32)
        assert np.allclose(st, want_st, rtol=3e-4, atol=3e-6), f"{tag}: PID state {st} vs {want_st}"
        for name, got, want in (("actor", am.mlp.get_flat(), mul.flat()), ("critic", cm.mlp.get_flat(), crl.flat()), ("cost critic", km.mlp.get_flat(), vcl.flat())):
            err = np.abs(got - want)
            assert err.max() < 2 * 3e-4 * 8 + 2e-6, f"{tag} {name}: {err.max()}"
            assert (err > 2e-6 + 1e-5 * np.abs(want)).mean() < 5e-3, f"{tag} {name}: {(err > 2e-6 + 1e-5 * np.abs(want)).sum()} coordinates off"
            t_ = dev(got); gl_ = [torch.empty_like(t_) for _ in range(world)]
            dist.all_gather(gl_, t_)
            assert all(torch.equal(gl_[0], g) for g in gl_), f"{tag} {name}: replicas diverged"
        gs = [torch.empty_like(state) for _ in range(world)]
        dist.all_gather(gs, state)
        assert all(torch.equal(gs[0], g) for g in gs), f"{tag}: PID state differs between ranks"
        if rank == 0:
            print(f"lagrange ppo update parity OK ({tag})")

    sac_check("nccl all-reduce")
    ddpg_check("nccl all-reduce")
    lagrange_check("nccl all-reduce")

    # ---- global whitening: every rank whitens its shard with the all-reduced moments
    x = np.random.default_rng(5).standard_normal(4000).astype(F32) * 3 + 1
    t = dev(x[rank::world].copy())
    ctx.check(ctx.lib.crux_whiten(ctx.h, ptr(t), t.numel()))
    assert np.allclose(t.cpu().numpy(), o.whiten(x)[rank::world], rtol=1e-4, atol=1e-5), "whiten"

    # ---- one-shot peer all-reduce over NVLink (CUDA IPC), many back-to-back calls (double-buffer race check)
    hb = (C.c_uint8 * 64)()
    ctx.check(ctx.lib.crux_peer_handle(ctx.h, hb, 16384))
    handles = [None] * world
    dist.all_gather_object(handles, bytes(hb))
    allh = (C.c_uint8 * (64 * world)).from_buffer_copy(b"".join(handles))
    dist.barrier()
    ctx.check(ctx.lib.crux_peer_init(ctx.h, rank, world, allh))
    dist.barrier()
    allreduce_check("peer", 50)
    ppo_check("fused LL peer exchange")      # reduce kernel stores {value, seq} words into every rank's region, Adam kernel sums them
    ppo_check("fused LL peer exchange, 2nd") # after early-stopped (skipped) exchanges: the device sequence numbers stayed in step
    os.environ["CRUX_NO_PEER_LL"] = "1"
    ppo_check("one-shot peer all-reduce between reduce and Adam")
    os.environ.pop("CRUX_NO_PEER_LL")
    ppo_check("fused LL peer exchange, 3rd")
    lagrange_check("one-shot peer all-reduce")
    allreduce_check("peer after ppo", 5)
    # timing of the two paths for the 44 KB gradient payload
    for tag in ("peer",):
        t = dev(np.ones(11213, F32))
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); dist.barrier(); a_.record()
        for _ in range(200):
            ctx.check(ctx.lib.crux_nccl_allreduce_f32(ctx.h, ptr(t), 11213))
        b_.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"{tag} all-reduce of 11213 floats: {a_.elapsed_time(b_) / 200 * 1e3:.1f} us per call")
    dist.barrier()
    if rank == 0:
        print("MULTIGPU OK", world, "ranks")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
