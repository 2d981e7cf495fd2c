"""CPU: host-side logic that needs no GPU -- schemas, schedules, logger periods, the synthetic environments."""
import numpy as np
import pytest


def test_mdp_data_schema(crux):
    # test/experience_buffer_tests.jl:8-20: default columns, extras by name, unknown key
    A = crux._abi
    d1 = crux.mdp_data(crux.ContinuousSpace(3), crux.ContinuousSpace(4), 100)
    assert list(d1) == ["s", "a", "sp", "r", "done", "episode_end"]
    assert d1["s"] == (A.F32, (3,), 0.0) and d1["a"] == (A.F32, (4,), 0.0) and d1["done"][0] == A.U8
    d2 = crux.mdp_data(crux.ContinuousSpace(3), crux.ContinuousSpace(4), 100, ["weight", "t", "advantage", "return", "logprob"])
    assert d2["weight"] == (A.F32, (1,), 1.0) and d2["t"][0] == A.I64 and d2["return"] == (A.F32, (1,), 0.0)
    with pytest.raises(KeyError):
        crux.mdp_data(crux.ContinuousSpace(3), crux.ContinuousSpace(4), 100, ["bad_key"])
    d3 = crux.mdp_data(crux.ContinuousSpace((2, 2), np.uint8), crux.DiscreteSpace(4), 10)  # :271-278
    assert d3["s"] == (A.U8, (2, 2), 0.0) and d3["a"][1] == (4,)


def test_split_batches(crux):
    assert crux.split_batches(100, [0.5, 0.5]) == [50, 50]
    assert crux.split_batches(100, [1 / 3, 1 / 3, 1 / 3]) == [34, 33, 33]
    with pytest.raises(AssertionError):
        crux.split_batches(100, 0.4)


def test_linear_decay_schedule(crux):
    # utils.jl:116-126; test/util_tests.jl:55-86
    l = crux.LinearDecaySchedule(1.0, 0.1, 10)
    assert l(0) == 1.0 and l(31) == 0.1 and abs(l(5) - 0.55) < 1e-12


def test_logger_elapsed(crux):
    # logging.jl:1-2, test/logging_tests.jl
    E = crux.LoggerParams.elapsed
    assert E(500, 500) and not E(499, 500) and E(0, 500)
    assert E((401, 600), 500) and not E((501, 600), 500) and E((1, 500), 500)


def test_spaces(crux):
    # test/spaces_tests.jl:15,24,34-40
    assert crux.tovec(3, crux.DiscreteSpace(4)).tolist() == [False, False, True, False]
    S = crux.ContinuousSpace(1, np.float32, np.float32(1), np.float32(2))
    assert crux.tovec(np.array([0.0], np.float32), S)[0] == np.float32(-0.5)


def test_host_linquad_matches_oracle_spec(crux):
    from oracle import crux_oracle as o
    spec = o.LinQuadSpec(17, 6, 0)
    A, B = crux.linquad_matrices(17, 6, 0)
    assert np.array_equal(A, spec.A) and np.array_equal(B, spec.B)
    env = crux.HostLinQuad(64, seed=3)
    s0 = env.reset().copy()
    assert np.all(np.abs(s0) <= 0.1)
    rng = np.random.default_rng(3)
    rng.random((64, 17), dtype=np.float32)  # the reset draw
    a = np.random.default_rng(1).standard_normal((64, 6)).astype(np.float32)
    sp, r, done = env.step(a)
    xi = rng.standard_normal((64, 17), dtype=np.float32)
    sp0, r0, d0 = spec.step(s0, a, xi)
    assert np.allclose(sp, sp0, atol=1e-6) and np.allclose(r, r0, atol=1e-6) and np.array_equal(done, d0)


def test_gridworld(crux):
    env = crux.SimpleGridWorld(1000, seed=0)
    s = env.reset()
    assert s.min() >= 1 and s.max() <= 10 and env.gamma == np.float32(0.95)
    env.state[:4] = [[4, 3], [4, 6], [9, 3], [8, 8]]
    env.state[4] = [1, 1]
    sp, r, done = env.step(np.zeros(1000, dtype=np.int64))
    assert r[:4].tolist() == [-10.0, -5.0, 10.0, 3.0] and done[:4].all() and (sp[:4] == -1).all()
    assert r[4] == 0 and not done[4]
    # transition statistics: the commanded move succeeds ~70 % of the time
    env.state[:] = [5, 5]
    sp, _, _ = env.step(np.full(1000, 3))  # :right
    assert abs((sp[:, 0] == 6).mean() - 0.7) < 0.05
    # walls keep the agent in place
    env.state[:] = [10, 10]
    sp, _, _ = env.step(np.full(1000, 3))
    assert sp[:, 0].max() == 10


def test_oracle_ppo_iteration_runs():
    from oracle.ppo_cpu import OraclePPO
    p = OraclePPO(8, 16, epochs=2, batch=32)
    D = p.iteration()
    assert D["s"].shape == (128, 17) and D["episode_end"][-8:].all()
    assert len(p.last["actor"]) == 8 and np.isfinite(p.last["actor"][-1]["actor_loss"])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints one JSON line with the keys the driver reads (small steps; CPU only)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CRUX_BENCH_TINY="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "higher_is_better", "cpu_baseline", "e2e", "config"):
        assert k in rec
    assert rec["impl"] == "reference" and rec["value"] > 0 and rec["e2e"]["h2d_bytes_per_step"] == 0


def test_native_host_linquad(crux):
    """C++ host env (csrc/host/linquad_host.cpp): the LinQuad specification with Philox noise, multi-threaded."""
    from oracle import crux_oracle as o
    spec = o.LinQuadSpec(17, 6, 0)
    n = 1000
    env = crux.NativeHostLinQuad(n, seed=9, n_threads=4)
    env1 = crux.NativeHostLinQuad(n, seed=9, n_threads=1)
    s0 = env.reset()
    assert np.array_equal(s0, env1.reset())               # thread count does not change the streams
    assert np.all(np.abs(s0) <= 0.1) and s0.std() > 0.04
    rng = np.random.default_rng(0)
    s = s0
    for t in range(5):
        a = rng.standard_normal((n, 6)).astype(np.float32)
        sp, r, done = env.step(a)
        sp1, r1, d1 = env1.step(a)
        assert np.array_equal(sp, sp1) and np.array_equal(r, r1) and np.array_equal(done, d1)
        mean = s @ spec.A.T + np.tanh(a) @ spec.B.T
        xi = (sp - mean) / 0.01
        assert abs(xi.mean()) < 0.03 and abs(xi.std() - 1) < 0.03
        want_r = 1 - (sp * sp).sum(1) / 17 - 0.1 * (a * a).sum(1) / 6
        assert np.allclose(r, want_r, atol=1e-5) and np.array_equal(done, np.abs(sp[:, 0]) > 5)
        s = sp
    idx = np.array([3, 500, 999], np.int32)
    o3 = env.reset(idx)
    assert o3.shape == (3, 17) and np.all(np.abs(o3) <= 0.1)
    a = np.zeros((n, 6), np.float32)
    sp, _, _ = env.step(a)
    # reset streams restart from their new initial state; the others continue from sp
    assert np.allclose(sp[3], o3[0] @ spec.A.T, atol=0.06) and np.allclose(sp[7], s[7] @ spec.A.T, atol=0.06)


def test_tfevents_writer_round_trip(crux, tmp_path):
    """logging.jl:4-9,52: TBLogger(dir, tb_increment) + log_value.  The file is standard TFRecord framing (length, masked CRC32C of the
    length, payload, masked CRC32C of the payload) around Event protobufs; known-answer checks pin the CRC and the framing."""
    from crux_b200 import logger as L
    assert L._crc32c(b"123456789") == 0xE3069283                     # the CRC-32C check value (RFC 3720 B.4)
    assert L._crc32c(bytes(32)) == 0x8A9136AA                        # RFC 3720 B.4: 32 bytes of zeros
    assert L._masked_crc(b"") == 0xA282EAD8                          # crc 0 -> rotate -> + kMaskDelta
    d = str(tmp_path / "run")
    lg = crux.TBLogger(d)
    assert lg.logdir == d
    lg.log_value("loss", 0.25, step=3)
    lg.log_value("eps", True, step=3)
    lg.log_value("ret", -1.5e3, step=1 << 40)
    lg.close()
    assert crux.read_scalars(lg.path) == [(3, "loss", 0.25), (3, "eps", 1.0), (1 << 40, "ret", -1500.0)]
    raw = open(lg.path, "rb").read()
    import struct
    n0 = struct.unpack("<Q", raw[:8])[0]
    assert b"brain.Event:2" in raw[12:12 + n0]                       # the version record comes first
    # tb_increment: an existing directory is never reused
    assert crux.tb_increment(d) == d + "_1"
    lg2 = crux.TBLogger(d + "/")
    assert lg2.logdir == d + "_1" and crux.tb_increment(d) == d + "_2"
    # a flipped payload byte is detected
    bad = bytearray(raw); bad[-6] ^= 1
    p2 = str(tmp_path / "bad"); open(p2, "wb").write(bytes(bad))
    import pytest
    with pytest.raises(ValueError):
        crux.read_scalars(p2)


def test_logger_params_defaults_and_log(crux, tmp_path):
    """logging.jl:12-25 defaults (period 500, two default fns, verbose) and Base.log (:30-58): writeout periods, elapsed gating, fns +
    data dicts + the exploration entry, aggregate_info (:60-66)."""
    with __import__("pytest").raises(RuntimeError):
        crux.LoggerParams(use_wandb=True, logger=None)
    p = crux.LoggerParams(dir=str(tmp_path / "log"), verbose=False)
    assert p.period == 500 and len(p.fns) == 2 and isinstance(p.logger, crux.TBLogger)
    calls = []

    class FakeSampler:
        class agent:
            pi_explore = crux.eps_greedy_policy(crux.LinearDecaySchedule(1.0, 0.1, 10), [1, 2, 3])
        def undiscounted_return(self, Neps=10):
            calls.append(Neps)
            return 7.0
    p = crux.LoggerParams(dir=str(tmp_path / "log"), period=10, verbose=False, sampler=FakeSampler(),
                          fns=[crux.log_undiscounted_return(5)], writeout={4: lambda **kw: calls.append(("w", kw["i"]))})
    p.log((1, 4), {"x": 1.0})             # writeout fires (4 elapsed), the period has not
    assert calls == [("w", 4)] and p.history == []
    p.log((5, 10), {"x": 2.0}, lambda **kw: {"y": lambda: 3.0})
    sched = crux.LinearDecaySchedule(1.0, 0.1, 10)
    assert p.history == [{"step": 10, "undiscounted_return": 7.0, "x": 2.0, "y": 3.0, "eps": sched(10)}]
    assert calls == [("w", 4), ("w", 10), 5]          # writeout: 8 lies inside 5..10 (it reports the range's last step), then the eval
    assert [(s, t) for s, t, _ in crux.read_scalars(p.logger.path)] == [(10, "undiscounted_return"), (10, "x"), (10, "y"), (10, "eps")]
    assert crux.aggregate_info([{"a": 1.0, "b": 2.0}, {"a": 3.0}]) == {"a": 2.0, "b": 2.0}
    fe = crux.FirstExplorePolicy(100, None, FakeSampler.agent.pi_explore)
    assert crux.log_exploration(fe)(i=5) == {"first_explore_on": True, "eps": crux.LinearDecaySchedule(1.0, 0.1, 10)(1)}


def test_context_helpers_without_a_device(crux, monkeypatch):
    """Host-only helpers of Context must not need a device (the 2-GPU bench died on a missing import inside peer_ll_active in round 2)."""
    from types import SimpleNamespace
    assert crux.Context.peer_ll_active(SimpleNamespace(peer_mapped=True)) is True
    assert crux.Context.peer_ll_active(SimpleNamespace()) is False
    monkeypatch.setenv("CRUX_NO_PEER_LL", "1")
    assert crux.Context.peer_ll_active(SimpleNamespace(peer_mapped=True)) is False
