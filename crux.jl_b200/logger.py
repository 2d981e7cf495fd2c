"""Host-side mirror of ``src/logging.jl``: ``LoggerParams`` with the reference's defaults, the built-in ``log_*`` functions and a
TensorBoard-compatible scalar writer standing in for ``TensorBoardLogger.TBLogger(dir, tb_increment)`` (logging.jl:4-9).

The writer emits standard ``events.out.tfevents.*`` files (TFRecord framing with masked CRC32C, ``Event`` / ``Summary`` protobufs
encoded by hand: wall_time, step, one ``simple_value`` per ``log_value`` call), which is all ``log_value(logger, name, v, step=i)``
(logging.jl:52) produces for the reals Crux logs.  Nothing here touches the device: evaluation functions call the sampler's own
metrics (sampler.jl:203-240), which run on the GPU.
"""
from __future__ import annotations

import os
import socket
import struct
import time

import numpy as np

# ------------------------------------------------------------------------------------------------ tfevents records
_CRC_TABLE = None


def _crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), the checksum of the TFRecord framing."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    crc = 0xFFFFFFFF
    for b in data:
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _masked_crc(data: bytes) -> int:
    c = _crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _field(num: int, wire: int, payload: bytes) -> bytes:
    head = _varint((num << 3) | wire)
    return head + (_varint(len(payload)) + payload if wire == 2 else payload)


def _event(wall_time: float, step: int | None = None, file_version: str | None = None, scalars=()) -> bytes:
    """``tensorflow.Event``: 1 wall_time (double), 2 step (int64), 3 file_version (string), 5 summary{1 value{1 tag, 2 simple_value}}."""
    ev = _field(1, 1, struct.pack("<d", wall_time))
    if step is not None:
        ev += _field(2, 0, _varint(int(step)))
    if file_version is not None:
        ev += _field(3, 2, file_version.encode())
    if scalars:
        summ = b"".join(_field(1, 2, _field(1, 2, tag.encode()) + _field(2, 5, struct.pack("<f", float(v)))) for tag, v in scalars)
        ev += _field(5, 2, summ)
    return ev


def _record(data: bytes) -> bytes:
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", _masked_crc(head)) + data + struct.pack("<I", _masked_crc(data))


def tb_increment(logdir: str) -> str:
    """``tb_increment`` of TensorBoardLogger.jl [3P]: an existing ``logdir`` is never touched, the run goes to ``logdir_1``, ``_2``…"""
    logdir = logdir.rstrip("/") or "/"
    if not os.path.exists(logdir):
        return logdir
    i = 1
    while os.path.exists(f"{logdir}_{i}"):
        i += 1
    return f"{logdir}_{i}"


class TBLogger:
    """``TBLogger(dir, tb_increment)`` (logging.jl:8): one event file per run directory, flushed after every record."""

    def __init__(self, logdir="log/", increment=True):
        self.logdir = tb_increment(logdir) if increment else (logdir.rstrip("/") or "/")
        os.makedirs(self.logdir, exist_ok=True)
        self.path = os.path.join(self.logdir, f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}.{os.getpid()}")
        self._f = open(self.path, "ab")
        self._f.write(_record(_event(time.time(), file_version="brain.Event:2")))
        self._f.flush()

    def log_value(self, name, value, step=0):
        """``log_value(logger, name, value; step)``: one scalar summary (Bool and integers are logged as their float value)."""
        self._f.write(_record(_event(time.time(), step=step, scalars=[(str(name), float(value))])))
        self._f.flush()

    def close(self):
        if not self._f.closed:
            self._f.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def read_scalars(path):
    """Reads an event file back (checking both CRCs of every record) -> list of (step, tag, value).  Used by the tests and by
    anyone who wants the curves without TensorBoard."""
    out = []
    with open(path, "rb") as f:
        buf = f.read()
    pos = 0

    def varint(b, p):
        n = s = 0
        while True:
            c = b[p]
            p += 1
            n |= (c & 0x7F) << s
            s += 7
            if not c & 0x80:
                return n, p

    def fields(b):
        p = 0
        while p < len(b):
            key, p = varint(b, p)
            num, wire = key >> 3, key & 7
            if wire == 0:
                v, p = varint(b, p)
            elif wire == 1:
                v, p = b[p:p + 8], p + 8
            elif wire == 5:
                v, p = b[p:p + 4], p + 4
            elif wire == 2:
                n, p = varint(b, p)
                v, p = b[p:p + n], p + n
            else:
                raise ValueError(f"unsupported wire type {wire}")
            yield num, wire, v

    while pos < len(buf):
        head = buf[pos:pos + 8]
        (n,) = struct.unpack("<Q", head)
        (hc,) = struct.unpack("<I", buf[pos + 8:pos + 12])
        data = buf[pos + 12:pos + 12 + n]
        (dc,) = struct.unpack("<I", buf[pos + 12 + n:pos + 16 + n])
        if hc != _masked_crc(head) or dc != _masked_crc(data):
            raise ValueError(f"corrupt record at byte {pos} of {path}")
        pos += 16 + n
        step = 0
        summaries = []
        for num, _, v in fields(data):
            if num == 2:
                step = v
            elif num == 5:
                summaries.append(v)
        for s in summaries:
            for num, _, val in fields(s):
                if num != 1:
                    continue
                tag, x = None, None
                for k, _, vv in fields(val):
                    if k == 1:
                        tag = vv.decode()
                    elif k == 2:
                        (x,) = struct.unpack("<f", vv)
                out.append((step, tag, x))
    return out


# ------------------------------------------------------------------------------------------------ LoggerParams
class LoggerParams:
    """logging.jl:12-25.  Defaults as the reference: ``dir = "log/"``, ``period = 500``, ``logger = TBLogger(dir, tb_increment)``,
    ``fns = [log_undiscounted_return(10), log_episode_averages([:r], period)]``, ``verbose = true``.  ``use_wandb`` raises like
    ``build_logger`` (:4-7).  ``logger=None`` keeps the records in ``history`` only (no files), which the parity tests and the
    benchmark use; every record is appended to ``history`` either way."""

    def __init__(self, dir="log/", period=500, use_wandb=False, config=None, project=None, entity=None, notes=None, logger="tb",
                 fns=None, writeout=None, verbose=True, sampler=None):
        if use_wandb:
            raise RuntimeError("Please run `using WeightsAndBiasLogger`")      # logging.jl:6
        self.dir, self.period, self.verbose, self.sampler = dir, int(period), verbose, sampler
        self.config, self.project, self.entity, self.notes = config, project, entity, notes
        self.logger = TBLogger(dir, increment=True) if logger == "tb" else logger
        self.fns = [log_undiscounted_return(10), log_episode_averages(["r"], self.period)] if fns is None else list(fns)
        self.writeout = dict(writeout or {})                                    # period => fn(i, s, dir, logger)
        self.history = []

    @staticmethod
    def elapsed(i, N):
        """logging.jl:1-2 (``i`` an int or an inclusive (lo, hi) range)."""
        if isinstance(i, tuple):
            lo, hi = i
            return hi // N > (lo - 1) // N
        return i % N == 0

    def log(self, i, *data, solver=None):
        """``Base.log(p::LoggerParams, i, data...; 𝒮)`` logging.jl:30-58."""
        last = i[1] if isinstance(i, tuple) else i
        for period, fn in self.writeout.items():                                # :33-35
            if self.elapsed(i, period):
                fn(i=last, s=self.sampler, dir=self.dir, logger=self.logger)
        if not self.elapsed(i, self.period):                                    # :38
            return
        step = last
        dicts = list(self.fns) + list(data)
        if self.sampler is not None:                                            # :43-46
            s0 = self.sampler[0] if isinstance(self.sampler, (list, tuple)) else self.sampler
            dicts.append(log_exploration(s0.agent.pi_explore))
        rec = {"step": step}
        for d in dicts:                                                         # :48-54
            d = d(s=self.sampler, i=step, solver=solver) if callable(d) else d
            for k, v in d.items():
                v = v() if callable(v) else v
                rec[str(k)] = v
                if self.logger is not None:
                    self.logger.log_value(str(k), v, step=step)
        self.history.append(rec)
        if self.verbose:
            print(f"Step: {step}" + "".join(f", {k}: {v}" for k, v in rec.items() if k != "step"))


def aggregate_info(infos):
    """logging.jl:60-66: per-key mean over the dicts that have the key."""
    keys = []
    for info in infos:
        keys += [k for k in info if k not in keys]
    return {k: float(np.mean([info[k] for info in infos if k in info])) for k in keys}


# ------------------------------------------------------------------------------------------------ built-in log functions
def log_performance(s, name, fn, **kw):
    """logging.jl:69-70: one entry per sampler of a vector (``name/T{i}``, 1-based) or a single entry."""
    if isinstance(s, (list, tuple)):
        return {f"{name}/T{j + 1}": fn(sj, **kw) for j, sj in enumerate(s)}
    return {name: fn(s, **kw)}


def log_discounted_return(Neps):
    """logging.jl:72."""
    return lambda s, **_: log_performance(s, "discounted_return", lambda x, **kw: x.discounted_return(**kw), Neps=Neps)


def log_undiscounted_return(*args, name="undiscounted_return"):
    """logging.jl:73-74: ``log_undiscounted_return(Neps)`` evaluates the logger's sampler, ``log_undiscounted_return(s, Neps)`` a given one
    (greedy episodes on a reset sampler)."""
    if len(args) == 2:
        s_fixed, Neps = args
        return lambda **_: log_performance(s_fixed, name, lambda x, **kw: x.undiscounted_return(**kw), Neps=Neps)
    (Neps,) = args or (10,)
    return lambda s, **_: log_performance(s, name, lambda x, **kw: x.undiscounted_return(**kw), Neps=Neps)


def log_failure(Neps):
    """logging.jl:75."""
    return lambda s, **_: log_performance(s, "failure_rate", lambda x, **kw: x.failure(**kw), Neps=Neps)


def log_metric_by_key(key, Neps):
    """logging.jl:76."""
    return lambda s, **_: log_performance(s, str(key), lambda x, **kw: x.metric_by_key(key, **kw), Neps=Neps)


def log_metrics_by_key(keys, Neps, **kw):
    """logging.jl:78-83."""
    return lambda s, **_: dict(zip(keys, s.metrics_by_key(list(keys), Neps=Neps, **kw)))


def log_validation_error(loss, D_val, name="validation_error"):
    """logging.jl:85."""
    return lambda s, **_: {name: loss(s.agent.pi, D_val)}


def log_exploration(policy, name=None):
    """logging.jl:87-96: ε of an ε-greedy policy, σ of a Gaussian-noise policy, the first-explore switch (+ its after-policy's entry
    evaluated at i = 1, as the reference does), nothing otherwise."""
    from .policies import FirstExplorePolicy, GaussianNoiseExplorationPolicy, MixedPolicy
    if isinstance(policy, MixedPolicy):
        return lambda i, **_: {name or "eps": float(policy.eps(i))}
    if isinstance(policy, GaussianNoiseExplorationPolicy):
        return lambda i, **_: {name or "noise_std": float(policy.sigma(i))}
    if isinstance(policy, FirstExplorePolicy):
        def f(i, **kw):
            d = {name or "first_explore_on": i < policy.N}
            if policy.after_policy is not None:
                d.update(log_exploration(policy.after_policy)(i=1, **kw))
            return d
        return f
    return lambda **_: {}


def _last_period_sum(buffer, key, idx0):
    import torch
    col = buffer[key]
    sel = torch.as_tensor(idx0, device=col.device)
    return float(col.index_select(0, sel).to(torch.float64).sum().item())


def log_episode_averages(keys, period):
    """logging.jl:99-112: ``sum(buffer[k][last period rows]) / sum(buffer[:episode_end][same rows])`` as ``avg_<k>``."""
    def f(solver=None, **_):
        d = {}
        buf = getattr(solver, "buffer", None)
        if buf is not None and len(buf) > 0:
            idx0 = np.asarray(buf.get_last_N_indices(period)) - 1
            ends = _last_period_sum(buf, "episode_end", idx0)
            for k in keys:
                d[f"avg_{k}"] = _last_period_sum(buf, k, idx0) / ends if ends else float("inf")
        return d
    return f


def log_experience_sums(keys, period):
    """logging.jl:114-127 (the reference names the entries ``avg_<k>`` too; kept)."""
    def f(solver=None, **_):
        d = {}
        buf = getattr(solver, "buffer", None)
        if buf is not None and len(buf) > 0:
            idx0 = np.asarray(buf.get_last_N_indices(period)) - 1
            for k in keys:
                d[f"avg_{k}"] = _last_period_sum(buf, k, idx0)
        return d
    return f
