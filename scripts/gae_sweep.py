"""GAE HBM GB/s sweep on the GPU box: python scripts/gae_sweep.py  (22 algorithmic bytes per transition)."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crux_b200 as crux
from crux_b200.device import ptr

ctx = crux.Context(0)
shapes = [(2048, 16384), (1024, 4096), (32, 4096)] if len(sys.argv) < 2 else [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=ctx.device)
for T, N in shapes:
    g = torch.Generator(device=ctx.device).manual_seed(2)
    r, vs, vsp = (torch.randn((T, N), device=ctx.device, generator=g) for _ in range(3))
    done = (torch.rand((T, N), device=ctx.device, generator=g) < 0.001).to(torch.uint8)
    ee = done.clone(); ee[999::1000] = 1; ee[-1] = 1
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    run = lambda: ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, ptr(r), ptr(done), ptr(ee), ptr(vs), ptr(vsp), T, N, 0.99, 0.95, ptr(adv), ptr(ret)))
    cfgs = [("scan", None), ("tma", None), ("tma", "0,0,0,0,1"), ("tma", "0,0,0,0,2")]
    if os.environ.get("SWEEP_SPL"):   # tile widths on narrow rollouts: 32-stream tiles (6th field 1) against 64-stream tiles and the scan
        cfgs = [("scan", None), ("tma", None), ("tma", "0,0,0,0,1,2"), ("tma", "32,6,0,1,1,1"), ("tma", "32,8,0,1,1,1"), ("tma", "64,4,0,1,1,1"),
                ("tma", "64,6,0,1,1,1"), ("tma", "64,7,0,1,1,1"), ("tma", "32,6,0,2,1,1"), ("tma", "64,3,0,2,1,1")]
    elif T >= 64:
        for cs, st, seg, per in itertools.product((16, 32), (3, 4, 6, 7, 8), (0, 16), (1, 2)):
            if st * cs * 128 * 14 * per > 226 * 1024 or os.environ.get("SWEEP_DEFAULT_ONLY"):
                continue
            cfgs.append(("tma", f"{cs},{st},{seg},{per}"))
    ref = None
    for mode, cfg in cfgs:
        os.environ["CRUX_GAE"] = mode
        if cfg: os.environ["CRUX_GAE_CFG"] = cfg
        else: os.environ.pop("CRUX_GAE_CFG", None)
        for _ in range(3): run()
        reps = 20 if T * N > 1 << 22 else 200
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        best, tot = 1e9, 0.0
        for _ in range(reps):
            if T * N * 22 < (200 << 20): flush.zero_()
            a.record(); run(); b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b); best = min(best, ms); tot += ms
        ms = tot / reps
        chk = float(adv.double().sum() + ret.double().sum())
        if ref is None: ref = chk
        print(f"[{T},{N}] {mode:4s} cfg={cfg or '-':14s} avg {ms*1e3:8.1f} us  {22*T*N/ms/1e6:7.1f} GB/s   best {22*T*N/best/1e6:7.1f} GB/s  checksum_delta={chk-ref:+.3e}", flush=True)
