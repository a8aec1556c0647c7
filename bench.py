#!/usr/bin/env python
"""bench.py — leapfrog-steps/sec of the many-chain NUTS hot path on B200 (BASELINE.json metric).

Headline workload (config.workload): BASELINE.json configs[1] — 1000-dim diagonal Gaussian (sigma_i = exp(lin(-1,1)), mu = 0.5),
1024 chains per GPU, maxdepth = 10, DiagNutsSettings defaults (num_tune = 400), seed 42, x0 ~ N(0,1).
Setup (untimed): nuts_set_position + the 400 tuning draws.  A "step" = one nuts_draw call of DRAWS_PER_STEP post-warmup
draws for every chain.  `value` = leapfrogs / device time of the draw kernel with the draws AND the 15 statistics written to HBM
buffers; `e2e` = the same steps through nuts_draw with pinned HOST buffers (D2H of every draw + all sampler statistics inside the
timed region; the chain state is resident between calls exactly like the reference's Chain, so there is no per-step H2D).

The same JSON line also carries (config.other_configs) device-timed sub-records of BASELINE configs 3, 4 and 5 - config 5 as the
65536-chain job split over the N ranks (strong scaling) -, the tuning phase and a whole default run (400 tuning + 1000 sampling
draws, wall clock through nuts_draw) of the headline workload.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1, one rank per GPU, weak scaling of configs[1])
  python bench.py --impl reference ...                     (the CPU oracle = C++ restatement of nuts-rs on all host cores)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM = 1000
CHAINS_PER_GPU = 1024
MAXDEPTH = 10
NUM_TUNE = 400
SEED = 42
DRAWS_PER_STEP = 50
METRIC = "leapfrog-steps/sec (all chains)"
UNIT = "leapfrog-steps/s"
WORKLOAD = "configs[1]: 1000-dim diagonal Gaussian, 1024 chains per GPU, maxdepth=10, num_tune=400 (untimed), post-warmup draws"

# BASELINE.json configs as concrete inputs (SURVEY.md §8d / BASELINE.md §3).  kind: NUTS_LOGP_* (1 diag Gaussian, 2 rank-1, 3 funnel)
CONFIGS = {
    "c2": dict(index=1, kind=1, dim=1000, chains=1024, num_tune=400, hbm_bound=True,
               what="1000-dim diagonal Gaussian, sigma = exp(lin(-1,1)), mu = 0.5, 1024 chains",
               model=lambda d: dict(mu=0.5, sigma=np.exp(np.linspace(-1.0, 1.0, d)))),
    "c3": dict(index=2, kind=3, dim=10, chains=8192, num_tune=400, hbm_bound=False,
               what="Neal's funnel (10-dim), 8192 chains - divergent tree-depth stress",
               model=lambda d: dict(funnel_scale=3.0)),
    "c4": dict(index=3, kind=1, dim=10000, chains=256, num_tune=1000, hbm_bound=True,
               what="10000-dim ill-conditioned Gaussian (sigma = 10^lin(-3,3)), diag mass-matrix tuning, 256 chains",
               model=lambda d: dict(mu=0.0, sigma=10.0 ** np.linspace(-3.0, 3.0, d))),
    "c5": dict(index=4, kind=2, dim=100, chains=65536, num_tune=400, hbm_bound=False,
               what="100-dim correlated Gaussian (Sigma = I + 0.5 11^T), 65536 chains sharded over the ranks",
               model=lambda d: dict(mu=0.0, rank1_scale=0.5)),
}


_REAL_STDOUT = None


def claim_stdout():
    """Only the JSON line may reach stdout: point fd 1 at stderr for the whole run (NCCL prints its version banner to stdout,
    whatever NCCL_DEBUG_FILE says) and keep the real stdout for emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def model_sigma():
    return np.exp(np.linspace(-1.0, 1.0, DIM))


def initial_positions(nchains, chain_offset, dim=DIM):
    # rows of N(0,1) keyed by the GLOBAL chain id (one generator per block of 1024 chains), so the shards of a multi-GPU run see
    # the rows of the unsharded run without generating everybody else's
    out = np.empty((nchains, dim))
    blk = 1024
    c = chain_offset
    while c < chain_offset + nchains:
        b = c // blk
        rows = np.random.default_rng([SEED, dim, b]).normal(size=(blk, dim))
        hi = min((b + 1) * blk, chain_offset + nchains)
        out[c - chain_offset:hi - chain_offset] = rows[c - b * blk:hi - b * blk]
        c = hi
    return out


def settings():
    from nuts_rs_b200 import _abi

    s = _abi.default_settings()
    s.num_tune = NUM_TUNE
    s.maxdepth = MAXDEPTH
    s.seed = SEED
    return s


def config_settings(cfg, num_tune=None):
    s = settings()
    s.num_tune = cfg["num_tune"] if num_tune is None else num_tune
    return s


def run_config(lib, name, nchains, chain_offset, device, steps, warmup, dps):
    """Device-timed sub-record of one BASELINE config on this rank: set_position + the config's tuning phase, then `steps` launches
    of `dps` post-warmup draws (draws + statistics into HBM buffers).  Returns per-rank numbers; the caller reduces over ranks."""
    import torch

    cfg = CONFIGS[name]
    d = cfg["dim"]
    math = lib.CudaMath(nchains, d, cfg["kind"], device=device, **cfg["model"](d))
    samp = lib.Sampler(math, config_settings(cfg), seed=SEED, chain_id_offset=chain_offset)
    st = samp.set_position(initial_positions(nchains, chain_offset, d))
    samp.draw_device(cfg["num_tune"])
    tune_ms, _ = samp.last_timing()
    lf_tune, _ = samp.counters()
    buf = torch.empty((dps, nchains, d), dtype=torch.float64, device="cuda")
    for _ in range(warmup):
        samp.draw_device(dps, buf.data_ptr())
    math.synchronize()
    lf0, _ = samp.counters()
    ms = 0.0
    for _ in range(steps):
        samp.draw_device(dps, buf.data_ptr())
        ms += samp.last_timing()[0]
    lf1, _ = samp.counters()
    samp.close()
    math.close()
    del buf
    return {"kernel_ms": ms, "leapfrogs": lf1 - lf0, "tune_ms": tune_ms, "tune_leapfrogs": lf_tune, "bad_init": int((st != 0).sum())}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML through pynvml (one
    sample per ~2 ms, the timed region is tens of ms); falls back to polling nvidia-smi when pynvml is unavailable."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device_index):
        self.rows = []  # (sm_mhz, sm_max_mhz, power_w, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap, timestamp)
        self.windows = []  # [t0, t1] of the timed regions
        self.stop = False
        self.idx = device_index
        self.nvml = None
        self.source = "nvidia-smi"
        try:
            import pynvml

            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; NVML's does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if device_index < len(ids) and ids[device_index].isdigit():
                    phys = int(ids[device_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.source = "nvml"
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def sample_nvml(self):
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        return (float(sm), float(mx), pw, bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap))

    def run(self):
        while not self.stop:
            try:
                if self.nvml is not None:
                    self.rows.append(self.sample_nvml() + (time.perf_counter(),))
                    time.sleep(0.002)
                    continue
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                              timeout=5).decode().strip()
                c = [x.strip() for x in out.split(",")]
                self.rows.append((float(c[0]), float(c[1]), float(c[2])) + tuple(x.lower().startswith("active") for x in c[3:7])
                                 + (time.perf_counter(),))
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def summary_between(self, t0, t1):
        rows = [r for r in self.rows if t0 <= r[7] <= t1]
        if not rows:
            return {"sm_mhz": None, "reasons": ["no sample in the window"]}
        sm = sorted(r[0] for r in rows)
        reasons = [name for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6))
                   if any(r[col] for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][1], "reasons": reasons, "samples": len(rows)}

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        # samples taken inside the timed regions (the sampler also runs through the warm-up steps right before them)
        rows = [r for r in self.rows if any(a <= r[7] <= b for a, b in self.windows)] or self.rows
        sm = sorted(r[0] for r in rows)
        reasons = [name for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6))
                   if any(r[col] for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][1], "reasons": reasons, "samples": len(rows),
                "samples_incl_warmup": len(self.rows), "power_w_max": max(r[2] for r in rows), "source": self.source}


def microbench_fixed_depth(lib, _abi, device, chain_offset, depth=6, draws=8):
    """SURVEY §8(d) microbench: fixed step size, unit mass matrix, no adaptation, mindepth = maxdepth (no U-turn checks):
    every draw is exactly 2^depth - 1 leapfrogs + checkpoints, nothing data-dependent - the pure leapfrog throughput of the engine."""
    s = settings()
    s.num_tune = 0
    s.maxdepth = depth
    s.mindepth = depth
    ss = s.adapt_options.step_size_settings
    ss.adapt_options.method = _abi.NUTS_STEPSIZE_FIXED
    ss.adapt_options.fixed_step = 0.05
    ss.has_jitter = 0
    math = lib.CudaMath(CHAINS_PER_GPU, DIM, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=model_sigma(), device=device)
    samp = lib.Sampler(math, s, seed=SEED, chain_id_offset=chain_offset)
    samp.set_position(initial_positions(CHAINS_PER_GPU, chain_offset))
    samp.draw_device(2)
    lf0, _ = samp.counters()
    ms = 0.0
    for _ in range(3):
        samp.draw_device(draws)
        ms += samp.last_timing()[0]
    lf1, _ = samp.counters()
    samp.close()
    math.close()
    rate = (lf1 - lf0) / (ms * 1e-3)
    return {"leapfrogs_per_s": rate, "depth": depth, "leapfrogs_per_draw": 2 ** depth - 1, "draws_per_launch": draws,
            "algorithmic_GBps": rate * 48 * DIM / 1e9}


def microbench_plane_leapfrog(lib, _abi, device, nchains=16384, steps=10, warmup=3):
    """The Tier-2 fused leapfrog `nuts_leapfrog` (plane kernel k_leapfrog_tma: rows staged by cp.async.bulk + mbarrier) at an
    HBM-resident size: 16384 chains x 1000 dims = 131 MB per plane, 10 planes per step (read z, v, grad_z, sigma, mean; write z', v',
    x', grad_x', grad_z' = 80 * dim bytes per chain).  Device time of the kernel from CUDA events on the context's stream
    (nuts_ctx_last_kernel_ms); the two points ping-pong, so every step streams 1.3 GB that the previous step did not leave in L2
    in a usable order (planes >> 126 MB L2)."""
    try:
        math = lib.CudaMath(nchains, DIM, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=model_sigma(), device=device)
        math.set_transform(np.exp(np.linspace(-0.5, 0.5, DIM)), 0.1)
        p, st = math.init_state(np.random.default_rng(0).normal(size=(nchains, DIM)))
        math.initialize_trajectory(p, True, SEED, 0, 0)
        q = lib.Point(math)
        out = {}
        for name, flag in (("tma", "1"), ("register_path", "0")):
            os.environ["NUTS_B200_PLANE_TMA"] = flag
            ms = []
            for i in range(warmup + steps):
                q, status, ee = math.leapfrog(p, 0.1, out=q)
                p, q = q, p
                if i >= warmup:
                    ms.append(math.last_kernel_ms())
            mean_ms = sum(ms) / len(ms)
            gbps = 80.0 * DIM * nchains / (mean_ms * 1e-3) / 1e9
            out[name] = {"kernel_ms": mean_ms, "kernel_ms_min": min(ms), "algorithmic_GBps": gbps, "diverged": int((status != 0).sum())}
        os.environ.pop("NUTS_B200_PLANE_TMA", None)
        math.close()
        return {"what": "nuts_leapfrog (Tier 2) on %d chains x %d dims, diagonal Gaussian, CUDA events around the kernel; "
                        "algorithmic bytes = 80 * dim per chain and step" % (nchains, DIM),
                "chains": nchains, "dim": DIM, "steps": steps, "algorithmic_bytes_per_launch": 80 * DIM * nchains, **out}
    except Exception as e:  # a measurement extra must never take the bench line down
        os.environ.pop("NUTS_B200_PLANE_TMA", None)
        return {"error": "%s: %s" % (type(e).__name__, e)}


def cpu_oracle_throughput(max_seconds, nthreads):
    """The oracle (C++ restatement of nuts-rs's CPU path, one chain per thread like the reference's rayon pool,
    src/sampler.rs:1287-1326) on a bounded sample of the same workload: `4 x cores` chains, 400 tuning draws (untimed),
    then post-warmup draws for ~max_seconds."""
    from nuts_rs_b200 import _abi
    from oracle import oracle as O

    nchains = 4 * nthreads
    model = O.Model(_abi.NUTS_LOGP_GAUSS_DIAG, DIM, mu=0.5, sigma=model_sigma())
    samp = O.Sampler(model, settings(), seed=SEED, nchains=nchains, nthreads=nthreads)
    st = samp.set_position(initial_positions(nchains, 0))
    assert (st == 0).all()
    samp.draw(NUM_TUNE, want_draws=False)
    total, elapsed, ndraws = 0, 0.0, 0
    block = 10
    while elapsed < max_seconds:
        t0 = time.perf_counter()
        _, arr = samp.draw(block, want_draws=True)
        elapsed += time.perf_counter() - t0
        total += int(arr["n_steps"].sum())
        ndraws += block
    return {"value": total / elapsed, "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": f"{nchains} chains x {ndraws} post-warmup draws of the same workload ({total} leapfrogs in {elapsed:.1f} s), "
                      f"oracle built -O3 -march=native, one chain per thread"}, total, elapsed


def run_reference(args, rank):
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    per_step_seconds = max(1.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    from nuts_rs_b200 import _abi
    from oracle import oracle as O

    nchains = 4 * nthreads
    model = O.Model(_abi.NUTS_LOGP_GAUSS_DIAG, DIM, mu=0.5, sigma=model_sigma())
    samp = O.Sampler(model, settings(), seed=SEED, nchains=nchains, nthreads=nthreads)
    samp.set_position(initial_positions(nchains, 0))
    samp.draw(NUM_TUNE, want_draws=False)
    # size a step so that it lasts about per_step_seconds
    t0 = time.perf_counter()
    _, arr = samp.draw(2, want_draws=True)
    dt = (time.perf_counter() - t0) / 2
    draws_per_step = max(1, int(per_step_seconds / max(dt, 1e-6)))
    for _ in range(args.warmup):
        samp.draw(draws_per_step, want_draws=True)
    total, t0 = 0, time.perf_counter()
    for _ in range(args.steps):
        _, arr = samp.draw(draws_per_step, want_draws=True)
        total += int(arr["n_steps"].sum())
    elapsed = time.perf_counter() - t0
    value = total / elapsed
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU arm: nuts-rs itself cannot be built here (Rust, no cargo in the image); this is the C++ "
                   "restatement of its CPU path (oracle/), one chain per host thread"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": f"{nchains} chains x {draws_per_step} post-warmup draws per step, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--draws-per-step", type=int, default=DRAWS_PER_STEP)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the sub-records of configs 3-5 and the whole-run entry")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    claim_stdout()

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from nuts_rs_b200 import _abi, lib

    if not lib.device_available():
        raise SystemExit("bench.py: no sm_100 GPU / libnuts_b200.so not usable: " + lib.load().nuts_last_error().decode())
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version banner must not precede the JSON line on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dps = args.draws_per_step
    N = CHAINS_PER_GPU
    chain_offset = rank * N  # global chain id => RNG stream (reference set_stream(chain_id + 1)); weak scaling: 1024 chains per GPU

    math = lib.CudaMath(N, DIM, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=model_sigma(), device=local_rank)
    samp = lib.Sampler(math, settings(), seed=SEED, chain_id_offset=chain_offset)
    x0 = initial_positions(N, chain_offset)
    t0 = time.perf_counter()
    status = samp.set_position(x0)
    assert (status == 0).all()
    samp.draw_device(NUM_TUNE)  # tuning phase, untimed
    math.synchronize()
    tune_wall = time.perf_counter() - t0
    tune_ms, _ = samp.last_timing()
    tune_leapfrogs, _ = samp.counters()

    dev_draws = torch.empty((dps, N, DIM), dtype=torch.float64, device="cuda")
    host_buf = lib.HostBuffer((dps, N, DIM))  # page-locked + device-mapped: nuts_draw lets the kernel write it directly
    host_draws = host_buf.array
    stats_struct, stats_arrays = lib.alloc_stats(dps, N)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm: `value`
    clocks = ClockSampler(local_rank)
    clocks.__enter__()  # samples through the warm-up and both timed regions; only samples inside the timed regions are reported
    for _ in range(args.warmup):
        samp.draw_device(dps, dev_draws.data_ptr())
    math.synchronize()
    lf0, _ = samp.counters()
    barrier()
    kernel_ms, launches = 0.0, 0
    w0 = time.perf_counter()
    for _ in range(args.steps):
        samp.draw_device(dps, dev_draws.data_ptr())
        ms, n = samp.last_timing()  # CUDA events around the kernel on the launching stream (waits for the kernel)
        kernel_ms += ms
        launches += n
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - w0)
    clocks.window(w0, time.perf_counter())
    lf1, _ = samp.counters()
    steps_dev = lf1 - lf0

    # ---------------- end-to-end arm through nuts_draw with host buffers: `e2e`
    import ctypes as C

    def e2e_step():
        lib._check(lib.load().nuts_draw(samp.h, dps, host_draws.ctypes.data_as(_abi.c_double_p), C.byref(stats_struct)))
        return int(stats_arrays["n_steps"].sum())

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    e0 = time.perf_counter()
    steps_e2e = 0
    for _ in range(args.steps):
        steps_e2e += e2e_step()
    barrier()
    e2e_s = time.perf_counter() - e0
    clocks.window(e0, time.perf_counter())
    d2h = host_draws.nbytes + sum(a.nbytes for a in stats_arrays.values())
    e2e_direct = samp.last_draw_direct()

    # ---------------- what the host side can absorb: every rank copies one step's draws device -> page-locked host at the same time
    # (cudaMemcpyAsync, nothing else running).  e2e cannot deliver draws faster than this; at N > 1 the ranks share the host's
    # PCIe root complexes and memory controllers.
    sink_ms = float("nan")
    try:
        dev_buf = torch.empty(host_draws.nbytes, dtype=torch.uint8, device="cuda")
        pin_buf = torch.empty(host_draws.nbytes, dtype=torch.uint8, pin_memory=True)
        pin_buf.copy_(dev_buf, non_blocking=True)
        torch.cuda.synchronize()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(3):
            pin_buf.copy_(dev_buf, non_blocking=True)
        ev[1].record()
        torch.cuda.synchronize()
        sink_ms = ev[0].elapsed_time(ev[1]) / 3.0
        barrier()
        del dev_buf, pin_buf
    except Exception as exc:  # pinned allocation can fail on a small host
        print(f"host sink ceiling not measured: {exc}", file=sys.stderr)

    # ---------------- the only exchange of the path: NCCL all-gather of every step's draws over NVLink, behind the C ABI
    # (nuts_gather_draws_*), on its own stream: the gather of step k runs beside the sampling of step k + 1
    gather = None
    try:
        def exchange(raw):
            box = [raw]
            if world > 1:
                dist.broadcast_object_list(box, src=0)
            return box[0]

        comm = lib.Comm(local_rank, world, rank, exchange)
        bufs = [dev_draws, torch.empty_like(dev_draws)]
        alls = [torch.empty((world,) + tuple(dev_draws.shape), dtype=torch.float64, device="cuda") for _ in range(2)]
        count = dev_draws.numel()

        def pipeline(nsteps):
            for k in range(nsteps):
                samp.draw_device(dps, bufs[k & 1].data_ptr())  # enqueued, not waited for
                if k > 0:
                    comm.gather_end()  # gather k-1 ran beside draw k
                comm.gather_begin(samp, bufs[k & 1].data_ptr(), alls[k & 1].data_ptr(), count)
            last_ms = comm.gather_end()
            math.synchronize()
            return last_ms

        pipeline(max(2, args.warmup))
        lfg0, _ = samp.counters()
        barrier()
        g0 = time.perf_counter()
        gather_last_ms = pipeline(args.steps)
        barrier()
        gather_s = time.perf_counter() - g0
        clocks.window(g0, time.perf_counter())
        lfg1, _ = samp.counters()
        same = bool(torch.equal(alls[(args.steps - 1) & 1][rank], bufs[(args.steps - 1) & 1]))
        gather = {"wall_s": gather_s, "leapfrogs": lfg1 - lfg0, "last_gather_ms": gather_last_ms, "own_shard_intact": same,
                  "bytes_received_per_rank_per_step": int(count * 8 * world)}
        comm.close()
        del bufs, alls
    except Exception as e:  # libnccl missing on a host: the sampling numbers stand, the gather line says why it is absent
        gather = {"error": str(e)[:200]}
    samp.close()
    math.close()

    # ---------------- a whole default run of the headline workload, wall clock through the public call (per rank: 1024 chains,
    # 400 tuning + 1000 sampling draws, every draw and statistic delivered to host memory in batches of `dps` draws)
    whole = None
    if not args.no_other_configs:
        math = lib.CudaMath(N, DIM, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=model_sigma(), device=local_rank)
        samp = lib.Sampler(math, settings(), seed=SEED, chain_id_offset=chain_offset)
        barrier()
        r0 = time.perf_counter()
        samp.set_position(x0)
        total_lf, tune_lf, done = 0, 0, 0
        while done < NUM_TUNE + 1000:
            lib._check(lib.load().nuts_draw(samp.h, dps, host_draws.ctypes.data_as(_abi.c_double_p), C.byref(stats_struct)))
            n = int(stats_arrays["n_steps"].sum())
            total_lf += n
            if done < NUM_TUNE:
                tune_lf += n
                r_tune = time.perf_counter() - r0
            done += dps
        barrier()
        whole_s = time.perf_counter() - r0
        clocks.window(r0, time.perf_counter())
        whole = {"wall_s": whole_s, "tune_wall_s": r_tune, "leapfrogs": total_lf, "tune_leapfrogs": tune_lf}
        samp.close()
        math.close()
    host_buf.close()
    del dev_draws

    # ---------------- the other BASELINE configs on this rank (device-timed sub-records); config 5 = 65536 chains over the ranks
    other = {}
    if not args.no_other_configs:
        from nuts_rs_b200 import sharding

        for name, k_steps, k_dps in (("c3", 5, 50), ("c4", 3, 10), ("c5", 3, 20)):
            cfg = CONFIGS[name]
            if name == "c5":
                off_c, n_c = sharding.shard_range(cfg["chains"], world, rank)
            else:
                n_c, off_c = cfg["chains"], rank * cfg["chains"]
            barrier()
            c0 = time.perf_counter()
            other[name] = run_config(lib, name, n_c, off_c, local_rank, k_steps, 2, k_dps)
            other[name].update(chains_this_rank=n_c, steps=k_steps, draws_per_step=k_dps)
            barrier()
            clocks.window(c0, time.perf_counter())
            other[name]["clocks"] = clocks.summary_between(c0, time.perf_counter())
    clocks.__exit__(None, None, None)

    micro = microbench_fixed_depth(lib, _abi, local_rank, chain_offset) if rank == 0 else None
    plane = microbench_plane_leapfrog(lib, _abi, local_rank) if rank == 0 else None

    # ---------------- reduce over ranks: max time, summed work
    names = sorted(other)
    tl = [kernel_ms, wall_ms, e2e_s * 1e3, tune_ms] + [other[n]["kernel_ms"] for n in names] + [other[n]["tune_ms"] for n in names]
    wl = [float(steps_dev), float(steps_e2e), float(tune_leapfrogs)] + [float(other[n]["leapfrogs"]) for n in names] \
        + [float(other[n]["tune_leapfrogs"]) for n in names]
    if whole:
        tl += [whole["wall_s"] * 1e3, whole["tune_wall_s"] * 1e3]
        wl += [float(whole["leapfrogs"]), float(whole["tune_leapfrogs"])]
    has_gather = gather is not None and "error" not in gather
    tl += [gather["wall_s"] * 1e3 if has_gather else 0.0, sink_ms if sink_ms == sink_ms else 0.0]
    wl += [float(gather["leapfrogs"]) if has_gather else 0.0]
    t = torch.tensor(tl, dtype=torch.float64, device="cuda")
    w = torch.tensor(wl, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
    t, w = t.tolist(), w.tolist()
    kernel_ms_max, wall_ms_max, e2e_ms_max, tune_ms_max = t[:4]
    steps_dev_all, steps_e2e_all, tune_lf_all = w[:3]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        value = steps_dev_all / (kernel_ms_max * 1e-3)
        alg_bytes_per_step = 48 * DIM  # read z, v, grad_z + write z', v', grad_z' per leapfrog per chain
        per_gpu_steps_per_launch = steps_dev / max(1, launches)
        avg_launch_ms = kernel_ms / max(1, launches)
        per_gpu_rate = per_gpu_steps_per_launch / (avg_launch_ms * 1e-3)
        achieved = per_gpu_rate * alg_bytes_per_step / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the same kernel on the same workload (10 draws x 1024 chains),
        # from the committed `ncu --set full` capture (profiles/); scaled to this run's draws per launch
        traffic, traffic_src, prof = None, None, {}
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "draw_kernel_ncu_latest.json")))
            traffic = float(prof["dram_traffic_bytes_per_launch"]) * dps / float(prof.get("draws_per_launch", 10))
            traffic_src = f"profiles/draw_kernel_ncu_latest.json ({prof.get('Kernel Name', '?')})"
        except Exception:
            pass
        # the ceilings that actually bind this kernel (the 48*d state never reaches DRAM: it lives in registers)
        ckpt_bytes = 16 * DIM  # one (z, v) checkpoint per leapfrog is what HAS to be written
        fp64_pct = prof.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
        ceilings = {
            "checkpoint_hbm": {"bytes_per_leapfrog": ckpt_bytes, "ceiling_leapfrogs_per_s": peak * 1e9 / ckpt_bytes,
                               "frac": per_gpu_rate * ckpt_bytes / 1e9 / peak,
                               "note": "the (z, v) checkpoint of every leaf (16*d bytes) is the only traffic a leapfrog must send to HBM"},
            "fp64_pipe": {"pct_of_peak_in_ncu_capture": None if fp64_pct is None else float(fp64_pct),
                          "source": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active of the committed ncu capture "
                                    "(profiles/draw_kernel_ncu_latest.json); scales with value / the capture's own rate",
                          "capture_leapfrogs_per_s": prof.get("leapfrogs_per_s")},
            "algorithmic_48d": {"bytes_per_leapfrog": alg_bytes_per_step, "frac": achieved / peak,
                                "note": "SURVEY 8(d) figure; saturates (> 1 possible) because these bytes stay on chip"},
        }
        cpu = None
        if not args.no_cpu_baseline:
            cpu, _, _ = cpu_oracle_throughput(args.cpu_seconds, os.cpu_count() or 1)
        other_out = {}
        for i, n in enumerate(names):
            cfg = CONFIGS[n]
            k_ms, k_tune_ms = t[4 + i], t[4 + len(names) + i]
            k_lf, k_tune_lf = w[3 + i], w[3 + len(names) + i]
            rate = k_lf / (k_ms * 1e-3)
            o = other[n]
            rec = {"workload": f"configs[{cfg['index']}]: {cfg['what']}", "value": rate, "unit": UNIT,
                   "scaling": "strong (65536 chains split over the ranks)" if n == "c5" else "weak (the config per GPU)",
                   "chains_rank0": o["chains_this_rank"], "chains_all_ranks": cfg["chains"] if n == "c5" else cfg["chains"] * world,
                   "dim": cfg["dim"], "num_tune": cfg["num_tune"], "steps": o["steps"], "draws_per_step": o["draws_per_step"],
                   "ms_per_step": k_ms / o["steps"], "tuning_phase_leapfrogs_per_s": k_tune_lf / (k_tune_ms * 1e-3),
                   "tuning_phase_ms": k_tune_ms, "bad_init_rank0": o["bad_init"], "clocks_rank0": o["clocks"],
                   "algorithmic_GBps_48d": rate * 48 * cfg["dim"] / 1e9,
                   "timing": "CUDA events around the draw kernel, summed over the steps, max over ranks"}
            if cfg["hbm_bound"]:
                rec["roofline"] = {"bound": "hbm", "achieved": rate / world * 48 * cfg["dim"] / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": rate / world * 48 * cfg["dim"] / 1e9 / peak,
                                   "checkpoint_hbm_frac": rate / world * 16 * cfg["dim"] / 1e9 / peak}
            else:
                rec["roofline"] = "not meaningful: the chain state fits on chip (SURVEY 8d); latency / occupancy bound"
            other_out[n] = rec
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": kernel_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": N, "dim": DIM, "draws_per_step": dps,
                       "leapfrogs_per_step_all_gpus": steps_dev_all / args.steps,
                       "cache": "inputs larger than L2: the checkpoint pools of the resident chains (one 16 kB slot per leaf) + 8 estimator / "
                                "13 state planes of 8 MB cycle through far more than the 126 MB L2; no flush between steps",
                       "timing": "CUDA events around the draw kernel on the launching stream, summed over K steps, max over ranks; the "
                                 "kernel writes the draws and all 15 statistics of Chain::draw to HBM buffers",
                       "wall_ms_per_step_incl_launch": wall_ms_max / args.steps,
                       "tuning_phase": {"draws": NUM_TUNE, "kernel_ms": tune_ms_max, "wall_s": tune_wall, "leapfrogs": tune_lf_all,
                                        "leapfrogs_per_s": tune_lf_all / max(tune_ms_max * 1e-3, 1e-9),
                                        "ratio_to_sampling_rate": tune_lf_all / max(tune_ms_max * 1e-3, 1e-9) / value}},
            "e2e": {"value": steps_e2e_all / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(d2h),
                    "d2h_GBps_per_rank": d2h * args.steps / (e2e_ms_max * 1e-3) / 1e9,
                    "host_sink_ceiling": {
                        "GBps_per_rank": (host_draws.nbytes / (t[-1] * 1e-3) / 1e9) if t[-1] > 0 else None,
                        "leapfrogs_per_s_all_gpus": (steps_e2e_all / args.steps / (t[-1] * 1e-3)) if t[-1] > 0 else None,
                        "what": "all ranks copy one step's draws device -> page-locked host at the same time with cudaMemcpyAsync and "
                                "nothing else running (max over ranks): the rate at which this box's host side absorbs draws; e2e "
                                "delivers the same bytes while sampling"},
                    "draws_path": "kernel writes the page-locked host buffer directly (posted PCIe writes overlapping the sampling)"
                                  if e2e_direct else "device staging buffer + cudaMemcpyAsync D2H after the kernel",
                    "note": "nuts_draw with page-locked host buffers: every draw [draws x chains x dim] f64 and all 15 statistics reach host "
                            "memory inside the timed region; chain state stays resident between calls like the reference's Chain (initial "
                            f"positions: one {x0.nbytes}-byte H2D in nuts_set_position, outside the steps)"},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src, "peak_source": peak_kind, "binding_ceilings": ceilings,
                         "note": "achieved = leapfrogs per launch x 48*dim algorithmic bytes / launch duration; the engine keeps z, v, grad in "
                                 "registers across leapfrogs, so algorithmic bytes are NOT DRAM bytes (frac > 1 is possible); see DESIGN.md"},
            "cpu_baseline": cpu,
        }
        if has_gather:
            line["config"]["draw_gather"] = {
                "what": "every step's draws all-gathered to every rank with NCCL over NVLink through the C ABI (nuts_gather_draws_begin / "
                        "_end) on the communicator's own stream, overlapping the sampling of the next step; wall clock over the K steps "
                        "incl. the last gather, max over ranks",
                "leapfrogs_per_s": w[-1] / (t[-2] * 1e-3), "ms_per_step": t[-2] / args.steps,
                "ratio_to_value": w[-1] / (t[-2] * 1e-3) / value, "last_gather_device_ms_rank0": gather["last_gather_ms"],
                "bytes_received_per_rank_per_step": gather["bytes_received_per_rank_per_step"],
                "gather_GBps_per_rank": gather["bytes_received_per_rank_per_step"] / max(gather["last_gather_ms"], 1e-9) / 1e6,
                "own_shard_intact_rank0": gather["own_shard_intact"]}
        elif gather is not None:
            line["config"]["draw_gather"] = gather
        line["config"]["microbench_fixed_step_no_turn_checks_rank0"] = micro
        if plane is not None:
            peak = line["roofline"]["peak"]
            for k in ("tma", "register_path"):
                if isinstance(plane.get(k), dict):
                    plane[k]["frac_of_hbm_peak"] = plane[k]["algorithmic_GBps"] / peak
            line["config"]["plane_leapfrog_rank0"] = plane
        line["config"]["other_configs"] = other_out
        if whole:
            k = 4 + 2 * len(names)
            line["config"]["whole_run"] = {
                "what": "nuts_set_position + 400 tuning + 1000 sampling draws of the headline workload per rank through nuts_draw, every "
                        "draw and statistic delivered to page-locked host memory in batches; wall clock, max over ranks",
                "wall_s": t[k] * 1e-3, "tuning_wall_s": t[k + 1] * 1e-3, "leapfrogs_all_gpus": w[3 + 2 * len(names)],
                "leapfrogs_per_s": w[3 + 2 * len(names)] / (t[k] * 1e-3),
                "tuning_share_of_wall": t[k + 1] / t[k]}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
