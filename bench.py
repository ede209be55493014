#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched PupperV3Env step (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path on 1 B200: BASELINE configs[1]
    torchrun --nproc-per-node N bench.py --gpus N ...        # N ranks, envs sharded (weak scaling): BASELINE configs[3]
    python bench.py --impl reference ...                      # the reference algorithm on the host cores (same config)

Workloads (`config.workload`), env kwargs of the reference's only complete set (reference test/test_environment.py:64-113),
Brax EpisodeWrapper/AutoResetWrapper semantics fused (episode_length 1000), synthetic actions 0.5*U(-1,1), a "step" = one
env step of every env (= 5 physics substeps, reference environment.py:179):

  N = 1   configs[1]: flat-ground velocity tracking, 4096 envs, full domain randomisation.  The same JSON line carries, under
          "configs", the single-GPU numbers of configs[3]'s per-GPU batch (65,536 envs: the weak-scaling base of the N > 1
          lines), configs[2] (obstacles.py box terrain with randomised pushes, 16,384 envs) and configs[4] (rollout
          collection with the policy MLP in the loop, 8192 envs).
  N > 1   configs[3]: 65,536 envs per GPU sharded over the N GPUs, full domain randomisation, with the path's only collective
          (NCCL SUM all-reduce of the 24-float episode accumulator, `parallel.EpisodeMetricsReducer`) issued every
          min(100, steps) steps INSIDE the timed region and reported as "collective_ms".

`--envs` overrides the batch for sweeps.  The reference itself (JAX/Brax/MJX) cannot be installed in this image (no wheels,
no network), so the `--impl reference` arm and `cpu_baseline` time the in-repo float32 C restatement of the same algorithm
(oracle/, kind "port") with OpenMP over envs on all host cores; that arm neither builds nor loads the CUDA library.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "env-steps/sec (device-timed)"
UNIT = "env-steps/s"
H = 2  # observation_history of the workload
COUNTERS = os.path.join(ROOT, "profiles", "r2_counters.json")


def b_alg(h: int) -> int:
    """Algorithmic bytes per env-step: action + state r/w + outputs + DR + obs history,
    plus 192 B for the fused Episode/AutoReset accounting."""
    return 4 * (355 + 36 * (2 * h - 1)) + 192


def default_envs(gpus: int) -> int:
    return 4096 if gpus <= 1 else 65536


def workload(envs: int, gpus: int, obstacles: bool = False, dr: bool = True) -> str:
    """config.workload -- the same string for both arms of a comparison (the driver checks that)."""
    if obstacles:
        head = "configs[2]: obstacles.py box terrain (10 boxes, reference test/test_environment.py:28-41) with randomised pushes"
    elif gpus > 1:
        head = "configs[3]: flat ground, envs sharded over the GPUs, NCCL all-reduce of the episode metrics"
    else:
        head = "configs[1]: flat-ground velocity tracking"
    return (f"{head}, {envs} envs/GPU x {gpus} GPU(s), {'full' if dr else 'no'} domain randomisation, H={H}, "
            f"fused episode/auto-reset, 5 substeps/step")


def host_cores() -> int:
    return len(os.sched_getaffinity(0))


def time_oracle(env, n, steps, warmup, use_dr, seed=0):
    """Times the float32 oracle (OpenMP over envs, all host cores). Returns (env-steps/s, seconds/step)."""
    import common
    from oracle import oracle
    from pupperv3_mjx_b200 import domain_randomization as dr, prng
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f32", n_threads=host_cores())  # torchrun pins OMP_NUM_THREADS=1: ask explicitly
    if use_dr:
        sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
        O.set_dr(common.dr_struct(sys_v))
    O.reset(common.env_keys(n, seed))
    acts = [common.actions(n, t) for t in range(4)]
    for t in range(warmup):
        O.step(acts[t % 4], episode=True)
    t0 = time.perf_counter()
    for t in range(steps):
        O.step(acts[t % 4], episode=True)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return n / dt, dt


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, False
        self.sm, self.reasons, self.sm_max = [], set(), None
        self.ready = threading.Event()  # set once NVML is initialised, so short timed regions still get samples

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
            self.ready.set()
            while not self.stop_flag:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.002)
        except Exception as e:  # NVML unavailable: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")
            self.ready.set()

    def result(self):
        import numpy as np
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def run_reference(args):
    """The reference algorithm on the host cores (float32 C port, all host threads), rank 0 only.  Nothing of the CUDA
    product is built, loaded or called here."""
    import __graft_entry__ as g
    g.build_oracle()
    import common
    n = min(args.envs, 4096)  # each CPU step is a bounded sample of the workload's env batch
    env = common.make_env()
    env.set_episode_params(1000, 1)
    cores = host_cores()
    value, dt = time_oracle(env, n, args.steps, args.warmup, use_dr=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(args.envs, args.gpus), "envs_per_gpu": args.envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step runs a {n}-env sample of the workload's batch; {args.steps} steps, float32 C restatement "
                                   f"(oracle/), OpenMP over envs on {cores} threads; the JAX/Brax/MJX reference is not installable here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def load_counters():
    """ncu-derived per-launch counters of the step kernel (profiles/r2_counters.json, written by tools/make_counters.py from
    a `--set full` capture).  They are only quoted when they were captured from the library build that is loaded now."""
    try:
        c = json.load(open(COUNTERS))
        import __graft_entry__ as g
        c["matches_loaded_library"] = (c.get("kernel_digest") == g.kernel_digest())
        return c
    except Exception:
        return {"matches_loaded_library": False}


def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import common
    from pupperv3_mjx_b200 import abi, domain_randomization as dr, parallel, runtime

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_runtime(env, n, use_dr=True):
        """Runtime of this rank's shard: env / DR keys are `split(PRNGKey(s), world*n)` sliced per rank
        (parallel.shard_keys), so the batch does not depend on the rank count."""
        rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, device=local, episode=True)
        if use_dr:
            sys_v, _ = dr.domain_randomize(env.sys, parallel.shard_keys(2, world * n, rank, world))
            rt.set_dr(sys_v)
        keys = parallel.shard_keys(0, world * n, rank, world)
        rt.reset(torch.from_numpy(np.ascontiguousarray(keys).view(np.int32)).to(dev))
        acts = [(torch.rand((n, 12), generator=gen, device=dev) - 0.5) for _ in range(8)]  # 0.5*U(-1,1)
        return rt, acts

    def timed_steps(rt, acts, steps, warmup, settle, do_flush, reducer=None, period=0):
        """W warm-up steps, then `steps` timed steps, each inside its own CUDA event pair on the launch stream (the L2 flush
        sits between the pairs).  Returns per-step ms, per-collective ms, launches, wall seconds."""
        for t in range(settle):  # untimed pre-roll: robots dropped by reset land and episodes de-synchronise
            rt.step(acts[t % len(acts)])
        for t in range(warmup):
            rt.step(acts[t % len(acts)])
        if reducer is not None:  # NCCL communicator set-up belongs to the warm-up, not to the first timed all-reduce
            reducer.reduce()
        barrier()
        launches0 = rt.launches
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        cev = []
        wall0 = time.perf_counter()
        for t in range(steps):
            if do_flush:
                flush.zero_()
            ev[t][0].record()
            rt.step(acts[t % len(acts)])
            if reducer is not None and (t + 1) % period == 0:
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                reducer.reduce()  # episode metrics: the only collective of the path, inside the step's event pair
                c1.record()
                cev.append((c0, c1))
            ev[t][1].record()
        barrier()
        wall = time.perf_counter() - wall0
        per_step = np.array([a.elapsed_time(b) for a, b in ev])
        coll = np.array([a.elapsed_time(b) for a, b in cev]) if cev else np.zeros(0)
        return per_step, coll, rt.launches - launches0, wall

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = args.envs
    W = max(args.warmup, 3)
    env = common.make_env(obstacles_on=args.obstacles)
    env.set_episode_params(1000, 1)
    rt, acts = make_runtime(env, n, use_dr=not args.no_dr)
    reducer = parallel.EpisodeMetricsReducer(rt.episode_field("totals")) if world > 1 else None
    period = max(1, min(100, args.steps))
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(timeout=10)
    per_step, coll, launches, wall = timed_steps(rt, acts, args.steps, W, args.settle, not args.no_flush, reducer, period)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms = max_over_ranks(float(np.mean(per_step)))
    coll_ms = max_over_ranks(float(np.mean(coll))) if len(coll) else None
    # the kernel alone (steps without a collective): what the roofline is quoted on
    has_coll = np.zeros(len(per_step), bool)
    if reducer is not None:
        has_coll[period - 1::period] = True
    kern = per_step[~has_coll] if (~has_coll).any() else per_step - (float(np.mean(coll)) if len(coll) else 0.0)
    kernel_ms = max_over_ranks(float(np.mean(kern)))
    value = world * n / (ms * 1e-3)

    # ---- end to end through the public API with host buffers -----------------------------------------------
    h_act = [a.cpu().pin_memory() for a in acts]
    h_out = torch.empty(n * (H * abi.OBS_DIM + 2), dtype=torch.float32).pin_memory()  # obs | reward | done, as the runtime packs them
    e2e_steps = max(10, min(args.steps, 200))
    for t in range(3):  # stream / staging-buffer set-up of the host path outside the timed region
        rt.step_host(h_act[t % len(acts)], h_out).synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(e2e_steps):
        # pinned host action in, obs + reward + done out; the host waits for the results (the policy needs obs before the
        # next action).  Batches >= 32768 envs are pipelined in env ranges so the PCIe copies overlap the kernels.
        rt.step_host(h_act[t % len(acts)], h_out).synchronize()
    e1.record()
    barrier()
    e2e_value = world * n / (max_over_ranks(e0.elapsed_time(e1) / e2e_steps) * 1e-3)
    report = parallel.episode_report(reducer.global_totals) if reducer is not None else parallel.episode_report(rt.episode_field("totals"))

    # ---- the other BASELINE configs, short runs of the same timing loop (no flag needed) ----------------------------
    configs = {}
    if not args.only_main and not args.obstacles:
        xs = 50
        if world == 1 and n != 65536:
            r2, a2 = make_runtime(env, 65536)
            ps, _, _, _ = timed_steps(r2, a2, xs, 3, args.settle, True)
            configs["configs[3] per-GPU batch on 1 GPU (weak-scaling base)"] = {
                "value": 65536 / (float(np.mean(ps)) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ps)), "envs": 65536, "steps": xs,
                "note": "flat ground, full DR, same timing loop (event pair per step, L2 flushed between steps)"}
            del r2, a2
        if world == 1:
            env_o = common.make_env(obstacles_on=True)
            env_o.set_episode_params(1000, 1)
            r2, a2 = make_runtime(env_o, 16384)
            ps, _, _, _ = timed_steps(r2, a2, xs, 3, args.settle, True)
            configs["configs[2] obstacle terrain"] = {
                "value": 16384 / (float(np.mean(ps)) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ps)), "envs": 16384, "steps": xs,
                "note": "10-box obstacles.py terrain, kicks on, full DR, same timing loop"}
            del r2, a2
        if world == 1:
            # one run whose state genuinely streams from HBM (H = 15: 5,596 + 192 B per env-step, 380 MB per step)
            env_h = common.make_env(observation_history=15)
            env_h.set_episode_params(1000, 1)
            r2, a2 = make_runtime(env_h, 65536)
            ps, _, _, _ = timed_steps(r2, a2, 30, 3, 20, True)
            bh = b_alg(15) * 65536
            configs["65,536 envs with observation_history = 15 (HBM-streaming variant)"] = {
                "value": 65536 / (float(np.mean(ps)) * 1e-3), "unit": UNIT, "ms_per_step": float(np.mean(ps)), "envs": 65536, "steps": 30,
                "algorithmic_GB_per_s": bh / (float(np.mean(ps)) * 1e-3) / 1e9,
                "note": "flat ground, full DR; the obs history alone is 283 MB per step, beyond the 126 MB L2"}
            del r2, a2
        # configs[4]: rollout collection with the policy MLP in the loop, 8192 envs per GPU (x N GPUs)
        import functools
        from pupperv3_mjx_b200 import rollout, wrappers
        en, T = 8192, 20
        env_r = common.make_env(device=local)
        rand = functools.partial(dr.domain_randomize, rng=parallel.shard_keys(2, world * en, rank, world))
        tenv = wrappers.wrap(env_r, episode_length=1000, randomization_fn=rand)
        st = tenv.reset(torch.from_numpy(np.ascontiguousarray(parallel.shard_keys(0, world * en, rank, world)).view(np.int32)).to(dev))
        pol_r = rollout.PolicyMLP.random(env_r.observation_size, precision=1, device=dev)
        roll = {}
        for mode in ("graph", "one_launch"):
            col = rollout.RolloutCollector(tenv, pol_r, st, T, use_cuda_graph=True, fused=mode == "one_launch")
            for _ in range(5):
                col.collect()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(10):
                col.collect()
            s1.record()
            barrier()
            roll[mode] = world * en * T * 10 / (max_over_ranks(s0.elapsed_time(s1)) * 1e-3)
        configs["configs[4] rollout collection"] = {
            "value": roll["graph"], "unit": UNIT, "envs_per_gpu": en, "unroll": T,
            "one_launch_variant": roll["one_launch"],
            "note": "policy MLP 72-256-128-128-128-12 (stand-in for the JAX policy: TF32 on the tcgen05 kernel, XLA's default float32 "
                    "matmul precision) + fused env step; per unroll ONE CUDA graph of 20 x (policy kernel that also files obs[t], env "
                    "step that writes reward[t] / done[t] into the trajectory), max over ranks; one_launch_variant = pupper_rollout "
                    "(the whole unroll as one kernel: policy phase + env step per CTA, steps chained on the device), kept opt-in "
                    "because it is slower (instruction-fetch bound, DESIGN.md section 4)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    bytes_per_launch = b_alg(H) * n
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    counters = load_counters()
    cnt = counters.get(f"envs_{n}", {}) if counters.get("matches_loaded_library") else {}
    flop_per_step = cnt.get("flop_per_env_step") or counters.get("flop_per_env_step_any_build")
    ffma_peak = runtime.measure_ffma_tflops(local)  # measured FP32 denominator (MEASURED_PEAKS.json has none)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload(n, world, args.obstacles, not args.no_dr), "envs_per_gpu": n, "settle_steps": args.settle,
                   "l2": "NOT flushed (diagnostic run)" if args.no_flush else "flushed between timed steps (256 MB memset outside the event pairs)",
                   "timing": "mean of per-step CUDA event pairs on the launch stream, max over ranks"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": cnt.get("dram_bytes_per_launch"), "peak_source": peak_src, "algorithmic_bytes_per_env_step": b_alg(H),
                     "kernel_ms": kernel_ms,
                     "counters": "profiles/r2_counters.json" + ("" if counters.get("matches_loaded_library") else " (captured from another build: not quoted)"),
                     "note": "the step is instruction-issue bound, not HBM bound (DESIGN.md); see fp32"},
        "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 12 * 4, "d2h_bytes_per_step": n * (H * abi.OBS_DIM + 2) * 4,
                "steps": e2e_steps, "note": "EnvRuntime.step_host: HOST buffers in and out, the host waits for every step's results; one kernel launch per "
                          "step and nothing else -- the kernel reads the pinned host action over PCIe (h2d bytes) and stores obs / reward / done "
                          "into the pinned host result buffer itself (d2h bytes; PupperStepOut.obs_copy), so the transfers are inside the timed "
                          "region and overlap the CTAs still computing; the device-resident path (`value`) is the product number"},
        "gpu_launches": launches,
        "ms_per_step_quantiles": {"p50": float(np.quantile(per_step, 0.5)), "p90": float(np.quantile(per_step, 0.9)),
                                  "max": float(per_step.max()), "note": "rank 0; steps in which an env takes a rare solver path run longer"},
        "wall_s_timed_region": wall,
        "episode_report": {k: report[k] for k in ("episodes", "sum_reward", "length", "terminations")},
    }
    if flop_per_step:
        line["fp32"] = {"flop_per_env_step": flop_per_step, "achieved_tflops": flop_per_step * n / (kernel_ms * 1e-3) / 1e12,
                        "peak_tflops": ffma_peak, "peak_source": "measured in this run: FFMA probe kernel, 8 chains/thread, best of 5 (nominal 74.4)",
                        "frac": flop_per_step * n / (kernel_ms * 1e-3) / 1e12 / ffma_peak,
                        "flop_source": "ncu source page of " + ("this build" if cnt else "an earlier build of the kernel")}
    if world > 1:
        line["collective_ms"] = coll_ms
        line["collectives_in_timed_region"] = int(has_coll.sum())
        line["collective"] = f"NCCL SUM all-reduce of the {abi.N_TOTALS}-float episode accumulator every {period} steps, inside the step's event pair"
    if configs:
        line["configs"] = configs
    if world == 1 and not args.skip_cpu:
        cores = host_cores()
        cpu_n, cpu_steps = min(n, 4096), 15
        v, dt = time_oracle(env, cpu_n, cpu_steps, 1, use_dr=not args.no_dr)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cpu_n} envs x {cpu_steps} steps of the same workload, float32 C restatement "
                                          f"(oracle/), OpenMP over envs on {cores} threads"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: 4096 = configs[1] at N=1, 65536 = configs[3] at N>1)")
    ap.add_argument("--settle", type=int, default=100, help="untimed pre-roll steps after reset (steady-state contacts)")
    ap.add_argument("--no-dr", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: keep L2 warm between steps (not a bench number)")
    ap.add_argument("--obstacles", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--only-main", action="store_true", help="skip the short runs of the other BASELINE configs")
    args = ap.parse_args()
    if args.envs is None:
        args.envs = default_envs(args.gpus)
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return  # rank 0 alone runs the CPU arm; the other ranks leave before any heavy import
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
