#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched PupperV3Env step (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path on 1 B200
    torchrun --nproc-per-node N bench.py --gpus N ...        # N ranks, envs sharded, weak scaling
    python bench.py --impl reference ...                      # the reference algorithm on the host cores

Workload (config.workload): BASELINE.json configs[1] -- flat-ground velocity tracking, 4096 envs per GPU
with full domain randomisation, env kwargs of the reference's only complete set
(reference test/test_environment.py:64-113), Brax EpisodeWrapper/AutoResetWrapper semantics fused
(episode_length 1000), synthetic actions 0.5*U(-1,1).  A "step" is one env step of every env
(= 5 physics substeps, reference environment.py:179).  `--envs` overrides the batch for sweeps; the extra
object reports the 16384 (configs[2] batch) and 65536 (configs[3] batch) flat-ground throughput too.

The reference itself (JAX/Brax/MJX) cannot be installed in this image (no wheels, no network), so the
`--impl reference` arm and `cpu_baseline` time the in-repo float32 C restatement of the same algorithm
(oracle/, kind "port") with OpenMP over envs on all host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "env-steps/sec (device-timed)"
UNIT = "env-steps/s"
H = 2  # observation_history of the workload


def b_alg(h: int) -> int:
    """Algorithmic bytes per env-step: action + state r/w + outputs + DR + obs history,
    plus 192 B for the fused Episode/AutoReset accounting."""
    return 4 * (355 + 36 * (2 * h - 1)) + 192


def make_env(obstacles=False):
    import common
    return common.make_env(obstacles_on=obstacles)


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
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(args.envs, 4096)  # each step is a bounded sample of the workload's env batch (the whole batch at the default size)
    env = make_env()
    env.set_episode_params(1000, 1)
    cores = host_cores()
    value, dt = time_oracle(env, n, args.steps, args.warmup, use_dr=True)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: flat ground, {args.envs} envs/GPU, full domain randomisation, H={H}, fused episode/auto-reset, "
                               f"5 substeps/step; each CPU step runs a {n}-env sample of that batch",
                   "envs_per_gpu": args.envs, "sample_envs": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} envs x {args.steps} steps of the same workload, float32 C restatement (oracle/), "
                                   f"OpenMP over envs on {cores} threads; the JAX/Brax/MJX reference is not installable here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_cuda(args):
    import torch
    import torch.distributed as dist
    import common
    from pupperv3_mjx_b200 import abi, domain_randomization as dr, prng, runtime

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.envs
    env = make_env(obstacles=args.obstacles)
    env.set_episode_params(1000, 1)
    rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, device=local, episode=True)
    # env keys / DR keys: split(PRNGKey(s), world*n) sliced per rank, so results do not depend on the rank count
    keys = prng.split(prng.PRNGKey(0), world * n)[rank * n:(rank + 1) * n]
    if not args.no_dr:
        sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), world * n)[rank * n:(rank + 1) * n])
        rt.set_dr(sys_v)
    rt.reset(torch.from_numpy(np.ascontiguousarray(keys).view(np.int32)).to(dev))
    n_act = 8
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    acts = [(torch.rand((n, 12), generator=g, device=dev) - 0.5) for _ in range(n_act)]  # 0.5*U(-1,1)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    totals = rt.episode_field("totals")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for t in range(args.settle):  # untimed pre-roll: robots dropped by reset land and episodes de-synchronise
        rt.step(acts[t % n_act])
    for t in range(max(args.warmup, 3)):
        rt.step(acts[t % n_act])
    if world > 1:  # NCCL communicator set-up belongs to the warm-up, not to the first timed all-reduce
        dist.all_reduce(totals)
        totals.div_(world)
    barrier()
    launches0 = rt.launches
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(timeout=10)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for t in range(args.steps):
        if not args.no_flush:
            flush.zero_()  # L2 flush between timed iterations (outside the per-step event pair)
        ev[t][0].record()
        rt.step(acts[t % n_act])
        if world > 1 and (t + 1) % 100 == 0:
            dist.all_reduce(totals)  # episode metrics: the only collective of the path
            totals.div_(world)  # keep the running sums bounded (replicated accumulators)
        ev[t][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    sampler.stop_flag = True
    sampler.join(timeout=2)
    per_step = np.array([a.elapsed_time(b) for a, b in ev])
    kernel_ms = float(np.mean(per_step))
    tt = torch.tensor([kernel_ms], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    launches = rt.launches - launches0
    value = world * n / (ms * 1e-3)

    # ---- end to end through the public API with host buffers -----------------------------------------------
    state = env_state_e2e = None
    h_act = [a.cpu().pin_memory() for a in acts]
    h_out = torch.empty(n * (H * abi.OBS_DIM + 2), dtype=torch.float32).pin_memory()  # obs | reward | done, as the runtime packs them
    e2e_steps = max(10, min(args.steps, 200))
    for t in range(3):  # stream / staging-buffer set-up of the host path outside the timed region
        rt.step_host(h_act[t % n_act], h_out).synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(e2e_steps):
        # pinned host action in, obs + reward + done out; the host waits for the results (the policy needs obs before the
        # next action).  Batches >= 32768 envs are pipelined in env ranges so the PCIe copies overlap the kernels.
        rt.step_host(h_act[t % n_act], h_out).synchronize()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    te = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n / (float(te.item()) * 1e-3)

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
    achieved = bytes_per_launch / (ms * 1e-3) / 1e9
    flop_per_step = 1.59e5  # executed FP32 flop per env-step of this kernel (ncu source page, profiles/r1_summary.md)
    traffic_4096 = 5.18e6  # dram__bytes_read+write per launch at 4096 envs from the ncu --set full capture (profiles/r1_summary.md)
    ffma_peak = runtime.measure_ffma_tflops(local)  # measured FP32 denominator (MEASURED_PEAKS.json has none)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"configs[2]: obstacles.py box terrain (10 boxes, reference test/test_environment.py:28-41) with randomised pushes, "
                                if args.obstacles else "configs[1]: flat ground, ") +
                               f"{n} envs/GPU, {'no' if args.no_dr else 'full'} domain randomisation, H={H}, "
                               f"fused episode/auto-reset, 5 substeps/step",
                   "envs_per_gpu": n, "settle_steps": args.settle, "l2": "NOT flushed (diagnostic run)" if args.no_flush else "flushed between timed steps (256 MB memset outside the event pairs)",
                   "timing": "mean of per-step CUDA event pairs on the launch stream, max over ranks"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic_4096 if n == 4096 else (9.2e7 if n == 65536 else None),  # 65536: 79.5 MB read + 12.5 MB written (same capture set) "peak_source": peak_src, "algorithmic_bytes_per_env_step": b_alg(H),
                     "note": "the step is FP32-pipe/latency bound, not HBM bound (DESIGN.md); see fp32"},
        "fp32": {"flop_per_env_step": flop_per_step, "achieved_tflops": flop_per_step * value / world / 1e12,
                 "peak_tflops": ffma_peak, "peak_source": "measured in this run: FFMA probe kernel, 8 chains/thread, best of 5 (nominal 74.4)",
                 "frac": flop_per_step * value / world / 1e12 / ffma_peak},
        "clocks": sampler.result(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 12 * 4, "d2h_bytes_per_step": n * (H * abi.OBS_DIM + 2) * 4,
                "steps": e2e_steps, "note": "EnvRuntime.step_host: pinned host action in, obs+reward+done out, host waits every step; "
                          f"{max(1, min(8, n // 16384))} pipelined env range(s)"},
        "gpu_launches": launches,
        "ms_per_step_quantiles": {"p50": float(np.quantile(per_step, 0.5)), "p90": float(np.quantile(per_step, 0.9)),
                                  "max": float(per_step.max()), "note": "rank 0; steps in which an env takes a rare solver path run longer"},
        "wall_s_timed_region": wall,
    }
    if world == 1 and not args.skip_cpu:
        cores = host_cores()
        cpu_n, cpu_steps = n, 15
        v, dt = time_oracle(env, cpu_n, cpu_steps, 1, use_dr=not args.no_dr)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cpu_n} envs x {cpu_steps} steps of the same workload, float32 C restatement "
                                          f"(oracle/), OpenMP over envs on {cores} threads"}
    if world == 1 and args.extra:
        extra = {}
        for en in (16384, 65536):
            r2 = runtime.EnvRuntime(env.model_desc, env.env_cfg, en, device=local, episode=True)
            sv, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), en))
            r2.set_dr(sv)
            r2.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), en)).view(np.int32)).to(dev))
            a2 = [(torch.rand((en, 12), generator=g, device=dev) - 0.5) for _ in range(4)]
            for t in range(args.settle + 5):
                r2.step(a2[t % 4])
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for t in range(50):
                r2.step(a2[t % 4])
            s1.record()
            torch.cuda.synchronize()
            extra[f"envs_{en}"] = {"value": en / (s0.elapsed_time(s1) / 50 * 1e-3), "unit": UNIT,
                                   "note": "flat ground, full DR, 50 back-to-back steps, state > L2 only at 65536"}
            del r2
        # BASELINE configs[2]: obstacles.py box terrain with randomised pushes, 16384 envs
        env_o = make_env(obstacles=True)
        env_o.set_episode_params(1000, 1)
        en = 16384
        r2 = runtime.EnvRuntime(env_o.model_desc, env_o.env_cfg, en, device=local, episode=True)
        sv, _ = dr.domain_randomize(env_o.sys, prng.split(prng.PRNGKey(2), en))
        r2.set_dr(sv)
        r2.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), en)).view(np.int32)).to(dev))
        a2 = [(torch.rand((en, 12), generator=g, device=dev) - 0.5) for _ in range(4)]
        for t in range(args.settle + 5):
            r2.step(a2[t % 4])
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for t in range(50):
            r2.step(a2[t % 4])
        s1.record()
        torch.cuda.synchronize()
        extra["envs_16384_obstacles"] = {"value": en / (s0.elapsed_time(s1) / 50 * 1e-3), "unit": UNIT,
                                         "note": "configs[2]: 10-box obstacles.py terrain, kicks on, full DR, 50 back-to-back steps"}
        del r2
        # H = 15 at 65,536 envs: 380 MB of state + history per step, i.e. the one case that streams from HBM
        import common as _c
        env15 = _c.make_env(observation_history=15)
        env15.set_episode_params(1000, 1)
        en = 65536
        r3 = runtime.EnvRuntime(env15.model_desc, env15.env_cfg, en, device=local, episode=True)
        sv, _ = dr.domain_randomize(env15.sys, prng.split(prng.PRNGKey(2), en))
        r3.set_dr(sv)
        r3.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), en)).view(np.int32)).to(dev))
        a3 = [(torch.rand((en, 12), generator=g, device=dev) - 0.5) for _ in range(4)]
        for t in range(args.settle + 5):
            r3.step(a3[t % 4])
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for t in range(50):
            r3.step(a3[t % 4])
        s1.record()
        torch.cuda.synchronize()
        v15 = en / (s0.elapsed_time(s1) / 50 * 1e-3)
        extra["envs_65536_H15"] = {"value": v15, "unit": UNIT, "algorithmic_GBps": v15 * b_alg(15) / 1e9,
                                   "note": f"observation_history=15: B_alg = {b_alg(15)} B/env-step, {en * b_alg(15) / 1e6:.0f} MB per step > L2"}
        del r3
        # BASELINE configs[4] (substitute): rollout collection with the fused policy-MLP kernel in the loop, 8192 envs, CUDA graph
        from pupperv3_mjx_b200 import rollout, wrappers
        import functools
        en, T = 8192, 20
        env_r = make_env()
        rand = functools.partial(dr.domain_randomize, rng=prng.split(prng.PRNGKey(2), en))
        tenv = wrappers.wrap(env_r, episode_length=1000, randomization_fn=rand)
        st = tenv.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), en)).view(np.int32)).to(dev))
        for tag, prec, what in (("rollout_8192", 3, "fused 3xTF32 kernel (mma.sync), float32-level accuracy"),
                                ("rollout_8192_tf32", 1, "TF32 on the tcgen05 / tensor-memory kernel (XLA's default float32 matmul precision)")):
            col = rollout.RolloutCollector(tenv, rollout.PolicyMLP.random(env_r.observation_size, precision=prec), st, T, use_cuda_graph=True)
            for _ in range(5):
                col.collect()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(10):
                col.collect()
            s1.record()
            torch.cuda.synchronize()
            extra[tag] = {"value": en * T * 10 / (s0.elapsed_time(s1) * 1e-3), "unit": UNIT,
                          "note": f"configs[4] substitute: unroll 20, policy MLP 72-256-128-128-128-12 (stand-in for the JAX policy; {what}) "
                                  "+ env step, one CUDA graph per unroll"}
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU (default: BASELINE configs[1])")
    ap.add_argument("--settle", type=int, default=100, help="untimed pre-roll steps after reset (steady-state contacts)")
    ap.add_argument("--no-dr", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="diagnostic: keep L2 warm between steps (not a bench number)")
    ap.add_argument("--obstacles", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time the 16384 / 65536 env batches (N=1)")
    args = ap.parse_args()
    import __graft_entry__ as g
    g.build()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
