#!/usr/bin/env python
"""bench.py - env steps/s of the fused MazeEnv.step kernel, beside the CPU restatement.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ID:NENVS] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is one MazeEnv.step over one batch of N_envs lock-step environments (per GPU). The default
workload is the configuration BASELINE.json's north-star target is quoted on: AntUMaze-v0 at 65 536
parallel environments per B200 (configs[2]); environments shard across ranks with no data-path
collective (weak scaling). Prints ONE JSON line (rank 0).

  value     whole-job env steps/s, inputs (actions) already resident in HBM, CUDA-event timed per
            step on the launching stream, L2 flushed between timed steps, max over ranks
  e2e       same metric through the C-ABI host call (mmz_step_host): pinned-host actions in,
            obs/reward/done/info out, copies inside the timed region
  roofline  algorithmic HBM bytes of one launch / its average duration vs the measured copy peak
            roofline.issue / roofline.fp32: the resources that actually bind (warp instructions and fp32 operations
            per launch from the committed ncu capture, profiles/kernel_counters.json, over the LIVE launch time)
  cpu_baseline  the reference's CPU implementation of the path on this box's host cores, bounded sample (rank 0,
            N=1): the REAL reference loop (gym + mujoco-py / mujoco through baseline/_ref, kind "reference") when
            tools/reference_loop.py --probe finds it importable, otherwise the fp64 CPU restatement (oracle/, kind
            "port": never to be read as mujoco-py)
  gather    (N > 1) a second timed loop with the observation all-gather of BASELINE configs[3]
  --impl reference  times that CPU arm alone, all host threads, same config and metric
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
for p in (os.path.join(ROOT, "mujoco-maze_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "env_steps_per_sec"
UNIT = "env steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default="AntUMaze-v0:65536", help="ENV_ID:ENVS_PER_GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--gather-obs", action="store_true", help="(N > 1) gather the observations inside the MAIN timed loop too")
    ap.add_argument("--gather", default="multicast", choices=["peer", "multicast", "nccl", "none"],
                    help="(N > 1) how the second timed loop gathers the observations: peer / multicast = stores from the step "
                         "kernel into every rank's gathered tensor (symmetric memory), nccl = all_gather_into_tensor after it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--ref-envs", type=int, default=0, help="--impl reference: environments per step (0 = auto)")
    return ap.parse_args()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def probe_reference():
    """Is the reference's own loop (gym + MuJoCo + baseline/_ref) runnable on this box? (subprocess: same package name)"""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_loop.py"), "--probe"],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return {"available": False, "why": f"probe failed: {type(e).__name__}: {e}"}


def time_real_reference(env_id, steps, warmup, procs):
    """gym.make(id); reset; step(action_space.sample()) - /root/reference/tests/test_envs.py:7-17 - one process per core."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "reference_loop.py"), "--time", env_id, "--steps", str(steps),
                        "--warmup", str(warmup), "--procs", str(procs)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-500:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def workload_config(model, env_id, n_envs):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": f"{env_id}, {n_envs} parallel envs per GPU", "env_id": env_id, "envs_per_gpu": n_envs,
            "frame_skip": int(model.frame_skip), "integrator": "RK4", "timestep": float(model.timestep),
            "max_episode_steps": int(model.max_episode_steps), "auto_reset": True,
            "actions": "uniform over ctrlrange, pre-generated (on device for the GPU arm)",
            "episode_phase": "all environments reset (seed 0) right before the warm-up steps"}


def make_model(env_id):
    import mujoco_maze  # noqa: F401
    from mujoco_maze import gym

    return gym.make(env_id, num_envs=1).unwrapped.model  # batched flavour: TimeLimit (1000) inside the kernel


def initial_state(model, n, seed):
    """reset_model distributions (ant.py:84-96 / point.py:71-81 / swimmer.py:55-68) in numpy, for the CPU legs."""
    rng = np.random.default_rng(seed)
    nq, nv = int(model.nq), int(model.nv)
    naq, nav = int(model.n_agent_q), int(model.n_agent_v)
    q = np.tile(np.asarray(model.qpos0, float)[:nq], (n, 1))
    v = np.zeros((n, nv))
    q[:, :naq] += rng.uniform(-0.1, 0.1, size=(n, naq))
    kind = int(model.reset_kind)
    if kind == 0:
        v[:, :nav] = 0.1 * rng.random((n, nav))
    elif kind == 1:
        v[:, :nav] = 0.1 * rng.normal(size=(n, nav))
    else:
        v[:, :nav] = rng.uniform(-0.1, 0.1, size=(n, nav))
    return q, v


def action_bounds(model):
    r = np.asarray(model.meta["act_ctrlrange"], float)
    return r[:, 0], r[:, 1]


def time_cpu(model, n_envs, steps, warmup, threads, seed=0):
    """steps/s of the CPU restatement on `threads` host threads over n_envs persistent environments."""
    from oracle import mmz_oracle

    lo, hi = action_bounds(model)
    rng = np.random.default_rng(seed + 1)
    batch = mmz_oracle.OracleBatch(model, n_envs)
    q, v = initial_state(model, n_envs, seed)
    batch.set_state(q, v)
    acts = [rng.uniform(lo, hi, size=(n_envs, len(lo))) for _ in range(min(8, warmup + steps))]
    for s in range(warmup):
        batch.step(acts[s % len(acts)], threads)
    t0 = time.perf_counter()
    for s in range(steps):
        batch.step(acts[(warmup + s) % len(acts)], threads)
    dt = time.perf_counter() - t0
    batch.close()
    return n_envs * steps / dt, dt


def cpu_baseline(model, env_id, target_seconds):
    cores = host_cores()
    p = probe_reference()
    if p.get("available"):
        r = time_real_reference(env_id, 100, 10, cores)  # the reference's own 100-step loop, one process per core
        scale = max(1, int(target_seconds * r["env_steps_per_sec"] / cores / 100))
        if scale > 1:
            r = time_real_reference(env_id, 100 * min(scale, 50), 10, cores)
        return {"value": r["env_steps_per_sec"], "unit": UNIT, "cores": cores, "kind": "reference",
                "sample": f"{cores} processes x {r['steps']} steps of gym.make('{env_id}') after {r['warmup']} warm-up steps, "
                          f"{p.get('mujoco')}, gym {p.get('gym')} (baseline/_ref)"}
    probe_envs = 8 * cores
    rate, _ = time_cpu(model, probe_envs, 2, 1, cores)
    steps = 10
    n_envs = int(max(probe_envs, min(65536, rate * target_seconds / steps)))
    n_envs = max(cores, n_envs // cores * cores)
    value, dt = time_cpu(model, n_envs, steps, 2, cores)
    return {
        "value": value, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{n_envs} {env_id} envs x {steps} steps after 2 warm-up steps ({dt:.1f} s), fp64 CPU restatement "
                  f"(oracle/mmz_oracle.c, OpenMP) - NOT mujoco-py: {p.get('why', 'reference not importable')}",
    }


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_bytes_per_env_step(model, with_info=True):
    """HBM bytes one env-step must move: persisted state rows read + written, action in, obs/reward/done(/info) out.

    State rows: qpos, qvel, qacc (solver warm start), 3 per observed body; counters t and n_reset (int32)."""
    rows = int(model.nq) + 2 * int(model.nv) + 3 * int(model.nobj)
    b = 2 * 4 * rows + 2 * 2 * 4 + 4 * int(model.nu) + 4 * int(model.obs_dim) + 4 + 1
    return b + (16 if with_info else 0)


def survey_bytes_per_env_step(model):
    """B = 4 (2 (nq + nv) + nu + obs_dim + 3) + 1 (no persisted warm start, counters or info)."""
    return 4 * (2 * (int(model.nq) + int(model.nv)) + int(model.nu) + int(model.obs_dim) + 3) + 1


def run_reference(args, env_id, n_envs):
    """The reference arm: the reference's CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = make_model(env_id)
    cores = host_cores()
    p = probe_reference()
    if p.get("available"):
        r = time_real_reference(env_id, max(100, args.steps), max(10, args.warmup), cores)
        value, dt, steps_timed = r["env_steps_per_sec"], r["steps"] / (r["env_steps_per_sec"] / cores), r["steps"]
        kind, dtype = "reference", "f64"
        sample = (f"{cores} processes x {r['steps']} steps of gym.make('{env_id}') (the reference's own loop, "
                  f"{p.get('mujoco')}, gym {p.get('gym')}, baseline/_ref)")
        ms = 1e3 * dt / steps_timed
        note = "reference arm = the unmodified reference through gym.make, one process per host core"
    else:
        if args.ref_envs:
            sample_envs = args.ref_envs
        else:
            rate, _ = time_cpu(model, 8 * cores, 2, 1, cores)
            budget = 150.0  # seconds for the whole run
            sample_envs = int(rate * budget / max(1, args.steps + args.warmup))
            sample_envs = max(cores, min(n_envs, sample_envs) // cores * cores)
        value, dt = time_cpu(model, sample_envs, args.steps, args.warmup, cores)
        kind, dtype = "port", "f64"
        sample = f"{sample_envs} envs per step x {args.steps} steps, fp64 CPU restatement (oracle/), OpenMP {cores} threads"
        ms = 1e3 * dt / args.steps
        note = ("reference arm = fp64 CPU restatement of MazeEnv.step (oracle/), NOT mujoco-py: " + p.get("why", ""))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": workload_config(model, env_id, n_envs),
        "details": {"note": note, "probe": p},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    env_id, n_envs = args.workload.split(":")
    n_envs = int(n_envs)
    if args.impl == "reference":
        run_reference(args, env_id, n_envs)
        return

    import torch
    import torch.distributed as dist

    from mujoco_maze.backend import BatchedSim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    model = make_model(env_id)
    sim = BatchedSim(model, n_envs, dev, auto_reset=True, env_offset=rank * n_envs)
    nu, od = sim.nu, sim.obs_dim
    lo, hi = action_bounds(model)
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    lo_t, hi_t = torch.tensor(lo, device=dev, dtype=torch.float32), torch.tensor(hi, device=dev, dtype=torch.float32)
    acts = [lo_t + (hi_t - lo_t) * torch.rand((n_envs, nu), generator=gen, device=dev) for _ in range(8)]
    obs = torch.empty((n_envs, od), device=dev)
    rew = torch.empty((n_envs,), device=dev)
    done = torch.empty((n_envs,), device=dev, dtype=torch.uint8)
    info = torch.empty((n_envs, 4), device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gather_state = {"mode": "none", "nccl_out": None, "peer": None}

    def set_gather(mode):
        """none | nccl (all_gather_into_tensor after the kernel) | peer / multicast (stores from the step kernel)"""
        if gather_state["peer"] is not None and mode not in ("peer", "multicast"):
            gather_state["peer"].close()
        if mode == "nccl" and gather_state["nccl_out"] is None:
            gather_state["nccl_out"] = torch.empty((world * n_envs, od), device=dev)
        if mode in ("peer", "multicast") and gather_state["peer"] is None:
            from mujoco_maze.sharding import PeerObsGatherer

            gather_state["peer"] = PeerObsGatherer(sim, world * n_envs, rank * n_envs, multicast=(mode == "multicast"))
        elif mode in ("peer", "multicast"):
            sim.set_obs_peers(gather_state["peer"]._ptrs, rank * n_envs, gather_state["peer"].multicast)
        gather_state["mode"] = mode

    def one_step(i):
        sim.step_into(acts[i % len(acts)], obs, rew, done, info)
        if gather_state["mode"] == "nccl":
            dist.all_gather_into_tensor(gather_state["nccl_out"], obs)
        elif gather_state["mode"] in ("peer", "multicast"):
            gather_state["peer"].sync()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(first, steps):
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        for i in range(steps):
            flush.zero_()
            starts[i].record()
            one_step(first + i)
            ends[i].record()
        barrier()
        total = torch.tensor([sum(s.elapsed_time(e) for s, e in zip(starts, ends))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item())

    if args.gather_obs and world > 1:
        set_gather(args.gather)
    sim.reset(seed=0)
    for i in range(args.warmup):
        one_step(i)
    barrier()

    # ---- device-resident timing: one CUDA-event pair per step, L2 flushed between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = sim.launch_count
    total_ms = timed_loop(args.warmup, args.steps)
    launches = sim.launch_count - launches0
    done_frac = float((done & 1).float().mean().item())
    unstable = int(((done & 4) != 0).sum().item())

    # ---- BASELINE configs[3]: the same loop with the observations of all ranks gathered into one tensor per rank
    gather = None
    if world > 1 and args.gather != "none" and not args.gather_obs:
        mode, why = args.gather, None
        try:
            set_gather(mode)
        except Exception as e:  # noqa: BLE001  (no symmetric memory on this box: fall back to the NCCL collective)
            why, mode = f"{type(e).__name__}: {e}"[:200], "nccl"
            set_gather(mode)
        sim.reset(seed=0)  # the same episode phase (and, the physics being deterministic, the same work) as the loop above
        for i in range(args.warmup):
            one_step(i)
        barrier()
        gsteps = args.steps
        g_ms = timed_loop(args.warmup, gsteps)
        check = None
        if mode != "nccl":  # the fused gather must equal the collective bit for bit
            ref = torch.empty((world * n_envs, od), device=dev)
            dist.all_gather_into_tensor(ref, obs)
            check = bool(torch.equal(ref, gather_state["peer"].out))
        if mode == "multicast" and not gather_state["peer"].multicast:
            mode = "peer"  # no NVLS multicast on this box: per-peer stores
        gather = {"ms_per_step": g_ms / gsteps, "ms_per_step_no_gather": total_ms / args.steps, "bytes_per_rank": n_envs * od * 4,
                  "method": {"peer": "stores from the step kernel into every rank's gathered tensor (symmetric memory, P2P)",
                             "multicast": "multimem.st from the step kernel through one NVLS multicast address",
                             "nccl": "dist.all_gather_into_tensor after the step kernel"}[mode],
                  "steps": gsteps, "equals_nccl_all_gather": check, "fallback_reason": why}
        set_gather("none")
        sim.set_obs_peers([], 0)

    # ---- end to end through the C-ABI host call: pinned host buffers, copies inside the timed region
    # (the same episode phase and the same actions as the device-resident loop above: the environments are reset again,
    # the warm-up steps go through the host call too)
    h_acts = [a.cpu().pin_memory() for a in acts]
    h_obs = torch.empty((n_envs, od), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((n_envs,), dtype=torch.float32).pin_memory()
    h_done = torch.empty((n_envs,), dtype=torch.uint8).pin_memory()
    h_info = torch.empty((n_envs, 4), dtype=torch.float32).pin_memory()
    e2e_steps = args.steps
    sim.reset(seed=0)
    for i in range(args.warmup):
        sim.step_host(h_acts[i % len(h_acts)], h_obs, h_rew, h_done, h_info)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + e2e_steps):
        sim.step_host(h_acts[i % len(h_acts)], h_obs, h_rew, h_done, h_info)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        sm_max_mhz = 1965.0
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                pk = json.load(f)
            peak, peak_src = float(pk["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
            sm_max_mhz = float(pk.get("sm_max_mhz", sm_max_mhz))
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        ms_per_step = total_ms / args.steps
        t_launch = ms_per_step * 1e-3
        b_full, b_survey = algorithmic_bytes_per_env_step(model), survey_bytes_per_env_step(model)
        achieved = b_full * n_envs / t_launch / 1e9
        # counters of ONE launch of this workload from the committed ncu capture (tools/ncu_counters.py); the live launch
        # time above turns them into the utilisation of the resources that actually bind the step
        counters, issue, fp32, traffic = None, None, None, None
        cpath = os.path.join(ROOT, "profiles", "kernel_counters.json")
        if os.path.exists(cpath):
            with open(cpath) as f:
                counters = json.load(f).get(args.workload)
        clk_hz = 1e6 * float((clocks or {}).get("sm_mhz") or sm_max_mhz)
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        if counters:
            traffic = counters["dram_bytes"]
            issue = {"warp_inst_per_launch": counters["warp_inst"], "frac": counters["warp_inst"] / (t_launch * n_sm * 4 * clk_hz),
                     "peak": "4 warp instructions / cycle / SM at the SM clock sampled during the timed region",
                     "stall_share_pct": counters.get("stall_share_pct"), "source": "profiles/kernel_counters.json <- " + counters["source"]}
            fp32 = {"flop_per_launch": counters["fp32_flop"], "achieved_tflops": counters["fp32_flop"] / t_launch / 1e12,
                    "peak_tflops": n_sm * 128 * 2 * clk_hz / 1e12,
                    "frac": counters["fp32_flop"] / t_launch / (n_sm * 128 * 2 * clk_hz),
                    "flop_per_env_step": counters["fp32_flop"] / n_envs}
        line = {
            "metric": METRIC, "value": world * n_envs * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(model, env_id, n_envs),
            "details": {"l2": "flushed (256 MiB write) between timed steps", "gather_obs_in_main_loop": gather_state["mode"] != "none" and bool(args.gather_obs),
                        "kernel": sim.kernel_config, "done_frac_last_step": done_frac, "unstable_last_step": unstable,
                        "line_search_tolerance": 1e-2},
            "clocks": clocks,
            "e2e": {"value": world * n_envs * e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": n_envs * nu * 4, "d2h_bytes_per_step": n_envs * (od * 4 + 4 + 1 + 16),
                    "steps": e2e_steps, "api": "mmz_step_host (C ABI, synchronous; pinned host buffers mapped into the device address space: one launch, the blocks read the actions and write the results over PCIe as they finish)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "traffic_note": "dram__bytes_read + write of one launch under ncu's default cache control (caches "
                                         "flushed before every replay, i.e. cold like this bench's flushed L2): the reads are "
                                         "the state tile + actions; the outputs stay in the 126 MB L2 past the end of the launch",
                         "bytes_per_env_step": b_full, "bytes_per_env_step_survey": b_survey,
                         "achieved_survey_bytes": b_survey * n_envs / t_launch / 1e9,
                         "kernel": sim.kernel_config.get("kernel", "maze_kernel"),
                         "issue": issue, "fp32": fp32,
                         "note": "HBM is the nominal bound of a one-pass state update and is 3 orders of magnitude away; "
                                 "what binds is the issue rate between block barriers and the waits at them (issue, fp32, "
                                 "stall_share_pct; DESIGN.md section 5)"},
        }
        if gather is not None:
            line["gather"] = gather
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(model, env_id, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
