#!/usr/bin/env python
"""Benchmark of the RSSM hot path (BASELINE.json metric): imagined latent steps/s, forward +
BPTT backward of Dreamer's actor loss, on synthetic latents of the named shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision fp16|bf16|fp32] [--rows R] [--no-cpu-baseline]

One "step" = one pass of the hot path over one batch of start states:
imagine_ahead (actor + 100-sample entropy + embed + GRU + prior + sample, T = 14 transitions)
-> reward and value heads -> lambda_return -> actor loss -> backward to the actor gradients
(+ one NCCL all-reduce of those gradients when N > 1).  The optimizer step is not part of the
metric.  Workload at every N: BASELINE.json configs[1] per GPU (weak scaling).

Prints ONE JSON line (see README / DESIGN.md for the keys).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# BASELINE.json configs[1]: Dreamer default RSSM
CFG = dict(Be=200, Hi=200, S=30, A=1, E=8, H=15, act="ELU")
ROWS_DEFAULT = 2500                # batch 50 x chunk 50 start states
DISCOUNT, LAMBDA, ENT_W = 0.995, 0.95, 1e-5
L2_FLUSH_BYTES = 256 << 20         # > 126 MB L2


def algorithmic_flops_per_row_step(d):
    """FLOP = 2*MAC; fwd + bwd of one imagined row-step (no recompute, no padding)."""
    Be, Hi, S, A = d["Be"], d["Hi"], d["S"], d["A"]
    embed = (S + A) * Be
    gru = 6 * Be * Be
    prior = Be * Hi + 2 * S * Hi
    head = (Be + S) * Hi + 3 * Hi * Hi + Hi
    actor = (Be + S) * Hi + 3 * Hi * Hi + 2 * A * Hi
    fwd = embed + gru + prior + actor + 2 * head
    bwd = (embed + gru + prior) + 2 * head + actor + (3 * Hi * Hi + 2 * A * Hi)
    return 2 * (fwd + bwd)


def kernel_macs_per_row_step(d):
    """Algorithmic MACs per imagined row-step attributed to each of the library's kernels
    (no recompute, no padding): the roofline numerators."""
    Be, Hi, S, A = d["Be"], d["Hi"], d["S"], d["A"]
    embed, gru, prior = (S + A) * Be, 6 * Be * Be, Be * Hi + 2 * S * Hi
    head = (Be + S) * Hi + 3 * Hi * Hi + Hi
    actor = (Be + S) * Hi + 3 * Hi * Hi + 2 * A * Hi
    return {"rollout_fwd": embed + gru + prior + actor, "mlp_fwd": 2 * head,
            "bptt": embed + gru + prior, "mlp_bwd": 2 * head + 3 * Hi * Hi + 2 * A * Hi,
            "wgrad": actor, "entropy": 0}


PROF_IDS = {"rollout_fwd": 0, "mlp_fwd": 1, "bptt": 2, "mlp_bwd": 3, "wgrad": 4, "entropy": 5}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        j = json.load(open(p))
        return dict(bf16_burst=j["bf16_tflops"], bf16_sustained=j["bf16_tflops_sustained"],
                    hbm=j["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons, smax = [], set(), None
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); smax = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=smax, reasons=sorted(reasons),
                    samples=len(sm))


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------
def cpu_actor_step(models, s0, b0, noise, d):
    from oracle import rssm_oracle as orc
    trans, actor, reward, value = models
    for v in actor.values():
        v.grad = None
    loss, _ = orc.actor_loss(trans, actor, reward, value, d["act"], 0.1, d["H"], s0[None], b0[None],
                             *noise, DISCOUNT, LAMBDA, ENT_W)
    loss.backward()
    return float(loss)


def cpu_setup(d, rows, seed=0):
    from oracle import rssm_oracle as orc
    trans, actor, reward, value = orc.make_models(seed, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    actor = {k: v.requires_grad_(True) for k, v in actor.items()}
    s0, b0 = orc.make_latents(seed, rows, d["Be"], d["S"])
    noise = orc.make_imagine_noise(seed, d["H"] - 1, rows, d["S"], d["A"])
    return (trans, actor, reward, value), s0, b0, noise


def time_cpu(d, rows, steps, warmup):
    models, s0, b0, noise = cpu_setup(d, rows)
    for _ in range(warmup):
        cpu_actor_step(models, s0, b0, noise, d)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_actor_step(models, s0, b0, noise, d)
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, rows, T = CFG, args.rows, CFG["H"] - 1
    steps = min(args.steps, 20)
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it is allowed
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    ts = time_cpu(d, rows, steps, max(1, min(args.warmup, 2)))
    mean = sum(ts) / len(ts)
    val = rows * T / mean
    cores = torch.get_num_threads()
    sample = (f"oracle port of Dreamer.imagine_ahead + heads + lambda_return + actor backward "
              f"(torch CPU, {cores} threads), full {rows} start states x {T} transitions per step, "
              f"{steps} timed steps")
    print(json.dumps({
        "impl": "reference", "metric": "imagined_latent_steps_per_sec_fwd_bwd", "value": val,
        "unit": "steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(rows, "fp32"),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


CEM_CFG = dict(B=1, C=1000, K=100, H=12, iters=10)      # BASELINE configs[2]


def cem_block(bd, orc, pu, dev, precision, with_cpu):
    """Secondary metric: CEM candidate evaluations/s of MPCPlanner.forward (BASELINE configs[2])."""
    d = dict(CFG, **CEM_CFG)
    trans, _, reward, _ = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, reward_sd=reward, device=dev)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    belief = torch.zeros(d["B"], d["Be"], device=dev)      # episode start (src/main.py:93-94)
    state = torch.zeros(d["B"], d["S"], device=dev)
    for _ in range(3):
        pl(belief, state)
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        pl(belief, state)                                  # draws its own noise, as the reference does
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / reps
    # the same plan (noise drawn inside) captured once and replayed as one CUDA graph
    cap = bd.CapturedStep(lambda: pl(belief, state))
    for _ in range(3):
        cap.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        cap.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = min(ms_eager, e0.elapsed_time(e1) / reps)
    out = {"metric": "cem_candidate_evaluations_per_sec", "value": d["B"] * d["C"] * d["iters"] / (ms * 1e-3),
           "ms_per_plan_eager": ms_eager,
           "unit": "candidate_evals/s", "ms_per_plan": ms, "candidate_steps_per_sec":
           d["B"] * d["C"] * d["iters"] * d["H"] / (ms * 1e-3), "config": dict(CEM_CFG, belief_size=d["Be"],
           state_size=d["S"], action_size=d["A"]), "precision": precision}
    if with_cpu:
        g = torch.Generator().manual_seed(0)
        ea = torch.randn(d["iters"], d["H"], d["B"], d["C"], d["A"], generator=g)
        es = torch.randn(d["iters"], d["H"], d["B"] * d["C"], d["S"], generator=g)
        b0, s0 = torch.zeros(d["B"], d["Be"]), torch.zeros(d["B"], d["S"])
        with torch.no_grad():
            orc.cem_plan(trans, reward, d["act"], 0.1, d["A"], d["H"], d["iters"], d["C"], d["K"], b0, s0, ea, es)
            t0 = time.perf_counter()
            for _ in range(3):
                orc.cem_plan(trans, reward, d["act"], 0.1, d["A"], d["H"], d["iters"], d["C"], d["K"], b0, s0, ea, es)
            cpu_s = (time.perf_counter() - t0) / 3
        out["cpu_baseline"] = {"value": d["B"] * d["C"] * d["iters"] / cpu_s, "unit": "candidate_evals/s",
                               "cores": torch.get_num_threads(), "kind": "port",
                               "sample": "3 full plans of the oracle port"}
    return out


OBS_CFG = dict(L=49, B=50, E=1024)      # BASELINE configs[3]


def observe_block(bd, orc, pu, dev, precision):
    """Secondary: TransitionModel.forward with observations (posterior pass, BASELINE configs[3]),
    forward + backward to all transition weights.  M = 50 rows: latency-bound, reported as time."""
    d = dict(CFG, **OBS_CFG)
    trans, _, _, _ = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    tm = pu.build_gpu_models(d, trans, device=dev).transition
    g = torch.Generator().manual_seed(0)
    L, B = d["L"], d["B"]
    s0, b0 = orc.make_latents(0, B, d["Be"], d["S"])
    s0, b0 = s0.to(dev), b0.to(dev)
    actions = (torch.rand(L, B, d["A"], generator=g) * 2 - 1).to(dev)
    emb = torch.randn(L, B, d["E"], generator=g).to(dev)
    nt = torch.ones(L, B, 1, device=dev)

    def step():
        for p_ in tm.parameters():
            p_.grad = None
        o = tm(s0, actions, b0, emb, nt)
        loss = o[0].mean() + o[3].mean() + o[4][0].mean() + o[4][1].mean() + o[2][0].mean()
        loss.backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_eager = e0.elapsed_time(e1) / reps
    cap = bd.CapturedStep(step)          # forward + backward of the pass as one CUDA graph
    for _ in range(3):
        cap.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        cap.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = min(ms_eager, e0.elapsed_time(e1) / reps)
    return {"metric": "observe_row_steps_per_sec", "value": L * B / (ms * 1e-3), "unit": "row-steps/s",
            "ms_per_pass": ms, "ms_per_pass_eager": ms_eager, "config": dict(OBS_CFG, belief_size=d["Be"], state_size=d["S"]),
            "precision": ("fp32 everywhere (persistent 16-CTA cluster kernels, packed FFMA2)" if precision == "fp32" else
                          "fp32 state and small layers; the two big contractions of the persistent cluster kernels on "
                          "TF32 mma.sync"), "pass": "fwd+bwd, full wgrad"}


def value_update_block(bd, orc, pu, dev, precision, rows, T):
    """Secondary: critic regression update on detached imagined (b, s): DenseModel
    forward + backward with weight gradients (src/dreamer.py:369-391)."""
    d = dict(CFG)
    _, _, _, value = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    critic = bd.DenseModel(d["Be"] + d["S"], d["Hi"], activation=d["act"]).to(dev)
    critic.load_state_dict(value)
    g = torch.Generator().manual_seed(0)
    b = torch.tanh(torch.randn(T, rows, d["Be"], generator=g)).to(dev)
    st = (0.5 * torch.randn(T, rows, d["S"], generator=g)).to(dev)
    target = torch.randn(T, rows, 1, generator=g).to(dev)

    def step():
        for p_ in critic.parameters():
            p_.grad = None
        v = critic(b, st)
        loss = 0.5 * ((v - target) ** 2).mean()          # -log N(target; v, 1) up to a constant
        loss.backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return {"metric": "value_update_rows_per_sec", "value": rows * T / (ms * 1e-3), "unit": "rows/s",
            "ms_per_update": ms, "rows": rows * T, "precision": precision}


def workload_config(rows, precision):
    return {"workload": "BASELINE configs[1]: Dreamer default RSSM imagine_ahead + reward/value heads"
                        " + lambda_return, fwd + BPTT actor loss",
            "belief_size": CFG["Be"], "hidden_size": CFG["Hi"], "state_size": CFG["S"],
            "action_size": CFG["A"], "start_states_per_gpu": rows, "planning_horizon": CFG["H"],
            "transitions": CFG["H"] - 1, "entropy_samples": 100, "precision": precision,
            "l2": "flushed between timed iterations (256 MiB write)"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import big_dreamer_b200 as bd
    from big_dreamer_b200 import dist as D_
    from big_dreamer_b200 import _lib
    from oracle import rssm_oracle as orc           # weights/latents recipe only (synthetic data)
    from tests import parity_utils as pu
    import torch.distributed as tdist

    rank, world, local = D_.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = bd.load_library()
    bd.set_precision(args.precision)
    d, rows, T = CFG, args.rows, CFG["H"] - 1

    trans, actor, reward, value = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, actor, reward, value, device=dev)
    pu.freeze(mods.transition, mods.reward, mods.critic)
    agent = pu.agent_ns(mods, d["H"])
    actor_params = list(mods.actor.parameters())
    # each rank owns its own start states (weak scaling): different seed per rank
    s0_h, b0_h = orc.make_latents(rank, rows, d["Be"], d["S"])
    s0_h, b0_h = s0_h.pin_memory(), b0_h.pin_memory()
    s0, b0 = s0_h.to(dev), b0_h.to(dev)
    noise = bd.draw_imagine_noise(T, rows, d["S"], d["A"], dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    CHUNK = 131072   # rows per forward/backward pass: rows are independent, so a large batch is
                     # processed in slices (bounds the saved-for-backward state); grads accumulate

    def compute(s0_, b0_, noise_):
        for p in actor_params:
            p.grad = None
        total, ents = None, []
        n = s0_.shape[0]
        for lo in range(0, n, CHUNK):
            hi = min(n, lo + CHUNK)
            nz = None if noise_ is None else {"eps_a": noise_["eps_a"][:, lo:hi], "eps_s": noise_["eps_s"][:, lo:hi],
                                              "eps_e": noise_["eps_e"][:, :, lo:hi]}
            if nz is not None and (lo > 0 or hi < n):
                nz = {k: v.contiguous() for k, v in nz.items()}
            beliefs, states, _, entropy = bd.imagine_ahead(agent, s0_[None, lo:hi], b0_[None, lo:hi], nz)
            rew = mods.reward(beliefs, states)
            val = mods.critic(beliefs, states)
            ret = bd.lambda_return(rew, val, val[-1], DISCOUNT, LAMBDA)
            loss = -(ret + ENT_W * entropy.unsqueeze(-1)).sum() / (T * n)     # = slice of the global mean
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
            ents.append(entropy.detach().mean() * ((hi - lo) / n))
        return total, torch.stack(ents).sum()

    def step(s0_, b0_, noise_):
        out = compute(s0_, b0_, noise_)
        D_.allreduce_grads(actor_params)
        return out

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        n0 = lib.bd_launch_count()
        for a, b in ev:
            flush.fill_(1)                    # evict L2 (untimed)
            a.record()
            fn()
            b.record()
        barrier()
        launches = lib.bd_launch_count() - n0
        ms = [a.elapsed_time(b) for a, b in ev]
        return ms, launches

    warmup = max(3, args.warmup)
    sampler = ClockSampler(local) if rank == 0 else None
    # (1) device-resident inputs: the kernel-side number (per-kernel CUDA events recorded by the
    # library on its launch stream during the same timed loop)
    import ctypes as C
    lib.bd_prof_enable(0)
    for _ in range(warmup):
        step(s0, b0, noise)
    lib.bd_prof_enable(1)
    ms_dev, launches = timed(lambda: step(s0, b0, noise), args.steps, 0)
    kernels = {}
    for name, kid in PROF_IDS.items():
        ms, n = C.c_float(0), C.c_int(0)
        lib.bd_prof_read(kid, C.byref(ms), C.byref(n))
        if n.value:
            kernels[name] = {"ms_per_step": ms.value / args.steps, "launches_per_step": n.value / args.steps}
    lib.bd_prof_enable(0)
    ms_eager = list(ms_dev)

    # (1b) the same step captured once as a CUDA graph (bd.CapturedStep) and replayed: one graph
    # launch per step instead of ~45 kernel launches + host work.  The gradient all-reduce (N > 1)
    # stays outside the graph.  This is the headline number; the eager pass above provides the
    # per-kernel CUDA-event times and the launch count.
    # (fp32 check mode: 728 launches per step; replaying that graph measured slower than the eager
    # launches, 18.8 vs 18.0 ms, so check mode is timed eagerly)
    use_graph = not args.no_graph and args.precision != "fp32"
    if use_graph:
        cap_dev = bd.CapturedStep(lambda: compute(s0, b0, noise))

        def graph_step():
            out = cap_dev.replay()
            D_.allreduce_grads(actor_params)
            return out
        ms_dev, _ = timed(graph_step, args.steps, warmup)

    # (2) end to end through the public API: host latents in pinned memory, H2D inside the timed
    # region, noise drawn by the API itself (as the reference does), D2H of the logged scalars
    # Every step copies one step's latents H2D (2.3 MB) and reads the two logged scalars back in one
    # D2H; the copy for step i+1 is enqueued on a copy stream while step i computes (what a training
    # loop's batch prefetch does), so only its tail is exposed.
    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch():
        with torch.cuda.stream(copy_stream):
            s = s0_h.to(dev, non_blocking=True)
            b = b0_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return s, b, ev
    nxt = [prefetch()]

    if use_graph:
        s_in, b_in = torch.empty_like(s0), torch.empty_like(b0)

        def e2e_fn():
            loss, ent = compute(s_in, b_in, None)          # noise drawn inside the captured step
            return torch.stack([loss.detach(), ent.detach()])
        cap_e2e = bd.CapturedStep(e2e_fn)

    def e2e_step():
        s, b, ev = nxt[0]
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        s.record_stream(cur)
        b.record_stream(cur)
        if use_graph:
            s_in.copy_(s, non_blocking=True)               # staged latents -> the graph's inputs
            b_in.copy_(b, non_blocking=True)
            res = cap_e2e.replay()
            D_.allreduce_grads(actor_params)
            nxt[0] = prefetch()
            return res.tolist()
        loss, ent = step(s, b, None)
        nxt[0] = prefetch()
        return torch.stack([loss.detach(), ent.detach()]).tolist()
    ms_e2e, _ = timed(e2e_step, args.steps, warmup)
    clocks = sampler.stop() if sampler else None

    def reduce_max(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    total_ms = reduce_max(sum(ms_dev))
    total_e2e_ms = reduce_max(sum(ms_e2e))
    if rank != 0:
        return
    ms_per_step = total_ms / args.steps
    value = world * rows * T / (ms_per_step * 1e-3)
    e2e_value = world * rows * T / (total_e2e_ms / args.steps * 1e-3)
    pk = peaks()
    flops = algorithmic_flops_per_row_step(d) * rows * T           # per GPU per step
    step_tflops = flops / (ms_per_step * 1e-3) / 1e12
    peak = pk["bf16_sustained"]
    macs = kernel_macs_per_row_step(d)
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    if dom is not None and macs[dom] > 0:
        kflops = 2 * macs[dom] * rows * T
        achieved = kflops / (kernels[dom]["ms_per_step"] * 1e-3) / 1e12
        note = (f"dominant kernel '{dom}': algorithmic {kflops / 1e9:.1f} GFLOP per step / its CUDA-event "
                f"time {kernels[dom]['ms_per_step']:.3f} ms (sum of its launches in a step); whole step: "
                f"{flops / 1e9:.1f} GFLOP -> {step_tflops:.1f} TFLOP/s = {step_tflops / peak:.4f} of peak; "
                f"peak = bf16 sustained of {pk['source']} peaks")
    else:
        achieved = step_tflops
        note = (f"algorithmic FLOPs of the whole step ({flops / 1e9:.1f} GFLOP) / CUDA-event step time; "
                f"peak = bf16 sustained of {pk['source']} peaks")
    out = {
        "metric": "imagined_latent_steps_per_sec_fwd_bwd", "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "tf32": "tf32", "fp16": "f16"}[args.precision],
        "data": "synthetic", "config": workload_config(rows, args.precision),
        "e2e": {"value": e2e_value, "unit": "steps/s",
                "h2d_bytes_per_step": int(s0_h.numel() * 4 + b0_h.numel() * 4),
                "d2h_bytes_per_step": 8, "ms_per_step": total_e2e_ms / args.steps,
                "note": "per step: one H2D of the start latents from pinned memory (enqueued on a copy "
                        "stream for the next step while this one computes), noise drawn by the API, "
                        "one D2H of (loss, entropy)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak,
                     # DRAM read+write bytes of that kernel per launch from the committed ncu --set
                     # full capture (profiles/r01b_ncu_full_bptt_c2_raw.csv: 157.7 MB read + 4.0 MB written), valid for the
                     # default workload only
                     "traffic": ({"rollout_fwd": 96.1e6, "bptt": 161.7e6}.get(dom)
                                 if (rows == ROWS_DEFAULT and args.precision == "fp16") else None),
                     "step_frac": step_tflops / peak, "note": note},
        "kernels": kernels,
        "clocks": clocks,
        "ms_min": min(ms_dev), "ms_median": statistics.median(ms_dev),
        "launch": ("one CUDA graph replay per step (bd.CapturedStep); gpu_launches counts the kernels inside "
                   "the replayed graphs" if use_graph else "eager: one launch per kernel"),
        "ms_per_step_eager": sum(ms_eager) / len(ms_eager),
    }
    if world == 1:
        out["cem"] = cem_block(bd, orc, pu, dev, args.precision, not args.no_cpu_baseline)
        out["observe"] = observe_block(bd, orc, pu, dev, args.precision)
        out["value_update"] = value_update_block(bd, orc, pu, dev, args.precision, min(rows, 131072), T)
    if world == 1 and not args.no_cpu_baseline:
        ts = time_cpu(d, rows, 5, 1)
        cores = torch.get_num_threads()
        out["cpu_baseline"] = {
            "value": rows * T / (sum(ts) / len(ts)), "unit": "steps/s", "cores": cores, "kind": "port",
            "sample": f"oracle port (torch CPU, {cores} threads) on the full workload "
                      f"({rows} start states x {T}), 1 warm-up + 5 timed steps"}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("BD_PRECISION", "fp16"),
                    help="fp16 (default: tcgen05, fp16 operands / fp32 accumulate), bf16, or fp32 check mode")
    ap.add_argument("--rows", type=int, default=ROWS_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time the eager (one launch per kernel) path only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    import torch.distributed as tdist
    if tdist.is_initialized():
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
