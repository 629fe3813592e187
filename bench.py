#!/usr/bin/env python
"""Benchmark of the RSSM hot path (BASELINE.json metric): imagined latent steps/s, forward +
BPTT backward of Dreamer's actor loss, and CEM candidate evaluations/s, on synthetic latents of
the named shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--precision fp16|bf16|fp32] [--rows R] [--no-cpu-baseline] [--no-extra]

One "step" = one pass of the hot path over one batch of start states:
imagine_ahead (actor + 100-sample entropy + embed + GRU + prior + sample, T = 14 transitions)
-> reward and value heads -> lambda_return -> actor loss -> backward to the actor gradients
(+ one NCCL all-reduce of those gradients when N > 1).  The optimizer step is not part of the
metric.  Headline workload at every N: BASELINE.json configs[1] per GPU (weak
scaling).  Secondary blocks in the same JSON line (every N unless noted):

    c5        BASELINE configs[4]: 2^17 start states per GPU (weak) and 2^20 in total (strong),
              each with its own per-kernel roofline
    cem       BASELINE configs[2]: MPCPlanner.forward, B = 1 and B = 10, candidates sharded over
              the ranks (one all-gather per iteration)
    observe   BASELINE configs[3] (N = 1: the path does not shard -- "replicas only")
    value_update   critic regression update (N = 1)
    act       acting path (N = 1): us per environment step, actor and CEM policies

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
CHUNK = 131072   # rows per forward/backward pass: rows are independent, so a large batch is
                 # processed in slices (bounds the saved-for-backward state); grads accumulate


def macs(d):
    """Algorithmic MACs per row-step of every building block (no recompute, no padding)."""
    Be, Hi, S, A = d["Be"], d["Hi"], d["S"], d["A"]
    return dict(embed=(S + A) * Be, gru=6 * Be * Be, prior=Be * Hi + 2 * S * Hi,
                head=(Be + S) * Hi + 3 * Hi * Hi + Hi,
                actor=(Be + S) * Hi + 3 * Hi * Hi + 2 * A * Hi,
                actor_dgrad=3 * Hi * Hi + 2 * A * Hi)


def algorithmic_flops_per_row_step(d):
    """fwd + bwd of one imagined row-step, FLOP = 2 * MAC."""
    m = macs(d)
    fwd = m["embed"] + m["gru"] + m["prior"] + m["actor"] + 2 * m["head"]
    bwd = (m["embed"] + m["gru"] + m["prior"]) + 2 * m["head"] + m["actor"] + m["actor_dgrad"]
    return 2 * (fwd + bwd)


def kernel_macs_per_row_step(d):
    """Algorithmic MACs per imagined row-step attributed to each of the library's kernel groups:
    the roofline numerators.  The fused rollout carries the heads (forward and dgrad) itself."""
    m = macs(d)
    rec = m["embed"] + m["gru"] + m["prior"]
    return {"rollout_fwd": rec + m["actor"], "mlp_fwd": 2 * m["head"], "bptt": rec,
            "mlp_bwd": 2 * m["head"] + m["actor_dgrad"], "wgrad": m["actor"], "entropy": 0,
            "rollout_fused_fwd": rec + m["actor"] + 2 * m["head"], "bptt_fused": rec + 2 * m["head"],
            "actor_dgrad": m["actor_dgrad"]}


PROF_IDS = {"rollout_fwd": 0, "mlp_fwd": 1, "bptt": 2, "mlp_bwd": 3, "wgrad": 4, "entropy": 5}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        j = json.load(open(p))
        return dict(bf16_burst=j["bf16_tflops"], bf16_sustained=j["bf16_tflops_sustained"],
                    hbm=j["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def ncu_traffic(kernel):
    """DRAM read+write bytes per launch of `kernel` at the headline workload, from the committed
    `ncu --set full` capture of this build (profiles/ncu_traffic.json, written by
    scripts/ncu_traffic.py from the .ncu-rep); None when no capture of this build is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.isfile(p):
        return None, None
    j = json.load(open(p))
    e = j.get("kernels", {}).get(kernel)
    return (e["dram_bytes_per_launch"], j.get("source")) if e else (None, None)


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


def host_threads():
    # torchrun exports OMP_NUM_THREADS=1; the CPU arms use every host core they are allowed
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except (AttributeError, RuntimeError):
        pass
    return torch.get_num_threads()


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference when its tree is importable
# (baseline/_ref or BD_REFERENCE_ROOT: kind "reference"), else the oracle port (kind "port")
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


def make_cpu_step(d, rows):
    """-> (callable running one actor-loss step on the host cores, kind, description)."""
    from oracle import ref_harness as rh
    from oracle import rssm_oracle as orc
    if rh.available():
        mods = rh.build_modules(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"], d["act"])
        s0, b0 = orc.make_latents(0, rows, d["Be"], d["S"])
        return (lambda: rh.ref_actor_step(mods, d["H"], s0[None], b0[None], DISCOUNT, LAMBDA, ENT_W),
                "reference", f"unmodified reference ({rh.REFERENCE_ROOT}): Dreamer.imagine_ahead + DenseModel "
                             "heads + lambda_return + actor_loss.backward(), its own noise draws")
    models, s0, b0, noise = cpu_setup(d, rows)
    return (lambda: cpu_actor_step(models, s0, b0, noise, d), "port",
            "oracle port of Dreamer.imagine_ahead + heads + lambda_return + actor backward")


def time_cpu(d, rows, steps, warmup):
    fn, kind, what = make_cpu_step(d, rows)
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return ts, kind, what


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, rows, T = CFG, args.rows, CFG["H"] - 1
    # a bounded sample: at most 20 timed steps and 2 warm-ups of the full workload (~0.2 s each)
    steps, warm = max(1, min(args.steps, 20)), max(1, min(args.warmup, 2))
    cores = host_threads()
    ts, kind, what = time_cpu(d, rows, steps, warm)
    mean = sum(ts) / len(ts)
    val = rows * T / mean
    sample = (f"{what} (torch CPU, {cores} threads), full {rows} start states x {T} transitions per step, "
              f"{steps} timed steps after {warm} warm-up")
    print(json.dumps({
        "impl": "reference", "metric": "imagined_latent_steps_per_sec_fwd_bwd", "value": val,
        "unit": "steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "requested": {"steps": args.steps, "warmup": args.warmup},
        "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(rows, "fp32"),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
class Ctx:
    """Everything the blocks share: device, library, distributed state, timing helpers."""

    def __init__(self, args):
        import big_dreamer_b200 as bd
        from big_dreamer_b200 import dist as D_
        from oracle import rssm_oracle as orc           # weights/latents recipe only (synthetic data)
        from tests import parity_utils as pu
        import torch.distributed as tdist
        self.bd, self.D, self.orc, self.pu, self.tdist = bd, D_, orc, pu, tdist
        self.rank, self.world, self.local = D_.init_from_env()
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.lib = bd.load_library()
        bd.set_precision(args.precision)
        self.args = args
        self.flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=self.dev)
        self.pk = peaks()

    def barrier(self):
        if self.world > 1:
            self.tdist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed warm-ups, then K steps, each bracketed by CUDA events on the launching stream, L2
        flushed (untimed) before each; barrier + synchronize on both sides.  -> (ms list, launches)."""
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        n0 = self.lib.bd_launch_count()
        for a, b in ev:
            self.flush.fill_(1)                    # evict L2 (untimed)
            a.record()
            fn()
            b.record()
        self.barrier()
        launches = self.lib.bd_launch_count() - n0
        return [a.elapsed_time(b) for a, b in ev], launches

    def reduce_max(self, x):
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.tdist.all_reduce(t, op=self.tdist.ReduceOp.MAX)
        return float(t.item())

    def read_kernels(self, steps):
        import ctypes as C
        out = {}
        for name, kid in PROF_IDS.items():
            ms, n = C.c_float(0), C.c_int(0)
            self.lib.bd_prof_read(kid, C.byref(ms), C.byref(n))
            if n.value:
                out[name] = {"ms_per_step": ms.value / steps, "launches_per_step": n.value / steps}
        return out


class ActorStep:
    """The actor-loss step of Dreamer.train_step (src/dreamer.py:304-367) on `rows` start states of
    this rank, through the public API (bd.imagine_and_returns: one forward launch chain, one backward)."""

    def __init__(self, cx: Ctx, rows: int, pinned: bool = False):
        bd, orc, pu = cx.bd, cx.orc, cx.pu
        d = CFG
        self.cx, self.rows, self.T = cx, rows, d["H"] - 1
        trans, actor, reward, value = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
        self.mods = pu.build_gpu_models(d, trans, actor, reward, value, device=cx.dev)
        pu.freeze(self.mods.transition, self.mods.reward, self.mods.critic)
        self.agent = pu.agent_ns(self.mods, d["H"])
        self.params = list(self.mods.actor.parameters())
        # each rank owns its own start states: different seed per rank
        s0_h, b0_h = orc.make_latents(cx.rank, rows, d["Be"], d["S"])
        if pinned:
            s0_h, b0_h = s0_h.pin_memory(), b0_h.pin_memory()
        self.s0_h, self.b0_h = s0_h, b0_h
        self.s0, self.b0 = s0_h.to(cx.dev), b0_h.to(cx.dev)
        self.noise = bd.draw_imagine_noise(self.T, rows, d["S"], d["A"], cx.dev)
        self.total_rows = rows           # rows the loss mean runs over on this rank

    def compute(self, s0_, b0_, noise_):
        bd, T = self.cx.bd, self.T
        for p in self.params:
            p.grad = None
        total, ents = None, []
        n = s0_.shape[0]
        for lo in range(0, n, CHUNK):
            hi = min(n, lo + CHUNK)
            nz = None if noise_ is None else {"eps_a": noise_["eps_a"][:, lo:hi], "eps_s": noise_["eps_s"][:, lo:hi],
                                              "eps_e": noise_["eps_e"][:, :, lo:hi]}
            if nz is not None and (lo > 0 or hi < n):
                nz = {k: v.contiguous() for k, v in nz.items()}
            out = bd.imagine_and_returns(self.agent, s0_[None, lo:hi], b0_[None, lo:hi], self.mods.reward,
                                         self.mods.critic, DISCOUNT, LAMBDA, nz)
            entropy, ret = out[3], out[6]
            loss = -(ret + ENT_W * entropy.unsqueeze(-1)).sum() / (T * n)     # = slice of the global mean
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
            ents.append(entropy.detach().mean() * ((hi - lo) / n))
        return total, torch.stack(ents).sum()

    def step(self, s0_=None, b0_=None, noise_="own"):
        out = self.compute(self.s0 if s0_ is None else s0_, self.b0 if b0_ is None else b0_,
                           self.noise if isinstance(noise_, str) else noise_)
        self.cx.D.allreduce_grads(self.params)
        return out


def roofline_of(kernels, rows, T, ms_per_step, pk, traffic_for=None):
    """Per-kernel and whole-step tensor roofline (algorithmic FLOPs / CUDA-event time / measured peak)."""
    d = CFG
    peak = pk["bf16_sustained"]
    flops = algorithmic_flops_per_row_step(d) * rows * T
    step_tflops = flops / (ms_per_step * 1e-3) / 1e12
    km = kernel_macs_per_row_step(d)
    fused = "mlp_fwd" not in kernels          # heads ride in the rollout / BPTT kernels
    per = {}
    for name, k in kernels.items():
        mac = km["rollout_fused_fwd" if (fused and name == "rollout_fwd") else
                 "bptt_fused" if (fused and name == "bptt") else name]
        if fused and name == "mlp_bwd":
            mac = km["actor_dgrad"]
        if mac <= 0 or k["ms_per_step"] <= 0:
            continue
        tf = 2 * mac * rows * T / (k["ms_per_step"] * 1e-3) / 1e12
        per[name] = {"ms_per_step": k["ms_per_step"], "tflops": tf, "frac": tf / peak}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"]) if kernels else None
    if dom in per:
        achieved = per[dom]["tflops"]
        note = (f"dominant kernel '{dom}': algorithmic FLOPs per step / its CUDA-event time "
                f"{kernels[dom]['ms_per_step']:.3f} ms (sum of its launches in a step); whole step: "
                f"{flops / 1e9:.1f} GFLOP -> {step_tflops:.1f} TFLOP/s = {step_tflops / peak:.4f} of peak; "
                f"peak = bf16 sustained of {pk['source']} peaks")
    else:
        achieved, note = step_tflops, (f"algorithmic FLOPs of the whole step ({flops / 1e9:.1f} GFLOP) / "
                                       f"CUDA-event step time; peak = bf16 sustained of {pk['source']} peaks")
    traffic, tsrc = ncu_traffic(dom) if traffic_for == "headline" and dom else (None, None)
    out = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
           "frac": achieved / peak, "traffic": traffic, "step_frac": step_tflops / peak,
           "step_tflops": step_tflops, "per_kernel": per, "note": note}
    if tsrc:
        out["traffic_source"] = tsrc
    return out


def measure_actor(cx: Ctx, st: ActorStep, steps, warmup, graph=True):
    """-> dict(ms list of the headline mode, eager ms, kernels, launches)."""
    lib = cx.lib
    lib.bd_prof_enable(0)
    for _ in range(warmup):
        st.step()
    lib.bd_prof_enable(1)
    ms_eager, launches = cx.timed(st.step, steps, 0)
    kernels = cx.read_kernels(steps)
    lib.bd_prof_enable(0)
    ms = list(ms_eager)
    if graph:
        cap = cx.bd.CapturedStep(lambda: st.compute(st.s0, st.b0, st.noise))

        def graph_step():
            out = cap.replay()
            cx.D.allreduce_grads(st.params)
            return out
        ms, _ = cx.timed(graph_step, steps, warmup)
    return dict(ms=ms, ms_eager=ms_eager, kernels=kernels, launches=launches)


def c5_block(cx: Ctx, precision):
    """BASELINE configs[4] (SURVEY hard part 5: the roofline claim belongs here): 2^17 start states per
    GPU (weak) and 2^20 start states in total sharded over the ranks (strong)."""
    T = CFG["H"] - 1
    out = {"config": {"horizon": CFG["H"], "transitions": T, "belief_size": CFG["Be"], "precision": precision,
                      "rows_per_pass": CHUNK}}
    for name, rows, scaling in (("weak_131072_per_gpu", 131072, "weak"),
                                ("strong_1048576_total", (1 << 20) // cx.world, "strong")):
        st = ActorStep(cx, rows)
        steps = 5 if rows <= 131072 else 3
        m = measure_actor(cx, st, steps, 2, graph=False)     # graph adds nothing at these step times
        tot = cx.reduce_max(sum(m["ms"]))
        ms = tot / steps
        rf = roofline_of(m["kernels"], rows, T, ms, cx.pk)
        out[name] = {"scaling": scaling, "rows_per_gpu": rows, "ms_per_step": ms,
                     "value": cx.world * rows * T / (ms * 1e-3), "unit": "steps/s",
                     "step_frac": rf["step_frac"], "step_tflops_per_gpu": rf["step_tflops"],
                     "roofline": {k: rf[k] for k in ("kernel", "achieved", "peak", "frac", "per_kernel")}}
        del st
        torch.cuda.empty_cache()
    out["step_frac"] = out["weak_131072_per_gpu"]["step_frac"]
    return out


CEM_CFG = dict(C=1000, K=100, H=12, iters=10)      # BASELINE configs[2]


def cem_block(cx: Ctx, precision, with_cpu):
    """CEM candidate evaluations/s of MPCPlanner.forward (BASELINE configs[2]) for B = 1 (acting) and
    B = 10 (evaluation, src/main.py:199-233); at N > 1 the candidates are sharded over the ranks
    (MPCPlanner._forward_sharded: one all-gather per iteration)."""
    bd, orc, pu = cx.bd, cx.orc, cx.pu
    d = dict(CFG, **CEM_CFG)
    trans, _, reward, _ = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, reward_sd=reward, device=cx.dev)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    m = macs(d)
    flop_eval = 2 * (m["embed"] + m["gru"] + m["prior"] + m["head"]) * d["H"]     # per candidate evaluation
    res = {}
    for B in (1, 10):
        belief = torch.zeros(B, d["Be"], device=cx.dev)      # episode start (src/main.py:93-94)
        state = torch.zeros(B, d["S"], device=cx.dev)
        gen = torch.Generator(device=cx.dev)

        def plan():
            gen.manual_seed(1234)                  # every rank draws the SAME global noise tensors
            return pl(belief, state, noise=pl.draw_noise(B, cx.dev, generator=gen))
        reps = 10
        ms_list, _ = cx.timed(plan, reps, 3)
        ms_eager = cx.reduce_max(sum(ms_list)) / reps
        ms = ms_eager
        if cx.world == 1:     # the single-rank plan (noise drawn inside) replayed as one CUDA graph
            cap = bd.CapturedStep(lambda: pl(belief, state))
            ms_g, _ = cx.timed(cap.replay, reps, 3)
            ms = min(ms_eager, sum(ms_g) / reps)
        evals = B * d["C"] * d["iters"]
        tf = evals * flop_eval / (ms * 1e-3) / 1e12
        res[f"B{B}"] = {"value": evals / (ms * 1e-3), "unit": "candidate_evals/s", "ms_per_plan": ms,
                        "ms_per_plan_eager": ms_eager,
                        "candidate_steps_per_sec": evals * d["H"] / (ms * 1e-3),
                        "candidates_per_gpu": d["C"] // cx.world,
                        "roofline": {"bound": "tensor", "achieved": tf, "peak": cx.pk["bf16_sustained"],
                                     "unit": "TFLOP/s", "frac": tf / cx.pk["bf16_sustained"],
                                     "note": f"{d['iters'] * d['H']} strictly serial rollout steps of "
                                             f"{B * d['C'] // cx.world} rows per GPU: latency-bound by construction"}}
    out = {"metric": "cem_candidate_evaluations_per_sec", "value": res["B1"]["value"],
           "unit": "candidate_evals/s", "ms_per_plan": res["B1"]["ms_per_plan"], "n_gpus": cx.world,
           "sharding": "candidates over ranks, all-gather of (returns, actions) per iteration" if cx.world > 1 else "none",
           "config": dict(CEM_CFG, belief_size=d["Be"], state_size=d["S"], action_size=d["A"]),
           "precision": precision, "B1": res["B1"], "B10": res["B10"]}
    if with_cpu and cx.rank == 0 and cx.world == 1:
        cores = host_threads()
        g = torch.Generator().manual_seed(0)
        ea = torch.randn(d["iters"], d["H"], 1, d["C"], d["A"], generator=g)
        es = torch.randn(d["iters"], d["H"], d["C"], d["S"], generator=g)
        b0, s0 = torch.zeros(1, d["Be"]), torch.zeros(1, d["S"])
        by_threads = {}
        for nt in sorted({1, cores}):
            torch.set_num_threads(nt)
            with torch.no_grad():
                orc.cem_plan(trans, reward, d["act"], 0.1, d["A"], d["H"], d["iters"], d["C"], d["K"], b0, s0, ea, es)
                t0 = time.perf_counter()
                n = 5 if nt > 1 else 2
                for _ in range(n):
                    orc.cem_plan(trans, reward, d["act"], 0.1, d["A"], d["H"], d["iters"], d["C"], d["K"], b0, s0, ea, es)
                by_threads[nt] = d["C"] * d["iters"] / ((time.perf_counter() - t0) / n)
        torch.set_num_threads(cores)
        best = max(by_threads, key=by_threads.get)
        out["cpu_baseline"] = {"value": by_threads[best], "unit": "candidate_evals/s", "cores": best, "kind": "port",
                               "by_threads": {str(k): v for k, v in by_threads.items()},
                               "sample": "full B=1 plans of the oracle port (1 warm-up + 5 timed at all threads, 2 at 1 thread)"}
    return out


OBS_CFG = dict(L=49, B=50, E=1024)      # BASELINE configs[3]


def observe_block(cx: Ctx, precision, with_cpu):
    """TransitionModel.forward with observations (posterior pass, BASELINE configs[3]), forward +
    backward to all transition weights.  M = 50 rows: latency-bound, reported as time."""
    bd, orc, pu = cx.bd, cx.orc, cx.pu
    d = dict(CFG, **OBS_CFG)
    trans, _, _, _ = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    tm = pu.build_gpu_models(d, trans, device=cx.dev).transition
    g = torch.Generator().manual_seed(0)
    L, B = d["L"], d["B"]
    s0c, b0c = orc.make_latents(0, B, d["Be"], d["S"])
    s0, b0 = s0c.to(cx.dev), b0c.to(cx.dev)
    actions_c = torch.rand(L, B, d["A"], generator=g) * 2 - 1
    emb_c = torch.randn(L, B, d["E"], generator=g)
    actions, emb = actions_c.to(cx.dev), emb_c.to(cx.dev)
    nt = torch.ones(L, B, 1, device=cx.dev)

    def step():
        for p_ in tm.parameters():
            p_.grad = None
        o = tm(s0, actions, b0, emb, nt)
        loss = o[0].mean() + o[3].mean() + o[4][0].mean() + o[4][1].mean() + o[2][0].mean()
        loss.backward()
    reps = 10
    ms_list, _ = cx.timed(step, reps, 3)
    ms_eager = sum(ms_list) / reps
    cap = bd.CapturedStep(step)          # forward + backward of the pass as one CUDA graph
    ms_g, _ = cx.timed(cap.replay, reps, 3)
    ms = min(ms_eager, sum(ms_g) / reps)
    Be, Hi, S, A, E = d["Be"], d["Hi"], d["S"], d["A"], d["E"]
    fwd_mac = (S + A) * Be + 6 * Be * Be + (Be * Hi + 2 * S * Hi) + ((Be + E) * Hi + 2 * S * Hi)
    flops = 2 * 3 * fwd_mac * L * B              # bwd ~ 2 x fwd (8.2 GFLOP per pass)
    tf = flops / (ms * 1e-3) / 1e12
    out = {"metric": "observe_row_steps_per_sec", "value": L * B / (ms * 1e-3), "unit": "row-steps/s",
           "ms_per_pass": ms, "ms_per_pass_eager": ms_eager, "us_per_step_direction": ms * 1e3 / (2 * L),
           "config": dict(OBS_CFG, belief_size=Be, state_size=S),
           "precision": precision, "pass": "fwd+bwd, full wgrad", "multi_gpu": "replicas only",
           "roofline": {"bound": "tensor", "achieved": tf, "peak": cx.pk["bf16_sustained"], "unit": "TFLOP/s",
                        "frac": tf / cx.pk["bf16_sustained"],
                        "note": f"{flops / 1e9:.2f} GFLOP per pass; 2 x {L} strictly serial steps of {B} rows: "
                                "latency-bound by construction (one row tile on the whole GPU)"}}
    if with_cpu and cx.rank == 0:
        cores = host_threads()
        sd = {k: v.clone().requires_grad_(True) for k, v in trans.items()}
        ep, eq = torch.randn(L, B, S, generator=g), torch.randn(L, B, S, generator=g)
        ntc = torch.ones(L, B, 1)

        def cpu_pass():
            for v in sd.values():
                v.grad = None
            o = orc.transition_forward(sd, d["act"], 0.1, s0c, actions_c, b0c, ep, emb_c, ntc, eq)
            (o[0].mean() + o[3].mean() + o[4][0].mean() + o[4][1].mean() + o[2][0].mean()).backward()
        cpu_pass()
        t0 = time.perf_counter()
        for _ in range(5):
            cpu_pass()
        cpu_s = (time.perf_counter() - t0) / 5
        out["cpu_baseline"] = {"value": L * B / cpu_s, "unit": "row-steps/s", "cores": cores, "kind": "port",
                               "ms_per_pass": cpu_s * 1e3,
                               "sample": "oracle port of TransitionModel.forward (observe mode) fwd + bwd, 1 warm-up + 5 passes"}
    return out


def value_update_block(cx: Ctx, precision, rows, T):
    """critic regression update on detached imagined (b, s): bd.value_update = critic forward
    + -Normal(v, 1).log_prob(target).mean() + backward with weight gradients (src/dreamer.py:369-391)."""
    bd, orc = cx.bd, cx.orc
    d = dict(CFG)
    _, _, _, value = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    critic = bd.DenseModel(d["Be"] + d["S"], d["Hi"], activation=d["act"]).to(cx.dev)
    critic.load_state_dict(value)
    g = torch.Generator().manual_seed(0)
    b = torch.tanh(torch.randn(T, rows, d["Be"], generator=g)).to(cx.dev)
    st = (0.5 * torch.randn(T, rows, d["S"], generator=g)).to(cx.dev)
    target = torch.randn(T, rows, 1, generator=g).to(cx.dev)

    def step():
        for p_ in critic.parameters():
            p_.grad = None
        return bd.value_update(critic, b, st, target)
    reps = 10
    ms_list, _ = cx.timed(step, reps, 3)
    ms = sum(ms_list) / reps
    m = macs(d)
    tf = 2 * 3 * m["head"] * rows * T / (ms * 1e-3) / 1e12
    return {"metric": "value_update_rows_per_sec", "value": rows * T / (ms * 1e-3), "unit": "rows/s",
            "ms_per_update": ms, "rows": rows * T, "precision": precision,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": cx.pk["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": tf / cx.pk["bf16_sustained"]}}


def act_block(cx: Ctx, precision, with_cpu):
    """SURVEY 8f-3, the acting path (src/planet.py:370-403 between the encoder and env.step), B = 1: one
    posterior step + policy + exploration noise per environment step, replayed as one CUDA graph
    (bd.ActPath).  Policies: Dreamer's actor (get_action) and PlaNet's CEM planner (BASELINE configs[2])."""
    bd, orc, pu = cx.bd, cx.orc, cx.pu
    d = dict(CFG, E=1024, **CEM_CFG)
    trans, actor, reward, _ = orc.make_models(0, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, actor, reward_sd=reward, device=cx.dev)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    pl.shard_candidates = False
    z = lambda n: torch.zeros(1, n, device=cx.dev)
    emb = torch.randn(1, d["E"], device=cx.dev)
    out = {"metric": "env_steps_per_sec", "unit": "steps/s", "batch": 1, "precision": precision,
           "path": "posterior step (TransitionModel.forward, L = 1, observe) + policy + exploration noise; "
                   "encoder and env.step excluded (out of scope)", "launch": "one CUDA graph replay per step"}
    for name, policy in (("actor", mods.actor), ("cem", pl)):
        act = bd.ActPath(mods.transition, policy, batch=1, action_noise=0.3)
        state = [z(d["Be"]), z(d["S"]), z(d["A"])]

        def step():
            b, s, a = act(state[0], state[1], state[2], emb, explore=True)
            state[0], state[1], state[2] = b, s, a
        reps = 50 if name == "actor" else 10
        ms_list, _ = cx.timed(step, reps, 3)
        ms = sum(ms_list) / reps
        out[name] = {"us_per_env_step": ms * 1e3, "value": 1e3 / ms}
    out["value"] = out["actor"]["value"]
    if with_cpu and cx.rank == 0:
        cores = host_threads()
        g = torch.Generator().manual_seed(0)
        b0, s0, a0 = torch.zeros(1, d["Be"]), torch.zeros(1, d["S"]), torch.zeros(1, d["A"])
        embc = torch.randn(1, d["E"], generator=g)
        ep, eq, ea = (torch.randn(1, d["S"], generator=g), torch.randn(1, d["S"], generator=g),
                      torch.randn(1, d["A"], generator=g))
        with torch.no_grad():
            for _ in range(3):
                orc.act_step(trans, actor, d["act"], 0.1, b0, s0, a0, embc, ep, eq, ea)
            t0 = time.perf_counter()
            for _ in range(200):
                orc.act_step(trans, actor, d["act"], 0.1, b0, s0, a0, embc, ep, eq, ea)
            cpu_s = (time.perf_counter() - t0) / 200
        out["cpu_baseline"] = {"value": 1.0 / cpu_s, "unit": "steps/s", "cores": cores, "kind": "port",
                               "us_per_env_step": cpu_s * 1e6,
                               "sample": "oracle port of the posterior step + get_action (actor policy), 200 steps"}
    return out


def run_ours(args):
    cx = Ctx(args)
    bd, D_ = cx.bd, cx.D
    rank, world, dev = cx.rank, cx.world, cx.dev
    d, rows, T = CFG, args.rows, CFG["H"] - 1
    warmup = max(3, args.warmup)
    st = ActorStep(cx, rows, pinned=True)

    sampler = ClockSampler(cx.local) if rank == 0 else None
    # (1) device-resident inputs.  Eager pass: per-kernel CUDA events recorded by the library on its
    # launch stream + the launch count; then the same step captured once as a CUDA graph
    # (bd.CapturedStep) and replayed: the headline `value`.  The gradient all-reduce (N > 1) stays
    # outside the graph.  (fp32 check mode: ~700 launches per step; replaying that graph measured slower
    # than the eager launches, so check mode is timed eagerly.)
    use_graph = not args.no_graph and args.precision != "fp32"
    m = measure_actor(cx, st, args.steps, warmup, graph=use_graph)
    ms_dev, ms_eager, kernels, launches = m["ms"], m["ms_eager"], m["kernels"], m["launches"]

    # (2) end to end through the public API: host latents in pinned memory, H2D inside the timed
    # region, noise drawn by the API itself (as the reference does), D2H of the logged scalars.
    # The copy for step i+1 is enqueued on a copy stream while step i computes (what a training
    # loop's batch prefetch does), so only its tail is exposed.
    copy_stream = torch.cuda.Stream(device=dev)

    def prefetch():
        with torch.cuda.stream(copy_stream):
            s = st.s0_h.to(dev, non_blocking=True)
            b = st.b0_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return s, b, ev
    nxt = [prefetch()]

    if use_graph:
        s_in, b_in = torch.empty_like(st.s0), torch.empty_like(st.b0)

        def e2e_fn():
            loss, ent = st.compute(s_in, b_in, None)          # noise drawn inside the captured step
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
            D_.allreduce_grads(st.params)
            nxt[0] = prefetch()
            return res.tolist()
        loss, ent = st.step(s, b, None)
        nxt[0] = prefetch()
        return torch.stack([loss.detach(), ent.detach()]).tolist()
    ms_e2e, _ = cx.timed(e2e_step, args.steps, warmup)
    clocks = sampler.stop() if sampler else None

    total_ms = cx.reduce_max(sum(ms_dev))
    total_e2e_ms = cx.reduce_max(sum(ms_e2e))
    ms_per_step = total_ms / args.steps
    value = world * rows * T / (ms_per_step * 1e-3)
    e2e_value = world * rows * T / (total_e2e_ms / args.steps * 1e-3)
    rf = roofline_of(kernels, rows, T, ms_per_step, cx.pk,
                     traffic_for="headline" if (rows == ROWS_DEFAULT and args.precision == "fp16") else None)
    out = {
        "metric": "imagined_latent_steps_per_sec_fwd_bwd", "value": value, "unit": "steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "bf16": "bf16", "tf32": "tf32", "fp16": "f16"}[args.precision],
        "data": "synthetic", "config": workload_config(rows, args.precision),
        "e2e": {"value": e2e_value, "unit": "steps/s",
                "h2d_bytes_per_step": int(st.s0_h.numel() * 4 + st.b0_h.numel() * 4),
                "d2h_bytes_per_step": 8, "ms_per_step": total_e2e_ms / args.steps,
                "note": "per step: one H2D of the start latents from pinned memory (enqueued on a copy "
                        "stream for the next step while this one computes), noise drawn by the API, "
                        "one D2H of (loss, entropy)"},
        "gpu_launches": int(launches),
        "roofline": rf,
        "kernels": kernels,
        "clocks": clocks,
        "ms_min": min(ms_dev), "ms_median": statistics.median(ms_dev),
        "launch": ("one CUDA graph replay per step (bd.CapturedStep); gpu_launches counts the library's kernel "
                   "launches of the eager pass over the same steps" if use_graph else "eager: one launch per kernel"),
        "ms_per_step_eager": sum(ms_eager) / len(ms_eager),
    }
    del st
    torch.cuda.empty_cache()
    if not args.no_extra:
        blocks = {"c5": lambda: c5_block(cx, args.precision),
                  "cem": lambda: cem_block(cx, args.precision, not args.no_cpu_baseline)}
        if world == 1:
            blocks["observe"] = lambda: observe_block(cx, args.precision, not args.no_cpu_baseline)
            blocks["value_update"] = lambda: value_update_block(cx, args.precision, min(rows, 131072), T)
            blocks["act"] = lambda: act_block(cx, args.precision, not args.no_cpu_baseline)
        for name, fn in blocks.items():
            try:
                out[name] = fn()
            except Exception as e:      # noqa: BLE001 - a failing secondary block must not lose the headline line
                out[name] = {"error": f"{type(e).__name__}: {e}"}
                cx.barrier()
            torch.cuda.empty_cache()
    if rank != 0:
        return
    if world == 1 and not args.no_cpu_baseline:
        cores = host_threads()
        ts, kind, what = time_cpu(d, rows, 5, 1)
        out["cpu_baseline"] = {
            "value": rows * T / (sum(ts) / len(ts)), "unit": "steps/s", "cores": cores, "kind": kind,
            "sample": f"{what} (torch CPU, {cores} threads) on the full workload "
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
    ap.add_argument("--no-extra", action="store_true", help="headline only (skip the c5 / cem / observe / value blocks)")
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
