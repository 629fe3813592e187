"""GPU parity at the BASELINE.json sizes themselves (VERDICT r1 "parity gaps"), the fused entries added
in round 2 (imagine_and_returns, value_update), the fp16 range guard, a negative shape test, the
2-rank NCCL checks and the patched, unmodified ``Dreamer.train_step`` on the device.  All through the C ABI."""
import os
import subprocess
import sys

import pytest
import torch

import big_dreamer_b200 as bd
from oracle import rssm_oracle as orc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(autouse=True)
def _restore():
    yield
    bd.set_precision("fp32")


# ----------------------------------------------------------------------------------------------
# c2 (BASELINE configs[1]) at its full size against the oracle: 2 500 start states, T = 14
# ----------------------------------------------------------------------------------------------
C2 = dict(Be=200, Hi=200, S=30, A=1, E=8, N=2500, H=15, act="ELU")


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_c2_full_size_vs_oracle(prec, tol):
    """beliefs, states, means, stds, entropy, reward, value, returns, loss and actor gradients of the whole
    actor-loss step at N = 2500 (the oracle does it in ~0.2 s on the host): fp32 check mode <= 1e-4,
    fp16 tcgen05 mode <= 1e-2 (north_star), err = max|delta| / max|ref| per tensor."""
    res = pu.run_imagine_case(C2, seed=11, precision=prec)
    errs = res["errors"]
    print(prec, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e < tol, (k, errs)


# ----------------------------------------------------------------------------------------------
# c3 (BASELINE configs[2]) in full: 1000 candidates x 12 steps x 10 iterations, top-100
# ----------------------------------------------------------------------------------------------
C3 = dict(Be=200, Hi=200, S=30, A=1, E=8, B=1, C=1000, K=100, H=12, iters=10, act="ELU")


def _c3_inputs(seed=5):
    d = C3
    trans, _, reward, _ = orc.make_models(seed, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    g = torch.Generator().manual_seed(seed + 100)
    s0, b0 = orc.make_latents(seed, d["B"], d["Be"], d["S"])
    ea = torch.randn(d["iters"], d["H"], d["B"], d["C"], d["A"], generator=g)
    es = torch.randn(d["iters"], d["H"], d["B"] * d["C"], d["S"], generator=g)
    with torch.no_grad():
        ref, trace = orc.cem_plan(trans, reward, d["act"], 0.1, d["A"], d["H"], d["iters"], d["C"], d["K"],
                                  b0, s0, ea, es, return_trace=True)
    return trans, reward, s0, b0, ea, es, ref, trace


def test_c3_full_fp32_elites_bit_exact_all_iterations():
    """north_star: "CEM elite indices bit-exact in fp32 check mode" -- at the named config, in every one
    of the 10 iterations (src/planner.py:28-90), plus the final action."""
    d = C3
    trans, reward, s0, b0, ea, es, ref, trace = _c3_inputs()
    bd.set_precision("fp32")
    mods = pu.build_gpu_models(d, trans, reward_sd=reward)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    out = pl(b0.cuda(), s0.cuda(), noise=dict(eps_act=ea.cuda(), eps_s=es.cuda()), trace=True)
    got = torch.sort(pl.last_trace["topk"].cpu(), dim=2)[0]
    for it in range(d["iters"]):
        want = torch.sort(trace[it]["topk"], dim=1)[0]
        assert torch.equal(got[it], want), f"iteration {it}: elite sets differ"
        assert pu.relerr(pl.last_trace["returns"][it], trace[it]["returns"]) < 1e-4
    assert pu.relerr(out, ref) < 1e-4


def test_c3_full_fp16_final_action_and_overlap():
    """fp16 tcgen05 mode at c3: 16-bit contractions may flip elites whose returns differ by less than the
    rounding error (SURVEY hard part 8), so the contract is a high elite overlap in EVERY iteration and the
    final action <= 1e-2 (abs) when the last iteration picked the oracle's elite set.  The final
    action is the mean of that iteration's K elites, so every elite that differs from the oracle's may move it
    by (a_i - a_j) / K with candidate actions spread over ~[-2, 2] at the late horizon steps (their standard
    deviation never collapses: they barely change the return): + 4 / K per flipped elite."""
    d = C3
    trans, reward, s0, b0, ea, es, ref, trace = _c3_inputs()
    bd.set_precision("fp16")
    mods = pu.build_gpu_models(d, trans, reward_sd=reward)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    out = pl(b0.cuda(), s0.cuda(), noise=dict(eps_act=ea.cuda(), eps_s=es.cuda()), trace=True)
    overlaps = []
    for it in range(d["iters"]):
        got = set(pl.last_trace["topk"][it, 0].cpu().tolist())
        overlaps.append(len(got & set(trace[it]["topk"][0].tolist())) / d["K"])
    err = float((out.cpu() - ref).abs().max())
    print("fp16 c3: final action abs err %.2e, elite overlap per iteration %s" % (err, overlaps))
    assert min(overlaps) >= 0.95, overlaps
    flipped = round((1.0 - overlaps[-1]) * d["K"])
    assert err < 1e-2 + flipped * 4.0 / d["K"], (err, flipped)


# ----------------------------------------------------------------------------------------------
# fp16 range: weights / latents scaled up until hidden activations leave fp16's range
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("scale", [30.0, 100.0])
def test_fp16_overflow_guard(scale):
    """fp16 operands saturate at +-65504 (cvt.satfinite) instead of becoming inf: a DenseModel whose hidden
    activations exceed fp16's range (weights x scale, inputs x scale) must give FINITE outputs; where the
    fp64 oracle's activations stay inside the range the result must still be accurate, and the bf16 mode
    (fp32 range) must be accurate everywhere."""
    g = torch.Generator().manual_seed(3)
    k1, k2, hid = 200, 30, 200
    sd = orc.make_mlp_sd(g, [k1 + k2] + [hid] * 4 + [1])
    sd["model.0.weight"] = sd["model.0.weight"] * scale
    sd["model.2.weight"] = sd["model.2.weight"] * scale
    x1 = torch.randn(257, k1, generator=g) * scale
    x2 = torch.randn(257, k2, generator=g)
    sd64 = {k: v.double() for k, v in sd.items()}
    ref = orc.dense(sd64, "ELU", x1.double(), x2.double())
    h1 = torch.nn.functional.elu(torch.nn.functional.linear(torch.cat([x1, x2], -1).double(), sd64["model.0.weight"],
                                                            sd64["model.0.bias"]))
    h2 = torch.nn.functional.elu(torch.nn.functional.linear(h1, sd64["model.2.weight"], sd64["model.2.bias"]))
    amax = float(max(h1.abs().max(), h2.abs().max()))
    dm = bd.DenseModel(k1 + k2, hid, 1, "ELU").cuda()
    dm.load_state_dict(sd)
    outs = {}
    for prec in ("fp16", "bf16"):
        bd.set_precision(prec)
        with torch.no_grad():
            outs[prec] = dm(x1.cuda(), x2.cuda()).cpu()
        assert torch.isfinite(outs[prec]).all(), f"{prec}: non-finite output (max |activation| {amax:.3g})"
    print(f"scale {scale}: max |hidden| {amax:.3g}; fp16 err {pu.relerr(outs['fp16'], ref.float()):.2e}, "
          f"bf16 err {pu.relerr(outs['bf16'], ref.float()):.2e}")
    assert pu.relerr(outs["bf16"], ref.float()) < 5e-2
    if amax < 6.0e4:
        assert pu.relerr(outs["fp16"], ref.float()) < 1e-2


def test_dense_model_rejects_wrong_width_in_fast_modes():
    """a wrong feature width must raise in every precision mode (it used to return numbers in
    the tensor-core modes)."""
    dm = bd.DenseModel(230, 200, 1, "ELU").cuda()
    for prec in ("fp32", "fp16"):
        bd.set_precision(prec)
        with pytest.raises(bd.BdError):
            dm(torch.zeros(8, 200, device="cuda"), torch.zeros(8, 31, device="cuda"))
        with pytest.raises(bd.BdError):
            dm(torch.zeros(8, 199, device="cuda"))


def test_actor_with_its_own_activation():
    """ActorModel.activation_function is independent of the transition model's
    (src/models.py:469-481, :131): actor = Tanh with an ELU RSSM must match the oracle (this
    combination runs the fp32 kernels in every mode)."""
    d = dict(Be=48, Hi=40, S=10, A=2, E=8, N=65, H=6, act="ELU")
    trans, actor, reward, value = orc.make_models(2, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    actor["model.8.bias"][d["A"]:] -= 6.0
    s0, b0 = orc.make_latents(2, d["N"], d["Be"], d["S"])
    ea, ee, es = orc.make_imagine_noise(2, d["H"] - 1, d["N"], d["S"], d["A"])
    mods = pu.build_gpu_models(d, trans, reward_sd=reward, value_sd=value)
    act_mod = pu.RefActor(d["Be"], d["S"], d["Hi"], d["A"], "Tanh")
    act_mod.load_state_dict(actor)
    mods.actor = act_mod.cuda()
    agent = pu.agent_ns(mods, d["H"])
    pu.freeze(mods.transition)
    for prec in ("fp32", "fp16"):
        bd.set_precision(prec)
        with torch.no_grad():
            b, s, _, ent = bd.imagine_ahead(agent, s0[None].cuda(), b0[None].cuda(),
                                            dict(eps_a=ea.cuda(), eps_e=ee.cuda(), eps_s=es.cuda()))
            rb, rs, _, rent = _oracle_imagine_mixed(trans, actor, d, s0, b0, ea, ee, es)
        assert pu.relerr(b, rb) < 1e-4 and pu.relerr(s, rs) < 1e-4 and pu.relerr(ent, rent) < 1e-3, prec


def _oracle_imagine_mixed(trans, actor, d, s0, b0, ea, ee, es):
    """orc.imagine_ahead with a different activation for the actor MLP."""
    belief, state = b0, s0
    B, S, E = [], [], []
    for t in range(d["H"] - 1):
        action, ent = orc.get_action(actor, "Tanh", belief, state, ea[t], ee[t])
        belief, state, _, _ = orc.transition_step(trans, d["act"], 0.1, state, action, belief, es[t])
        B.append(belief); S.append(state); E.append(ent)
    return torch.stack(B), torch.stack(S), None, torch.stack(E)


# ----------------------------------------------------------------------------------------------
# fused entries
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,tol,cluster", [("fp32", 1e-4, None), ("fp16", 1e-2, None), ("fp16", 1e-2, "1"),
                                              ("fp16", 1e-2, "2"), ("bf16", 5e-2, None)])
@pytest.mark.parametrize("d", [dict(Be=200, Hi=200, S=30, A=1, E=8, N=300, H=15, act="ELU"),
                               dict(Be=32, Hi=32, S=30, A=1, E=8, N=130, H=15, act="ELU"),
                               dict(Be=48, Hi=40, S=10, A=3, E=8, N=129, H=7, act="Tanh")])
def test_imagine_and_returns_fused_vs_oracle(d, prec, tol, cluster, monkeypatch):
    """bd.imagine_and_returns (heads + lambda_return fused with the rollout: SURVEY 8b level L2) against the
    oracle's actor-loss block: every output and the actor gradients (src/dreamer.py:313-363).  `cluster`
    forces the column-split cluster size of the rollout engine (1 = one CTA per row tile)."""
    if cluster:
        monkeypatch.setenv("BD_TC_CLUSTER", cluster)
    bd.set_precision(prec)
    trans, actor, reward, value = orc.make_models(7, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    actor["model.8.bias"][d["A"]:] -= 6.0
    s0, b0 = orc.make_latents(7, d["N"], d["Be"], d["S"])
    ea, ee, es = orc.make_imagine_noise(7, d["H"] - 1, d["N"], d["S"], d["A"])
    mods = pu.build_gpu_models(d, trans, actor, reward, value)
    frozen = pu.freeze(mods.transition, mods.reward, mods.critic)
    agent = pu.agent_ns(mods, d["H"])
    noise = dict(eps_a=ea.cuda(), eps_e=ee.cuda(), eps_s=es.cuda())
    beliefs, states, (means, stds), entropy, rew, val, ret = bd.imagine_and_returns(
        agent, s0[None].cuda(), b0[None].cuda(), mods.reward, mods.critic, 0.995, 0.95, noise, fused=True)
    loss = -(ret + 1e-5 * entropy.unsqueeze(-1)).mean()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in mods.actor.named_parameters()}
    for p in frozen:
        p.requires_grad_(True)
    inter = dict(beliefs=beliefs, states=states, means=means, stds=stds, entropy=entropy, reward=rew, value=val,
                 returns=ret)
    ref = pu.oracle_actor_loss(d, trans, actor, reward, value, s0, b0, ea, ee, es, dtype=torch.float64)
    errs = pu.compare_actor_loss((loss.detach(), {k: v.detach() for k, v in inter.items()}, grads), ref)
    print(prec, d, {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e < tol, (k, errs)


@pytest.mark.parametrize("save_rows", ["0", "1000000"])
@pytest.mark.parametrize("cluster", ["1", "4"])
@pytest.mark.parametrize("prec,tol", [("fp16", 1e-2), ("bf16", 5e-2)])
@pytest.mark.parametrize("d", [dict(Be=200, Hi=200, S=30, A=1, E=8, N=300, H=15, act="ELU"),
                               dict(Be=48, Hi=40, S=10, A=3, E=8, N=129, H=7, act="Tanh")])
def test_actor_backward_from_saved_hidden_images(d, prec, tol, cluster, save_rows, monkeypatch):
    """The batched actor backward of imagine_ahead reads the hidden activations the rollout saved per (step, row
    tile) (row counts >= BD_ACTOR_SAVE_MIN_ROWS, default 2048) or recomputes the actor's forward pass: both against
    the oracle's actor gradients (src/dreamer.py:363), with and without column-split clusters, on row counts that
    are not a multiple of the 128-row tile."""
    monkeypatch.setenv("BD_ACTOR_SAVE_MIN_ROWS", save_rows)
    monkeypatch.setenv("BD_TC_CLUSTER", cluster)
    res = pu.run_imagine_case(d, seed=5, precision=prec, oracle_dtype=torch.float64)
    print(prec, cluster, save_rows, {k: f"{v:.2e}" for k, v in res["errors"].items()})
    for k, e in res["errors"].items():
        assert e < tol, (k, res["errors"])


@pytest.mark.parametrize("prec,tol", [("fp16", 1e-2), ("bf16", 5e-2)])
def test_mlp_backward_cta_pair_mode(prec, tol, monkeypatch):
    """BD_TC_PAIR2=1: the MLP backward launched in clusters of two CTAs -- tcgen05.mma.cta_group::2 (M = 256: each CTA's
    own 128-row tile, each CTA holding half of every weight stage), multicast commits, the peer's stage completions
    relayed to the leader.  Same oracle comparison as the default launch (actor gradients of the whole actor-loss
    step, src/dreamer.py:363; the heads' and the actor's backward all take even tile counts here)."""
    monkeypatch.setenv("BD_TC_PAIR2", "1")
    monkeypatch.setenv("BD_ACTOR_SAVE_MIN_ROWS", "0")
    d = dict(Be=200, Hi=200, S=30, A=1, E=8, N=512, H=15, act="ELU")
    res = pu.run_imagine_case(d, seed=7, precision=prec, oracle_dtype=torch.float64)
    print(prec, {k: f"{v:.2e}" for k, v in res["errors"].items()})
    for k, e in res["errors"].items():
        assert e < tol, (k, res["errors"])


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_cem_deferred_reward_head_matches_fused(prec, monkeypatch):
    """The CEM rollout with the reward head as ONE batched MLP pass over all (step, candidate) latents (the default
    while the row tiles leave SMs idle) against the head fused into every rollout step (BD_CEM_FUSED_HEAD=1): the same
    16-bit operands and accumulation order, so returns, elite sets and the planned action agree (src/planner.py:53-90)."""
    d = dict(Be=200, Hi=200, S=30, A=1, E=8, B=1, C=1000, K=100, H=12, iters=4, act="ELU")
    monkeypatch.setenv("BD_CEM_FUSED_HEAD", "1")
    fused = pu.run_cem_case(d, seed=4, precision=prec)
    monkeypatch.setenv("BD_CEM_FUSED_HEAD", "0")
    deferred = pu.run_cem_case(d, seed=4, precision=prec)
    print(prec, fused["action_err"], deferred["action_err"], fused["returns_err"], deferred["returns_err"])
    assert abs(fused["returns_err"] - deferred["returns_err"]) < 1e-4
    assert abs(fused["action_err"] - deferred["action_err"]) < 1e-3
    assert deferred["returns_err"] < (1e-2 if prec == "fp16" else 5e-2)


@pytest.mark.parametrize("frozen", [False, True])
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("rows_shape", [(14, 2500), (3, 77)])
def test_heads_pair_matches_separate_calls(prec, rows_shape, frozen, monkeypatch):
    """bd_heads_forward (reward and value models on the same latents as ONE launch, src/dreamer.py:321-322) against the
    two separate DenseModel calls: outputs, input gradients and head weight gradients (heads trainable: the backward
    of the pair is one bd_mlp_backward per head on the hidden images the paired forward left; heads frozen as in the
    behaviour step, :320: bd_heads_backward, both dgrad chains and the sum of their input gradients in one launch)."""
    from big_dreamer_b200 import modules as M
    bd.set_precision(prec)
    d = dict(Be=200, Hi=200, S=30, A=1, E=8, act="ELU")
    trans, actor, reward, value = orc.make_models(2, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, actor, reward, value)
    g = torch.Generator().manual_seed(5)
    T, N = rows_shape
    b = torch.randn(T, N, d["Be"], generator=g).cuda().requires_grad_(True)
    s = torch.randn(T, N, d["S"], generator=g).cuda().requires_grad_(True)
    wr = torch.randn(T, N, 1, generator=g).cuda()
    wv = torch.randn(T, N, 1, generator=g).cuda()
    params = list(mods.reward.parameters()) + list(mods.critic.parameters())
    if frozen:
        pu.freeze(mods.reward, mods.critic)
        params = []

    def run(pair):
        monkeypatch.setenv("BD_HEADS_PAIR", "1" if pair else "0")
        for t in [b, s] + params:
            t.grad = None
        r, v = M.heads_pair(mods.reward, mods.critic, b, s)
        ((r * wr).sum() + (v * wv).sum()).backward()
        return [r.detach().clone(), v.detach().clone(), b.grad.clone(), s.grad.clone()] + [p.grad.clone() for p in params]

    one, two = run(True), run(False)
    assert one[0].shape == (T, N, 1) and one[1].shape == (T, N, 1)
    for x, y in zip(one, two):
        assert pu.relerr(x, y) < 2e-3, pu.relerr(x, y)


def test_entropy_four_rows_per_thread_vs_oracle_and_scalar_kernel(monkeypatch):
    """The 100-sample policy entropy (src/models.py:725-733) at a row count that takes the four-rows-per-thread
    kernel (T * N > 2^16, N a multiple of 4): against the fp64 oracle in the well-conditioned regime and against the
    one-row-per-thread kernel (BD_ENT_SCALAR) on the same inputs, entropy values and actor gradients."""
    d = dict(Be=32, Hi=32, S=30, A=1, E=8, N=5000, H=15, act="ELU")
    res = pu.run_imagine_case(d, seed=3, precision="fp16", oracle_dtype=torch.float64, small_std=True)
    print({k: f"{v:.2e}" for k, v in res["errors"].items()})
    assert res["errors"]["entropy"] < 1e-3, res["errors"]
    for k, e in res["errors"].items():
        assert e < 1e-2, (k, res["errors"])
    monkeypatch.setenv("BD_ENT_SCALAR", "1")
    res1 = pu.run_imagine_case(d, seed=3, precision="fp16", oracle_dtype=torch.float64, small_std=True)
    assert abs(res1["errors"]["entropy"] - res["errors"]["entropy"]) < 1e-5, (res1["errors"], res["errors"])


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_value_update_vs_reference_fixture(prec, tol):
    """bd.value_update against the reference's own critic regression block (src/dreamer.py:369-391;
    fixture frozen from the unmodified reference by oracle/make_golden.py): loss and critic gradients,
    without and with the use_discount weighting."""
    fx = torch.load(os.path.join(GOLDEN, "value_update.pt"))
    dm = fx["dims"]
    bd.set_precision(prec)
    critic = bd.DenseModel(dm["Be"] + dm["S"], dm["Hi"], activation="ELU").cuda()
    critic.load_state_dict(fx["critic"])
    b, s, t, w = (fx[k].cuda() for k in ("beliefs", "states", "target", "discount"))
    for case in fx["cases"]:
        for p in critic.parameters():
            p.grad = None
        loss = bd.value_update(critic, b, s, t, w if case["weighted"] else None)
        assert abs(float(loss) - float(case["loss"])) < tol * max(1.0, abs(float(case["loss"])))
        for k, p in critic.named_parameters():
            assert pu.relerr(p.grad, case["grads"][k]) < tol, (k, case["weighted"])
    # gradients of one update share ONE flat buffer (single all-reduce) and accumulate like autograd
    ptrs = {p.grad.untyped_storage().data_ptr() for p in critic.parameters()}
    assert len(ptrs) == 1
    g0 = critic.model[0].weight.grad.clone()
    bd.value_update(critic, b, s, t, None)
    assert pu.relerr(critic.model[0].weight.grad - g0, fx["cases"][0]["grads"]["model.0.weight"]) < 5 * tol


# ----------------------------------------------------------------------------------------------
# multi-rank (NCCL) checks: skipped with fewer than 2 GPUs
# ----------------------------------------------------------------------------------------------
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_sharded_equals_single():
    """2 ranks over NCCL: the candidate-sharded CEM plan == the single-rank plan, and the row-sharded actor
    gradients after dist.allreduce_grads == the single-GPU gradients (scripts/dist_check.py)."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(ROOT, "scripts", "dist_check.py")],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DIST_CHECK_OK" in r.stdout


# ----------------------------------------------------------------------------------------------
# L1 drop-in on hardware: the UNMODIFIED reference's Dreamer.train_step on patched modules
# ----------------------------------------------------------------------------------------------
def _ref_available():
    from oracle import ref_harness as rh
    return rh.available()


@pytest.mark.skipif(not _ref_available(), reason="reference tree not importable (baseline/_ref or BD_REFERENCE_ROOT)")
@pytest.mark.parametrize("fused", [False, True])
def test_patched_reference_train_step_on_gpu(fused):
    """SURVEY 8c whole-agent oracle: dreamer.Dreamer(params, env).train_step() of the unmodified reference
    with bd.patch() applied runs on the device (planet.device / dreamer.device set to cuda -- the
    reference's `from utils import device` froze them at None) and logs the same scalars as
    the unpatched reference run on the CPU from the same seed and replay buffer.  The two runs draw
    different Gaussian noise (CPU vs CUDA generators), so the comparison is statistical: each logged loss
    is a mean over >= 2 450 rows."""
    from oracle import ref_harness as rh
    params = rh.default_params()
    params.update(experience_size=400, batch_size=20, seq_len=20, seed_steps=0)
    logs = {}
    for mode in ("reference_cpu", "patched_gpu"):
        with rh.agent_modules() as m:
            try:
                dev = torch.device("cuda") if mode == "patched_gpu" else None
                m.planet.device = m.dreamer.device = dev
                if mode == "patched_gpu":
                    bd.set_precision("fp16")
                    bd.patch(fused=fused)
                torch.manual_seed(0)
                agent = m.dreamer.Dreamer(params, rh.FakeEnv(action_size=2))
                names = ("transition_model", "reward_model", "critic", "critic_target", "actor", "encoder",
                         "observation_model")
                if mode == "patched_gpu":
                    assert isinstance(agent.transition_model, bd.TransitionModel)
                    assert isinstance(agent.critic_target, bd.DenseModel)
                    for n in names:                       # same initial weights as the CPU run
                        getattr(agent, n).load_state_dict(logs["_sd"][n])
                else:
                    logs["_sd"] = {n: {k: v.detach().clone() for k, v in getattr(agent, n).state_dict().items()}
                                   for n in names}
                rh.fill_buffer(agent, 300, seed=1)
                torch.manual_seed(1)
                import numpy as np
                np.random.seed(1)
                before = [p.detach().clone() for p in agent.actor.parameters()]
                n0 = bd.load_library().bd_launch_count()
                logs[mode] = agent.train_step()
                assert any(not torch.equal(a, b.detach()) for a, b in zip(before, agent.actor.parameters()))
                if mode == "patched_gpu":
                    from big_dreamer_b200 import modules as M_
                    assert bd.load_library().bd_launch_count() > n0          # the library's kernels ran
                    assert (M_._fused_record is not None) == fused           # fused results were handed out
            finally:
                bd.unpatch()
                bd.set_precision("fp32")
                m.planet.device = m.dreamer.device = None
    ref, got = logs["reference_cpu"], logs["patched_gpu"]
    print("reference (CPU):", ref)
    print("patched (GPU):  ", got)
    for k in ("observation_loss", "reward_loss", "kl_loss", "model_loss", "actor_loss", "policy_entropy", "value_loss"):
        assert k in got and got[k] == got[k], k                  # present and not NaN
        assert abs(got[k] - ref[k]) <= 0.08 * abs(ref[k]) + 0.05, (k, got[k], ref[k])


# ----------------------------------------------------------------------------------------------
# acting path
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("d", [dict(Be=200, Hi=200, S=30, A=1, E=1024, B=1, act="ELU"),
                               dict(Be=48, Hi=40, S=10, A=3, E=32, B=5, act="Tanh")])
def test_act_path_vs_oracle(d, deterministic, prec, tol):
    """bd.ActPath (posterior step + Dreamer.get_action + exploration noise, src/planet.py:370-403) with
    explicit noise against oracle.act_step; SampleDist.mode picks the oracle's sample."""
    bd.set_precision(prec)
    trans, actor, _, _ = orc.make_models(5, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    actor["model.8.bias"][d["A"]:] -= 3.0
    mods = pu.build_gpu_models(d, trans, actor)
    g = torch.Generator().manual_seed(3)
    B = d["B"]
    s0, b0 = orc.make_latents(5, B, d["Be"], d["S"])
    a0 = torch.rand(B, d["A"], generator=g) * 2 - 1
    emb = torch.randn(B, d["E"], generator=g)
    ep, eq = torch.randn(B, d["S"], generator=g), torch.randn(B, d["S"], generator=g)
    ea = torch.randn(*((100, B, d["A"]) if deterministic else (B, d["A"])), generator=g)
    ex = torch.randn(B, d["A"], generator=g)
    with torch.no_grad():
        bo, po, ao = orc.act_step(trans, actor, d["act"], 0.1, b0, s0, a0, emb, ep, eq, ea, deterministic)
    ao_x = torch.clamp(ao + 0.3 * ex, -1, 1)
    act = bd.ActPath(mods.transition, mods.actor, batch=B, action_noise=0.3, deterministic=deterministic)
    noise = dict(eps_prior=ep.cuda(), eps_post=eq.cuda(), eps_act=ea.cuda(), eps_explore=ex.cuda())
    bg, pg, ag = act(b0.cuda(), s0.cuda(), a0.cuda(), emb.cuda(), explore=True, noise=noise)
    assert pu.relerr(bg, bo) < tol and pu.relerr(pg, po) < tol
    if deterministic and prec != "fp32":
        # 16-bit contractions may pick another of the 100 samples when two log-probabilities are closer than
        # the rounding error: the action must then be ANOTHER sample of the oracle's set
        mean, std = orc.actor_mean_std(actor, d["act"], bo, po)
        cand = torch.clamp(torch.tanh(mean[None] + ea * std[None]) + 0.3 * ex[None], -1, 1)      # (100,B,A)
        assert float((cand - ag.cpu()[None]).abs().amax(dim=2).amin(dim=0).max()) < tol
    else:
        assert float((ag.cpu() - ao_x).abs().max()) < tol
    # the replayed graph (noise drawn inside): same belief (no noise enters it), actions in range
    bg2, pg2, ag2 = act(b0.cuda(), s0.cuda(), a0.cuda(), emb.cuda(), explore=True)
    torch.cuda.synchronize()
    assert pu.relerr(bg2, bo) < tol
    assert bool(torch.isfinite(pg2).all()) and float(ag2.abs().max()) <= 1.0
    bg3, _, _ = act(bg2, pg2, ag2, emb.cuda(), explore=True)          # outputs fed back as the next inputs
    assert bool(torch.isfinite(bg3).all())


def test_act_path_planner_policy():
    """ActPath with the CEM planner as policy (PlaNet): explicit noise reproduces MPCPlanner on the
    posterior latents; the captured graph replays."""
    bd.set_precision("fp32")
    d = dict(Be=32, Hi=32, S=30, A=2, E=16, B=2, C=64, K=8, H=4, iters=2, act="ELU")
    trans, _, reward, _ = orc.make_models(2, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, reward_sd=reward)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    g = torch.Generator().manual_seed(4)
    B = d["B"]
    s0, b0 = orc.make_latents(2, B, d["Be"], d["S"])
    a0 = torch.rand(B, d["A"], generator=g) * 2 - 1
    emb = torch.randn(B, d["E"], generator=g)
    ep, eq = torch.randn(B, d["S"], generator=g), torch.randn(B, d["S"], generator=g)
    ea = torch.randn(d["iters"], d["H"], B, d["C"], d["A"], generator=g)
    es = torch.randn(d["iters"], d["H"], B * d["C"], d["S"], generator=g)
    with torch.no_grad():
        bo, _, _, po, _ = orc.transition_forward(trans, d["act"], 0.1, s0, a0[None], b0, ep[None], emb[None], None,
                                                 eq[None])
        ref = orc.cem_plan(trans, reward, d["act"], 0.1, d["A"], d["H"], d["iters"], d["C"], d["K"], bo[0], po[0],
                           ea, es)
    act = bd.ActPath(mods.transition, pl, batch=B)
    noise = dict(eps_prior=ep.cuda(), eps_post=eq.cuda(), planner=dict(eps_act=ea.cuda(), eps_s=es.cuda()))
    bg, pg, ag = act(b0.cuda(), s0.cuda(), a0.cuda(), emb.cuda(), noise=noise)
    assert pu.relerr(bg, bo[0]) < 1e-4 and pu.relerr(pg, po[0]) < 1e-4
    assert pu.relerr(ag, ref) < 1e-4
    bg2, _, ag2 = act(b0.cuda(), s0.cuda(), a0.cuda(), emb.cuda())
    torch.cuda.synchronize()
    assert pu.relerr(bg2, bo[0]) < 1e-4 and bool(torch.isfinite(ag2).all())
