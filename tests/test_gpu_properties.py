"""Size-independent properties at BASELINE.json's full sizes (the oracle is too slow there):
rows are independent, so evaluating a batch in pieces must reproduce the same rows
bit for bit; repeated calls are deterministic; the tcgen05 path agrees with the fp32 check mode."""
import pytest
import torch

import big_dreamer_b200 as bd
from oracle import rssm_oracle as orc
from tests import parity_utils as pu

pytestmark = pytest.mark.gpu
D = dict(Be=200, Hi=200, S=30, A=1, E=8, H=15, act="ELU")


@pytest.fixture(autouse=True)
def _restore():
    yield
    bd.set_precision("fp32")


def _setup(N, seed=0):
    trans, actor, reward, value = orc.make_models(seed, D["Be"], D["S"], D["A"], D["Hi"], D["E"])
    actor["model.8.bias"][D["A"]:] -= 6.0
    mods = pu.build_gpu_models(D, trans, actor, reward, value)
    pu.freeze(mods.transition, mods.reward, mods.critic)
    s0, b0 = orc.make_latents(seed, N, D["Be"], D["S"])
    noise = bd.draw_imagine_noise(D["H"] - 1, N, D["S"], D["A"], "cuda")
    return mods, s0.cuda(), b0.cuda(), noise


def _imagine(mods, s0, b0, noise):
    with torch.no_grad():
        b, s, (m, sd), e = bd.imagine_ahead(pu.agent_ns(mods, D["H"]), s0[None], b0[None], noise)
    return b, s, m, sd, e


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
@pytest.mark.parametrize("N", [2500, 2 ** 14])
def test_rows_are_independent_and_deterministic(prec, N):
    """c2 (2500 start states) and the first c5 size (2^14): splitting the batch at a non-tile
    boundary gives bit-identical rows, and two runs agree exactly."""
    bd.set_precision(prec)
    mods, s0, b0, noise = _setup(N)
    full = _imagine(mods, s0, b0, noise)
    again = _imagine(mods, s0, b0, noise)
    for x, y in zip(full, again):
        assert torch.equal(x, y)
    cut = N // 2 + 37
    sl = lambda lo, hi: {k: v[:, :, lo:hi].contiguous() if k == "eps_e" else v[:, lo:hi].contiguous()
                         for k, v in noise.items()}
    lo = _imagine(mods, s0[:cut], b0[:cut], sl(0, cut))
    hi = _imagine(mods, s0[cut:], b0[cut:], sl(cut, N))
    for f, a, b in zip(full, lo, hi):
        assert torch.equal(f[:, :cut], a) and torch.equal(f[:, cut:], b)


def test_tc_matches_check_mode_at_c2_size():
    """Full c2 step (fwd + BPTT): fp16 tcgen05 path vs the fp32 check mode on the same inputs."""
    mods, s0, b0, noise = _setup(2500, seed=1)
    mods.actor.requires_grad_(True)
    res = {}
    for prec in ("fp32", "fp16"):
        bd.set_precision(prec)
        res[prec] = pu.gpu_actor_loss(mods, D["H"], s0, b0, noise)
    errs = pu.compare_actor_loss(res["fp16"], res["fp32"])
    for k, e in errs.items():
        assert e < 1e-2, errs
    assert res["fp16"][1]["beliefs"].shape == (14, 2500, 200)


def test_lambda_return_linearity_full_size():
    """lambda_return is linear in (reward, value): R(a x + b y) = a R(x) + b R(y) (2^16 rows)."""
    g = torch.Generator(device="cuda").manual_seed(0)
    T, N = 14, 2 ** 16
    r1, v1, r2, v2 = (torch.randn(T, N, 1, device="cuda", generator=g) for _ in range(4))
    R = lambda r, v: bd.lambda_return(r, v, v[-1], 0.995, 0.95)
    lhs = R(2.0 * r1 - 0.5 * r2, 2.0 * v1 - 0.5 * v2)
    rhs = 2.0 * R(r1, v1) - 0.5 * R(r2, v2)
    assert float((lhs - rhs).abs().max()) < 1e-4
    # lambda = 0 is the 1-step return, lambda = 1 the discounted Monte-Carlo return + bootstrap
    one = bd.lambda_return(r1, v1, v1[-1], 0.9, 0.0)
    nxt = torch.cat([v1[1:], v1[-1:]], 0)
    assert float((one - (r1 + 0.9 * nxt)).abs().max()) < 1e-5


def test_cem_plan_c3_size_invariants():
    """BASELINE configs[2] sizes: the plan is deterministic given the noise, elites are K distinct
    candidates, and the returned action is the mean of the first-step elite actions."""
    d = dict(D, B=1, C=1000, K=100, H=12, iters=10)
    trans, _, reward, _ = orc.make_models(2, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    mods = pu.build_gpu_models(d, trans, reward_sd=reward)
    pl = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    noise = pl.draw_noise(d["B"], "cuda")
    b0, s0 = torch.zeros(1, 200, device="cuda"), torch.zeros(1, 30, device="cuda")
    for prec in ("fp32", "fp16"):
        bd.set_precision(prec)
        a1 = pl(b0, s0, noise=noise, trace=True)
        topk = pl.last_trace["topk"].clone()
        a2 = pl(b0, s0, noise=noise)
        assert torch.equal(a1, a2)
        assert a1.shape == (1, 1) and bool(torch.isfinite(a1).all())
        for it in range(d["iters"]):
            assert len(set(topk[it, 0].tolist())) == d["K"]
            assert int(topk[it, 0].min()) >= 0 and int(topk[it, 0].max()) < d["C"]


@pytest.mark.parametrize("prec", ["fp16", "fp32"])
def test_captured_step_matches_eager(prec):
    """bd.CapturedStep: the whole actor-loss step (imagine_ahead + heads + lambda_return + backward)
    replayed as one CUDA graph gives the bit-identical loss and the same actor gradients (up to the
    order of the fp32 atomics that end the weight-gradient kernels) as the eager launches,
    also after new start latents are copied into the graph's static inputs."""
    bd.set_precision(prec)
    N = 300
    mods, s0, b0, noise = _setup(N)
    agent = pu.agent_ns(mods, D["H"])
    params = list(mods.actor.parameters())

    def fn(s_, b_):
        for p in params:
            p.grad = None
        beliefs, states, _, ent = bd.imagine_ahead(agent, s_[None], b_[None], noise)
        rew, val = mods.reward(beliefs, states), mods.critic(beliefs, states)
        ret = bd.lambda_return(rew, val, val[-1], 0.995, 0.95)
        loss = -(ret + 1e-5 * ent.unsqueeze(-1)).mean()
        loss.backward()
        return [loss.detach()] + [p.grad for p in params]     # the graph's static outputs

    s_in, b_in = s0.clone(), b0.clone()
    step = bd.CapturedStep(fn, [s_in, b_in])
    for scale in (1.0, 0.5):
        s1, b1 = s0 * scale, b0 * scale
        out_g = [t.clone() for t in step(s1, b1)]
        out_e = fn(s1, b1)              # eager launches (rebinds .grad to fresh tensors)
        torch.cuda.synchronize()
        assert torch.equal(out_g[0], out_e[0])
        for g, e in zip(out_g[1:], out_e[1:]):     # weight gradients end in fp32 atomics: order-dependent
            assert pu.relerr(g, e) < 1e-5


@pytest.mark.parametrize("prec", ["fp32", "fp16"])
def test_observe_pass_is_deterministic(prec):
    """The persistent cluster kernels exchange activations between 16 CTAs through L2 behind cluster
    barriers; a missing ordering would show up as run-to-run differences.  Forward outputs and the input
    gradients (no atomics on their path) must be bit-identical across repeated runs; 3 clusters, ragged."""
    bd.set_precision(prec)
    d = dict(Be=200, Hi=200, S=30, A=2, E=96, act="ELU")
    trans = orc.make_models(7, d["Be"], d["S"], d["A"], d["Hi"], d["E"])[0]
    tm = pu.build_gpu_models(d, trans).transition
    g = torch.Generator().manual_seed(1)
    L, B = 12, 150
    s0 = (0.5 * torch.randn(B, d["S"], generator=g)).cuda()
    b0 = torch.tanh(torch.randn(B, d["Be"], generator=g)).cuda()
    actions = (torch.rand(L, B, d["A"], generator=g) * 2 - 1).cuda()
    emb = torch.randn(L, B, d["E"], generator=g).cuda()
    nt = (torch.rand(L, B, 1, generator=g) > 0.1).float().cuda()
    noise = dict(eps_prior=torch.randn(L, B, d["S"], generator=g).cuda(),
                 eps_post=torch.randn(L, B, d["S"], generator=g).cuda())

    def run():
        a = [s0.clone().requires_grad_(True), b0.clone().requires_grad_(True), emb.clone().requires_grad_(True)]
        o = tm(a[0], actions, a[1], a[2], nt, noise=noise)
        outs = [o[0], o[1], o[2][0], o[2][1], o[3], o[4][0], o[4][1]]
        sum(t.square().mean() for t in outs).backward()
        return [t.detach().clone() for t in outs] + [a[0].grad.clone(), a[1].grad.clone()]

    ref = run()
    for _ in range(4):
        for x, y in zip(run(), ref):
            assert torch.equal(x, y)


def test_actor_grads_add_over_row_slices_when_backward_is_chunked():
    """40 000 start states x 14 steps = 560 000 actor rows: above the 2^18 rows the batched actor backward
    takes per pass, so its two-segment input (start latents for step 0, rollout outputs after) is walked in
    chunks.  Rows are independent, so the actor gradients of the whole batch must equal the sum over four
    slices of 10 000 rows (each a single pass)."""
    bd.set_precision("fp16")
    N = 40000
    mods, s0, b0, noise = _setup(N)
    agent = pu.agent_ns(mods, D["H"])
    params = list(mods.actor.parameters())

    def grads(lo, hi):
        for p in params:
            p.grad = None
        nz = {"eps_a": noise["eps_a"][:, lo:hi].contiguous(), "eps_s": noise["eps_s"][:, lo:hi].contiguous(),
              "eps_e": noise["eps_e"][:, :, lo:hi].contiguous()}
        b, s, _, ent = bd.imagine_ahead(agent, s0[None, lo:hi], b0[None, lo:hi], nz)
        rew, val = mods.reward(b, s), mods.critic(b, s)
        ret = bd.lambda_return(rew, val, val[-1], 0.995, 0.95)
        (-(ret + 1e-5 * ent.unsqueeze(-1)).sum() / (14 * N)).backward()
        return [p.grad.clone() for p in params]

    full = grads(0, N)
    parts = [grads(lo, lo + N // 4) for lo in range(0, N, N // 4)]
    for i, g in enumerate(full):
        ref = sum(p[i] for p in parts)
        err = pu.relerr(g, ref)
        print("param", i, f"{err:.2e}")
        assert err < 2e-3, i
