"""Shared parity helpers: run the CUDA path and the CPU oracle on identical weights,
inputs and noise, and report err = max|delta| / max|ref| per tensor."""
import types

import torch

import big_dreamer_b200 as bd
from big_dreamer_b200 import modules as M
from oracle import rssm_oracle as orc


def relerr(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


class RefActor(torch.nn.Module):
    """Attribute-compatible stand-in for the reference's ActorModel (src/models.py:466-503):
    the B200 imagine_ahead only reads .model, ._min_std, ._mean_scale, .raw_init_std."""

    def __init__(self, Be, S, Hi, A, act="ELU", n_layers=4, min_std=1e-4, init_std=5.0,
                 mean_scale=5.0):
        super().__init__()
        self.model = M.build_mlp(Be + S, Hi, 2 * A, n_layers, act)
        self._min_std, self._init_std, self._mean_scale = min_std, init_std, mean_scale
        self.raw_init_std = torch.log(torch.exp(torch.tensor(init_std)) - 1)
        self.action_distribution = "Gaussian"


def build_gpu_models(d, trans_sd, actor_sd=None, reward_sd=None, value_sd=None, device="cuda"):
    tm = bd.TransitionModel(d["Be"], d["S"], d["A"], d["Hi"], d["E"], d["act"])
    tm.load_state_dict(trans_sd)
    out = types.SimpleNamespace(transition=tm.to(device))
    if actor_sd is not None:
        actor = RefActor(d["Be"], d["S"], d["Hi"], d["A"], d["act"])
        actor.load_state_dict(actor_sd)
        out.actor = actor.to(device)
    for name, sd in (("reward", reward_sd), ("critic", value_sd)):
        if sd is not None:
            m = bd.DenseModel(d["Be"] + d["S"], d["Hi"], activation=d["act"])
            m.load_state_dict(sd)
            setattr(out, name, m.to(device))
    return out


def agent_ns(mods, H):
    return types.SimpleNamespace(transition_model=mods.transition, actor=mods.actor,
                                 planning_horizon=H, latent_distribution="Gaussian")


def freeze(*modules):
    ps = [p for m in modules for p in m.parameters()]
    for p in ps:
        p.requires_grad_(False)
    return ps


def gpu_actor_loss(mods, H, s0, b0, noise, discount=0.995, lambda_=0.95, entropy_weight=1e-5):
    """Behaviour-learning block of Dreamer.train_step (src/dreamer.py:313-363) on the CUDA path.
    Returns loss, intermediates, actor grads."""
    frozen = freeze(mods.transition, mods.reward, mods.critic)
    for p in mods.actor.parameters():
        p.grad = None
    agent = agent_ns(mods, H)
    beliefs, states, (means, stds), entropy = bd.imagine_ahead(agent, s0[None], b0[None], noise)
    reward = mods.reward(beliefs, states)
    value = mods.critic(beliefs, states)
    returns = bd.lambda_return(reward, value, value[-1], discount, lambda_)
    objective = returns
    if entropy_weight != -1:
        objective = objective + entropy_weight * entropy.unsqueeze(-1)
    loss = -objective.mean()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in mods.actor.named_parameters()}
    for p in frozen:
        p.requires_grad_(True)
    inter = dict(beliefs=beliefs, states=states, means=means, stds=stds, entropy=entropy,
                 reward=reward, value=value, returns=returns)
    return loss.detach(), {k: v.detach() for k, v in inter.items()}, grads


def oracle_actor_loss(d, trans_sd, actor_sd, reward_sd, value_sd, s0, b0, ea, ee, es,
                      discount=0.995, lambda_=0.95, entropy_weight=1e-5, dtype=torch.float32):
    c = lambda sd: {k: v.to(dtype) for k, v in sd.items()}
    asd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in actor_sd.items()}
    loss, inter = orc.actor_loss(c(trans_sd), asd, c(reward_sd), c(value_sd), d["act"], 0.1, d["H"],
                                 s0[None].to(dtype), b0[None].to(dtype), ea.to(dtype), ee.to(dtype),
                                 es.to(dtype), discount, lambda_, entropy_weight)
    loss.backward()
    return loss.detach(), {k: v.detach() for k, v in inter.items()}, \
        {k: v.grad for k, v in asd.items()}


def compare_actor_loss(gpu, ref):
    (lg, ig, gg), (lr, ir, gr) = gpu, ref
    # the loss is a mean of the returns (+ 1e-5 entropy): its error is measured against the returns' scale
    # (|loss| itself can be arbitrarily close to zero, e.g. for a single row)
    scale = max(float(lr.abs()), float(ir["returns"].abs().mean()))
    errs = {"loss": float((lg.detach().float().cpu() - lr.detach().float().cpu()).abs()) / max(scale, 1e-30)}
    for k in ("beliefs", "states", "means", "stds", "entropy", "reward", "value", "returns"):
        errs[k] = relerr(ig[k], ir[k])
    errs["actor_grads"] = max(relerr(gg[k], gr[k]) for k in gr)
    return errs


def run_imagine_case(d, seed=0, precision="fp32", small_std=True, oracle_dtype=torch.float32):
    """Synthetic weights/latents/noise -> CUDA path vs oracle."""
    bd.set_precision(precision)
    trans, actor, reward, value = orc.make_models(seed, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    if small_std:  # well-conditioned entropy regime (SURVEY.md hard part 7)
        actor["model.8.bias"][d["A"]:] -= 6.0
    s0, b0 = orc.make_latents(seed, d["N"], d["Be"], d["S"])
    ea, ee, es = orc.make_imagine_noise(seed, d["H"] - 1, d["N"], d["S"], d["A"])
    mods = build_gpu_models(d, trans, actor, reward, value)
    noise = dict(eps_a=ea.cuda(), eps_e=ee.cuda(), eps_s=es.cuda())
    gpu = gpu_actor_loss(mods, d["H"], s0.cuda(), b0.cuda(), noise)
    ref = oracle_actor_loss(d, trans, actor, reward, value, s0, b0, ea, ee, es, dtype=oracle_dtype)
    return dict(errors=compare_actor_loss(gpu, ref), gpu=gpu, ref=ref)


def run_cem_case(d, seed=0, precision="fp32"):
    bd.set_precision(precision)
    trans, _, reward, _ = orc.make_models(seed, d["Be"], d["S"], d["A"], d["Hi"], d["E"])
    g = torch.Generator().manual_seed(seed + 7)
    s0, b0 = orc.make_latents(seed, d["B"], d["Be"], d["S"])
    ea = torch.randn(d["iters"], d["H"], d["B"], d["C"], d["A"], generator=g)
    es = torch.randn(d["iters"], d["H"], d["B"] * d["C"], d["S"], generator=g)
    mods = build_gpu_models(d, trans, reward_sd=reward)
    planner = bd.MPCPlanner(d["A"], d["H"], d["iters"], d["C"], d["K"], mods.transition, mods.reward)
    out = planner(b0.cuda(), s0.cuda(), noise=dict(eps_act=ea.cuda(), eps_s=es.cuda()), trace=True)
    with torch.no_grad():
        ref, trace = orc.cem_plan(trans, reward, d["act"], 0.1, d["A"], d["H"], d["iters"], d["C"],
                                  d["K"], b0, s0, ea, es, return_trace=True)
    topk_gpu = torch.sort(planner.last_trace["topk"].cpu(), dim=2)[0]
    topk_ref = torch.stack([t["topk"] for t in trace])
    ret_err = max(relerr(planner.last_trace["returns"][i], trace[i]["returns"])
                  for i in range(d["iters"]))
    return dict(action_err=relerr(out, ref), elites_equal=bool(torch.equal(topk_gpu, topk_ref)),
                returns_err=ret_err, action=out.cpu(), ref=ref)
