"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the build
container, where /root/reference exists):

    python -m oracle.make_golden

Each fixture holds the reference modules' state_dicts, the synthetic inputs, the
noise tape, and the reference's own outputs in fp32 (what the product must
match) and fp64 (error budgeting).  Fixtures travel to the GPU box; the
reference does not.
"""
import os

import torch

from oracle import ref_harness as rh
from oracle import rssm_oracle as orc

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _sd(mod, dtype=torch.float32):
    return {k: v.detach().clone().to(dtype) for k, v in mod.state_dict().items()}


def _to(x, dtype):
    if isinstance(x, torch.Tensor):
        return x.to(dtype) if x.is_floating_point() else x
    if isinstance(x, (tuple, list)):
        return type(x)(_to(v, dtype) for v in x)
    if isinstance(x, dict):
        return {k: _to(v, dtype) for k, v in x.items()}
    return x


def imagine_case(name, Be, Hi, S, A, E, N, act, H, small_std, seed=0):
    mods = rh.build_modules(seed, Be, S, A, Hi, E, act)
    if small_std:
        with torch.no_grad():
            mods.actor.model[8].bias[A:] -= 6.0
    s0, b0 = orc.make_latents(seed, N, Be, S)
    ea, ee, es = orc.make_imagine_noise(seed, H - 1, N, S, A)
    fx = dict(dims=dict(Be=Be, Hi=Hi, S=S, A=A, E=E, N=N, H=H, act=act), seed=seed,
              transition=_sd(mods.transition), actor=_sd(mods.actor), reward=_sd(mods.reward),
              critic=_sd(mods.critic), prev_state=s0, prev_belief=b0, eps_a=ea, eps_e=ee, eps_s=es,
              discount=0.995, lambda_=0.95, entropy_weight=1e-5)
    loss, inter, grads = rh.ref_actor_loss(mods, H, s0[None], b0[None], ea, ee, es)
    fx["ref32"] = dict(loss=loss, grads=grads, **inter)
    # fp64 run of the same reference code on the same tape
    m64 = rh.build_modules(seed, Be, S, A, Hi, E, act, torch.float64)
    for dst, src in ((m64.transition, mods.transition), (m64.actor, mods.actor),
                     (m64.reward, mods.reward), (m64.critic, mods.critic)):
        dst.load_state_dict(_sd(src, torch.float64))
    d = torch.float64
    loss, inter, grads = rh.ref_actor_loss(m64, H, s0[None].to(d), b0[None].to(d), ea.to(d),
                                           ee.to(d), es.to(d))
    fx["ref64"] = _to(dict(loss=loss, grads=grads, **inter), torch.float32)
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, "loss", float(fx["ref32"]["loss"]))


def transition_case(name, Be, Hi, S, A, E, B, act, L, seed=1):
    mods = rh.build_modules(seed, Be, S, A, Hi, E, act)
    g = torch.Generator().manual_seed(seed + 10)
    s0, b0 = orc.make_latents(seed, B, Be, S)
    actions = torch.rand(L, B, A, generator=g) * 2 - 1
    ep, eq = torch.randn(L, B, S, generator=g), torch.randn(L, B, S, generator=g)
    emb = torch.randn(L, B, E, generator=g)
    nt = (torch.rand(L, B, 1, generator=g) > 0.15).float()
    fx = dict(dims=dict(Be=Be, Hi=Hi, S=S, A=A, E=E, B=B, L=L, act=act), transition=_sd(mods.transition),
              init_state=s0, init_belief=b0, actions=actions, eps_prior=ep, eps_post=eq,
              embeddings=emb, nonterminals=nt)
    with torch.no_grad():
        r = rh.ref_transition(mods, s0, actions, b0, ep)
        fx["prior_only"] = dict(beliefs=r[0], prior_states=r[1], prior_means=r[2][0],
                                prior_stds=r[2][1])
        r = rh.ref_transition(mods, s0, actions, b0, ep, emb, nt, eq)
        fx["observe"] = dict(beliefs=r[0], prior_states=r[1], prior_means=r[2][0],
                             prior_stds=r[2][1], posterior_states=r[3], posterior_means=r[4][0],
                             posterior_stds=r[4][1])
    # observe backward through the reference: loss = sum of all outputs * fixed cotangents
    for p in mods.transition.parameters():
        p.grad = None
    cot = [torch.randn(L, B, d, generator=g) for d in (Be, S, S, S, S, S, S)]
    s0g, b0g = s0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    embg = emb.clone().requires_grad_(True)
    r = rh.ref_transition(mods, s0g, actions, b0g, ep, embg, nt, eq)
    outs = [r[0], r[1], r[2][0], r[2][1], r[3], r[4][0], r[4][1]]
    sum((o * c).sum() for o, c in zip(outs, cot)).backward()
    fx["observe_bwd"] = dict(cotangents=cot, d_init_state=s0g.grad, d_init_belief=b0g.grad,
                             d_embeddings=embg.grad,
                             grads={k: p.grad.clone() for k, p in mods.transition.named_parameters()})
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name)


def cem_case(name, Be, Hi, S, A, E, B, C, K, H, iters, act="ELU", seed=2):
    mods = rh.build_modules(seed, Be, S, A, Hi, E, act)
    g = torch.Generator().manual_seed(seed + 20)
    s0, b0 = orc.make_latents(seed, B, Be, S)
    ea = torch.randn(iters, H, B, C, A, generator=g)
    es = torch.randn(iters, H, B * C, S, generator=g)
    out = rh.ref_cem(mods, A, H, iters, C, K, b0, s0, ea, es)
    # per-iteration trace from the oracle restatement (pinned to the reference on
    # the final action by tests/test_oracle_vs_reference.py); fp64 for elite margins
    d = torch.float64
    with torch.no_grad():
        o32, trace = orc.cem_plan(_sd(mods.transition), _sd(mods.reward), act, 0.1, A, H, iters, C,
                                  K, b0, s0, ea, es, return_trace=True)
        o64, trace64 = orc.cem_plan(_sd(mods.transition, d), _sd(mods.reward, d), act, 0.1, A, H,
                                    iters, C, K, b0.to(d), s0.to(d), ea.to(d), es.to(d),
                                    return_trace=True)
    assert torch.allclose(o32, out, atol=1e-6)
    fx = dict(dims=dict(Be=Be, Hi=Hi, S=S, A=A, E=E, B=B, C=C, K=K, H=H, iters=iters, act=act),
              transition=_sd(mods.transition), reward=_sd(mods.reward), belief=b0, state=s0,
              eps_act=ea, eps_s=es, ref_action=out, trace32=trace, trace64=_to(trace64, torch.float32),
              ref_action64=o64.float())
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, out.flatten()[:4])


def kl_case(name="kl_loss", L=7, B=9, S=30, seed=4):
    """Reference _kl_loss values and gradients: Planet form, Dreamer without and with balancing;
    free_nats chosen so that some rows sit below the floor and some above."""
    g = torch.Generator().manual_seed(seed)
    mk = lambda: torch.randn(L, B, S, generator=g) * 0.3
    sd = lambda: torch.rand(L, B, S, generator=g) * 0.5 + 0.4
    base = dict(post_mean=mk(), post_std=sd(), prior_mean=mk(), prior_std=sd())
    fx = dict(inputs=base, cases=[])
    for agent, fn, bal in (("planet", 3.0, -1), ("dreamer", 3.0, -1), ("dreamer", 0.05, 0.8),
                           ("dreamer", 3.0, 0.8), ("dreamer", 0.1, 0.3)):
        t = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        loss = rh.ref_kl_loss(agent, (t["post_mean"], t["post_std"]), (t["prior_mean"], t["prior_std"]), fn, bal)
        (loss.sum() * 1.7).backward()
        fx["cases"].append(dict(agent=agent, free_nats=fn, kl_balance=bal, loss=loss.detach().clone(),
                                grads={k: (v.grad.clone() if v.grad is not None else torch.zeros_like(v))
                                       for k, v in t.items()}))
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, [float(c["loss"].sum()) for c in fx["cases"]])


def value_update_case(name="value_update", Be=200, Hi=200, S=30, T=5, N=77, seed=6):
    """Reference critic regression block (src/dreamer.py:369-391): loss + critic gradients, without
    and with the use_discount weighting."""
    mods = rh.build_modules(seed, Be, S, 1, Hi, 8, "ELU")
    g = torch.Generator().manual_seed(seed)
    b = torch.tanh(torch.randn(T, N, Be, generator=g))
    st = 0.5 * torch.randn(T, N, S, generator=g)
    target = torch.randn(T, N, 1, generator=g)
    disc = torch.cumprod(0.995 * torch.round(torch.rand(T, N, 1, generator=g) * 0.6 + 0.45), 0)
    fx = dict(dims=dict(Be=Be, Hi=Hi, S=S, T=T, N=N), critic=_sd(mods.critic), beliefs=b, states=st,
              target=target, discount=disc, cases=[])
    for w in (None, disc):
        loss, grads = rh.ref_value_update(mods, b, st, target, w)
        fx["cases"].append(dict(weighted=w is not None, loss=loss, grads=grads))
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, [float(c["loss"]) for c in fx["cases"]])


def main():
    os.makedirs(OUT, exist_ok=True)
    only = os.environ.get("BD_GOLDEN_ONLY")
    if only == "kl":
        return kl_case()
    if only == "value":
        return value_update_case()
    value_update_case()
    kl_case()
    # c1 (README Pendulum sizes), well-conditioned entropy + default-init entropy
    imagine_case("imagine_c1", 32, 32, 30, 1, 64, 48, "ELU", 15, small_std=True)
    imagine_case("imagine_c1_init", 32, 32, 30, 1, 64, 32, "ELU", 15, small_std=False)
    imagine_case("imagine_odd", 48, 40, 10, 3, 32, 37, "ReLU", 6, small_std=True)
    imagine_case("imagine_tanh", 40, 48, 12, 2, 32, 21, "Tanh", 5, small_std=True)
    imagine_case("imagine_default", 200, 200, 30, 1, 8, 12, "ELU", 15, small_std=True)
    imagine_case("imagine_default_a6", 200, 200, 30, 6, 8, 6, "ELU", 8, small_std=True)
    transition_case("transition_c1", 32, 32, 30, 1, 64, 10, "ELU", 9)
    transition_case("transition_odd", 48, 40, 10, 3, 32, 7, "ReLU", 6)
    cem_case("cem_small", 32, 32, 30, 2, 16, 2, 64, 8, 5, 4)
    cem_case("cem_c3_like", 32, 32, 30, 1, 16, 1, 1000, 100, 4, 3)
    g = torch.Generator().manual_seed(9)
    r, v = torch.randn(14, 33, 1, generator=g), torch.randn(14, 33, 1, generator=g)
    torch.save(dict(reward=r, value=v, discount=0.995, lambda_=0.95,
                    returns=rh.ref_lambda_return(r, v, v[-1], 0.995, 0.95)),
               os.path.join(OUT, "lambda_return.pt"))


if __name__ == "__main__":
    main()
