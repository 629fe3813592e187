"""Oracle restatement vs the frozen reference outputs in tests/golden/ (these
fixtures were produced by oracle/make_golden.py from the unmodified reference).
Runs everywhere (no GPU, no /root/reference needed)."""
import glob
import os

import pytest
import torch

from oracle import rssm_oracle as orc

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relerr(a, b):
    return float((a.detach() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def load(name):
    return torch.load(os.path.join(G, name + ".pt"), weights_only=False)


IMAGINE = sorted(os.path.basename(p)[:-3] for p in glob.glob(os.path.join(G, "imagine_*.pt")))


@pytest.mark.parametrize("name", IMAGINE)
def test_imagine_fixture(name):
    fx = load(name)
    d = fx["dims"]
    asd = {k: v.clone().requires_grad_(True) for k, v in fx["actor"].items()}
    loss, inter = orc.actor_loss(fx["transition"], asd, fx["reward"], fx["critic"], d["act"], 0.1,
                                 d["H"], fx["prev_state"][None], fx["prev_belief"][None],
                                 fx["eps_a"], fx["eps_e"], fx["eps_s"], fx["discount"],
                                 fx["lambda_"], fx["entropy_weight"])
    loss.backward()
    ref = fx["ref32"]
    ent_tol = 5e-2 if name.endswith("_init") else 1e-4
    for k in ("beliefs", "states", "means", "stds", "reward", "value", "returns"):
        assert relerr(inter[k], ref[k]) < 1e-5, k
    assert relerr(inter["entropy"], ref["entropy"]) < ent_tol
    assert relerr(loss, ref["loss"]) < 1e-5
    for k, g in ref["grads"].items():
        assert relerr(asd[k].grad, g) < 2e-4, k


@pytest.mark.parametrize("name", ["transition_c1", "transition_odd"])
def test_transition_fixture(name):
    fx = load(name)
    d = fx["dims"]
    with torch.no_grad():
        o = orc.transition_forward(fx["transition"], d["act"], 0.1, fx["init_state"], fx["actions"],
                                   fx["init_belief"], fx["eps_prior"])
        r = fx["prior_only"]
        assert relerr(o[0], r["beliefs"]) < 1e-5 and relerr(o[1], r["prior_states"]) < 1e-5
        o = orc.transition_forward(fx["transition"], d["act"], 0.1, fx["init_state"], fx["actions"],
                                   fx["init_belief"], fx["eps_prior"], fx["embeddings"],
                                   fx["nonterminals"], fx["eps_post"])
        r = fx["observe"]
        assert relerr(o[0], r["beliefs"]) < 1e-5 and relerr(o[3], r["posterior_states"]) < 1e-5
        assert relerr(o[4][1], r["posterior_stds"]) < 1e-5


@pytest.mark.parametrize("name", ["cem_small", "cem_c3_like"])
def test_cem_fixture(name):
    fx = load(name)
    d = fx["dims"]
    with torch.no_grad():
        o, trace = orc.cem_plan(fx["transition"], fx["reward"], d["act"], 0.1, d["A"], d["H"],
                                d["iters"], d["C"], d["K"], fx["belief"], fx["state"],
                                fx["eps_act"], fx["eps_s"], return_trace=True)
    assert relerr(o, fx["ref_action"]) < 1e-5
    for t, t32 in zip(trace, fx["trace32"]):
        assert torch.equal(t["topk"], t32["topk"])


def test_lambda_return_fixture():
    fx = load("lambda_return")
    o = orc.lambda_return(fx["reward"], fx["value"], fx["value"][-1], fx["discount"], fx["lambda_"])
    assert torch.equal(o, fx["returns"])


def test_kl_loss_fixture():
    """Oracle kl_loss against the reference's Planet/Dreamer._kl_loss values and gradients
    (no balancing, balancing above the free-nats floor, balancing at the floor)."""
    fx = load("kl_loss")
    for c in fx["cases"]:
        t = {k: v.clone().requires_grad_(True) for k, v in fx["inputs"].items()}
        loss = orc.kl_loss((t["post_mean"], t["post_std"]), (t["prior_mean"], t["prior_std"]),
                           torch.full((1,), c["free_nats"]), c["kl_balance"])
        assert loss.shape == c["loss"].shape
        assert relerr(loss, c["loss"]) < 1e-6
        (loss.sum() * 1.7).backward()
        for k, g in c["grads"].items():
            got = t[k].grad if t[k].grad is not None else torch.zeros_like(t[k])
            assert float((got - g).abs().max()) <= 1e-6 * float(g.abs().max()) + 1e-12, (c, k)
