"""The learner-side restatements (oracle.actor_reference, oracle.critic_reference,
oracle.discounted_returns_reference) and the fused kernels against outputs of the REAL reference's
`marlnav.models` (tests/golden/ref_models.npz, produced by make_golden.py:models_case from
/root/reference/marlnav/models.py:14-56,113-115,131-148 on fixed weights and inputs, the standard
normal draws of MultivariateNormal injected)."""
import numpy as np
import pytest
import torch

import golden_replay as gr


def _case(z, tag):
    t = lambda k: torch.from_numpy(np.ascontiguousarray(z[f"{tag}_{k}"]))
    actor = {k[len(tag) + 7:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(f"{tag}_actor.")}
    critic = {k[len(tag) + 8:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith(f"{tag}_critic.")}
    return t, actor, critic


@pytest.mark.parametrize("tag", ["a", "b"])
def test_actor_restatement_matches_reference_models(oracle, tag):
    """oracle.actor_reference == Actor.forward -> dist.sample() -> dist.log_prob() of the reference
    (same torch ops; the float32 GEMM's summation order differs between hosts, thread counts and
    even buffer alignments -- one run in this container missed a 1e-6 bound -- hence the parity
    contract's 1e-5, not bits)."""
    _, z = gr.load("ref_models")
    t, actor, _ = _case(z, tag)
    obs = t("obs")
    act, lp, mu, var = oracle.actor_reference(obs, actor, t("eps"))
    np.testing.assert_allclose(mu.numpy(), z[f"{tag}_mu"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(var.numpy(), z[f"{tag}_var"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(act.numpy(), z[f"{tag}_actions"], rtol=1e-5, atol=4e-6)
    np.testing.assert_allclose(lp.numpy(), z[f"{tag}_log_probs"], rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_critic_restatement_matches_reference_models(oracle, tag):
    _, z = gr.load("ref_models")
    t, _, critic = _case(z, tag)
    got = oracle.critic_reference(t("obs"), critic)
    np.testing.assert_allclose(got.numpy(), z[f"{tag}_values"], rtol=1e-5, atol=4e-6)


def test_returns_restatement_matches_reference_process_rewards(oracle):
    """MAPPO._process_rewards (models.py:131-148): backward scan bit-exact in float64, and its
    std/mean normalisation."""
    _, z = gr.load("ref_models")
    rew, done = torch.from_numpy(z["ret_rewards"].copy()), torch.from_numpy(z["ret_done"].copy())
    want = torch.from_numpy(z["ret_returns"].copy())
    got = oracle.discounted_returns_reference(rew, done, float(z["ret_gamma"]))
    assert got.dtype == torch.float64 and torch.equal(got, want)
    std, mean = torch.std_mean(got.reshape(-1))
    np.testing.assert_allclose(((got - mean) / (std + 1e-12)).numpy(), z["ret_normalized"], rtol=1e-12, atol=1e-12)
    assert abs(float(mean) - float(z["ret_mean"])) < 1e-12


# ----------------------------------------------------------------------------- GPU: kernels vs the reference

@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_fused_actor_matches_reference_models(tag):
    """marlnav_actor_sample_f32 against the reference's own Actor / MultivariateNormal outputs,
    1e-5 (float32 dot products in a different association than torch's GEMM)."""
    import marlnav_b200 as mb
    _, z = gr.load("ref_models")
    t, actor, _ = _case(z, tag)
    obs = t("obs")
    act, lp, mu, var = mb.FusedActor(actor, seed=1).act(obs.cuda(), eps=t("eps"), want_moments=True)
    np.testing.assert_allclose(mu.cpu().numpy(), z[f"{tag}_mu"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(var.cpu().numpy(), z[f"{tag}_var"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(act.cpu().numpy(), z[f"{tag}_actions"], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), z[f"{tag}_log_probs"], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_fused_critic_matches_reference_models(tag):
    import marlnav_b200 as mb
    _, z = gr.load("ref_models")
    t, _, critic = _case(z, tag)
    got = mb.FusedCritic(critic)(t("obs").cuda()).cpu()
    np.testing.assert_allclose(got.numpy(), z[f"{tag}_values"], rtol=2e-5, atol=1e-5)


@pytest.mark.gpu
def test_discounted_returns_match_reference_process_rewards():
    import marlnav_b200 as mb
    _, z = gr.load("ref_models")
    rew, done = torch.from_numpy(z["ret_rewards"].copy()), torch.from_numpy(z["ret_done"].copy())
    got = mb.discounted_returns(rew.cuda(), done.cuda(), float(z["ret_gamma"])).cpu()
    assert torch.equal(got, torch.from_numpy(z["ret_returns"].copy()))
    norm = mb.discounted_returns(rew.cuda(), done.cuda(), float(z["ret_gamma"]), normalize=True).cpu()
    np.testing.assert_allclose(norm.numpy(), z["ret_normalized"], rtol=1e-9, atol=1e-9)
