"""Strided / DDIM sampler tables (tamf_b200.diffusion) against outputs of the REFERENCE's own respace.py and
gaussian_diffusion.ddim_sample (tests/golden/spaced_arch_mdm.npz, written by `python -m oracle.make_golden spaced`).
CPU only: the update rule  x' = c1 x0 + c2 x + sigma eps  is evaluated in torch on the reference's own pred_xstart."""
import numpy as np
import pytest
import torch

from tamf_b200 import synth
from tamf_b200.diffusion import (SpacedDiffusion, create_gaussian_diffusion, get_named_beta_schedule,
                                 space_timesteps)


@pytest.mark.parametrize("spec", ["ddim50", "ddim25", "100", "10,20,30"])
def test_space_timesteps_matches_reference(golden, spec):
    g = golden("spaced_arch_mdm.npz")
    assert sorted(space_timesteps(1000, spec)) == g["steps_" + spec.replace(",", "_")].tolist()


def test_space_timesteps_errors():
    with pytest.raises(ValueError, match="integer stride"):
        space_timesteps(1000, "ddim999")
    with pytest.raises(ValueError, match="cannot divide"):
        space_timesteps(10, [20])
    assert space_timesteps(1000, [1000]) == set(range(1000))  # what create_gaussian_diffusion passes by default


def test_spaced_tables_match_reference(golden):
    g = golden("spaced_arch_mdm.npz")
    d = create_gaussian_diffusion(1000, "cosine", timestep_respacing="ddim50")
    assert isinstance(d, SpacedDiffusion) and d.num_timesteps == 50 and d.original_num_steps == 1000
    assert d.timestep_map == g["map50"].tolist()
    for k in ("betas", "alphas_cumprod", "posterior_mean_coef1", "posterior_mean_coef2", "posterior_log_variance_clipped"):
        np.testing.assert_allclose(getattr(d, k), g["d50_" + k], rtol=1e-12, atol=0)
    full = create_gaussian_diffusion(1000, "cosine")
    assert full.num_timesteps == 1000 and full.timestep_map == list(range(1000))


def _x(g):
    B, T = int(g["B"]), int(g["T"])
    return torch.randn(B, 99, 1, T, generator=torch.Generator().manual_seed(5)), (B, 99, 1, T)


@pytest.mark.parametrize("eta", [0.0, 0.5])
def test_ddim_rule_reproduces_reference_ddim_sample(golden, eta):
    """ddim_sample (gaussian_diffusion.py:642-690) folded to three tables: fp32 re-association only."""
    g = golden("spaced_arch_mdm.npz")
    x, shape = _x(g)
    d = SpacedDiffusion(space_timesteps(1000, "ddim50"), get_named_beta_schedule("cosine", 1000))
    c1, c2, sigma = d.ddim_rule(eta)
    assert sigma[0] == 0 and (eta > 0 or float(sigma.abs().max()) == 0.0)
    for i in (49, 20, 1, 0):
        x0 = torch.from_numpy(g[f"x0_i{i}"])
        out = c1[i] * x0 + c2[i] * x + sigma[i] * synth.step_noise(int(g["noise_seed"]), i, shape)
        ref = g[f"ddim_eta{eta}_i{i}"]
        assert np.abs(out.numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), (eta, i)
    assert abs(float(c1[0]) - 1.0) < 1e-6 and abs(float(c2[0])) < 1e-6  # last step returns x0


def test_ancestral_rule_on_strided_process(golden):
    g = golden("spaced_arch_mdm.npz")
    x, shape = _x(g)
    d = SpacedDiffusion(space_timesteps(1000, "ddim50"), get_named_beta_schedule("cosine", 1000))
    c1, c2, sigma = d.ancestral_rule()
    for i in (49, 0):
        x0 = torch.from_numpy(g[f"x0_i{i}"])
        out = c1[i] * x0 + c2[i] * x + sigma[i] * synth.step_noise(int(g["noise_seed"]), i, shape)
        assert np.abs(out.numpy() - g[f"anc_i{i}"]).max() <= 2e-5 * max(1.0, np.abs(g[f"anc_i{i}"]).max())
