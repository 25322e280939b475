"""Drop-in for the reference's DDPM sampler surface.

Reference: src/oakink2_tamf/model/diffusion_util.py:5-31 (create_gaussian_diffusion: 1000 steps, cosine, predict x0,
fixed-small variance, no respacing) and model/diffusion/gaussian_diffusion.py:20-62 (schedules), :116-161 (tables),
:412-460 (p_sample), :506-640 (p_sample_loop / _progressive).  Only the ancestral START_X / FIXED_SMALL path the
launchers use is implemented; DDIM/PLMS/VB branches are out of scope (never called)."""
from __future__ import annotations

import math

import numpy as np
import torch


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.0):
    """gaussian_diffusion.py:20-43."""
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    if schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """gaussian_diffusion.py:45-62."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


class GaussianDiffusion:
    """Schedule tables in float64 exactly as GaussianDiffusion.__init__ (gaussian_diffusion.py:116-161), after the
    identity respacing SpacedDiffusion applies (respace.py:69-83)."""

    def __init__(self, betas):
        betas = np.array(betas, dtype=np.float64)
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        ac0 = np.cumprod(1.0 - betas, axis=0)
        last, nb = 1.0, []
        for a in ac0:  # respace.py:76-81 with use_timesteps = all
            nb.append(1 - a / last)
            last = a
        betas = np.array(nb, dtype=np.float64)
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        self.timestep_map = list(range(self.num_timesteps))
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)

    # ---- q(x_t | x_0), gaussian_diffusion.py:188-207 ----
    def q_sample(self, x_start, t, noise=None):
        if noise is None:
            noise = torch.randn_like(x_start)
        a = torch.from_numpy(self.sqrt_alphas_cumprod).to(t.device)[t].float().view(-1, *([1] * (x_start.ndim - 1)))
        b = torch.from_numpy(self.sqrt_one_minus_alphas_cumprod).to(t.device)[t].float().view(
            -1, *([1] * (x_start.ndim - 1)))
        return a * x_start + b * noise

    def _check_supported(self, clip_denoised, denoised_fn, cond_fn, cond_fn_with_grad, randomize_class):
        if clip_denoised or denoised_fn is not None or cond_fn is not None or cond_fn_with_grad or randomize_class:
            raise NotImplementedError(
                "tamf_b200 implements the sampler configuration the reference launchers use "
                "(clip_denoised=False, no denoised_fn / cond_fn; launch/sample.py:218-229)")

    def p_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                 const_noise=False, noise=None):
        """gaussian_diffusion.py:412-460 -> {"sample", "pred_xstart"}; t [B] (all equal, as p_sample_loop builds it)."""
        self._check_supported(clip_denoised, denoised_fn, cond_fn, False, False)
        batch = (model_kwargs or {})["batch"]
        if noise is None:
            noise = torch.randn_like(x)
        if const_noise:
            noise = noise[[0]].repeat(x.shape[0], 1, 1, 1)
        tv = int(t[0])
        if not bool((t == tv).all()):
            raise NotImplementedError("p_sample: per-row timesteps are not used by p_sample_loop")
        return model.p_sample_step(x, tv, batch, noise=noise)

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                      randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False, seed=None):
        """gaussian_diffusion.py:506-571.  With no per-step hooks (dump_steps / const_noise) the whole chain runs as
        CUDA-graph replays with in-kernel Philox noise; otherwise step by step with torch-drawn noise."""
        self._check_supported(clip_denoised, denoised_fn, cond_fn, cond_fn_with_grad, randomize_class)
        batch = (model_kwargs or {})["batch"]
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise.to(device) if noise is not None else torch.randn(*shape, device=device)
        if skip_timesteps and init_image is None:
            init_image = torch.zeros_like(img)
        t_first = self.num_timesteps - skip_timesteps - 1
        if init_image is not None:
            my_t = torch.ones([shape[0]], device=device, dtype=torch.long) * t_first
            img = self.q_sample(init_image, my_t, img)
        img = img.to(torch.float32).contiguous().clone()
        if dump_steps is None and not const_noise and not progress:
            return model.p_sample_chain(img, t_first, 0, batch, seed=seed)
        dump = []
        it = range(t_first, -1, -1)
        if progress:
            from tqdm.auto import tqdm
            it = tqdm(it)
        for i, t in enumerate(it):
            n = torch.randn_like(img)
            if const_noise:
                n = n[[0]].repeat(shape[0], 1, 1, 1)
            img = model.p_sample_step(img, t, batch, noise=n)["sample"]
            if dump_steps is not None and i in dump_steps:
                dump.append(img.clone())
        return dump if dump_steps is not None else img


SpacedDiffusion = GaussianDiffusion


def create_gaussian_diffusion(diffusion_steps, noise_schedule, sigma_small=True):
    """model/diffusion_util.py:5-31."""
    if not sigma_small:
        raise NotImplementedError("FIXED_LARGE variance is not used by the reference launchers")
    return GaussianDiffusion(get_named_beta_schedule(noise_schedule, diffusion_steps, 1.0))
