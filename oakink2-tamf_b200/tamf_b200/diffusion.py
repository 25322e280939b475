"""Drop-in for the reference's diffusion sampler surface.

Reference: src/oakink2_tamf/model/diffusion_util.py:5-31 (create_gaussian_diffusion: 1000 steps, cosine, predict x0,
fixed-small variance, no respacing) and model/diffusion/gaussian_diffusion.py:20-62 (schedules), :116-161 (tables),
:412-460 (p_sample), :506-640 (p_sample_loop / _progressive), :642-690 + :766-870 (ddim_sample / ddim_sample_loop),
model/diffusion/respace.py:8-57 (space_timesteps), :60-111 (SpacedDiffusion).  START_X / FIXED_SMALL only (what
create_gaussian_diffusion builds); the launchers use the ancestral sampler over all 1000 steps, the strided and DDIM
samplers are the ones the reference file offers next to it.  PLMS / VB terms / training losses are
out of scope.

Every sampler here is the same device rule  x_{i-1} = c1[i] x0 + c2[i] x_i + sigma[i] eps  fused into the last GEMM of
the denoiser; a sampler differs only in the three tables and the timestep map it installs
(tamf_denoiser_set_sampler)."""
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


def space_timesteps(num_timesteps, section_counts):
    """respace.py:8-57: the set of original timesteps a strided process keeps.  "ddimN" = the one fixed integer stride
    that yields exactly N steps; otherwise per-section counts (list or comma-separated string), each section strided
    with a fractional step and rounded."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                kept = range(0, num_timesteps, stride)
                if len(kept) == want:
                    return set(kept)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(tok) for tok in section_counts.split(",")]
    nsec = len(section_counts)
    base, extra = divmod(num_timesteps, nsec)
    kept, first = [], 0
    for k, count in enumerate(section_counts):
        size = base + (1 if k < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        step = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(first + round(pos))
            pos += step
        first += size
    return set(kept)


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    """gaussian_diffusion.py:45-62."""
    n = num_diffusion_timesteps
    return np.array([min(1 - alpha_bar((i + 1) / n) / alpha_bar(i / n), max_beta) for i in range(n)])


def _x_T(shape, device, seed):
    """x_T = th.randn(*shape) (gaussian_diffusion.py:604): torch's global generator like the reference, or -- when the
    caller passes a chain seed -- a generator seeded with it, so that a seeded chain is reproducible end to end."""
    if seed is None:
        return torch.randn(*shape, device=device)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) & 0x7FFFFFFFFFFFFFFF)
    return torch.randn(*shape, device=device, generator=g)


class GaussianDiffusion:
    """Schedule tables in float64 exactly as GaussianDiffusion.__init__ (gaussian_diffusion.py:116-161), after the
    respacing SpacedDiffusion applies (respace.py:69-83): `use_timesteps` = the original steps to keep (None = all)."""

    def __init__(self, betas, use_timesteps=None):
        betas = np.array(betas, dtype=np.float64)
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.original_num_steps = int(betas.shape[0])
        self.use_timesteps = set(range(self.original_num_steps)) if use_timesteps is None else set(use_timesteps)
        ac0 = np.cumprod(1.0 - betas, axis=0)
        last, nb, tmap = 1.0, [], []
        for i, a in enumerate(ac0):  # respace.py:76-81
            if i in self.use_timesteps:
                nb.append(1 - a / last)
                last = a
                tmap.append(i)
        betas = np.array(nb, dtype=np.float64)
        self.betas = betas
        self.num_timesteps = int(betas.shape[0])
        self.timestep_map = tmap
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)

    # ---- the (c1, c2, sigma) tables of a sampler, fp32 like the reference's per-step lookups (:1275) ----
    def ancestral_rule(self):
        """p_sample with FIXED_SMALL variance (gaussian_diffusion.py:209-229, 305-310, 448-459)."""
        f = lambda a: torch.from_numpy(np.asarray(a)).float()
        sigma = torch.exp(0.5 * f(self.posterior_log_variance_clipped))
        sigma[0] = 0.0  # nonzero_mask
        return f(self.posterior_mean_coef1), f(self.posterior_mean_coef2), sigma

    def ddim_rule(self, eta=0.0):
        """ddim_sample (gaussian_diffusion.py:642-690) with eps re-derived from x0 (:331-335) folded in:
             eps  = (A x_t - x0) / Bm,  A = sqrt(1/abar_t), Bm = sqrt(1/abar_t - 1)
             x'   = sqrt(abar_prev) x0 + k eps + sigma noise,   k = sqrt(1 - abar_prev - sigma^2)
           =>  c1 = sqrt(abar_prev) - k / Bm,   c2 = k A / Bm.   sigma is evaluated in fp32 like the reference."""
        f = lambda a: torch.from_numpy(np.asarray(a)).float()
        ab, abp = f(self.alphas_cumprod), f(self.alphas_cumprod_prev)
        sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
        k = torch.sqrt(1 - abp - sigma ** 2).double()
        A, Bm = f(self.sqrt_recip_alphas_cumprod).double(), f(self.sqrt_recipm1_alphas_cumprod).double()
        c1 = torch.sqrt(abp).double() - k / Bm
        c2 = k * A / Bm
        sigma = sigma.clone()
        sigma[0] = 0.0  # nonzero_mask
        return c1.float(), c2.float(), sigma

    def _install(self, model, kind, eta=0.0):
        """Hands this process's update rule to the device path (no-op when it is already installed)."""
        key = (kind, float(eta), self.original_num_steps, tuple(self.timestep_map))
        if getattr(model, "_rule_key", None) == key and getattr(model, "_handle", None) is not None:
            return  # already installed: the step-by-step API calls this once per step
        c1, c2, sigma = self.ancestral_rule() if kind == "ancestral" else self.ddim_rule(eta)
        model.set_sampler_rule(key, c1, c2, sigma, self.timestep_map, self.original_num_steps)

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
        self._install(model, "ancestral")
        return model.p_sample_step(x, tv, batch, noise=noise)

    def ddim_sample(self, model, x, t, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None, eta=0.0,
                    noise=None):
        """gaussian_diffusion.py:642-690 -> {"sample", "pred_xstart"}.  `noise` (extension) replaces th.randn_like."""
        self._check_supported(clip_denoised, denoised_fn, cond_fn, False, False)
        batch = (model_kwargs or {})["batch"]
        if noise is None:
            noise = torch.randn_like(x)
        tv = int(t[0])
        if not bool((t == tv).all()):
            raise NotImplementedError("ddim_sample: per-row timesteps are not used by ddim_sample_loop")
        self._install(model, "ddim", eta)
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
        img = noise.to(device) if noise is not None else _x_T(shape, device, seed)
        if skip_timesteps and init_image is None:
            init_image = torch.zeros_like(img)
        t_first = self.num_timesteps - skip_timesteps - 1
        if init_image is not None:
            my_t = torch.ones([shape[0]], device=device, dtype=torch.long) * t_first
            img = self.q_sample(init_image, my_t, img)
        img = img.to(torch.float32).contiguous().clone()
        self._install(model, "ancestral")
        if dump_steps is None and not const_noise and not progress:
            return model.p_sample_chain(img, t_first, 0, batch, seed=seed)
        dump = []
        it = range(t_first, -1, -1)
        if progress:
            from tqdm.auto import tqdm
            it = tqdm(it)
        with model.cond_scope(batch, shape[0], shape[3], img.device):  # conditioning once for the whole loop
            for i, t in enumerate(it):
                n = torch.randn_like(img)
                if const_noise:
                    n = n[[0]].repeat(shape[0], 1, 1, 1)
                img = model.p_sample_step(img, t, batch, noise=n)["sample"]
                if dump_steps is not None and i in dump_steps:
                    dump.append(img.clone())
        return dump if dump_steps is not None else img


    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False, seed=None):
        """gaussian_diffusion.py:766-870.  The whole chain is CUDA-graph replays of the fused step (in-kernel Philox
        noise when eta > 0); with progress=True step by step with torch-drawn noise."""
        if dump_steps is not None or const_noise:
            raise NotImplementedError()  # as the reference (:791-794)
        self._check_supported(clip_denoised, denoised_fn, cond_fn, cond_fn_with_grad, randomize_class)
        batch = (model_kwargs or {})["batch"]
        if device is None:
            device = next(model.parameters()).device
        assert isinstance(shape, (tuple, list))
        img = noise.to(device) if noise is not None else _x_T(shape, device, seed)
        if skip_timesteps and init_image is None:
            init_image = torch.zeros_like(img)
        t_first = self.num_timesteps - skip_timesteps - 1
        if init_image is not None:
            my_t = torch.ones([shape[0]], device=device, dtype=torch.long) * t_first
            img = self.q_sample(init_image, my_t, img)
        img = img.to(torch.float32).contiguous().clone()
        self._install(model, "ddim", eta)
        if not progress:
            return model.p_sample_chain(img, t_first, 0, batch, seed=seed)
        from tqdm.auto import tqdm
        with model.cond_scope(batch, shape[0], shape[3], img.device):
            for t in tqdm(range(t_first, -1, -1)):
                img = model.p_sample_step(img, t, batch, noise=torch.randn_like(img))["sample"]
        return img


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:60-111: a process that keeps `use_timesteps` of a base process; the model sees the ORIGINAL
    timesteps through `timestep_map` (_WrappedModel, :114-119) -- on the device that is the gathered timestep-token
    table tamf_denoiser_set_sampler builds."""

    def __init__(self, use_timesteps, betas, **kwargs):
        for k in ("model_mean_type", "model_var_type", "loss_type", "rescale_timesteps"):
            kwargs.pop(k, None)  # START_X / FIXED_SMALL / MSE / no rescaling: the only combination built (diffusion_util.py)
        if kwargs:
            raise TypeError(f"unexpected arguments {sorted(kwargs)}")
        super().__init__(betas, use_timesteps=use_timesteps)


def create_gaussian_diffusion(diffusion_steps, noise_schedule, sigma_small=True, timestep_respacing=""):
    """model/diffusion_util.py:5-31 (`timestep_respacing` is a local there, fixed to ""; exposed here: "ddim50",
    "100", "10,20,30" ... as respace.py:8-57 reads it)."""
    if not sigma_small:
        raise NotImplementedError("FIXED_LARGE variance is not used by the reference launchers")
    betas = get_named_beta_schedule(noise_schedule, diffusion_steps, 1.0)
    if not timestep_respacing:
        timestep_respacing = [diffusion_steps]
    return SpacedDiffusion(use_timesteps=space_timesteps(diffusion_steps, timestep_respacing), betas=betas)
