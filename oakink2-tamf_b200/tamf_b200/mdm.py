"""Drop-in for `InterationSegmentMDM` (MF-MDM G): same constructor arguments, state_dict keys and
forward(x, timesteps, batch) signature as src/oakink2_tamf/model/interaction_segment_mdm.py:12-174, running on
libtamf_b200 (tcgen05 GEMMs + fused attention / LayerNorm / posterior kernels).

The torch sub-modules below are weight CONTAINERS with the reference's parameter names (so checkpoints saved by
util/state_util.py:22-39 load with strict=False exactly as launch/sample.py:190-196 does); they are never executed."""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .diffusion import GaussianDiffusion, create_gaussian_diffusion


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("weight container: tamf_b200 runs the CUDA library, not torch modules")


def _positional_table(d_model: int, max_len: int = 5000) -> torch.Tensor:
    """PositionalEncoding buffer (interaction_segment_mdm.py:186-193), [max_len,1,d]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe.unsqueeze(0).transpose(0, 1).contiguous()


def build_encoder_container(d, nhead, ff, dropout, activation, num_layers):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=nhead, dim_feedforward=ff, dropout=dropout,
                                           activation=activation)
        return nn.TransformerEncoder(layer, num_layers=num_layers)


def layer_weight_structs(enc: nn.TransformerEncoder, keep: list):
    """Host fp32 pointers of every encoder layer in the order include/tamf_b200.h tamf_layer_weights declares."""
    L = len(enc.layers)
    arr = (_lib.TamfLayerWeights * L)()
    for l, lay in enumerate(enc.layers):
        ts = [lay.self_attn.in_proj_weight, lay.self_attn.in_proj_bias, lay.self_attn.out_proj.weight,
              lay.self_attn.out_proj.bias, lay.linear1.weight, lay.linear1.bias, lay.linear2.weight, lay.linear2.bias,
              lay.norm1.weight, lay.norm1.bias, lay.norm2.weight, lay.norm2.bias]
        for (name, _), t in zip(_lib.TamfLayerWeights._fields_, ts):
            c = t.detach().to("cpu", torch.float32).contiguous()
            keep.append(c)
            setattr(arr[l], name, c.data_ptr())
    return arr


class InterationSegmentMDM(nn.Module):
    def __init__(self, input_dim=99, obj_input_dim=9, hand_shape_dim=10, obj_embed_dim=768, latent_dim=256,
                 ff_size=1024, num_layers=8, num_heads=4, dropout=0.1, activation="gelu", clip_dim=512,
                 clip_version="ViT-B/32", text_encoder: Optional[Callable] = None, diffusion_steps: int = 1000,
                 noise_schedule: str = "cosine", **kargs):
        super().__init__()
        if activation != "gelu":
            raise NotImplementedError("tamf_b200 implements activation='gelu' (config/arch_*.yml)")
        self.latent_dim, self.ff_size, self.num_layers, self.num_heads = latent_dim, ff_size, num_layers, num_heads
        self.dropout, self.activation, self.clip_dim = dropout, activation, clip_dim
        self.input_feats, self.obj_input_feats = input_dim, obj_input_dim
        self.hand_shape_feats, self.obj_embed_feats = hand_shape_dim, obj_embed_dim
        self.cond_mask_prob = kargs.get("cond_mask_prob", 0.0)
        d = latent_dim
        # ---- weight containers, reference names (interaction_segment_mdm.py:46-79) ----
        self.hand_side_process = _Holder()
        self.hand_side_process.register_buffer("rh_embed", torch.zeros(d))
        lh = torch.zeros(d)
        lh[0] = 1.0
        self.hand_side_process.register_buffer("lh_embed", lh)
        self.hand_shape_process = _Holder()
        self.hand_shape_process.shape_embed = nn.Linear(hand_shape_dim, d)
        self.obj_embed_process = _Holder()
        self.obj_embed_process.embedding = nn.Linear(obj_embed_dim, d)
        self.input_process = _Holder()
        self.input_process.poseEmbedding = nn.Linear(input_dim, d)
        self.obj_input_process = _Holder()
        self.obj_input_process.poseEmbedding = nn.Linear(obj_input_dim, d)
        self.input_merge = nn.Sequential(nn.Linear(d * 2, d), nn.SiLU(), nn.Linear(d, d))
        self.sequence_pos_encoder = _Holder()
        self.sequence_pos_encoder.register_buffer("pe", _positional_table(d))
        self.seqTransEncoder = build_encoder_container(d, num_heads, ff_size, dropout, activation, num_layers)
        self.embed_timestep = _Holder()
        self.embed_timestep.time_embed = nn.Sequential(nn.Linear(d, d), nn.SiLU(), nn.Linear(d, d))
        self.embed_timestep.sequence_pos_encoder = self.sequence_pos_encoder  # shared, as in the reference (:72)
        self.embed_text = nn.Linear(clip_dim, d)
        self.output_process = _Holder()
        self.output_process.poseFinal = nn.Linear(d, input_dim)
        self.clip_version = clip_version
        self._text_encoder = text_encoder
        self._clip_model = None
        # ---- runtime state ----
        self._diffusion = create_gaussian_diffusion(diffusion_steps, noise_schedule)
        self._handle = None
        self._handle_dev = None
        self._ws = None
        self._bound = None
        self._cond_scope = None
        self._rule_key = None
        self._keep = []

    # ---- reference surface ----
    def parameters_wo_clip(self):
        return [p for name, p in self.named_parameters() if not name.startswith("clip_model.")]

    def load_state_dict(self, state_dict, strict=True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._drop_handle()
        return res

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._drop_handle()
        return r

    def encode_text(self, raw_text):
        """[B, clip_dim] fp32 CLIP text features (interaction_segment_mdm.py:111-132).  `text_encoder` (constructor)
        overrides; otherwise the vendored `clip` package is loaded like load_and_freeze_clip (:84-97)."""
        device = next(self.parameters()).device
        if self._text_encoder is not None:
            return self._text_encoder(raw_text).to(device=device, dtype=torch.float32)
        if self._clip_model is None:
            try:
                import clip
            except ImportError as e:  # pragma: no cover
                raise RuntimeError("no `clip` package: pass text_encoder=... to InterationSegmentMDM") from e
            m, _ = clip.load(self.clip_version, device="cpu", jit=False)
            clip.model.convert_weights(m)
            self._clip_model = m.eval().to(device)
            self._clip = clip
        texts = self._clip.tokenize(raw_text, context_length=22, truncate=True).to(device)
        texts = torch.cat([texts, torch.zeros([texts.shape[0], 77 - 22], dtype=texts.dtype, device=device)], dim=1)
        with torch.no_grad():
            return self._clip_model.encode_text(texts).float()

    # ---- library plumbing ----
    def _drop_handle(self):
        if self._handle is not None:
            _lib.lib().tamf_denoiser_destroy(self._handle)
        self._handle, self._bound, self._cond_scope, self._ws, self._rule_key = None, None, None, None, None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _ensure_handle(self, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("tamf_b200.InterationSegmentMDM runs on a B200 (model.to('cuda')); no CPU fallback")
        if self._handle is not None and self._handle_dev == device:
            return
        self._drop_handle()
        keep = []
        host = lambda t: keep.append(t.detach().to("cpu", torch.float32).contiguous()) or keep[-1].data_ptr()
        cfg = _lib.TamfCfg(self.input_feats, self.obj_input_feats, self.hand_shape_feats, self.obj_embed_feats,
                           self.latent_dim, self.ff_size, self.num_layers, self.num_heads, self.clip_dim,
                           self._diffusion.num_timesteps)
        w = _lib.TamfGWeights()
        pairs = dict(shape=self.hand_shape_process.shape_embed, objemb=self.obj_embed_process.embedding,
                     pose=self.input_process.poseEmbedding, objtraj=self.obj_input_process.poseEmbedding,
                     merge0=self.input_merge[0], merge2=self.input_merge[2], time0=self.embed_timestep.time_embed[0],
                     time2=self.embed_timestep.time_embed[2], text=self.embed_text, final=self.output_process.poseFinal)
        for k, lin in pairs.items():
            setattr(w, k + "_w", host(lin.weight))
            setattr(w, k + "_b", host(lin.bias))
        pe = self.sequence_pos_encoder.pe[:, 0]
        w.pe, w.pe_rows = host(pe), pe.shape[0]
        layers = layer_weight_structs(self.seqTransEncoder, keep)
        w.layers = C.cast(layers, C.POINTER(_lib.TamfLayerWeights))
        f64 = lambda a: keep.append(np.ascontiguousarray(a, np.float64)) or keep[-1].ctypes.data
        w.posterior_mean_coef1 = f64(self._diffusion.posterior_mean_coef1)
        w.posterior_mean_coef2 = f64(self._diffusion.posterior_mean_coef2)
        w.posterior_log_variance_clipped = f64(self._diffusion.posterior_log_variance_clipped)
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().tamf_denoiser_create(C.byref(cfg), C.byref(w), C.byref(h)), "tamf_denoiser_create")
        self._handle, self._handle_dev = h, device
        del keep, layers

    def _ensure_bound(self, B: int, T: int, device):
        self._ensure_handle(device)
        if self._bound == (B, T):
            return
        L = _lib.lib()
        nbytes = L.tamf_denoiser_workspace_bytes(self._handle, B, T)
        self._ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        base = (self._ws.data_ptr() + 255) & ~255
        with torch.cuda.device(device):
            _lib.check(L.tamf_denoiser_bind(self._handle, B, T, C.c_void_p(base), nbytes), "tamf_denoiser_bind")
        self._bound, self._cond_scope = (B, T), None

    @staticmethod
    def hand_side_ids(hand_side):
        ids = []
        for hs in hand_side:
            if hs == "rh":
                ids.append(0)
            elif hs == "lh":
                ids.append(1)
            else:
                raise ValueError(f"unexpected hand_side: {hs}")  # interaction_segment_mdm.py:284
        return ids

    def set_cond(self, batch: dict, B: int, T: int, device):
        """Conditioning is constant over a reverse chain: the fused sampler entries compute it once per chain (the
        reference recomputes it, CLIP included, at every step -- interaction_segment_mdm.py:141-166).  There is no
        implicit cache: tensor contents cannot be told apart by identity (a freed dict id / allocator address is reused
        by the next item of a per-item loop), so every call recomputes unless the caller has pinned this very batch
        object with `cond_scope(batch)`."""
        self._ensure_bound(B, T, device)
        if self._cond_scope is not None and self._cond_scope[0] is batch and self._cond_scope[1] == (B, T, device):
            return
        ts = [batch["shape"], batch["obj_traj"], batch["obj_embedding"]]
        shape, traj, emb = (_lib.dev_f32(t, device) for t in ts)
        if shape.shape != (B, T, self.hand_shape_feats):
            raise ValueError(f"batch['shape'] must be [B,T,{self.hand_shape_feats}], got {tuple(shape.shape)}")
        nobj = traj.shape[1]
        if traj.shape != (B, nobj, T, self.obj_input_feats) or emb.shape != (B, nobj, self.obj_embed_feats):
            raise ValueError("batch['obj_traj'] / batch['obj_embedding'] have inconsistent shapes")
        side = torch.tensor(self.hand_side_ids(batch["hand_side"]), dtype=torch.int32, device=device)
        text = self.encode_text(batch["text"]).contiguous()
        with torch.cuda.device(device):
            _lib.check(_lib.lib().tamf_denoiser_set_cond(self._handle, _lib.ptr(text), _lib.ptr(side), _lib.ptr(shape),
                                                         _lib.ptr(traj), _lib.ptr(emb), nobj, _lib.stream_ptr(device)),
                       "tamf_denoiser_set_cond")

    @contextlib.contextmanager
    def cond_scope(self, batch: dict, B: int, T: int, device):
        """Pins `batch` as the conditioning of every step issued inside the `with` block: computed once on entry, and
        `set_cond` of this same object (held alive here, so its identity cannot be recycled) is a no-op until exit.  The
        step-by-step sampler loops use it; the caller must not mutate the batch tensors inside the block."""
        self._cond_scope = None
        self.set_cond(batch, B, T, device)
        self._cond_scope = (batch, (B, T, device))
        try:
            yield self
        finally:
            self._cond_scope = None

    def forward(self, x, timesteps, batch):
        """x [B,99,1,T] fp32, timesteps [B] int -> predicted x0 [B,99,1,T] (interaction_segment_mdm.py:134-174)."""
        B, nf, one, T = x.shape
        if nf != self.input_feats or one != 1:
            raise ValueError(f"x must be [B,{self.input_feats},1,T], got {tuple(x.shape)}")
        dev = x.device
        self.set_cond(batch, B, T, dev)
        xc = x.detach().to(torch.float32).contiguous()
        t32 = timesteps.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty_like(xc)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tamf_denoiser_forward(self._handle, _lib.ptr(xc), _lib.ptr(t32), _lib.ptr(out),
                                                        _lib.stream_ptr(dev)), "tamf_denoiser_forward")
        return out

    # ---- fused sampler entry points used by tamf_b200.diffusion ----
    def set_sampler_rule(self, key, c1, c2, sigma, timestep_map, original_num_steps):
        """Installs x_{i-1} = c1[i] x0 + c2[i] x_i + sigma[i] eps with model timesteps timestep_map[i]
        (tamf_denoiser_set_sampler); `key` identifies the rule so that repeated calls are free."""
        dev = next(self.parameters()).device
        self._ensure_handle(dev)
        if self._rule_key == key:
            return
        if original_num_steps != self._diffusion.num_timesteps:
            raise ValueError(f"diffusion process has {original_num_steps} base steps, the model was built for "
                             f"{self._diffusion.num_timesteps}")
        f32 = lambda t: torch.as_tensor(t).detach().to("cpu", torch.float32).contiguous()
        c1, c2, sigma = f32(c1), f32(c2), f32(sigma)
        tmap = torch.tensor(list(timestep_map), dtype=torch.int32)
        K = int(tmap.numel())
        if not (c1.numel() == c2.numel() == sigma.numel() == K):
            raise ValueError("sampler tables and timestep_map must have the same length")
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tamf_denoiser_set_sampler(self._handle, K, _lib.ptr(c1), _lib.ptr(c2), _lib.ptr(sigma),
                                                            _lib.ptr(tmap)), "tamf_denoiser_set_sampler")
        self._rule_key = key

    def p_sample_step(self, x, t: int, batch, noise=None, seed: int = 0):
        B, _, _, T = x.shape
        dev = x.device
        self.set_cond(batch, B, T, dev)
        x_io = x.detach().to(torch.float32).contiguous().clone()
        x0 = torch.empty_like(x_io)
        n = None if noise is None else noise.to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tamf_p_sample_step(self._handle, _lib.ptr(x_io), int(t), _lib.ptr(n), int(seed),
                                                     _lib.ptr(x0), _lib.stream_ptr(dev)), "tamf_p_sample_step")
        return {"sample": x_io, "pred_xstart": x0}

    def p_sample_chain(self, x_T, t_start: int, t_end: int, batch, seed=None):
        """x_T [B,99,1,T] is consumed (updated in place) and returned as the sample at t_end."""
        B, _, _, T = x_T.shape
        dev = x_T.device
        self.set_cond(batch, B, T, dev)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tamf_p_sample_chain(self._handle, _lib.ptr(x_T), int(t_start), int(t_end), int(seed),
                                                      _lib.stream_ptr(dev)), "tamf_p_sample_chain")
        return x_T

    def sample_host(self, batch: dict, seed: int = 0, x_T: Optional[torch.Tensor] = None) -> torch.Tensor:
        """End-to-end p_sample_loop on HOST buffers (tamf_p_sample_loop_host): conditioning tensors are read from
        (pinned) host memory, the finished sample [B,99,1,T] is written back to host memory."""
        dev = next(self.parameters()).device
        B, T = batch["shape"].shape[0], batch["shape"].shape[1]
        self._ensure_bound(B, T, dev)
        host = lambda t: t.detach().to("cpu", torch.float32).contiguous()
        shape, traj, emb = host(batch["shape"]), host(batch["obj_traj"]), host(batch["obj_embedding"])
        text = self.encode_text(batch["text"]).to("cpu").contiguous()
        side = torch.tensor(self.hand_side_ids(batch["hand_side"]), dtype=torch.int32)
        out = torch.empty((B, self.input_feats, 1, T), dtype=torch.float32).pin_memory()
        xT = None if x_T is None else host(x_T)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tamf_p_sample_loop_host(
                self._handle, _lib.ptr(text), _lib.ptr(side), _lib.ptr(shape), _lib.ptr(traj), _lib.ptr(emb),
                traj.shape[1], _lib.ptr(xT), int(seed), _lib.ptr(out), _lib.stream_ptr(dev)),
                "tamf_p_sample_loop_host")
        self._cond_scope = None  # the host entry installed its own conditioning
        return out
