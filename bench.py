#!/usr/bin/env python
"""bench.py -- sampled motion sequences/sec, full 1000-step reverse chain, MF-MDM G arch_mdm_l (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (libtamf_b200.so, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference on the host cores (oracle/_ref, else the oracle port), rank 0 only
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1: one rank per GPU, batch-sharded, NCCL gather

A "step" is one pass of the hot path over one batch: conditioning (once per sample) + x_T ~ N(0,I) + the 1000
p_sample steps for B=64 synthetic sequences per GPU (BASELINE.json configs[1]).  One JSON line on stdout (rank 0).

  value     whole-job sequences/s with every input resident in HBM when the timed region starts
  e2e       the same through the reference-facing call (InterationSegmentMDM.sample_host -> tamf_p_sample_loop_host)
            with HOST buffers: H2D of the conditioning and D2H of the samples inside the timed region
  roofline  the dominant kernel of the step, measured live INSIDE the captured step graph (globaltimer stamps at every
            kernel's dependency-wait end and last CTA exit, tamf_denoiser_profile_graph): algorithmic FLOPs per launch /
            in-graph duration vs the measured sustained bf16 peak (MEASURED_PEAKS.json); `isolated` repeats it with CUDA
            events around eager launches against the burst peak; `step` is the whole evaluation
    --sequences S   BASELINE.json configs[3]: S sequences batch-sharded over the ranks (strong scaling)
    --config refine BASELINE.json configs[2]: the MF-MDM R forward
  cpu_baseline  the oracle port of the reference's PyTorch algorithm on the host cores (bounded sample, extrapolated)

CLIP `encode_text` is library code that stays in PyTorch and needs downloaded weights (absent offline): both arms
take the [B,512] text feature as an input (synthetic), i.e. the reference arm is NOT charged the per-step CLIP
evaluation the real reference performs -- the conservative choice for the speed-up.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oakink2-tamf_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "sampled motion sequences/sec, full reverse chain, arch_mdm_l"
UNIT = "sequences/s"
ARCH = "arch_mdm_l"
T_FRAMES, NOBJ, DIFF_STEPS = 160, 2, 1000


def flops_per_seq_step(cfg, T=T_FRAMES):
    """SURVEY.md 8d closed form: algorithmic FLOPs of one denoiser evaluation for one sequence."""
    d, ff, L = cfg["latent_dim"], cfg["ff_size"], cfg["num_layers"]
    S = T + 5
    return L * (8 * S * d * d + 4 * S * S * d + 4 * S * d * ff) + 4 * T * 99 * d + 4 * T * d * d


def kernel_classes(cfg, B, T=T_FRAMES, chain=True, stack=False):
    """(name, algorithmic FLOPs per launch) in the launch order tamf_denoiser_profile_step reports.  chain=True: the
    layer-kernel form of the encoder (csrc/layer_chain.cuh): in_proj of layer 0, then per layer attention | out_proj+LN1 ->
    linear1+GELU -> linear2+LN2 -> in_proj of the next layer in ONE kernel; chain=False: the five-kernel layer (TAMF_CHAIN=0)."""
    d, ff, L = cfg["latent_dim"], cfg["ff_size"], cfg["num_layers"]
    S = T + 5
    M, Mf = B * S, B * T
    out = [("prep", 0), ("embed_a", 2 * Mf * 99 * d), ("embed_b", 2 * Mf * d * d)]
    f_in, f_att, f_out, f_l1, f_l2 = 2 * M * 3 * d * d, 4 * B * S * S * d, 2 * M * d * d, 2 * M * d * ff, 2 * M * ff * d
    if stack:  # TAMF_CHAIN=2: one persistent attention launch + one persistent layer-kernel launch for all layers
        out += [("in_proj0", f_in), ("attention_all_layers", L * f_att),
                ("layers_ln1_l1_ln2_inproj_all", L * (f_out + f_l1 + f_l2) + (L - 1) * f_in)]
    elif chain:
        out.append(("in_proj0", f_in))
        for l in range(L):
            out += [("attention", f_att), ("layer_ln1_l1_ln2_inproj", f_out + f_l1 + f_l2 + (f_in if l + 1 < L else 0))]
    else:
        for _ in range(L):
            out += [("in_proj", f_in), ("attention", f_att), ("out_proj_ln", f_out), ("linear1_gelu", f_l1),
                    ("linear2_ln", f_l2)]
    out.append(("final_posterior", 2 * Mf * d * 99))
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm=j["hbm_gbs"], burst=j["bf16_tflops"], sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, burst=1590.0, sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])), mx.append(float(c[1])), pw.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def host_threads():
    """All the host cores the reference arm may use.  torch.distributed.run exports OMP_NUM_THREADS=1 to every rank when
    nproc > 1, which would leave the CPU arm single-threaded: the count is taken from the machine, not the environment."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


_CPU_STATE = {}


_KIND_NOTE = {
    "reference": "the reference's own InterationSegmentMDM + GaussianDiffusion.p_sample from oracle/_ref, CLIP text tower "
                 "(random init) evaluated every step as the reference does",
    "port": "oracle port of the reference (oracle/tamf_oracle.py), CLIP text tower excluded (text feature given)",
}
REF_DIR = os.path.join(ROOT, "oracle", "_ref")  # private copy of the reference's modules (oracle/make_ref.py), git-ignored


def reference_kind():
    """"reference": oracle/_ref holds the reference's own modules (built by oracle/make_ref.py where /root/reference
    exists; travels to the GPU box) -> the CPU arm runs them unmodified.  "port": only the oracle restatement is there."""
    return "reference" if os.path.isdir(os.path.join(REF_DIR, "src", "oakink2_tamf")) else "port"


def _reference_sample(B, n_diff):
    """The reference itself on the host cores: its InterationSegmentMDM (arch_mdm_l, the bench's random-init weights) and
    its GaussianDiffusion.p_sample (gaussian_diffusion.py:412-460), called as launch/sample.py:216-228 does -- including
    the CLIP text tower the reference evaluates at every step (random-init ViT-B/32 from oracle/ref_shims.py: the trained
    weights are a network download)."""
    import torch
    if "ref" not in _CPU_STATE:
        os.environ["TAMF_REFERENCE_ROOT"] = REF_DIR
        import warnings

        from oracle import ref_shims
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ns = ref_shims.install()
        _CPU_STATE["ref"] = ns
    ns = _CPU_STATE["ref"]
    from tamf_b200 import synth
    key = ("ref", B)
    if key not in _CPU_STATE:
        import warnings
        cfg = synth.ARCH[ARCH]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model = ns.mdm.InterationSegmentMDM(**cfg)
        missing, unexpected = model.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
        assert not unexpected and all(k.startswith("clip_model") for k in missing), (missing[:4], unexpected[:4])
        model.eval()
        diffusion = ns.diffusion_util.create_gaussian_diffusion(diffusion_steps=DIFF_STEPS, noise_schedule="cosine")
        _CPU_STATE[key] = (model, diffusion, synth.make_batch(B, T_FRAMES, nobj=NOBJ, seed=0))
    model, diffusion, batch = _CPU_STATE[key]
    x = torch.randn(B, 99, 1, T_FRAMES, generator=torch.Generator().manual_seed(0))
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n_diff):
            t = torch.full((B,), DIFF_STEPS - 1 - i, dtype=torch.long)
            x = diffusion.p_sample(model, x, t, clip_denoised=False, model_kwargs={"batch": batch})["sample"]
    return (time.perf_counter() - t0) / n_diff


def cpu_reference_sample(B, n_diff):
    """The reference's per-step algorithm (InterationSegmentMDM.forward + p_sample) on the host cores: `n_diff` diffusion
    steps at batch B, arch_mdm_l, torch intra-op threads = every host core.  Runs the reference's own modules from
    oracle/_ref when they are there (reference_kind()), else the oracle port.  Returns seconds per diffusion step."""
    import torch
    if torch.get_num_threads() != host_threads():
        torch.set_num_threads(host_threads())
    if reference_kind() == "reference":
        return _reference_sample(B, n_diff)

    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    if B not in _CPU_STATE:
        cfg = synth.ARCH[ARCH]
        batch = synth.make_batch(B, T_FRAMES, nobj=NOBJ, seed=0)
        _CPU_STATE[B] = (cfg, synth.g_state_dict(cfg, seed=0), batch, synth.text_features(batch["text"]),
                         orc.diffusion_tables(DIFF_STEPS))
    cfg, sd, batch, text, tab = _CPU_STATE[B]
    x = torch.randn(B, 99, 1, T_FRAMES, generator=torch.Generator().manual_seed(0))
    t0 = time.perf_counter()
    with torch.no_grad():
        for i in range(n_diff):
            t = DIFF_STEPS - 1 - i
            x0 = orc.g_forward(sd, cfg, x, torch.full((B,), t, dtype=torch.long), batch, text)
            x = orc.p_sample_update(tab, x, x0, t, torch.randn_like(x))
    return (time.perf_counter() - t0) / n_diff


def workload_config(B, world, chain_steps=DIFF_STEPS):
    """The `config` object of BOTH arms (identical dicts: the driver compares them)."""
    return {"workload": f"MF-MDM G {ARCH}, batch {B} synthetic sequences per GPU, T={T_FRAMES}, nobj={NOBJ}, "
                        f"full {chain_steps}-step reverse chain (BASELINE.json configs[1])",
            "global_batch": world * B, "sequences_per_gpu": B, "frames": T_FRAMES, "objects": NOBJ,
            "diffusion_steps": chain_steps, "weights": "random init",
            "text_features": "synthetic prompts; trained CLIP weights are not available offline (GPU arm: hash features, "
                             "evaluated once per batch; reference arm: whatever oracle/_ref or the port does, see its "
                             "cpu_baseline.sample)"}


def run_reference(args):
    """`--impl reference`: the reference on every host core -- its own modules from oracle/_ref when the snapshot carries
    them (oracle/make_ref.py, built where /root/reference exists), else the oracle port.  Rank 0 only; other ranks exit 0
    without work.  Each timed step is a bounded sample of the workload -- `n_diff` of the 1000 diffusion steps at the full
    batch, extrapolated linearly -- sized from a first measured step so that the whole --steps/--warmup run stays within
    --ref-budget-s."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    os.environ.pop("OMP_NUM_THREADS", None)
    B = args.batch
    first = cpu_reference_sample(B, 1)  # also the first warm-up step
    per_call = max(args.ref_budget_s / max(1, args.steps + args.warmup), first)
    n_diff = int(max(1, min(args.ref_diff_steps, per_call / first)))
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_sample(B, 1)
    per = [cpu_reference_sample(B, n_diff) for _ in range(args.steps)]
    sec_per_diff = sum(per) / len(per)
    ms_per_step = sec_per_diff * DIFF_STEPS * 1e3
    value = B / (sec_per_diff * DIFF_STEPS)
    cores = host_threads()
    kind = reference_kind()
    sample = (f"{n_diff} of {DIFF_STEPS} diffusion steps at B={B} per timed step (x{args.steps} steps, "
              f"{sec_per_diff:.3f} s per diffusion step), extrapolated linearly to the full chain; " + _KIND_NOTE[kind])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "notes": "reference arm on the host cores, rank 0 only (kind: see cpu_baseline); ms_per_step is the linear "
                 "extrapolation of the bounded sample to the full 1000-step chain",
    }
    print(json.dumps(line), flush=True)


def _dist_setup():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; tamf_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, local, dev


def in_graph_breakdown(model, L, x, classes, batch, n_steps=8, warm_steps=500):
    """Per-kernel figures INSIDE the captured step (tamf_denoiser_profile_graph: globaltimer stamps at every kernel's
    entry / end of its dependency wait / exit, the same graph and programmatic launches as the chain).  Returns
    {class: {"us": busy time ready->exit summed over its launches, "launches", "flops"}}, the step period and the share of
    the period no kernel of ours was past its dependency wait (launch gaps + dependency waits)."""
    import ctypes as C
    from tamf_b200 import _lib
    cap = 128
    en, rd, ex = (C.c_double * cap)(), (C.c_double * cap)(), (C.c_double * cap)()
    n_out, step_us = C.c_int(0), C.c_double(0)
    # The board throttles under its power cap within tens of milliseconds: queue half a chain ahead (asynchronous), so
    # that the stamped replays run at the SUSTAINED clock of the timed chain and not at the boost clock of an idle GPU.
    model.p_sample_chain(x, DIFF_STEPS - 1, DIFF_STEPS - warm_steps, batch, seed=5)
    _lib.check(L.tamf_denoiser_profile_graph(model._handle, _lib.ptr(x), 700, n_steps, 11, en, rd, ex, cap, C.byref(n_out),
                                             C.byref(step_us), _lib.stream_ptr(x.device)), "profile_graph")
    n = n_out.value
    assert n == len(classes), (n, len(classes))
    per = {}
    busy_until, covered = 0.0, 0.0
    for k, (name, fl) in enumerate(classes):
        c = per.setdefault(name, {"us": 0.0, "launches": 0, "flops": 0})
        c["us"] += ex[k] - rd[k]
        c["launches"] += 1
        c["flops"] += fl
        lo, hi = max(rd[k], busy_until), ex[k]
        if hi > lo:
            covered += hi - lo
            busy_until = hi
    return per, step_us.value, 1.0 - covered / max(step_us.value, 1e-9)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import tamf_b200
    from tamf_b200 import _lib, shard, synth

    world, rank, local, dev = _dist_setup()
    L = _lib.lib()
    cfg = synth.ARCH[ARCH]
    B, T = args.batch, T_FRAMES
    strong = args.sequences > 0
    model = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    model.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    model = model.eval().to(dev)
    tamf_b200.create_gaussian_diffusion(DIFF_STEPS, "cosine")._install(model, "ancestral")
    host_batch = synth.make_batch(B, T, nobj=NOBJ, seed=100 + rank)
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    dev_batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    x = torch.empty(B, 99, 1, T, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)
    t_end = DIFF_STEPS - args.chain_steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if strong:
        # configs[3]: a FIXED set of sequences, contiguous index ranges per rank (launch/sample.py:198-199), chains of B
        # on ONE x buffer (one captured graph: timestep counter and seed live in device memory), samples copied into the
        # rank's output block, one NCCL all_gather at the end
        mine = shard.shard_range(args.sequences, rank, world)
        chains = list(shard.batches(mine, B))
        if any(len(c) != B for c in chains):
            raise SystemExit("--sequences must be a multiple of world_size * batch for the strong-scaling run")
        out = torch.empty(len(mine), T, 99, device=dev)
        gathered = None

        def device_step(seed):
            flush.zero_()
            for i, ids in enumerate(chains):
                model.set_cond(dev_batch, B, T, dev)  # conditioning: once per chain
                _lib.check(L.tamf_philox_normal(_lib.ptr(x), x.numel(), seed * 100003 + ids.start, DIFF_STEPS,
                                                _lib.stream_ptr(dev)), "x_T")
                model.p_sample_chain(x, DIFF_STEPS - 1, t_end, dev_batch, seed=seed * 100003 + ids.start)
                out[ids.start - mine.start: ids.stop - mine.start] = x.permute(0, 3, 1, 2).squeeze(3)  # extract_sample.py:32
            return shard.gather_samples(out, args.sequences)
        seqs_per_step = args.sequences
    else:
        gathered = torch.empty(world * B, 99, 1, T, device=dev) if world > 1 else None

        def device_step(seed):
            flush.zero_()  # L2 flush between timed iterations (the step's own working set, 180 MB, also exceeds L2)
            model.set_cond(dev_batch, B, T, dev)  # conditioning: once per sample batch
            _lib.check(L.tamf_philox_normal(_lib.ptr(x), x.numel(), seed, DIFF_STEPS, _lib.stream_ptr(dev)), "x_T")
            model.p_sample_chain(x, DIFF_STEPS - 1, t_end, dev_batch, seed=seed)
            if world > 1:
                dist.all_gather_into_tensor(gathered, x)  # the one collective of the path: final gather of the samples
        seqs_per_step = world * B

    # ---------------- device-resident throughput ----------------
    if strong:  # one short chain warms the graph / workspace; a warm-up "step" would be the whole job
        model.set_cond(dev_batch, B, T, dev)
        model.p_sample_chain(x.normal_(), DIFF_STEPS - 1, DIFF_STEPS - 10, dev_batch, seed=1)
    else:
        for i in range(args.warmup):
            device_step(1000 + i)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = L.tamf_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        device_step(2000 + i)
    e1.record(stream)
    barrier()
    launches = L.tamf_kernel_launch_count() - n0
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    assert torch.isfinite(x).all(), "non-finite samples"
    value = seqs_per_step * args.steps / (ms_total * 1e-3)

    # ---------------- end to end through the host-buffer API ----------------
    h2d = sum(pinned[k].numel() * 4 for k in ("shape", "obj_traj", "obj_embedding")) + B * 512 * 4 + B * 4
    d2h = B * 99 * T * 4
    e2e = None
    if not strong:
        e2e_steps = max(1, args.steps)
        model.sample_host(pinned, seed=1)  # warm-up (graph already captured; the workspace binding is reused)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            out_h = model.sample_host(pinned, seed=3000 + i)
        torch.cuda.synchronize(dev)
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert torch.isfinite(out_h).all()
        e2e = {"value": world * B * e2e_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "path": "InterationSegmentMDM.sample_host -> tamf_p_sample_loop_host: pinned host conditioning in, host samples out"}

    # ---------------- roofline: in-graph per-kernel figures (live) + isolated launches ----------------
    roof = None
    if rank == 0:
        import ctypes as C
        pk = peaks()
        model.set_cond(dev_batch, B, T, dev)
        classes = kernel_classes(cfg, B)
        ms_buf, n_out = (C.c_float * 64)(), C.c_int(0)
        _lib.check(L.tamf_denoiser_profile_step(model._handle, _lib.ptr(x), 500, 7, ms_buf, 64, C.byref(n_out),
                                                _lib.stream_ptr(dev)), "profile_step")
        if n_out.value == 7:  # TAMF_CHAIN=2: the stack form
            classes = kernel_classes(cfg, B, stack=True)
        elif n_out.value != len(classes):  # TAMF_CHAIN=0: the five-kernel layer of round 1
            classes = kernel_classes(cfg, B, chain=False)
        iso = {}
        for r in range(args.profile_reps + 2):
            _lib.check(L.tamf_denoiser_profile_step(model._handle, _lib.ptr(x), 500, 7, ms_buf, 64, C.byref(n_out),
                                                    _lib.stream_ptr(dev)), "profile_step")
            if r >= 2:
                for i, (name, _) in enumerate(classes):
                    iso[name] = iso.get(name, 0.0) + ms_buf[i] / args.profile_reps
        per, step_us, idle = in_graph_breakdown(model, L, x, classes, dev_batch)
        top = max(per, key=lambda k: per[k]["us"])
        c = per[top]
        achieved = c["flops"] / (c["us"] * 1e-6) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(top)
        step_fl = flops_per_seq_step(cfg) * B
        chain_ms = ms_total / args.steps / args.chain_steps / (len(chains) if strong else 1)
        roof = {
            "bound": "tensor", "kernel": top, "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
            "frac": achieved / pk["sustained"], "traffic": traffic,
            "peak_source": pk["src"] + " (sustained: the kernel is timed inside the captured step)",
            "launch_us_in_graph": c["us"] / c["launches"], "launches_per_step": c["launches"],
            "flops_per_launch": c["flops"] / c["launches"], "share_of_step": c["us"] / step_us,
            "how": "globaltimer stamps (end of dependency wait -> last CTA exit) of every launch inside the captured "
                   "step graph (tamf_denoiser_profile_graph), replayed right behind 500 steps of the chain so that the "
                   "clock is the sustained one",
            "isolated": {"launch_ms": iso[top] / c["launches"],
                         "achieved": c["flops"] / (iso[top] * 1e-3) / 1e12, "peak": pk["burst"],
                         "frac": c["flops"] / (iso[top] * 1e-3) / 1e12 / pk["burst"],
                         "how": "CUDA events around each eager launch (includes ~6 us of launch gap per kernel)"},
            "step": {"achieved": step_fl / (chain_ms * 1e-3) / 1e12, "ms_per_denoiser_eval": chain_ms,
                     "frac": step_fl / (chain_ms * 1e-3) / 1e12 / pk["sustained"], "peak": pk["sustained"]},
            "kernels_in_graph_us": {k: round(v["us"], 2) for k, v in per.items()},
            "kernels_in_graph_tflops": {k: round(v["flops"] / (v["us"] * 1e-6) / 1e12, 1) for k, v in per.items() if v["flops"]},
            "in_graph_step_us": round(step_us, 2), "in_graph_idle_frac": round(idle, 4),
            "kernels_isolated_ms": {k: round(v, 4) for k, v in iso.items()},
        }

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec = cpu_reference_sample(B, args.ref_diff_steps)
        cpu = {"value": B / (sec * DIFF_STEPS), "unit": UNIT, "cores": host_threads(), "kind": reference_kind(),
               "host_cpus": os.cpu_count(),
               "sample": f"{args.ref_diff_steps} of {DIFF_STEPS} diffusion steps at B={B} ({sec:.2f} s each), extrapolated "
                         f"linearly; " + _KIND_NOTE[reference_kind()]}

    if rank == 0:
        config = workload_config(B, world, args.chain_steps)
        if strong:
            config["workload"] = (f"MF-MDM G {ARCH} sampling batch-sharded over {world} B200, {args.sequences} synthetic "
                                  f"sequences in chains of {B}, T={T_FRAMES}, nobj={NOBJ}, full {args.chain_steps}-step "
                                  "chains, NCCL gather of outputs (BASELINE.json configs[3])")
            config["global_batch"] = args.sequences
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "measurement": {"parallelism": (f"dp{world}: contiguous sequence ranges per rank, one NCCL all_gather of the "
                                            "samples") if world > 1 else "single GPU",
                            "l2": "256 MB flush between timed steps; step working set 180 MB > 126 MB L2",
                            "noise": "in-kernel Philox4x32-10", "timing": "CUDA events on the launch stream, max over ranks"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_refine(args):
    """BASELINE.json configs[2]: MF-MDM R (arch_refine) forward with ManoLayer FK and the hand->object NN query, batch 64,
    T=160, one object of 8192 points.  A step = one SegmentRefineModel.forward (3 FK + 3 NN + transformer) per rank."""
    import numpy as np
    import torch
    import torch.distributed as dist

    import tamf_b200
    from tamf_b200 import _lib, synth

    world, rank, local, dev = _dist_setup()
    L = _lib.lib()
    B, T, P, nobj = args.batch, T_FRAMES, 8192, 1
    cfg = synth.ARCH["arch_refine"]
    m = tamf_b200.SegmentRefineModel("unused", **cfg, use_pc=True,
                                     mano_assets={"right": synth.mano_assets("right"), "left": synth.mano_assets("left")})
    m.load_state_dict(synth.r_state_dict(cfg, 0), strict=False)
    m = m.eval().to(dev)
    host = synth.make_batch(B, T, nobj=nobj, seed=1 + rank, npoints=P, with_pointcloud=True)
    batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step():
        flush.zero_()
        return m(batch)
    # the first forwards of a process pay allocator growth and lazy kernel loading (36 ms, then 16 ms: tools/diag_refine.py)
    args.warmup = max(args.warmup, 5)
    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = L.tamf_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        out = step()
    e1.record(stream)
    barrier()
    launches = L.tamf_kernel_launch_count() - n0
    clk = clocks.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    assert all(torch.isfinite(v).all() for v in out.values())
    value = world * B * args.steps / (ms_total * 1e-3)
    # e2e: host tensors in (pinned), the 13-key dict back on the host
    keys_in = ("sample_pose_repr", "pose_repr", "shape", "obj_traj", "obj_embedding")
    h2d = sum(pinned[k].numel() * pinned[k].element_size() for k in keys_in) + B * nobj * P * 12
    # (the results land in pinned host buffers allocated once: 681 MB per step through pageable memory ran at 2 GB/s)
    host_out = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in out.items()}
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o = m({k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in pinned.items()})
        for k, v in o.items():
            host_out[k].copy_(v, non_blocking=True)
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())
    roof = None
    if rank == 0:
        # dominant kernel: the block-pruned hand->object search (3 calls per forward), HBM roofline on the algorithmic bytes
        from tamf_b200.chamfer import H2OIndex
        hv = out["sample_hand_verts"].contiguous()
        traj = batch["obj_traj"].float().contiguous()
        pts = [np.asarray(o_, np.float32)[: len(l)] for o_, l in zip(batch["obj_pointcloud"], batch["obj_list"])]
        oix = H2OIndex(pts, dev)
        dist_t = torch.empty((B, T, 778), device=dev)
        idx_t = torch.empty((B, T, 778), dtype=torch.int64, device=dev)

        def nn_q():
            _lib.check(L.tamf_h2o_dist_indexed(_lib.ptr(hv), _lib.ptr(traj), _lib.ptr(oix.index),
                                               _lib.C.c_void_p(oix.first.data_ptr()), B, T, 778, traj.shape[1], P,
                                               _lib.ptr(dist_t), _lib.ptr(idx_t), _lib.stream_ptr(dev)), "h2o indexed")
        for _ in range(3):
            nn_q()
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a0.record(stream)
        for _ in range(reps):
            nn_q()
        a1.record(stream)
        torch.cuda.synchronize(dev)
        nn_ms = a0.elapsed_time(a1) / reps
        N = B * T
        nn_bytes = N * (778 * 12 + 778 * 12 + 36.0) + oix.total_obj * P * 12.0  # 18.7 KB + 36 B per frame
        pk = peaks()
        ach = nn_bytes / (nn_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "h2o_pruned_kernel (+ finalize)", "achieved": ach, "peak": pk["hbm"],
                "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": None, "peak_source": pk["src"],
                "launch_ms": nn_ms, "launches_per_step": 3, "share_of_step": 3 * nn_ms / (ms_total / args.steps),
                "note": "exact block-pruned search: instruction bound (DESIGN.md 4.6), far from the HBM roofline by design"}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import tamf_oracle as orc
        import torch as _t
        _t.set_num_threads(host_threads())
        nb = 4
        sub = {k: (v[:nb] if isinstance(v, (torch.Tensor, list)) else v) for k, v in host.items()}
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.r_forward(synth.r_state_dict(cfg, 0), cfg, sub, synth.mano_assets("right"), synth.mano_assets("left"))
        sec = time.perf_counter() - t0
        cpu = {"value": nb / sec, "unit": "sequences/s", "cores": host_threads(), "kind": "port", "host_cpus": os.cpu_count(),
               "sample": f"{nb} of {B} sequences through the oracle's r_forward ({sec:.1f} s)"}
    if rank == 0:
        line = {
            "metric": "refined motion sequences/sec, SegmentRefineModel forward (3 FK + 3 NN + transformer)", "value": value,
            "unit": "sequences/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (FK, NN) / bf16 (transformer)", "data": "synthetic",
            "config": {"workload": f"MF-MDM R arch_refine, batch {B} synthetic sequences per GPU, T={T}, {nobj} object x {P} "
                                   "points, use_pc (BASELINE.json configs[2])", "global_batch": world * B,
                       "weights": "random init"},
            "measurement": {"l2": "256 MB flush between timed steps", "timing": "CUDA events on the launch stream"},
            "e2e": {"value": world * B * args.steps / float(dt.item()), "unit": "sequences/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="sample", choices=["sample", "refine"],
                    help="sample: BASELINE.json configs[1] (the metric); refine: configs[2] (MF-MDM R forward)")
    ap.add_argument("--sequences", type=int, default=0,
                    help="> 0: BASELINE.json configs[3] -- a FIXED set of sequences batch-sharded over the ranks (strong scaling)")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--chain-steps", type=int, default=DIFF_STEPS, help="debug only: shorter chains are not the metric")
    ap.add_argument("--ref-diff-steps", type=int, default=24,
                    help="at most this many diffusion steps per timed CPU sample (about 10 s of host work at B=64)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="wall-time budget of the whole reference-arm run; the per-step sample shrinks to fit")
    ap.add_argument("--profile-reps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 1 if args.sequences > 0 else (10 if args.config == "refine" else 3)
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "refine":
        run_refine(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
