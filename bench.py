#!/usr/bin/env python
"""bench.py -- sampled motion sequences/sec, full 1000-step reverse chain, MF-MDM G arch_mdm_l (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (libtamf_b200.so, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port), rank 0 only
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1: one rank per GPU, batch-sharded, NCCL gather

A "step" is one pass of the hot path over one batch: conditioning (once per sample) + x_T ~ N(0,I) + the 1000
p_sample steps for B=64 synthetic sequences per GPU (BASELINE.json configs[1]).  One JSON line on stdout (rank 0).

  value     whole-job sequences/s with every input resident in HBM when the timed region starts
  e2e       the same through the reference-facing call (InterationSegmentMDM.sample_host -> tamf_p_sample_loop_host)
            with HOST buffers: H2D of the conditioning and D2H of the samples inside the timed region
  roofline  the dominant kernel of the step (per-kernel CUDA-event times from tamf_denoiser_profile_step, live),
            algorithmic FLOPs per launch / mean launch duration vs the measured bf16 peak (MEASURED_PEAKS.json)
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


def kernel_classes(cfg, B, T=T_FRAMES, chain=True):
    """(name, algorithmic FLOPs per launch) in the launch order tamf_denoiser_profile_step reports.  chain=True: the
    layer-kernel form of the encoder (csrc/layer_chain.cuh): in_proj of layer 0, then per layer attention | out_proj+LN1 ->
    linear1+GELU -> linear2+LN2 -> in_proj of the next layer in ONE kernel; chain=False: the five-kernel layer (TAMF_CHAIN=0)."""
    d, ff, L = cfg["latent_dim"], cfg["ff_size"], cfg["num_layers"]
    S = T + 5
    M, Mf = B * S, B * T
    out = [("prep", 0), ("embed_a", 2 * Mf * 99 * d), ("embed_b", 2 * Mf * d * d)]
    f_in, f_att, f_out, f_l1, f_l2 = 2 * M * 3 * d * d, 4 * B * S * S * d, 2 * M * d * d, 2 * M * d * ff, 2 * M * ff * d
    if chain:
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


def cpu_reference_sample(B, n_diff):
    """Oracle port of the reference's per-step algorithm (InterationSegmentMDM.forward + p_sample) on the host cores:
    `n_diff` diffusion steps at batch B, arch_mdm_l, torch intra-op threads = every host core.  Returns seconds per
    diffusion step."""
    import torch

    from oracle import tamf_oracle as orc
    from tamf_b200 import synth
    if torch.get_num_threads() != host_threads():
        torch.set_num_threads(host_threads())
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
            "diffusion_steps": chain_steps, "weights": "random init", "text_features": "synthetic (CLIP tower excluded)"}


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port; the Python reference cannot travel to the GPU box
    and pytorch3d/CLIP weights are absent) on every host core.  Rank 0 only; other ranks exit 0 without work.  Each timed
    step is a bounded sample of the workload -- `n_diff` of the 1000 diffusion steps at the full batch, extrapolated
    linearly -- sized from a first measured step so that the whole --steps/--warmup run stays within --ref-budget-s."""
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
    sample = (f"{n_diff} of {DIFF_STEPS} diffusion steps at B={B} per timed step (x{args.steps} steps, "
              f"{sec_per_diff:.3f} s per diffusion step), extrapolated linearly to the full chain; CLIP text tower "
              f"excluded (text feature given)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "notes": "reference arm = oracle port on the host cores, rank 0 only; ms_per_step is the linear extrapolation of "
                 "the bounded sample to the full 1000-step chain",
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import tamf_b200
    from tamf_b200 import _lib, synth

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
    L = _lib.lib()
    cfg = synth.ARCH[ARCH]
    B, T = args.batch, T_FRAMES
    model = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    model.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    model = model.eval().to(dev)
    host_batch = synth.make_batch(B, T, nobj=NOBJ, seed=100 + rank)
    pinned = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    dev_batch = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host_batch.items()}
    x = torch.empty(B, 99, 1, T, device=dev)
    gathered = torch.empty(world * B, 99, 1, T, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_step(seed):
        flush.zero_()  # L2 flush between timed iterations (the step's own working set, 180 MB, also exceeds L2)
        model.set_cond(dev_batch, B, T, dev)  # conditioning: once per sample batch
        _lib.check(L.tamf_philox_normal(_lib.ptr(x), x.numel(), seed, DIFF_STEPS, _lib.stream_ptr(dev)), "x_T")
        model.p_sample_chain(x, DIFF_STEPS - 1, DIFF_STEPS - args.chain_steps, dev_batch, seed=seed)
        if world > 1:
            dist.all_gather_into_tensor(gathered, x)  # the one collective of the path: final gather of the samples

    # ---------------- device-resident throughput ----------------
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
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---------------- end to end through the host-buffer API ----------------
    h2d = sum(pinned[k].numel() * 4 for k in ("shape", "obj_traj", "obj_embedding")) + B * 512 * 4 + B * 4
    d2h = B * 99 * T * 4
    e2e_steps = max(1, args.steps)
    model.sample_host(pinned, seed=1)  # warm-up (graph already captured; new workspace binding is reused)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        out = model.sample_host(pinned, seed=3000 + i)
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(dt.item())
    assert torch.isfinite(out).all()

    # ---------------- per-kernel times (live, CUDA events on the launch stream) -> roofline ----------------
    roof, per_class = None, None
    if rank == 0:
        import ctypes as C
        classes = kernel_classes(cfg, B)
        ms_buf = (C.c_float * 64)()
        n_out = C.c_int(0)
        acc = [0.0] * len(classes)
        reps = 0
        model.set_cond(dev_batch, B, T, dev)
        for r in range(args.profile_reps + 2):
            _lib.check(L.tamf_denoiser_profile_step(model._handle, _lib.ptr(x), 500, 7, ms_buf, 64, C.byref(n_out),
                                                    _lib.stream_ptr(dev)), "profile_step")
            if n_out.value != len(classes):  # TAMF_CHAIN=0: the five-kernel layer of round 1
                classes = kernel_classes(cfg, B, chain=False)
                acc = [0.0] * len(classes)
            assert n_out.value == len(classes)
            if r >= 2:
                reps += 1
                for i in range(len(classes)):
                    acc[i] += ms_buf[i]
        per_class = {}
        for (name, fl), a in zip(classes, acc):
            c = per_class.setdefault(name, {"ms": 0.0, "launches": 0, "flops": 0})
            c["ms"] += a / reps
            c["launches"] += 1
            c["flops"] += fl
        step_ms_sum = sum(c["ms"] for c in per_class.values())
        top = max(per_class, key=lambda k: per_class[k]["ms"])
        c = per_class[top]
        pk = peaks()
        achieved = c["flops"] / (c["ms"] * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(top)
        roof = {"bound": "tensor", "kernel": top, "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["sustained"], "traffic": traffic, "peak_source": pk["src"] + " (sustained)",
                "launch_ms": c["ms"] / c["launches"], "share_of_step": c["ms"] / step_ms_sum,
                "flops_per_launch": c["flops"] / c["launches"]}
        step_fl = flops_per_seq_step(cfg) * B
        chain_ms = ms_total / args.steps / args.chain_steps
        roof["step"] = {"achieved": step_fl / (chain_ms * 1e-3) / 1e12, "ms_per_denoiser_eval": chain_ms,
                        "frac": step_fl / (chain_ms * 1e-3) / 1e12 / pk["sustained"]}
        roof["kernels_ms"] = {k: round(v["ms"], 4) for k, v in per_class.items()}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        sec = cpu_reference_sample(B, args.ref_diff_steps)
        cpu = {"value": B / (sec * DIFF_STEPS), "unit": UNIT, "cores": _t.get_num_threads(), "kind": "port",
               "host_cpus": os.cpu_count(),
               "sample": f"{args.ref_diff_steps} of {DIFF_STEPS} diffusion steps at B={B} ({sec:.2f} s each), extrapolated "
                         f"linearly; CLIP text tower excluded"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"MF-MDM G {ARCH}, batch {B} synthetic sequences per GPU, T={T_FRAMES}, nobj={NOBJ}, "
                                   f"full {args.chain_steps}-step reverse chain (BASELINE.json configs[1])",
                       "global_batch": world * B, "parallelism": f"dp{world} (batch-sharded chains, NCCL all_gather of "
                                                                 "the samples)" if world > 1 else "single GPU",
                       "l2": "256 MB flush between timed steps; step working set 180 MB > 126 MB L2",
                       "weights": "random init", "noise": "in-kernel Philox4x32-10"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--chain-steps", type=int, default=DIFF_STEPS, help="debug only: shorter chains are not the metric")
    ap.add_argument("--ref-diff-steps", type=int, default=24,
                    help="at most this many diffusion steps per timed CPU sample (about 10 s of host work at B=64)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="wall-time budget of the whole reference-arm run; the per-step sample shrinks to fit")
    ap.add_argument("--profile-reps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
