#!/usr/bin/env python
"""BASELINE.json configs[3]: arch_mdm_l sampling of a FIXED set of synthetic sequences (default 8192) batch-sharded over
the N GPUs of one box -- contiguous index ranges per rank exactly like launch/sample.py:198-199
(tamf_b200.shard.shard_range), batches of 64 per rank, every batch a full 1000-step reverse chain, ONE NCCL
all_gather of the finished [n_local,160,99] samples at the end (strong scaling: total work fixed).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_shard.py [--sequences 8192]

Rank 0 prints one JSON line: sequences/s over the whole job (CUDA events on each rank, max over ranks, gather included).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oakink2-tamf_b200"))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--chain-steps", type=int, default=1000, help="debug only: shorter chains are not the metric")
    a = ap.parse_args()
    import tamf_b200
    from tamf_b200 import _lib, shard, synth
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.ARCH["arch_mdm_l"]
    T = 160
    model = tamf_b200.InterationSegmentMDM(**cfg, text_encoder=synth.text_features)
    model.load_state_dict(synth.g_state_dict(cfg, seed=0), strict=False)
    model = model.eval().to(dev)
    diffusion = tamf_b200.create_gaussian_diffusion(1000, "cosine")
    mine = shard.shard_range(a.sequences, rank, world)
    # two synthetic conditioning batches alternate (content does not change the timing; set_cond runs per batch)
    conds = [{k: (v.to(dev) if isinstance(v, torch.Tensor) else v)
              for k, v in synth.make_batch(a.batch, T, nobj=2, seed=200 + rank * 2 + i).items()} for i in range(2)]
    out = torch.empty(len(mine), T, 99, device=dev)
    L = _lib.lib()

    def run_shard():
        for i, ids in enumerate(shard.batches(mine, a.batch)):
            nb = len(ids)
            cb = conds[i % 2] if nb == a.batch else {k: (v[:nb] if isinstance(v, (torch.Tensor, list)) else v)
                                                      for k, v in conds[i % 2].items()}
            x = torch.empty(nb, 99, 1, T, device=dev)
            _lib.check(L.tamf_philox_normal(_lib.ptr(x), x.numel(), 7000 + ids.start, 1000, _lib.stream_ptr(dev)), "x_T")
            diffusion._install(model, "ancestral")
            model.p_sample_chain(x, 999, 1000 - a.chain_steps, cb, seed=9000 + ids.start)
            out[ids.start - mine.start: ids.stop - mine.start] = x.permute(0, 3, 1, 2).squeeze(3)  # extract_sample.py:32
        return shard.gather_samples(out, a.sequences)

    # warm-up: one short batch (graph capture, workspace binding), then the timed job
    warm = torch.empty(a.batch, 99, 1, T, device=dev).normal_()
    model.p_sample_chain(warm, 999, 990, conds[0], seed=1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.current_stream(dev)
    e0.record(stream)
    full = run_shard()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    assert full.shape == (a.sequences, T, 99) and torch.isfinite(full).all()
    if rank == 0:
        sec = float(ms.item()) * 1e-3
        print(json.dumps({
            "metric": "sampled motion sequences/sec, full reverse chain, arch_mdm_l", "value": a.sequences / sec,
            "unit": "sequences/s", "n_gpus": world, "seconds": sec, "scaling": "strong", "dtype": "bf16",
            "data": "synthetic", "higher_is_better": True,
            "config": {"workload": f"arch_mdm_l sampling batch-sharded over {world} B200, {a.sequences} synthetic "
                                   f"sequences, batches of {a.batch}, full {a.chain_steps}-step chains, NCCL gather of "
                                   "outputs (BASELINE.json configs[3])",
                       "per_rank_sequences": len(mine), "gather": f"all_gather of [{len(mine)},160,99] fp32 per rank"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
