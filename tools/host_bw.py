#!/usr/bin/env python
"""Host <-> device copy ceiling of the box at N GPUs (diagnostic for bench.py's e2e scaling; not part of the library).

    python tools/host_bw.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/host_bw.py

Every rank copies between its own page-locked host buffers and its GPU with plain cudaMemcpyAsync -- all ranks at the same
time -- in three modes: H2D only, D2H only, both directions at once in bench.py's e2e ratio (40.6 MB up : 29.5 MB down per
field with 8-bit frames; 73.7 : 29.5 with fp32 frames).  Rank 0 prints one JSON line with per-rank and whole-box GB/s.
The e2e rate bench.py can reach at N GPUs is bounded by  (whole-box GB/s) / (bytes per field)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from bench import _bind_to_gpu_numa_node
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    aff = _bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    chunk = 64 << 20
    n_up, n_dn = 8, 8
    hu = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(n_up)]
    hd = [torch.empty(chunk, dtype=torch.uint8).pin_memory() for _ in range(n_dn)]
    for t in hu + hd:
        t.fill_(1)  # touch: first-touch NUMA placement happens here, after the affinity call
    du = [torch.empty(chunk, dtype=torch.uint8, device="cuda") for _ in range(n_up)]
    dd = [torch.ones(chunk, dtype=torch.uint8, device="cuda") for _ in range(n_dn)]
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up_chunks, dn_chunks, reps=6):
        best = None
        for rep in range(reps):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(s_up):
                for k in range(up_chunks):
                    du[k % n_up].copy_(hu[k % n_up], non_blocking=True)
            with torch.cuda.stream(s_dn):
                for k in range(dn_chunks):
                    hd[k % n_dn].copy_(dd[k % n_dn], non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            if rep > 0 and (best is None or dt < best):
                best = dt
        return best

    res = {}
    for name, up, dn in (("h2d_only", 32, 0), ("d2h_only", 0, 32), ("both_u8_ratio", 22, 16), ("both_f32_ratio", 40, 16)):
        dt = run(up, dn)
        res[name] = {"per_rank_gbs_up": up * chunk / dt / 1e9, "per_rank_gbs_down": dn * chunk / dt / 1e9,
                     "box_gbs_total": world * (up + dn) * chunk / dt / 1e9}
    mb_u8, mb_f32 = 40.6 + 29.5, 73.7 + 29.5
    res["implied_e2e_ceiling_fields_per_s"] = {
        "u8_frames": res["both_u8_ratio"]["box_gbs_total"] * 1e3 / mb_u8,
        "f32_frames": res["both_f32_ratio"]["box_gbs_total"] * 1e3 / mb_f32,
        "note": "whole box; bytes per 2560x1440 field = %.1f MB (8-bit frames) / %.1f MB (fp32 frames), up + down" % (mb_u8, mb_f32)}
    if rank == 0:
        print(json.dumps({"tool": "host_bw", "n_gpus": world, "host_cpus": os.cpu_count(), "affinity_rank0": aff, "chunk_mb": chunk >> 20,
                          "results": res}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
