#!/usr/bin/env python
"""Development probe (2+ GPUs, torchrun): can a kernel of ours read a peer's buffer through
torch symmetric memory, and at what bandwidth?  Not part of the product."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem
    from gmlm_b200.ops import gather_rows, scatter_add_rows_

    rows, feat = 2_000_000, 256
    t = symm_mem.empty((rows, feat), dtype=torch.bfloat16, device=dev)
    t.copy_(torch.full((rows, feat), float(rank + 1), dtype=torch.bfloat16, device=dev))
    t[:, 0] = torch.arange(rows, device=dev).remainder(251).to(torch.bfloat16)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD.group_name)
    print(f"[{rank}] rendezvous ok: {type(hdl).__name__}", flush=True)
    peer = (rank + 1) % world
    pt = hdl.get_buffer(peer, (rows, feat), torch.bfloat16)
    print(f"[{rank}] peer buffer ptr {pt.data_ptr():#x} local {t.data_ptr():#x}", flush=True)
    hdl.barrier()
    ids = torch.randperm(rows, device=dev)[:1_500_000]
    out = gather_rows(pt, ids)
    torch.cuda.synchronize()
    ok = bool((out[:, 1] == float(peer + 1)).all()) and bool((out[:, 0] == ids.remainder(251).to(torch.bfloat16)).all())
    print(f"[{rank}] remote gather correct: {ok}", flush=True)
    for name, fn in (("random-row pull", lambda: gather_rows(pt, ids, out=out)),
                     ("contiguous pull-add", lambda: scatter_add_rows_(out, ids_seq, pt[: ids.numel()]))):
        ids_seq = torch.arange(ids.numel(), device=dev)
        hdl.barrier()
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        hdl.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"[{rank}] {name}: {ms:.3f} ms  {ids.numel() * feat * 2 / ms / 1e6:.0f} GB/s", flush=True)
    # barrier cost
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        hdl.barrier()
    b.record()
    torch.cuda.synchronize()
    print(f"[{rank}] symm barrier: {a.elapsed_time(b) / 20 * 1e3:.1f} us", flush=True)
    # all peers at once (what the halo pull does): each rank pulls a slice from every other rank
    if world > 2:
        bufs = [hdl.get_buffer(p, (rows, feat), torch.bfloat16) for p in range(world)]
        per = ids.numel() // (world - 1)
        streams = [torch.cuda.Stream() for _ in range(world)]
        hdl.barrier()
        torch.cuda.synchronize()
        a.record()
        for it in range(5):
            for p in range(world):
                if p == rank:
                    continue
                gather_rows(bufs[p], ids[:per], out=out[(p if p < rank else p - 1) * per:][:per])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"[{rank}] pull from all {world-1} peers (serial launches): {ms:.3f} ms "
              f"{per * (world - 1) * feat * 2 / ms / 1e6:.0f} GB/s", flush=True)
    hdl.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
