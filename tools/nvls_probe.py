#!/usr/bin/env python
"""Probe (under torchrun on N GPUs of one box): does this box give us NVSwitch multicast memory through torch's
symmetric-memory rendezvous, and what does an in-switch (multimem) all-reduce of a row-parallel partial cost next to
NCCL's?  Prints one JSON line on rank 0.  Plumbing only -- nothing here is on the product path.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def timed(fn, iters, dev):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world, "torch": torch.__version__}
    try:
        import torch.distributed._symmetric_memory as symm
        out["backend"] = str(symm.get_backend(dev)) if hasattr(symm, "get_backend") else None
        M, N = 8192, 4096
        t = symm.empty((M, N), dtype=torch.bfloat16, device=dev)
        h = symm.rendezvous(t, dist.group.WORLD.group_name)
        out["multicast_ptr"] = int(getattr(h, "multicast_ptr", 0) or 0)
        out["buffer_ptrs"] = len(h.buffer_ptrs)
        out["signal_pad_size"] = int(getattr(h, "signal_pad_size", 0))
        out["buffer_size"] = int(getattr(h, "buffer_size", 0))
        t.normal_()
        ref = t.clone()
        dist.all_reduce(ref)
        torch.cuda.synchronize()
        plain = torch.randn((M, N), device=dev).to(torch.bfloat16)
        out["nccl_us"] = timed(lambda: dist.all_reduce(plain), 30, dev)
        name = dist.group.WORLD.group_name
        for op in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
            try:
                f = getattr(torch.ops.symm_mem, op)
                out[op + "_us"] = timed(lambda: f(t, "sum", name), 30, dev)
            except Exception as e:  # noqa: BLE001
                out[op + "_err"] = repr(e)[:200]
    except Exception as e:  # noqa: BLE001
        out["error"] = repr(e)[:400]
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
