#!/usr/bin/env python
"""Model-level measurements of BASELINE configs 3-5 on the B200 box (one JSON line per config on stdout).

  config 3  python tools/bench_models.py prefill [--layers 32] [--batch 8] [--seq 2048]
            full Llama-3-8B prefill through QLlamaDecoderLayer, random-init weights, tokens/s
  config 4  torchrun --nproc-per-node N tools/bench_models.py qwen_tp [--tokens 16384]
            one Qwen2.5-32B-shaped layer, tensor parallel N ways (column qkv/gate_up, row o/down + all-reduce)
  config 5  [torchrun --nproc-per-node N] tools/bench_models.py mixtral_ep [--tokens 16384]
            Mixtral-8x7B MoE block (8 experts, top-2), experts sharded over N ranks

Timing: CUDA events around `iters` forwards after warm-up, barrier + synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from micromix_b200 import mixedgemm  # noqa: E402
from micromix_b200 import model_shapes as S  # noqa: E402


def setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        from micromix_b200.parallel_utils import init_tensor_parallel
        rank, world, dev = init_tensor_parallel("nccl")
    else:
        rank, dev = 0, torch.device("cuda", 0)
        torch.cuda.set_device(dev)
    return rank, world, dev


def timed(fn, iters, warmup, world, dev):
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = mixedgemm.launch_count()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, (mixedgemm.launch_count() - l0) // iters


def layer_flops(cfg, tokens, moe=False):
    h, i, d = cfg["hidden_size"], cfg["intermediate_size"], cfg["head_dim"]
    nq, nkv = cfg["num_attention_heads"], cfg["num_key_value_heads"]
    lin = h * (nq + 2 * nkv) * d + nq * d * h
    lin += (cfg["num_experts_per_tok"] if moe else 1) * 3 * h * i
    return 2.0 * tokens * lin


def prefill(args, rank, world, dev):
    from micromix_b200.qLlamaLayer import QLlamaDecoderLayer
    cfg = S.LLAMA3_8B
    group = dist.group.WORLD if world > 1 else None
    layers = []
    for i in range(args.layers):
        layer = S.make_layer(cfg, dev, seed=i)
        idx, p6, p8 = S.make_calibration(cfg, i)
        layers.append(QLlamaDecoderLayer(layer, False, p8, p6, idx, i, tp_group=group, fused=args.fused))
        del layer
    torch.cuda.empty_cache()
    b, s = args.batch, args.seq
    g = torch.Generator(device=dev).manual_seed(721)
    x0 = torch.randn(b, s, cfg["hidden_size"], generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    pos = S.rope_tables(cfg, b, s, dev)

    @torch.no_grad()
    def fwd():
        x = x0
        for l in layers:
            x = l(x, position_embeddings=pos)[0]
        return x

    ms, launches = timed(fwd, args.iters, 2, world, dev)
    y = fwd()
    assert torch.isfinite(y.float()).all()
    tokens = b * s
    scale = cfg["num_hidden_layers"] / args.layers
    return {"config": "Llama-3-8B full prefill, random-init weights, synthetic reorder_index, "
                      f"batch {b} x seq {s}", "layers_run": args.layers, "ms_per_prefill": ms * scale,
            "tokens_per_s": tokens / (ms * scale) * 1e3, "linear_tflops": layer_flops(cfg, tokens) * args.layers / ms / 1e9,
            "n_gpus": world, "parallelism": f"tp{world}" if world > 1 else "single", "mmx_launches_per_forward": launches,
            "fused_norm_act": bool(args.fused),
            "note": "decoder layers only (no embedding / lm_head, as in the reference's layer-wise eval); attention = SDPA"}


def qwen_tp(args, rank, world, dev):
    from micromix_b200.qQwenLayer import QQwen2DecoderLayer
    cfg = S.QWEN25_32B
    group = dist.group.WORLD if world > 1 else None
    layer = S.make_layer(cfg, dev, seed=0)
    idx, p6, p8 = S.make_calibration(cfg, 0)
    b, s = max(1, args.tokens // 2048), 2048
    ws = None
    sp = bool(args.sp or args.tpr) and world > 1 and args.tp_reduce == "fused"
    if world > 1 and args.tp_reduce == "fused":
        from micromix_b200.parallel_utils import PeerWorkspace
        ws = PeerWorkspace(b * s, cfg["hidden_size"], group=group, device=dev,
                           gather=(b * s, max(cfg["hidden_size"], -(-cfg["intermediate_size"] // world // 128) * 128))
                           if sp else None)
    q = QQwen2DecoderLayer(layer, False, p8, p6, idx, 0, tp_group=group, workspace=ws, sequence_parallel=sp,
                           fused=bool(args.fused), token_parallel_rows=bool(args.tpr))
    del layer
    torch.cuda.empty_cache()
    b, s = max(1, args.tokens // 2048), 2048
    g = torch.Generator(device=dev).manual_seed(721)
    x0 = torch.randn(b, s, cfg["hidden_size"], generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    pos = S.rope_tables(cfg, b, s, dev)
    if sp:
        lo, hi = ws.shard_range(b * s)
        x0 = x0.reshape(b * s, -1)[lo:hi].unsqueeze(0).contiguous()  # this rank's token rows
    fwd = lambda: q(x0, position_embeddings=pos)
    ms, launches = timed(fwd, args.iters, 3, world, dev)
    # the same forward replayed from a CUDA graph: no Python / launch overhead between the layer's ~40 kernels
    ms_graph = None
    try:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fwd()
        torch.cuda.synchronize()
        ms_graph, _ = timed(graph.replay, max(args.iters, 5), 2, world, dev)
    except Exception as e:  # noqa: BLE001
        ms_graph = f"capture failed: {e!r}"[:120]
    tokens = b * s
    how = "NCCL all-reduce" if ws is None else (f"sequence parallel: fused GEMM->reduce-scatter + multicast all-gather of packed "
                                                 f"codes ({ws.mode})" if sp else f"fused GEMM->all-reduce ({ws.mode})")
    if args.tpr and sp:
        how = "token-parallel o/down (replicated MXFP4 weights, all-to-all of packed codes) + multicast all-gather of codes"
    best = ms_graph if isinstance(ms_graph, float) else ms
    return {"config": f"Qwen2.5-32B-shaped decoder layer, {tokens} tokens, tensor parallel {world} (column qkv/gate_up, "
                      f"row o/down + {how})", "ms_per_layer": ms, "ms_per_layer_cuda_graph": ms_graph,
            "tokens_per_s_per_layer": tokens / best * 1e3, "linear_tflops": layer_flops(cfg, tokens) / best / 1e9,
            "n_gpus": world, "mmx_launches_per_forward": launches, "tp_status": ws.status() if ws is not None else None}


def mixtral_ep(args, rank, world, dev):
    from micromix_b200.qMixtralLayer import QMixtralSparseMoeBlock
    cfg = S.MIXTRAL_8X7B
    group = dist.group.WORLD if world > 1 else None
    layer = S.make_layer(cfg, dev, seed=0, moe=True)
    idx, p6, p8 = S.make_calibration(cfg, 0, moe=True)
    blk = QMixtralSparseMoeBlock(layer.block_sparse_moe, p8, p6, idx, 0, ep_group=group, fused=args.fused,
                                 grouped=None if not args.loop else False)
    del layer
    torch.cuda.empty_cache()
    tokens = args.tokens
    g = torch.Generator(device=dev).manual_seed(721)
    x0 = torch.randn(1, tokens, cfg["hidden_size"], generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    ms, launches = timed(lambda: blk(x0), args.iters, 3, world, dev)
    ms_graph = None
    if blk.grouped:  # no host synchronisation anywhere: the whole block is one CUDA graph
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                blk(x0)
            torch.cuda.synchronize()
            ms_graph, _ = timed(graph.replay, max(args.iters, 5), 2, world, dev)
        except Exception as e:  # noqa: BLE001
            ms_graph = f"capture failed: {e!r}"[:120]
    best = ms_graph if isinstance(ms_graph, float) else ms
    flops = 2.0 * tokens * cfg["num_experts_per_tok"] * 3 * cfg["hidden_size"] * cfg["intermediate_size"]
    out = {"config": f"Mixtral-8x7B expert FFN (8 experts, top-2), {tokens} tokens, expert parallel {world}",
           "path": "grouped (one quantize + one GEMM launch per projection, fused token gather, combine kernel)" if blk.grouped
           else "python loop over experts (the reference's op sequence)",
           "ms_per_block": ms, "ms_per_block_cuda_graph": ms_graph, "tokens_per_s": tokens / best * 1e3,
           "expert_tflops": flops / best / 1e9, "expert_tflops_per_gpu": flops / best / 1e9 / world, "n_gpus": world,
           "mmx_launches_per_forward": launches, "fused_act": bool(args.fused)}
    try:  # against the measured MX tensor-pipe peak of this GPU (burst), 5:2:1 split
        import ctypes
        lib = mixedgemm._lib.load()
        pk = []
        for kind in (0, 1, 2):
            t = ctypes.c_double()
            lib.mmx_debug_mma_peak(kind, 2000, 0, 3, ctypes.byref(t), None)
            pk.append(t.value)
        peak = 1.0 / (0.625 / pk[0] + 0.25 / pk[1] + 0.125 / pk[2])
        out["mx_peak_burst_tflops_split_weighted"] = peak
        out["expert_roofline_frac_per_gpu"] = out["expert_tflops_per_gpu"] / peak
    except Exception:  # noqa: BLE001
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["prefill", "qwen_tp", "mixtral_ep"])
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2048)
    ap.add_argument("--tokens", type=int, default=16384)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--tp-reduce", default="fused", choices=["fused", "nccl"], help="qwen_tp: row-parallel reduction")
    ap.add_argument("--loop", action="store_true", help="mixtral_ep: the per-expert Python loop instead of the grouped path")
    ap.add_argument("--sp", action="store_true", help="qwen_tp: sequence-parallel layer (reduce-scatter + all-gather of codes)")
    ap.add_argument("--tpr", action="store_true", help="qwen_tp: sequence-parallel layer with token-parallel o / down")
    ap.add_argument("--fused", action="store_true", help="prefill: RMSNorm and SiLU*up run inside the quantizers (QDecoderLayer(fused=True))")
    args = ap.parse_args()
    rank, world, dev = setup()
    try:
        res = {"prefill": prefill, "qwen_tp": qwen_tp, "mixtral_ep": mixtral_ep}[args.what](args, rank, world, dev)
        if rank == 0:
            print(json.dumps(res), flush=True)
    finally:
        if world > 1 and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
