#!/usr/bin/env python
"""bench.py -- the hot path of MicroMix on B200: reorder+quantize -> three-segment mixed-MX GEMM.

Workload (BASELINE.json configs[1]): the four linears of one Llama-3-8B decoder layer
(qkv 6144x4096, o 4096x4096, gate_up 28672x4096, down 4096x14336) over M tokens, 5-bit average split
(p4,p6,p8) = K*(5/8, 2/8, 1/8), synthetic activations / random-init weights / synthetic reorder_index.
A "step" = one pass of the hot path over that batch: for each linear, mmx_reorder_quantize_x then mmx_matmul
(8 kernel launches), weights pre-quantized to MXFP4 offline exactly as QLinearLayer.__init__ does.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--tokens M] [--impl ours|reference]

N > 1 (launched by torchrun, one rank per GPU; NCCL is the plumbing): tensor parallel as in BASELINE north_star --
qkv/gate_up column-parallel (no collective), o/down row-parallel with the bf16 [M,4096] partials summed over the ranks.
The sum is fused into the GEMM by default (`--tp-reduce fused`: micromix_b200/csrc/tp_reduce.cu, peer pushes at tp=2,
in-switch multimem reduction at tp>=4); `--tp-reduce nccl` is mmx_matmul + ncclAllReduce, the baseline it is measured
against.  The step is replayed from a CUDA graph (eager launches from N Python processes would bound a step whose kernels
are 15-70 us each).  Total work is fixed as N grows ("scaling": "strong"); value = whole-job TFLOP/s =
2*M*sum(N*K) / max-over-ranks time.  stdout carries exactly one JSON line.

`--impl reference` times the reference's algorithm on the host cores (the CPU oracle port: the reference has no CPU
implementation of its own and its GEMM cannot run on sm_100) on a bounded token sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "mixed-MX GEMM TFLOPS & prefill tokens/s, Llama-3-8B linears, 1/2/4/8 B200"
_RESULT_FD = None  # the process's original stdout; fd 1 itself is pointed at stderr while the bench runs (see main)


def emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)
# (name, N, K, parallel mode)
LINEARS = [("qkv", 6144, 4096, "col"), ("o", 4096, 4096, "row"), ("gate_up", 28672, 4096, "col"),
           ("down", 4096, 14336, "row")]


def split_for(K):
    """5.0 average bits: p4 = 5/8 K, p6 = 2/8 K, p8 = 1/8 K, multiples of 128."""
    p8 = (K // 8) // 128 * 128
    p6 = (K // 4) // 128 * 128
    return K - p6 - p8, p6, p8


def flops_per_token():
    return sum(2 * n * k for _, n, k, _ in LINEARS)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def bind_to_gpu_numa_node(index):
    """Pin this process (and, by first touch, its pinned host buffers) to the NUMA node the GPU hangs off: with one process
    per GPU, eight ranks staging their host copies through one socket's memory controllers is what bounded the round-1
    end-to-end numbers at N > 1.  Best effort: returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]  # 0000:xx:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return "no NUMA node reported for the GPU"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"GPU on NUMA node {node}, none of its CPUs allowed here"
        os.sched_setaffinity(0, allowed)
        return f"bound to NUMA node {node} ({len(allowed)} CPUs)"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({e!r})"[:120]


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self.active = threading.Event()
        self.stop_flag = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                except Exception:
                    pass
            time.sleep(0.004)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "samples": len(s),
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world, emit_line=True):
    """The reference's algorithm on the host cores (oracle port; `kind`: "port").  Rank 0 only.  Returns the line."""
    if rank != 0:
        return None
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers as H
    O = H.O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Ms = args.cpu_tokens
    prep = []
    for name, N, K, _ in LINEARS:
        sp = split_for(K)
        idx = H.make_index(K, seed=0)
        x = H.make_activations(Ms, K, idx)
        w = H.make_weights(N, K)
        b = O.reorder_quantize(H.bits(w), idx.numpy(), *sp, "w4")  # offline, like QLinearLayer.__init__
        wd = [torch.from_numpy(O.dequant(b[i], b[3 + i], N, sp[i], 4)) for i in range(3)]
        prep.append((H.bits(x), idx.numpy(), sp, wd, N))

    def step():
        for xb, idx, sp, wd, N in prep:
            a = O.reorder_quantize(xb, idx, *sp, "x")
            acc = None
            for i, f in enumerate((4, 6, 8)):
                if sp[i] == 0:
                    continue
                ad = torch.from_numpy(O.dequant(a[i], a[3 + i], Ms, sp[i], f))
                part = ad @ wd[i].T
                acc = part if acc is None else acc + part
            acc.to(torch.bfloat16)

    for _ in range(args.warmup_ref):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps_ref):
        step()
    dt = (time.perf_counter() - t0) / args.steps_ref
    tflops = Ms * flops_per_token() / dt / 1e12
    sample = f"{Ms} of {args.tokens} tokens per step through all four linears, weights pre-dequantised"
    line = {"impl": "reference", "metric": METRIC, "value": tflops, "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "mxfp4/6/8 (fake-quant in fp32)", "data": "synthetic",
            "tokens_per_s": Ms / dt,
            "config": workload_config(args, world),
            "cpu_baseline": {"value": tflops, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tflops, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if emit_line:
        emit(line)
    return line


def workload_config(args, world):
    return {"workload": f"Llama-3-8B decoder-layer linears qkv(6144x4096) o(4096x4096) gate_up(28672x4096) "
                        f"down(4096x14336), M={args.tokens} tokens, split p4:p6:p8 = 5:2:1 (5.0 avg bits), "
                        f"quantize + mixed GEMM per linear",
            "tokens": args.tokens, "split_4096": list(split_for(4096)), "split_14336": list(split_for(14336)),
            "parallelism": f"tp{world}" if world > 1 else "single",
            "tp_reduce": (args.tp_reduce if world > 1 and args.impl == "ours" else None),
            "reference_arm_sample": (f"{args.cpu_tokens} of {args.tokens} tokens per step, {args.steps_ref} timed steps "
                                     "(a bounded CPU sample: same_steps is false by design)"
                                     if args.impl == "reference" else None),
            "tp_chunks": (args.tp_chunks if world > 1 and args.impl == "ours" else None),
            "options": (args.set_option or None),
            "l2": "inputs+weights+outputs per step (>1 GB) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------------ our arm
class HotLinear:
    """One (possibly TP-sharded) linear with every buffer preallocated; calls the C ABI directly."""

    def __init__(self, name, N, K, mode, M, rank, world, dev, lib, seed, workspace=None, share=None, chunk=0):
        import torch
        from micromix_b200 import mixedgemm
        from micromix_b200.parallel_utils import column_shard_range, row_shard_plan
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers as H
        self.name, self.mode, self.lib, self.M, self.world, self.seed = name, mode, lib, M, world, seed
        self.ws = workspace if (world > 1 and mode == "row") else None
        self._c_out = ctypes.c_void_p()
        if share is not None:
            # another token chunk of the same (sharded) linear: same quantized weights, permutation and split
            self.N, self.K, self.split, self.idx, self.W = share.N, share.K, share.split, share.idx, share.W
            p4, p6, p8 = self.split
        else:
            idx = H.make_index(K, seed=seed)
            g = torch.Generator(device=dev).manual_seed(1234 + seed)
            w = (torch.randn(N, K, generator=g, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
            p4, p6, p8 = split_for(K)
            if world > 1 and mode == "col":
                n0, n1 = column_shard_range(N, world, rank)
                w = w[n0:n1].contiguous()
            elif world > 1 and mode == "row":
                k0, k1, idx, p4, p6, p8 = row_shard_plan(idx, p6, p8, world, rank)
                w = w[:, k0:k1].contiguous()
            self.N, self.K = w.shape
            self.split = (p4, p6, p8)
            self.idx = idx.to(dev)
            self.W = mixedgemm.reorder_quantize_w4(w, self.idx, p4, p6, p8)
            del w
        gx = torch.Generator(device=dev).manual_seed(721 + seed + 97 * rank * (mode == "row") + 7919 * chunk)
        gain = 1.0 + 31.0 * (torch.arange(self.K, device=dev, dtype=torch.float32) / self.K) ** 8
        x = torch.randn(M, self.K, generator=gx, device=dev, dtype=torch.float32)
        xg = torch.empty_like(x)
        xg[:, self.idx.long()] = x * gain
        self.x = xg.to(torch.bfloat16)
        del x, xg
        u8 = dict(dtype=torch.uint8, device=dev)
        self.A = [torch.empty((M, w_), **u8) for w_ in (p4 // 2, p6 // 4 * 3, p8)]
        self.SFA = [torch.empty((int(lib.mmx_sf_bytes_act(M, k)),), **u8) for k in (p4, p6, p8)]
        self.out = torch.empty((M, self.N), dtype=torch.bfloat16, device=dev)
        self.flops = 2.0 * M * self.N * self.K
        self.qbytes = 2.0 * M * self.K + M * (p4 / 2 + p6 * 3 / 4 + p8) + M * self.K / 32
        # GEMM: packed A + scales, MXFP4 B + scales, bf16 C -- each touched once
        self.gbytes = M * (p4 / 2 + p6 * 3 / 4 + p8) + M * self.K / 32 + self.N * self.K / 2 + self.N * self.K / 32 + 2.0 * M * self.N
        p = lambda t: t.data_ptr() if t.numel() else None
        self._qargs = (p(self.x), M, self.K, p(self.idx), p4, p6, p8, p(self.A[0]), p(self.A[1]), p(self.A[2]),
                       p(self.SFA[0]), p(self.SFA[1]), p(self.SFA[2]))
        W = self.W
        self._margs = (p(self.A[0]), p(W[0]), p(self.A[1]), p(W[1]), p(self.A[2]), p(W[2]), p(self.SFA[0]), p(W[3]),
                       p(self.SFA[1]), p(W[4]), p(self.SFA[2]), p(W[5]), M, self.N, p4, p6, p8, 1, None, p(self.out))

    def enable_sp(self, ws):
        """Sequence-parallel step (tp_mode "sp"): a column-parallel linear quantizes only THIS rank's rows of X and
        multicasts the packed codes into every rank's gather channel (its GEMM waits per source rank); a row-parallel
        linear ends in a reduce-scatter (this rank keeps its rows of the sum)."""
        self.sp_ws = ws
        p4, p6, p8 = self.split
        if self.mode == "col":
            lo, hi = ws.shard_range(self.M)
            self.x_shard = self.x[lo:hi].contiguous()
            self._views = (ctypes.c_void_p * 6)()
            pq = lambda t: t.data_ptr() if t.numel() else self.idx.data_ptr()
            self._gargs = (ws.ctx, pq(self.x_shard), self.M, self.K, self.idx.data_ptr(), p4, p6, p8, None, 0.0, self._views)
            p = lambda t: t.data_ptr() if t.numel() else None
            W = self.W
            self._mgargs = (ws.ctx, p(W[0]), p(W[1]), p(W[2]), p(W[3]), p(W[4]), p(W[5]), self.M, self.N, p4, p6, p8, 1,
                            None, self.out.data_ptr())
        else:
            self._row0, self._rows = ctypes.c_int64(), ctypes.c_int64()

    def enable_tpr(self, ws, rank, world):
        """Token-parallel form of a row-parallel linear (tp_mode "tpr", the alternative design): replicated MXFP4 weight
        in the rank-blocked channel order, all-to-all of the packed activation codes, full-K GEMM on this rank's rows."""
        import torch
        from micromix_b200 import mixedgemm
        from micromix_b200.parallel_utils import token_parallel_plan
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers as H
        if self.mode != "row":
            return self.enable_sp(ws)
        self.sp_ws = None
        if getattr(self, "tpr", None) is None:
            dev = self.x.device
            Kf = self.K * world
            idx = H.make_index(Kf, seed=self.seed)
            g = torch.Generator(device=dev).manual_seed(1234 + self.seed)  # the same weight HotLinear sharded at construction
            w = (torch.randn(self.N, Kf, generator=g, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
            _, p6, p8 = split_for(Kf)
            perm, tot, shards = token_parallel_plan(idx, p6, p8, world)
            me = shards[rank]
            assert tuple(me["split"]) == tuple(self.split) and torch.equal(me["index"].to(dev), self.idx)
            Wt = mixedgemm.reorder_quantize_w4(w, perm.to(dev), *tot)
            del w
            lo, hi = ws.shard_range(self.M)
            out = torch.empty((max(hi - lo, 1), self.N), dtype=torch.bfloat16, device=dev)
            self.tpr = dict(W=Wt, tot=tot, off=me["offset"], perm=perm.to(dev), out=out, lo=lo, hi=hi,
                            c_tot=(ctypes.c_int32 * 3)(*tot), c_off=(ctypes.c_int32 * 3)(*me["offset"]))
            p = lambda t: t.data_ptr() if t.numel() else None
            p4, p6l, p8l = self.split
            self._xargs = (ws.ctx, self.x.data_ptr(), self.M, self.K, self.idx.data_ptr(), p4, p6l, p8l, self.tpr["c_tot"],
                           self.tpr["c_off"], None)
            self._row0, self._rows = ctypes.c_int64(), ctypes.c_int64()
            self._yargs = (ws.ctx, p(Wt[0]), p(Wt[1]), p(Wt[2]), p(Wt[3]), p(Wt[4]), p(Wt[5]), self.M, self.N, tot[0], tot[1],
                           tot[2], 1, None, out.data_ptr(), ctypes.byref(self._row0), ctypes.byref(self._rows))
        self.tpr_on = True

    def run(self, stream, events=None):
        """quantize + GEMM of this rank's shard (fused mode: + the reduction, inside the GEMM launch pair)."""
        lib = self.lib
        if events is not None:
            events[0].record()
        sp = getattr(self, "sp_ws", None)
        if getattr(self, "tpr_on", False) and self.mode == "row":
            rc = lib.mmx_tp_quantize_alltoall(*self._xargs, stream)
            if events is not None:
                events[1].record()
            rc |= lib.mmx_tp_matmul_exchanged(*self._yargs, stream)
            if events is not None:
                events[2].record()
            if rc:
                raise RuntimeError(lib.mmx_last_error().decode())
            return
        if sp is not None and self.mode == "col":
            rc = lib.mmx_tp_quantize_allgather(*self._gargs, stream)
            if events is not None:
                events[1].record()
            rc |= lib.mmx_tp_matmul_gathered(*self._mgargs, stream)
        else:
            rc = lib.mmx_reorder_quantize_x(*self._qargs, stream)
            if events is not None:
                events[1].record()
            if sp is not None:
                rc |= lib.mmx_matmul_reduce_scatter(sp.ctx, *self._margs[:-1], ctypes.byref(self._c_out),
                                                    ctypes.byref(self._row0), ctypes.byref(self._rows), stream)
            elif self.ws is not None:
                # fused: GEMM epilogue pushes partial tiles to their owner rank over NVLink, co-resident reducer kernel
                rc |= lib.mmx_matmul_allreduce(self.ws.ctx, *self._margs[:-1], ctypes.byref(self._c_out), stream)
            else:
                rc |= lib.mmx_matmul(*self._margs, stream)
        if events is not None:
            events[2].record()
        if rc:
            raise RuntimeError(lib.mmx_last_error().decode())

    def result(self):
        """The linear's output as this rank holds it after run(): (tensor, row0) -- row-parallel fused outputs live in
        the peer workspace (all rows, or this rank's rows under sequence parallelism)."""
        import torch
        from micromix_b200.parallel_utils import _DeviceBytes
        sp = getattr(self, "sp_ws", None)
        if getattr(self, "tpr_on", False) and self.mode == "row":
            return self.tpr["out"][: self.tpr["hi"] - self.tpr["lo"]], self.tpr["lo"]
        if self.mode == "row" and (sp is not None or self.ws is not None):
            rows = int(self._rows.value) if sp is not None else self.M
            row0 = int(self._row0.value) if sp is not None else 0
            t = torch.as_tensor(_DeviceBytes(self._c_out.value, rows * self.N * 2, "|u1"), device=self.out.device)
            return t.view(torch.bfloat16).view(rows, self.N), row0
        return self.out, 0

    def reduce_async(self):
        """Row-parallel partial sum over the ranks: NCCL all-reduce on NCCL's own stream (returns the Work to wait on
        before the result is consumed).  Nothing to do for column-parallel linears and in fused mode."""
        import torch.distributed as dist
        if self.ws is None and self.world > 1 and self.mode == "row":
            return dist.all_reduce(self.out, async_op=True)
        return None


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from micromix_b200 import _lib, mixedgemm
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: micromix_b200 has no CPU path (use --impl reference for "
                           "the host-core baseline)")
    lib = _lib.load()
    for kv in args.set_option:
        k, v = kv.split("=")
        if lib.mmx_set_option(k.encode(), int(v)):
            raise ValueError(lib.mmx_last_error().decode())
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_note = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else "off"
    M = args.tokens
    peaks = load_peaks()
    # token chunks: at N > 1 the step runs as C micro-batches so that the all-reduce of one chunk's row-parallel
    # linear (NCCL stream) overlaps the other chunks' quantize + GEMM; every chunk keeps a real layer's dependency
    # chain qkv -> o -> all-reduce -> gate_up -> down -> all-reduce (gate_up of chunk c waits for ITS o all-reduce)
    C = max(1, args.tp_chunks) if world > 1 else 1
    if M % (C * 128):
        C = 1
    Mc = M // C
    ws, ws_note = None, None
    tp_mode = "none"
    if world > 1 and args.tp_reduce == "fused":
        from micromix_b200.parallel_utils import PeerWorkspace
        n_row = max(N for _, N, _, mode in LINEARS if mode == "row")
        k_col = max(K for _, _, K, mode in LINEARS if mode == "col")
        # (the token-parallel alternative exchanges [M / tp, K] activations of the ROW linears through the same channel)
        k_col = max(k_col, -(-max(K for _, _, K, mode in LINEARS if mode == "row") // world // 128) * 128)
        want_sp = args.tp_mode in ("auto", "sp", "tpr") and C == 1
        try:
            # sequence parallel needs the NVSwitch multicast mapping (gather channel); the constructor is collective and
            # raises on EVERY rank when any rank cannot map it
            ws = PeerWorkspace(Mc, n_row, device=dev, gather=(Mc, k_col) if want_sp else None)
            tp_mode = ("tpr" if args.tp_mode == "tpr" else "sp") if want_sp else "ar"
        except Exception as e:  # noqa: BLE001
            ws_note = f"sequence-parallel workspace unavailable ({e!r})"[:200]
            if args.tp_mode in ("sp", "tpr"):
                raise
            try:
                ws = PeerWorkspace(Mc, n_row, device=dev)
                tp_mode = "ar"
            except Exception as e2:  # noqa: BLE001 -- no peer mapping on this box: every rank must agree on the fallback
                ws_note = f"peer workspace unavailable ({e2!r})"[:200]
        ok = torch.tensor([1 if ws is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not int(ok.item()):
            if ws is not None:
                ws.close()
            ws, args.tp_reduce, tp_mode = None, "nccl", "ar"
            ws_note = ws_note or "peer workspace unavailable on another rank"
    elif world > 1:
        tp_mode = "ar"
    if world > 1 and args.gemm_ctas > 0:
        lib.mmx_set_option(b"gemm_ctas", args.gemm_ctas)  # leave SMs to the concurrent NCCL kernel
    chunks = []
    for c in range(C):
        chunks.append([HotLinear(n, N, K, mode, Mc, rank, world, dev, lib, seed=i, workspace=ws,
                                 share=(chunks[0][i] if c else None), chunk=c)
                       for i, (n, N, K, mode) in enumerate(LINEARS)])
    lins = chunks[0]
    total_flops = M * flops_per_token()  # whole job, all ranks together
    mx_peak = measure_mx_peak(lib)

    def run_mode(tp_mode):
        """Everything that is timed, for one tensor-parallel mode: warm-up, real-rank parity, the K timed steps, the evented
        per-kernel pass, the end-to-end leg.  Returns the JSON line of that mode."""
        for l in lins:
            l.tpr_on = False
            if tp_mode == "sp":
                l.enable_sp(ws)
            elif tp_mode == "tpr":
                l.enable_tpr(ws, rank, world)
            else:
                l.sp_ws = None

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def issue_step(ev=None):
            """One pass of the hot path over the batch.  ev[c][li] = three events around quantize / GEMM."""
            stream = torch.cuda.current_stream().cuda_stream
            if world == 1:
                for li, l in enumerate(lins):
                    l.run(stream, ev[0][li] if ev is not None else None)
                return
            pend = [None] * C
            for half in ((0, 1), (2, 3)):  # attention linears, MLP linears
                for c in range(C):
                    if pend[c] is not None:
                        pend[c].wait()  # this chunk's previous all-reduce: overlapped the other chunks' kernels
                        pend[c] = None
                    for li in half:
                        chunks[c][li].run(stream, ev[c][li] if ev is not None else None)
                    pend[c] = chunks[c][half[1]].reduce_async()
            for c in range(C):
                if pend[c] is not None:
                    pend[c].wait()

        def new_events():
            return [[[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in LINEARS] for _ in range(C)]

        for _ in range(max(args.warmup, 3)):
            issue_step()
        barrier()
        tp_parity = check_tp_parity(lins, lib, ws, tp_mode, rank, world, dev) if world > 1 else None
        # N > 1: the step is replayed from a CUDA graph (8C kernels + 2C NCCL all-reduces per replay) -- at tp=8 the
        # kernels are 10-70 us each and eager launches from 8 Python processes would bound the step
        graph, graph_note = None, None
        if world > 1 and not args.no_graph:
            try:
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    issue_step()
                torch.cuda.current_stream().wait_stream(side)
                barrier()
                graph = torch.cuda.CUDAGraph()
                l0 = mixedgemm.launch_count()
                with torch.cuda.graph(graph, stream=side):
                    issue_step()
                launches_per_step = mixedgemm.launch_count() - l0
                barrier()
                for _ in range(3):
                    graph.replay()
                barrier()
            except Exception as e:  # noqa: BLE001 -- report, then time the eager step instead
                graph, graph_note = None, f"capture failed: {e!r}"[:200]
                torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        sampler.start()
        n_ev = args.steps if graph is None else max(3, min(args.steps, 10))
        ev = [new_events() for _ in range(n_ev)]
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = mixedgemm.launch_count()
        barrier()
        sampler.active.set()
        t_start.record()
        if graph is not None:
            for s in range(args.steps):
                graph.replay()
        else:
            for s in range(args.steps):
                issue_step(ev[s])
        t_end.record()
        barrier()
        if graph is None:
            sampler.active.clear()  # (graph mode: keep sampling through the evented pass of the same step below)
        launches = (mixedgemm.launch_count() - launches0) if graph is None else launches_per_step * args.steps
        ms_total = t_start.elapsed_time(t_end)
        ms_step = ms_total / args.steps
        if world > 1:
            t = torch.tensor([ms_step], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_step = float(t.item())
        if graph is not None:
            # per-kernel device times cannot be taken inside a graph replay: an eager, evented pass of the same step
            for s in range(n_ev):
                issue_step(ev[s])
            barrier()
            sampler.active.clear()
        # per-kernel device time from the events (N = 1: inside the timed region itself)
        cells = [(s, c, li) for s in range(n_ev) for c in range(C) for li in range(len(LINEARS))]
        q_ms = sum(ev[s][c][li][0].elapsed_time(ev[s][c][li][1]) for s, c, li in cells)
        g_ms = sum(ev[s][c][li][1].elapsed_time(ev[s][c][li][2]) for s, c, li in cells)
        per_lin = {}
        for li, l in enumerate(lins):
            gq = sum(ev[s][c][li][0].elapsed_time(ev[s][c][li][1]) for s in range(n_ev) for c in range(C)) / (n_ev * C)
            gg = sum(ev[s][c][li][1].elapsed_time(ev[s][c][li][2]) for s in range(n_ev) for c in range(C)) / (n_ev * C)
            per_lin[l.name] = {"M": l.M, "N": l.N, "K": l.K, "quant_us": gq * 1e3, "gemm_us": gg * 1e3,
                               "gemm_includes_allreduce": l.ws is not None,
                               "quant_gbs": l.qbytes / gq / 1e6, "gemm_tflops": l.flops / gg / 1e9}
        # the GEMM roofline counts launches that are GEMMs only: in fused mode a row-parallel "GEMM" interval is
        # GEMM + all-reduce (reported per linear, not against the tensor peak)
        pure = [li for li, l in enumerate(lins) if l.ws is None]
        g_pure_ms = sum(ev[s][c][li][1].elapsed_time(ev[s][c][li][2]) for s, c, li in cells if li in pure)
        pure_flops = C * sum(lins[li].flops for li in pure)
        gemm_tflops = pure_flops * n_ev / g_pure_ms / 1e9
        # split-weighted tensor peak from the MEASURED MX issue rates of this GPU (mmx_debug_mma_peak): the FP4 segment runs as
        # kind::mxf4, the FP6 / FP8 segments as kind::mxf8f6f4.  The GEMMs are timed inside a long step -> sustained figures.
        sus = tuple(mx_peak[f"{n}_sustained_tflops"] for n in ("mxf4", "mxf6", "mxf8"))
        bur = tuple(mx_peak[f"{n}_burst_tflops"] for n in ("mxf4", "mxf6", "mxf8"))

        def tmin_of(l, a):
            return 2.0 * l.M * l.N * (l.split[0] / a[0] + l.split[1] / a[1] + l.split[2] / a[2])

        tmin = C * sum(tmin_of(lins[li], sus) for li in pure)
        peak_eff = pure_flops / tmin  # TFLOP/s
        peak_burst = pure_flops / (C * sum(tmin_of(lins[li], bur) for li in pure))
        p_bf16 = peaks["bf16_tflops"]
        peak_proxy = pure_flops / (C * sum(tmin_of(lins[li], (4 * p_bf16, 2 * p_bf16, 2 * p_bf16)) for li in pure))
        for li in pure:
            l = lins[li]
            per_lin[l.name]["gemm_roofline_frac"] = per_lin[l.name]["gemm_tflops"] / (l.flops / tmin_of(l, sus))
            per_lin[l.name]["gemm_roofline_frac_vs_burst_peak"] = per_lin[l.name]["gemm_tflops"] / (l.flops / tmin_of(l, bur))
        for l in lins:
            per_lin[l.name]["quant_roofline_frac"] = per_lin[l.name]["quant_gbs"] / peaks["hbm_gbs"]
        quant_gbs = C * sum(l.qbytes for l in lins) * n_ev / q_ms / 1e6
        # DRAM bytes per launch from the ncu --set full capture of this same command (profiles/, tools/ncu_traffic.py), next to
        # the algorithmic bytes of each launch; N = 1 only (the capture is a one-GPU run)
        traffic = traffic_q = None
        prof = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if world == 1 and os.path.exists(prof):
            try:
                tj = json.load(open(prof))
                for l in lins:
                    g = tj.get("gemm", {}).get(l.name)
                    if g:
                        per_lin[l.name]["gemm_dram_bytes"] = g["dram_bytes"]
                        per_lin[l.name]["gemm_algorithmic_bytes"] = l.gbytes
                    qd = tj.get("quantize", {}).get(l.name)
                    if qd:
                        per_lin[l.name]["quant_dram_bytes"] = qd["dram_bytes"]
                        per_lin[l.name]["quant_algorithmic_bytes"] = l.qbytes
                if all("gemm_dram_bytes" in per_lin[l.name] for l in lins):
                    traffic = sum(per_lin[l.name]["gemm_dram_bytes"] for l in lins)
                if all("quant_dram_bytes" in per_lin[l.name] for l in lins):
                    traffic_q = sum(per_lin[l.name]["quant_dram_bytes"] for l in lins)
            except Exception:
                traffic = traffic_q = None

        # ---- e2e: the plugin call a user makes (QLinearLayer.forward) with HOST buffers, copies inside the timed region
        e2e = None if args.no_e2e else measure_e2e(args, rank, world, dev, chunks, total_flops)

        clocks = sampler.summary()
        sampler.stop_flag.set()
        tp_status, ws_mode = None, None
        if ws is not None:
            ws_mode = ws.mode
            tp_status = ws.status()  # 0 = no reducer wait ever timed out
        line = {"metric": METRIC, "value": total_flops / ms_step / 1e9, "unit": "TFLOP/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "mxfp4/mxfp6/mxfp8 x mxfp4 -> fp32 acc -> bf16",
                "data": "synthetic", "tokens_per_s": M / ms_step * 1e3, "config": workload_config(args, world),
                "gpu_launches": int(launches), "e2e": e2e, "clocks": clocks, "host_numa": numa_note,
                "roofline": {"kernel": "mixed_gemm_kernel", "bound": "tensor", "achieved": gemm_tflops, "peak": peak_eff,
                             "unit": "TFLOP/s", "frac": gemm_tflops / peak_eff, "traffic": traffic,
                             "traffic_note": "DRAM bytes of the step's GEMM launches (sum; per launch under per_linear) from the "
                                             "ncu capture in profiles/r02_traffic.json" if traffic else None,
                             "algorithmic_bytes": sum(l.gbytes for l in lins) if traffic else None,
                             "launches_counted": [lins[li].name for li in pure],
                             "frac_vs_burst_peak": gemm_tflops / peak_burst, "frac_vs_bf16_proxy": gemm_tflops / peak_proxy,
                             "mx_peak_measured": mx_peak,
                             "peak_note": "split-weighted over the linears' (p4, p6, p8) from the MEASURED sustained issue rates "
                                          "of kind::mxf4 / kind::mxf8f6f4 (E3M2, E4M3) on this GPU (mmx_debug_mma_peak: 0.45 s of "
                                          "back-to-back MMAs on smem-resident operands, power-capped clocks; the same probe as a "
                                          f"0.5 ms burst gives {peak_burst:.0f}); the round-1 proxy (4x / 2x the "
                                          f"{peaks['source']} cuBLAS bf16 burst {p_bf16}) would give {peak_proxy:.0f} TFLOP/s"},
                "roofline_quantize": {"kernel": "reorder_quantize_kernel", "bound": "hbm", "achieved": quant_gbs,
                                      "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": quant_gbs / peaks["hbm_gbs"],
                                      "traffic": traffic_q,
                                      "algorithmic_bytes": sum(l.qbytes for l in lins) if traffic_q else None,
                                      "peak_note": f"{peaks['source']} copy bandwidth"},
                "share": {"gemm": g_ms / ms_total if world == 1 else None, "quantize": q_ms / ms_total if world == 1 else None},
                "per_linear": per_lin}
        if tp_status is not None:
            line["tp_fused_status"] = tp_status
        if tp_parity is not None:
            line["tp_parity"] = tp_parity
        if world > 1:
            # what the row-parallel collectives put on NVLink, per rank and call, against the measured peer-copy rate
            wire = {}
            for l in lins:
                if l.mode != "row":
                    continue
                full = l.M * l.N * 2.0
                if tp_mode == "tpr":
                    p4_, p6_, p8_ = l.split
                    egress = (world - 1) / world * l.M * (p4_ / 2 + p6_ * 3 / 4 + p8_ + l.K / 32)
                    note = "all-to-all of packed MX codes + scales of this rank's K slice (no partial sums on the wire)"
                elif tp_mode == "sp":
                    egress = full * (world - 1) / world  # partial rows pulled by / pushed to their owners
                    note = "reduce-scatter: (tp-1)/tp of the bf16 partial leaves each rank, 1/tp of the sum comes back"
                elif ws is not None and ws.mode == "switch":
                    egress = full * (world - 1) / world + full / world
                    note = "in-switch all-reduce: partials out + own result tiles multicast"
                else:
                    egress = 2.0 * full * (world - 1) / world
                    note = "all-reduce: partials to owners + results to every rank"
                us = per_lin[l.name]["gemm_us"]
                wire[l.name] = {"egress_bytes_per_rank": egress, "gemm_plus_collective_us": us,
                                "egress_gbs_over_that_interval": egress / us / 1e3, "note": note}
            line["tp_wire"] = {"per_linear": wire, "peer_copy_peak_gbs": 770.0,
                               "peak_note": "measured peer copy per direction on this pool (B200_PROFILING.md); round 1 "
                                            "measured 744 GB/s (profiles/r01_nvlink_p2p_bw.log)"}
        if world > 1:
            line["tp"] = {"reduce": args.tp_reduce, "mode": tp_mode, "fused_mode": ws_mode, "fallback_note": ws_note,
                          "mode_note": {"sp": "sequence parallel: row-parallel GEMMs end in a reduce-scatter, column-parallel "
                                              "GEMMs start from an NVSwitch-multicast all-gather of the packed MX codes that each "
                                              "rank quantized for ITS rows", "ar": "row-parallel GEMM -> all-reduce, every rank "
                                              "quantizes the replicated activation"}.get(tp_mode),
                          "chunks": C, "cuda_graph": graph is not None, "graph_note": graph_note,
                          "gemm_ctas": args.gemm_ctas or None,
                          "per_kernel_times": "eager evented pass after the timed region" if graph is not None
                          else "events inside the timed region"}
        if world == 1 and not args.no_prefill:
            try:
                line["prefill"] = measure_prefill(args, dev)
                line["prefill_tokens_per_s"] = line["prefill"]["tokens_per_s"]
            except Exception as e:  # noqa: BLE001 -- the headline line must survive a failure of the secondary leg
                line["prefill"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_moe:
            try:
                line["moe"] = measure_moe(args, dev)
                line["moe_tokens_per_s"] = line["moe"]["tokens_per_s"]
            except Exception as e:  # noqa: BLE001 -- the headline line must survive a failure of the secondary leg
                line["moe"] = {"error": repr(e)[:300]}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        return line

    # N > 1 with the fused reduction: the row-parallel design has two forms here -- GEMM -> all-reduce ("ar") and the
    # sequence-parallel GEMM -> reduce-scatter + all-gather of packed codes ("sp").  `--tp-mode auto` measures BOTH (each its
    # own K timed steps) and reports the faster one as the line, the other under "tp_modes_measured".
    modes = [tp_mode]
    if world > 1 and ws is not None and tp_mode == "sp" and args.tp_mode == "auto":
        modes = ["ar", "sp", "tpr"]
    lines = {m: run_mode(m) for m in modes}
    # the LINE is the faster of the two forms of the north_star's row-parallel design; the token-parallel alternative
    # ("tpr": replicated o / down weights, all-to-all of packed codes) is measured the same way and reported NEXT to it
    best = min((m for m in lines if m != "tpr" or len(lines) == 1), key=lambda m: lines[m]["ms_per_step"])
    line = lines[best]
    if "tpr" in lines and best != "tpr":
        t = lines["tpr"]
        line["alt_token_parallel"] = {
            "what": "o / down token-parallel: replicated MXFP4 weights, rank-local quantize of the K slice, all-to-all of "
                    "the packed codes, full-K GEMM on each rank's token rows; qkv / gate_up as in the sequence-parallel form",
            "ms_per_step": t["ms_per_step"], "value": t["value"], "unit": "TFLOP/s", "tp_parity": t.get("tp_parity"),
            "e2e": t.get("e2e"), "tp_wire": t.get("tp_wire"),
            "per_linear_us": {k: [round(v["quant_us"], 1), round(v["gemm_us"], 1)] for k, v in t["per_linear"].items()}}
    if len(lines) > 1:
        line["tp_modes_measured"] = {m: {"ms_per_step": l["ms_per_step"], "value": l["value"],
                                         "e2e_value": (l.get("e2e") or {}).get("value"),
                                         "e2e_ms_per_step": (l.get("e2e") or {}).get("ms_per_step"),
                                         "tp_parity_ok": (l.get("tp_parity") or {}).get("ok"),
                                         "per_linear_us": {k: [round(v["quant_us"], 1), round(v["gemm_us"], 1)]
                                                           for k, v in l["per_linear"].items()}}
                                     for m, l in lines.items()}
        line["tp_mode_choice"] = f"{best}: the faster of the modes measured in this run (each timed over its own {args.steps} steps)"
    if ws is not None:
        ws.close()
    if rank == 0:
        emit(line)


def measure_mx_peak(lib):
    """The MEASURED dense MX tensor-pipe peak of this GPU (mmx_debug_mma_peak: every SM issues back-to-back block-scaled
    tcgen05 MMAs on shared-memory-resident operands): burst = a ~0.5 ms kernel, sustained = a ~5 ms kernel under the power
    cap.  kind::mxf4 carries the FP4 segment, kind::mxf8f6f4 the FP6 and FP8 segments (same rate for both)."""
    out = {}
    for name, kind in (("mxf4", 0), ("mxf6", 1), ("mxf8", 2)):
        # burst: best of 3 kernels of ~0.5 ms; sustained: 40 back-to-back kernels of ~11 ms (~0.45 s, power-capped clocks)
        for tag, stages, reps in (("burst", 2000, 3), ("sustained", 40000, -40)):
            t, ms = ctypes.c_double(), ctypes.c_double()
            rc = lib.mmx_debug_mma_peak(kind, stages, 0, reps, ctypes.byref(t), ctypes.byref(ms))
            if rc:
                raise RuntimeError(lib.mmx_last_error().decode())
            out[f"{name}_{tag}_tflops"] = t.value
    return out


def _bf16_ulps(a, b):
    """Largest distance in bf16 rounding steps between two bf16 tensors (ordered-integer view of the bit patterns)."""
    import torch
    ka, kb = a.contiguous().view(torch.int16).to(torch.int32), b.contiguous().view(torch.int16).to(torch.int32)
    ka = torch.where(ka < 0, -32768 - ka, ka)
    kb = torch.where(kb < 0, -32768 - kb, kb)
    return int((ka - kb).abs().max().item()) if ka.numel() else 0


def check_tp_parity(lins, lib, ws, tp_mode, rank, world, dev):
    """Real-rank parity of every tensor-parallel linear, BEFORE the timed region (the driver's scaling run must
    carry it).  Column-parallel: the product path (sequence parallel: rows quantized by their owner rank, packed codes
    multicast, gathered GEMM) must equal mmx_reorder_quantize_x + mmx_matmul on the full activation of THIS rank bit for
    bit.  Row-parallel: the fused GEMM -> all-reduce / reduce-scatter must equal bf16(sum over ranks, fp32, rank order) of
    the ranks' mmx_matmul partials -- exactly on the push data path, within one bf16 rounding step on the in-switch path
    (the NVSwitch fixes the summation order) -- and is also compared with mmx_matmul + ncclAllReduce."""
    import torch
    import torch.distributed as dist
    from micromix_b200 import mixedgemm
    stream = torch.cuda.current_stream().cuda_stream
    exact_expected = ws is not None and ws.mode == "push"  # (ncclAllReduce sums in its own order: within one step)
    out = {"mode": tp_mode, "data_path": (ws.mode if ws is not None else "nccl"), "linears": {}}
    ok_all = True
    for l in lins:
        l.run(stream)
        torch.cuda.synchronize()
        got, row0 = l.result()
        got = got.clone()
        if ws is None and l.mode == "row":
            dist.all_reduce(got)  # --tp-reduce nccl: the step's reduction is the separate ncclAllReduce
        if getattr(l, "tpr_on", False) and l.mode == "row":
            # token-parallel: the rows this rank holds must equal, bit for bit, the rows of the single-GPU linear over the
            # full K with the rank-blocked permutation (full activation = the ranks' K slices side by side)
            parts = [torch.empty_like(l.x) for _ in range(world)]
            dist.all_gather(parts, l.x)
            xf = torch.cat(parts, dim=1)
            del parts
            a = mixedgemm.reorder_quantize_x(xf, l.tpr["perm"], *l.tpr["tot"])
            W = l.tpr["W"]
            ref = mixedgemm.matmul(a[0], W[0], a[1], W[1], a[2], W[2], a[3], W[3], a[4], W[4], a[5], W[5])
            ulps = _bf16_ulps(got, ref[row0:row0 + got.shape[0]])
            del xf, a, ref
            ok = ulps == 0
            rec = {"kind": "token-parallel (all-to-all of packed codes, replicated weight)", "rows": [row0, row0 + got.shape[0]],
                   "max_bf16_steps_vs_single_gpu_full_K_linear": ulps, "bound": 0}
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            worst = torch.tensor([ulps], device=dev)
            dist.all_reduce(worst, op=dist.ReduceOp.MAX)
            rec["ok_on_all_ranks"] = bool(int(flag.item()))
            rec["worst_rank_bf16_steps"] = int(worst.item())
            ok_all = ok_all and rec["ok_on_all_ranks"]
            out["linears"][l.name] = rec
            continue
        a = mixedgemm.reorder_quantize_x(l.x, l.idx, *l.split)
        W = l.W
        part = mixedgemm.matmul(a[0], W[0], a[1], W[1], a[2], W[2], a[3], W[3], a[4], W[4], a[5], W[5])
        if l.mode == "col":
            ulps = _bf16_ulps(got, part)
            ok = ulps == 0
            rec = {"kind": "column-parallel", "rows": got.shape[0], "max_bf16_steps_vs_local_quantize+matmul": ulps}
        else:
            parts = [torch.empty_like(part) for _ in range(world)]
            dist.all_gather(parts, part)
            acc = parts[0].float()
            for q in parts[1:]:
                acc += q.float()
            ref = acc.to(torch.bfloat16)
            del parts, acc
            nccl = part.clone()
            dist.all_reduce(nccl)
            rows = got.shape[0]
            ulps = _bf16_ulps(got, ref[row0:row0 + rows])
            ulps_nccl = _bf16_ulps(got, nccl[row0:row0 + rows])
            ok = ulps == 0 if exact_expected else ulps <= 1
            if ws is None:
                # --tp-reduce nccl (the baseline arm): ncclAllReduce rounds to bf16 at every hop of its ring / tree, so near
                # a cancellation its sum can sit on the other side of zero -- bounded in absolute terms instead of in steps
                err = float((got.float() - ref[row0:row0 + rows].float()).abs().max() / ref.float().pow(2).mean().sqrt())
                ok = err <= 1e-1  # (measured at tp = 4: a few per cent of the rms at the worst of 33 M elements)
            rec = {"kind": "row-parallel", "rows": [row0, row0 + rows], "max_bf16_steps_vs_fp32_rank_order_sum": ulps,
                   "max_bf16_steps_vs_matmul+ncclAllReduce": ulps_nccl,
                   "bound": "max |diff| <= 1e-1 rms (NCCL arm)" if ws is None else (0 if exact_expected else 1)}
            if ws is None:
                rec["max_abs_diff_over_rms_vs_fp32_rank_order_sum"] = err
            if tp_mode != "sp":  # all-reduce: every rank must hold the same bits
                h = got.view(torch.int16).to(torch.int64).sum().reshape(1)
                hs = [torch.empty_like(h) for _ in range(world)]
                dist.all_gather(hs, h)
                rec["identical_on_all_ranks"] = all(int(x.item()) == int(h.item()) for x in hs)
                ok = ok and rec["identical_on_all_ranks"]
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        worst = torch.tensor([ulps], device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        rec["ok_on_all_ranks"] = bool(int(flag.item()))
        rec["worst_rank_bf16_steps"] = int(worst.item())
        ok_all = ok_all and rec["ok_on_all_ranks"]
        out["linears"][l.name] = rec
    out["ok"] = ok_all
    return out


def measure_e2e(args, rank, world, dev, chunks, total_flops):
    import torch
    import torch.distributed as dist
    import torch.nn as nn
    from micromix_b200 import mixedgemm
    from micromix_b200.qLinearLayer import QLinearLayer
    M = args.tokens
    lins = [l for ch in chunks for l in ch]  # every (token chunk, linear) of the step
    layers, xin, yout, yrows = [], [], [], []
    for l in lins:
        q = QLinearLayer.__new__(QLinearLayer)  # reuse the already-quantized shard instead of re-quantizing
        nn.Module.__init__(q)
        q.in_features, q.out_features, q.bias = l.K, l.N, None
        q.p4_num, q.p6_num, q.p8_num = l.split
        q.register_buffer("reorder_index", l.idx, persistent=False)
        q.BN, q.BS, q.BO, q.SFBN, q.SFBS, q.SFBO = l.W
        layers.append(q)
        sp = getattr(l, "sp_ws", None)
        if sp is not None and l.mode == "col":
            # sequence parallel: a rank uploads only ITS rows of the replicated activation (the packed codes of the other
            # rows arrive over NVLink) -- every activation byte crosses PCIe once per box, not once per rank
            lo, hi = sp.shard_range(l.M)
            xin.append(l.x[lo:hi].cpu().pin_memory())
        else:
            xin.append(l.x.cpu().pin_memory())
        # a row-parallel result: each rank returns its 1/N of the rows (sequence parallel: exactly the rows it owns; all-
        # reduce: the result is replicated and the host needs it once); a column-parallel shard is returned whole
        if (sp is not None or getattr(l, "tpr_on", False)) and l.mode == "row":
            r0, r1 = l.ws.shard_range(l.M)
        else:
            r0, r1 = (l.M * rank // world, l.M * (rank + 1) // world) if (world > 1 and l.mode == "row") else (0, l.M)
        yrows.append((r0, r1))
        yout.append(torch.empty((r1 - r0, l.N), dtype=torch.bfloat16).pin_memory())
    h2d = sum(x.numel() * 2 for x in xin)
    d2h = sum(y.numel() * 2 for y in yout)

    # three streams: H2D copies, the layers, D2H copies -- a step is bound by the slower PCIe direction, not their sum
    s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    # Steps are PIPELINED: the host enqueues step s+1 while step s is still copying / computing, and synchronises once at
    # the end of the timed region (every step's H2D and D2H copies are inside it).  Tensors that cross streams are handed to
    # the allocator with record_stream; a result that lives in the peer workspace (fused all-reduce / reduce-scatter) is
    # overwritten by the same linear's next call, so that call waits for the previous step's D2H of it.
    prev_out = {}

    def step():
        for i, (q, x, y, l, (r0, r1)) in enumerate(zip(layers, xin, yout, lins, yrows)):
            with torch.cuda.stream(s_in):
                xd = x.to(dev, non_blocking=True)
                e_in = torch.cuda.Event()
                e_in.record(s_in)
            xd.record_stream(s_run)
            s_run.wait_event(e_in)
            with torch.cuda.stream(s_run):
                sp = getattr(l, "sp_ws", None)
                tpr = getattr(l, "tpr_on", False) and l.mode == "row"
                in_ws = l.mode == "row" and (sp is not None or l.ws is not None) and not tpr
                if in_ws and i in prev_out:
                    s_run.wait_event(prev_out[i])
                if tpr:
                    l.ws.quantize_alltoall(xd, l.M, q.reorder_index, l.split, l.tpr["tot"], l.tpr["off"])
                    yd, _ = l.ws.matmul_exchanged(l.M, l.tpr["W"], l.tpr["tot"])
                    r0, r1 = 0, yd.shape[0]  # this rank's rows
                elif sp is not None and l.mode == "col":
                    sp.quantize_allgather(xd, l.M, q.reorder_index, q.p4_num, q.p6_num, q.p8_num)
                    yd = sp.matmul_gathered(l.M, l.W, q.p4_num, q.p6_num, q.p8_num)
                elif sp is not None:
                    a = mixedgemm.reorder_quantize_x(xd, q.reorder_index, q.p4_num, q.p6_num, q.p8_num)
                    yd, _ = sp.matmul_reduce_scatter(a, l.W)
                    r0, r1 = 0, yd.shape[0]  # already this rank's rows
                elif l.ws is not None:
                    a = mixedgemm.reorder_quantize_x(xd, q.reorder_index, q.p4_num, q.p6_num, q.p8_num)
                    yd = l.ws.matmul_allreduce(a, l.W)
                else:
                    yd = q(xd.view(1, l.M, -1))
                    if world > 1 and l.mode == "row":
                        dist.all_reduce(yd)
                e_run = torch.cuda.Event()
                e_run.record(s_run)
            if not in_ws:
                yd.record_stream(s_out)
            s_out.wait_event(e_run)
            with torch.cuda.stream(s_out):
                y.copy_(yd.reshape(-1, l.N)[r0:r1], non_blocking=True)
                e_out = torch.cuda.Event()
                e_out.record(s_out)
            prev_out[i] = e_out

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n = max(3, min(args.steps, 8))
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    # What bounds it: the same bytes over the same two copy streams with NO kernels in between (all ranks at once, so the
    # contention for the host's memory / PCIe root ports is the same as in the timed step).
    ydev = [torch.empty((y.shape[0], y.shape[1]), dtype=torch.bfloat16, device=dev) for y in yout]

    def copies_only():
        for x, y, yd in zip(xin, yout, ydev):
            with torch.cuda.stream(s_in):
                xd = x.to(dev, non_blocking=True)
            with torch.cuda.stream(s_out):
                y.copy_(yd, non_blocking=True)
            del xd

    copies_only()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        copies_only()
    torch.cuda.synchronize()
    dc = (time.perf_counter() - t0) / n
    if world > 1:
        t = torch.tensor([dc], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dc = float(t.item())
    if dc >= 0.75 * dt:
        bound = ("host<->device copies: the step's bytes alone (no kernels, all ranks copying at once) take %.2f ms of the "
                 "%.2f ms step" % (dc * 1e3, dt * 1e3))
    else:
        bound = ("not the copies (%.2f ms alone of a %.2f ms step): host-side enqueue and cross-rank launch skew"
                 % (dc * 1e3, dt * 1e3))
    return {"value": total_flops / dt / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "tokens_per_s": M / dt,
            "steps_timed": n, "copies_only_ms": dc * 1e3, "bound": bound,
            "pcie_gbps_per_rank": {"h2d": h2d / dt / 1e9, "d2h": d2h / dt / 1e9,
                                   "copies_only_h2d": h2d / dc / 1e9, "copies_only_d2h": d2h / dc / 1e9},
            "api": "QLinearLayer.forward on pinned host tensors; per linear: H2D copy, quantize, GEMM (+ reduction), D2H "
                   "copy, on three streams (copy-in / layers / copy-out), steps pipelined, one synchronize at the end of the "
                   "timed region; bytes are per rank.  Sequence-parallel mode: a rank uploads only ITS rows of a replicated "
                   "activation and returns only its rows of a row-parallel result (every byte crosses PCIe once per box)"}


def measure_moe(args, dev):
    """BASELINE config 5 in the driver's bench line (N = 1): the Mixtral-8x7B sparse-MoE block -- 8 experts, top-2, 16384
    tokens, random-init weights, synthetic calibration -- through QMixtralSparseMoeBlock(fused=True): routing kernel, ONE grouped
    quantize (token gather fused), ONE grouped GEMM over w1||w3 whose epilogue emits w2's quantized operand, ONE grouped GEMM
    over w2, combine kernel; the whole block replayed from one CUDA graph.  (Expert parallel over 8 GPUs:
    tools/bench_models.py mixtral_ep under torchrun, profiles/r02_mixtral_ep8_grouped_final.json.)"""
    import torch
    from micromix_b200 import mixedgemm
    from micromix_b200 import model_shapes as S
    from micromix_b200.qMixtralLayer import QMixtralSparseMoeBlock
    cfg = S.MIXTRAL_8X7B
    tokens = args.moe_tokens
    layer = S.make_layer(cfg, dev, seed=0, moe=True)
    idx, p6, p8 = S.make_calibration(cfg, 0, moe=True)
    blk = QMixtralSparseMoeBlock(layer.block_sparse_moe, p8, p6, idx, 0, fused=True)
    del layer
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev).manual_seed(721)
    x0 = torch.randn(1, tokens, cfg["hidden_size"], generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    with torch.no_grad():
        for _ in range(2):
            y = blk(x0)[0]
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(y.float()).all())
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    l0 = mixedgemm.launch_count()
    with torch.cuda.graph(graph, stream=side), torch.no_grad():
        blk(x0)
    launches = mixedgemm.launch_count() - l0
    torch.cuda.synchronize()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = 2.0 * tokens * cfg["num_experts_per_tok"] * 3 * cfg["hidden_size"] * cfg["intermediate_size"]
    out = {"config": f"Mixtral-8x7B sparse-MoE block (8 experts, top-2), {tokens} tokens, one GPU, random-init weights, "
                     "synthetic calibration", "path": "grouped" if blk.grouped else "expert loop",
           "act_epilogue": bool(getattr(blk, "act_epilogue", False)), "ms_per_block": ms, "tokens_per_s": tokens / ms * 1e3,
           "expert_tflops": flops / ms / 1e9, "mmx_launches_per_block": int(launches), "cuda_graph": True,
           "output_finite": finite}
    del graph, blk, y
    torch.cuda.empty_cache()
    return out


def measure_prefill(args, dev):
    """BASELINE config 3 -- the metric's "prefill tokens/s" half: ALL 32 Llama-3-8B decoder layers (random-init weights,
    synthetic reorder_index, 5-bit split), batch 8 x seq 2048 = 16384 tokens (/root/reference/prof_micromix.sh:1), through
    QLlamaDecoderLayer(fused=True) -- RMSNorm inside the quantizer, SiLU*up inside down_proj's quantizer -- replayed from
    ONE CUDA graph.  fused=True: SiLU*up + down_proj's quantizer run in the gate_up GEMM's epilogue, the residual adds in the
    o_proj / down_proj epilogues, RoPE in place on the qkv output (library kernel); attention is torch SDPA (cuDNN)."""
    import torch
    from micromix_b200 import mixedgemm
    from micromix_b200 import model_shapes as S
    from micromix_b200.qLlamaLayer import QLlamaDecoderLayer
    cfg = S.LLAMA3_8B
    n_layers, b, s = cfg["num_hidden_layers"], args.prefill_batch, args.prefill_seq
    layers = []
    for i in range(n_layers):
        layer = S.make_layer(cfg, dev, seed=i)
        idx, p6, p8 = S.make_calibration(cfg, i)
        layers.append(QLlamaDecoderLayer(layer, False, p8, p6, idx, i, fused=True))
        del layer
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev).manual_seed(721)
    x0 = torch.randn(b, s, cfg["hidden_size"], generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    pos = S.rope_tables(cfg, b, s, dev)

    @torch.no_grad()
    def fwd():
        x = x0
        for l in layers:
            x = l(x, position_embeddings=pos)[0]
        return x

    for _ in range(2):
        y = fwd()
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(y.float()).all())
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    l0 = mixedgemm.launch_count()
    with torch.cuda.graph(graph, stream=side):
        y = fwd()
    launches = mixedgemm.launch_count() - l0
    torch.cuda.synchronize()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    n = args.prefill_iters
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    tokens = b * s
    lin_flops = tokens * flops_per_token() * n_layers
    out = {"config": f"Llama-3-8B full prefill: {n_layers} decoder layers, batch {b} x seq {s} = {tokens} tokens, "
                     "random-init weights, synthetic reorder_index, 5.0-bit split", "layers_run": n_layers,
           "ms_per_prefill": ms, "tokens_per_s": tokens / ms * 1e3, "linear_tflops": lin_flops / ms / 1e9,
           "iters": n, "cuda_graph": True, "mmx_launches_per_prefill": int(launches), "output_finite": finite,
           "fused_norm_act": True,
           "note": "decoder layers only (no embedding / lm_head, as in the reference's layer-wise eval); attention = torch "
                   "SDPA (cuDNN); RoPE = the library's in-place kernel; SiLU*up + quantize and the residual adds run in GEMM epilogues; "
                   "one CUDA-graph replay per prefill"}
    del graph, layers, y
    torch.cuda.empty_cache()
    return out


def cpu_baseline(args):
    a = argparse.Namespace(**vars(args))
    a.steps_ref, a.warmup_ref = 2, 1
    return run_reference(a, 0, 1, emit_line=False)["cpu_baseline"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--tokens", type=int, default=8192)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-tokens", type=int, default=2048, help="token sample for the host-core baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tp-chunks", type=int, default=1,
                    help="N>1: token micro-batches per step; a chunk's all-reduce overlaps the other chunks' kernels")
    ap.add_argument("--gemm-ctas", type=int, default=0, help="N>1: cap the persistent GEMM grid (SMs left to NCCL)")
    ap.add_argument("--no-graph", action="store_true", help="N>1: launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (sweeps)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA node")
    ap.add_argument("--no-prefill", action="store_true", help="skip the 32-layer Llama-3-8B prefill leg (N = 1)")
    ap.add_argument("--no-moe", action="store_true", help="skip the Mixtral-8x7B MoE-block leg (N = 1)")
    ap.add_argument("--moe-tokens", type=int, default=16384)
    ap.add_argument("--prefill-batch", type=int, default=8)
    ap.add_argument("--prefill-seq", type=int, default=2048)
    ap.add_argument("--prefill-iters", type=int, default=5)
    ap.add_argument("--set-option", action="append", default=[], metavar="KEY=VALUE",
                    help="mmx_set_option(KEY, VALUE) before the run (tuning sweeps)")
    ap.add_argument("--tp-mode", default="auto", choices=["auto", "sp", "ar", "tpr"],
                    help="N>1, fused reduction: sp = sequence parallel (reduce-scatter + all-gather of packed codes through "
                         "NVSwitch multicast; the default when the box offers multicast memory), ar = all-reduce")
    ap.add_argument("--tp-reduce", default="fused", choices=["nccl", "fused"],
                    help="row-parallel reduction at N>1: our GEMM->all-reduce over NVLink peer / NVSwitch multicast memory "
                         "(default; falls back to NCCL, and says so, if the peer workspace cannot be mapped), or mmx_matmul "
                         "+ NCCL all-reduce")
    args = ap.parse_args()
    args.steps_ref = max(1, min(args.steps, 3))
    args.warmup_ref = max(1, min(args.warmup, 1))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "ours":
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    # stdout carries the JSON line and nothing else: native libraries print there too (NCCL's version banner), so fd 1
    # is pointed at stderr for the duration and the line goes to a private duplicate of the original stdout
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        # the host-core arm uses every core it can: torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin the
        # OpenMP loops of the C oracle (and MKL) to one thread -- undo that before any OpenMP runtime is loaded
        for var in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[var] = str(os.cpu_count() or 1)
        run_reference(args, rank, world)
        return
    if world > 1:
        from micromix_b200.parallel_utils import init_tensor_parallel
        init_tensor_parallel("nccl")
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    main()
