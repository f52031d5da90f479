#!/usr/bin/env python
"""GPU bring-up script (run on the B200 box through gpurun): stage-by-stage checks, each in its own process so a
fault in one stage cannot take the others down.  Writes a log and artefacts into gpurun_out/."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def stage_quant():
    import numpy as np
    import torch
    import helpers as H
    from micromix_b200 import mixedgemm
    O = H.O
    dev = torch.device("cuda:0")
    res = []
    cases = [(1, 4096, (2560, 1024, 512)), (7, 4096, (4096, 0, 0)), (127, 4096, (0, 0, 4096)),
             (129, 4096, (0, 4096, 0)), (300, 1024, (512, 256, 256)), (2048, 4096, (2560, 1024, 512)),
             (256, 14336, (8960, 3584, 1792)), (200, 27648, (17280, 6912, 3456)), (130, 128, (0, 128, 0))]
    for (M, K, (KN, KS, KO)) in cases:
        idx = H.make_index(K, seed=K)
        x = H.make_activations(M, K, idx)
        for mode, fn in (("x", mixedgemm.reorder_quantize_x), ("w4", mixedgemm.reorder_quantize_w4),
                         ("w", mixedgemm.reorder_quantize_w)):
            got = fn(x.to(dev), idx.to(dev), KN, KS, KO)
            torch.cuda.synchronize()
            ref = O.reorder_quantize(H.bits(x), idx.numpy(), KN, KS, KO, mode)
            ok = True
            detail = []
            for i in range(3):
                eq = np.array_equal(H.u8(got[i]), ref[i])
                ok &= eq
                if not eq:
                    g = H.u8(got[i]); bad = np.argwhere(g != ref[i])
                    detail.append(f"q{i}: {bad.shape[0]} bytes differ, first {bad[0].tolist()} got {g[tuple(bad[0])]} ref {ref[i][tuple(bad[0])]}")
            for i, k in zip(range(3, 6), (KN, KS, KO)):
                g = H.u8(got[i])
                mask = O.sf_valid_mask(M, k, g.shape[0])
                if g.shape != ref[i].shape:
                    ok = False; detail.append(f"sf{i} shape {g.shape} vs {ref[i].shape}")
                    continue
                eq = np.array_equal(g[mask], ref[i][mask])
                ok &= eq
                if not eq:
                    bad = np.argwhere((g != ref[i]) & mask)
                    detail.append(f"sf{i}: {bad.shape[0]} differ first off {bad[0].tolist()} got {g[bad[0][0]]} ref {ref[i][bad[0][0]]}")
            res.append((M, K, KN, KS, KO, mode, ok))
            print(f"quant M={M} K={K} split=({KN},{KS},{KO}) mode={mode}: {'OK' if ok else 'MISMATCH'} {' | '.join(detail)}", flush=True)
    # reference kernel (the reference's own reorder.cu compiled for sm_100a), if it travelled with the snapshot
    refso = os.path.join(ROOT, "oracle", "_ref", "libref_reorder.so")
    if os.path.exists(refso):
        import ctypes
        R = ctypes.CDLL(refso)
        R.ref_reorder_quantize.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6
        golden = {}
        for (M, K, (KN, KS, KO), tag) in [(128, 11008, (11008 - 1024, 1024 - 128, 128), "testpy"),
                                          (300, 4096, (2560, 1024, 512), "mixed"), (64, 3072, (1024, 1024, 1024), "thirds")]:
            idx = H.make_index(K, seed=1, identity=(tag == "testpy"))
            x = H.make_testpy_activations(M, K, KN, KS, KO) if tag == "testpy" else H.make_activations(M, K, idx)
            xd, idd = x.to(dev), idx.to(dev)
            for mode_i, mode in enumerate(("x", "w", "w4")):
                fm = (4, 4, 4) if mode == "w4" else (4, 6, 8)
                is_act = mode == "x"
                q = [torch.zeros((M, k * f // 8), dtype=torch.uint8, device=dev) for k, f in zip((KN, KS, KO), fm)]
                sf = [torch.zeros((O.sf_bytes(M, k, is_act),), dtype=torch.uint8, device=dev) for k in (KN, KS, KO)]
                torch.cuda.synchronize()
                rc = R.ref_reorder_quantize(mode_i, xd.data_ptr(), M, idd.data_ptr(), KN, KS, KO, *[t.data_ptr() for t in q], *[t.data_ptr() for t in sf])
                torch.cuda.synchronize()
                ours = {"x": mixedgemm.reorder_quantize_x, "w": mixedgemm.reorder_quantize_w, "w4": mixedgemm.reorder_quantize_w4}[mode](xd, idd, KN, KS, KO)
                orc = O.reorder_quantize(H.bits(x), idx.numpy(), KN, KS, KO, mode)
                ok_ours, ok_orc = True, True
                for i in range(3):
                    ok_ours &= bool(torch.equal(q[i], ours[i])); ok_orc &= np.array_equal(H.u8(q[i]), orc[i])
                for i, k in zip(range(3), (KN, KS, KO)):
                    mask = O.sf_valid_mask(M, k, sf[i].numel())
                    r_ = H.u8(sf[i]); ok_ours &= np.array_equal(r_[mask], H.u8(ours[3 + i])[:r_.shape[0]][mask]); ok_orc &= np.array_equal(r_[mask], orc[3 + i][mask])
                print(f"REFKERNEL rc={rc} tag={tag} mode={mode}: reference==ours {ok_ours}  reference==oracle {ok_orc}", flush=True)
                if tag != "testpy" or mode == "x":
                    for i in range(3):
                        golden[f"{tag}_{mode}_q{i}"] = H.u8(q[i]); golden[f"{tag}_{mode}_sf{i}"] = H.u8(sf[i])
        np.savez_compressed(os.path.join(OUT, "ref_reorder_golden.npz"), **golden)
        print("wrote golden", flush=True)
    else:
        print("REFKERNEL: oracle/_ref/libref_reorder.so not present", flush=True)
    # timing
    for (M, K) in [(2048, 4096), (16384, 4096), (16384, 14336)]:
        KN, KS, KO = H.SPLITS[K]
        idx = H.make_index(K).to(dev)
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        for rows in (0, 4, 2):
            from micromix_b200 import _lib
            _lib.load().mmx_set_option(b"quant_rows", rows)
            try:
                for _ in range(3):
                    mixedgemm.reorder_quantize_x(x, idx, KN, KS, KO)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    mixedgemm.reorder_quantize_x(x, idx, KN, KS, KO)
                e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 20 * 1e3
                byts = 2 * M * K + M * (KN // 2 + KS * 3 // 4 + KO) + M * K // 32
                print(f"quant timing M={M} K={K} rows={rows}: {us:.1f} us  {byts / us / 1e3:.1f} GB/s", flush=True)
            except Exception as e:
                print(f"quant timing M={M} K={K} rows={rows}: ERROR {e}", flush=True)
        _lib.load().mmx_set_option(b"quant_rows", 0)


def stage_gemm(tx_mode: int):
    import numpy as np
    import torch
    import helpers as H
    from micromix_b200 import mixedgemm, _lib
    O = H.O
    L = _lib.load()
    L.mmx_set_option(b"gemm_watchdog", 1)
    L.mmx_set_option(b"gemm_tx_mode", tx_mode)
    dev = torch.device("cuda:0")
    import ctypes
    def status():
        buf = (ctypes.c_uint32 * 8)()
        L.mmx_gemm_debug_status(buf, 8)
        return [hex(v) for v in buf]
    cases = [(128, 256, (256, 0, 0)), (128, 256, (0, 0, 128)), (128, 256, (0, 128, 0)), (128, 256, (128, 0, 0)),
             (256, 256, (256, 0, 0)), (256, 256, (0, 0, 128)), (256, 256, (0, 128, 0)), (256, 256, (128, 0, 0)),
             (256, 256, (512, 0, 0)), (256, 256, (0, 256, 0)), (256, 256, (0, 0, 256)), (129, 128, (384, 0, 0)),
             (256, 256, (256, 128, 128)), (200, 384, (384, 256, 128)), (1, 128, (128, 128, 128)),
             (300, 1024, (2560, 1024, 512)), (2048, 4096, (2560, 1024, 512))]
    allok = True
    for (M, N, (KN, KS, KO)) in cases:
        K = KN + KS + KO
        idx = H.make_index(K, seed=3)
        x = H.make_activations(M, K, idx)
        w = H.make_weights(N, K)
        for sym in (False, True):
            a = mixedgemm.reorder_quantize_x(x.to(dev), idx.to(dev), KN, KS, KO)
            b = (mixedgemm.reorder_quantize_w if sym else mixedgemm.reorder_quantize_w4)(w.to(dev), idx.to(dev), KN, KS, KO)
            if sym and KS == 0 and KO == 0:
                continue
            c = mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5])
            torch.cuda.synchronize()
            st = status()
            an = [H.u8(t) for t in a]; bn = [H.u8(t) for t in b]
            ref = O.matmul(an[0], bn[0], an[1], bn[1], an[2], bn[2], an[3], bn[3], an[4], bn[4], an[5], bn[5], chain=False)
            mx, mean = H.rel_err(H.bits(c), ref)
            exact = float(np.mean(H.bits(c) == ref))
            ok = mx <= 1e-2 and mean <= 1e-3 and all(s == "0x0" for s in st)
            allok &= ok
            print(f"gemm tx={tx_mode} M={M} N={N} split=({KN},{KS},{KO}) sym={sym}: max={mx:.3e} mean={mean:.3e} exact={exact:.4f} status={st[:3]} {'OK' if ok else 'FAIL'}", flush=True)
            if not ok and M * N <= 128 * 256:
                g = O.bf16_bits_to_f32(H.bits(c)); r = O.bf16_bits_to_f32(ref)
                print("   got[0,:8]", g[0, :8], "\n   ref[0,:8]", r[0, :8], "\n   got[64,128:136]", g[min(64, M - 1), 128:136], "\n   ref", r[min(64, M - 1), 128:136], flush=True)
    print("GEMM_ALL_OK" if allok else "GEMM_SOME_FAIL", flush=True)
    if allok:
        L.mmx_set_option(b"gemm_watchdog", 0)
        for (M, N, K) in [(2048, 4096, 4096), (8192, 4096, 4096), (8192, 14336, 4096), (8192, 4096, 14336), (16384, 28672, 4096)]:
            KN, KS, KO = H.SPLITS[K]
            idx = H.make_index(K).to(dev)
            x = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
            w = (torch.randn(N, K, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16)
            a = mixedgemm.reorder_quantize_x(x, idx, KN, KS, KO)
            b = mixedgemm.reorder_quantize_w4(w, idx, KN, KS, KO)
            out = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
            for _ in range(3):
                mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], out=out)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            print(f"gemm timing M={M} N={N} K={K}: {us:.1f} us  {2.0 * M * N * K / us / 1e6:.1f} TFLOPS", flush=True)
            for (opt, val, tag) in ((b"gemm_cta_group", 1, "single-CTA kernel"), (b"gemm_ctas", 74, "74 CTAs")):
                L.mmx_set_option(opt, val)
                for _ in range(2):
                    mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], out=out)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    mixedgemm.matmul(a[0], b[0], a[1], b[1], a[2], b[2], a[3], b[3], a[4], b[4], a[5], b[5], out=out)
                e1.record(); torch.cuda.synchronize()
                us2 = e0.elapsed_time(e1) / 10 * 1e3
                print(f"     [{tag}] {us2:.1f} us  {2.0 * M * N * K / us2 / 1e6:.1f} TFLOPS", flush=True)
                L.mmx_set_option(opt, 0)


def run_stage(name, *args, timeout=600):
    cmd = [sys.executable, os.path.abspath(__file__), "--stage", name, *map(str, args)]
    t0 = time.time()
    try:
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
        out, rc = p.stdout, p.returncode
    except subprocess.TimeoutExpired as e:
        out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), -999
    print(f"===== stage {name} {args} rc={rc} {time.time() - t0:.1f}s =====\n{out}", flush=True)
    return rc, out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="all")
    ap.add_argument("arg", nargs="*")
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    if a.stage == "quant":
        stage_quant()
    elif a.stage == "gemm":
        stage_gemm(int(a.arg[0]))
    else:
        subprocess.run(["nvidia-smi", "--query-gpu=name,clocks.sm,clocks.max.sm,memory.total", "--format=csv"])
        run_stage("quant")
        rc, out = run_stage("gemm", 0)
