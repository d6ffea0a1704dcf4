#!/usr/bin/env python
"""bench.py — the headline measurement: decode tokens/s of the Gemma quantized hot path on B200,
with the mat-vec kernels' HBM roofline fraction and the reference's CPU path timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one decode token through the whole path (embedding row -> L layers -> logits -> greedy
token) at a pinned context length n_kv = ctx (BASELINE.json config 2: Gemma-2B, Q4_0, batch 1,
ctx 2048).  Weights (1.41 GB) are larger than the 126 MB L2, so every step re-streams them from HBM;
no L2 flush is needed.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            j = json.load(open(p))
            return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (pynvml, every 100 ms)."""

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


MODELS = {"2b": "GEMMA_2B", "7b": "GEMMA_7B"}
WTYPES = {"q4_0": 2, "q8_0": 8}


def workload_name(args):
    return f"gemma-{args.model} {args.wtype} batch-1 decode, n_kv={args.ctx} pinned"


def matvec_bytes_per_step(dims, wtype):
    """algorithmic bytes of the mat-vec launches of one token: weights at 18|34 B per 32
    + Q8_0 activations read + fp32 outputs written."""
    bpb = 18 if wtype == 2 else 34
    E, L, H, Hkv, d, F, V = (dims[k] for k in ("n_embd", "n_layer", "n_head", "n_head_kv", "head_dim", "n_ff", "n_vocab"))
    QD, KD = H * d, Hkv * d
    mats = [(QD + 2 * KD, E), (E, QD), (2 * F, E), (E, F)]
    per_layer = sum(n * (k // 32) * bpb + (k // 32) * 34 + 4 * n for n, k in mats)
    per_layer -= 4 * 2 * F - (F // 32) * 34  # gate/up write Q8_0 blocks, not fp32
    return L * per_layer + V * (E // 32) * bpb + (E // 32) * 34 + 4 * V


def cpu_reference_tokens_per_s(args, n_tokens, warm=1, budget_s=None):
    """The reference's CPU path restated by the oracle (thread_pool row split, scalar vec_dot), on the
    same workload; ORACLE USE: this is the checker timed as a baseline, never the product.
    budget_s bounds the timed region: tokens stop early once it is spent (the count actually timed is
    returned), so the run ends within minutes whatever the host's core count."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    import gemma_ggml_b200 as G
    dims = dict(getattr(G, MODELS[args.model]), n_ctx=args.ctx)
    threads = os.cpu_count() or 1
    m = O.Model(wtype=WTYPES[args.wtype], n_threads=threads, **dims)
    m.fill_kv_synth(99)
    tok = np.array([4242], dtype=np.int32)
    for _ in range(warm):
        m.forward(tok, args.ctx - 1, O.REF)
    t0 = time.perf_counter()
    done = 0
    for _ in range(n_tokens):
        logits = m.forward(tok, args.ctx - 1, O.REF)
        tok[0] = O.argmax(logits)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    m.close()
    return done / dt, threads, dt, done


def bench_config(args, world, lowering, kernels_per_step):
    """the `config` object of a line: the SAME keys for both arms (the driver compares them)"""
    return {"workload": workload_name(args), "parallelism": f"tp{world}" if world > 1 else "single",
            "l2": "weights (1.41 GB for 2B Q4_0) > 126 MB L2: re-streamed from HBM every step, no flush needed",
            "lowering": lowering, "graph": "one CUDA graph per token", "kernels_per_step": kernels_per_step}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be compiled
    (un-vendored ggml), so this is the oracle's restatement of its thread_pool mul_mat path
    with all host threads (kind "port": scalar vec_dot, not ggml's hand-written SIMD).  One step = one full decode
    token of the same workload; the timed region is capped at ~150 s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all0 = time.perf_counter()
    warm = max(1, args.warmup)
    # warm-up tokens are bounded in time as well (a token takes ~1 s on 16 cores): the count actually run is reported
    tps_w, threads, dt_w, warm_done = cpu_reference_tokens_per_s(args, warm, warm=0, budget_s=20.0)
    tps, threads, dt, done = cpu_reference_tokens_per_s(args, args.steps, warm=0, budget_s=150.0)
    line = {
        "impl": "reference", "metric": "decode tokens/s", "value": tps, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": done, "warmup_run": warm_done,
        "ms_per_step": 1000.0 * dt / done,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8 dot (dp4a), fp32 accumulate",
        "data": "synthetic (seeded random-init Gemma-architecture weights in ggml quant format, synthetic KV)",
        "config": bench_config(args, args.gpus, "fused persistent decode kernel", 1),
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port",
                         "sample": f"{done} full decode tokens (of {args.steps} requested; timed region capped at 150 s), "
                                   f"oracle restatement of src/hpc.cpp thread_pool path (scalar ggml vec_dot, "
                                   f"gcc -O3 AVX2 autovec; upstream hand-written SIMD kernels not reproduced)"},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_all0,
    }
    print(json.dumps(line), flush=True)


def time_decode(G, L, m, steps, warmup, pos, barrier, dist):
    """K graph-replayed steps, inputs resident, CUDA events on the launching stream; max over ranks"""
    m.set_state(4242, pos, 0)
    m.run_steps(warmup)
    barrier()
    L.ggb_prof_enable(1)
    G.prof_reset()
    barrier()
    L.ggb_prof_start(b"bench_timed_region")
    m.run_steps(steps)
    L.ggb_prof_stop(b"bench_timed_region")
    barrier()
    ms, _ = G.prof_get("bench_timed_region")
    L.ggb_prof_enable(0)
    if dist is not None:
        import torch
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def rank_bytes(m, dims, ctx, world):
    """algorithmic bytes THIS GPU streams per token: its weight pass streams (row shards + replicated phases; 18 | 34 B
    per 32 weights, the same figure as SURVEY 8d for one GPU) + the KV rows of the heads whose cache it holds"""
    kv = dims["n_layer"] * 2 * m.rank_kv_heads * dims["head_dim"] * 2 * ctx
    return m.rank_stream_bytes + kv


def sub_record_7b(G, L, args, dist, rank, world, barrier, peak):
    """BASELINE config 4 on the same run: Gemma-7B Q4_0, n_kv 2048, tensor parallel over the same ranks"""
    dims = dict(G.GEMMA_7B, n_ctx=2048)
    m = G.GemmaModel(G.make_config(wtype=2, **dims))
    if world > 1:
        m.init_tensor_parallel(dist, rank, world)
    m.init_synthetic_weights(); m.init_kv_cache(); m.fill_kv_cache_synthetic(99)
    m.build_compute_graph(capture_acts=False, fused=True)
    steps = max(16, min(args.steps, 128))
    ms = time_decode(G, L, m, steps, max(3, min(args.warmup, 8)), 2047, barrier, dist)
    rb = rank_bytes(m, dims, 2048, world)
    step_s = ms / steps / 1000.0
    rec = {"workload": "gemma-7b q4_0 batch-1 decode, n_kv=2048 pinned", "value": steps / (ms / 1000.0), "unit": "tokens/s",
           "ms_per_step": ms / steps, "steps": steps, "parallelism": f"tp{world}" if world > 1 else "single",
           "per_gpu_bytes_per_step": rb, "per_gpu_gbs": rb / step_s / 1e9, "per_gpu_frac_of_peak": rb / step_s / 1e9 / peak}
    m.close()
    return rec


def sub_record_prefill(G, L, args):
    """BASELINE config 3 on the same run (1 GPU): Gemma-2B Q8_0, prefill of 512 tokens through the tcgen05 GEMMs"""
    T = 512
    dims = dict(G.GEMMA_2B, n_ctx=T)
    m = G.GemmaModel(G.make_config(wtype=8, **dims))
    m.init_synthetic_weights(); m.init_kv_cache(); m.build_compute_graph(capture_acts=False, fused=True)
    prompt = ((1 + 7919 * np.arange(T)) % dims["n_vocab"]).astype(np.int32)
    for _ in range(3):
        m.prefill(prompt, 0)
    G.sync()
    steps = 5
    t0 = time.perf_counter()            # the step as a user runs it: no profiling probes
    for _ in range(steps):
        m.prefill(prompt, 0)
    G.sync()
    wall = time.perf_counter() - t0
    L.ggb_prof_enable(1)                # then the same steps with cudaEvent probes around every operator: GEMM time
    G.prof_reset()
    for _ in range(steps):
        m.prefill(prompt, 0)
    G.sync()
    gemm_ms, _ = G.prof_get("op_mul_mat")
    L.ggb_prof_enable(0)
    peak_tf = 1396.9
    try:
        peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    ach = m.prefill_gemm_flops / (gemm_ms / steps / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    rec = {"workload": "gemma-2b q8_0 prefill of 512 tokens (host call: H2D ids, per-op launches, D2H token)",
           "value": T * steps / wall, "unit": "tokens/s", "ms_per_step": 1000.0 * wall / steps,
           "gemm_tflops": ach, "gemm_frac_of_tensor_peak": ach / peak_tf, "gemm_ms_per_step": gemm_ms / steps,
           "tensor_peak_tflops": peak_tf, "kernel": "gemm_bf16w_kernel (tcgen05.mma kind::f16, TMEM accumulators)"}
    m.close()
    return rec


def mul_mat_dropin_us(G):
    """the plugin boundary itself: the reference's mul_mat entry (host buffers in, host buffer out, src/hpc.h:22-32)
    per Gemma-2B Q4_0 shape, one column: us per call after the weight is resident (H2D of the Q8_0 row + kernel + D2H)"""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    out = {}
    L = G.cuda_lib()
    rng = np.random.default_rng(5)
    for name, (K, N) in {"q/o 2048x2048": (2048, 2048), "gate/up 2048x16384": (2048, 16384), "down 16384x2048": (16384, 2048)}.items():
        nb = K // 32
        w = rng.integers(0, 256, size=(N, nb * 18), dtype=np.uint8)
        w[:, 0::18] = 0; w[:, 1::18] = 0x1c  # fp16 scales ~ 0.004
        act = rng.integers(0, 256, size=(nb * 34,), dtype=np.uint8)
        act[0::34] = 0; act[1::34] = 0x20
        dst = np.zeros(N, dtype=np.float32)
        call = lambda: L.ggb_mul_mat(N, 1, 1, nb * 18, 1, N * 4, N * 4, nb * 34, K, w.ctypes.data_as(C.c_void_p),
                                     dst.ctypes.data_as(C.c_void_p), 2, act.ctypes.data_as(C.c_char_p))
        if call() != 0:
            out[name] = None
            continue
        for _ in range(5):
            call()
        t0 = time.perf_counter()
        for _ in range(50):
            call()
        out[name] = (time.perf_counter() - t0) / 50 * 1e6
        L.ggb_mul_mat_forget(w.ctypes.data_as(C.c_void_p))
    return out


def kquant_matvec_record(G, peak_gbs):
    """SURVEY.md 8f-2 on the record: the resident (packed) K-quant mat-vecs on the tied-embedding shape of Gemma-2B
    (256000 x 2048, one Q8_K activation row), device-timed; algorithmic bytes = ggml's block bytes (144 / 210 per 256)."""
    L = G.cuda_lib()
    rng = np.random.default_rng(11)
    n, k = 256000, 2048
    nb = k // 256
    out = {}
    x = rng.standard_normal((1, k)).astype(np.float32)
    act = G.quantize_q8_K(x)
    ad = G.DeviceBuffer.from_numpy(act)
    dst = G.DeviceBuffer(4 * n)
    for name, size, d_off in (("q4_K", 144, 0), ("q6_K", 210, 208)):
        w = rng.integers(0, 256, size=(n, nb, size), dtype=np.uint8)
        w[:, :, d_off] = 0
        w[:, :, d_off + 1] = 0x1c            # fp16 super-scale ~ 0.004
        if name == "q4_K":
            w[:, :, 2] = 0
            w[:, :, 3] = 0x1c
        w = np.ascontiguousarray(w.reshape(n, nb * size))
        wd = G.DeviceBuffer.from_numpy(w)
        if name == "q4_K":
            packed = G.DeviceBuffer(w.nbytes)
            G.check(L.ggb_op_q4k_repack(wd.ptr, n, w.shape[1], k, packed.ptr), "repack")
            call = lambda: L.ggb_op_mul_mat_q4k_packed(packed.ptr, n, k, ad.ptr, act.shape[1], 1, dst.ptr, n)
        else:
            packed = G.DeviceBuffer(L.ggb_q6k_packed_bytes(n, k))
            G.check(L.ggb_op_q6k_repack(wd.ptr, n, w.shape[1], k, packed.ptr), "repack")
            call = lambda: L.ggb_op_mul_mat_q6k_packed(packed.ptr, n, k, ad.ptr, act.shape[1], 1, dst.ptr, n)
        for _ in range(5):
            G.check(call(), "kquant mat-vec")
        G.sync()
        L.ggb_prof_enable(1); G.prof_reset()
        L.ggb_prof_start(b"kq")
        for _ in range(40):
            call()
        L.ggb_prof_stop(b"kq")
        G.sync()
        ms, _ = G.prof_get("kq")
        L.ggb_prof_enable(0)
        us = ms * 1e3 / 40
        gbs = w.nbytes / (us * 1e-6) / 1e9
        out[name] = {"shape": "256000 x 2048 (resident, packed form)", "us_per_launch": us, "algorithmic_gbs": gbs,
                     "frac_of_hbm_peak": gbs / peak_gbs}
        del wd, packed
    return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import gemma_ggml_b200 as G

    dist = None
    if world > 1:
        # NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION: the line rank 0 prints must stay the only one
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local_rank)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod
    G.init(local_rank)
    L = G.cuda_lib()

    def barrier():
        G.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- multi-GPU: parity first (tiny MQA + MHA models, both quant types, fused lowering, against the oracle, bit for
    # bit) - ORACLE USE: the checker, before and outside every timed region
    tp_parity = None
    if world > 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        import tp_worker
        try:
            n = tp_worker.check(G, O, dist, rank, world, lowerings=(True,))
            tp_parity = f"ok ({n} cases: small mqa / mha / thin models x q4_0/q8_0, tokens and logits bit-identical to the oracle)"
        except AssertionError as e:
            tp_parity = f"FAILED: {e}"
        barrier()

    wtype = WTYPES[args.wtype]
    dims = dict(getattr(G, MODELS[args.model]), n_ctx=args.ctx)
    cfg = G.make_config(wtype=wtype, **dims)
    m = G.GemmaModel(cfg)
    if world > 1:
        m.init_tensor_parallel(dist, rank, world)
    m.init_synthetic_weights()
    m.init_kv_cache()
    m.fill_kv_cache_synthetic(99)
    m.build_compute_graph(capture_acts=False, fused=(args.mode == "fused"))
    fused = m.is_fused
    kernels_per_step = m.kernels_per_step
    pos = args.ctx - 1

    # ---- value: K graph-replayed steps, inputs resident, CUDA events on the launching stream
    with ClockSampler(local_rank) as clocks:
        ms = time_decode(G, L, m, args.steps, args.warmup, pos, barrier, dist)
    tok_s = args.steps / (ms / 1000.0)

    # ---- e2e: the host-buffer call a user makes (pinned H2D of token+pos, graph, D2H of the token)
    e2e_steps = max(8, min(args.steps, 128))
    for _ in range(3):
        m.decode_step(4242, pos)
    barrier()
    t0 = time.perf_counter()
    tok = 4242
    for _ in range(e2e_steps):
        tok = m.decode_step(tok, pos)
    G.sync()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_tok_s = e2e_steps / e2e_s

    # ---- roofline of the dominant kernel: PER GPU (what this GPU streams per launch / its launch time / its peak)
    peak, peak_src = load_peaks()
    kv_bytes = dims["n_layer"] * 2 * dims["n_head_kv"] * dims["head_dim"] * 2 * args.ctx
    step_s = ms / args.steps / 1000.0
    per_class = {}
    wname = "Q4_0" if wtype == 2 else "Q8_0"
    if fused:
        # one persistent kernel per token: algorithmic bytes per launch = weights + KV this GPU reads,
        # duration = the CUDA-event time of the timed region / launches
        step_bytes = (m.weight_bytes + kv_bytes) if world == 1 else rank_bytes(m, dims, args.ctx, world)
        achieved = step_bytes / step_s / 1e9
        quad = "true" if args.model == "7b" else "false"
        roof_kernel = f"decoder_kernel<{wname}, QUAD={quad}, PROF=false> (persistent, one cooperative launch per token)"
        roof_bytes, roof_launches, roof_us = step_bytes, 1.0, step_s * 1e6
    else:
        # per-operator lowering: cudaEvent probes around every launch of an eager pass, mat-vec class
        prof_steps = 8
        m.set_state(4242, pos, 0)
        m.run_steps_profiled(2)
        G.prof_reset()
        m.run_steps_profiled(prof_steps)
        for name in ("op_mul_mat", "op_attn", "op_rms_norm", "op_rope", "op_get_rows", "op_argmax"):
            tms, calls = G.prof_get(name)
            per_class[name] = {"ms_per_step": tms / prof_steps, "launches_per_step": calls / prof_steps}
        step_bytes = (m.weight_bytes + kv_bytes) / world
        roof_bytes = matvec_bytes_per_step(dims, wtype) / world
        mv_ms = per_class["op_mul_mat"]["ms_per_step"]
        roof_launches = per_class["op_mul_mat"]["launches_per_step"]
        achieved = roof_bytes / (mv_ms / 1000.0) / 1e9 if mv_ms > 0 else 0.0
        roof_kernel = f"matvec_kernel<{wname}> / gate_up_kernel<{wname}> (all weight mat-vecs of a step, eager + cudaEvent probes)"
        roof_us = 1000.0 * mv_ms / roof_launches if roof_launches else None
    traffic = None
    if world == 1:  # one ncu --set full capture of this kernel on this workload (profiles/traffic.json); not valid for shards
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and args.model == "2b" and args.wtype == "q4_0" and args.ctx == 2048:
            try:
                traffic = json.load(open(tpath)).get("fused" if fused else "ops")
            except Exception:
                traffic = None
    exp_mode = int(L.ggb_decoder_exp_computed(m.H.gm_decoder(m.h))) if fused else None
    m.close()

    # ---- sub-records measured in the same run
    extra = {}
    if fused and args.model == "2b" and not args.no_sub_records:
        try:
            extra["config4_7b_q4_0"] = sub_record_7b(G, L, args, dist, rank, world, barrier, peak)
        except Exception as e:
            extra["config4_7b_q4_0"] = {"failed": str(e)}
        if world == 1:
            try:
                extra["config3_prefill_2b_q8_0"] = sub_record_prefill(G, L, args)
            except Exception as e:
                extra["config3_prefill_2b_q8_0"] = {"failed": str(e)}
            try:
                extra["mul_mat_dropin_us_per_call"] = mul_mat_dropin_us(G)
            except Exception as e:
                extra["mul_mat_dropin_us_per_call"] = {"failed": str(e)}
            try:
                extra["kquant_matvec"] = kquant_matvec_record(G, peak)
            except Exception as e:  # noqa: BLE001
                extra["kquant_matvec"] = {"failed": str(e)}
    if rank != 0:
        return
    line = {
        "metric": "decode tokens/s", "value": tok_s, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int8 dot (dp4a), fp32 accumulate",
        "data": "synthetic (seeded random-init Gemma-architecture weights in ggml quant format, synthetic KV)",
        "config": bench_config(args, world, "fused persistent decode kernel" if fused else "per-operator kernels", kernels_per_step),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 16, "d2h_bytes_per_step": 4,
                "steps": e2e_steps},
        "gpu_launches": kernels_per_step * args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "kernel": roof_kernel, "algorithmic_bytes_per_launch": roof_bytes / roof_launches,
                     "launches_per_step": roof_launches, "avg_launch_us": roof_us,
                     "scope": "per GPU: bytes this GPU streams per launch / its launch time / its measured peak",
                     "step_frac_of_8tbs": step_bytes / step_s / 1e9 / 8000.0},
        "parity": "tokens and logits bit-identical to the oracle's scalar ggml order (tests/test_model_gpu.py; config 1 "
                  "at full size from tests/golden/)",
        "softmax_exp": {2: "computed (fp32 expf, proved equal to the fp16 table)", 1: "computed (double exp, proved equal)",
                        0: "fp16 table lookup", None: None}[exp_mode],
        "op_classes": per_class,
    }
    if tp_parity is not None:
        line["tp_parity"] = tp_parity
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        try:
            tps, threads, dt, done = cpu_reference_tokens_per_s(args, args.cpu_tokens, budget_s=30.0)
            line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port",
                                    "sample": f"{done} full decode tokens of the same workload in {dt:.1f} s "
                                              f"(oracle restatement of the reference thread_pool path, scalar vec_dot; "
                                              f"upstream ggml's hand-written SIMD kernels are not reproduced)"}
        except Exception as e:  # the baseline must not take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)


def run_prefill(args):
    """Secondary workload (BASELINE.json config 3): batched prefill of T tokens through the dequant -> bf16
    tcgen05 GEMMs.  value = prompt tokens/s; roofline = tensor pipe (GEMM flops / cudaEvent time of the GEMM
    launches vs the measured dense bf16 peak)."""
    import gemma_ggml_b200 as G
    G.init(0)
    L = G.cuda_lib()
    T = args.prefill_tokens
    dims = dict(getattr(G, MODELS[args.model]), n_ctx=max(args.ctx, T))
    m = G.GemmaModel(G.make_config(wtype=WTYPES[args.wtype], **dims))
    m.init_synthetic_weights(); m.init_kv_cache(); m.build_compute_graph(capture_acts=False, fused=True)
    prompt = ((1 + 7919 * np.arange(T)) % dims["n_vocab"]).astype(np.int32)
    for _ in range(max(3, args.warmup if args.warmup < 8 else 3)):
        m.prefill(prompt, 0)
    G.sync()
    steps = max(1, min(args.steps, 20))
    with ClockSampler(0) as clocks:       # the step as a user runs it: no profiling probes
        t0 = time.perf_counter()
        for _ in range(steps):
            m.prefill(prompt, 0)
        G.sync()
        wall = time.perf_counter() - t0
    L.ggb_prof_enable(1)                  # then the same steps with cudaEvent probes around every operator
    G.prof_reset()
    for _ in range(steps):
        m.prefill(prompt, 0)
    G.sync()
    gemm_ms, gemm_calls = G.prof_get("op_mul_mat")
    classes = {}
    for name in ("op_mul_mat", "op_attn", "op_rms_norm", "op_rope", "op_cvt", "op_gelu", "op_cpy", "op_get_rows", "op_argmax"):
        tms, calls = G.prof_get(name)
        classes[name] = {"ms_per_step": tms / steps, "launches_per_step": calls / steps}
    L.ggb_prof_enable(0)
    flops = m.prefill_gemm_flops
    peak_tf = 1396.9
    try:
        peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
    except Exception:
        pass
    ach = flops / (gemm_ms / steps / 1000.0) / 1e12 if gemm_ms > 0 else 0.0
    line = {
        "metric": "prefill tokens/s", "value": T * steps / wall, "unit": "tokens/s", "n_gpus": 1, "steps": steps,
        "warmup": 3, "ms_per_step": 1000.0 * wall / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 x bf16 -> fp32 (weights dequantised from " + args.wtype + ")",
        "data": "synthetic (seeded random-init Gemma-architecture weights in ggml quant format)",
        "config": {"workload": f"gemma-{args.model} {args.wtype} prefill of {T} tokens (host call: H2D ids, per-op launches, D2H token)",
                   "timing": "value / ms_per_step: eager per-op launches, no probes; roofline and op_classes: a second pass of the same steps with cudaEvent probes per operator (the probes add ~1 ms of launch gaps per step)"},
        "clocks": clocks.summary(),
        "roofline": {"bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                     "traffic": None, "kernel": "gemm_bf16w_kernel (tcgen05.mma kind::f16, TMEM accumulators)",
                     "algorithmic_flops_per_step": flops, "gemm_ms_per_step": gemm_ms / steps,
                     "launches_per_step": gemm_calls / steps, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained"},
        "op_classes": classes,
    }
    m.close()
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="2b", choices=list(MODELS))
    ap.add_argument("--wtype", default="q4_0", choices=list(WTYPES))
    ap.add_argument("--ctx", type=int, default=2048)
    ap.add_argument("--mode", default="fused", choices=["fused", "ops", "prefill"])
    ap.add_argument("--prefill-tokens", type=int, default=512)
    ap.add_argument("--cpu-tokens", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "prefill":
        run_prefill(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
