#!/usr/bin/env python
"""bench.py -- RVC v2/48k inference hot path on B200 (BASELINE.json metric: 48 kHz output samples / s).

One "step" = one 10 s / 16 kHz utterance through the whole path (config #2: v2/48k model, RMVPE f0,
100k-vector IVF2564,Flat index, k=8, index_rate 0.75, x_pad=3 => 16 s of model compute, 479 040 output samples):
high-pass filtfilt + reflect pad -> HuBERT features -> IVF-Flat search + blend || RMVPE f0 -> f0 post-processing ->
SynthesizerTrnMs768NSFsid.infer -> RMS mix + int16.

  value   device-resident: the utterance (float32) already in HBM, CUDA events around the product path's own body
          (Pipeline._dev_body, the very function VC.vc_single runs), replayed as one CUDA graph
  e2e     VC.vc_single (the reference's public entry) with HOST numpy audio in and host int16 audio out:
          pinned H2D of the utterance, the same body, D2H of the int16 result, all host work inside the timed region
  --impl reference   the reference's CPU path (oracle restatement, pinned against the reference's own modules)
                     on the box's host cores for the same metric/config.
Synthetic seeded weights / audio / index (no assets, no network): "data": "synthetic".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

OUT_SAMPLES = 479040          # 160000 in @16k -> (2*799)*480 - 2*3*48000 out @48k
UTT_SECONDS = 10.0
ALGO_FLOPS = (246.0 + 117.1 + 1833.0) * 1e9     # BASELINE.md section 2, offline fp16-config column


class Cfg:  # the object Pipeline reads (configs/config.py:219-230, fp16 "6G" preset)
    x_pad, x_query, x_center, x_max, is_half = 3, 10, 60, 65, True

    def __init__(self, device):
        self.device = device
        self.rmvpe_state_dict = None


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.samples, self.reasons, self.maxmhz, self.stop_flag = gpu_index, [], set(), 0, False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(", ")
                self.samples.append(float(o[0])); self.maxmhz = float(o[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), o[2:6]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.maxmhz or None,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


CPU_SAMPLE_SECONDS = 2.0      # seconds of MODEL compute per CPU sample (the 10 s utterance is 16 s of compute at x_pad=3)


def make_cpu_sample(audio, idx):
    """A bounded sample of the same workload for the CPU arm: a CPU_SAMPLE_SECONDS-long slice of the padded utterance goes
    through the whole reference path (RMVPE f0 -> HuBERT -> IVF search + blend -> synthesizer); credited output samples =
    the slice's share of the utterance's 479 040 samples (per-second work is identical, attention is the only
    super-linear term and favours the short slice)."""
    from scipy import signal
    from oracle import pipeline as OP, rmvpe as ORM, weights as OW
    pipe = OP.OraclePipeline(48000, 3, 10, 60, 65, OW.hubert_weights(777), OW.rmvpe_weights(4321), OW.synth_weights(1234), OW.V2_48K_CONFIG)
    a = signal.filtfilt(OP.bh, OP.ah, audio)
    audio_pad = np.pad(a, (48000, 48000), mode="reflect").astype(np.float32)
    n = int(CPU_SAMPLE_SECONDS * 16000)
    chunk = np.ascontiguousarray(audio_pad[96000: 96000 + n])
    big = idx.reconstruct_n(0, idx.ntotal)
    credit = OUT_SAMPLES * n / audio_pad.shape[0]

    def step():
        with torch.no_grad():
            pitch, pitchf = ORM.calculate(pipe.rw, chunk, n // 160, 0)
            pt = torch.tensor(pitch).unsqueeze(0).long()
            pf = torch.tensor(pitchf.astype(np.float32)).unsqueeze(0)
            return pipe.vc(torch.tensor([0]), chunk, pt, pf, idx, big, 0.75, "v2", 0.33)
    return step, credit


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path (oracle restatement) on the host cores."""
    if rank != 0:
        return
    from oracle import ivf as OI, pipeline as OP, weights as OW
    cores = min(os.cpu_count(), 16)        # more threads than this only adds fork/join overhead on these small convolutions
    torch.set_num_threads(cores)
    audio = OW.synth_voice(UTT_SECONDS, seed=0).numpy()
    vec = OW.index_vectors(100000, 768, 0).numpy()
    idx = OI.build_ivf(vec, None, seed=0, exact_assign=False)
    step, credit = make_cpu_sample(audio, idx)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps * credit / dt
    line = {"impl": "reference", "metric": "48kHz audio samples/sec (v2/48k infer, RMVPE, IVF index)", "value": v, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "rtf_x": v / 48000.0,
            "config": {"workload": "configs[1]: v2/48k, RMVPE f0, 100k-vec IVF2564,Flat k=8 rate 0.75, 10s utterance, x_pad=3 (16s compute)"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"each step = a {CPU_SAMPLE_SECONDS:g} s slice (of 16 s) of the padded utterance through the whole reference path, credited {credit:.0f} output samples; torch CPU fp32, {cores} threads of {os.cpu_count()} cores"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    if args.warmup < 3:          # timing hygiene: never fewer than 3 untimed warm-up steps (reported as run)
        args.warmup = 3
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from rvc_b200 import _lib, synthetic as SY
    from rvc_b200.engine import Index
    from rvc_b200.index_build import build_ivf_layout
    from infer.modules.vc.modules import VC
    from infer.modules.vc.utils import HubertB200
    _lib.init(local_rank)

    # ---- model containers (seeded synthetic checkpoints) ----
    cfg = Cfg(str(dev))
    cfg.rmvpe_state_dict = SY.rmvpe_weights(4321)
    vc = VC(cfg)
    vc.hubert_model = HubertB200(SY.hubert_weights(777), dev)
    cpt = SY.synth_cpt(1234, "v2")
    vc.get_vc(cpt)
    # index: rank 0 builds the layout; the optional one-time broadcast over NCCL shares it (the only collective of the path)
    if rank == 0:
        lay = build_ivf_layout(SY.index_vectors(100000, 768, 0).numpy(), None, seed=0, device=str(dev))
    if world > 1:
        from rvc_b200.dist_utils import broadcast_layout
        lay = broadcast_layout(lay if rank == 0 else None, 0, str(dev))
    index = Index.from_oracle_layout(lay, local_rank)

    audio = SY.synth_voice(UTT_SECONDS, seed=rank).numpy()          # each rank converts its own utterances (weak scaling)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def e2e_step():
        info, out = vc.vc_single(0, audio, 0, None, "rmvpe", index, "", 0.75, 3, 0, 0.25, 0.33)
        assert out is not None, info
        return out[1]

    # ---- device-resident step: the product path's own body (Pipeline._dev_body: filtfilt -> pad -> [RMVPE -> f0 post] ||
    # [HuBERT -> retrieval] -> synthesizer -> RMS mix -> int16) with the utterance already in HBM: same kernels, same stream
    # structure and the same data dependencies (the synthesizer consumes the pitch RMVPE produced in this very step) ----
    pipe = vc.pipeline
    x_dev = torch.from_numpy(np.divide(audio, max(1.0, np.abs(audio).max() / 0.95)).astype(np.float32)).to(dev)
    body_args = (vc.hubert_model, vc.net_g, torch.tensor(0).unsqueeze(0).long(), [0, 0, 0], 0, index, index.vectors, 0.75, 1, 48000,
                 0.25, "v2", 0.33, True)

    def dev_step(use_side=True):
        if use_side:
            return pipe._dev_body(x_dev, *body_args)
        real = pipe._side
        pipe._side = torch.cuda.current_stream()
        try:
            return pipe._dev_body(x_dev, *body_args)
        finally:
            pipe._side = real

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            flush.fill_(1)                      # L2 flush between timed iterations (outside the event pair)
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        ms = sum(s.elapsed_time(e) for s, e in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(t.item())

    # the device-resident step is a fixed sequence of ~510 launches: capture it once (both streams) and replay the CUDA graph
    for _ in range(2):
        dev_step()
    torch.cuda.synchronize()
    use_graph = os.environ.get("RVCB_BENCH_GRAPH", "1") == "1"
    launches = 0
    step_fn = dev_step
    if use_graph:
        try:
            graph = torch.cuda.CUDAGraph()
            l0 = _lib.lib().rvcb_launch_count()
            with torch.cuda.graph(graph):
                graph_out = dev_step()
            launches = _lib.lib().rvcb_launch_count() - l0
            step_fn = graph.replay
        except Exception as e:      # capture is an optimisation of launch overhead only; report and fall back to eager launches
            print(f"[bench] CUDA graph capture failed ({e}); timing eager launches", file=sys.stderr)
            use_graph = False
            torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.lib().rvcb_launch_count()
    dev_ms = timed(step_fn, args.steps, args.warmup)
    if not use_graph:
        launches = (_lib.lib().rvcb_launch_count() - l0) // (args.steps + args.warmup)
    e2e_ms = timed(e2e_step, args.steps, args.warmup)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel family (gemm_tc): live CUDA events around every launch of one step set ----
    import ctypes as C
    _lib.check(_lib.lib().rvcb_prof_begin())
    for _ in range(3):
        dev_step(use_side=False)      # serial, so every launch's event pair times that kernel alone
    gms, gn = C.c_double(0), C.c_ulonglong(0)
    _lib.check(_lib.lib().rvcb_prof_end(C.byref(gms), C.byref(gn)))
    gemm_ms_per_step = gms.value / 3
    cls = [(C.c_double * 3)() for _ in range(4)]
    _lib.check(_lib.lib().rvcb_prof_classes(*cls))
    ws_ms, ws_n, ws_flops, ws_bytes = cls[0][1] / 3, cls[1][1] / 3, cls[2][1] / 3, cls[3][1] / 3
    pk, pk_src = peaks()
    achieved = ALGO_FLOPS / (gemm_ms_per_step * 1e-3) / 1e12
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))

    value = world * args.steps * OUT_SAMPLES / (dev_ms * 1e-3)
    e2e_v = world * args.steps * OUT_SAMPLES / (e2e_ms * 1e-3)
    line = {
        "metric": "48kHz audio samples/sec (v2/48k infer, RMVPE, IVF index)", "value": value, "unit": "samples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate+residuals",
        "data": "synthetic", "rtf_x_per_gpu": value / world / 48000.0,
        "config": {"workload": "configs[1]: v2/48k, RMVPE f0, 100k-vec IVF2564,Flat k=8 rate 0.75, 10s utterance, x_pad=3 (16s compute)",
                   "l2": "256 MiB flush between timed iterations", "utterances_per_gpu_per_step": 1,
                   "device_step": f"CUDA graph replay of the {int(launches)}-launch step" if use_graph else "eager launches"},
        "e2e": {"value": e2e_v, "unit": "samples/s", "h2d_bytes_per_step": int(160000 * 4 + 8),   # the utterance (float32, one pinned copy) + speaker id
                "d2h_bytes_per_step": int(OUT_SAMPLES * 2),  # mixed + normalised waveform, int16
                "ms_per_step": e2e_ms / args.steps, "rtf_x_per_gpu": e2e_v / world / 48000.0,
                "api": "infer.modules.vc.VC.vc_single (host numpy in, host int16 out)"},
        "gpu_launches": int(launches * args.steps),
        "gpu_launches_per_step": int(launches),
        "clocks": sampler.summary(),
        # dominant kernel by work: the weight-stationary vocoder convolution (44 % of the utterance's FLOPs); it is HBM-bound
        # in this unfused layer-by-layer design: algorithmic bytes = activations in + weights + fp32 residual in + outputs
        "roofline": {"bound": "hbm", "achieved": ws_bytes / (ws_ms * 1e-3) / 1e9, "peak": pk.get("hbm_gbs"), "unit": "GB/s",
                     "frac": ws_bytes / (ws_ms * 1e-3) / 1e9 / pk.get("hbm_gbs"),
                     "traffic": 244.0e6, "traffic_note": "dram read+write of one stage-2 c2 launch (ncu --set full, profiles/prof_r1u_ws2_metrics.txt: 147.4 MB read + 96.6 MB written, 47.2 us) vs 294 MB algorithmic; part of the fp16/fp32 output is still in L2 when the kernel ends",
                     "kernel": "gemm_ws2_kernel<*> / gemm_ws_kernel<*> (weight-stationary vocoder resblock convolutions, stages 1-3)", "launches_per_step": ws_n,
                     "avg_launch_us": ws_ms / max(ws_n, 1) * 1e3, "algorithmic_bytes_per_step": ws_bytes, "peak_source": pk_src,
                     "tensor_view": {"achieved_tflops": ws_flops / (ws_ms * 1e-3) / 1e12, "frac_of_bf16_sustained": ws_flops / (ws_ms * 1e-3) / 1e12 / peak}},
        "roofline_all_gemm": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "traffic": None, "kernel": "gemm_tc_kernel<*> + gemm_ws2_kernel<*> + gemm_ws_kernel<*> (all tcgen05 implicit-GEMM launches of one utterance, timed serially)",
                     "launches_per_step": int(gn.value // 3), "ms_per_step": gemm_ms_per_step, "peak_source": pk_src,
                     "algorithmic_flops_per_step": ALGO_FLOPS},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # bounded CPU sample: ONE full utterance through the oracle pipeline on the host cores
        from oracle import ivf as OI
        cpu_threads = min(os.cpu_count(), 16)
        torch.set_num_threads(cpu_threads)
        class _L:  # reuse the already built layout for the oracle index (membership is data, not arithmetic)
            pass
        assign = np.empty(lay.vectors.shape[0], dtype=np.int64)
        for l in range(len(lay.list_off) - 1):
            assign[lay.list_ids[lay.list_off[l]:lay.list_off[l + 1]]] = l
        oidx = OI.IVFFlat(lay.centroids, lay.vectors, assign)
        step, credit = make_cpu_sample(audio, oidx)
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": credit / dt, "unit": "samples/s", "cores": cpu_threads, "kind": "port",
                                "sample": f"one {CPU_SAMPLE_SECONDS:g} s slice (of 16 s) of the padded utterance through the whole reference path, credited "
                                          f"{credit:.0f} output samples; torch CPU fp32, {cpu_threads} threads of {os.cpu_count()} cores, no warm-up"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
