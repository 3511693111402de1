#!/usr/bin/env python
"""Parity report at BASELINE config #2's own shapes (10 s utterance, x_pad=3 -> 799 HuBERT frames, 1601 RMVPE frames, T = 1598,
767 040 decoder samples, 100 k-vector IVF2564 index): the sm_100a path and the reference's own fp16 GPU path (oracle modules eager
on the GPU, oracle/gpu_ref.py) against the fp32 CPU oracle, stage by stage and end to end with shared pitch + noise.

    python tools/parity_report.py [--seconds 10] [--nindex 100000] [--out gpurun_out/parity.json]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "retrieval-based-voice-conversion-webui_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def stats(a, b):
    d = np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))
    return {"max": float(d.max()), "mean": float(d.mean()), "rms_rel": float(np.sqrt((d ** 2).mean()) / (np.sqrt((np.asarray(b, dtype=np.float64) ** 2).mean()) + 1e-30))}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--nindex", type=int, default=100000)
    ap.add_argument("--x-pad", type=int, default=3)
    ap.add_argument("--threads", type=int, default=32)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity.json"))
    args = ap.parse_args()
    torch.set_num_threads(min(args.threads, os.cpu_count()))
    from scipy import signal
    from oracle import gpu_ref as GR, ivf as OI, pipeline as OP, rmvpe as ORM, synth as OS, weights as OW
    from infer.modules.vc.pipeline import Pipeline
    from infer.modules.vc.utils import HubertB200
    from rvc.synthesizer import get_synthesizer
    from rvc_b200.engine import Index, Synth

    xp = args.x_pad
    hw, rw, sw = OW.hubert_weights(777), OW.rmvpe_weights(4321), OW.synth_weights(1234)
    audio = OW.synth_voice(args.seconds, seed=0).numpy()
    t0 = time.time()
    idx = OI.build_ivf(OW.index_vectors(args.nindex, 768, 0).numpy(), None, seed=0, exact_assign=False)
    rep = {"config": {"seconds": args.seconds, "x_pad": xp, "nindex": args.nindex, "nlist": int(idx.centroids.shape[0])}}
    print(f"index built in {time.time() - t0:.1f}s", flush=True)

    # ---------------- fp32 CPU oracle (the checker) ----------------
    op = OP.OraclePipeline(48000, xp, 10, 60, 65, hw, rw, sw, OW.V2_48K_CONFIG, noise_seed=3)
    t0 = time.time()
    with torch.no_grad():
        ref = op.pipeline(0, audio.copy(), 0, "rmvpe", idx, 0.75, 1, 48000, 0, 0.25, "v2", 0.33)
    rep["oracle_cpu_seconds"] = time.time() - t0
    tap = op.taps[0]
    pitch, pitchf = op.pitch[0].numpy(), op.pitchf[0].numpy()
    rep["shapes"] = {"hubert_frames": int(tap["feats_hubert"].shape[1]), "T": int(tap["phone"].shape[1]), "out_samples": int(ref.shape[0])}
    print("oracle done", rep["oracle_cpu_seconds"], rep["shapes"], flush=True)
    a_f = signal.filtfilt(OP.bh, OP.ah, audio)
    audio_pad = np.pad(a_f, (16000 * xp, 16000 * xp), mode="reflect").astype(np.float32)
    with torch.no_grad():
        ref_wave = OS.synth_infer(sw, OW.V2_48K_CONFIG, tap["phone"], torch.tensor([tap["phone"].shape[1]]), torch.tensor([0]),
                                  op.pitch[:, :tap["phone"].shape[1]], op.pitchf[:, :tap["phone"].shape[1]], tap["noise"][0], tap["noise"][1])[0, 0].numpy()

    class Cfg:
        x_pad, x_query, x_center, x_max, is_half = xp, 10, 60, 65, True
        device = "cuda:0"
        rmvpe_state_dict = rw

    def product(tag):
        r = {}
        hub = HubertB200(hw, "cuda:0")
        net_g, _ = get_synthesizer(OW.synth_cpt(1234, "v2"), "cuda:0")
        gidx = Index.from_oracle_layout(idx)
        pipe = Pipeline(48000, Cfg())
        feats = hub.extract_features(source=torch.from_numpy(audio_pad)[None].cuda(), output_layer=12)[0][0]
        r["hubert_feats"] = stats(feats.cpu().numpy(), tap["feats_hubert"][0].numpy())
        _, I = gidx.search_device(feats, 8)
        I = I.cpu().numpy()
        r["retrieval_from_own_feats"] = {"top1_same": float((I[:, 0] == tap["ix"][:, 0]).mean()), "all8_same": float((I == tap["ix"]).mean()),
                                         "rows_all8_same": float((I == tap["ix"]).all(1).mean())}
        D2, I2 = gidx.search_device(tap["feats_hubert"][0].cuda(), 8)
        r["retrieval_from_oracle_feats"] = {"I_bit_exact": bool(np.array_equal(I2.cpu().numpy(), tap["ix"])),
                                            "D_bit_exact": bool(np.array_equal(D2.cpu().numpy(), tap["score"]))}
        # synthesizer alone at T = 1598 on the oracle's own phone / pitch / noise
        T = tap["phone"].shape[1]
        syn = Synth(sw, OW.V2_48K_CONFIG, 768)
        w = syn.infer(tap["phone"][0].cuda(), 0, op.pitch[0, :T].cuda(), op.pitchf[0, :T].cuda(), tap["noise"][0][0].cuda(), tap["noise"][1].reshape(-1).cuda()).cpu().numpy()
        r["synth_waveform_T%d" % T] = stats(w, ref_wave)
        # RMVPE f0 at 1601 frames
        c2, f2 = pipe.f0_gen.calculate(audio_pad, len(pitch), 0, "rmvpe", 3)
        both = (f2[: len(pitchf)] > 0) & (pitchf > 0)
        r["rmvpe"] = {"coarse_same": float((c2[: len(pitch)] == pitch).mean()), "median_rel_f0": float(np.median(np.abs(f2[: len(pitchf)][both] / pitchf[both] - 1))),
                      "voiced_same": float(((f2[: len(pitchf)] > 0) == (pitchf > 0)).mean())}
        # end to end, shared pitch track + noise
        net_g.set_noise(*tap["noise"])
        out = pipe.pipeline(hub, net_g, 0, audio.copy(), [0, 0, 0], 0, (pitch, pitchf.astype(np.float64)), gidx, 0.75, 2, 3, 48000, 0, 0.25, "v2", 0.33)
        e = stats(out / 32768.0, ref / 32768.0)
        r["e2e_shared_pitch_noise_fullscale"] = e
        rep[tag] = r
        print(tag, json.dumps(r), flush=True)

    product("b200_fp16")
    if os.environ.get("RVCB_HUBERT_SPLIT") is None:
        os.environ["RVCB_HUBERT_SPLIT"] = "1"
        try:
            product("b200_hubert_split")
        except Exception as e:      # the switch may not exist in this build
            rep["b200_hubert_split"] = {"error": str(e)}
        del os.environ["RVCB_HUBERT_SPLIT"]

    # ---------------- the reference's own GPU path (eager torch, cuDNN/cuBLAS), fp16 and fp32 ----------------
    for half in (True, False):
        tag = "reference_gpu_fp16" if half else "reference_gpu_fp32"
        g = GR.GpuReference(hw, rw, sw, OW.V2_48K_CONFIG, idx, "cuda:0", half, xp)
        r = {}
        with torch.no_grad():
            feats = g.hubert(audio_pad)
            r["hubert_feats"] = stats(feats[0].float().cpu().numpy(), tap["feats_hubert"][0].numpy())
            _, ix = g.retrieve(feats, 0.75)
            r["retrieval_from_own_feats"] = {"top1_same": float((ix[:, 0] == tap["ix"][:, 0]).mean()), "all8_same": float((ix == tap["ix"]).mean()),
                                             "rows_all8_same": float((ix == tap["ix"]).all(1).mean())}
            _, ix0 = g.idx.search(tap["feats_hubert"][0].numpy(), 8)
            r["blas_search_vs_lane_order_oracle_same_feats"] = float((ix0 == tap["ix"]).mean())
            T = tap["phone"].shape[1]
            w = g.synth(tap["phone"].to(g.dev, g.dt), T, op.pitch[:, :T].to(g.dev), op.pitchf[:, :T].to(g.dev), tap["noise"])[0, 0].float().cpu().numpy()
            r["synth_waveform_T%d" % T] = stats(w, ref_wave)
            c2, f2 = g.f0(audio_pad, len(pitch))
            both = (f2[: len(pitchf)] > 0) & (pitchf > 0)
            r["rmvpe"] = {"coarse_same": float((c2[: len(pitch)] == pitch).mean()), "median_rel_f0": float(np.median(np.abs(f2[: len(pitchf)][both] / pitchf[both] - 1))),
                          "voiced_same": float(((f2[: len(pitchf)] > 0) == (pitchf > 0)).mean())}
            out = g.convert(audio.copy(), noise=tap["noise"], pitch_override=(pitch, pitchf))
            r["e2e_shared_pitch_noise_fullscale"] = stats(out / 32768.0, ref / 32768.0)
            # timing of the eager GPU path (whole utterance, host in -> host out), after warm-up
            for _ in range(2):
                g.convert(audio.copy())
            ts = []
            for _ in range(5):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g.convert(audio.copy(), timed=False)
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            g.convert(audio.copy(), timed=True)
            r["ms_per_utterance_host_to_host"] = {"median": float(np.median(ts)), "min": float(min(ts)), "stages_ms": g.stage_ms}
        rep[tag] = r
        print(tag, json.dumps(r), flush=True)
        del g
        torch.cuda.empty_cache()

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(rep, open(args.out, "w"), indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
