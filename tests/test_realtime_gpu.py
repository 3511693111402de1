"""GPU parity of the realtime path (BASELINE config #3): ``infer.lib.rtrvc.RVC.infer`` against the oracle restatement of
infer/lib/rtrvc.py:134-260 over consecutive rolling-window blocks (eager, CUDA-graph capture, replay), and the device-side tail
of gui.py's audio callback (envelope mix + SOLA, gui.py:1024-1087) against its oracle restatement."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

WIN, BLK, SKIP, RET = 43520, 2560, 250, 21        # gui.py sizes at 48 kHz, block 0.16 s, extra 2.5 s, crossfade 0.05 s


def _world(n_blocks):
    from oracle import ivf as OI, rtrvc as ORT, weights as OW
    from infer.lib.rtrvc import RVC
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    hw, rw, sw = OW.hubert_weights(777), OW.rmvpe_weights(4321), OW.synth_weights(1234)
    idx = OI.build_ivf(OW.index_vectors(2000, 768, 1).numpy(), None, seed=0, exact_assign=True)
    stream = OW.synth_voice(2.72 + 0.16 * n_blocks + 0.1, seed=9).numpy()
    orc = ORT.OracleRVC(hw, rw, sw, OW.V2_48K_CONFIG, idx, 0.5, key=0, noise_seed=7)
    rt = RVC(0, 0, OW.synth_cpt(1234, "v2"), Index.from_oracle_layout(idx), 0.5, device="cuda:0", hubert_model=HubertB200(hw, "cuda:0"),
             rmvpe_state_dict=rw)
    return orc, rt, stream


def test_rtrvc_blocks_match_oracle_given_the_pitch_ring():
    """HuBERT on the rolling window, last-frame duplication, tail-only retrieval (frames >= skip_head // 2), x2 upsample and the
    skip_head / return_length synthesizer variant, block by block, with the oracle's pitch ring slice and noise draws."""
    orc, rt, stream = _world(3)
    for b in range(3):
        win = stream[b * BLK: b * BLK + WIN]
        ref = orc.infer(win, BLK, SKIP, RET)
        tap = orc.taps[-1]
        rt.net_g.set_noise(*tap["noise"])
        y = rt.infer(torch.from_numpy(win).cuda(), BLK, SKIP, RET, (tap["pitch"][0].numpy(), tap["pitchf"][0].numpy())).cpu().numpy()
        assert y.shape == ref.shape == (RET * 480,)
        err = np.abs(y - ref).max()
        print(f"[parity] realtime block {b} (shared pitch ring + noise): max abs err {err:.3e}")
        assert err < 1e-3, (b, err)          # measured 3.6e-4 .. 5.1e-4


def test_rtrvc_rmvpe_path_pitch_ring_and_graph_replay():
    """The full block as gui.py drives it ("rmvpe"): RMVPE on the window tail, pitch-ring roll and ``pitch[3:-1]`` write
    (rtrvc.py:209-217) -- first block eager, second captured into a CUDA graph, then replays; shared noise through device buffers
    that the captured graph reads in place."""
    orc, rt, stream = _world(5)
    nb = None
    for b in range(5):
        win = stream[b * BLK: b * BLK + WIN]
        ref = orc.infer(win, BLK, SKIP, RET)
        tap = orc.taps[-1]
        if nb is None:
            nb = (torch.empty_like(tap["noise"][0], device="cuda"), torch.empty_like(tap["noise"][1], device="cuda"))
        nb[0].copy_(tap["noise"][0]); nb[1].copy_(tap["noise"][1])
        rt.net_g._noise[:] = [nb]
        y = rt.infer(torch.from_numpy(win).cuda(), BLK, SKIP, RET, "rmvpe").cpu().numpy()
        rt.net_g._noise.clear()
        cp, cpf = rt.cache_pitch.cpu().numpy(), rt.cache_pitchf.cpu().numpy()
        op_, opf = orc.cache_pitch.numpy(), orc.cache_pitchf.numpy()
        assert np.array_equal(cpf > 0, opf > 0)                                   # same voiced frames, same ring positions
        same = (cp == op_).mean()
        both = (cpf > 0) & (opf > 0)
        rel = np.median(np.abs(cpf[both] / opf[both] - 1)) if both.any() else 0.0
        rms = np.sqrt(np.mean((y - ref) ** 2)) / np.sqrt(np.mean(ref ** 2))
        print(f"[parity] realtime block {b} (own RMVPE, {'graph' if any('graph' in e for e in rt._graphs.values()) else 'eager'}): "
              f"ring coarse same {same:.4f}, median |df0|/f0 {rel:.2e}, waveform rel RMS err {rms:.3e}")
        assert same >= 0.97 and rel < 1e-3
        assert rms < 5e-3          # measured 7e-4 .. 8e-4: f0 differences of ~1e-5 integrate into the NSF sine phase over the 2.7 s window
    assert any("graph" in e for e in rt._graphs.values())


@pytest.mark.parametrize("rate", [1.0, 0.25])
def test_realtime_tail_matches_oracle(rate):
    from oracle import rtrvc as ORT, weights as OW
    from infer.modules.gui import RealtimeTail
    tail = RealtimeTail(48000, 0.16, 0.05, 2.5, "cuda:0")
    assert (tail.block_frame, tail.sola_buffer_frame, tail.sola_search_frame, tail.extra_frame) == (7680, 1920, 480, 120000)
    assert (tail.block_frame_16k, tail.skip_head, tail.return_length, tail.input_frames_16k) == (BLK, SKIP, RET, WIN)
    ot = ORT.SolaTail(tail.block_frame, tail.sola_buffer_frame, tail.sola_search_frame)
    n = tail.return_length * tail.zc
    sig = OW.synth_voice(4.0, sr=48000, seed=21)
    inp = OW.synth_voice(4.0, sr=48000, seed=22)
    offs = []
    for b in range(6):
        start = b * tail.block_frame + (37 * b) % 300              # drifting alignment: SOLA has to find a non-trivial offset
        y, x = sig[start: start + n].clone(), inp[start: start + n + 500].clone()
        yo = ORT.envelope_mix(y, x, tail.zc, rate) if rate < 1 else y
        ref, off = ot.step(yo)
        got = tail.process(y.cuda(), x.cuda(), rate, want_offset=True).cpu()
        offs.append(off)
        assert int(tail.last_offset.item()) == off, (b, int(tail.last_offset.item()), off)
        assert (got - ref).abs().max().item() < 2e-5, (b, (got - ref).abs().max().item())
        assert (tail.sola_buffer.cpu() - ot.sola_buffer).abs().max().item() < 2e-5
    assert len(set(offs[1:])) > 1


def test_phase_vocoder_crossfade_matches_oracle():
    """use_pv (gui.py:27-48, 1078-1083): the SOLA step with the phase-vocoder cross-fade, block after block, against the oracle's
    operation-for-operation restatement.  The reference evaluates cos(w * t + phi) with w * t up to ~6000 rad in float32 (2.4e-4 rad
    of rounding per term); the kernel reduces the 2 pi f i / n part exactly, so it is judged against the float64 evaluation of the
    same formula as well: it must be at least as close to it as the float32 oracle is."""
    from oracle import rtrvc as ORT, weights as OW
    from infer.modules.gui import RealtimeTail
    tail = RealtimeTail(48000, 0.16, 0.05, 2.5, "cuda:0", use_pv=True)
    ot = ORT.SolaTail(tail.block_frame, tail.sola_buffer_frame, tail.sola_search_frame, use_pv=True)
    n = tail.return_length * tail.zc
    sig = OW.synth_voice(4.0, sr=48000, seed=23)
    for b in range(5):
        start = b * tail.block_frame + (41 * b) % 300
        y = sig[start: start + n].clone()
        prev = ot.sola_buffer.clone()
        ref, off = ot.step(y)
        got = tail.process(y.cuda(), None, 1.0, want_offset=True).cpu()
        assert int(tail.last_offset.item()) == off
        head = slice(0, tail.sola_buffer_frame)
        fi, fo = ORT.fade_windows(tail.sola_buffer_frame)
        ref64 = ORT.phase_vocoder(prev.double(), y[off: off + tail.sola_buffer_frame].double(), fo.double(), fi.double())
        e_gpu64 = (got[head].double() - ref64).abs().max().item()
        e_ref64 = (ref[head].double() - ref64).abs().max().item()
        e = (got - ref).abs().max().item()
        print(f"[parity] phase-vocoder block {b}: offset {off}, vs oracle {e:.2e}; vs float64 formula: kernel {e_gpu64:.2e}, float32 oracle {e_ref64:.2e}")
        if b == 0:
            # the previous tail is all zeros: rfft(0) has zero magnitude and the formula's start phase angle(fa) is then decided by
            # the SIGN of the zeros the FFT library happens to return (angle(-0 + 0j) = pi): an artefact of the very first block only
            assert e < 2e-3
        else:
            assert e < 2e-4 and e_gpu64 <= max(2.0 * e_ref64, 2e-5)
        assert (got[tail.sola_buffer_frame:] - ref[tail.sola_buffer_frame:]).abs().max().item() == 0.0     # untouched samples are copies
        assert (tail.sola_buffer.cpu() - ot.sola_buffer).abs().max().item() < 2e-3


@pytest.mark.parametrize("cfg", [
    dict(samplerate=48000, I_noise_reduce=True, O_noise_reduce=True, rms_mix_rate=0.5, threhold=-60.0),
    dict(samplerate=40000, I_noise_reduce=False, O_noise_reduce=False, rms_mix_rate=1.0, threhold=-45.0),
])
def test_realtime_block_matches_oracle_callback(cfg):
    """The whole device side of gui.py's audio callback (gui.py:940-1090) as ONE object / one CUDA graph per block -- input ring,
    TorchGate on input and output, resamplers (incl. model rate != device rate), RVC.infer with its own RMVPE, envelope mix, SOLA --
    against the oracle restatement (OracleCallback) over consecutive blocks: eager, capture, replays.  Noise draws are shared
    through device buffers the captured graph reads in place."""
    from oracle import ivf as OI, rtrvc as ORT, weights as OW
    from infer.lib.rtrvc import RVC
    from infer.modules.gui import RealtimeBlock
    from infer.modules.vc.utils import HubertB200
    from rvc_b200.engine import Index
    sr = cfg["samplerate"]
    hw, rw, sw = OW.hubert_weights(777), OW.rmvpe_weights(4321), OW.synth_weights(1234)
    idx = OI.build_ivf(OW.index_vectors(2000, 768, 1).numpy(), None, seed=0, exact_assign=True)
    orc_rvc = ORT.OracleRVC(hw, rw, sw, OW.V2_48K_CONFIG, idx, 0.5, key=0, noise_seed=11)
    orc = ORT.OracleCallback(orc_rvc, samplerate=sr, block_time=0.16, crossfade_time=0.05, extra_time=2.5, I_noise_reduce=cfg["I_noise_reduce"],
                             O_noise_reduce=cfg["O_noise_reduce"], rms_mix_rate=cfg["rms_mix_rate"], threhold=cfg["threhold"])
    rt = RVC(0, 0, OW.synth_cpt(1234, "v2"), Index.from_oracle_layout(idx), 0.5, device="cuda:0", hubert_model=HubertB200(hw, "cuda:0"),
             rmvpe_state_dict=rw)
    blk = RealtimeBlock(rt, samplerate=sr, block_time=0.16, crossfade_time=0.05, extra_time=2.5, I_noise_reduce=cfg["I_noise_reduce"],
                        O_noise_reduce=cfg["O_noise_reduce"], rms_mix_rate=cfg["rms_mix_rate"], threhold=cfg["threhold"], device="cuda:0")
    assert (blk.block_frame, blk.skip_head, blk.return_length) == (orc.block_frame, orc.skip_head, orc.return_length)
    n_blocks = 5
    mic = OW.synth_voice(0.16 * n_blocks + 0.1, sr=sr, seed=31).numpy()
    mic[: orc.block_frame // 2] *= 0.001                      # a quiet stretch for the threshold gate to act on
    nb, agree, errs = None, 0, []
    for b in range(n_blocks):
        ind = mic[b * orc.block_frame: (b + 1) * orc.block_frame]
        ref = orc.block(ind)
        tap = orc_rvc.taps[-1]
        if nb is None:
            nb = (torch.empty_like(tap["noise"][0], device="cuda"), torch.empty_like(tap["noise"][1], device="cuda"))
        nb[0].copy_(tap["noise"][0]); nb[1].copy_(tap["noise"][1])
        rt.net_g._noise[:] = [nb]
        got = blk.process(ind)
        rt.net_g._noise.clear()
        assert got.shape == ref.shape == (orc.block_frame,)
        res_err = (blk.input_wav_res.cpu() - orc.last["input_wav_res"]).abs().max().item()
        off = int(blk.tail.last_offset.item()) if blk.tail.last_offset is not None else None
        rms = float(np.sqrt(np.mean((got - ref) ** 2)) / max(np.sqrt(np.mean(ref ** 2)), 1e-6))
        print(f"[parity] callback block {b} sr={sr}: 16 kHz window max err {res_err:.2e}, SOLA offset {off} vs {orc.last['offset']}, "
              f"block rel RMS err {rms:.3e} ({'graph' if any('graph' in e for e in blk._graphs.values()) else 'eager'})")
        assert res_err < 2e-4
        if off is None or off == orc.last["offset"]:
            agree += 1
            errs.append(rms)
    assert any("graph" in e for e in blk._graphs.values())
    assert agree >= n_blocks - 1 and max(errs) < 1e-2, (agree, errs)
