"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs, weights and
mode flags.  Bar (BASELINE.json north_star): max-abs <= 1e-3 on the enhanced magnitude spectrogram, inputs
peak-normalised so magnitudes reach ~48.  Wav outputs are held to the same absolute bar."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL_MAG = 1e-3
TOL_WAV = 1e-3


@pytest.fixture(scope="module")
def blob(weights):
    from nunet_b200.weights import pack_blob
    return pack_blob(weights)


@pytest.fixture(scope="module")
def oracles(weights):
    from oracle.nunet_oracle import Oracle
    return {m: Oracle(weights, ctfa_mode=m) for m in ("causal_avg32", "frame_div32")}


def _engine(blob, **kw):
    from nunet_b200.engine import NunetEngine
    return NunetEngine(blob, **kw)


def test_library_is_the_cuda_one():
    from nunet_b200 import _lib
    L = _lib.lib()
    assert L.nunet_abi_version() == 1
    assert torch.cuda.get_device_capability(0)[0] == 10


@pytest.mark.parametrize("mode", ["causal_avg32", "frame_div32"])
def test_offline_matches_golden_and_oracle(blob, oracles, golden_io, mode):
    wav = golden_io["wav"]
    eng = _engine(blob, max_frames=2 * 40, ctfa_mode=mode)
    y, est = eng.forward_wav(torch.from_numpy(wav).cuda())
    est, y = est.cpu().numpy(), y.cpu().numpy()
    assert eng.last_launch_count > 100
    ref = golden_io[f"est_{mode}"]
    assert ref.max() > 20.0
    assert np.abs(est - ref).max() <= TOL_MAG
    assert np.abs(y - golden_io[f"wav_{mode}"]).max() <= TOL_WAV
    with torch.no_grad():
        y2, est2 = oracles[mode].forward_wav(wav)
    assert np.abs(est - est2.numpy()).max() <= TOL_MAG
    assert (est[:, :, 0] == 0).all()                      # DC bin zero-padded (proposed.py:617)


@pytest.mark.parametrize("B,T", [(1, 1), (1, 2), (3, 33), (5, 9), (2, 70)])
def test_forward_mag_ragged_shapes(blob, oracles, B, T):
    """Frame counts that do not fill the time tiles, odd batch sizes, T = 1 (no history at all)."""
    from nunet_b200.synth import synth_clips
    o = oracles["causal_avg32"]
    wav = synth_clips(B, 512 + 256 * (T - 1), first_clip=7)
    mags, _ = o.stft(torch.from_numpy(wav))
    mag = mags[:, :, 1:].contiguous()
    with torch.no_grad():
        ref = o.net(mag[..., None]).squeeze(-1)
    eng = _engine(blob, max_frames=B * T + 5, ctfa_mode="causal_avg32")
    out = eng.forward_mag(mag.cuda()).cpu()
    assert float((out - ref).abs().max()) <= TOL_MAG


def test_wav_length_not_multiple_of_hop(blob, oracles):
    from nunet_b200.synth import synth_clips
    wav = synth_clips(2, 512 + 256 * 6 + 131, first_clip=3)   # trailing 131 samples are ignored by the STFT
    eng = _engine(blob, max_frames=2 * 7)
    y, est = eng.forward_wav(torch.from_numpy(wav).cuda())
    with torch.no_grad():
        y2, est2 = oracles["causal_avg32"].forward_wav(wav)
    assert y.shape == y2.shape and est.shape == est2.shape
    assert float((est.cpu() - est2).abs().max()) <= TOL_MAG
    assert float((y.cpu() - y2).abs().max()) <= TOL_WAV


def test_host_call_equals_device_call(blob):
    from nunet_b200.synth import synth_clips
    wav = synth_clips(3, 512 + 256 * 15, first_clip=11)
    eng = _engine(blob, max_frames=3 * 16)
    y, est = eng.forward_wav(torch.from_numpy(wav).cuda())
    out_wav = torch.empty(y.shape, dtype=torch.float32).pin_memory()
    out_mag = torch.empty(est.shape, dtype=torch.float32).pin_memory()
    eng.forward_wav_host(torch.from_numpy(wav).pin_memory(), out_wav, out_mag)
    assert torch.equal(out_wav, y.cpu()) and torch.equal(out_mag, est.cpu())


def test_batch_independence_and_causality(blob):
    """Size-independent properties: a clip's output does not depend on its batch neighbours, and frame t does
    not depend on samples after frame t (the whole model is causal)."""
    from nunet_b200.synth import synth_clips
    T = 48
    wav = synth_clips(6, 512 + 256 * (T - 1), first_clip=20)
    eng = _engine(blob, max_frames=6 * T)
    _, est = eng.forward_wav(torch.from_numpy(wav).cuda(), want_wav=False)
    _, est1 = eng.forward_wav(torch.from_numpy(wav[4:5]).cuda(), want_wav=False)
    assert torch.equal(est[4:5], est1)
    wav2 = wav.copy()
    wav2[:, 512 + 256 * 30:] = 0.0                     # destroy everything after frame 30's window
    _, est2 = eng.forward_wav(torch.from_numpy(wav2).cuda(), want_wav=False)
    assert torch.equal(est[:, :31], est2[:, :31])
    assert not torch.equal(est[:, 32:], est2[:, 32:])


def test_capacity_and_argument_errors(blob):
    from nunet_b200._lib import NunetError
    eng = _engine(blob, max_frames=8)
    out = eng.forward_mag(torch.zeros(3, 3, 256, device="cuda"))    # 9 frames on an 8-frame arena: two sub-batches, no error
    assert out.shape == (3, 3, 256) and bool(torch.isfinite(out).all())
    with pytest.raises(ValueError):
        eng.forward_mag(torch.zeros(3, 3, 255, device="cuda"))
    with pytest.raises(NunetError):
        eng.stream_step_mag(torch.zeros(1, 256, device="cuda"))     # streaming disabled
    with pytest.raises(NunetError) as ei:
        _engine(b"garbage" * 10, max_frames=8)
    assert "magic" in str(ei.value)
    with pytest.raises(NunetError):
        eng.forward_wav(torch.zeros(1, 300, device="cuda"))         # shorter than one frame


# --------------------------------------------------------------------------------------------- streaming
def test_streaming_step_mag_matches_frame_graph(blob, oracles):
    """S streams x 12 steps of the one-frame stateful graph (converter_proposed.py:188-867) vs the oracle's
    frame_step with carried history; also checks every exported history tensor after the last step."""
    from nunet_b200.synth import synth_clips
    o = oracles["frame_div32"]
    S, steps = 3, 12
    wav = synth_clips(S, 512 + 256 * (steps - 1), first_clip=40)
    mags, _ = o.stft(torch.from_numpy(wav))
    mag = mags[:, :, 1:].contiguous()                   # [S, steps, 256]
    eng = _engine(blob, max_streams=S)
    eng.stream_reset()
    state = o.zero_state(S)
    worst = 0.0
    for t in range(steps):
        feed = {"input": mag[:, t].reshape(S, 1, 256, 1)}
        feed.update({k.replace("_cur", "_prev"): v for k, v in state.items()})
        with torch.no_grad():
            res = o.frame_step(feed)
        ref = res.pop("model_out").reshape(S, 256)
        state = res
        out = eng.stream_step_mag(mag[:, t].contiguous().cuda()).cpu()
        worst = max(worst, float((out - ref).abs().max()))
    assert worst <= TOL_MAG
    from nunet_b200.interpreter import _engine_to_ref
    names = eng.state_names()
    assert len(names) == 130
    for n in names:
        ref = state[_engine_to_ref(n, "cur")]
        for s in (0, S - 1):
            got = eng.state_export(s, n)
            assert np.abs(got - ref[s].reshape(-1).numpy()).max() <= TOL_MAG, n


def test_streaming_with_ctfa_history_equals_offline(blob):
    """Extension flag: with 31 frames of TA carried per stream the streaming engine computes the offline graph."""
    from nunet_b200.synth import synth_clips
    S, T = 2, 45
    wav = synth_clips(S, 512 + 256 * (T - 1), first_clip=50)
    off = _engine(blob, max_frames=S * T, ctfa_mode="causal_avg32")
    _, est = off.forward_wav(torch.from_numpy(wav).cuda(), want_wav=False)
    # magnitudes exactly as the offline engine saw them
    mag = torch.from_numpy(off.debug_read("mag").reshape(S, T, 256)).cuda()
    eng = _engine(blob, max_streams=S, stream_ctfa_history=True)
    eng.stream_reset()
    outs = [eng.stream_step_mag(mag[:, t].contiguous()) for t in range(T)]
    out = torch.stack(outs, dim=1)
    # offline runs box-loaded tiles and CTA pairs, streaming the slot-table variants of the same split-half kernel: same
    # function, different summation order in the 128-channel units
    assert float((out - est[:, :, 1:]).abs().max()) <= TOL_MAG


def test_streaming_wav_loop_matches_interpreter_loop(blob, oracles, weights):
    """real_time_speech_enhancer (interpreter_proposed.py:15-370): hop in, hop out, DC 'edge' padding."""
    from nunet_b200.interpreter import Interpreter, real_time_speech_enhancer
    from nunet_b200.synth import synth_clips
    noisy = synth_clips(1, 256 * 14, first_clip=60)[0]
    mags = []
    with torch.no_grad():
        ref, _ = oracles["frame_div32"].real_time_speech_enhancer(noisy, dc_pad="edge", collect_mag=mags)
    it = Interpreter(weights=weights)
    out, times = real_time_speech_enhancer(noisy, it)
    assert out.shape == ref.shape and len(times) == 13
    assert np.abs(out - ref).max() <= TOL_WAV


def test_signature_runner_contract(weights, oracles):
    """131 tensors in, 131 out, reference names; foreign history arrays are honoured (import path)."""
    from nunet_b200.interpreter import Interpreter
    from nunet_b200.state_table import STATE_SHAPES
    o = oracles["frame_div32"]
    it = Interpreter(weights=weights)
    it.allocate_tensors()
    sig = it.get_signature_list()
    assert list(sig) == ["nutls_lstm_sm"] and len(sig["nutls_lstm_sm"]["inputs"]) == 131
    run = it.get_signature_runner("nutls_lstm_sm")
    rng = np.random.default_rng(5)
    state = {k: np.zeros(s, np.float32) for k, s in STATE_SHAPES.items()}
    ostate = o.zero_state(1)
    for step in range(4):
        x = rng.uniform(0, 30, (1, 1, 256, 1)).astype(np.float32)
        feed = {k.replace("_cur", "_prev"): v for k, v in state.items()}
        if step == 2:   # hand back copies: forces the import path
            feed = {k: v.copy() for k, v in feed.items()}
        out = run(input=x, **feed)
        assert len(out) == 131 and out["model_out"].shape == (1, 1, 256, 1)
        ofeed = {"input": torch.from_numpy(x)}
        ofeed.update({k.replace("_cur", "_prev"): v for k, v in ostate.items()})
        with torch.no_grad():
            ores = o.frame_step(ofeed)
        assert np.abs(out["model_out"] - ores["model_out"].numpy()).max() <= TOL_MAG
        ostate = {k: v for k, v in ores.items() if k != "model_out"}
        state = {k: v for k, v in out.items() if k != "model_out"}
        for k in ("msfe6_ee_cur1", "msfe4_dd2_cur3", "msfe6_de_cur1", "state_h", "msfe3_de_c"):
            assert out[k].shape == tuple(STATE_SHAPES[k])
            assert np.abs(out[k] - ostate[k].numpy()).max() <= TOL_MAG, k
    with pytest.raises(ValueError):
        run(input=x)                                   # missing history tensors
    with pytest.raises(ValueError):
        it.get_signature_runner("nope")


def test_models_surface(weights, golden_io):
    """models.NUTLS_LSTM(opt).build_model() -> model(x, training=False) (test_interface.py:45,58)."""
    from nunet_b200 import models
    from nunet_b200.options import default_options
    m = models.NUTLS_LSTM(default_options())
    model = m.build_model()
    model.load_weights(weights)
    y = model(golden_io["wav"], training=False)
    assert isinstance(y, np.ndarray)
    assert np.abs(y - golden_io["wav_causal_avg32"]).max() <= TOL_WAV
    with pytest.raises(NotImplementedError):
        model(golden_io["wav"], training=True)
    fm = m.tflite_model().load_weights(weights)
    out = fm(np.full((1, 1, 256, 1), 3.0, np.float32))
    assert out.shape == (1, 1, 256, 1) and np.isfinite(out).all()


def test_enhancement_quality_on_reference_excerpt(blob):
    """Functional pin: on the reference's own noisy/clean pair the enhanced output must be much closer to clean
    than the noisy input is."""
    import os
    from oracle.nunet_oracle import min_max_norm, si_sdr
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "wav_excerpt.npz"))
    noisy = g["noisy"].astype(np.float64) / 32768.0
    clean = g["clean"].astype(np.float64) / 32768.0
    x = min_max_norm(noisy).astype(np.float32)[None]
    eng = _engine(blob, max_frames=128)
    y, _ = eng.forward_wav(torch.from_numpy(x).cuda())
    y = y.cpu().numpy()[0]
    before, after = si_sdr(clean[:len(y)], x[0, :len(y)]), si_sdr(clean[:len(y)], y)
    assert after > before + 8.0, (before, after)


# ---------------------------------------------------------------------------------------------------
# Parity with the reference's DEPLOYED graph: the shipped nutls_lstm.tflite executed by oracle/tflite_graph.py
# (committed fixture tests/golden/o2_lstm.npz), engine loaded with the same (dequantised) weights.
def test_streaming_engine_matches_shipped_tflite_graph(tflite_weights, golden_o2):
    from nunet_b200.weights import pack_blob
    mag = golden_o2["mag"]
    eng = _engine(pack_blob(tflite_weights), max_streams=2, ctfa_mode="frame_div32")
    eng.stream_reset()
    outs = []
    for t in range(mag.shape[0]):
        m = torch.from_numpy(np.stack([mag[t], mag[t]])).cuda()       # two identical streams
        outs.append(eng.stream_step_mag(m).cpu().numpy())
    est = np.stack(outs)                                               # [T, 2, 256]
    ref = golden_o2["model_out"]
    assert ref.max() > 20.0
    assert np.abs(est[:, 0] - ref).max() <= TOL_MAG
    assert (est[:, 0] == est[:, 1]).all()
    for k in ("msfe6_ee_cur1", "msfe4_dd2_cur3", "state_h", "msfe5_en_c"):
        name = k.replace("_cur", "_") if "_cur" in k else k
        got = eng.state_export(1, name)
        assert np.abs(got - golden_o2[f"state_{k}"].reshape(-1)).max() <= TOL_MAG, k


def test_offline_engine_matches_shipped_tflite_graph(tflite_weights, golden_o2):
    """Zero history == zero padding: the offline tensor-core path in frame_div32 mode computes the deployed graph."""
    from nunet_b200.weights import pack_blob
    mag = torch.from_numpy(golden_o2["mag"])[None].contiguous().cuda()
    eng = _engine(pack_blob(tflite_weights), max_frames=mag.shape[1], ctfa_mode="frame_div32")
    est = eng.forward_mag(mag).cpu().numpy()[0]
    assert np.abs(est - golden_o2["model_out"]).max() <= TOL_MAG


# ---------------------------------------------------------------------------------------------------
# dilated-dense baseline (BASELINE configs[3]): offline engine vs the shipped nutls.tflite graph and vs the oracle
def test_ddb_offline_engine_matches_shipped_tflite_graph(ddb_weights, golden_o2_ddb):
    from nunet_b200._lib import NUNET_VARIANT_DDB
    from nunet_b200.weights import VARIANT_DDB, pack_blob
    mag = torch.from_numpy(golden_o2_ddb["mag"])[None].contiguous().cuda()
    eng = _engine(pack_blob(ddb_weights, VARIANT_DDB), max_frames=mag.shape[1], ctfa_mode="frame_div32", variant=NUNET_VARIANT_DDB)
    est = eng.forward_mag(mag).cpu().numpy()[0]
    ref = golden_o2_ddb["model_out"]
    assert np.abs(est - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), np.abs(est - ref).max()


@pytest.mark.parametrize("B,T", [(2, 40), (3, 7), (1, 1)])
def test_ddb_offline_engine_matches_oracle_random_weights(B, T):
    """Seeded random weights give O(1..10) activations in every layer (the shipped baseline file barely lights up the
    network), both CTFA modes, batch > 1 (the dilation must not read across clip boundaries)."""
    from nunet_b200._lib import NUNET_VARIANT_DDB
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import VARIANT_DDB, pack_blob, random_ddb_weights
    from oracle.nunet_oracle import Oracle
    w = random_ddb_weights(3)
    wav = synth_clips(B, 512 + 256 * (T - 1), first_clip=11)
    for mode in ("causal_avg32", "frame_div32"):
        o = Oracle(w, ctfa_mode=mode, variant="ddb")
        with torch.no_grad():
            y_ref, est_ref = o.forward_wav(wav)
        eng = _engine(pack_blob(w, VARIANT_DDB), max_frames=B * T, ctfa_mode=mode, variant=NUNET_VARIANT_DDB)
        y, est = eng.forward_wav(torch.from_numpy(wav).cuda())
        scale = max(1.0, float(est_ref.abs().max()) / 48.0)       # the 1e-3 bar is stated for peaks of ~48
        assert float((est.cpu() - est_ref).abs().max()) <= TOL_MAG * scale
        assert float((y.cpu() - y_ref).abs().max()) <= TOL_WAV * scale
        eng.close()


def test_ddb_models_surface(ddb_weights, golden_o2_ddb):
    """models.NUTLS(opt).build_model() (nunet_tls.py:1007); no streaming form yet."""
    from nunet_b200 import models
    from nunet_b200.options import default_options
    m = models.NUTLS(default_options())
    model = m.build_model().load_weights(ddb_weights)
    from nunet_b200.synth import synth_clips
    y = model(synth_clips(1, 512 + 256 * 9), training=False)
    assert y.shape == (1, 9 * 256 + 512) and np.isfinite(y).all()
    fm = m.tflite_model().load_weights(ddb_weights)
    out = fm(np.full((1, 1, 256, 1), 3.0, np.float32))
    assert out.shape == (1, 1, 256, 1) and np.isfinite(out).all()


def test_ddb_streaming_engine_matches_shipped_tflite_graph(ddb_weights, golden_o2_ddb):
    """One-frame stateful form of the dilated-dense baseline (converter_nunet_tls.py:292-1526) vs the shipped
    nutls.tflite executed frame by frame: 72 steps exercise the 32-step look-back of the deepest layer."""
    from nunet_b200._lib import NUNET_VARIANT_DDB
    from nunet_b200.weights import VARIANT_DDB, pack_blob
    mag = golden_o2_ddb["mag"]
    eng = _engine(pack_blob(ddb_weights, VARIANT_DDB), max_streams=2, ctfa_mode="frame_div32", variant=NUNET_VARIANT_DDB)
    eng.stream_reset()
    outs = []
    for t in range(mag.shape[0]):
        outs.append(eng.stream_step_mag(torch.from_numpy(np.stack([mag[t], mag[t]])).cuda()).cpu().numpy())
    est = np.stack(outs)
    ref = golden_o2_ddb["model_out"]
    assert np.abs(est[:, 0] - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), np.abs(est[:, 0] - ref).max()
    assert (est[:, 0] == est[:, 1]).all()


def test_ddb_streaming_state_contract_random_weights():
    """40 steps of S = 3 streams vs the oracle's one-frame graph with carried history; then every one of the 208
    reference-named history tensors (interpreter_nunet_tls.py:36-289) is exported and compared, and a foreign history
    is imported into one stream."""
    from nunet_b200._lib import NUNET_VARIANT_DDB
    from nunet_b200.interpreter import _engine_to_ref
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import VARIANT_DDB, pack_blob, random_ddb_weights
    from oracle.nunet_oracle import Oracle
    w = random_ddb_weights(5)
    S, T = 3, 40
    o = Oracle(w, ctfa_mode="frame_div32", variant="ddb")
    mags, _ = o.stft(torch.from_numpy(synth_clips(S, 512 + 256 * (T - 1), first_clip=21)))
    mag = mags[:, :, 1:].contiguous()
    eng = _engine(pack_blob(w, VARIANT_DDB), max_streams=S, ctfa_mode="frame_div32", variant=NUNET_VARIANT_DDB)
    eng.stream_reset()
    state = o.zero_state(S)
    worst, peak = 0.0, 0.0
    for t in range(T):
        feed = {"input": mag[:, t].reshape(S, 1, 256, 1)}
        feed.update({k.replace("_cur", "_prev"): v for k, v in state.items()})
        with torch.no_grad():
            res = o.frame_step(feed)
        ref = res.pop("model_out").reshape(S, 256)
        state = res
        out = eng.stream_step_mag(mag[:, t].contiguous().cuda()).cpu()
        worst = max(worst, float((out - ref).abs().max()))
        peak = max(peak, float(ref.abs().max()))
    scale = max(1.0, peak / 48.0)
    assert worst <= TOL_MAG * scale, (worst, peak)
    names = eng.state_names()
    assert len(names) == 208
    for n in names:
        ref = state[_engine_to_ref(n, "cur")]
        for s_ in (0, S - 1):
            got = eng.state_export(s_, n)
            assert np.abs(got - ref[s_].reshape(-1).numpy()).max() <= TOL_MAG * scale, n
    # import: stream 0's history into stream 1, then both must produce the same next frame
    for n in names:
        eng.state_import(1, n, eng.state_export(0, n))
    nxt = torch.stack([mag[0, T - 1], mag[0, T - 1], mag[2, T - 1]]).contiguous().cuda()
    y = eng.stream_step_mag(nxt).cpu().numpy()
    assert np.abs(y[0] - y[1]).max() <= 1e-5 * scale


def test_ddb_signature_runner_contract(ddb_weights, golden_o2_ddb):
    """Interpreter(...).get_signature_runner('nutls') with 209 keyword tensors (interpreter_nunet_tls.py:543-549)."""
    from nunet_b200.interpreter import Interpreter
    it = Interpreter(weights=ddb_weights, variant="ddb")
    it.allocate_tensors()
    sig = it.get_signature_list()
    assert list(sig) == ["nutls"] and len(sig["nutls"]["inputs"]) == 209 and len(sig["nutls"]["outputs"]) == 209
    run = it.get_signature_runner("nutls")
    with pytest.raises(ValueError):
        it.get_signature_runner("nutls_lstm_sm")
    from nunet_b200.state_table import STATE_SHAPES_DDB
    state = {k.replace("_cur", "_prev"): np.zeros(s, np.float32) for k, s in STATE_SHAPES_DDB.items()}
    mag = golden_o2_ddb["mag"]
    for t in range(3):
        out = run(input=mag[t].reshape(1, 1, 256, 1), **state)
        assert np.abs(out["model_out"].reshape(256) - golden_o2_ddb["model_out"][t]).max() <= 1e-4 * 2
        state = {k.replace("_cur", "_prev"): v for k, v in out.items() if k != "model_out"}
    assert out["ddb_cur6"].shape == (1, 32, 4, 192) and out["msfe3_en_ddb_cur_in"].shape == (1, 1, 1, 32)


@pytest.mark.parametrize("variant", ["lstm", "ddb"])
def test_streaming_graph_with_parallel_chains_equals_plain_launches(blob, ddb_weights, variant, monkeypatch):
    """The streaming step replayed as a CUDA graph -- here also split into two parallel chains over disjoint stream
    ranges (NUNET_STREAM_SPLIT) -- computes exactly what kernel-by-kernel launches compute (wav and mag entry points,
    both variants)."""
    from nunet_b200._lib import NUNET_VARIANT_DDB, NUNET_VARIANT_LSTM
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import VARIANT_DDB, pack_blob
    S, steps = 160, 7
    if variant == "ddb":
        b, v = pack_blob(ddb_weights, VARIANT_DDB), NUNET_VARIANT_DDB
    else:
        b, v = blob, NUNET_VARIANT_LSTM
    wav = np.tile(synth_clips(8, 256 * steps, first_clip=90), (S // 8, 1))
    wav *= np.linspace(0.5, 1.0, S, dtype=np.float32)[:, None]          # every stream different
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NUNET_DEBUG_KNOBS", "1")
        monkeypatch.setenv("NUNET_STREAM_GRAPH", mode)
        monkeypatch.setenv("NUNET_STREAM_SPLIT", "2")      # two chains of 80 streams inside the captured step
        eng = NunetEngine(b, max_streams=S, variant=v)
        eng.stream_reset()
        ys = []
        for t in range(steps):
            ys.append(eng.stream_step_wav(torch.from_numpy(wav[:, 256 * t:256 * (t + 1)]).cuda()).cpu().numpy().copy())
        mags = torch.from_numpy(np.abs(wav[:, :256]) * 20).cuda()
        ys.append(eng.stream_step_mag(mags).cpu().numpy().copy())
        ys.append(eng.stream_step_mag(mags).cpu().numpy().copy())
        outs[mode] = ys
        eng.close()
    for a, g in zip(outs["0"], outs["1"]):
        assert np.isfinite(g).all() and np.abs(a - g).max() <= 1e-6


@pytest.mark.parametrize("knob", ["NUNET_TC3_TWIN", "NUNET_CTFA_GATE4"])
@pytest.mark.parametrize("variant", ["lstm", "ddb"])
def test_kernel_variants_do_not_change_results(blob, ddb_weights, variant, knob, monkeypatch):
    """Two round-2 changes are pure re-arrangements and must not move a bit -- whole clips and time chunks:
    NUNET_TC3_TWIN: inner stride-2 convs read an [even | odd] copy of their inputs that the producers write next to the
    bin-ordered one (Tc3Params::out2); switched off they read the bin-ordered tensors through strided tensor-map boxes.
    NUNET_CTFA_GATE4: the CTFA gate kernel handles four frames per warp with packed fp32x2 arithmetic in the summation order of
    the one-frame-per-warp kernel it replaces."""
    from nunet_b200._lib import NUNET_VARIANT_DDB, NUNET_VARIANT_LSTM
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import VARIANT_DDB, pack_blob
    if variant == "ddb":
        b, v = pack_blob(ddb_weights, VARIANT_DDB), NUNET_VARIANT_DDB
    else:
        b, v = blob, NUNET_VARIANT_LSTM
    B, T = 5, 45
    wav = torch.from_numpy(synth_clips(B, 512 + 256 * (T - 1), first_clip=300)).cuda()
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NUNET_DEBUG_KNOBS", "1")
        monkeypatch.setenv(knob, mode)
        eng = NunetEngine(b, max_frames=B * T, variant=v)
        y, est = eng.forward_wav(wav)
        cut = NunetEngine(b, max_frames=B * T, variant=v, chunk_frames=16)
        y2, est2 = cut.forward_wav(wav)
        outs[mode] = [t.cpu().numpy().copy() for t in (y, est, y2, est2)]
        eng.close()
        cut.close()
    for a, g in zip(outs["1"], outs["0"]):
        assert np.isfinite(a).all() and np.array_equal(a, g)


def test_streaming_lstm_group_kernel_equals_per_stream_kernel(blob, monkeypatch):
    """A streaming step runs the LSTM bottlenecks on lstm_stream_kernel (16 streams per CTA share the weight reads); every sum
    is formed as in lstm_block_kernel (one CTA per stream, NUNET_LSTM_STREAM=0), so the two agree bit for bit -- 37 streams
    (two full groups and a partial one), 6 hops, states carried."""
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    S, steps = 37, 6
    wav = synth_clips(S, 256 * steps, first_clip=410)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("NUNET_DEBUG_KNOBS", "1")
        monkeypatch.setenv("NUNET_LSTM_STREAM", mode)
        eng = NunetEngine(blob, max_streams=S)
        eng.stream_reset()
        outs[mode] = [eng.stream_step_wav(torch.from_numpy(wav[:, 256 * t:256 * (t + 1)]).cuda()).cpu().numpy().copy() for t in range(steps)]
        eng.close()
    for a, g in zip(outs["1"], outs["0"]):
        assert np.isfinite(a).all() and np.array_equal(a, g)


def test_full_size_batch_properties(blob, oracles):
    """BASELINE configs[1] at FULL size (256 clips x 4 s = 63 744 frames, the bench workload), checked through properties
    that do not need an oracle run of that size: (a) copies of one clip anywhere in the batch come out bit-identical
    (tiles straddle clips, CTA pairs work on rows of different clips, boxes zero-fill the time pads -- none of it may
    leak between clips), (b) the first clips equal a 2-clip run of another engine bit for bit, (c) those two clips match
    the oracle within the bar, (d) everything is finite."""
    from nunet_b200.synth import synth_clips
    B, N, T = 256, 64000, 249
    pool = synth_clips(32, N, first_clip=200)
    wav = np.tile(pool, (B // 32, 1))
    eng = _engine(blob, max_frames=B * T)
    y, est = eng.forward_wav(torch.from_numpy(wav).cuda())
    y, est = y.cpu().numpy(), est.cpu().numpy()
    eng.close()
    assert est.shape == (B, T, 257) and np.isfinite(est).all() and np.isfinite(y).all()
    for k in range(1, B // 32):
        assert np.array_equal(est[:32], est[32 * k:32 * (k + 1)]), k
        assert np.array_equal(y[:32], y[32 * k:32 * (k + 1)]), k
    small = _engine(blob, max_frames=2 * T)
    y2, est2 = small.forward_wav(torch.from_numpy(wav[:2]).cuda())
    assert np.array_equal(est[:2], est2.cpu().numpy()) and np.array_equal(y[:2], y2.cpu().numpy())
    with torch.no_grad():
        y_ref, est_ref = oracles["causal_avg32"].forward_wav(wav[:2])
    assert np.abs(est[:2] - est_ref.numpy()).max() <= TOL_MAG
    assert np.abs(y[:2] - y_ref.numpy()).max() <= TOL_WAV


def test_full_size_streaming_properties(blob):
    """BASELINE configs[2] at FULL size (1024 concurrent streams): copies of one stream anywhere among the 1024 stay
    bit-identical over 8 hops (graph replay included), and equal the same streams run alone on a 16-stream engine."""
    from nunet_b200.synth import synth_clips
    S, steps = 1024, 8
    pool = synth_clips(16, 256 * steps, first_clip=300)
    wav = np.tile(pool, (S // 16, 1))
    big, small = _engine(blob, max_streams=S), _engine(blob, max_streams=16)
    big.stream_reset()
    small.stream_reset()
    for t in range(steps):
        hop = wav[:, 256 * t:256 * (t + 1)]
        yb = big.stream_step_wav(torch.from_numpy(hop).cuda()).cpu().numpy()
        ys = small.stream_step_wav(torch.from_numpy(hop[:16]).cuda()).cpu().numpy()
        assert np.isfinite(yb).all()
        for k in range(1, S // 16):
            assert np.array_equal(yb[:16], yb[16 * k:16 * (k + 1)]), (t, k)
        assert np.array_equal(yb[:16], ys), t
