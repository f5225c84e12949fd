"""Offline calls of any size on a fixed arena: sub-batches of whole clips and time chunks with carried history (conv rows,
LSTM h / c, the 31-frame attention window) -- the engine's form of the reference's memory bound (`options.py:42` chunk_size;
the one-frame graph `converter_proposed.py:188-867` is the chunk = 1 limit).  Chunked results must be BIT-IDENTICAL to
unchunked ones; BASELINE configs[4]'s 1024 clips per GPU as 4 x 256 and 8192 clips on one GPU run on the 256-clip arena."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def blob(weights):
    from nunet_b200.weights import pack_blob
    return pack_blob(weights)


def _engine(blob, **kw):
    from nunet_b200.engine import NunetEngine
    return NunetEngine(blob, **kw)


@pytest.mark.parametrize("mode", ["causal_avg32", "frame_div32"])
@pytest.mark.parametrize("chunk", [1, 7, 16, 40])
def test_time_chunks_are_bit_identical(blob, mode, chunk):
    """3 clips x 70 frames whole, and in time chunks of 1 / 7 / 16 / 40 frames (the last chunk ragged; 7 and 16 are shorter
    than the 32-frame attention window, 40 longer; 1 is the streaming limit)."""
    from nunet_b200.synth import synth_clips
    B, T = 3, 70
    wav = torch.from_numpy(synth_clips(B, 512 + 256 * (T - 1), first_clip=500)).cuda()
    whole = _engine(blob, max_frames=B * T, ctfa_mode=mode)
    y0, e0 = whole.forward_wav(wav)
    n0 = whole.last_launch_count
    cut = _engine(blob, max_frames=B * T, ctfa_mode=mode, chunk_frames=chunk)
    y1, e1 = cut.forward_wav(wav)
    assert cut.last_launch_count > n0
    mag = e0[:, :, 1:].contiguous().abs()
    m0, m1 = whole.forward_mag(mag), cut.forward_mag(mag)     # the network alone (magnitudes in / out) takes the same path
    if chunk == 1:
        # one-frame chunks run the streaming kernel variants (no tensor-map boxes, so no CTA pairs: the 128-channel units add
        # their three partial products in another order) -- equal to fp32 rounding, not bit for bit
        assert float((e0 - e1).abs().max()) <= 1e-4 and float((y0 - y1).abs().max()) <= 1e-4
        assert float((m0 - m1).abs().max()) <= 1e-4
        return
    assert torch.equal(e0, e1)
    assert torch.equal(y0, y1)
    assert torch.equal(m0, m1)


def test_clip_longer_than_the_arena_and_sub_batches(blob):
    """max_frames smaller than one clip (time chunks of max_frames, one clip at a time) and smaller than the batch
    (sub-batches of whole clips): both bit-identical to an arena that holds the whole call."""
    from nunet_b200.synth import synth_clips
    B, T = 5, 90
    wav = torch.from_numpy(synth_clips(B, 512 + 256 * (T - 1), first_clip=520)).cuda()
    y0, e0 = _engine(blob, max_frames=B * T).forward_wav(wav)
    for cap in (64, 2 * T + 11):       # 64 < T: chunks of 64 + 26 frames per clip; 191: two clips per sub-batch
        y1, e1 = _engine(blob, max_frames=cap).forward_wav(wav)
        assert torch.equal(e0, e1) and torch.equal(y0, y1), cap
    # host entry point: staged through the handle's buffers in sub-batches (and grown for a clip that exceeds them)
    small = _engine(blob, max_frames=40)
    out = small.forward_wav_host(wav.cpu().numpy())[0]
    assert np.array_equal(out, y0.cpu().numpy())


def test_ten_minute_clip_on_a_small_arena(blob):
    """One 10-minute clip (37 499 frames) on an arena of 4096 frames: runs, is finite, and -- the model being causal --
    its first frames equal those of the clip's first 20 s processed alone."""
    from nunet_b200.synth import synth_clips
    n = 16000 * 600
    T = 1 + (n - 512) // 256
    piece = synth_clips(1, 16000 * 20, first_clip=540)[0]
    wav = np.tile(piece, 30)[None, :n].copy()
    eng = _engine(blob, max_frames=4096)
    y, est = eng.forward_wav(torch.from_numpy(wav).cuda())
    assert est.shape == (1, T, 257) and bool(torch.isfinite(est).all()) and bool(torch.isfinite(y).all())
    T20 = 1 + (16000 * 20 - 512) // 256
    y20, est20 = eng.forward_wav(torch.from_numpy(wav[:, :16000 * 20]).cuda())
    assert torch.equal(est[:, :T20], est20)
    assert torch.equal(y[:, :(T20 - 1) * 256], y20[:, :(T20 - 1) * 256])


def test_1024_clips_as_four_sub_batches_of_256(blob):
    """BASELINE configs[4]: 1024 clips per GPU on the 256-clip arena = one call, four sub-batches; bit-identical to four
    separate 256-clip calls (and copies of a clip agree wherever they sit)."""
    from nunet_b200.synth import synth_clips
    N, T = 64000, 249
    pool = synth_clips(32, N, first_clip=560)
    wav = torch.from_numpy(np.tile(pool, (32, 1))).cuda()          # 1024 clips
    wav[256:512] *= 0.5                                            # make the four quarters different
    wav[512:768] *= 0.25
    wav[768:] *= 0.125
    eng = _engine(blob, max_frames=256 * T)
    y, est = eng.forward_wav(wav)
    assert eng.last_launch_count > 4 * 150
    for q in range(4):
        yq, eq = eng.forward_wav(wav[256 * q:256 * (q + 1)].contiguous())
        assert torch.equal(est[256 * q:256 * (q + 1)], eq), q
        assert torch.equal(y[256 * q:256 * (q + 1)], yq), q
    assert torch.equal(est[:32], est[32:64])


def test_8192_clips_on_one_gpu(blob):
    """8192 clips x 4 s (2 040 000 frames) in one call on the 256-clip arena: no NUNET_ENOMEM, finite, copies identical."""
    from nunet_b200.synth import synth_clips
    N, T = 64000, 249
    pool = synth_clips(16, N, first_clip=580)
    wav = torch.from_numpy(pool).cuda().repeat(512, 1)
    eng = _engine(blob, max_frames=256 * T)
    y, _ = eng.forward_wav(wav, want_mag=False)
    assert y.shape == (8192, (T - 1) * 256 + 512) and bool(torch.isfinite(y).all())
    assert torch.equal(y[:16], y[-16:]) and torch.equal(y[:16], y[4000:4016])


def test_ddb_variant_sub_batches_and_time_chunks(ddb_weights):
    """Dilated-dense variant: sub-batches of whole clips, and time chunks whose history is the last 32 frames of every dilated
    layer's input (the deepest dilation) plus one row for the two causal (2,3) convs of each block -- bit-identical to one pass.
    Chunks of 5 and 20 frames are shorter than that history, 48 longer."""
    from nunet_b200._lib import NUNET_VARIANT_DDB
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import VARIANT_DDB, pack_blob
    blob = pack_blob(ddb_weights, VARIANT_DDB)
    B, T = 4, 100
    wav = torch.from_numpy(synth_clips(B, 512 + 256 * (T - 1), first_clip=600)).cuda()
    y0, e0 = NunetEngine(blob, max_frames=B * T, variant=NUNET_VARIANT_DDB).forward_wav(wav)
    y1, e1 = NunetEngine(blob, max_frames=T + 3, variant=NUNET_VARIANT_DDB).forward_wav(wav)     # one clip per sub-batch
    assert torch.equal(e0, e1) and torch.equal(y0, y1)
    for chunk in (5, 20, 48):
        y2, e2 = NunetEngine(blob, max_frames=B * T, variant=NUNET_VARIANT_DDB, chunk_frames=chunk).forward_wav(wav)
        assert torch.equal(e0, e2) and torch.equal(y0, y2), chunk
    y3, e3 = NunetEngine(blob, max_frames=64, variant=NUNET_VARIANT_DDB).forward_wav(wav)        # clip longer than the arena
    assert torch.equal(e0, e3) and torch.equal(y0, y3)
