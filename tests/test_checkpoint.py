"""Stream checkpoints: history of running streams under the reference's tensor names, restorable into
another engine or handed to a signature runner."""
import json

import numpy as np
import pytest
import torch


class _FakeEngine:
    """Host-only stand-in with the engine's history interface (names from the plan, values = a counter)."""

    def __init__(self, variant=0, streams=2):
        from nunet_b200.interpreter import _engine_to_ref
        from nunet_b200.state_table import STATE_SHAPES, STATE_SHAPES_DDB
        shapes = STATE_SHAPES_DDB if variant else STATE_SHAPES
        import re
        self._numel = {}
        for ref, sh in shapes.items():
            name = re.sub(r"_cur_(in|out)$", r"_\1", ref)
            name = re.sub(r"_cur(\d+)$", r"_\1", name)
            assert _engine_to_ref(name, "cur") == ref
            self._numel[name] = int(np.prod(sh))
        self._numel.update(in_buffer=512, out_buffer=512)
        self.variant, self.max_streams = variant, streams
        self.ctfa_mode, self.dc_mode, self.stream_ctfa_history = "frame_div32", "edge", False
        self.mem = {(s, n): np.full(k, 1000.0 * s + i, np.float32) for s in range(streams)
                    for i, (n, k) in enumerate(self._numel.items())}

    def state_names(self):
        return [n for n in self._numel if n not in ("in_buffer", "out_buffer")]

    def state_numel(self, name):
        return self._numel[name]

    def state_export(self, s, name):
        return self.mem[(s, name)].copy()

    def state_import(self, s, name, value):
        self.mem[(s, name)] = np.asarray(value, np.float32).reshape(-1).copy()


@pytest.mark.parametrize("variant,count", [(0, 130), (1, 208)])
def test_checkpoint_names_shapes_and_roundtrip(tmp_path, variant, count):
    from nunet_b200 import checkpoint
    from nunet_b200.state_table import STATE_SHAPES, STATE_SHAPES_DDB
    shapes = STATE_SHAPES_DDB if variant else STATE_SHAPES
    src = _FakeEngine(variant, streams=3)
    path = str(tmp_path / "ck.npz")
    checkpoint.save(src, path, stream_ids=[2, 0])
    with np.load(path) as z:
        ck = {k: z[k] for k in z.files}
    meta = json.loads(str(ck["meta"]))
    assert meta["format"] == checkpoint.FORMAT and meta["streams"] == [2, 0]
    assert len(ck) == count + 2 + 1
    for ref, sh in shapes.items():                       # reference names and shapes, stacked over streams
        assert ck[ref].shape == (2,) + tuple(sh), ref
    feed = checkpoint.signature_feed(ck, k=1)             # what a TFLite runner takes besides `input`
    assert len(feed) == count and all("_cur" not in k for k in feed)
    assert feed["msfe6_ee_prev1"].shape == (1, 1, 256, 64) if not variant else True
    dst = _FakeEngine(variant, streams=4)
    assert checkpoint.load(dst, path, stream_ids=[3, 1]) == [3, 1]
    for n in src._numel:
        assert np.array_equal(dst.mem[(3, n)], src.mem[(2, n)]) and np.array_equal(dst.mem[(1, n)], src.mem[(0, n)])
    with pytest.raises(ValueError):
        checkpoint.load(_FakeEngine(1 - variant), path)   # other variant
    with pytest.raises(ValueError):
        checkpoint.load(dst, path, stream_ids=[9, 1])     # target out of range
    with pytest.raises(ValueError):
        checkpoint.load(dst, path, stream_ids=[0])        # count mismatch


@pytest.mark.gpu
@pytest.mark.parametrize("history", [False, True])
def test_checkpoint_resume_continues_the_stream(tmp_path, weights, history):
    """A stream saved after 9 hops and restored into a NEW engine (other slot, other step count) continues where it
    left off.  Not bit-exact by construction: history rows live on the device as fp16 (hi, lo) pairs, the checkpoint
    holds their fp32 sum, and re-splitting a sum whose lo half is exactly half an ulp of hi picks the other
    representation of the same value -- hence 1e-5 here (the parity bar of the path is 1e-3)."""
    from nunet_b200 import checkpoint
    from nunet_b200.engine import NunetEngine
    from nunet_b200.synth import synth_clips
    from nunet_b200.weights import pack_blob
    blob = pack_blob(weights)
    S, n1, n2 = 2, 9, 6
    wav = synth_clips(S, 256 * (n1 + n2), first_clip=70)
    a = NunetEngine(blob, max_streams=S, stream_ctfa_history=history)
    a.stream_reset()
    hop = np.empty((S, 256), np.float32)
    for t in range(n1):
        a.stream_step_wav_host(wav[:, 256 * t:256 * (t + 1)], hop)
    path = str(tmp_path / "streams.npz")
    checkpoint.save(a, path)
    ref = np.stack([a.stream_step_wav_host(wav[:, 256 * t:256 * (t + 1)]).copy() for t in range(n1, n1 + n2)], 1)
    b = NunetEngine(blob, max_streams=4, stream_ctfa_history=history)
    b.stream_reset()
    for t in range(3):                                   # the new engine has a different step count and dirty slots
        b.stream_step_wav_host(wav[[0, 1, 0, 1], 256 * t:256 * (t + 1)])
    checkpoint.load(b, path, stream_ids=[3, 1])
    got = []
    for t in range(n1, n1 + n2):
        x = np.zeros((4, 256), np.float32)
        x[3], x[1] = wav[0, 256 * t:256 * (t + 1)], wav[1, 256 * t:256 * (t + 1)]
        y = b.stream_step_wav_host(x)
        got.append(np.stack([y[3], y[1]]))
    got = np.stack(got, 1)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-5


@pytest.mark.gpu
def test_checkpoint_migrates_to_the_frame_graph(weights):
    """Engine -> checkpoint -> signature feed -> the one-frame graph (oracle restatement of converter_proposed.py:188-867)
    continues the stream within the parity bar."""
    from nunet_b200 import checkpoint
    from nunet_b200.engine import NunetEngine
    from nunet_b200.weights import pack_blob
    from oracle.nunet_oracle import Oracle
    o = Oracle(weights, ctfa_mode="frame_div32")
    rng = np.random.default_rng(11)
    eng = NunetEngine(pack_blob(weights), max_streams=1)
    eng.stream_reset()
    mags = rng.uniform(0, 30, (8, 1, 256)).astype(np.float32)
    for t in range(5):
        eng.stream_step_mag(torch.from_numpy(mags[t]).cuda())
    ck = checkpoint.capture(eng)
    feed = {k: torch.from_numpy(v) for k, v in checkpoint.signature_feed(ck).items()}
    for t in range(5, 8):
        ours = eng.stream_step_mag(torch.from_numpy(mags[t]).cuda()).cpu().numpy()
        feed["input"] = torch.from_numpy(mags[t].reshape(1, 1, 256, 1))
        with torch.no_grad():
            res = o.frame_step(feed)
        assert np.abs(ours.reshape(-1) - res.pop("model_out").numpy().reshape(-1)).max() <= 1e-3
        feed = {k.replace("_cur", "_prev"): v for k, v in res.items()}
