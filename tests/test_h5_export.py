"""Artifact emission: the engine's weights written back as the reference's Keras `.h5`."""
import numpy as np
import pytest


def test_h5_writer_roundtrip_small(tmp_path):
    from nunet_b200.h5_reader import read_h5, read_h5_attrs
    from nunet_b200.h5_writer import write_h5
    rng = np.random.default_rng(0)
    ds = {"/a/a/kernel:0": rng.standard_normal((2, 3, 4, 5)).astype(np.float32), "/a/a/bias:0": rng.standard_normal(5).astype(np.float32),
          "/b/b/cell_1/kernel:0": rng.standard_normal((7, 84)).astype(np.float32)}
    ds.update({f"/many/many/w{i}:0": np.full((i + 1,), float(i), np.float32) for i in range(37)})      # several symbol nodes
    attrs = {"": {"layer_names": ["a", "empty", "b", "many"], "backend": "tensorflow", "keras_version": "2.12.0"},
             "/a": {"weight_names": ["a/kernel:0", "a/bias:0"]}, "/empty": {"weight_names": []},
             "/b": {"weight_names": ["b/cell_1/kernel:0"]}, "/many": {"weight_names": [f"many/w{i}:0" for i in range(37)]}}
    path = str(tmp_path / "w.h5")
    write_h5(path, ds, attrs)
    back = read_h5(path)
    assert set(back) == set(ds)
    for k, v in ds.items():
        assert back[k].dtype == np.float32 and np.array_equal(back[k], v), k
    got = read_h5_attrs(path)
    for g, a in attrs.items():
        assert got[g] == a, g
    assert "/empty" in got and got["/a/a"] == {}


def test_export_reproduces_the_reference_inventory_and_reimports_bit_exactly(weights, tmp_path):
    """weights -> nutls_lstm.h5 -> weights: same role-named set, same engine blob; the file lists the same layers, weight
    names and dataset shapes as the file Keras wrote (table extracted from it: data/keras_layout_lstm.json)."""
    from nunet_b200.h5_reader import read_h5, read_h5_attrs
    from nunet_b200.keras_export import export_lstm_h5, load_layout
    from nunet_b200.weights import lstm_weights_from_h5, pack_blob
    path = str(tmp_path / "nutls_lstm.h5")
    export_lstm_h5(weights, path)
    layout = load_layout()
    raw = read_h5(path)
    assert len(raw) == len(layout["datasets"]) == 772
    for d in layout["datasets"]:
        assert list(raw[d["path"]].shape) == d["shape"], d["path"]
    attrs = read_h5_attrs(path)
    assert sorted(attrs) == layout["groups"]
    assert len(attrs[""]["layer_names"]) == 340 and attrs[""]["backend"] == "tensorflow" and attrs[""]["keras_version"] == "2.12.0"
    for g, a in layout["attrs"].items():
        assert attrs[g] == a, g
    again = lstm_weights_from_h5(path)
    assert set(again) == set(weights)
    for k in weights:
        assert np.array_equal(again[k], np.asarray(weights[k], np.float32).reshape(again[k].shape)), k
    assert pack_blob(again) == pack_blob(weights)
    bad = dict(weights)
    bad.pop("out_conv/kernel")
    with pytest.raises(KeyError):
        export_lstm_h5(bad, str(tmp_path / "bad.h5"))


def test_model_save_weights_mirrors_keras(weights, tmp_path):
    """models.NUTLS_LSTM(opt).build_model().load_weights(set).save_weights(path) -> a file load_weights accepts again
    (no GPU needed: the engine is only created on the first call)."""
    from types import SimpleNamespace
    from nunet_b200 import models
    from nunet_b200.weights import pack_blob
    opt = SimpleNamespace(win_len=512, fft_len=512, hop_len=256, batch_size=2, chunk_size=16000)
    m = models.NUTLS_LSTM(opt).build_model()
    with pytest.raises(RuntimeError):
        m.save_weights(str(tmp_path / "x.h5"))
    m.load_weights(weights)
    path = str(tmp_path / "saved.h5")
    m.save_weights(path)
    m2 = models.NUTLS_LSTM(opt).build_model().load_weights(path)
    assert m2._blob == pack_blob(weights)
