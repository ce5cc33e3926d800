"""Burn `.mpk` record reading/writing: layout against az_weight_name, round trip, tolerant reader."""
import msgpack
import numpy as np
import pytest

import alphazero_chess_b200 as az
from alphazero_chess_b200 import mpk


def test_layout_matches_the_abi():
    layout = mpk.weight_layout()
    assert [n for n, _ in layout] == az.weight_names()
    w = az.random_weights(seed=0)
    assert [int(np.prod(s)) for _, s in layout] == [a.size for a in w]


def test_round_trip(tmp_path):
    w = az.random_weights(seed=5, randomize_bn=True)
    path = tmp_path / "model.mpk"
    mpk.save_mpk(path, w)
    got = mpk.load_mpk(path)
    assert len(got) == 144 and all(np.array_equal(a, b) for a, b in zip(got, w))
    rec = msgpack.unpackb(path.read_bytes(), raw=False)
    assert set(rec) == {"metadata", "item"}
    conv = rec["item"]["res_blocks"][3]["conv2"]
    assert conv["weight"]["param"]["shape"] == [128, 128, 3, 3] and conv["weight"]["param"]["dtype"] == "F32"
    assert isinstance(conv["weight"]["param"]["bytes"], bytes) and conv["stride"] == [None, None]
    assert rec["item"]["value_linear_1"]["weight"]["param"]["shape"] == [512, 64]      # burn's [d_input, d_output]


def test_reader_accepts_other_element_types_and_legacy_tensors(tmp_path):
    w = az.random_weights(seed=6)
    path = tmp_path / "model.mpk"
    mpk.save_mpk(path, w)
    rec = msgpack.unpackb(path.read_bytes(), raw=False)
    b = rec["item"]["input_bn"]["gamma"]["param"]
    b["bytes"] = list(np.frombuffer(b["bytes"], "<f4").astype("<f8").tobytes())       # bytes as a sequence of integers, f64
    b["dtype"] = "F64"
    v = rec["item"]["value_linear_2"]["bias"]
    rec["item"]["value_linear_2"]["bias"] = {"id": "x", "param": {"value": [0.25], "shape": [1]}}
    h = rec["item"]["policy_bn"]["beta"]["param"]
    f32 = np.frombuffer(h["bytes"], "<f4")
    h["bytes"] = (f32.view(np.uint32) >> 16).astype("<u2").tobytes()
    h["dtype"] = "BF16"
    path.write_bytes(msgpack.packb(rec, use_bin_type=True))
    got = mpk.load_mpk(path)
    names = az.weight_names()
    assert np.array_equal(got[names.index("input_bn.gamma")], w[names.index("input_bn.gamma")])
    assert got[names.index("value_linear_2.bias")][0] == 0.25 and v is not None
    want = (f32.view(np.uint32) & 0xFFFF0000).view(np.float32)
    assert np.array_equal(got[names.index("policy_bn.beta")], want)


def test_missing_field_and_wrong_shape_are_errors(tmp_path):
    w = az.random_weights(seed=6)
    path = tmp_path / "model.mpk"
    mpk.save_mpk(path, w)
    rec = msgpack.unpackb(path.read_bytes(), raw=False)
    del rec["item"]["value_bn"]["running_var"]
    path.write_bytes(msgpack.packb(rec, use_bin_type=True))
    with pytest.raises(ValueError, match="value_bn.running_var"):
        mpk.load_mpk(path)
    mpk.save_mpk(path, w)
    rec = msgpack.unpackb(path.read_bytes(), raw=False)
    rec["item"]["input_conv"]["weight"]["param"]["shape"] = [19, 128, 3, 3]
    path.write_bytes(msgpack.packb(rec, use_bin_type=True))
    with pytest.raises(ValueError, match="input_conv.weight"):
        mpk.load_mpk(path)
