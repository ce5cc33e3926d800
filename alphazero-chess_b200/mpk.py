"""Burn `.mpk` checkpoints: `NamedMpkFileRecorder::<FullPrecisionSettings>` as used by
`model.save_file(..)` (training.rs:63,266-270) and `load_model` (main.rs:109-116), read into / written from the 144 f32
arrays `az_load_weights` takes (names and order of `az_weight_name`; burn's own tensor layouts, so no transposition).

Format, restated from burn 0.18's recorder (burn is an un-vendored dependency, Cargo.toml:7; no checkpoint ships with the
reference, so this is "parity unpinned" and covered by structural tests and round trips only):
`rmp_serde::encode::write_named` of `BurnRecord { metadata, item }`; `item` mirrors the module tree of agent.rs:11-66
field by field (a `Vec` is an array); a parameter is `{"id": str, "param": {"bytes": bin, "shape": [..], "dtype": "F32"}}`;
constants (`stride`, `kernel_size`, `dilation`, `groups`, `padding`, `momentum`, `epsilon`) carry no data and are nil.
The reader is tolerant: it follows the field names, accepts `bytes` as bin or as a list of integers, F64/F16/BF16 element
types, and the older `{"value": [...], "shape": [...]}` tensor form.
"""
import numpy as np

NUM_RES_BLOCKS = 10      # parameters.rs:13
NUM_FILTERS = 128        # parameters.rs:14


def _conv(cin, cout, k):
    return [("weight", (cout, cin, k, k)), ("bias", (cout,))]


def _bn(c):
    return [("gamma", (c,)), ("beta", (c,)), ("running_mean", (c,)), ("running_var", (c,))]


def weight_layout():
    """[(dotted name, shape)] in az_weight_name order (agent.rs:50-66 field order)."""
    out = []

    def add(prefix, fields):
        out.extend((f"{prefix}.{n}", s) for n, s in fields)

    add("input_conv", _conv(19, NUM_FILTERS, 3))
    add("input_bn", _bn(NUM_FILTERS))
    for b in range(NUM_RES_BLOCKS):
        for half in ("1", "2"):
            add(f"res_blocks.{b}.conv{half}", _conv(NUM_FILTERS, NUM_FILTERS, 3))
            add(f"res_blocks.{b}.bn{half}", _bn(NUM_FILTERS))
    add("policy_conv_1", _conv(NUM_FILTERS, 32, 1))
    add("policy_bn", _bn(32))
    add("policy_conv_2", _conv(32, 64, 1))
    add("value_conv", _conv(NUM_FILTERS, 8, 1))
    add("value_bn", _bn(8))
    add("value_linear_1", [("weight", (512, 64)), ("bias", (64,))])      # burn Linear: [d_input, d_output]
    add("value_linear_2", [("weight", (64, 1)), ("bias", (1,))])
    return out


def _bf16_to_f32(raw):
    return (np.frombuffer(raw, "<u2").astype(np.uint32) << 16).view(np.float32)


def _tensor_to_f32(node, name):
    """A serialized tensor (TensorData, or ParamSerde wrapping one) -> flat f32 array and its shape."""
    if isinstance(node, dict) and "param" in node:
        node = node["param"]
    if not isinstance(node, dict) or "shape" not in node:
        raise ValueError(f"{name}: not a tensor record")
    shape = tuple(int(x) for x in node["shape"])
    if "bytes" in node:
        raw = node["bytes"]
        raw = bytes(raw) if not isinstance(raw, (bytes, bytearray)) else bytes(raw)
        dtype = node.get("dtype", "F32")
        if isinstance(dtype, dict):                       # externally tagged enum form
            dtype = next(iter(dtype))
        if dtype == "F32":
            data = np.frombuffer(raw, "<f4")
        elif dtype == "F64":
            data = np.frombuffer(raw, "<f8").astype(np.float32)
        elif dtype == "F16":
            data = np.frombuffer(raw, "<f2").astype(np.float32)
        elif dtype == "BF16":
            data = _bf16_to_f32(raw)
        else:
            raise ValueError(f"{name}: unsupported dtype {dtype}")
    elif "value" in node:
        data = np.asarray(node["value"], np.float32)
    else:
        raise ValueError(f"{name}: tensor record without data")
    if data.size != int(np.prod(shape)):
        raise ValueError(f"{name}: {data.size} elements for shape {shape}")
    return np.ascontiguousarray(data, np.float32), shape


def _walk(item, dotted):
    node = item
    for part in dotted.split("."):
        node = node[int(part)] if isinstance(node, (list, tuple)) else node[part]
    return node


def load_mpk(path):
    """-> the 144 flat f32 arrays for az_load_weights / Engine.load_weights."""
    import msgpack
    with open(path, "rb") as f:
        rec = msgpack.unpackb(f.read(), raw=False, strict_map_key=False)
    item = rec["item"] if isinstance(rec, dict) and "item" in rec else rec
    arrays = []
    for name, shape in weight_layout():
        try:
            node = _walk(item, name)
        except (KeyError, IndexError, TypeError) as exc:
            raise ValueError(f"{path}: field {name} is missing from the record") from exc
        data, got = _tensor_to_f32(node, name)
        if got != shape:
            raise ValueError(f"{path}: {name} has shape {got}, expected {shape}")
        arrays.append(data)
    return arrays


def _param(arr, shape, pid):
    return {"id": str(pid), "param": {"bytes": np.ascontiguousarray(arr, "<f4").tobytes(), "shape": list(shape), "dtype": "F32"}}


def save_mpk(path, arrays):
    """Writes the 144 arrays as the record `AlphaZero::save_file` would produce (see the module docstring)."""
    import msgpack
    layout = weight_layout()
    if len(arrays) != len(layout):
        raise ValueError(f"expected {len(layout)} arrays")
    it = iter(range(len(layout)))

    def take(n):
        out = []
        for _ in range(n):
            i = next(it)
            name, shape = layout[i]
            a = np.asarray(arrays[i], np.float32).reshape(-1)
            if a.size != int(np.prod(shape)):
                raise ValueError(f"{name}: {a.size} elements for shape {shape}")
            out.append((name.rsplit(".", 1)[1], _param(a, shape, 0x5EED0000 + i)))
        return out

    def conv():
        d = dict(take(2))
        d.update({"stride": [None, None], "kernel_size": [None, None], "dilation": [None, None], "groups": None, "padding": None})
        return d

    def bn():
        d = dict(take(4))
        d.update({"momentum": None, "epsilon": None})
        return d

    item = {"input_conv": conv(), "input_bn": bn(), "res_blocks": []}
    for _ in range(NUM_RES_BLOCKS):
        item["res_blocks"].append({"conv1": conv(), "bn1": bn(), "conv2": conv(), "bn2": bn()})
    item["policy_conv_1"] = conv()
    item["policy_bn"] = bn()
    item["policy_conv_2"] = conv()
    item["value_conv"] = conv()
    item["value_bn"] = bn()
    item["value_linear_1"] = dict(take(2))
    item["value_linear_2"] = dict(take(2))
    rec = {"metadata": {"float": "f32", "int": "i32",
                        "format": "burn_core::record::file::NamedMpkFileRecorder<burn_core::record::settings::FullPrecisionSettings>",
                        "version": "0.18.0", "settings": "burn_core::record::settings::FullPrecisionSettings"},
           "item": item}
    with open(path, "wb") as f:
        f.write(msgpack.packb(rec, use_bin_type=True))
