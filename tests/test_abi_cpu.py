"""The C-ABI shared library loads without a GPU and exports every symbol include/az_b200.h declares."""
import ctypes
import os
import re

import numpy as np

import alphazero_chess_b200 as az

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    with open(os.path.join(ROOT, "include", "az_b200.h")) as f:
        text = f.read()
    names = sorted(set(re.findall(r"^(?:int|void|const char\*|int64_t|uint64_t)\s+(az_[a-z0-9_]+)\(", text, re.M)))
    assert len(names) >= 25
    L = az.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_struct_layouts_and_host_helpers():
    assert ctypes.sizeof(az.Config) == 72
    assert az.POSITION_DTYPE.itemsize == 72 and az.SAMPLE_DTYPE.itemsize == 1120
    c = az.default_config()
    assert (c.num_simulations, c.temperature_annealing, c.num_halfmoves, c.num_fullmoves, c.repetitions, c.seed) == (256, 15, 100, 200, 3, 42)
    assert abs(c.c_puct - 3.0) < 1e-7 and abs(c.dirichlet_alpha - 0.3) < 1e-7 and abs(c.dirichlet_epsilon - 0.25) < 1e-7
    p = az.start_position()
    assert int(p["roles"][0]) == 0x00FF00000000FF00 and p["castling"] == 15 and p["fullmoves"] == 1 and p["ep_square"] == -1
    names, sizes = az.weight_names(), az.weight_sizes()
    assert len(names) == 144 and names[0] == "input_conv.weight" and names[-1] == "value_linear_2.bias"
    assert sum(sizes) == 3_024_777  # parameter count of the 10x128 network incl. BatchNorm statistics
    w = az.random_weights(seed=1)
    assert [a.size for a in w] == sizes


def test_no_gpu_means_an_error_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        return
    try:
        az.Engine(max_games=4)
    except az.EngineError as e:
        assert "az_engine_create" in str(e)
    else:
        raise AssertionError("an engine was created without a GPU")


def test_rust_bindings_list_every_header_function():
    """bindings/rust/az-b200-sys/src/lib.rs (source only, no Rust toolchain here) and the block in INTEGRATION.md declare
    every function of the header except the az_dbg_* test entry points."""
    with open(os.path.join(ROOT, "include", "az_b200.h")) as f:
        names = set(re.findall(r"^(?:int|void|const char\*|int64_t|uint64_t)\s+(az_[a-z0-9_]+)\(", f.read(), re.M))
    names = {n for n in names if not n.startswith("az_dbg_")}
    for rel in (("bindings", "rust", "az-b200-sys", "src", "lib.rs"), ("INTEGRATION.md",)):
        with open(os.path.join(ROOT, *rel)) as f:
            declared = set(re.findall(r"pub fn (az_[a-z0-9_]+)\(", f.read()))
        assert names <= declared, (rel, sorted(names - declared))
        assert declared <= names, (rel, sorted(declared - names))


def test_oracle_weight_catalogue_equals_the_abi():
    """bench.py --impl reference builds its random-init network from the oracle's own catalogue (it must not load the CUDA
    library); the catalogue and the generator have to agree with the product's."""
    from oracle import pyoracle as orc

    import alphazero_chess_b200 as az
    cat = orc.weight_catalogue()
    assert [n for n, _ in cat] == az.weight_names()
    assert [s for _, s in cat] == az.weight_sizes()
    a, b = orc.random_weights(seed=42), az.random_weights(seed=42)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_header_is_plain_c_and_links(tmp_path):
    """include/az_b200.h is a C header (the boundary a cgo / Rust-sys / ctypes binding consumes): a C99 program compiled
    with gcc links against libaz_b200.so, reads the defaults of parameters.rs and the struct sizes the bindings assume."""
    import subprocess

    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include "az_b200.h"
int main(void) {
    az_config c;
    az_position p;
    az_config_default(&c);
    az_position_start(&p);
    printf("%s|%d|%g|%g|%u|%zu|%zu|%zu|%zu|%d|%lld\n", az_version(), c.num_simulations, (double)c.c_puct, (double)c.temperature,
           c.temperature_annealing, sizeof(az_config), sizeof(az_position), sizeof(az_sample), sizeof(az_selfplay_stats),
           (int)p.castling, (long long)az_weight_size(0));
    return 0;
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(az.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libaz_b200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().split("|")
    assert out[1:] == ["256", "3", "1", "15", "72", "72", "1120", str(8 * 14), "15", str(128 * 19 * 9)], out
    assert ctypes.sizeof(az.SelfplayStats) == 8 * 14 and ctypes.sizeof(az.Config) == 72
