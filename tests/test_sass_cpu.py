"""The built library really contains the Blackwell paths the design claims (B200_PROFILING.md: SASS mnemonics that prove them):
tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG / UBLKCP, tcgen05.commit = UTCBAR -- in the tower, the input convolution AND
the policy/value heads.  Runs on the CPU box (cuobjdump reads the cross-compiled sm_100a code)."""
import os
import re
import shutil
import subprocess

import pytest

import alphazero_chess_b200 as az

LIB = os.path.join(os.path.dirname(az.__file__), "libaz_b200.so")


@pytest.fixture(scope="module")
def sass_by_kernel():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    az.lib()  # builds the library if needed
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
        elif cur is not None:
            kernels[cur].append(line)
    return {k: "\n".join(v) for k, v in kernels.items()}


def _one(sass_by_kernel, needle):
    hits = [k for k in sass_by_kernel if needle in k]
    assert hits, f"no kernel named *{needle}* in {LIB}"
    return hits


@pytest.mark.parametrize("needle", ["conv_tower_kernelILi2E", "conv3x3_tc2_kernelILi1ELi64ELi0E", "k_heads_tc"])
def test_hot_network_kernels_are_tcgen05(sass_by_kernel, needle):
    for k in _one(sass_by_kernel, needle):
        s = sass_by_kernel[k]
        assert re.search(r"\bUTC[A-Z]*MMA", s), f"{k}: no tcgen05.mma"
        assert "LDTM" in s, f"{k}: no tcgen05.ld"
        assert "UTMALDG" in s, f"{k}: no TMA tensor load"
        assert "UTCBAR" in s, f"{k}: no tcgen05.commit"


def test_cta_pair_kernels_use_cta_group_2(sass_by_kernel):
    for k in _one(sass_by_kernel, "conv_tower_kernelILi2E"):
        assert ".2CTA" in sass_by_kernel[k]


def test_heads_stage_fc1_weights_with_bulk_copies(sass_by_kernel):
    for k in _one(sass_by_kernel, "k_heads_tc"):
        assert "UBLKCP" in sass_by_kernel[k]          # cp.async.bulk: fc1 weights into the drained activation ring
        assert "HMMA" in sass_by_kernel[k]            # the value head's [32 x 512] x [512 x 64] stays on mma.sync


def test_search_kernel_uses_no_tensor_cores(sass_by_kernel):
    """k_advance is integer / latency work: there is no GEMM shape in it to reach for (the tier's rule: do not reshape such paths)."""
    for k in _one(sass_by_kernel, "k_advanceILi7E"):
        assert not re.search(r"\bUTC[A-Z]*MMA|\bHMMA", sass_by_kernel[k])
