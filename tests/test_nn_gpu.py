"""agent.rs forward on the GPU: fp32 path within 1e-5, bf16 tensor-core path within 1e-2 (absolute) of a plain
PyTorch fp32 reference, on random-init weights with randomised BatchNorm statistics."""
import numpy as np
import pytest

import alphazero_chess_b200 as az
from helpers import orc, random_playouts, torch_reference_forward

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def weights():
    return az.random_weights(seed=3, randomize_bn=True)


@pytest.fixture(scope="module")
def positions():
    p, _ = random_playouts(200, seed=8, max_plies=100)
    return p


def test_fp32_path_matches_torch(weights, positions):
    with az.Engine(max_games=256, precision=1) as e:
        e.load_weights(weights)
        planes = e.encode(positions)
        pol, val = e.forward_planes(planes)
        rp, rv = torch_reference_forward(weights, planes)
        assert np.abs(pol - rp).max() <= 1e-5
        assert np.abs(val - rv).max() <= 1e-5
        assert np.allclose(pol.sum(1), 1.0, atol=1e-4)
        pol2, val2 = e.forward(positions)  # to_tensor + forward in one call
        assert np.array_equal(pol, pol2) and np.array_equal(val, val2)


def test_bf16_path_matches_torch(weights, positions):
    with az.Engine(max_games=256, precision=0) as e:
        e.load_weights(weights)
        planes = e.encode(positions)
        pol, val = e.forward(positions)
        rp, rv = torch_reference_forward(weights, planes)
        assert np.abs(pol - rp).max() <= 1e-2
        assert np.abs(val - rv).max() <= 1e-2
        # the tolerance above is loose for a ~1/4096 policy; the relative error is what bf16 really costs
        rel = np.abs(pol - rp).max() / rp.max()
        print(f"\nbf16 path: max |dp| {np.abs(pol - rp).max():.2e} (rel to max p {rel:.2e}), max |dv| {np.abs(val - rv).max():.2e}")
        assert rel < 0.2
        # batch invariance: a row does not depend on what else is in the batch (needed for visit-count parity)
        for i in (0, 1, 77, 199):
            p1, v1 = e.forward(positions[i])
            assert np.array_equal(p1[0], pol[i]) and v1[0] == val[i]
        p3, v3 = e.forward(positions[5:8])
        assert np.array_equal(p3, pol[5:8]) and np.array_equal(v3, val[5:8])
        pp, vv = e.forward_planes(planes[:9])
        assert np.array_equal(pp, pol[:9]) and np.array_equal(vv, val[:9])


def test_no_weights_is_an_error(positions):
    with az.Engine(max_games=16) as e:
        with pytest.raises(az.EngineError):
            e.forward(positions[:2])
        with pytest.raises(az.EngineError):
            e.search(positions[:2], num_simulations=4)


def test_search_with_network_matches_oracle_given_identical_outputs(weights):
    """Visit counts are bit-exact when the oracle's tree is fed the very outputs the GPU network produces."""
    roots, histories = random_playouts(24, seed=31, max_plies=60)
    keep = [i for i in range(len(roots)) if orc.outcome(roots[i]) == 0][:12]
    with az.Engine(max_games=64, num_simulations=48, precision=0) as e:
        e.load_weights(weights)

        def cb(ctx, pos_ptr, pol_ptr, val_ptr):
            pos = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * 72).from_address(pos_ptr)).view(az.POSITION_DTYPE)
            p, v = e.forward(pos)
            np.ctypeslib.as_array(pol_ptr, (4096,))[:] = p[0]
            val_ptr[0] = float(v[0])

        ev = orc.make_evaluator("callback", callback=orc.EVAL_FN(cb))
        prm = orc.make_params(num_simulations=48)
        r = roots[keep]
        visits, scores, depth = e.search(r, num_simulations=48, want_scores=True)
        for k, i in enumerate(keep):
            v, s, d, _ = orc.search(roots[i], prm, ev)
            assert np.array_equal(visits[k], v), k
            assert np.array_equal(scores[k], s), k
            assert depth[k] == d


_MODE_SCRIPT = """
import sys, hashlib
sys.path.insert(0, {tests!r})
import _pkg  # noqa: F401
import alphazero_chess_b200 as az
from helpers import orc, random_playouts
p, _ = random_playouts(150, seed=8, max_plies=100)
live = p[[i for i in range(len(p)) if orc.outcome(p[i]) == 0][:40]]
with az.Engine(max_games=256, precision=0) as e:
    e.load_weights(az.random_weights(seed=3, randomize_bn=True))
    pol, val = e.forward(p)
    visits = e.search(live, num_simulations=24)[0]   # the search kernel writes its own planes (encode_bf16_warp)
print("DIGEST", hashlib.sha256(pol.tobytes() + val.tobytes() + visits.tobytes()).hexdigest())
"""


def test_launch_modes_give_identical_bits():
    """AZ_TOWER_FUSED = 0 (21 launches), 1 (input convolution + fused tower), 2 (everything in one launch) run the same
    MMAs in the same order, so the network's outputs must not differ by a single bit."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    digests = []
    # the last two also force three board ranges inside the tower launch (ranges with and without tiles for a CTA pair)
    # and AZ_INPUT_K32=1 (32-channel plane layout: the two K blocks it drops per tap only ever added exact zeros)
    # and AZ_TOWER_WIDE=1 / 2 (one 16-file / 10-file TMA box per channel half instead of three 8-file boxes: the same MMAs in the same order)
    for mode, split, k32, wide in (("0", None, "0", "0"), ("1", None, "0", "0"), ("2", None, "0", "0"), ("1", "3", "0", "0"), ("2", "3", "0", "0"),
                                   ("1", None, "1", "0"), ("0", None, "1", "0"), ("2", "3", "1", "0"), ("1", None, "1", "1"), ("1", "3", "0", "1"), ("1", None, "1", "2"), ("1", "3", "1", "2"), ("0", None, "1", "2")):
        env = dict(os.environ, AZ_TOWER_FUSED=mode, AZ_INPUT_K32=k32, AZ_TOWER_WIDE=wide)
        if split:
            env["AZ_TOWER_SPLIT"] = split
        out = subprocess.run([sys.executable, "-c", _MODE_SCRIPT.format(tests=here)], env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        digests.append([l for l in out.stdout.splitlines() if l.startswith("DIGEST")][0])
    assert len(set(digests)) == 1, digests


def test_heads_beyond_one_value_batch_per_cta(weights, positions):
    """k_heads_tc batches the value head per 16 tiles (32 boards) of a CTA and stages the fc1 weights into the drained activation
    ring only when a CTA has ONE such batch (<= 4736 boards on 148 SMs).  4801 boards (odd: the last tile holds one board) take the
    other path: several batches per CTA, weights from global memory.  Rows must equal the same boards evaluated alone."""
    reps = (4801 + len(positions) - 1) // len(positions)
    big = np.concatenate([positions] * reps)[:4801]
    with az.Engine(max_games=4801, precision=0) as e:
        e.load_weights(weights)
        pol, val = e.forward(big)
        assert np.allclose(pol.sum(1), 1.0, atol=1e-3)
        for lo in (0, 200, 4600):
            idx = np.arange(lo, min(lo + 200, 4801))
            p1, v1 = e.forward(big[idx])
            assert np.array_equal(pol[idx], p1) and np.array_equal(val[idx], v1), lo
        p_last, v_last = e.forward(big[4800])
        assert np.array_equal(pol[4800], p_last[0]) and val[4800] == v_last[0]


_HEADS_SCRIPT = """
import sys
import numpy as np
sys.path.insert(0, {tests!r})
import _pkg  # noqa: F401
import alphazero_chess_b200 as az
from helpers import random_playouts
p, _ = random_playouts(300, seed=8, max_plies=100)
with az.Engine(max_games=300, precision=0) as e:
    e.load_weights(az.random_weights(seed=3, randomize_bn=True))
    pol, val = e.forward(p)
np.save({out!r}, np.concatenate([pol.ravel(), val.ravel()]))
"""


def test_both_head_kernels_agree(tmp_path):
    """AZ_HEADS_TC=1 (tcgen05 heads, default) and AZ_HEADS_TC=0 (warp-level mma.sync heads) accumulate in different orders, so
    their outputs may differ in the last bits -- but by far less than the 1e-2 the bf16 path is allowed against fp32."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    outs = []
    for mode in ("1", "0"):
        out = str(tmp_path / f"heads{mode}.npy")
        r = subprocess.run([sys.executable, "-c", _HEADS_SCRIPT.format(tests=here, out=out)], env=dict(os.environ, AZ_HEADS_TC=mode),
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(out))
    d = np.abs(outs[0] - outs[1]).max()
    print(f"\ntcgen05 heads vs mma.sync heads: max |d| = {d:.2e}")
    assert d <= 2e-4
