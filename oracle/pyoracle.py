"""ctypes access to the CPU oracle (TEST INFRASTRUCTURE ONLY: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline and
--impl reference legs).  The product package never imports this module."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_az.so")

POSITION_DTYPE = np.dtype(
    [("roles", "<u8", 6), ("colors", "<u8", 2), ("turn", "u1"), ("castling", "u1"), ("ep_square", "i1"), ("reserved", "u1"),
     ("halfmoves", "<u2"), ("fullmoves", "<u2")]
)
ACTION_SPACE = 4096


class SearchParams(ctypes.Structure):
    _fields_ = [("num_simulations", ctypes.c_int32), ("c_puct", ctypes.c_float), ("dirichlet_alpha", ctypes.c_float),
                ("dirichlet_eps", ctypes.c_float), ("temperature_annealing", ctypes.c_uint32), ("seed", ctypes.c_uint64),
                ("temperature", ctypes.c_float)]


EVAL_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float))


class Evaluator(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("stub_seed", ctypes.c_uint64), ("net", ctypes.c_void_p), ("callback", EVAL_FN),
                ("callback_ctx", ctypes.c_void_p)]


class EpisodeStats(ctypes.Structure):
    _fields_ = [("n_steps", ctypes.c_int32), ("result", ctypes.c_int32), ("simulations", ctypes.c_int64), ("evals", ctypes.c_int64),
                ("cache_hits", ctypes.c_int64), ("seconds", ctypes.c_double)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = ctypes.CDLL(LIB)
        L.orc_perft.restype = ctypes.c_uint64
        L.orc_rng_u64.restype = ctypes.c_uint64
        L.orc_rng_u64.argtypes = [ctypes.c_uint64] * 5
        L.orc_det_log.restype = ctypes.c_double
        L.orc_det_log.argtypes = [ctypes.c_double]
        L.orc_det_exp.restype = ctypes.c_double
        L.orc_det_exp.argtypes = [ctypes.c_double]
        L.orc_net_create.restype = ctypes.c_void_p
        L.orc_cache_create.restype = ctypes.c_void_p
        L.orc_replay_create.restype = ctypes.c_void_p
        L.orc_cache_size.restype = ctypes.c_uint64
        L.orc_net_array_size.restype = ctypes.c_uint64
        L.orc_init()
        _lib = L
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else ctypes.c_void_p(0)


def from_fen(fen):
    p = np.zeros(1, POSITION_DTYPE)
    assert lib().orc_pos_from_fen(fen.encode(), _p(p)) == 0, fen
    return p[0]


def startpos():
    p = np.zeros(1, POSITION_DTYPE)
    lib().orc_startpos(_p(p))
    return p[0]


def _one(pos):
    return np.ascontiguousarray(np.atleast_1d(pos), POSITION_DTYPE)


def legal_moves(pos):
    p = _one(pos)
    mv = np.zeros(256, np.uint16)
    ix = np.zeros(256, np.uint16)
    n = lib().orc_legal_moves(_p(p), _p(mv), _p(ix))
    return mv[:n].copy(), ix[:n].copy()


def perft(pos, depth):
    return int(lib().orc_perft(_p(_one(pos)), int(depth)))


def minimax_scores(pos, depth):
    """chess.rs:295-318: -negamax(child, depth - 1) for every legal move, in legal-move order."""
    out = np.zeros(256, np.int32)
    n = lib().orc_minimax_scores(_p(_one(pos)), int(depth), _p(out))
    return out[:n].copy()


def perft_batch(positions, depth, threads):
    p = np.ascontiguousarray(positions, POSITION_DTYPE)
    out = np.zeros(p.shape[0], np.uint64)
    lib().orc_perft_batch(_p(p), p.shape[0], int(depth), int(threads), _p(out))
    return out


def outcome(pos):
    return lib().orc_outcome(_p(_one(pos)))


def play_encoded(pos, mv):
    p = _one(pos).copy()
    assert lib().orc_play_encoded(_p(p), int(mv)) == 0
    return p[0]


def play_move(pos, action_index, history=None):
    """chess.rs play_move via a policy index; returns (new position, GameResult)."""
    p = _one(pos).copy()
    h = np.ascontiguousarray(history, POSITION_DTYPE) if history is not None and len(history) else None
    r = lib().orc_play_move(_p(p), _p(h), 0 if h is None else h.shape[0], int(action_index))
    return p[0], r


def move_to_index(pos, mv):
    return lib().orc_move_to_index(_p(_one(pos)), int(mv))


def index_to_move(pos, index):
    out = ctypes.c_uint16(0)
    ok = lib().orc_index_to_move(_p(_one(pos)), int(index), ctypes.byref(out))
    return out.value if ok else None


def to_tensor(pos):
    out = np.zeros(19 * 64, np.float32)
    lib().orc_to_tensor(_p(_one(pos)), _p(out))
    return out.reshape(19, 8, 8)


def playout_corpus(n, seed=42, max_plies=80, roots=None, with_history=True):
    """Seeded random-playout positions generated inside the oracle library: (positions [n],
    history [total] or None, hist_offsets [n + 1]); each game's history ends in its position."""
    r = np.ascontiguousarray(roots if roots is not None else [startpos()], POSITION_DTYPE)
    pos = np.zeros(n, POSITION_DTYPE)
    offs = np.zeros(n + 1, np.uint32)
    lib().orc_playout_corpus.restype = ctypes.c_int64
    hist = None
    if with_history:
        cap = int(n) * (max_plies + 1)
        hist = np.zeros(cap, POSITION_DTYPE)
        total = lib().orc_playout_corpus(ctypes.c_uint64(seed), int(n), int(max_plies), _p(r), int(r.shape[0]), _p(pos), _p(hist),
                                         ctypes.c_int64(cap), _p(offs))
        hist = hist[:total]
    else:
        lib().orc_playout_corpus(ctypes.c_uint64(seed), int(n), int(max_plies), _p(r), int(r.shape[0]), _p(pos), ctypes.c_void_p(0),
                                 ctypes.c_int64(0), _p(offs))
    return pos, hist, offs


def legal_moves_batch(positions):
    p = np.ascontiguousarray(positions, POSITION_DTYPE)
    n = p.shape[0]
    moves = np.zeros((n, 256), np.uint16)
    index = np.zeros((n, 256), np.uint16)
    count = np.zeros(n, np.int32)
    lib().orc_legal_moves_batch(_p(p), n, _p(moves), _p(index), _p(count))
    return moves, index, count


def to_tensor_batch(positions):
    p = np.ascontiguousarray(positions, POSITION_DTYPE)
    out = np.zeros((p.shape[0], 19, 8, 8), np.float32)
    lib().orc_to_tensor_batch(_p(p), p.shape[0], _p(out))
    return out


def play_move_batch(positions, action_index, history=None, hist_offsets=None):
    p = np.ascontiguousarray(positions, POSITION_DTYPE).copy()
    a = np.ascontiguousarray(action_index, np.uint16)
    res = np.zeros(p.shape[0], np.int32)
    h = np.ascontiguousarray(history, POSITION_DTYPE) if history is not None else None
    ho = np.ascontiguousarray(hist_offsets, np.uint32) if hist_offsets is not None else None
    lib().orc_play_move_batch(_p(p), p.shape[0], _p(h), _p(ho), _p(a), _p(res))
    return p, res


def index_to_move_batch(positions, index):
    p = np.ascontiguousarray(positions, POSITION_DTYPE)
    ix = np.ascontiguousarray(index, np.uint16)
    out = np.zeros(p.shape[0], np.uint16)
    lib().orc_index_to_move_batch(_p(p), p.shape[0], _p(ix), _p(out))
    return out


def stub_eval(seed, pos):
    pol = np.zeros(ACTION_SPACE, np.float32)
    val = ctypes.c_float(0)
    lib().orc_stub_eval(ctypes.c_uint64(seed), _p(_one(pos)), _p(pol), ctypes.byref(val))
    return pol, val.value


def dirichlet(seed, game, ply, alpha, n):
    out = np.zeros(n, np.float32)
    lib().orc_dirichlet(ctypes.c_uint64(seed), ctypes.c_uint64(game), ctypes.c_uint64(ply), ctypes.c_float(alpha), int(n), _p(out))
    return out


class Net:
    """fp32 CPU network (oracle/net.cpp) built from the same 144 arrays the engine loads."""

    def __init__(self, arrays):
        self._arrs = [np.ascontiguousarray(a, np.float32).ravel() for a in arrays]
        ptrs = (ctypes.c_void_p * len(self._arrs))(*[a.ctypes.data for a in self._arrs])
        self.h = lib().orc_net_create(ptrs, len(self._arrs))
        assert self.h, "orc_net_create failed"

    def forward_positions(self, positions):
        p = np.ascontiguousarray(np.atleast_1d(positions), POSITION_DTYPE)
        pol = np.zeros((p.shape[0], ACTION_SPACE), np.float32)
        val = np.zeros(p.shape[0], np.float32)
        lib().orc_net_forward_pos(ctypes.c_void_p(self.h), _p(p), p.shape[0], _p(pol), _p(val))
        return pol, val

    def forward_planes(self, planes):
        x = np.ascontiguousarray(planes, np.float32).reshape(-1, 19 * 64)
        pol = np.zeros((x.shape[0], ACTION_SPACE), np.float32)
        val = np.zeros(x.shape[0], np.float32)
        lib().orc_net_forward_planes(ctypes.c_void_p(self.h), _p(x), x.shape[0], _p(pol), _p(val))
        return pol, val

    def __del__(self):
        try:
            lib().orc_net_destroy(ctypes.c_void_p(self.h))
        except Exception:
            pass


def make_evaluator(kind="stub", stub_seed=0, net=None, callback=None):
    ev = Evaluator()
    if kind == "stub":
        ev.kind, ev.stub_seed = 0, stub_seed
    elif kind == "net":
        ev.kind, ev.net = 1, net.h
    else:
        ev.kind = 2
        ev.callback = callback
    return ev


def make_params(num_simulations=256, c_puct=3.0, alpha=0.3, eps=0.25, anneal=15, seed=42, temperature=1.0):
    return SearchParams(num_simulations, c_puct, alpha, eps, anneal, seed, temperature)


def improved_policy(visits, temperature=1.0):
    """tree.rs:173-177: visits^(1/T) / sum."""
    v = np.ascontiguousarray(visits, np.float32)
    out = np.zeros(ACTION_SPACE, np.float32)
    lib().orc_improved_policy(_p(v), ctypes.c_float(temperature), _p(out))
    return out


def search(root, params, evaluator, history=None, noise_game=-1, noise_ply=0):
    """MCTree::init + monte_carlo_tree_search.  Returns (visits [4096], scores [4096], depth, evals)."""
    r = _one(root)
    h = np.ascontiguousarray(history, POSITION_DTYPE) if history is not None and len(history) else None
    visits = np.zeros(ACTION_SPACE, np.float32)
    scores = np.zeros(ACTION_SPACE, np.float32)
    depth = ctypes.c_int(0)
    evals = ctypes.c_long(0)
    rc = lib().orc_search(_p(r), _p(h), 0 if h is None else h.shape[0], ctypes.byref(params), ctypes.byref(evaluator),
                          ctypes.c_int64(noise_game), ctypes.c_int64(noise_ply), _p(visits), _p(scores), ctypes.byref(depth),
                          ctypes.byref(evals))
    assert rc == 0
    return visits, scores, depth.value, evals.value


def selfplay_episode(params, evaluator, game_id, max_steps=512, cache=None, want_visits=True):
    """run_episode (training.rs:294-338).  Returns dict(positions, visits, final_value, depth, action, stats)."""
    pos = np.zeros(max_steps, POSITION_DTYPE)
    visits = np.zeros((max_steps, ACTION_SPACE), np.float32) if want_visits else None
    fv = np.zeros(max_steps, np.float32)
    depth = np.zeros(max_steps, np.int32)
    action = np.zeros(max_steps, np.int32)
    st = EpisodeStats()
    rc = lib().orc_selfplay_episode(ctypes.byref(params), ctypes.byref(evaluator), ctypes.c_uint64(game_id),
                                    ctypes.c_void_p(cache) if cache else ctypes.c_void_p(0), int(max_steps), _p(pos), _p(visits), _p(fv),
                                    _p(depth), _p(action), ctypes.byref(st))
    assert rc == 0
    n = st.n_steps
    return dict(positions=pos[:n], visits=None if visits is None else visits[:n], final_value=fv[:n], depth=depth[:n],
                action=action[:n], stats=st)


def cache_create():
    return lib().orc_cache_create()


def cache_destroy(c):
    lib().orc_cache_destroy(ctypes.c_void_p(c))


class Replay:
    """ReplayBuffer of memory.rs restated (oracle/replay.cpp)."""

    def __init__(self, capacity):
        self.h = lib().orc_replay_create(int(capacity))

    def add(self, state, improved_policy, final_value):
        pol = np.ascontiguousarray(improved_policy, np.float32)
        return lib().orc_replay_add(ctypes.c_void_p(self.h), _p(_one(state)), _p(pol), ctypes.c_float(final_value))

    def get(self, state):
        pol = np.zeros(ACTION_SPACE, np.float32)
        val = ctypes.c_float(0)
        n = lib().orc_replay_get(ctypes.c_void_p(self.h), _p(_one(state)), _p(pol), ctypes.byref(val))
        return pol, val.value, n

    def __len__(self):
        return lib().orc_replay_len(ctypes.c_void_p(self.h))

    def __del__(self):
        try:
            lib().orc_replay_destroy(ctypes.c_void_p(self.h))
        except Exception:
            pass


# ---- random-init weights without the product library (bench.py --impl reference must not load libaz_b200.so) ---------------
def weight_catalogue():
    """(name, size) of the 144 arrays of AlphaZero::new (agent.rs:69-110) in record order, restated independently of
    include/az_b200.h's az_weight_name / az_weight_size (tests/test_abi_cpu.py checks that the two agree)."""
    bn = ["gamma", "beta", "running_mean", "running_var"]
    out = [("input_conv.weight", 128 * 19 * 9), ("input_conv.bias", 128)] + [(f"input_bn.{k}", 128) for k in bn]
    for b in range(10):
        for c, n in (("conv1", "bn1"), ("conv2", "bn2")):
            out += [(f"res_blocks.{b}.{c}.weight", 128 * 128 * 9), (f"res_blocks.{b}.{c}.bias", 128)]
            out += [(f"res_blocks.{b}.{n}.{k}", 128) for k in bn]
    out += [("policy_conv_1.weight", 32 * 128), ("policy_conv_1.bias", 32)] + [(f"policy_bn.{k}", 32) for k in bn]
    out += [("policy_conv_2.weight", 64 * 32), ("policy_conv_2.bias", 64)]
    out += [("value_conv.weight", 8 * 128), ("value_conv.bias", 8)] + [(f"value_bn.{k}", 8) for k in bn]
    out += [("value_linear_1.weight", 512 * 64), ("value_linear_1.bias", 64), ("value_linear_2.weight", 64), ("value_linear_2.bias", 1)]
    return out


_FAN_IN = {"input_conv": 19 * 9, "conv1": 128 * 9, "conv2": 128 * 9, "policy_conv_1": 128, "policy_conv_2": 32, "value_conv": 128,
           "value_linear_1": 512, "value_linear_2": 64}


def random_weights(seed=42):
    """The same random-init network as the product's random_weights(seed) (conv / linear U(-k, k), k = 1/sqrt(fan_in);
    BatchNorm gamma 1, beta 0, mean 0, var 1), generated without touching the CUDA library."""
    rng = np.random.default_rng(seed)
    out = []
    for name, size in weight_catalogue():
        layer, field = name.split(".")[-2:]
        if field in ("weight", "bias"):
            k = 1.0 / np.sqrt(_FAN_IN[layer])
            out.append(rng.uniform(-k, k, size).astype(np.float32))
        elif field in ("gamma", "running_var"):
            out.append(np.ones(size, np.float32))
        else:
            out.append(np.zeros(size, np.float32))
    return out
