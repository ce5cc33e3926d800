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
