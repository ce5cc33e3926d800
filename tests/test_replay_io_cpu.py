"""memory.rs:100-115 file framing (bincode 2 standard configuration), checked against hand-assembled bytes."""
import struct

import numpy as np
import pytest

import alphazero_chess_b200 as az
from alphazero_chess_b200 import replay_io as rio

START = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1"
AFTER_E4 = "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq - 0 1"     # no black pawn attacks e3: pseudo-legal ep is empty
EP = "rnbqkbnr/ppp1pppp/8/8/3pP3/8/PPPP1PPP/RNBQKBNR b KQkq e3 0 3"


def test_varint_rules():
    cases = {0: b"\x00", 250: b"\xfa", 251: b"\xfb\xfb\x00", 65535: b"\xfb\xff\xff", 65536: b"\xfc\x00\x00\x01\x00",
             2 ** 32: b"\xfd" + struct.pack("<Q", 2 ** 32)}
    for v, raw in cases.items():
        assert rio.encode_varint(v) == raw
        assert rio.decode_varint(raw, 0) == (v, len(raw))


@pytest.mark.parametrize("fen", [START, AFTER_E4, EP, "8/8/4k3/8/8/4K3/8/8 w - - 99 120", "r3k2r/8/8/8/8/8/8/R3K2R b Kq - 3 40"])
def test_fen_round_trip(fen):
    assert rio.position_to_fen(az.position_from_fen(fen)) == fen


def test_file_matches_hand_assembled_bytes(tmp_path):
    pos = np.array([az.position_from_fen(START), az.position_from_fen(EP)], az.POSITION_DTYPE)
    policy = np.zeros((2, 4096), np.float32)
    policy[0, 796] = 1.0
    policy[1, 5] = 0.25
    policy[1, 4095] = 0.75
    value = np.array([0.5, -1.0], np.float32)
    visits = np.array([1, 300], np.uint32)
    path = tmp_path / "replay_buffer"
    rio.write_file(path, [(pos, policy, value, visits)], 2)
    want = b"\x02"                                                    # HashMap len
    for k, fen in enumerate((START, EP)):
        want += bytes([len(fen)]) + fen.encode()                      # Fen as a string
        want += policy[k].astype("<f4").tobytes()                     # BigArray: 4096 raw f32, no length
        want += struct.pack("<f", value[k])
        want += b"\x01" if k == 0 else b"\xfb" + struct.pack("<H", 300)   # usize varint
    want += b"\x02" + bytes([len(START)]) + START.encode() + bytes([len(EP)]) + EP.encode()   # VecDeque order
    assert path.read_bytes() == want
    p2, pol2, val2, vis2 = rio.read_file(path)
    assert p2.tobytes() == pos.tobytes() and np.array_equal(pol2, policy) and np.array_equal(val2, value) and np.array_equal(vis2, visits)


def test_map_order_is_free_but_queue_order_rules(tmp_path):
    """A HashMap iterates in arbitrary order; FIFO order comes from the `order` queue alone."""
    row = np.zeros(4096, "<f4")
    body = b"\x02"
    for fen, val in ((EP, 2.0), (START, 1.0)):                        # map lists EP first
        body += bytes([len(fen)]) + fen.encode() + row.tobytes() + struct.pack("<f", val) + b"\x07"
    body += b"\x02" + bytes([len(START)]) + START.encode() + bytes([len(EP)]) + EP.encode()
    path = tmp_path / "rb"
    path.write_bytes(body)
    pos, _, val, vis = rio.read_file(path)
    assert rio.position_to_fen(pos[0]) == START and list(val) == [1.0, 2.0] and list(vis) == [7, 7]
