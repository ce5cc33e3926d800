"""ReplayBuffer::save / ReplayBuffer::load (memory.rs:100-115): the reference's `checkpoint/replay_buffer` file.

The file is `bincode::serde::encode_into_std_write(self, .., bincode::config::standard())` of

    struct ReplayBuffer { buffer: HashMap<Fen, MemoryEntry>, order: VecDeque<Fen> }            (memory.rs:26-30)
    struct MemoryEntry  { policy: [f32; 4096] (BigArray), value: f32, visit_count: usize }      (memory.rs:18-24)

bincode 2.0.1 (Cargo.toml:17) with the standard configuration, restated from its published format specification:
little endian; u16..u64/usize as a variable-length integer (one byte below 251, else a marker byte 251/252/253 followed
by u16/u32/u64); f32 as four bytes; a map or sequence as its varint length followed by the items; a string as its
varint byte length followed by UTF-8; a struct, a tuple and a BigArray as their fields with no framing.  shakmaty's `Fen`
(serde feature) is serialised as its `Display` string, here the FEN of `Fen::from_position(.., EnPassantMode::PseudoLegal)`
(memory.rs:42).  No file written by the reference exists in this image, so this framing is "parity unpinned": it is
checked against hand-assembled bytes and by round trips only.
"""
import struct

import numpy as np

from . import ACTION_SPACE, POSITION_DTYPE, position_from_fen

_ROLE_CHARS = "pnbrqk"


def position_to_fen(pos):
    """FEN of an az_position the way shakmaty's Fen prints a standard-chess Setup (castling as KQkq)."""
    roles = [int(x) for x in pos["roles"]]
    white = int(pos["colors"][0])
    rows = []
    for rank in range(7, -1, -1):
        row, empty = "", 0
        for file in range(8):
            bit = 1 << (rank * 8 + file)
            ch = None
            for r in range(6):
                if roles[r] & bit:
                    ch = _ROLE_CHARS[r].upper() if white & bit else _ROLE_CHARS[r]
            if ch is None:
                empty += 1
            else:
                row += (str(empty) if empty else "") + ch
                empty = 0
        rows.append(row + (str(empty) if empty else ""))
    c = int(pos["castling"])
    castling = "".join(ch for k, ch in enumerate("KQkq") if c >> k & 1) or "-"
    ep = int(pos["ep_square"])
    ep_s = "-" if ep < 0 else "abcdefgh"[ep & 7] + str((ep >> 3) + 1)
    return f"{'/'.join(rows)} {'wb'[int(pos['turn'])]} {castling} {ep_s} {int(pos['halfmoves'])} {int(pos['fullmoves'])}"


# ------------------------------------------------------------------------------------------------ bincode primitives
def encode_varint(v):
    if v < 251:
        return bytes([v])
    if v < 1 << 16:
        return b"\xfb" + struct.pack("<H", v)
    if v < 1 << 32:
        return b"\xfc" + struct.pack("<I", v)
    return b"\xfd" + struct.pack("<Q", v)


def decode_varint(buf, off):
    b = buf[off]
    if b < 251:
        return b, off + 1
    if b == 251:
        return struct.unpack_from("<H", buf, off + 1)[0], off + 3
    if b == 252:
        return struct.unpack_from("<I", buf, off + 1)[0], off + 5
    if b == 253:
        return struct.unpack_from("<Q", buf, off + 1)[0], off + 9
    raise ValueError("unsupported varint marker (u128)")


def _encode_str(s):
    raw = s.encode()
    return encode_varint(len(raw)) + raw


def _decode_str(buf, off):
    n, off = decode_varint(buf, off)
    return bytes(buf[off: off + n]).decode(), off + n


# ------------------------------------------------------------------------------------------------------ file format
def write_file(path, pages, n_entries):
    """pages: iterable of (positions, policy [k,4096] f32, value [k] f32, visit_count [k]) in FIFO order (oldest first)."""
    fens = []
    with open(path, "wb") as f:
        f.write(encode_varint(n_entries))                      # HashMap length; any iteration order is a valid file
        for pos, policy, value, visits in pages:
            policy = np.ascontiguousarray(policy, "<f4")
            for k in range(len(pos)):
                fen = position_to_fen(pos[k])
                fens.append(fen)
                f.write(_encode_str(fen))
                f.write(policy[k].tobytes())
                f.write(struct.pack("<f", float(value[k])))
                f.write(encode_varint(int(visits[k])))
        if len(fens) != n_entries:
            raise ValueError("page contents do not add up to n_entries")
        f.write(encode_varint(len(fens)))                      # VecDeque<Fen> order, front = oldest
        for fen in fens:
            f.write(_encode_str(fen))


def read_file(path):
    """-> (positions, policy [n,4096], value [n], visit_count [n]) ordered oldest first (the `order` queue)."""
    buf = memoryview(np.fromfile(path, np.uint8)).cast("B")
    n, off = decode_varint(buf, 0)
    policy = np.empty((n, ACTION_SPACE), np.float32)
    value = np.empty(n, np.float32)
    visits = np.empty(n, np.uint32)
    slot_of = {}
    for i in range(n):
        fen, off = _decode_str(buf, off)
        slot_of[fen] = i
        policy[i] = np.frombuffer(buf, "<f4", ACTION_SPACE, off)
        off += 4 * ACTION_SPACE
        value[i] = struct.unpack_from("<f", buf, off)[0]
        off += 4
        v, off = decode_varint(buf, off)
        visits[i] = v
    m, off = decode_varint(buf, off)
    if m != n:
        raise ValueError(f"order queue holds {m} keys, the map {n}")
    order = np.empty(n, np.int64)
    pos = np.zeros(n, POSITION_DTYPE)
    for i in range(n):
        fen, off = _decode_str(buf, off)
        order[i] = slot_of[fen]
        pos[i] = position_from_fen(fen)
    if off != len(buf):
        raise ValueError("trailing bytes after the replay buffer")
    return pos, policy[order], value[order], visits[order]
