// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// Restates ReplayBuffer (memory.rs:27-118): a map FEN -> {policy[4096], value, visit_count} with running means over
// repeated positions (memory.rs:44-58) and a FIFO of the unique keys that evicts the oldest at REPLAY_BUFFER_SIZE
// (memory.rs:60-74).  The FEN key (EnPassantMode::PseudoLegal) is restated as the tuple board / turn / castling /
// pseudo-legal ep / halfmoves / fullmoves.  sample() is random in the reference (choose_multiple on an unseeded rng),
// so only its contract (min(batch, len) distinct entries) is restated by the tests.
#include "oracle.hpp"
#include <cstring>
#include <deque>
#include <string>
#include <unordered_map>

namespace {

struct Entry {
    std::array<float, orc::ACTION_SPACE> policy;
    float value;
    size_t visit_count;
};

struct Replay {
    std::unordered_map<std::string, Entry> buffer;
    std::deque<std::string> order;
    size_t capacity;
};

std::string fen_key(const orc::Pos& p) {
    struct K { orc::u64 b[8]; uint8_t turn, castling; int8_t ep; uint8_t pad; uint16_t hm, fm; } k;
    std::memset(&k, 0, sizeof k);
    std::memcpy(k.b, p.role, 48); std::memcpy(k.b + 6, p.color, 16);
    k.turn = p.turn; k.castling = p.castling; k.ep = (int8_t)orc::pseudo_legal_ep_square(p); k.hm = p.halfmoves; k.fm = p.fullmoves;
    return std::string((const char*)&k, sizeof k);
}

}  // namespace

extern "C" {

void* orc_replay_create(int capacity) { Replay* r = new Replay; r->capacity = (size_t)capacity; return r; }
void orc_replay_destroy(void* h) { delete (Replay*)h; }
int orc_replay_len(void* h) { return (int)((Replay*)h)->buffer.size(); }

// ReplayBuffer::add (memory.rs:41-76); returns 1 for a new unique position, 0 for an update
int orc_replay_add(void* h, const orc::Pos* state, const float* improved_policy, float final_value) {
    Replay* r = (Replay*)h;
    std::string key = fen_key(*state);
    auto it = r->buffer.find(key);
    if (it != r->buffer.end()) {
        Entry& e = it->second;
        float old_count = (float)e.visit_count;
        float new_total_count = old_count + 1.0f;
        e.value = (e.value * old_count + final_value) / new_total_count;
        for (int i = 0; i < orc::ACTION_SPACE; i++) e.policy[i] = (e.policy[i] * old_count + improved_policy[i]) / new_total_count;
        e.visit_count += 1;
        return 0;
    }
    if (r->order.size() >= r->capacity) {
        r->buffer.erase(r->order.front());
        r->order.pop_front();
    }
    Entry e;
    std::memcpy(e.policy.data(), improved_policy, sizeof(float) * orc::ACTION_SPACE);
    e.value = final_value;
    e.visit_count = 1;
    r->buffer.emplace(key, e);
    r->order.push_back(key);
    return 1;
}

// returns visit_count (0 if absent)
int orc_replay_get(void* h, const orc::Pos* state, float* policy_out, float* value_out) {
    Replay* r = (Replay*)h;
    auto it = r->buffer.find(fen_key(*state));
    if (it == r->buffer.end()) return 0;
    std::memcpy(policy_out, it->second.policy.data(), sizeof(float) * orc::ACTION_SPACE);
    *value_out = it->second.value;
    return (int)it->second.visit_count;
}

}  // extern "C"
