// Device-resident replay buffer (memory.rs:27-118): positions de-duplicated by their FEN identity (pseudo-legal ep,
// counters included), running means of policy and value over repeated visits, FIFO eviction of the oldest UNIQUE
// position at capacity, uniform sampling without replacement.  It consumes the az_sample records self-play leaves in
// device memory, so a generation's positions never travel to the host.
//
// `add` is order dependent in the reference: the running mean (old*n + new)/(n+1) is evaluated in f32 in the order the
// steps arrive, and which position is evicted depends on the insertion order of the unique ones.  To stay bit-exact AND
// parallel the batch is applied in two phases per chunk:
//   A  k_replay_resolve (one warp): decides for every step, in order, which slot it lands in, whether it is a new unique
//      position (evicting the oldest when full) and the visit count it meets.  Only 64-byte keys and counters move; the
//      warp probes the hash table for 32 steps at once and then replays their decisions in order from shared memory
//      (a speculative hit is void if an earlier step of the window evicted that slot, a miss if an earlier step of the
//      window brought the same position), so the DRAM/L2 latency of a probe is paid once per 32 steps.
//   B  k_replay_apply (one CTA per touched slot): the steps that hit the same slot form a chain (prev/next links written
//      by A); the CTA of the chain's last step walks it in order with the 16 KB policy row in registers.  Chains are
//      independent, and a chain whose slot was evicted again later in the chunk is never touched.
// Evicted positions leave stale entries in the open-addressing table (their slot now holds another key, so a probe just
// walks past them); k_replay_rebuild re-inserts the live slots when too many have accumulated.
// Sampling and batch assembly (planes, policy rows, values) are parallel over the batch.
#include "engine.h"
#include "mcts.h"
#include "det_math.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace azb {

struct ReplayPtrs {
    DPos* keys;            // [cap] slot -> key (FEN identity)
    float* policy;         // [cap][4096] running mean of improved_policy
    float* value;          // [cap]
    uint32_t* visits;      // [cap] visit_count
    int32_t* table;        // [tsize] open addressing, linear probing: slot or -1 (entries of evicted positions go stale)
    int32_t* state;        // [4]: head (next slot, = oldest when full), len, new_unique (of the last add), stale entries
    int32_t* last_touch;   // [cap] index (within the current chunk) of the last step that touched the slot, -1 = none
    int cap, tmask;
};

struct StepPtrs {          // per step of the current chunk, written by phase A, read by phase B
    int32_t* slot;
    uint32_t* old;         // visit count the step meets (0 = new unique position)
    int32_t* prev;         // previous / next step of the chunk on the same slot (-1 = none)
    int32_t* next;
};

}  // namespace azb

struct az_replay {
    az_engine* eng = nullptr;
    azb::ReplayPtrs p{};
    az_sample* d_staging = nullptr;   // host-sample staging for az_replay_add
    size_t staging_cap = 0;
    int32_t* d_idx = nullptr;         // sample indices
    float* d_planes = nullptr;        // [max_batch][19][64]
    float* d_pol = nullptr;           // [max_batch][4096]
    float* d_val = nullptr;           // [max_batch]
    az_position* d_pos = nullptr;     // [max_batch] export/import page (lazy)
    int max_batch = 0;
    azb::StepPtrs sp{};
    int chunk = 0;                    // steps per phase A/B round
    int window = 0;                   // steps probed together by phase A
    int force_slow = 0;               // AZ_REPLAY_FORCE_SLOW=1: phase A always takes its in-order path (tests)
};

namespace azb {

__device__ __forceinline__ int replay_find(const ReplayPtrs& r, const DPos& key, u64 h) {
    for (uint32_t i = (uint32_t)h & r.tmask;; i = (i + 1) & r.tmask) {
        const int s = r.table[i];
        if (s < 0) return -1;
        if (fen_key_equal(r.keys[s], key)) return s;
    }
}
__device__ __forceinline__ void replay_table_insert(const ReplayPtrs& r, u64 h, int slot) {
    uint32_t i = (uint32_t)h & r.tmask;
    while (r.table[i] >= 0) i = (i + 1) & r.tmask;
    r.table[i] = slot;
}

// Phase A.  One warp; lanes < W each take one step of the window.
__global__ void __launch_bounds__(32) k_replay_resolve(ReplayPtrs r, StepPtrs sp, const az_sample* __restrict__ samples, int n, int W,
                                                       int force_slow) {
    __shared__ DPos s_key[32];
    __shared__ int s_slot[32], s_tpos[32];
    __shared__ uint32_t s_cnt[32];
    const int lane = threadIdx.x;
    int head = r.state[0], len = r.state[1], new_unique = 0, stale = r.state[3];
    for (int i0 = 0; i0 < n; i0 += W) {
        const int m = min(W, n - i0);
        const bool active = lane < m;
        const int gi = i0 + lane;
        DPos key{};
        int s = -1, lt = -1;
        uint32_t pos = 0, v_old = 0;
        if (active) {   // speculative probe against the table as it stands before this window
            key = fen_key_of(dpos_from_wire(samples[gi].position));
            pos = (uint32_t)fen_key_hash(key) & r.tmask;
            for (int guard = 0;; pos = (pos + 1) & r.tmask) {
                s = r.table[pos];
                if (s < 0 || fen_key_equal(r.keys[s], key)) break;
                if (++guard > r.tmask) __trap();   // the table always keeps empty entries (k_replay_rebuild)
            }
            if (s >= 0) { v_old = r.visits[s]; lt = r.last_touch[s]; }
            s_key[lane] = key;
        }
        __syncwarp();
        // pairwise relations inside the window, by shuffling the 64-bit hashes (full keys are compared only on a hash match)
        const u64 h = active ? fen_key_hash(key) : 0ULL;
        const bool spec_hit = s >= 0;
        int dup = -1, root = -1, rank = 0;   // latest / first earlier step with the same position, how many there are
        bool last = true, tconf = false;
        for (int l2 = 0; l2 < m; l2++) {
            const u64 h2 = __shfl_sync(0xffffffffu, h, l2);
            const uint32_t pos2 = __shfl_sync(0xffffffffu, pos, l2);
            const bool hit2 = __shfl_sync(0xffffffffu, (int)spec_hit, l2) != 0;
            if (!active || l2 == lane) continue;
            const bool same = h2 == h && fen_key_equal(s_key[l2], key);
            if (l2 < lane) {
                if (same) { if (root < 0) root = l2; dup = l2; rank++; }
                else if (!spec_hit && !hit2 && pos2 == pos) tconf = true;   // two new positions want the same empty table entry
            } else if (same) {
                last = false;
            }
        }
        int my_slot = -1, my_prev = -1, my_tpos = -1;
        uint32_t my_old = 0;
        bool my_ins = false;
        const int head0 = head;
        // fast path: assume no speculative hit was evicted inside this window; then everything is a prefix sum
        const bool opt_ins = active && dup < 0 && !spec_hit;
        const unsigned ins_mask = __ballot_sync(0xffffffffu, opt_ins);
        const int ins_before = __popc(ins_mask & ((1u << lane) - 1u));
        int d = s - head0;
        if (d < 0) d += r.cap;
        const bool violated = force_slow || (active && ((dup < 0 && spec_hit && d < ins_before) || tconf));
        if (!__any_sync(0xffffffffu, violated)) {
            if (active && dup < 0) {
                my_ins = opt_ins;
                my_slot = opt_ins ? (head0 + ins_before) % r.cap : s;
                my_old = opt_ins ? 0u : v_old;
                my_prev = opt_ins ? -1 : lt;
                my_tpos = opt_ins ? (int)pos : -1;
            }
            const int src = root >= 0 ? root : 0;
            const int root_slot = __shfl_sync(0xffffffffu, my_slot, src);
            const uint32_t root_old = __shfl_sync(0xffffffffu, my_old, src);
            if (active && dup >= 0) { my_slot = root_slot; my_old = root_old + (uint32_t)rank; my_prev = i0 + dup; }
            const int total = __popc(ins_mask), grow = min(total, r.cap - len);
            len += grow; stale += total - grow;   // order.pop_front() for every insertion beyond capacity
            head = (head0 + total) % r.cap;
            new_unique += total;
        } else {
            // slow path (rare): the decisions strictly in order
            int ins = 0;
            for (int j = 0; j < m; j++) {
                int ins_j = 0;
                if (lane == j) {
                    bool hit = false;
                    if (dup >= 0) { my_slot = s_slot[dup]; my_old = s_cnt[dup]; my_prev = i0 + dup; hit = true; }
                    else if (spec_hit && d >= ins) { my_slot = s; my_old = v_old; my_prev = lt; hit = true; }   // else: evicted earlier in this window
                    if (!hit) {   // ReplayBuffer::add's else branch (memory.rs:60-74)
                        my_ins = true; ins_j = 1; my_slot = head; my_old = 0; my_prev = -1;
                        uint32_t tp = pos;
                        bool check_global = spec_hit;    // the probe stopped on a (now void) hit, not on an empty entry
                        if (check_global) tp = (tp + 1) & r.tmask;
                        for (;;) {
                            bool taken = false;
                            for (int l2 = 0; l2 < j; l2++) taken |= s_tpos[l2] == (int)tp;
                            if (!taken && (!check_global || r.table[tp] < 0)) break;
                            tp = (tp + 1) & r.tmask;
                            check_global = true;
                        }
                        my_tpos = (int)tp;
                    }
                    s_slot[j] = my_slot; s_cnt[j] = my_old + 1; s_tpos[j] = my_tpos;
                }
                ins_j = __shfl_sync(0xffffffffu, ins_j, j);
                if (ins_j) {
                    if (len >= r.cap) stale++; else len++;   // order.pop_front(): the oldest unique position goes
                    head = head + 1 == r.cap ? 0 : head + 1;
                    ins++; new_unique++;
                }
                __syncwarp();
            }
        }
        if (active) {
            sp.slot[gi] = my_slot; sp.old[gi] = my_old; sp.prev[gi] = my_prev;
            if (my_prev >= 0) sp.next[my_prev] = gi;
            if (my_ins) { r.keys[my_slot] = key; r.table[my_tpos] = my_slot; }
            if (last) { r.visits[my_slot] = my_old + 1; r.last_touch[my_slot] = gi; }   // the window's last step on this position
        }
        __syncwarp();
    }
    if (lane == 0) { r.state[0] = head; r.state[1] = len; r.state[2] += new_unique; r.state[3] = stale; }
}

// Phase B.  Block i acts only if step i is the last one of the chunk on its slot; it then owns that slot.
__global__ void __launch_bounds__(256) k_replay_apply(ReplayPtrs r, StepPtrs sp, const az_sample* __restrict__ samples, int n, float inv_t) {
    __shared__ float s_new[AZ_ACTION_SPACE];
    __shared__ float s_w[AZ_MAX_MOVES];
    __shared__ float s_sum;
    __shared__ int s_cur;
    const int tail = blockIdx.x, t = threadIdx.x;
    const int slot = sp.slot[tail];
    if (r.last_touch[slot] != tail) return;
    if (t == 0) {
        int h = tail, guard = 0;
        while (sp.old[h] != 0 && sp.prev[h] >= 0) {
            h = sp.prev[h];
            if (++guard > n) __trap();   // a chain cannot be longer than the chunk: broken links must not hang the device
        }
        s_cur = h;
    }
    __syncthreads();
    int cur = s_cur;
    float row[16], val = 0.0f;
    int steps = 0;
    float* pol = r.policy + (size_t)slot * AZ_ACTION_SPACE;
    if (sp.old[cur] != 0) {   // the chain continues an entry stored before this chunk
#pragma unroll
        for (int q = 0; q < 16; q++) row[q] = pol[t + 256 * q];
        val = r.value[slot];
    }
    for (;;) {
        const az_sample* sm = samples + cur;
        for (int k = t; k < AZ_ACTION_SPACE; k += 256) s_new[k] = 0.0f;
        __syncthreads();
        // improved_policy (tree.rs:173-177): weights = visits^(1/T), normalised by their sum taken in index order (the pairs
        // of a sample are sorted by policy index).  At T = 1 the sum is the sample's own simulation count, whatever engine,
        // configuration or az_search call produced it.
        const int nv = min((int)sm->n_visits, AZ_MAX_MOVES);
        for (int k = t; k < nv; k += 256) s_w[k] = pow_inv_temperature((float)sm->count[k], inv_t);
        __syncthreads();
        if (t == 0) {
            float acc = 0.0f;
            for (int k = 0; k < nv; k++) acc = __fadd_rn(acc, s_w[k]);
            s_sum = acc;
        }
        __syncthreads();
        const float wsum = s_sum;
        for (int k = t; k < nv; k += 256) s_new[sm->index[k]] = __fdiv_rn(s_w[k], wsum);
        __syncthreads();
        const uint32_t old = sp.old[cur];
        if (old == 0) {
#pragma unroll
            for (int q = 0; q < 16; q++) row[q] = s_new[t + 256 * q];
            val = sm->final_value;
        } else {   // memory.rs:45-54
            const float old_count = (float)old, total = __fadd_rn(old_count, 1.0f);
#pragma unroll
            for (int q = 0; q < 16; q++) row[q] = __fdiv_rn(__fadd_rn(__fmul_rn(row[q], old_count), s_new[t + 256 * q]), total);
            val = __fdiv_rn(__fadd_rn(__fmul_rn(val, old_count), sm->final_value), total);
        }
        __syncthreads();
        if (cur == tail) break;
        cur = sp.next[cur];
        if (cur < 0 || cur >= n || ++steps > n) __trap();
    }
#pragma unroll
    for (int q = 0; q < 16; q++) pol[t + 256 * q] = row[q];
    if (t == 0) r.value[slot] = val;
}

// Table maintenance: phase 0 clears, phase 1 re-inserts the live slots, phase 2 resets the stale counter -- each only if
// more than cap/2 stale entries have accumulated (the table has at least 4 * cap entries).
__global__ void __launch_bounds__(256) k_replay_rebuild(ReplayPtrs r, int phase) {
    if (r.state[3] <= r.cap / 2) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (phase == 0) {
        if (i <= r.tmask) r.table[i] = -1;
    } else if (phase == 1) {
        if (i < r.state[1]) {   // live slots are [0, len): the ring is dense
            uint32_t pos = (uint32_t)fen_key_hash(r.keys[i]) & r.tmask;
            while (atomicCAS(&r.table[pos], -1, i) != -1) pos = (pos + 1) & r.tmask;
        }
    } else if (i == 0) {
        r.state[3] = 0;
    }
}

// training batch assembly: planes (to_tensor of the stored position), policy rows, values for the chosen slots
__global__ void __launch_bounds__(256) k_replay_gather(ReplayPtrs r, const int32_t* __restrict__ idx, int n, float* __restrict__ planes,
                                                       float* __restrict__ policy, float* __restrict__ value) {
    const int b = blockIdx.x;
    if (b >= n) return;
    // idx holds FIFO ranks (0 = oldest), so a buffer reloaded from a file samples like the one that was saved
    const int len = r.state[1];
    const int slot = ((len >= r.cap ? r.state[0] : 0) + idx[b]) % r.cap;
    const DPos p = r.keys[slot];   // the key IS the position as the reference rebuilds it from the FEN (memory.rs:90)
    const u64 occ = occupied(p);
    const u64 ours = meta_turn(p.meta) == 0 ? p.white : occ ^ p.white;
    const int pep = meta_ep(p.meta);  // already the pseudo-legal ep square
    for (int e = threadIdx.x; e < AZ_NUM_PLANES * 64; e += 256)
        planes[(size_t)b * AZ_NUM_PLANES * 64 + e] = plane_value(p, e >> 6, e & 63, pep, ours, occ ^ ours);
    const float4* src = reinterpret_cast<const float4*>(r.policy + (size_t)slot * AZ_ACTION_SPACE);
    float4* dst = reinterpret_cast<float4*>(policy + (size_t)b * AZ_ACTION_SPACE);
    for (int k = threadIdx.x; k < AZ_ACTION_SPACE / 4; k += 256) dst[k] = src[k];
    if (threadIdx.x == 0) value[b] = r.value[slot];
}

// test hook: entry of one position
__global__ void k_replay_get(ReplayPtrs r, const az_position* __restrict__ wire, float* __restrict__ policy, float* __restrict__ value,
                             uint32_t* __restrict__ visits) {
    __shared__ int s_slot;
    if (threadIdx.x == 0) {
        const DPos key = fen_key_of(dpos_from_wire(wire[0]));
        s_slot = replay_find(r, key, fen_key_hash(key));
        visits[0] = s_slot >= 0 ? r.visits[s_slot] : 0u;
        value[0] = s_slot >= 0 ? r.value[s_slot] : 0.0f;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < AZ_ACTION_SPACE; k += blockDim.x)
        policy[k] = s_slot >= 0 ? r.policy[(size_t)s_slot * AZ_ACTION_SPACE + k] : 0.0f;
}

// persistence (memory.rs:100-115): entries [first, first + n) in FIFO order (oldest first) out of / into the ring
__global__ void __launch_bounds__(256) k_replay_export(ReplayPtrs r, int first, int n, az_position* __restrict__ pos, float* __restrict__ policy,
                                                       float* __restrict__ value, uint32_t* __restrict__ visits) {
    const int b = blockIdx.x;
    if (b >= n) return;
    const int len = r.state[1];
    const int oldest = len >= r.cap ? r.state[0] : 0;
    const int slot = (oldest + first + b) % r.cap;
    const float4* src = reinterpret_cast<const float4*>(r.policy + (size_t)slot * AZ_ACTION_SPACE);
    float4* dst = reinterpret_cast<float4*>(policy + (size_t)b * AZ_ACTION_SPACE);
    for (int k = threadIdx.x; k < AZ_ACTION_SPACE / 4; k += 256) dst[k] = src[k];
    if (threadIdx.x == 0) { pos[b] = dpos_to_wire(r.keys[slot]); value[b] = r.value[slot]; visits[b] = r.visits[slot]; }
}

// entries appended in order as the newest ones (an entry whose position is already present replaces it in place)
__global__ void __launch_bounds__(1024) k_replay_import(ReplayPtrs r, int n, const az_position* __restrict__ pos, const float* __restrict__ policy,
                                                        const float* __restrict__ value, const uint32_t* __restrict__ visits) {
    __shared__ int s_slot;
    const int t = threadIdx.x;
    int head = r.state[0], len = r.state[1], stale = r.state[3];
    for (int i = 0; i < n; i++) {
        if (t == 0) {
            const DPos key = fen_key_of(dpos_from_wire(pos[i]));
            const u64 h = fen_key_hash(key);
            int slot = replay_find(r, key, h);
            if (slot < 0) {
                slot = head;
                if (len >= r.cap) stale++;   // the evicted position's table entry goes stale
                else len++;
                r.keys[slot] = key;
                replay_table_insert(r, h, slot);
                head = head + 1 == r.cap ? 0 : head + 1;
            }
            r.value[slot] = value[i];
            r.visits[slot] = visits[i];
            s_slot = slot;
        }
        __syncthreads();
        float* pol = r.policy + (size_t)s_slot * AZ_ACTION_SPACE;
        for (int k = t; k < AZ_ACTION_SPACE; k += 1024) pol[k] = policy[(size_t)i * AZ_ACTION_SPACE + k];
        __syncthreads();
    }
    if (t == 0) { r.state[0] = head; r.state[1] = len; r.state[3] = stale; }
}

static inline uint64_t host_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace azb

using namespace azb;

extern "C" {

int az_replay_create(az_engine* e, int capacity, int max_batch, az_replay** out) {
    if (!e || !out || capacity <= 0 || max_batch <= 0) return AZ_ERR_INVALID_ARGUMENT;
    cudaSetDevice(e->cfg.device);
    az_replay* rp = new az_replay;
    rp->eng = e;
    rp->max_batch = max_batch;
    *out = rp;
    ReplayPtrs& p = rp->p;
    p.cap = capacity;
    int ts = 1024;
    while (ts < 4 * capacity + 4 * max_batch) ts <<= 1;   // live + stale entries stay below half of this (k_replay_rebuild)
    p.tmask = ts - 1;
    rp->window = std::max(1, std::min(32, capacity / 2));   // an eviction can never reach back into the same window
    rp->chunk = std::max(rp->window, capacity / 2);
    if (const char* v = getenv("AZ_REPLAY_FORCE_SLOW")) rp->force_slow = atoi(v);
    if (const char* v = getenv("AZ_REPLAY_WINDOW")) rp->window = std::max(1, std::min(rp->window, atoi(v)));
    AZ_CUDA(e, cudaMalloc(&p.keys, (size_t)capacity * sizeof(DPos)));
    AZ_CUDA(e, cudaMalloc(&p.policy, (size_t)capacity * AZ_ACTION_SPACE * sizeof(float)));
    AZ_CUDA(e, cudaMalloc(&p.value, (size_t)capacity * sizeof(float)));
    AZ_CUDA(e, cudaMalloc(&p.visits, (size_t)capacity * sizeof(uint32_t)));
    AZ_CUDA(e, cudaMalloc(&p.table, (size_t)ts * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&p.state, 4 * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&p.last_touch, (size_t)capacity * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&rp->sp.slot, (size_t)rp->chunk * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&rp->sp.old, (size_t)rp->chunk * sizeof(uint32_t)));
    AZ_CUDA(e, cudaMalloc(&rp->sp.prev, (size_t)rp->chunk * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&rp->sp.next, (size_t)rp->chunk * sizeof(int32_t)));
    AZ_CUDA(e, cudaMemset(p.table, 0xFF, (size_t)ts * sizeof(int32_t)));
    AZ_CUDA(e, cudaMemset(p.state, 0, 4 * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&rp->d_idx, (size_t)max_batch * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&rp->d_planes, (size_t)max_batch * AZ_NUM_PLANES * 64 * sizeof(float)));
    AZ_CUDA(e, cudaMalloc(&rp->d_pol, (size_t)max_batch * AZ_ACTION_SPACE * sizeof(float)));
    AZ_CUDA(e, cudaMalloc(&rp->d_val, (size_t)max_batch * sizeof(float)));
    return AZ_OK;
}

void az_replay_destroy(az_replay* rp) {
    if (!rp) return;
    cudaSetDevice(rp->eng->cfg.device);
    cudaStreamSynchronize(rp->eng->stream);
    cudaFree(rp->p.keys); cudaFree(rp->p.policy); cudaFree(rp->p.value); cudaFree(rp->p.visits); cudaFree(rp->p.table);
    cudaFree(rp->p.last_touch); cudaFree(rp->sp.slot); cudaFree(rp->sp.old); cudaFree(rp->sp.prev); cudaFree(rp->sp.next);
    cudaFree(rp->p.state); cudaFree(rp->d_staging); cudaFree(rp->d_idx); cudaFree(rp->d_planes); cudaFree(rp->d_pol); cudaFree(rp->d_val); cudaFree(rp->d_pos);
    delete rp;
}

// re-insert the live slots into a cleared table once enough stale entries have piled up (the kernels decide on the device)
static int replay_maintain(az_replay* rp) {
    az_engine* e = rp->eng;
    const int ts = rp->p.tmask + 1;
    e->n_launches += 3;
    k_replay_rebuild<<<(ts + 255) / 256, 256, 0, e->stream>>>(rp->p, 0);
    k_replay_rebuild<<<(rp->p.cap + 255) / 256, 256, 0, e->stream>>>(rp->p, 1);
    k_replay_rebuild<<<1, 32, 0, e->stream>>>(rp->p, 2);
    AZ_CUDA(e, cudaGetLastError());
    return 0;
}

static int replay_add_dev(az_replay* rp, const az_sample* d_samples, int n, int* new_unique_out) {
    az_engine* e = rp->eng;
    AZ_CUDA(e, cudaMemsetAsync(rp->p.state + 2, 0, sizeof(int32_t), e->stream));   // new_unique of this call
    for (int off = 0; off < n; off += rp->chunk) {
        const int m = std::min(rp->chunk, n - off);
        AZ_CUDA(e, cudaMemsetAsync(rp->p.last_touch, 0xFF, (size_t)rp->p.cap * sizeof(int32_t), e->stream));
        AZ_CUDA(e, cudaMemsetAsync(rp->sp.next, 0xFF, (size_t)m * sizeof(int32_t), e->stream));
        e->n_launches += 2;
        k_replay_resolve<<<1, 32, 0, e->stream>>>(rp->p, rp->sp, d_samples + off, m, rp->window, rp->force_slow);
        k_replay_apply<<<m, 256, 0, e->stream>>>(rp->p, rp->sp, d_samples + off, m, 1.0f / e->cfg.temperature);
        AZ_CUDA(e, cudaGetLastError());
        int r = replay_maintain(rp);
        if (r) return r;
    }
    int32_t st[4] = {0, 0, 0, 0};
    AZ_CUDA(e, cudaMemcpyAsync(st, rp->p.state, sizeof st, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    if (new_unique_out) *new_unique_out = n > 0 ? st[2] : 0;
    return AZ_OK;
}

int az_replay_add(az_replay* rp, const az_sample* samples, int n, int* new_unique_out) {
    if (!rp || n < 0 || (n > 0 && !samples)) return AZ_ERR_INVALID_ARGUMENT;
    az_engine* e = rp->eng;
    cudaSetDevice(e->cfg.device);
    if ((size_t)n > rp->staging_cap) {
        cudaFree(rp->d_staging); rp->d_staging = nullptr; rp->staging_cap = 0;
        AZ_CUDA(e, cudaMalloc(&rp->d_staging, std::max<size_t>(n, 4096) * sizeof(az_sample)));
        rp->staging_cap = std::max<size_t>(n, 4096);
    }
    if (n > 0) AZ_CUDA(e, cudaMemcpyAsync(rp->d_staging, samples, (size_t)n * sizeof(az_sample), cudaMemcpyHostToDevice, e->stream));
    return replay_add_dev(rp, rp->d_staging, n, new_unique_out);
}

int az_replay_add_dev(az_replay* rp, const az_sample* samples_dev, int n, int* new_unique_out) {
    if (!rp || n < 0 || (n > 0 && !samples_dev)) return AZ_ERR_INVALID_ARGUMENT;
    cudaSetDevice(rp->eng->cfg.device);
    return replay_add_dev(rp, samples_dev, n, new_unique_out);
}

int az_replay_add_pending(az_replay* rp, int* n_added_out, int* new_unique_out) {
    if (!rp) return AZ_ERR_INVALID_ARGUMENT;
    az_engine* e = rp->eng;
    cudaSetDevice(e->cfg.device);
    const az_sample* d_samples = nullptr;
    int n = 0;
    int r = search_pending_samples(e, &d_samples, &n);
    if (r) return r;
    r = replay_add_dev(rp, d_samples, n, new_unique_out);
    if (r) return r;
    if (n_added_out) *n_added_out = n;
    return search_clear_pending(e);
}

int az_replay_len(az_replay* rp, int* len_out) {
    if (!rp || !len_out) return AZ_ERR_INVALID_ARGUMENT;
    az_engine* e = rp->eng;
    cudaSetDevice(e->cfg.device);
    int32_t st[4];
    AZ_CUDA(e, cudaMemcpyAsync(st, rp->p.state, sizeof st, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    *len_out = st[1];
    return AZ_OK;
}

// ReplayBuffer::sample (memory.rs:78-97); `to_device`: the three outputs are device pointers (no host hop on the way to the trainer)
static int replay_sample(az_replay* rp, int batch_size, uint64_t seed, float* planes_out, float* policy_out, float* value_out, int* n_out,
                         bool to_device) {
    if (!rp || !n_out || batch_size < 0) return AZ_ERR_INVALID_ARGUMENT;
    az_engine* e = rp->eng;
    cudaSetDevice(e->cfg.device);
    if (batch_size > rp->max_batch) return set_err(e, AZ_ERR_CAPACITY, "batch larger than the replay buffer's max_batch");
    int len = 0;
    int r = az_replay_len(rp, &len);
    if (r) return r;
    const int n = std::min(batch_size, len);   // effective_batch_size (memory.rs:79)
    *n_out = n;
    if (n == 0) return AZ_OK;
    // choose_multiple: n distinct live entries, uniformly (partial Fisher-Yates over the FIFO ranks [0, len))
    std::vector<int32_t> pool(len);
    for (int i = 0; i < len; i++) pool[i] = i;
    uint64_t s = host_splitmix64(seed);
    for (int i = 0; i < n; i++) {
        s = host_splitmix64(s);
        const int j = i + (int)(s % (uint64_t)(len - i));
        std::swap(pool[i], pool[j]);
    }
    AZ_CUDA(e, cudaMemcpyAsync(rp->d_idx, pool.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    e->n_launches++;
    k_replay_gather<<<n, 256, 0, e->stream>>>(rp->p, rp->d_idx, n, rp->d_planes, rp->d_pol, rp->d_val);
    AZ_CUDA(e, cudaGetLastError());
    const cudaMemcpyKind kind = to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (planes_out) AZ_CUDA(e, cudaMemcpyAsync(planes_out, rp->d_planes, (size_t)n * AZ_NUM_PLANES * 64 * 4, kind, e->stream));
    if (policy_out) AZ_CUDA(e, cudaMemcpyAsync(policy_out, rp->d_pol, (size_t)n * AZ_ACTION_SPACE * 4, kind, e->stream));
    if (value_out) AZ_CUDA(e, cudaMemcpyAsync(value_out, rp->d_val, (size_t)n * 4, kind, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    return AZ_OK;
}

int az_replay_sample(az_replay* rp, int batch_size, uint64_t seed, float* planes_out, float* policy_out, float* value_out, int* n_out) {
    return replay_sample(rp, batch_size, seed, planes_out, policy_out, value_out, n_out, false);
}

int az_replay_sample_dev(az_replay* rp, int batch_size, uint64_t seed, float* planes_dev, float* policy_dev, float* value_dev, int* n_out) {
    return replay_sample(rp, batch_size, seed, planes_dev, policy_dev, value_dev, n_out, true);
}

int az_replay_export(az_replay* rp, int first, int n, az_position* pos_out, float* policy_out, float* value_out, uint32_t* visits_out,
                     int* n_out) {
    if (!rp || !n_out || first < 0 || n < 0 || !pos_out || !policy_out || !value_out || !visits_out) return AZ_ERR_INVALID_ARGUMENT;
    az_engine* e = rp->eng;
    cudaSetDevice(e->cfg.device);
    if (n > rp->max_batch) return set_err(e, AZ_ERR_CAPACITY, "page larger than the replay buffer's max_batch");
    int len = 0;
    int r = az_replay_len(rp, &len);
    if (r) return r;
    const int m = std::max(0, std::min(n, len - first));
    *n_out = m;
    if (m == 0) return AZ_OK;
    if (!rp->d_pos) AZ_CUDA(e, cudaMalloc(&rp->d_pos, (size_t)rp->max_batch * sizeof(az_position)));
    e->n_launches++;
    k_replay_export<<<m, 256, 0, e->stream>>>(rp->p, first, m, rp->d_pos, rp->d_pol, rp->d_val, reinterpret_cast<uint32_t*>(rp->d_idx));
    AZ_CUDA(e, cudaGetLastError());
    AZ_CUDA(e, cudaMemcpyAsync(pos_out, rp->d_pos, (size_t)m * sizeof(az_position), cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(policy_out, rp->d_pol, (size_t)m * AZ_ACTION_SPACE * 4, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(value_out, rp->d_val, (size_t)m * 4, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(visits_out, rp->d_idx, (size_t)m * 4, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    return AZ_OK;
}

int az_replay_import(az_replay* rp, int n, const az_position* pos, const float* policy, const float* value, const uint32_t* visits) {
    if (!rp || n < 0 || (n > 0 && (!pos || !policy || !value || !visits))) return AZ_ERR_INVALID_ARGUMENT;
    az_engine* e = rp->eng;
    cudaSetDevice(e->cfg.device);
    if (n > rp->max_batch) return set_err(e, AZ_ERR_CAPACITY, "page larger than the replay buffer's max_batch");
    if (n == 0) return AZ_OK;
    for (int i = 0; i < n; i++)
        if (visits[i] == 0) return set_err(e, AZ_ERR_INVALID_ARGUMENT, "a stored entry has visit_count >= 1");
    if (!rp->d_pos) AZ_CUDA(e, cudaMalloc(&rp->d_pos, (size_t)rp->max_batch * sizeof(az_position)));
    AZ_CUDA(e, cudaMemcpyAsync(rp->d_pos, pos, (size_t)n * sizeof(az_position), cudaMemcpyHostToDevice, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(rp->d_pol, policy, (size_t)n * AZ_ACTION_SPACE * 4, cudaMemcpyHostToDevice, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(rp->d_val, value, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(rp->d_idx, visits, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
    e->n_launches++;
    k_replay_import<<<1, 1024, 0, e->stream>>>(rp->p, n, rp->d_pos, rp->d_pol, rp->d_val, reinterpret_cast<const uint32_t*>(rp->d_idx));
    AZ_CUDA(e, cudaGetLastError());
    if (int r = replay_maintain(rp)) return r;
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    return AZ_OK;
}

int az_replay_get(az_replay* rp, const az_position* pos, float* policy_out, float* value_out, uint32_t* visit_count_out) {
    if (!rp || !pos || !policy_out || !value_out || !visit_count_out) return AZ_ERR_INVALID_ARGUMENT;
    az_engine* e = rp->eng;
    cudaSetDevice(e->cfg.device);
    AZ_CUDA(e, cudaMemcpyAsync(e->d_wire, pos, sizeof(az_position), cudaMemcpyHostToDevice, e->stream));
    k_replay_get<<<1, 256, 0, e->stream>>>(rp->p, e->d_wire, rp->d_pol, rp->d_val, reinterpret_cast<uint32_t*>(rp->d_idx));
    AZ_CUDA(e, cudaGetLastError());
    AZ_CUDA(e, cudaMemcpyAsync(policy_out, rp->d_pol, AZ_ACTION_SPACE * 4, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(value_out, rp->d_val, 4, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaMemcpyAsync(visit_count_out, rp->d_idx, 4, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    return AZ_OK;
}

}  // extern "C"
