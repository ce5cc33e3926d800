// Device-resident replay buffer (memory.rs:27-118): positions de-duplicated by their FEN identity (pseudo-legal ep,
// counters included), running means of policy and value over repeated visits, FIFO eviction of the oldest UNIQUE
// position at capacity, uniform sampling without replacement.  It consumes the az_sample records self-play leaves in
// device memory, so a generation's positions never travel to the host.
//
// `add` is order dependent in the reference: the running mean (old*n + new)/(n+1) is evaluated in f32 in the order the
// steps arrive, and which position is evicted depends on the insertion order of the unique ones.  To stay bit-exact one
// CTA applies a batch strictly in order (a step costs about a microsecond: a hash probe by one thread, then a 16 KB
// read-modify-write of the dense policy row by 1024 threads); it runs once per generation, off the self-play hot path.
// Sampling and batch assembly (planes, policy rows, values) are parallel over the batch.
#include "engine.h"
#include "mcts.h"
#include <algorithm>
#include <cstring>
#include <vector>

namespace azb {

struct ReplayPtrs {
    DPos* keys;            // [cap] slot -> key (FEN identity)
    float* policy;         // [cap][4096] running mean of improved_policy
    float* value;          // [cap]
    uint32_t* visits;      // [cap] visit_count
    int32_t* table;        // [tsize] open addressing, linear probing with backward-shift deletion: slot or -1
    int32_t* state;        // [4]: head (next slot, = oldest when full), len, new_unique (of the last add), reserved
    int cap, tmask;
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
// remove the table entry pointing at `slot` (backward-shift deletion keeps probe sequences intact without tombstones)
__device__ __forceinline__ void replay_table_erase(const ReplayPtrs& r, int slot) {
    uint32_t i = (uint32_t)fen_key_hash(r.keys[slot]) & r.tmask;
    while (r.table[i] != slot) i = (i + 1) & r.tmask;
    uint32_t hole = i;
    for (uint32_t j = (hole + 1) & r.tmask;; j = (j + 1) & r.tmask) {
        const int s = r.table[j];
        if (s < 0) break;
        const uint32_t home = (uint32_t)fen_key_hash(r.keys[s]) & r.tmask;
        // the entry at j may move into the hole if its home position does not lie cyclically in (hole, j]
        const bool between = hole <= j ? (home > hole && home <= j) : (home > hole || home <= j);
        if (!between) { r.table[hole] = s; hole = j; }
    }
    r.table[hole] = -1;
}

// ReplayBuffer::add for a batch of EpisodeSteps, strictly in order (memory.rs:41-76)
__global__ void __launch_bounds__(1024) k_replay_add(ReplayPtrs r, const az_sample* __restrict__ samples, int n, float sims) {
    __shared__ float s_new[AZ_ACTION_SPACE];
    __shared__ int s_slot, s_is_new;
    __shared__ float s_old;
    const int t = threadIdx.x;
    int head = r.state[0], len = r.state[1], new_unique = 0;
    for (int i = 0; i < n; i++) {
        const az_sample* sm = samples + i;
        for (int k = t; k < AZ_ACTION_SPACE; k += 1024) s_new[k] = 0.0f;
        __syncthreads();
        const int nv = sm->n_visits;
        for (int k = t; k < nv; k += 1024) s_new[sm->index[k]] = __fdiv_rn((float)sm->count[k], sims);   // improved_policy
        if (t == 0) {
            const DPos key = fen_key_of(dpos_from_wire(sm->position));
            const u64 h = fen_key_hash(key);
            int slot = replay_find(r, key, h);
            if (slot >= 0) {
                const float old_count = (float)r.visits[slot], total = __fadd_rn(old_count, 1.0f);
                r.value[slot] = __fdiv_rn(__fadd_rn(__fmul_rn(r.value[slot], old_count), sm->final_value), total);
                r.visits[slot] += 1;
                s_is_new = 0; s_old = old_count;
            } else {
                slot = head;
                if (len >= r.cap) replay_table_erase(r, slot);   // order.pop_front(): the oldest unique position goes
                else len++;
                r.keys[slot] = key;
                r.value[slot] = sm->final_value;
                r.visits[slot] = 1;
                replay_table_insert(r, h, slot);
                head = head + 1 == r.cap ? 0 : head + 1;
                new_unique++;
                s_is_new = 1; s_old = 0.0f;
            }
            s_slot = slot;
        }
        __syncthreads();
        float* pol = r.policy + (size_t)s_slot * AZ_ACTION_SPACE;
        if (s_is_new) {
            for (int k = t; k < AZ_ACTION_SPACE; k += 1024) pol[k] = s_new[k];
        } else {
            const float old_count = s_old, total = __fadd_rn(old_count, 1.0f);
            for (int k = t; k < AZ_ACTION_SPACE; k += 1024) pol[k] = __fdiv_rn(__fadd_rn(__fmul_rn(pol[k], old_count), s_new[k]), total);
        }
        __syncthreads();
    }
    if (t == 0) { r.state[0] = head; r.state[1] = len; r.state[2] = new_unique; }
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
    int head = r.state[0], len = r.state[1];
    for (int i = 0; i < n; i++) {
        if (t == 0) {
            const DPos key = fen_key_of(dpos_from_wire(pos[i]));
            const u64 h = fen_key_hash(key);
            int slot = replay_find(r, key, h);
            if (slot < 0) {
                slot = head;
                if (len >= r.cap) replay_table_erase(r, slot);
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
    if (t == 0) { r.state[0] = head; r.state[1] = len; }
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
    while (ts < 2 * capacity) ts <<= 1;
    p.tmask = ts - 1;
    AZ_CUDA(e, cudaMalloc(&p.keys, (size_t)capacity * sizeof(DPos)));
    AZ_CUDA(e, cudaMalloc(&p.policy, (size_t)capacity * AZ_ACTION_SPACE * sizeof(float)));
    AZ_CUDA(e, cudaMalloc(&p.value, (size_t)capacity * sizeof(float)));
    AZ_CUDA(e, cudaMalloc(&p.visits, (size_t)capacity * sizeof(uint32_t)));
    AZ_CUDA(e, cudaMalloc(&p.table, (size_t)ts * sizeof(int32_t)));
    AZ_CUDA(e, cudaMalloc(&p.state, 4 * sizeof(int32_t)));
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
    cudaFree(rp->p.state); cudaFree(rp->d_staging); cudaFree(rp->d_idx); cudaFree(rp->d_planes); cudaFree(rp->d_pol); cudaFree(rp->d_val); cudaFree(rp->d_pos);
    delete rp;
}

static int replay_add_dev(az_replay* rp, const az_sample* d_samples, int n, int* new_unique_out) {
    az_engine* e = rp->eng;
    if (n > 0) {
        e->n_launches++;
        k_replay_add<<<1, 1024, 0, e->stream>>>(rp->p, d_samples, n, (float)e->cfg.num_simulations);
        AZ_CUDA(e, cudaGetLastError());
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

int az_replay_sample(az_replay* rp, int batch_size, uint64_t seed, float* planes_out, float* policy_out, float* value_out, int* n_out) {
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
    if (planes_out) AZ_CUDA(e, cudaMemcpyAsync(planes_out, rp->d_planes, (size_t)n * AZ_NUM_PLANES * 64 * 4, cudaMemcpyDeviceToHost, e->stream));
    if (policy_out) AZ_CUDA(e, cudaMemcpyAsync(policy_out, rp->d_pol, (size_t)n * AZ_ACTION_SPACE * 4, cudaMemcpyDeviceToHost, e->stream));
    if (value_out) AZ_CUDA(e, cudaMemcpyAsync(value_out, rp->d_val, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
    AZ_CUDA(e, cudaStreamSynchronize(e->stream));
    return AZ_OK;
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
