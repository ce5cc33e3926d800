//! Safe wrapper over `az-b200-sys` for the reference crate (agent.rs / training.rs / memory.rs / validation.rs).
//!
//! SOURCE ONLY: never compiled in the image this project is built in (no Rust toolchain); the C ABI underneath is what
//! the test-suite exercises (through the Python and C++ mirrors).  The types here are the ABI's own (`az_position`,
//! `az_sample`); converting `shakmaty::Chess` to and from `az_position` is a field copy that belongs in the reference
//! crate, so this crate has no dependency on shakmaty.
use az_b200_sys as sys;
use std::ffi::CStr;
use std::ptr;

pub const ACTION_SPACE: usize = 4096; // parameters.rs:3
pub const MAX_MOVES: usize = 256;

#[derive(Debug)]
pub struct Error { pub code: i32, pub message: String }
pub type Result<T> = std::result::Result<T, Error>;

/// One engine per GPU; calls are serialised by the owner (the reference's single inference-loop thread).
pub struct Engine { raw: *mut sys::az_engine, cfg: sys::az_config }
unsafe impl Send for Engine {}

impl Engine {
    pub fn default_config() -> sys::az_config {
        let mut c = std::mem::MaybeUninit::<sys::az_config>::uninit();
        unsafe { sys::az_config_default(c.as_mut_ptr()); c.assume_init() }
    }
    pub fn new(cfg: sys::az_config) -> Result<Engine> {
        let mut raw: *mut sys::az_engine = ptr::null_mut();
        let rc = unsafe { sys::az_engine_create(&cfg, &mut raw) };
        let e = Engine { raw, cfg };
        if rc != 0 { return Err(e.error(rc)); }
        Ok(e)
    }
    fn error(&self, code: i32) -> Error {
        let message = if self.raw.is_null() { String::from("az_engine_create failed") } else {
            unsafe { CStr::from_ptr(sys::az_last_error(self.raw)) }.to_string_lossy().into_owned()
        };
        Error { code, message }
    }
    fn check(&self, rc: i32) -> Result<()> { if rc == 0 { Ok(()) } else { Err(self.error(rc)) } }
    pub fn config(&self) -> &sys::az_config { &self.cfg }

    /// load_model (main.rs:109-116): the 144 tensors of the burn record in `az_weight_name` order.
    pub fn load_weights(&mut self, arrays: &[&[f32]]) -> Result<()> {
        let ptrs: Vec<*const f32> = arrays.iter().map(|a| a.as_ptr()).collect();
        self.check(unsafe { sys::az_load_weights(self.raw, ptrs.as_ptr(), ptrs.len() as i32) })
    }

    /// process_batch (training.rs:380-422) without the channel plumbing: N positions in, N (policy, value) out.
    pub fn forward(&mut self, positions: &[sys::az_position]) -> Result<(Vec<f32>, Vec<f32>)> {
        let n = positions.len();
        let mut policy = vec![0f32; n * ACTION_SPACE];
        let mut value = vec![0f32; n];
        self.check(unsafe { sys::az_forward(self.raw, n as i32, positions.as_ptr(), policy.as_mut_ptr(), value.as_mut_ptr()) })?;
        Ok((policy, value))
    }

    /// Chess::legal_moves() for a batch: (wire moves, policy indices, counts).
    pub fn legal_moves(&mut self, positions: &[sys::az_position]) -> Result<(Vec<u16>, Vec<u16>, Vec<i32>)> {
        let n = positions.len();
        let (mut moves, mut index, mut count) = (vec![0u16; n * MAX_MOVES], vec![0u16; n * MAX_MOVES], vec![0i32; n]);
        self.check(unsafe { sys::az_movegen(self.raw, n as i32, positions.as_ptr(), moves.as_mut_ptr(), index.as_mut_ptr(), count.as_mut_ptr()) })?;
        Ok((moves, index, count))
    }

    /// MCTree::init(model, state, noise) + monte_carlo_tree_search (validation.rs:39-40): `history[i]` holds every position
    /// already counted in that game's `pos_count` (including the root).  Returns visits [n][4096] and max_subtree_depth [n];
    /// `visits / sims` is the improved policy (tree.rs:173-175 with TEMPERATURE = 1).
    pub fn search(&mut self, roots: &[sys::az_position], history: &[Vec<sys::az_position>], sims: i32,
                  noise: Option<(&[u64], &[u32])>) -> Result<(Vec<f32>, Vec<i32>)> {
        let n = roots.len();
        let mut flat = Vec::new();
        let mut offs = vec![0u32; n + 1];
        for (i, h) in history.iter().enumerate() { flat.extend_from_slice(h); offs[i + 1] = flat.len() as u32; }
        let (ids, plies) = match noise { Some((g, p)) => (g.as_ptr(), p.as_ptr()), None => (ptr::null(), ptr::null()) };
        let mut visits = vec![0f32; n * ACTION_SPACE];
        let mut depth = vec![0i32; n];
        self.check(unsafe { sys::az_search(self.raw, n as i32, roots.as_ptr(), flat.as_ptr(), offs.as_ptr(), sims, ids, plies,
                                           visits.as_mut_ptr(), ptr::null_mut(), depth.as_mut_ptr()) })?;
        Ok((visits, depth))
    }

    /// run_all_episodes (training.rs:340-378): EXACTLY `n_games` self-play games (ids first_game_id ..), each played to
    /// completion; the callback receives the finished games' steps (values already back-filled) as they become available.
    /// Returns the average batch size (= evaluations per wave).
    pub fn run_all_episodes<F: FnMut(&[sys::az_sample])>(&mut self, n_games: i32, first_game_id: u64, mut sink: F) -> Result<f32> {
        self.check(unsafe { sys::az_selfplay_begin_n(self.raw, n_games, first_game_id, n_games as u64) })?;
        let mut buf: Vec<sys::az_sample> = Vec::with_capacity((n_games as usize * 128).max(1 << 16));
        let mut stats = sys::az_selfplay_stats::default();
        let mut waves = 0u64;
        loop {
            self.check(unsafe { sys::az_selfplay_step(self.raw, 64, &mut stats) })?;
            waves += 64;
            if stats.pending_samples > 0 {
                let mut n = 0i32;
                self.check(unsafe { sys::az_selfplay_drain(self.raw, buf.as_mut_ptr(), buf.capacity() as i32, &mut n) })?;
                unsafe { buf.set_len(n as usize) };
                sink(&buf);
            }
            if stats.active_games == 0 && stats.pending_samples == 0 { break; }   // every game of the generation has ended
        }
        Ok(stats.evaluations as f32 / waves as f32)
    }

    /// get_best_move up to its random tie-break (chess.rs:295-318): negamax score of every legal move of `pos`.
    pub fn minimax_scores(&mut self, pos: &sys::az_position, depth: i32) -> Result<Vec<i32>> {
        let mut scores = vec![0i32; MAX_MOVES];
        let mut count = 0i32;
        self.check(unsafe { sys::az_minimax(self.raw, 1, pos, depth, scores.as_mut_ptr(), &mut count) })?;
        scores.truncate(count as usize);
        Ok(scores)
    }
}

impl Drop for Engine {
    fn drop(&mut self) { if !self.raw.is_null() { unsafe { sys::az_engine_destroy(self.raw) } } }
}

/// improved_policy of one EpisodeStep (training.rs:303-308) from the sparse record the engine emits.
pub fn improved_policy(sample: &sys::az_sample, sims: u32) -> Box<[f32; ACTION_SPACE]> {
    let mut policy = Box::new([0f32; ACTION_SPACE]);
    for k in 0..sample.n_visits as usize { policy[sample.index[k] as usize] = sample.count[k] as f32 / sims as f32; }
    policy
}

/// ReplayBuffer (memory.rs:26-118), device resident.  The buffer must not outlive its engine.
pub struct ReplayBuffer { raw: *mut sys::az_replay }

impl ReplayBuffer {
    pub fn new(engine: &mut Engine, capacity: i32, max_batch: i32) -> Result<ReplayBuffer> {
        let mut raw: *mut sys::az_replay = ptr::null_mut();
        let rc = unsafe { sys::az_replay_create(engine.raw, capacity, max_batch, &mut raw) };
        if rc != 0 { return Err(engine.error(rc)); }
        Ok(ReplayBuffer { raw })
    }
    /// add() for every finished self-play step still in device memory; returns (steps, new unique positions).
    pub fn add_pending(&mut self) -> (i32, i32) {
        let (mut n, mut nu) = (0i32, 0i32);
        unsafe { sys::az_replay_add_pending(self.raw, &mut n, &mut nu) };
        (n, nu)
    }
    pub fn len(&self) -> usize {
        let mut n = 0i32;
        unsafe { sys::az_replay_len(self.raw, &mut n) };
        n as usize
    }
    /// sample(batch_size) (memory.rs:78-97): stacked planes [n,19,8,8], policies [n,4096] and values [n].
    pub fn sample(&mut self, batch_size: usize, seed: u64) -> (usize, Vec<f32>, Vec<f32>, Vec<f32>) {
        let (mut planes, mut policy, mut value) = (vec![0f32; batch_size * 19 * 64], vec![0f32; batch_size * ACTION_SPACE], vec![0f32; batch_size]);
        let mut n = 0i32;
        unsafe { sys::az_replay_sample(self.raw, batch_size as i32, seed, planes.as_mut_ptr(), policy.as_mut_ptr(), value.as_mut_ptr(), &mut n) };
        let n = n as usize;
        planes.truncate(n * 19 * 64); policy.truncate(n * ACTION_SPACE); value.truncate(n);
        (n, planes, policy, value)
    }
}

impl Drop for ReplayBuffer {
    fn drop(&mut self) { if !self.raw.is_null() { unsafe { sys::az_replay_destroy(self.raw) } } }
}
