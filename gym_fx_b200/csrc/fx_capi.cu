// fx_capi.cu -- the C-ABI of libfxenv.so (include/fxenv.h): handle management, device memory, kernel launches.
// No torch types, no exceptions across the boundary; every entry point returns a status code.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include "fx_kernels.cuh"
#include "fx_policy.cuh"

namespace {

thread_local std::string g_create_error;

struct SlabPlan {
  size_t bytes = 0;
  size_t add(size_t nbytes) {
    const size_t off = bytes;
    bytes += (nbytes + 255) & ~(size_t)255;
    return off;
  }
};

}  // namespace

struct FxEnv {
  FxKernelParams P;
  int device = 0;
  unsigned char* slab = nullptr;
  size_t slab_bytes = 0;
  double* candles_dev[FXENV_MAX_PAIRS] = {};
  double* stats_dev[FXENV_MAX_PAIRS] = {};
  int64_t* minutes_dev[FXENV_MAX_PAIRS] = {};
  bool loaded[FXENV_MAX_PAIRS] = {};
  bool tame[FXENV_MAX_PAIRS] = {};
  bool was_reset = false;
  bool first_reset = true;
  int timeline_steps = 0;
  int force_engine = -1;           // FXENV_ENGINE (timing experiments): 0 graph of single steps, 1 persistent launch
  bool seq_tracked = true;         // fx_rollout_kernel's seq[] / ticket words hold (seq_base, ticket_base): no memset needed
  unsigned seq_base = 0u, ticket_base = 0u;
  int64_t launches = 0;
  std::string err;
  // fxenv_step_host staging
  static constexpr int kHostSlices = 8;
  cudaStream_t hstream = nullptr, hcopy = nullptr;
  cudaEvent_t hev[kHostSlices] = {};
  void* h_actions = nullptr;
  float* h_obs = nullptr;
  float* h_reward = nullptr;
  uint8_t* h_term = nullptr;
  // fxenv_step_many graph cache (one entry: the last pointer set)
  // fxenv_step_many, graph engine: the two most recent launch sequences, keyed by the pointer set and sizes
  struct CachedGraph {
    cudaGraphExec_t exec = nullptr;
    const void* actions = nullptr;
    float* obs = nullptr;
    float* reward = nullptr;
    uint8_t* term = nullptr;
    int steps = 0, slots = 0;
    uint64_t used = 0;
  } graphs[2];
  uint64_t graph_clock = 0;
};

namespace {

int fail(FxEnv* env, int code, const std::string& msg) {
  if (env) env->err = msg; else g_create_error = msg;
  return code;
}

int cuda_fail(FxEnv* env, cudaError_t e, const char* what) {
  return fail(env, FXENV_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define FX_CUDA(env, call)                                   \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return cuda_fail(env, e__, #call); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int validate(const FxConfig& c, std::string& why) {
  char buf[256];
#define BAD(...) do { snprintf(buf, sizeof buf, __VA_ARGS__); why = buf; return FXENV_E_INVALID; } while (0)
  if (c.struct_size != (int32_t)sizeof(FxConfig)) BAD("FxConfig.struct_size %d != %d (ABI mismatch)", c.struct_size, (int)sizeof(FxConfig));
  if (c.num_envs < 1) BAD("num_envs must be >= 1");
  if (c.num_pairs < 1 || c.num_pairs > FXENV_MAX_PAIRS) BAD("num_pairs must be in 1..%d", FXENV_MAX_PAIRS);
  if (c.n_cols < 5 || c.n_cols > FXENV_MAX_COLS) BAD("n_cols must be in 5..%d", FXENV_MAX_COLS);
  if (c.order_capacity < 0 || c.order_capacity > 512) BAD("order_capacity must be in 0..512");
  if (c.window_size < 1) BAD("window_size must be >= 1");
  if (c.price_col < 0 || c.price_col >= c.n_cols) BAD("price_col out of range");
  if (!(c.slippage_perc >= 0.0 && c.slippage_perc < 1.0)) BAD("slippage_perc must be in [0, 1)");
  if (!(c.leverage > 0.0)) BAD("leverage must be > 0");
  if (c.strategy < 0 || c.strategy > FX_STRATEGY_ATR_SLTP) BAD("unknown strategy %d", c.strategy);
  if (c.strategy == FX_STRATEGY_ATR_SLTP && (c.atr_period < 1 || c.atr_period > 64)) BAD("atr_period must be in 1..64");
  if (c.strategy != FX_STRATEGY_DEFAULT && c.strat_position_size == 0.0 && !c.use_rel_volume) BAD("bracket strategies need position_size != 0");
  if (c.preproc < 0 || c.preproc > FX_PREPROC_FEATURE_WINDOW) BAD("unknown preprocessor %d", c.preproc);
  if (c.preproc == FX_PREPROC_FEATURE_WINDOW) {
    if (c.n_features < 1 || c.n_features > FXENV_MAX_FEATURES) BAD("n_features must be in 1..%d", FXENV_MAX_FEATURES);
    for (int i = 0; i < c.n_features; i++)
      if (c.feature_cols[i] < 0 || c.feature_cols[i] >= c.n_cols) BAD("feature_cols[%d] out of range", i);
    if (c.scaling < 0 || c.scaling > FX_SCALING_EXPANDING) BAD("unknown feature scaling %d", c.scaling);
    if (c.scaling == FX_SCALING_ROLLING && c.scaling_window < 1) BAD("feature_scaling_window must be >= 1");
  }
  if (c.reward < 0 || c.reward > FX_REWARD_DD) BAD("unknown reward %d", c.reward);
  if (c.reward == FX_REWARD_SHARPE && (c.sharpe_window < 1 || c.sharpe_window > 4096)) BAD("sharpe window must be in 1..4096");
  if (c.reward_initial_cash == 0.0) BAD("reward_initial_cash must be non-zero (the reference substitutes 1.0)");
  if (c.episode_bars < 0) BAD("episode_bars must be >= 0");
#undef BAD
  return FXENV_OK;
}

int require_ready(FxEnv* env, bool need_reset) {
  if (!env) return FXENV_E_INVALID;
  for (int p = 0; p < env->P.cfg.num_pairs; p++)
    if (!env->loaded[p]) return fail(env, FXENV_E_STATE, "candles for pair " + std::to_string(p) + " not loaded");
  if (need_reset && !env->was_reset) return fail(env, FXENV_E_STATE, "Call reset() before step().");
  return FXENV_OK;
}

void drop_graph(FxEnv* env) {
  for (auto& g : env->graphs)
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
}

}  // namespace

extern "C" {

int fxenv_abi_version(void) { return FXENV_ABI_VERSION; }

const char* fxenv_last_error(const FxEnv* env) { return env ? env->err.c_str() : g_create_error.c_str(); }

int fxenv_destroy(FxEnv* env);

int fxenv_create(const FxConfig* cfg, FxEnv** out) {
  if (!cfg || !out) return fail(nullptr, FXENV_E_INVALID, "null argument");
  *out = nullptr;
  std::string why;
  int rc = validate(*cfg, why);
  if (rc != FXENV_OK) return fail(nullptr, rc, why);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(nullptr, FXENV_E_CUDA, std::string("no CUDA device available (libfxenv has no CPU path): ") +
                                           (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0"));
  FxEnv* env = new (std::nothrow) FxEnv();
  if (!env) return fail(nullptr, FXENV_E_NOMEM, "out of host memory");
  memset(&env->P, 0, sizeof env->P);
  env->P.cfg = *cfg;
  FxConfig& c = env->P.cfg;
  if (cudaGetDevice(&env->device) != cudaSuccess) { delete env; return fail(nullptr, FXENV_E_CUDA, "cudaGetDevice failed"); }
  int cap = c.order_capacity == 0 ? 128 : c.order_capacity;
  cap = (cap + 31) & ~31;
  c.order_capacity = cap;
  env->P.cap = cap;
  env->P.obs_dim = (int32_t)fx_obs_dim(c);
  env->P.inv_initial_cash = 1.0 / (c.initial_cash != 0.0 ? c.initial_cash : 1.0);
  if (const char* dv = getenv("FXENV_DEBUG")) env->P.debug = atoi(dv);  // timing experiments only
  if (const char* tv = getenv("FXENV_TIMING")) {
    if (atoi(tv)) {
      cudaMalloc(&env->P.timing, (size_t)c.num_envs * 2 * FX_NSTAMP * sizeof(long long));
      cudaMemset(env->P.timing, 0, (size_t)c.num_envs * 2 * FX_NSTAMP * sizeof(long long));
    }
  }
  env->P.fast_features = 0;
  if (c.preproc == FX_PREPROC_FEATURE_WINDOW && c.n_features == 5 && c.n_cols == 5) {
    env->P.fast_features = 5;
    for (int i = 0; i < 5; i++) if (c.feature_cols[i] != i) env->P.fast_features = 0;
  }
  if (const char* tl = getenv("FXENV_TIMELINE")) {  // debug, timing build: per-ticket start / end stamps of fxenv_step_many
    if (atoi(tl) > 0) {
      env->timeline_steps = atoi(tl);
      const size_t nb = (size_t)env->timeline_steps * c.num_envs * 2 * sizeof(long long);
      cudaMalloc(&env->P.timeline, nb);
      cudaMemset(env->P.timeline, 0, nb);
    }
  }
  if (const char* fe = getenv("FXENV_ENGINE")) env->force_engine = (fe[0] == 'p') ? 1 : (fe[0] == 'g' ? 0 : -1);
  env->P.any_binary = 0;
  for (int i = 0; i < c.n_features; i++) if (c.feature_binary[i]) env->P.any_binary = 1;
  if (c.preproc != FX_PREPROC_FEATURE_WINDOW) env->P.any_binary = 0;
  env->P.lean = 0;  // decided in fxenv_load_candles (needs to know that the data are finite)
  ce = fx_configure_kernels(env->P);
  if (const char* rb = getenv("FXENV_ROLLOUT_BLOCKS"))  // timing experiments only: grid of the persistent launch
    if (atoi(rb) > 0 && atoi(rb) < env->P.resident_blocks) env->P.resident_blocks = atoi(rb);
  if (ce != cudaSuccess) { fxenv_destroy(env); return cuda_fail(nullptr, ce, "window_size * n_cols too large for shared memory"); }
  // one slab for the whole per-env state (snapshot == one memcpy)
  const size_t N = (size_t)c.num_envs;
  const size_t ring = (c.reward == FX_REWARD_SHARPE) ? (size_t)c.sharpe_window : 1;
  SlabPlan plan;
  const size_t capP = (size_t)cap + FXO_SLACK;
  size_t o_d[9], o_start, o_i[11], o_flags, o_ring, o_welford, o_meta, o_p0, o_p1, o_sz, o_nbar, o_rstats;
  for (int i = 0; i < 9; i++) o_d[i] = plan.add(N * 8);
  o_start = plan.add(N * 8);
  o_nbar = plan.add(N * 6 * 8);
  o_rstats = plan.add(N * FX_RS_N * 8);
  for (int i = 0; i < 11; i++) o_i[i] = plan.add(N * 4);
  o_flags = plan.add(N * 4);
  o_ring = plan.add(N * ring * 8);
  o_welford = plan.add(N * FXENV_MAX_FEATURES * 2 * 8);
  o_meta = plan.add(N * capP * 4);
  o_p0 = plan.add(N * capP * 8);
  o_p1 = plan.add(N * capP * 8);
  o_sz = plan.add(N * capP * 8);
  env->slab_bytes = plan.bytes;
  ce = cudaMalloc(&env->slab, plan.bytes);
  if (ce != cudaSuccess) { fxenv_destroy(env); return cuda_fail(nullptr, ce, "cudaMalloc(state slab)"); }
  cudaMemset(env->slab, 0, plan.bytes);
  ce = cudaMalloc(&env->P.seq, (N + 1) * sizeof(int32_t));
  if (ce != cudaSuccess) { fxenv_destroy(env); return cuda_fail(nullptr, ce, "cudaMalloc(seq)"); }  // frees what was allocated so far
  cudaMemset(env->P.seq, 0, (N + 1) * sizeof(int32_t));
  FxDeviceState& st = env->P.st;
  unsigned char* b = env->slab;
  double** dcols[9] = {&st.cash, &st.psize, &st.pprice, &st.equity, &st.prev_equity, &st.price,
                       &st.commission_paid, &st.dd_peak, &st.sub_need};
  for (int i = 0; i < 9; i++) *dcols[i] = reinterpret_cast<double*>(b + o_d[i]);
  st.start = reinterpret_cast<int64_t*>(b + o_start);
  st.nbar = reinterpret_cast<double*>(b + o_nbar);
  st.rstats = reinterpret_cast<double*>(b + o_rstats);
  int32_t** icols[11] = {&st.t, &st.total_bars, &st.position, &st.bar_index, &st.trades, &st.n_orders,
                         &st.sh_len, &st.sh_head, &st.sh_last_step, &st.dd_last_step, &st.n_acc};
  for (int i = 0; i < 11; i++) *icols[i] = reinterpret_cast<int32_t*>(b + o_i[i]);
  st.flags = reinterpret_cast<uint32_t*>(b + o_flags);
  st.sh_ring = reinterpret_cast<double*>(b + o_ring);
  st.welford = reinterpret_cast<double*>(b + o_welford);
  st.o_meta = reinterpret_cast<uint32_t*>(b + o_meta);
  st.o_p0 = reinterpret_cast<double*>(b + o_p0);
  st.o_p1 = reinterpret_cast<double*>(b + o_p1);
  st.o_sz = reinterpret_cast<double*>(b + o_sz);
  *out = env;
  return FXENV_OK;
}

int fxenv_destroy(FxEnv* env) {
  if (!env) return FXENV_OK;
  DeviceGuard g(env->device);
  drop_graph(env);
  for (int p = 0; p < FXENV_MAX_PAIRS; p++) {
    cudaFree(env->candles_dev[p]); cudaFree(env->stats_dev[p]); cudaFree(env->minutes_dev[p]);
  }
  cudaFree(env->slab);
  cudaFree(env->P.seq);
  cudaFree(env->P.timing);
  cudaFree(env->P.timeline);
  cudaFree(env->h_actions); cudaFree(env->h_obs); cudaFree(env->h_reward); cudaFree(env->h_term);
  if (env->hstream) cudaStreamDestroy(env->hstream);
  if (env->hcopy) cudaStreamDestroy(env->hcopy);
  for (auto& ev : env->hev) if (ev) cudaEventDestroy(ev);
  delete env;
  return FXENV_OK;
}

int fxenv_load_candles(FxEnv* env, int pair_id, const double* candles_host, int64_t T, const int64_t* minutes_host) {
  if (!env) return FXENV_E_INVALID;
  const FxConfig& c = env->P.cfg;
  if (pair_id < 0 || pair_id >= c.num_pairs) return fail(env, FXENV_E_INVALID, "pair_id out of range");
  if (!candles_host) return fail(env, FXENV_E_INVALID, "candles_host is null");
  // app/env.py:64-65: the data must be longer than the window
  if (T < (int64_t)c.window_size + 2) return fail(env, FXENV_E_INVALID, "input data is empty or too short for the configured window");
  DeviceGuard g(env->device);
  drop_graph(env);
  cudaFree(env->candles_dev[pair_id]); cudaFree(env->stats_dev[pair_id]); cudaFree(env->minutes_dev[pair_id]);
  env->candles_dev[pair_id] = nullptr; env->stats_dev[pair_id] = nullptr; env->minutes_dev[pair_id] = nullptr;
  env->loaded[pair_id] = false;
  const size_t bytes = (size_t)T * c.n_cols * 8;
  FX_CUDA(env, cudaMalloc(&env->candles_dev[pair_id], bytes + 32));  // tail padding: TMA bulk copies are 16-B granular
  FX_CUDA(env, cudaMemset(reinterpret_cast<char*>(env->candles_dev[pair_id]) + bytes, 0, 32));
  FX_CUDA(env, cudaMemcpy(env->candles_dev[pair_id], candles_host, bytes, cudaMemcpyHostToDevice));
  if (minutes_host) {
    FX_CUDA(env, cudaMalloc(&env->minutes_dev[pair_id], (size_t)T * 8));
    FX_CUDA(env, cudaMemcpy(env->minutes_dev[pair_id], minutes_host, (size_t)T * 8, cudaMemcpyHostToDevice));
  }
  if (c.preproc == FX_PREPROC_FEATURE_WINDOW && c.scaling == FX_SCALING_ROLLING) {
    FX_CUDA(env, cudaMalloc(&env->stats_dev[pair_id], (size_t)T * c.n_features * 16));
    FX_CUDA(env, fx_launch_stats(c, env->candles_dev[pair_id], env->stats_dev[pair_id], T, 0));
    env->launches++;
    FX_CUDA(env, cudaDeviceSynchronize());
  }
  FxPairTable& tb = env->P.pair[pair_id];
  tb.candles = env->candles_dev[pair_id];
  tb.stats = env->stats_dev[pair_id];
  tb.minutes = env->minutes_dev[pair_id];
  tb.T = T;
  env->loaded[pair_id] = true;
  // finite and far from overflow => the observation code may skip its NaN fix-up (FxKernelParams::tame_data)
  bool tame = true;
  for (size_t i = 0, n = (size_t)T * c.n_cols; i < n && tame; i++) tame = fabs(candles_host[i]) < 1e100;  // false for NaN / inf
  env->tame[pair_id] = tame;
  env->P.tame_data = 1;
  for (int p = 0; p < c.num_pairs; p++) if (env->loaded[p] && !env->tame[p]) env->P.tame_data = 0;
  bool all_loaded = true;
  for (int p = 0; p < c.num_pairs; p++) all_loaded = all_loaded && env->loaded[p];
  env->P.lean = (all_loaded && !getenv("FXENV_NO_LEAN") && fx_config_is_lean(env->P)) ? 1 : 0;
  return FXENV_OK;
}

int64_t fxenv_obs_dim(const FxEnv* env) { return env ? env->P.obs_dim : -1; }

int fxenv_reset(FxEnv* env, const int64_t* start_bar_dev, const uint8_t* mask_dev, void* stream) {
  int rc = require_ready(env, false);
  if (rc) return rc;
  DeviceGuard g(env->device);
  FX_CUDA(env, fx_launch_reset(env->P, start_bar_dev, env->first_reset ? nullptr : mask_dev, env->first_reset ? 1 : 0,
                               (cudaStream_t)stream));
  env->launches++;
  env->first_reset = false;
  env->was_reset = true;
  return FXENV_OK;
}

int fxenv_observe(FxEnv* env, float* obs_dev, void* stream) {
  int rc = require_ready(env, true);
  if (rc) return rc;
  if (!obs_dev) return fail(env, FXENV_E_INVALID, "obs_dev is null");
  DeviceGuard g(env->device);
  FX_CUDA(env, fx_launch_observe(env->P, obs_dev, (cudaStream_t)stream));
  env->launches++;
  return FXENV_OK;
}

int fxenv_step(FxEnv* env, const void* actions_dev, float* obs_dev, float* reward_dev, uint8_t* terminated_dev,
               double* reward64_dev, void* stream) {
  int rc = require_ready(env, true);
  if (rc) return rc;
  if (!actions_dev || !obs_dev || !reward_dev || !terminated_dev) return fail(env, FXENV_E_INVALID, "null I/O pointer");
  DeviceGuard g(env->device);
  FX_CUDA(env, fx_launch_step(env->P, actions_dev, obs_dev, reward_dev, reward64_dev, terminated_dev, (cudaStream_t)stream));
  env->launches++;
  return FXENV_OK;
}

static bool batch_uses_rollout(const FxEnv* env, int n_steps) {
  // measured (B200, round 2, us/step persistent vs graph of single steps): cfg2 4096 envs 11.3 vs 24.7; cfg2 shape at 16384
  // envs 47.9 vs 50.7; cfg3 (16384 envs, 7.2 KB rows) 70.3 vs 70.7; cfg5 (8192 envs, 14 KB rows) 78.9 vs 83.8 -- the
  // persistent launch wins or ties everywhere, so every batch of more than one step uses it
  bool rollout = n_steps > 1;
  if (env->force_engine == 0) rollout = false;             // FXENV_ENGINE=graph / persistent: A/B of the engines
  if (env->force_engine == 1 && n_steps > 1) rollout = true;
  if (env->P.debug & (4 | 8)) rollout = false;             // FXENV_DEBUG: force the graph of single steps (A/B timing)
  if ((env->P.debug & 16) && n_steps > 1) rollout = true;  // FXENV_DEBUG & 16: force the persistent launch
  return rollout;
}

int fxenv_step_many_engine(const FxEnv* env, int n_steps) {
  if (!env) return FXENV_E_INVALID;
  return batch_uses_rollout(env, n_steps) ? 1 : 0;
}

int fxenv_step_many(FxEnv* env, int n_steps, const void* actions_dev, float* obs_dev, int obs_slots, float* reward_dev,
                    uint8_t* terminated_dev, void* stream_) {
  int rc = require_ready(env, true);
  if (rc) return rc;
  if (n_steps < 1 || obs_slots < 1) return fail(env, FXENV_E_INVALID, "n_steps and obs_slots must be >= 1");
  if (!actions_dev || !obs_dev || !reward_dev || !terminated_dev) return fail(env, FXENV_E_INVALID, "null I/O pointer");
  DeviceGuard g(env->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t N = (size_t)env->P.cfg.num_envs, D = (size_t)env->P.obs_dim;
  auto enqueue = [&](cudaStream_t s) -> cudaError_t {
    for (int k = 0; k < n_steps; k++) {
      const char* a = reinterpret_cast<const char*>(actions_dev) + (size_t)k * N * 4;
      cudaError_t e = fx_launch_step(env->P, a, obs_dev + (size_t)(k % obs_slots) * N * D, reward_dev + (size_t)k * N,
                                     nullptr, terminated_dev + (size_t)k * N, s);
      if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
  };
  // Two ways to run a batch: (a) ONE persistent launch whose warps pull (step, env) tickets and honour per-env
  // dependencies (fx_rollout_kernel) -- the tail of a step (envs with many fills) overlaps the next step; (b) a CUDA graph
  // of K single-step launches -- wins for large observation rows once the device is saturated (see batch_uses_rollout).
  const bool rollout = batch_uses_rollout(env, n_steps);
  if (rollout) {
    if ((unsigned long long)N * (unsigned long long)n_steps >= (1ull << 31))
      return fail(env, FXENV_E_INVALID, "num_envs * n_steps must be < 2^31 per fxenv_step_many call");
    // the per-env sequence words and the ticket counter are epoch-based: the host knows what they hold after every
    // launch, so no memset is needed between batches.  Inside a stream capture (the launch may be replayed any number
    // of times) that knowledge is lost: such launches, and every launch after one, zero the words first.
    cudaStreamCaptureStatus rcs = cudaStreamCaptureStatusNone;
    if (stream != nullptr) cudaStreamIsCapturing(stream, &rcs);
    if (rcs != cudaStreamCaptureStatusNone) env->seq_tracked = false;
    const bool tracked = env->seq_tracked;
    if (env->P.timeline && n_steps > env->timeline_steps) return fail(env, FXENV_E_INVALID, "FXENV_TIMELINE smaller than n_steps");
    const FxChunkPlan plan = fx_rollout_plan(env->P, n_steps);
    FX_CUDA(env, fx_launch_rollout(env->P, actions_dev, obs_dev, obs_slots, reward_dev, terminated_dev, plan,
                                   tracked ? env->seq_base : 0u, tracked ? env->ticket_base : 0u, !tracked, stream));
    if (tracked) {
      env->seq_base += (unsigned)n_steps;
      env->ticket_base += (unsigned)(N * (size_t)plan.n_rounds) + (unsigned)(fx_rollout_blocks(env->P) * FX_WARPS);
    }
    env->launches += 1;
    return FXENV_OK;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (stream != nullptr) cudaStreamIsCapturing(stream, &cs);
  if (cs != cudaStreamCaptureStatusNone || n_steps == 1 || stream == nullptr) {
    // already inside someone else's capture (e.g. a torch CUDA graph), or nothing to amortise: plain launches
    FX_CUDA(env, enqueue(stream));
    env->launches += n_steps;
    return FXENV_OK;
  }
  FxEnv::CachedGraph* slot = nullptr;
  for (auto& g : env->graphs)
    if (g.exec && g.actions == actions_dev && g.obs == obs_dev && g.reward == reward_dev && g.term == terminated_dev &&
        g.steps == n_steps && g.slots == obs_slots) slot = &g;
  if (!slot) {
    slot = (env->graphs[0].used <= env->graphs[1].used) ? &env->graphs[0] : &env->graphs[1];  // least recently used
    if (slot->exec) { cudaGraphExecDestroy(slot->exec); slot->exec = nullptr; }
    cudaGraph_t graph = nullptr;
    FX_CUDA(env, cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = enqueue(stream);
    cudaError_t e2 = cudaStreamEndCapture(stream, &graph);
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(env, e, "capture: fx_launch_step"); }
    if (e2 != cudaSuccess) return cuda_fail(env, e2, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&slot->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { slot->exec = nullptr; return cuda_fail(env, e, "cudaGraphInstantiate"); }
    slot->actions = actions_dev; slot->obs = obs_dev; slot->reward = reward_dev; slot->term = terminated_dev;
    slot->steps = n_steps; slot->slots = obs_slots;
  }
  slot->used = ++env->graph_clock;
  FX_CUDA(env, cudaGraphLaunch(slot->exec, stream));
  env->launches += n_steps;
  return FXENV_OK;
}

int fxenv_step_host(FxEnv* env, const void* actions_host, float* obs_host, float* reward_host, uint8_t* terminated_host) {
  int rc = require_ready(env, true);
  if (rc) return rc;
  if (!actions_host || !obs_host || !reward_host || !terminated_host) return fail(env, FXENV_E_INVALID, "null I/O pointer");
  DeviceGuard g(env->device);
  const size_t N = (size_t)env->P.cfg.num_envs, D = (size_t)env->P.obs_dim;
  if (!env->hstream) {
    FX_CUDA(env, cudaStreamCreateWithFlags(&env->hstream, cudaStreamNonBlocking));
    FX_CUDA(env, cudaStreamCreateWithFlags(&env->hcopy, cudaStreamNonBlocking));
    for (auto& ev : env->hev) FX_CUDA(env, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    FX_CUDA(env, cudaMalloc(&env->h_actions, N * 4));
    FX_CUDA(env, cudaMalloc(&env->h_obs, N * D * 4));
    FX_CUDA(env, cudaMalloc(&env->h_reward, N * 4));
    FX_CUDA(env, cudaMalloc(&env->h_term, N));
  }
  // The call is PCIe-bound (the observation rows: 3.6 KB per env).  The envs are stepped in up to 8 slices, slice i's
  // rows travelling to the host (copy engine, second stream) while slice i + 1 is being computed, so that only the
  // first slice's kernel time is exposed in front of the transfer.
  cudaStream_t s = env->hstream, sc = env->hcopy;
  int slices = (int)(N / 1024);
  if (slices > 2) slices = 2;  // measured (cfg2, 4096 envs): 1 slice 13.0 M env-steps/s, 8 slices 12.6 M (every extra DMA costs)
  if (const char* hs = getenv("FXENV_HOST_SLICES")) slices = atoi(hs);  // timing experiments only
  if (slices < 1) slices = 1;
  if (slices > FxEnv::kHostSlices) slices = FxEnv::kHostSlices;
  const size_t per = ((N + slices - 1) / slices + 31) & ~(size_t)31;
  FX_CUDA(env, cudaMemcpyAsync(env->h_actions, actions_host, N * 4, cudaMemcpyHostToDevice, s));
  for (int i = 0; i < slices; i++) {
    const size_t e0 = (size_t)i * per, e1 = (e0 + per < N) ? e0 + per : N;
    if (e0 >= e1) break;
    FX_CUDA(env, fx_launch_step(env->P, env->h_actions, env->h_obs, env->h_reward, nullptr, env->h_term, s, (int)e0, (int)e1));
    env->launches++;
    FX_CUDA(env, cudaEventRecord(env->hev[i], s));
    FX_CUDA(env, cudaStreamWaitEvent(sc, env->hev[i], 0));
    FX_CUDA(env, cudaMemcpyAsync(obs_host + e0 * D, env->h_obs + e0 * D, (e1 - e0) * D * 4, cudaMemcpyDeviceToHost, sc));
  }
  FX_CUDA(env, cudaMemcpyAsync(reward_host, env->h_reward, N * 4, cudaMemcpyDeviceToHost, sc));
  FX_CUDA(env, cudaMemcpyAsync(terminated_host, env->h_term, N, cudaMemcpyDeviceToHost, sc));
  FX_CUDA(env, cudaStreamSynchronize(sc));  // sc's last copy waited for the last slice's kernel: s is idle too
  return FXENV_OK;
}

int fxenv_get_info(FxEnv* env, FxInfoPtrs* out) {
  if (!env || !out) return FXENV_E_INVALID;
  const FxDeviceState& st = env->P.st;
  out->equity = st.equity; out->prev_equity = st.prev_equity; out->price = st.price; out->cash = st.cash;
  out->position_size = st.psize; out->position_price = st.pprice; out->commission_paid = st.commission_paid;
  out->position = st.position; out->bar_index = st.bar_index; out->total_bars = st.total_bars;
  out->trades = st.trades; out->n_orders = st.n_orders; out->flags = st.flags;
  out->run_stats = st.rstats;
  return FXENV_OK;
}

// A snapshot = header + the raw state slab.  The header pins what the slab's layout depends on, so that a blob taken
// from a different configuration / order capacity / candle table length / library version is refused instead of being
// reinterpreted.
struct FxStateHeader {
  uint32_t magic, abi, config_bytes, order_capacity;
  uint64_t slab_bytes, config_hash;
  int64_t table_rows[FXENV_MAX_PAIRS];
};

static FxStateHeader state_header(const FxEnv* env) {
  FxStateHeader h;
  memset(&h, 0, sizeof h);
  h.magic = 0x32535846u;  // "FXS2"
  h.abi = FXENV_ABI_VERSION;
  h.config_bytes = (uint32_t)sizeof(FxConfig);
  h.order_capacity = (uint32_t)env->P.cap;
  h.slab_bytes = env->slab_bytes;
  uint64_t x = 1469598103934665603ull;  // FNV-1a over the resolved config
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&env->P.cfg);
  for (size_t i = 0; i < sizeof(FxConfig); i++) { x ^= b[i]; x *= 1099511628211ull; }
  h.config_hash = x;
  for (int p = 0; p < FXENV_MAX_PAIRS; p++) h.table_rows[p] = env->P.pair[p].T;
  return h;
}

int64_t fxenv_state_bytes(const FxEnv* env) { return env ? (int64_t)(sizeof(FxStateHeader) + env->slab_bytes) : -1; }

int fxenv_get_state(FxEnv* env, void* buf_host, int64_t nbytes) {
  if (!env || !buf_host) return FXENV_E_INVALID;
  if (nbytes != fxenv_state_bytes(env)) return fail(env, FXENV_E_INVALID, "state buffer size mismatch");
  DeviceGuard g(env->device);
  FX_CUDA(env, cudaDeviceSynchronize());
  const FxStateHeader h = state_header(env);
  memcpy(buf_host, &h, sizeof h);
  FX_CUDA(env, cudaMemcpy(static_cast<char*>(buf_host) + sizeof h, env->slab, env->slab_bytes, cudaMemcpyDeviceToHost));
  return FXENV_OK;
}

int fxenv_set_state(FxEnv* env, const void* buf_host, int64_t nbytes) {
  if (!env || !buf_host) return FXENV_E_INVALID;
  if (nbytes != fxenv_state_bytes(env)) return fail(env, FXENV_E_INVALID, "state buffer size mismatch");
  FxStateHeader got;
  memcpy(&got, buf_host, sizeof got);
  const FxStateHeader want = state_header(env);
  if (got.magic != want.magic || got.abi != want.abi) return fail(env, FXENV_E_INVALID, "not a state blob of this library version");
  if (memcmp(&got, &want, sizeof got) != 0)
    return fail(env, FXENV_E_INVALID, "state blob was taken from a different configuration (config / order capacity / candle tables differ)");
  DeviceGuard g(env->device);
  FX_CUDA(env, cudaDeviceSynchronize());
  FX_CUDA(env, cudaMemcpy(env->slab, static_cast<const char*>(buf_host) + sizeof got, env->slab_bytes, cudaMemcpyHostToDevice));
  env->was_reset = true;
  env->first_reset = false;
  return FXENV_OK;
}

int64_t fxenv_launch_count(const FxEnv* env) { return env ? env->launches : -1; }

/* debug (FXENV_TIMELINE=K, timing build): copies the [K][num_envs][2] ticket start / end stamps; returns K or <0 */
int fxenv_debug_timeline(FxEnv* env, long long* out_host) {
  if (!env || !out_host || !env->P.timeline) return FXENV_E_STATE;
  DeviceGuard g(env->device);
  cudaDeviceSynchronize();
  if (cudaMemcpy(out_host, env->P.timeline, (size_t)env->timeline_steps * env->P.cfg.num_envs * 2 * sizeof(long long),
                 cudaMemcpyDeviceToHost) != cudaSuccess) return FXENV_E_CUDA;
  return env->timeline_steps;
}

/* debug / tests (pure host arithmetic, no CUDA call): how fxenv_step_many would cut n_steps into ticket rounds for
 * num_envs envs on a device holding resident_warps warps of the rollout kernel.  starts[0..rounds] receives the first
 * step of every round (starts[rounds] = n_steps); returns the number of rounds, or <0 if `cap` entries are not enough. */
int fxenv_debug_rollout_plan(int num_envs, int resident_warps, int n_steps, int* starts, int cap) {
  if (num_envs < 1 || resident_warps < 1 || n_steps < 1 || !starts) return FXENV_E_INVALID;
  FxKernelParams P = {};
  P.cfg.num_envs = num_envs;
  P.resident_blocks = (resident_warps + FX_WARPS - 1) / FX_WARPS;
  const FxChunkPlan pl = fx_rollout_plan(P, n_steps);
  if (pl.n_rounds + 1 > cap) return FXENV_E_INVALID;
  int r = 0;
  for (; r < pl.n_uniform; r++) starts[r] = r * pl.chunk;
  for (int t = 0; r <= pl.n_rounds; r++, t++) starts[r] = pl.tail_start[t];
  return pl.n_rounds;
}

/* debug (FXENV_TIMING=1): copies the [num_envs][FX_NSTAMP] phase stamps of the last step; returns FX_NSTAMP or <0 */
int fxenv_debug_timings(FxEnv* env, long long* out_host) {
  if (!env || !out_host || !env->P.timing) return FXENV_E_STATE;
  DeviceGuard g(env->device);
  cudaDeviceSynchronize();
  if (cudaMemcpy(out_host, env->P.timing, (size_t)env->P.cfg.num_envs * 2 * FX_NSTAMP * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess)
    return FXENV_E_CUDA;
  return FX_NSTAMP;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- closed loop (policy)
struct FxPolicy {
  FxEnv* env = nullptr;
  int k_pad = 0;                       // obs_dim padded to a multiple of 64 (bf16 row stride of the observation copy)
  uint16_t* w1 = nullptr;              // bf16 [256][k_pad]
  uint16_t* w2 = nullptr;              // bf16 [256][256]
  float* fparams = nullptr;            // b1[256] | b2[256] | head_w[4][256] | head_b[4]
  uint16_t* obs16[2] = {nullptr, nullptr};  // bf16 [num_envs][k_pad], double buffered
  uint16_t* h1 = nullptr;              // bf16 [num_envs padded to whole tiles][256]: the two halves of h1 meet here
  float* head_part = nullptr;          // float4 [num_envs padded]: partial head sums of the second CTA of a pair
  int32_t* sync = nullptr;             // act_flag[tiles] | done_cnt[tiles] | timeouts[1] (FxTileSync), zeroed per rollout
  int tiles = 0;
  int32_t* scratch_act = nullptr;      // bootstrap evaluation: action / logp are discarded
  float* scratch_logp = nullptr;
  static constexpr int kGroups = 4;     // env groups of a rollout (see enqueue_rollout)
  cudaStream_t side[kGroups - 1] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kGroups - 1] = {};
  CUtensorMap map_obs[2], map_w1, map_w2, map_h1;
  FxPolicyDev dev;
  bool has_weights = false;
  struct { cudaGraphExec_t exec = nullptr; FxRollout io; } cached;
};

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D bf16 [rows][cols] row-major tensor, box = 64 columns (128 bytes) x box_rows rows, 128-byte swizzle, OOB -> 0
int make_map(FxEnv* env, CUtensorMap* map, void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn)
      return fail(env, FXENV_E_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(env, FXENV_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return FXENV_OK;
}

// The envs are independent, so a rollout is run as up to 4 env GROUPS, each its own chain
//   policy(group, t) -> step(group, t) -> policy(group, t + 1) -> ...
// on its own stream (forked from / joined into the caller's stream; inside a capture this becomes parallel branches of
// the graph).  While one group's 128-row policy tiles occupy a few SMs, the other groups' env steps use the rest of the
// device, and a group of ~2048 envs steps in a single wave instead of the two of a 4096-env launch.  More, smaller groups
// do not pay: every group adds two kernel nodes per step to the graph, and a group's policy tiles need SMs of their own.
// (Also measured without gain: step kernels of 16 envs per CTA, so that a group's step leaves whole SMs free.)
cudaError_t enqueue_rollout(FxEnv* env, FxPolicy* pol, const FxRollout& io, cudaStream_t s) {
  const size_t N = (size_t)env->P.cfg.num_envs, D = (size_t)env->P.obs_dim;
  const int H = io.horizon, slots = io.obs_slots;
  static const int skip = [] { const char* v = getenv("FXENV_ROLLOUT_SKIP"); return v ? atoi(v) : 0; }();  // timing experiments:
                                                                          // 1 = no env steps, 2 = no policy evaluations
  // FXENV_TILE_SYNC=1: per-tile hand-over between the two kernels (FxTileSync) instead of whole-kernel dependencies.
  // Off by default: bit-identical results, but measured 27.9 vs 29.5 us/step at 1024 envs and 38.1 vs 37.5 at 4096 (two env
  // groups) -- the ~2 us per kernel boundary it removes is paid back by a thousand polling warps.
  const char* tsv = getenv("FXENV_TILE_SYNC");
  const bool tile_sync = !skip && tsv && atoi(tsv) != 0;
  cudaError_t e = cudaMemsetAsync(pol->sync, 0, (2 * (size_t)pol->tiles + 1) * sizeof(int32_t), s);
  if (e != cudaSuccess) return e;
  e = fx_launch_observe(env->P, io.obs, s, pol->obs16[0], pol->k_pad);  // the current observation, both copies
  if (e != cudaSuccess) return e;
  int groups = (int)(N / 2048);  // measured at 4096 envs (cfg4 shape), us/step: 41.2 with 1 group, 37.7 with 2, 39.5 with 4
  if (groups > FxPolicy::kGroups) groups = FxPolicy::kGroups;
  if (const char* ge = getenv("FXENV_ROLLOUT_GROUPS")) { const int g = atoi(ge); if (g >= 1 && g <= FxPolicy::kGroups) groups = g; }  // measurements
  if (groups < 1 || (env->P.debug & 64)) groups = 1;
  const size_t per = ((N + groups - 1) / groups + FX_POLICY_TILE_M - 1) / FX_POLICY_TILE_M * FX_POLICY_TILE_M;
  if (groups > 1) {
    e = cudaEventRecord(pol->ev_fork, s);
    if (e != cudaSuccess) return e;
  }
  for (int g = 0; g < groups; g++) {
    const size_t e0 = (size_t)g * per, e1 = (e0 + per < N) ? e0 + per : N;
    if (e0 >= e1) break;
    cudaStream_t sg = (g == 0) ? s : pol->side[g - 1];
    if (g > 0) {
      e = cudaStreamWaitEvent(sg, pol->ev_fork, 0);
      if (e != cudaSuccess) return e;
    }
    for (int t = 0; t <= H; t++) {
      const bool last = (t == H);  // the bootstrap evaluation: value only
      if (!(skip & 2))
      e = fx_launch_policy(pol->map_obs[t & 1], pol->map_w1, pol->map_w2, pol->map_h1, pol->dev, (int)N, pol->k_pad,
                           (!last && io.gumbel) ? io.gumbel + (size_t)t * N * 3 : nullptr, io.seed, (unsigned)t,
                           last ? pol->scratch_act : io.actions + (size_t)t * N, last ? pol->scratch_logp : io.logp + (size_t)t * N,
                           io.value + (size_t)t * N, sg, (int)e0, (int)e1, tile_sync);
      if (e != cudaSuccess) return e;
      if (last) break;
      const FxTileSync ts = {pol->dev.act_flag, pol->sync + pol->tiles, pol->dev.timeouts, t + 1};
      if (!(skip & 1))
      e = fx_launch_step(env->P, io.actions + (size_t)t * N, io.obs + (size_t)((t + 1) % slots) * N * D, io.reward + (size_t)t * N,
                         nullptr, io.done + (size_t)t * N, sg, (int)e0, (int)e1, pol->obs16[(t + 1) & 1], pol->k_pad,
                         tile_sync ? &ts : nullptr);
      if (e != cudaSuccess) return e;
    }
    if (g > 0) {
      e = cudaEventRecord(pol->ev_join[g - 1], sg);
      if (e != cudaSuccess) return e;
      e = cudaStreamWaitEvent(s, pol->ev_join[g - 1], 0);
      if (e != cudaSuccess) return e;
    }
  }
  return cudaSuccess;
}

}  // namespace

extern "C" {

int fxenv_policy_destroy(FxPolicy* pol) {
  if (!pol) return FXENV_OK;
  DeviceGuard g(pol->env->device);
  if (pol->cached.exec) cudaGraphExecDestroy(pol->cached.exec);
  cudaFree(pol->w1); cudaFree(pol->w2); cudaFree(pol->fparams); cudaFree(pol->obs16[0]); cudaFree(pol->obs16[1]);
  cudaFree(pol->h1); cudaFree(pol->head_part); cudaFree(pol->sync);
  cudaFree(pol->scratch_act); cudaFree(pol->scratch_logp);
  for (auto& st : pol->side) if (st) cudaStreamDestroy(st);
  if (pol->ev_fork) cudaEventDestroy(pol->ev_fork);
  for (auto& ev : pol->ev_join) if (ev) cudaEventDestroy(ev);
  delete pol;
  return FXENV_OK;
}

/* polls of the last fxenv_rollout's per-tile hand-over that gave up (0 unless something is broken); synchronises */
int fxenv_policy_sync_timeouts(FxPolicy* pol) {
  if (!pol) return FXENV_E_INVALID;
  DeviceGuard g(pol->env->device);
  int32_t v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(&v, pol->sync + 2 * pol->tiles, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return FXENV_E_CUDA;
  return (int)v;
}

int fxenv_policy_create(FxEnv* env, FxPolicy** out) {
  if (!env || !out) return FXENV_E_INVALID;
  *out = nullptr;
  if (env->P.cfg.action_mode != FX_ACTION_DISCRETE) return fail(env, FXENV_E_INVALID, "the fused policy samples discrete actions");
  DeviceGuard g(env->device);
  FxPolicy* pol = new (std::nothrow) FxPolicy();
  if (!pol) return fail(env, FXENV_E_NOMEM, "out of host memory");
  pol->env = env;
  const size_t N = (size_t)env->P.cfg.num_envs, D = (size_t)env->P.obs_dim;
  pol->k_pad = (int)((D + 63) / 64 * 64);
  const size_t KP = (size_t)pol->k_pad, Hd = FX_POLICY_HIDDEN;
  bool ok = cudaMalloc(&pol->w1, Hd * KP * 2) == cudaSuccess && cudaMalloc(&pol->w2, Hd * Hd * 2) == cudaSuccess &&
            cudaMalloc(&pol->fparams, (2 * Hd + 4 * Hd + 4) * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&pol->obs16[0], N * KP * 2) == cudaSuccess && cudaMalloc(&pol->obs16[1], N * KP * 2) == cudaSuccess &&
            cudaMalloc(&pol->scratch_act, N * 4) == cudaSuccess && cudaMalloc(&pol->scratch_logp, N * 4) == cudaSuccess;
  const size_t NP = (N + FX_POLICY_TILE_M - 1) / FX_POLICY_TILE_M * FX_POLICY_TILE_M;  // whole 128-env tiles
  ok = ok && cudaMalloc(&pol->h1, NP * Hd * 2) == cudaSuccess && cudaMalloc(&pol->head_part, NP * 4 * sizeof(float)) == cudaSuccess;
  pol->tiles = (int)(NP / FX_POLICY_TILE_M);
  ok = ok && cudaMalloc(&pol->sync, (2 * (size_t)pol->tiles + 1) * sizeof(int32_t)) == cudaSuccess;
  if (!ok) { fxenv_policy_destroy(pol); return fail(env, FXENV_E_CUDA, "cudaMalloc(policy buffers) failed"); }
  // the K padding of the observation copies is never written by the env kernels: zero it once
  cudaMemset(pol->obs16[0], 0, N * KP * 2);
  cudaMemset(pol->obs16[1], 0, N * KP * 2);
  pol->dev.b1 = pol->fparams; pol->dev.b2 = pol->fparams + Hd; pol->dev.head_w = pol->fparams + 2 * Hd;
  pol->dev.head_b = pol->fparams + 6 * Hd;
  pol->dev.h1 = pol->h1; pol->dev.head_part = reinterpret_cast<float4*>(pol->head_part);
  pol->dev.dbg = env->P.timeline;  // (nullptr unless FXENV_TIMELINE is set; only the timing build looks at it)
  pol->dev.act_flag = pol->sync; pol->dev.done_cnt = pol->sync + pol->tiles; pol->dev.timeouts = pol->sync + 2 * pol->tiles;
  cudaMemset(pol->sync, 0, (2 * (size_t)pol->tiles + 1) * sizeof(int32_t));
  int rc = make_map(env, &pol->map_obs[0], pol->obs16[0], N, KP, FX_POLICY_TILE_M);
  if (!rc) rc = make_map(env, &pol->map_obs[1], pol->obs16[1], N, KP, FX_POLICY_TILE_M);
  if (!rc) rc = make_map(env, &pol->map_w1, pol->w1, Hd, KP, FX_POLICY_HIDDEN / 2);   // a CTA loads its half of the units
  if (!rc) rc = make_map(env, &pol->map_w2, pol->w2, Hd, Hd, FX_POLICY_HIDDEN / 2);
  if (!rc) rc = make_map(env, &pol->map_h1, pol->h1, NP, Hd, FX_POLICY_TILE_M);
  if (rc) { fxenv_policy_destroy(pol); return rc; }
  cudaError_t ce = fx_policy_configure();
  for (auto& st : pol->side) if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&pol->ev_fork, cudaEventDisableTiming);
  for (auto& ev : pol->ev_join) if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (ce != cudaSuccess) { fxenv_policy_destroy(pol); return cuda_fail(env, ce, "fx_policy_configure / streams"); }
  *out = pol;
  return FXENV_OK;
}

int fxenv_policy_set_weights(FxPolicy* pol, const FxPolicyWeights* w, void* stream_) {
  if (!pol || !w) return FXENV_E_INVALID;
  FxEnv* env = pol->env;
  if (!w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->w_pi || !w->b_pi || !w->w_v || !w->b_v) return fail(env, FXENV_E_INVALID, "null weight pointer");
  DeviceGuard g(env->device);
  cudaStream_t s = (cudaStream_t)stream_;
  const int D = env->P.obs_dim, Hd = FX_POLICY_HIDDEN;
  FX_CUDA(env, fx_policy_pack(w->w1, pol->w1, Hd, D, pol->k_pad, s));
  FX_CUDA(env, fx_policy_pack(w->w2, pol->w2, Hd, Hd, Hd, s));
  float* f = pol->fparams;
  FX_CUDA(env, cudaMemcpyAsync(f, w->b1, Hd * 4, cudaMemcpyDeviceToDevice, s));
  FX_CUDA(env, cudaMemcpyAsync(f + Hd, w->b2, Hd * 4, cudaMemcpyDeviceToDevice, s));
  FX_CUDA(env, cudaMemcpyAsync(f + 2 * Hd, w->w_pi, 3 * Hd * 4, cudaMemcpyDeviceToDevice, s));
  FX_CUDA(env, cudaMemcpyAsync(f + 5 * Hd, w->w_v, Hd * 4, cudaMemcpyDeviceToDevice, s));
  FX_CUDA(env, cudaMemcpyAsync(f + 6 * Hd, w->b_pi, 3 * 4, cudaMemcpyDeviceToDevice, s));
  FX_CUDA(env, cudaMemcpyAsync(f + 6 * Hd + 3, w->b_v, 4, cudaMemcpyDeviceToDevice, s));
  env->launches += 2;
  pol->has_weights = true;
  return FXENV_OK;
}

int fxenv_rollout(FxEnv* env, FxPolicy* pol, const FxRollout* io, void* stream_) {
  int rc = require_ready(env, true);
  if (rc) return rc;
  if (!pol || !io || pol->env != env) return fail(env, FXENV_E_INVALID, "policy does not belong to this env");
  if (!pol->has_weights) return fail(env, FXENV_E_STATE, "fxenv_policy_set_weights has not been called");
  if (io->horizon < 1 || io->obs_slots < 2) return fail(env, FXENV_E_INVALID, "horizon must be >= 1 and obs_slots >= 2");
  if (!io->obs || !io->actions || !io->logp || !io->value || !io->reward || !io->done) return fail(env, FXENV_E_INVALID, "null I/O pointer");
  DeviceGuard g(env->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t nl = 2 * (int64_t)io->horizon + 2;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (stream != nullptr) cudaStreamIsCapturing(stream, &cs);
  if (cs != cudaStreamCaptureStatusNone || stream == nullptr || (env->P.debug & 32)) {  // inside a capture / legacy stream: plain launches
    FX_CUDA(env, enqueue_rollout(env, pol, *io, stream));
    env->launches += nl;
    return FXENV_OK;
  }
  if (!pol->cached.exec || memcmp(&pol->cached.io, io, sizeof(FxRollout)) != 0) {
    if (pol->cached.exec) { cudaGraphExecDestroy(pol->cached.exec); pol->cached.exec = nullptr; }
    cudaGraph_t graph = nullptr;
    FX_CUDA(env, cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    cudaError_t e = enqueue_rollout(env, pol, *io, stream);
    cudaError_t e2 = cudaStreamEndCapture(stream, &graph);
    if (e != cudaSuccess) { if (graph) cudaGraphDestroy(graph); return cuda_fail(env, e, "capture: rollout"); }
    if (e2 != cudaSuccess) return cuda_fail(env, e2, "cudaStreamEndCapture");
    e = cudaGraphInstantiate(&pol->cached.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { pol->cached.exec = nullptr; return cuda_fail(env, e, "cudaGraphInstantiate"); }
    pol->cached.io = *io;
  }
  FX_CUDA(env, cudaGraphLaunch(pol->cached.exec, stream));
  env->launches += nl;
  return FXENV_OK;
}

}  // extern "C"
