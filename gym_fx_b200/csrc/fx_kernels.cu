// fx_kernels.cu -- sm_100a kernels of the fused gym-fx env.step().
//
// fx_step_kernel: ONE launch = one env.step() for all N envs.  One warp owns one env:
//   1. scalar state + the env's order table are pulled into registers / per-warp shared memory (coalesced);
//   2. backtrader's per-bar broker pass: bracket activation and the trigger test of every live order run
//      lane-parallel (one order per lane, ballot -> hit mask); the few orders that do trade are then executed
//      in FIFO order by uniform scalar fp64 code (fx_core.cuh), exactly like BackBroker.next();
//   3. the strategy plugin's apply_action appends new orders; bridge publish; reward (pnl / sharpe / dd);
//   4. the observation row ([W,F] z-scored features | prices | returns | 4 agent scalars) is streamed by all 32
//      lanes: coalesced fp64 reads of the L2-resident candle table, fp64 math, coalesced fp32 stores.
// No tensor cores: there is no contraction on this path; it is bound by HBM stores of the observation rows.
//
// Reference call stack being replaced: app/env.py:131-172 -> app/bt_bridge.py:119-150 -> strategy/reward/
// preprocessor plugins + backtrader (see fx_core.cuh for the per-function citations).
#include <cuda_runtime.h>

#include "fx_kernels.cuh"

#define FX_FULL 0xffffffffu

namespace {

struct WarpSmem {
  double *p0, *p1, *sz, *mean, *rcp, *ring;
  uint32_t *meta, *hit;
};

__device__ __forceinline__ WarpSmem fx_carve(unsigned char* base, int cap, int ring_len) {
  WarpSmem w;
  double* d = reinterpret_cast<double*>(base);
  w.p0 = d; d += cap;
  w.p1 = d; d += cap;
  w.sz = d; d += cap;
  w.mean = d; d += FXENV_MAX_FEATURES;
  w.rcp = d; d += FXENV_MAX_FEATURES;
  w.ring = d; d += ring_len;
  uint32_t* u = reinterpret_cast<uint32_t*>(d);
  w.meta = u; u += cap;
  w.hit = u;
  return w;
}

__device__ __forceinline__ double fx_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FX_FULL, v, o);
  return v;
}

__device__ __forceinline__ void fx_load_regs(const FxDeviceState& st, int env, FxEnvRegs& e) {
  e.cash = st.cash[env]; e.psize = st.psize[env]; e.pprice = st.pprice[env]; e.value = st.value[env];
  e.equity = st.equity[env]; e.prev_equity = st.prev_equity[env]; e.price = st.price[env];
  e.commission_paid = st.commission_paid[env];
  e.position = st.position[env]; e.bar_index = st.bar_index[env]; e.trades = st.trades[env];
  e.flags = st.flags[env];
}

__device__ __forceinline__ void fx_store_regs(const FxDeviceState& st, int env, const FxEnvRegs& e) {
  st.cash[env] = e.cash; st.psize[env] = e.psize; st.pprice[env] = e.pprice; st.value[env] = e.value;
  st.equity[env] = e.equity; st.prev_equity[env] = e.prev_equity; st.price[env] = e.price;
  st.commission_paid[env] = e.commission_paid;
  st.position[env] = e.position; st.bar_index[env] = e.bar_index; st.trades[env] = e.trades;
  st.flags[env] = e.flags;
}

// GymFxEnv.reset (app/env.py:102-129): fresh bridge/broker/strategy; broker.next() on bar 0 with nothing pending
// (value = cash) and the first _publish_obs.  Reward-plugin state persists across episodes, like the plugin
// instance does in the reference (its `step <= last_step` rule then clears it on the next compute_reward).
__device__ __forceinline__ void fx_reset_regs(const FxConfig& c, FxEnvRegs& e, double close0) {
  e.cash = c.initial_cash; e.value = c.initial_cash; e.psize = 0.0; e.pprice = 0.0;
  e.equity = c.initial_cash; e.prev_equity = c.initial_cash; e.commission_paid = 0.0;
  e.trades = 0; e.position = 0; e.flags = 0u;
  e.price = close0; e.bar_index = 1;
}

__device__ __forceinline__ int32_t fx_total_bars(const FxConfig& c, int64_t T, int64_t start) {
  int64_t tb = T - start;
  if (c.episode_bars > 0 && c.episode_bars < tb) tb = c.episode_bars;
  return (int32_t)tb;
}

// ---- observation row: preprocessor.make_observation in the flat VecEnv layout -------------------------------------
__device__ __forceinline__ void fx_write_obs(const FxKernelParams& P, const FxPairTable& tb, int lane, const WarpSmem& ws,
                                             const FxEnvRegs& e, int32_t total_bars, int64_t start,
                                             float* __restrict__ out) {
  const FxConfig& c = P.cfg;
  const int W = c.window_size, C = c.n_cols;
  int s = e.bar_index;
  if (s < 0) s = 0;
  if (s > total_bars) s = total_bars;  // app/env.py:228
  int left = s - W;
  if (left < 0) left = 0;
  const int have = s - left;
  const int pad = W - have;  // left padding with the first available row
  const double* __restrict__ base = tb.candles + start * (int64_t)C;
  int off = 0;
  if (c.preproc == FX_PREPROC_FEATURE_WINDOW) {
    const int F = c.n_features;
    int hl = 0, hn = 0;
    if (c.scaling == FX_SCALING_ROLLING) { hl = s - c.scaling_window; if (hl < 0) hl = 0; hn = s - hl; }
    else if (c.scaling == FX_SCALING_EXPANDING) { hl = 0; hn = s; }
    const bool scale = (c.scaling != FX_SCALING_NONE) && hn >= 2;
    if (scale) {
      if (c.scaling == FX_SCALING_ROLLING && hn == c.scaling_window && tb.stats != nullptr) {
        // full rolling window: per-bar statistics precomputed at load time (pure function of the bar)
        if (lane < F) {
          const double* sp = tb.stats + ((start + s - 1) * (int64_t)F + lane) * 2;
          ws.mean[lane] = sp[0];
          ws.rcp[lane] = sp[1];
        }
      } else {
        // warm-up (history shorter than the scaling window) or expanding z-score: two-pass mean / population std
        for (int f = 0; f < F; f++) {
          const int col = c.feature_cols[f];
          double acc = 0.0;
          for (int k = lane; k < hn; k += 32) acc += base[(int64_t)(hl + k) * C + col];
          const double m = fx_warp_sum(acc) / (double)hn;
          double a2 = 0.0;
          for (int k = lane; k < hn; k += 32) { const double d = base[(int64_t)(hl + k) * C + col] - m; a2 += d * d; }
          double sd = sqrt(fx_warp_sum(a2) / (double)hn);
          if (sd < 1e-8) sd = 1.0;
          if (lane == 0) { ws.mean[f] = m; ws.rcp[f] = 1.0 / sd; }
        }
      }
      __syncwarp();
    }
    const float clipf = (float)c.feature_clip;
    const bool do_clip = c.feature_clip > 0.0;
    const int total = W * F;
    if (P.fast_features && pad == 0) {
      // feature columns == all table columns and no padding: the [W][F] block is one contiguous span of the table
      const double* __restrict__ src = base + (int64_t)left * C;
      for (int j = lane; j < total; j += 32) {
        const int f = j % F;
        const double x = __ldg(src + j);
        const float v = (scale && !c.feature_binary[f]) ? (float)((x - ws.mean[f]) * ws.rcp[f]) : (float)x;
        __stcs(out + j, fx_clip_nan(v, clipf, do_clip));
      }
    } else {
      for (int j = lane; j < total; j += 32) {
        const int w = j / F, f = j - w * F;
        int k = w - pad;
        if (k < 0) k = 0;
        const double x = __ldg(base + (int64_t)(left + k) * C + c.feature_cols[f]);
        const float v = (scale && !c.feature_binary[f]) ? (float)((x - ws.mean[f]) * ws.rcp[f]) : (float)x;
        __stcs(out + j, fx_clip_nan(v, clipf, do_clip));
      }
    }
    off = total;
  }
  const bool inc_price = (c.preproc == FX_PREPROC_DEFAULT) || c.include_price_window;
  const bool inc_agent = (c.preproc == FX_PREPROC_DEFAULT) || c.include_agent_state;
  const int pc = c.price_col;
  if (inc_price) {
    for (int w = lane; w < W; w += 32) {
      int k = w - pad;
      if (k < 0) k = 0;
      const double p = __ldg(base + (int64_t)(left + k) * C + pc);
      double prev = p;
      if (w > 0) {
        int k1 = w - 1 - pad;
        if (k1 < 0) k1 = 0;
        prev = __ldg(base + (int64_t)(left + k1) * C + pc);
      }
      __stcs(out + off + w, (float)p);
      __stcs(out + off + W + w, (w == 0) ? 0.0f : (float)(p - prev));
    }
    off += 2 * W;
  }
  if (inc_agent && lane == 0) {
    const double last = base[(int64_t)(left + have - 1) * C + pc];
    double ref;
    if (c.preproc == FX_PREPROC_DEFAULT) ref = last;                       // default_preprocessor.py:63
    else ref = inc_price ? (double)(float)last : e.price;                  // feature_window_preprocessor.py:218-222
    float sc[4];
    fx_agent_scalars(c, e, total_bars, ref, sc);
    out[off + 0] = sc[0]; out[off + 1] = sc[1]; out[off + 2] = sc[2]; out[off + 3] = sc[3];
  }
}

// ---- the fused step ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FX_WARPS_PER_BLOCK * 32)
fx_step_kernel(const __grid_constant__ FxKernelParams P, const void* __restrict__ actions, float* __restrict__ obs,
               float* __restrict__ reward, double* __restrict__ reward64, uint8_t* __restrict__ terminated) {
  extern __shared__ __align__(16) unsigned char fx_smem[];
  const FxConfig& c = P.cfg;
  const FxDeviceState& st = P.st;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * FX_WARPS_PER_BLOCK + warp;
  if (env >= c.num_envs) return;
  const int cap = P.cap;
  const int ring_len = (c.reward == FX_REWARD_SHARPE) ? c.sharpe_window : 0;
  const WarpSmem ws = fx_carve(fx_smem + (size_t)warp * P.smem_per_warp, cap, ring_len);
  const int pair = env % c.num_pairs;
  const FxPairTable& tb = P.pair[pair];
  const int C = c.n_cols;
  float* __restrict__ obs_row = obs + (int64_t)env * P.obs_dim;

  FxEnvRegs e;
  fx_load_regs(st, env, e);
  int32_t t = st.t[env];
  int32_t total_bars = st.total_bars[env];
  int64_t start = st.start[env];
  int n = st.n_orders[env];

  int action;
  if (c.action_mode == FX_ACTION_CONTINUOUS) action = fx_coerce_continuous(c, reinterpret_cast<const float*>(actions)[env]);
  else action = fx_coerce_discrete(reinterpret_cast<const int32_t*>(actions)[env]);

  // --- already terminated: the reference answers (obs, 0.0, True) without touching plugins (app/env.py:137-138)
  if (e.flags & FX_FLAG_TERMINATED) {
    if (c.auto_reset) {
      // build-side extension: next-step auto reset (the env restarts its episode window)
      total_bars = fx_total_bars(c, tb.T, start);
      fx_reset_regs(c, e, tb.candles[start * (int64_t)C + 3]);
      if (lane == 0) {
        fx_store_regs(st, env, e);
        st.t[env] = 0; st.total_bars[env] = total_bars; st.n_orders[env] = 0;
        reward[env] = 0.0f; terminated[env] = 0;
        if (reward64) reward64[env] = 0.0;
      }
      fx_write_obs(P, tb, lane, ws, e, total_bars, start, obs_row);
      return;
    }
    if (lane == 0) {
      reward[env] = 0.0f; terminated[env] = 1;
      if (reward64) reward64[env] = 0.0;
    }
    fx_write_obs(P, tb, lane, ws, e, total_bars, start, obs_row);
    return;
  }

  // --- step <-> bar timeline: the first step does not advance; later steps advance or exhaust
  bool exhausted = false, advance = false;
  if (!(e.flags & FX_FLAG_STARTED)) e.flags |= FX_FLAG_STARTED;
  else if (t + 1 >= total_bars) exhausted = true;  // strategy.stop(): bridge state unchanged (app/bt_bridge.py:152-155)
  else { t += 1; advance = true; }

  FxBar b;
  {
    const double* r = tb.candles + (start + t) * (int64_t)C;
    b.o = r[0]; b.h = r[1]; b.l = r[2]; b.c = r[3];
  }
  FxOrderTab tab;
  tab.meta = ws.meta; tab.p0 = ws.p0; tab.p1 = ws.p1; tab.sz = ws.sz;
  tab.n = n; tab.cap = cap; tab.dirty_from = n;
  const int64_t obase = (int64_t)env * cap;

  if (advance && n > 0) {
    // ---- BackBroker.next(): stage the order table; lane-parallel activation + trigger test
    int first_sub = n, first_changed = n;
    const int nch = (n + 31) >> 5;
    for (int ch = 0; ch < nch; ch++) {
      const int k = ch * 32 + lane;
      bool hit = false, sub = false, changed = false;
      if (k < n) {
        const uint32_t m0 = st.o_meta[obase + k];
        const uint32_t m = fx_entry_begin_bar(m0);
        const double p0 = st.o_p0[obase + k], p1 = st.o_p1[obase + k], sz = st.o_sz[obase + k];
        ws.meta[k] = m; ws.p0[k] = p0; ws.p1[k] = p1; ws.sz[k] = sz;
        changed = (m != m0);
        sub = (m & FXO_SUBMITTED) != 0u;
        hit = fx_entry_hits(m, p0, p1, sz, b);
      }
      const uint32_t hm = __ballot_sync(FX_FULL, hit);
      const uint32_t sm = __ballot_sync(FX_FULL, sub);
      const uint32_t cm = __ballot_sync(FX_FULL, changed);
      if (lane == 0) ws.hit[ch] = hm;
      if (sm && first_sub == n) first_sub = ch * 32 + __ffs(sm) - 1;
      if (cm && first_changed == n) first_changed = ch * 32 + __ffs(cm) - 1;
    }
    __syncwarp();
    tab.dirty_from = first_changed;
    fx_check_submitted(c, e, tab, first_sub);
    __syncwarp();
    // ---- FIFO walk over the entries that trade on this bar (uniform scalar code)
    for (int ch = 0; ch < nch; ch++) {
      uint32_t m = ws.hit[ch];
      while (m) {
        const int k = ch * 32 + __ffs(m) - 1;
        m &= m - 1;
        fx_exec_entry(c, e, tab, k, b);
      }
    }
    __syncwarp();
    // ---- stable compaction of finished entries (keeps FIFO order == array order)
    if (tab.dirty_from < n) {
      int w = 0;
      for (int ch = 0; ch < nch; ch++) {
        const int k = ch * 32 + lane;
        uint32_t m = 0u; double p0 = 0.0, p1 = 0.0, sz = 0.0;
        bool keep = false;
        if (k < n) { m = ws.meta[k]; p0 = ws.p0[k]; p1 = ws.p1[k]; sz = ws.sz[k]; keep = !(m & FXO_DEAD); }
        const uint32_t km = __ballot_sync(FX_FULL, keep);
        __syncwarp();
        if (keep) {
          const int dst = w + __popc(km & ((1u << lane) - 1u));
          ws.meta[dst] = m; ws.p0[dst] = p0; ws.p1[dst] = p1; ws.sz[dst] = sz;
        }
        w += __popc(km);
        __syncwarp();
      }
      tab.n = w;
    }
  }
  if (advance) fx_mark_to_market(c, e, b.c);

  if (!exhausted) {
    // ---- strategy plugin (BTBridgeStrategy._apply_action) at bar t
    double atr = 0.0;
    bool atr_ready = false;
    if (c.strategy == FX_STRATEGY_ATR_SLTP && action != 0) {
      // simple-mean ATR over the env's TR deque; TR(k) is a pure function of the table, so the
      // deque is rebuilt from the last min(t+1, period) bars in deque order with Python's compensated sum()
      const int period = c.atr_period;
      const int nb = (t + 1 < period) ? t + 1 : period;
      double s_ = 0.0, comp = 0.0;
      for (int j = 0; j < nb; j++) {
        const int k = t - nb + 1 + j;
        const double* r = tb.candles + (start + k) * (int64_t)C;
        const double prevc = (k > 0) ? r[3 - C] : 0.0;
        const double tr = fx_true_range(r[1], r[2], prevc, k > 0);
        if (j == 0) s_ = tr; else fx_neumaier_add(s_, comp, tr);
      }
      atr = fx_neumaier_done(s_, comp) / (double)nb;
      atr_ready = nb >= period;
    }
    const bool has_min = (tb.minutes != nullptr);
    const int64_t minutes = (c.session_filter && has_min) ? tb.minutes[start + t] : 0;
    fx_apply_action(c, e, tab, action, b, pair, atr, atr_ready, has_min, minutes);
    fx_publish(e, b.c, t);
    if (e.equity <= c.min_equity) e.flags |= FX_FLAG_TERMINATED | FX_FLAG_BROKE;  // app/bt_bridge.py:140-143
  } else {
    e.flags |= FX_FLAG_TERMINATED | FX_FLAG_EXHAUSTED;
  }
  __syncwarp();

  // ---- reward plugin (app/env.py:148-155)
  double r;
  if (c.reward == FX_REWARD_PNL) {
    r = fx_reward_pnl(c, e);
  } else if (c.reward == FX_REWARD_DD) {
    double peak = st.dd_peak[env];
    int32_t last = st.dd_last_step[env];
    r = fx_reward_dd(c, e, peak, last);
    if (lane == 0) { st.dd_peak[env] = peak; st.dd_last_step[env] = last; }
  } else {
    const int Wn = c.sharpe_window;
    double* gring = st.sh_ring + (int64_t)env * Wn;
    int32_t len = st.sh_len[env], head = st.sh_head[env], last = st.sh_last_step[env];
    for (int k = lane; k < Wn; k += 32) ws.ring[k] = gring[k];
    __syncwarp();
    const double ret = (e.equity - e.prev_equity) / c.reward_initial_cash;
    const int32_t len0 = len, head0 = head;
    int slot;  // where the new return lands (same rule as fx_sharpe_push)
    if (e.bar_index <= last) slot = 0; else slot = (len0 == Wn) ? head0 : (head0 + len0) % Wn;
    const int nn = fx_sharpe_push(ws.ring, Wn, len, head, last, e.bar_index, ret);
    r = fx_sharpe_eval(ws.ring, Wn, nn, head, c.annualization_factor);
    if (lane == 0) {
      gring[slot] = ret;
      st.sh_len[env] = len; st.sh_head[env] = head; st.sh_last_step[env] = last;
    }
  }
  const bool term = ((e.flags & FX_FLAG_TERMINATED) != 0u) || (e.equity <= c.min_equity);  // app/env.py:157

  // ---- write back: scalars, outputs, the stale tail of the order table
  if (lane == 0) {
    fx_store_regs(st, env, e);
    st.t[env] = t;
    st.n_orders[env] = tab.n;
    reward[env] = (float)r;
    if (reward64) reward64[env] = r;
    terminated[env] = term ? 1 : 0;
  }
  for (int k = tab.dirty_from + lane; k < tab.n; k += 32) {
    st.o_meta[obase + k] = ws.meta[k]; st.o_p0[obase + k] = ws.p0[k];
    st.o_p1[obase + k] = ws.p1[k]; st.o_sz[obase + k] = ws.sz[k];
  }

  // ---- observation (app/env.py:160 -> preprocessor.make_observation)
  fx_write_obs(P, tb, lane, ws, e, total_bars, start, obs_row);
}

__global__ void fx_reset_kernel(const __grid_constant__ FxKernelParams P, const int64_t* __restrict__ start_bar,
                                const uint8_t* __restrict__ mask, int first) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  const FxConfig& c = P.cfg;
  if (env >= c.num_envs) return;
  const FxDeviceState& st = P.st;
  if (first) {  // plugin instances are brand new
    st.sh_len[env] = 0; st.sh_head[env] = 0; st.sh_last_step[env] = -1; st.dd_last_step[env] = -1;
    st.dd_peak[env] = 0.0;
  }
  if (mask && !mask[env]) return;
  const FxPairTable& tb = P.pair[env % c.num_pairs];
  int64_t start = start_bar ? start_bar[env] : st.start[env];
  if (start < 0) start = 0;
  if (start > tb.T - 1) start = tb.T - 1;
  FxEnvRegs e;
  fx_reset_regs(c, e, tb.candles[start * (int64_t)c.n_cols + 3]);
  fx_store_regs(st, env, e);
  st.start[env] = start;
  st.t[env] = 0;
  st.total_bars[env] = fx_total_bars(c, tb.T, start);
  st.n_orders[env] = 0;
}

__global__ void __launch_bounds__(FX_WARPS_PER_BLOCK * 32)
fx_observe_kernel(const __grid_constant__ FxKernelParams P, float* __restrict__ obs) {
  extern __shared__ __align__(16) unsigned char fx_smem[];
  const FxConfig& c = P.cfg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = blockIdx.x * FX_WARPS_PER_BLOCK + warp;
  if (env >= c.num_envs) return;
  const int ring_len = (c.reward == FX_REWARD_SHARPE) ? c.sharpe_window : 0;
  const WarpSmem ws = fx_carve(fx_smem + (size_t)warp * P.smem_per_warp, P.cap, ring_len);
  FxEnvRegs e;
  fx_load_regs(P.st, env, e);
  fx_write_obs(P, P.pair[env % c.num_pairs], lane, ws, e, P.st.total_bars[env], P.st.start[env],
               obs + (int64_t)env * P.obs_dim);
}

// Per-bar rolling z-score statistics (feature_window_preprocessor._scale_window :96-124 for a FULL window):
// stats[g][f] = {mean, 1/std} over rows (g-S, g], population std, std < 1e-8 -> 1.  One thread per (bar, feature).
__global__ void fx_stats_kernel(FxConfig c, const double* __restrict__ candles, double* __restrict__ stats, int64_t T) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int F = c.n_features, S = c.scaling_window, C = c.n_cols;
  if (idx >= T * F) return;
  const int64_t g = idx / F;
  const int f = (int)(idx - g * F);
  double m = 0.0, rc = 1.0;
  if (g + 1 >= S) {
    const double* p = candles + (g + 1 - S) * C + c.feature_cols[f];
    double acc = 0.0;
    for (int k = 0; k < S; k++) acc += p[(int64_t)k * C];
    m = acc / (double)S;
    double a2 = 0.0;
    for (int k = 0; k < S; k++) { const double d = p[(int64_t)k * C] - m; a2 += d * d; }
    double sd = sqrt(a2 / (double)S);
    if (sd < 1e-8) sd = 1.0;
    rc = 1.0 / sd;
  }
  stats[idx * 2 + 0] = m;
  stats[idx * 2 + 1] = rc;
}

}  // namespace

size_t fx_smem_per_warp(const FxConfig& cfg, int cap) {
  const int ring_len = (cfg.reward == FX_REWARD_SHARPE) ? cfg.sharpe_window : 0;
  size_t b = (size_t)cap * 3 * 8 + 2 * FXENV_MAX_FEATURES * 8 + (size_t)ring_len * 8 + (size_t)cap * 4 + 8 * 4;
  return (b + 15) & ~(size_t)15;
}

// dynamic shared memory above the 48 KB default needs an explicit opt-in per kernel
cudaError_t fx_configure_kernels(size_t smem_per_block) {
  if (smem_per_block <= 48 * 1024) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fx_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_per_block);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fx_observe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_per_block);
}

cudaError_t fx_launch_step(const FxKernelParams& P, const void* actions, float* obs, float* reward, double* reward64,
                           uint8_t* terminated, cudaStream_t stream) {
  const int N = P.cfg.num_envs;
  const int blocks = (N + FX_WARPS_PER_BLOCK - 1) / FX_WARPS_PER_BLOCK;
  const size_t smem = (size_t)P.smem_per_warp * FX_WARPS_PER_BLOCK;
  fx_step_kernel<<<blocks, FX_WARPS_PER_BLOCK * 32, smem, stream>>>(P, actions, obs, reward, reward64, terminated);
  return cudaGetLastError();
}

cudaError_t fx_launch_reset(const FxKernelParams& P, const int64_t* start_bar, const uint8_t* mask, int first,
                            cudaStream_t stream) {
  const int N = P.cfg.num_envs;
  fx_reset_kernel<<<(N + 127) / 128, 128, 0, stream>>>(P, start_bar, mask, first);
  return cudaGetLastError();
}

cudaError_t fx_launch_observe(const FxKernelParams& P, float* obs, cudaStream_t stream) {
  const int N = P.cfg.num_envs;
  const int blocks = (N + FX_WARPS_PER_BLOCK - 1) / FX_WARPS_PER_BLOCK;
  const size_t smem = (size_t)P.smem_per_warp * FX_WARPS_PER_BLOCK;
  fx_observe_kernel<<<blocks, FX_WARPS_PER_BLOCK * 32, smem, stream>>>(P, obs);
  return cudaGetLastError();
}

cudaError_t fx_launch_stats(const FxConfig& cfg, const double* candles, double* stats, int64_t T, cudaStream_t stream) {
  const int64_t total = T * cfg.n_features;
  const int threads = 128;
  fx_stats_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(cfg, candles, stats, T);
  return cudaGetLastError();
}
