/*
 * include/fxenv.h -- C-ABI of libfxenv.so: the B200-native vectorised gym-fx env.step() hot path.
 *
 * The reference (harveybc/gym-fx) is pure Python and has no FFI of its own; the boundary below is what a
 * Python binding (ctypes) of its per-tick path would bind.  Each entry point cites the
 * reference interface it replaces (paths relative to the reference tree):
 *
 *   fxenv_create / fxenv_destroy   GymFxEnv.__init__ / close            app/env.py:36-97, 177-178
 *   fxenv_load_candles             data_feed load_data -> dataframe     data_feed_plugins/default_data_feed.py:36-56
 *                                  + GymFxEnv.dataframe / total_bars    app/env.py:63-68
 *   fxenv_reset                    GymFxEnv.reset                       app/env.py:102-129 (+ bt_bridge.py:33-66)
 *   fxenv_step / fxenv_step_host   GymFxEnv.step                        app/env.py:131-172
 *                                  (BTBridgeStrategy.next               app/bt_bridge.py:119-150,
 *                                   strategy apply_action               strategy_plugins/direct_{fixed,atr}_sltp.py,
 *                                   backtrader BackBroker.next          [external],
 *                                   reward compute_reward               reward_plugins/{pnl,sharpe,dd_penalized}_reward.py,
 *                                   preprocessor make_observation       preprocessor_plugins/{default,feature_window}_preprocessor.py)
 *   fxenv_observe                  GymFxEnv._make_observation           app/env.py:226-242
 *   fxenv_get_info                 GymFxEnv._make_info                  app/env.py:244-254
 *   fxenv_get_state/set_state      (no counterpart: env snapshot)
 *
 * Conventions: every function returns 0 on success, <0 on error (see FXENV_E_*); fxenv_last_error() gives the
 * message.  No exceptions or aborts cross the ABI.  The library owns env state (device struct-of-arrays) behind
 * an opaque handle; the CALLER owns every I/O buffer and passes raw pointers (+ a cudaStream_t as void*).
 * "_dev" pointers are device memory, "_host" pointers host memory (pinned for best speed).  A handle is bound
 * to the CUDA device that was current at fxenv_create and is not thread-safe.  No host synchronisation happens
 * inside fxenv_step/fxenv_reset/fxenv_observe: work is enqueued on the caller's stream.
 */
#ifndef FXENV_H_
#define FXENV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FXENV_ABI_VERSION 2
#define FXENV_MAX_PAIRS 8
#define FXENV_MAX_FEATURES 16
#define FXENV_MAX_COLS 16

/* status codes */
#define FXENV_OK 0
#define FXENV_E_INVALID (-1)   /* bad argument / unsupported configuration */
#define FXENV_E_CUDA (-2)      /* a CUDA runtime call failed */
#define FXENV_E_STATE (-3)     /* call out of order (e.g. step before candles were loaded / before reset) */
#define FXENV_E_NOMEM (-4)

/* enumerations used in FxConfig */
enum { FX_ACTION_DISCRETE = 0, FX_ACTION_CONTINUOUS = 1 };
enum { FX_STRATEGY_DEFAULT = 0, FX_STRATEGY_FIXED_SLTP = 1, FX_STRATEGY_ATR_SLTP = 2 };
enum { FX_PREPROC_DEFAULT = 0, FX_PREPROC_FEATURE_WINDOW = 1 };
enum { FX_SCALING_NONE = 0, FX_SCALING_ROLLING = 1, FX_SCALING_EXPANDING = 2 };
enum { FX_REWARD_PNL = 0, FX_REWARD_SHARPE = 1, FX_REWARD_DD = 2 };
enum { FX_SIZE_FX_UNITS = 0, FX_SIZE_NOTIONAL = 1 };

/* per-env status bits (FxInfoPtrs.flags) */
#define FX_FLAG_STARTED 1u        /* the first step() (which does not advance the bar) has happened */
#define FX_FLAG_TERMINATED 2u     /* bridge.terminated                    app/bt_bridge.py:141,154 */
#define FX_FLAG_EXHAUSTED 4u      /* data ran out (strategy.stop())       app/bt_bridge.py:152-155 */
#define FX_FLAG_BROKE 8u          /* equity <= min_equity                 app/bt_bridge.py:203-204 */
#define FX_FLAG_ORDER_OVERFLOW 16u/* order table full: an order was dropped (the reference is unbounded) */
#define FX_FLAG_TRADE_PRICE_OWN 32u /* statistics bookkeeping: the open trade's average price differs from the position's */

/*
 * The reference resolves one "dict of everything" at call time (app/config.py:1-45, plugin_params of every
 * plugin, app/main.py:42-45).  The host side resolves it ONCE, per each plugin's own precedence rules, into
 * this POD.  All doubles are used exactly as the reference uses the corresponding Python float.
 */
typedef struct FxConfig {
  int32_t struct_size;              /* = sizeof(FxConfig); checked by fxenv_create */
  int32_t num_envs;
  int32_t num_pairs;                /* candle tables; env i trades pair (i % num_pairs) */
  int32_t n_cols;                   /* float64 columns per candle row; cols 0..4 = OPEN,HIGH,LOW,CLOSE,VOLUME */
  int32_t order_capacity;           /* order-table entries per env (0 = default 128) */
  int32_t auto_reset;               /* 1: an env that terminated at step k is reset at step k+1 */
  int64_t episode_bars;             /* bars per episode window; 0 = from the start bar to the end of the table */

  /* GymFxEnv                                                         app/env.py:56-80 */
  double initial_cash;
  double position_size;             /* default order flow size         app/bt_bridge.py:172 */
  double min_equity;
  int32_t action_mode;              /* FX_ACTION_*                     app/env.py:71-80,187-204 */
  int32_t _pad0;
  double continuous_action_threshold;

  /* default_broker -> backtrader BackBroker                          broker_plugins/default_broker.py:35-53 */
  double commission;                /* fraction of notional */
  double leverage;
  double slippage_perc;             /* fraction of price per fill, set_slippage_perc(perc, slip_open/limit/match=True) */
  int32_t children_same_bar;        /* 0 (backtrader: bracket children activate next cycle) | 1 */

  /* strategy plugin                                                  strategy_plugins/direct_{fixed,atr}_sltp.py */
  int32_t strategy;                 /* FX_STRATEGY_* */
  double strat_position_size;
  double sl_pips, tp_pips, pip_size;              /* direct_fixed_sltp.py:24-29 */
  double pair_pip_size[FXENV_MAX_PAIRS];          /* per-pair override of pip_size (0 = use pip_size) */
  int32_t atr_period;                             /* direct_atr_sltp.py:49-86 */
  int32_t use_rel_volume;
  double k_sl, k_tp;
  double rel_volume, strat_leverage, min_order_volume, max_order_volume;
  int32_t size_mode;                /* FX_SIZE_* */
  int32_t use_min_frac, use_max_frac;
  int32_t session_filter;
  double min_sltp_frac, max_sltp_frac;
  int32_t entry_dow_start, entry_hour_start, force_close_dow, force_close_hour;

  /* preprocessor plugin                                              preprocessor_plugins/ *.py */
  int32_t preproc;                  /* FX_PREPROC_* */
  int32_t window_size;
  int32_t price_col;                /* column index of price_column */
  int32_t n_features;
  int32_t feature_cols[FXENV_MAX_FEATURES];
  int32_t feature_binary[FXENV_MAX_FEATURES];
  int32_t scaling;                  /* FX_SCALING_* */
  int32_t scaling_window;
  int32_t include_price_window, include_agent_state;
  double feature_clip;
  double obs_position_size;         /* config.get("position_size", 1.0) as read by the preprocessors */

  /* reward plugin                                                    reward_plugins/ *.py */
  int32_t reward;                   /* FX_REWARD_* */
  int32_t sharpe_window;
  double reward_initial_cash;       /* float(config["initial_cash"]) or 1.0 */
  double reward_scale;
  double annualization_factor;
  double penalty_lambda;
} FxConfig;

typedef struct FxEnv FxEnv;

/* Device pointers to the per-env info columns (arrays of num_envs), valid until fxenv_destroy.
 * Mirrors GymFxEnv._make_info (app/env.py:244-254) + step()'s additions (:162-166). */
typedef struct FxInfoPtrs {
  const double* equity;
  const double* prev_equity;        /* pnl = equity - prev_equity */
  const double* price;
  const double* cash;
  const double* position_size;      /* signed units held */
  const double* position_price;
  const double* commission_paid;
  const int32_t* position;          /* -1 / 0 / +1 */
  const int32_t* bar_index;         /* bridge.bar_index = len(data) = local bar + 1 */
  const int32_t* total_bars;
  const int32_t* trades;
  const int32_t* n_orders;          /* live order-table entries */
  const uint32_t* flags;            /* FX_FLAG_* */
  /* [num_envs][FXENV_RUN_STATS] float64: what backtrader's DrawDown / TradeAnalyzer / SQN analyzers (attached by
   * app/bt_bridge.py:230-234) hold for the current episode -- the inputs of GymFxEnv.summary()
   * (app/env.py:256-271 -> metrics_plugins/default_metrics.py:48-60).  Field order: FXENV_RS_*. */
  const double* run_stats;
} FxInfoPtrs;

#define FXENV_RUN_STATS 12
enum {
  FXENV_RS_DD_MAXVALUE = 0, /* running peak of the broker value */
  FXENV_RS_DD_MAX_MONEY,    /* drawdown.max.moneydown */
  FXENV_RS_DD_MAX_PCT,      /* drawdown.max.drawdown (percent) */
  FXENV_RS_TR_PNL, FXENV_RS_TR_COMM, FXENV_RS_TR_PRICE, /* the open trade (TR_PRICE valid only with FX_FLAG_TRADE_PRICE_OWN) */
  FXENV_RS_PNL_NET,         /* trades.pnl.net.total (average = / closed trades = FxInfoPtrs.trades) */
  FXENV_RS_PNL_SQ,          /* sum of squares of the closed trades' net pnl: with n = FxInfoPtrs.trades, mean = PNL_NET / n,
                             * sqn = sqrt(n) * mean / sqrt(PNL_SQ / n - mean^2)  (n > 1) */
  FXENV_RS_SPARE,
  FXENV_RS_OPENED,          /* trades.total.total */
  FXENV_RS_WON, FXENV_RS_LOST
};

int fxenv_abi_version(void);

/* Replaces GymFxEnv.__init__ (app/env.py:36-97: config resolution, spaces, min_equity) and, per reset, build_cerebro /
 * build_bt_broker (app/bt_bridge.py:207-238, broker_plugins/default_broker.py:35-53).  Validates the POD config and
 * allocates the per-env device state; fails (no CPU path) when no CUDA device is available. */
int fxenv_create(const FxConfig* cfg, FxEnv** out);
/* Replaces GymFxEnv.close (app/env.py:177-178). */
int fxenv_destroy(FxEnv* env);
const char* fxenv_last_error(const FxEnv* env); /* env may be NULL: error of the last failed fxenv_create */

/* Replaces data_feed.load_data + build_bt_feed (data_feed_plugins/default_data_feed.py:36-79: the dataframe the env
 * keeps, app/env.py:62-67, incl. its "too short for the window" ValueError).
 * Copies a host float64 [T, n_cols] row-major candle table (and optional int64 [T] minutes-since-epoch
 * timestamps, needed only by the ATR session filter) to the device, and precomputes the per-bar rolling
 * z-score statistics.  Synchronous. */
int fxenv_load_candles(FxEnv* env, int pair_id, const double* candles_host, int64_t T, const int64_t* minutes_host);

/* Replaces the observation_space bookkeeping of app/env.py:81-90.
 * Number of float32 per observation row and the offsets of its parts (flat layout):
 * [features W*F | prices W | returns W | position | equity_norm | unrealized_pnl_norm | steps_remaining_norm] */
int64_t fxenv_obs_dim(const FxEnv* env);

/* Replaces GymFxEnv.reset (app/env.py:102-129: new bridge / broker / feed, first publish at bar 0).
 * start_bar_dev: int64 [num_envs] first bar (row of the pair's table) of each env's episode window, or NULL to
 * keep the current ones (all 0 after create).  mask_dev: uint8 [num_envs], reset only where != 0, or NULL = all. */
int fxenv_reset(FxEnv* env, const int64_t* start_bar_dev, const uint8_t* mask_dev, void* stream);

/* Replaces _make_observation for the current state (app/env.py:226-242 -> preprocessor.make_observation).
 * Writes the observation of the current state (what reset() returns). obs_dev: float32 [num_envs, obs_dim]. */
int fxenv_observe(FxEnv* env, float* obs_dev, void* stream);

/* Replaces GymFxEnv.step (app/env.py:131-172) and everything below it: BTBridgeStrategy.next (app/bt_bridge.py:119-150),
 * the strategy / reward / preprocessor plugins and backtrader's broker pass.
 * One env.step() for every env.  actions_dev: int32 [num_envs] (discrete) or float32 [num_envs] (continuous).
 * obs_dev float32 [num_envs, obs_dim]; reward_dev float32 [num_envs]; terminated_dev uint8 [num_envs].
 * reward64_dev: optional float64 [num_envs] copy of the reward before the float32 cast (may be NULL). */
int fxenv_step(FxEnv* env, const void* actions_dev, float* obs_dev, float* reward_dev, uint8_t* terminated_dev,
               double* reward64_dev, void* stream);

/* n_steps consecutive steps with the actions of the whole batch supplied up front (replayed / random / scripted
 * drivers: strategy_plugins/default_strategy.py:38-53, the loop of app/main.py:57-65); step k reads
 * actions_dev + k*num_envs and writes reward/terminated at k*num_envs; obs rows go to
 * obs_dev + (k % obs_slots)*num_envs*obs_dim (obs_slots >= 1).  Results are identical to n_steps calls of fxenv_step.
 * Two engines (fxenv_step_many_engine): 1 = one persistent launch whose warps pull (round of consecutive steps, env)
 * tickets and honour per-env dependencies only (every batch of more than one step); 0 = a CUDA graph of n_steps
 * single-step launches, cached by pointer set. */
int fxenv_step_many(FxEnv* env, int n_steps, const void* actions_dev, float* obs_dev, int obs_slots,
                    float* reward_dev, uint8_t* terminated_dev, void* stream);

/* Which engine fxenv_step_many would use for a batch of n_steps (1 persistent launch / 0 graph of steps), <0 on error. */
int fxenv_step_many_engine(const FxEnv* env, int n_steps);

/* Same as fxenv_step for callers that hold numpy / host buffers like the reference's own loop (app/main.py:57-65,
 * tools/smoke_test.py:79-83).  Reference-facing call with HOST buffers: H2D actions, one step, D2H obs/reward/terminated, then waits. */
int fxenv_step_host(FxEnv* env, const void* actions_host, float* obs_host, float* reward_host,
                    uint8_t* terminated_host);

/* Replaces GymFxEnv._make_info (app/env.py:244-254): zero-copy device views instead of a dict of Python floats. */
int fxenv_get_info(FxEnv* env, FxInfoPtrs* out);

/* No reference counterpart (its env cannot be cloned: one daemon thread per instance, app/bt_bridge.py:30-66).
 * Snapshot / restore of the whole env state (host buffer of fxenv_state_bytes() bytes). Synchronous. */
int64_t fxenv_state_bytes(const FxEnv* env);
int fxenv_get_state(FxEnv* env, void* buf_host, int64_t nbytes);
int fxenv_set_state(FxEnv* env, const void* buf_host, int64_t nbytes);

/* Kernels launched by this handle since creation (bench.py's gpu_launches). */
int64_t fxenv_launch_count(const FxEnv* env);

/* ---- closed loop: a policy on the device between the steps ------------------------------------------------------
 * Replaces the caller loop of app/main.py:57-65 (`action = strategy.decide_action(obs, info, step); env.step(action)`)
 * for a learned actor-critic (BASELINE configs[3]: PPO MLP(256,256)): observation rows never leave the GPU, and the
 * policy is one fused tensor-core kernel per step (tcgen05 / TMEM / TMA, gym_fx_b200/csrc/fx_policy.cu), chained to the
 * env step kernel by programmatic dependent launches.  Discrete action mode only.
 *
 *   h1 = tanh(obs W1^T + b1), h2 = tanh(h1 W2^T + b2), logits = h2 Wpi^T + bpi (3), value = h2 wv + bv
 *   action = argmax(logits + Gumbel noise), logp = log_softmax(logits)[action]
 * The two hidden layers run in bfloat16 with float32 accumulation (the env step writes a bfloat16 copy of each row for
 * this purpose); biases, heads, sampling and log-prob in float32. */
typedef struct FxPolicy FxPolicy;

/* DEVICE pointers to float32 parameters in torch.nn.Linear layout ([out][in] row-major). */
typedef struct FxPolicyWeights {
  const float* w1;   /* [256][obs_dim] */
  const float* b1;   /* [256] */
  const float* w2;   /* [256][256] */
  const float* b2;   /* [256] */
  const float* w_pi; /* [3][256] */
  const float* b_pi; /* [3] */
  const float* w_v;  /* [256] */
  const float* b_v;  /* [1] */
} FxPolicyWeights;

/* Buffers of one rollout of `horizon` steps (all DEVICE, caller-owned).  Step t: the policy reads the observation of
 * slot t % obs_slots, writes actions/logp/value at [t], the env step writes reward/done at [t] and the next observation
 * into slot (t + 1) % obs_slots; value[horizon] is the value of the last observation (bootstrap).  obs_slots >= 2
 * (horizon + 1 keeps every observation for the learner). */
typedef struct FxRollout {
  int32_t horizon;
  int32_t obs_slots;
  float* obs;           /* [obs_slots][num_envs][obs_dim] */
  int32_t* actions;     /* [horizon][num_envs] */
  float* logp;          /* [horizon][num_envs] */
  float* value;         /* [horizon + 1][num_envs] */
  float* reward;        /* [horizon][num_envs] */
  uint8_t* done;        /* [horizon][num_envs] */
  const float* gumbel;  /* [horizon][num_envs][3] Gumbel(0,1) noise, or NULL: counter-based generator from `seed` */
  uint64_t seed;
} FxRollout;

int fxenv_policy_create(FxEnv* env, FxPolicy** out);
/* Converts / copies the parameters into the policy's own device buffers (stream-ordered; call after every optimiser step). */
int fxenv_policy_set_weights(FxPolicy* pol, const FxPolicyWeights* weights_dev, void* stream);
int fxenv_policy_destroy(FxPolicy* pol);
/* `horizon` closed-loop steps starting from the env's current state; everything is enqueued on `stream`
 * (2 * horizon + 2 kernels; the launch sequence is cached as a CUDA graph per buffer set). */
int fxenv_rollout(FxEnv* env, FxPolicy* pol, const FxRollout* io, void* stream);
/* Inside a rollout the policy kernel and the env-step kernel hand 128-env tiles to each other through flags in device
 * memory instead of whole-kernel dependencies; a poll that is never answered gives up after ~0.3 s instead of hanging
 * the device.  Returns how many polls of the LAST rollout gave up (0 unless something is broken; then that rollout's
 * results are invalid), or <0.  Synchronises the device. */
int fxenv_policy_sync_timeouts(FxPolicy* pol);

#ifdef __cplusplus
}
#endif
#endif /* FXENV_H_ */
