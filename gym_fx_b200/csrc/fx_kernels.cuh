// fx_kernels.cuh -- device-side data layout shared by fx_kernels.cu (kernels) and fx_capi.cu (C-ABI host code).
#pragma once

#include "fx_core.cuh"

// One candle table per currency pair, resident in HBM (and, at benchmark sizes, in the 126 MB L2):
//   candles  float64 [T][n_cols] row-major (AoS rows: a W-row window of all columns is ONE contiguous span,
//            so the flat [W][F] observation block maps 1:1 onto a contiguous read when F == n_cols)
//   stats    float64 [T][F][2] = {mean, 1/std} of the rolling z-score window ENDING at that bar (precomputed at
//            load for every bar that has a full window behind it; pure function of the bar)
//   minutes  int64 [T] minutes since the Unix epoch (session filter only) or nullptr
struct FxPairTable {
  const double* candles;
  const double* stats;
  const int64_t* minutes;
  int64_t T;
};

// Per-env state: struct-of-arrays over N envs.  The info columns of the C-ABI (FxInfoPtrs) point straight in here.
// Invariants at kernel boundaries: bar_index == t + 1, position == sign(psize), price == CLOSE[start + t], and the
// broker's cached value == equity -- so the step kernel only LOADS cash/psize/pprice/equity (+ counters) and
// re-derives the rest; prev_equity / price / position / bar_index are stored for the info columns.
struct FxDeviceState {
  double *cash, *psize, *pprice;                          // broker: cash, position size / average price
  double *equity, *prev_equity, *price, *commission_paid; // bridge (app/bt_bridge.py:30-66)
  double* dd_peak;                                        // dd_penalized_reward._peak
  double* sub_need;                                       // check_submitted cash bound of the entries [n_acc, n_orders)
  double* nbar;                                           // [N][6] {open, high, low, close, price_col value, -} of the candle the NEXT step call
                                                          // works on (row min(t+1, total_bars-1), or row t right after a reset): saves the
                                                          // dependent cursor -> candle round trip at the head of every step
  int64_t* start;                                         // first bar (table row) of the episode window
  int32_t *t, *total_bars, *position, *bar_index, *trades, *n_orders, *n_acc;
  int32_t *sh_len, *sh_head, *sh_last_step, *dd_last_step;
  uint32_t* flags;
  double* rstats;    // [N][FX_RS_N] end-of-run statistics: DrawDown / TradeAnalyzer / SQN state (fx_core.cuh FX_RS_*)
  double* sh_ring;   // [N][sharpe_window]
  double* welford;   // [N][FXENV_MAX_FEATURES][2] running {mean, M2} of each feature column over rows [0, s)
  uint32_t* o_meta;  // [N][cap + FXO_SLACK]  order table, entry-major per env (a warp scans one env coalesced)
  double *o_p0, *o_p1, *o_sz;
};

#define FX_NSTAMP 12

// How fx_rollout_kernel cuts the steps [0, n_steps) of a batch into rounds (one ticket = one env for the steps of one
// round): n_uniform rounds of `chunk` steps, then rounds [tail_start[i], tail_start[i + 1]); n_rounds in total.
#define FX_PLAN_TAIL 7
struct FxChunkPlan {
  int32_t chunk, n_uniform, n_rounds;
  int32_t tail_start[FX_PLAN_TAIL + 1];
};

struct FxKernelParams {
  FxConfig cfg;
  FxPairTable pair[FXENV_MAX_PAIRS];
  FxDeviceState st;
  double inv_initial_cash;  // 1 / (initial_cash or 1.0)
  int32_t* seq;             // [N + 1] per-env sequence words of a fxenv_step_many batch + the ticket counter (fx_rollout_kernel)
  long long* timing;        // debug (FXENV_TIMING=1): [N][FX_NSTAMP] clock64() phase stamps of the last step, else nullptr
  long long* timeline;      // debug (FXENV_TIMELINE=K, timing build): [K][N][2] globaltimer at start / end of every ticket
  int32_t obs_dim;
  int32_t cap;              // logical order-table capacity (multiple of 32); arrays hold cap + FXO_SLACK
  int32_t debug;            // timing experiments only (env FXENV_DEBUG, bit mask): 1 skip obs windows, 2 skip broker /
                            // strategy / reward, 4 plain launches (no programmatic dependent launch), 8 step_many always as
                            // the graph of single steps, 16 step_many always as the persistent launch
  int32_t resident_blocks;  // CTAs of the persistent rollout kernel resident at once on this device (SMs x occupancy)
  int32_t tame_data;        // 1: every loaded table value is finite and |x| < 1e100 (no NaN can arise in a z-score)
  int32_t fast_features;    // 5: F == n_cols == 5 with identity columns (the [W][5] block is one contiguous span)
  int32_t any_binary;       // 1: some feature column is a binary pass-through (feature_binary[])
  int32_t lean;             // 1: the configuration qualifies for the specialised kernels (fx_config_is_lean)
  int32_t num_sms;          // SMs of the device (fx_rollout_kernel: CTA b is the (b / num_sms)-th CTA of its SM)
};

// warps (= envs) per CTA of the step kernel.  One warp per CTA lets the second wave back-fill SM slots as soon as a
// single env finishes (measured, cfg2: 4 warps/CTA 26.1 / 64.7 us per step at 4096 / 16384 envs, 2: 25.9 / 64.7,
// 1: 24.2 / 57.9).
#ifndef FX_WARPS
#define FX_WARPS 1
#endif
// CTAs per SM the step kernel is compiled for.  Measured on B200 (cfg2, us/step at 4096 / 16384 envs):
//   (with 4 warps per CTA) 8 (64 regs, 220 B of spills) 30.4 / 73.1 | 6: 29.8 / 71.1 | 5: 28.1 / 69.8 |
//   4 (128 regs, no spills) 26.2 / 65.0 | 3: 28.7 / 75.9.  The spill-free build wins although 4096 envs then run as two waves of 16 warps per SM.
#ifndef FX_MIN_BLOCKS
#define FX_MIN_BLOCKS (16 / FX_WARPS)   // 16 warps per SM => 128 registers, no spills
#endif

// host-callable launchers (fx_kernels.cu)
// one step of the envs [env_begin, env_end) (env_end < 0: all); the array arguments are the bases for env 0
// Fine-grained hand-over between the policy kernel and the env-step kernel of a closed-loop rollout (fxenv_rollout), per
// 128-env tile, instead of whole-kernel dependencies: the step kernel of step t starts an env as soon as the policy has
// published the tile's actions (act_flag[tile] >= t + 1), and the policy kernel of step t + 1 starts a tile as soon as
// its envs have finished step t (done_cnt[tile] == envs of the tile x (t + 1)).  Both kernels are launched with the
// programmatic-dependent-launch attribute and are resident together; a dependent grid is only launched once every CTA
// of its predecessor has started, so whoever is waited for is always running.  nullptr members: plain kernel order.
struct FxTileSync {
  int32_t* act_flag;   // [tiles] written by the policy kernel (release), polled by the step kernel (acquire)
  int32_t* done_cnt;   // [tiles] incremented by the step kernel (release), polled by the policy kernel (acquire)
  int32_t* timeouts;   // [1] number of polls that gave up (a bug or a lost launch: results are then invalid; tests assert 0)
  int32_t epoch;       // t + 1
};
#define FX_SYNC_TILE 128
#define FX_SYNC_MAX_POLLS (1 << 22)   // x ~64 ns: a poll gives up after ~0.3 s instead of hanging the device  // obs16: optional bf16 copy of the rows

cudaError_t fx_launch_step(const FxKernelParams& P, const void* actions, float* obs, float* reward, double* reward64,
                           uint8_t* terminated, cudaStream_t stream, int env_begin = 0, int env_end = -1,
                           uint16_t* obs16 = nullptr, int stride16 = 0, const FxTileSync* sync = nullptr);

cudaError_t fx_launch_reset(const FxKernelParams& P, const int64_t* start_bar, const uint8_t* mask, int first,
                            cudaStream_t stream);
cudaError_t fx_launch_observe(const FxKernelParams& P, float* obs, cudaStream_t stream, uint16_t* obs16 = nullptr,
                              int stride16 = 0);
cudaError_t fx_launch_stats(const FxConfig& cfg, const double* candles, double* stats, int64_t T, cudaStream_t stream);
cudaError_t fx_configure_kernels(FxKernelParams& P);
bool fx_config_is_lean(const FxKernelParams& P);
cudaError_t fx_launch_rollout(const FxKernelParams& P, const void* actions, float* obs, int obs_slots, float* reward,
                              uint8_t* terminated, const FxChunkPlan& plan, unsigned seq_base, unsigned ticket_base,
                              bool reset_words, cudaStream_t stream);
int fx_rollout_blocks(const FxKernelParams& P);
FxChunkPlan fx_rollout_plan(const FxKernelParams& P, int n_steps);  // the host's ticket accounting needs n_rounds
