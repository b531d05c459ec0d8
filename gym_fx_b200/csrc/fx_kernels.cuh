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
struct FxDeviceState {
  double *cash, *psize, *pprice, *value;                  // broker: cash, position size/price, cached value
  double *equity, *prev_equity, *price, *commission_paid; // bridge (app/bt_bridge.py:30-66)
  double* dd_peak;                                        // dd_penalized_reward._peak
  int64_t* start;                                         // first bar (table row) of the episode window
  int32_t *t, *total_bars, *position, *bar_index, *trades, *n_orders;
  int32_t *sh_len, *sh_head, *sh_last_step, *dd_last_step;
  uint32_t* flags;
  double* sh_ring;   // [N][sharpe_window]
  uint32_t* o_meta;  // [N][cap]   order table, entry-major per env (a warp scans one env's entries coalesced)
  double *o_p0, *o_p1, *o_sz;
};

struct FxKernelParams {
  FxConfig cfg;
  FxPairTable pair[FXENV_MAX_PAIRS];
  FxDeviceState st;
  int32_t obs_dim;
  int32_t cap;          // order-table capacity (multiple of 32)
  int32_t smem_per_warp;
  int32_t fast_features;  // 1: feature columns are 0..F-1 == all table columns (contiguous window block)
};

#define FX_WARPS_PER_BLOCK 4

// host-callable launchers (fx_kernels.cu)
cudaError_t fx_launch_step(const FxKernelParams& P, const void* actions, float* obs, float* reward, double* reward64,
                           uint8_t* terminated, cudaStream_t stream);
cudaError_t fx_launch_reset(const FxKernelParams& P, const int64_t* start_bar, const uint8_t* mask, int first,
                            cudaStream_t stream);
cudaError_t fx_launch_observe(const FxKernelParams& P, float* obs, cudaStream_t stream);
cudaError_t fx_launch_stats(const FxConfig& cfg, const double* candles, double* stats, int64_t T, cudaStream_t stream);
size_t fx_smem_per_warp(const FxConfig& cfg, int cap);
cudaError_t fx_configure_kernels(size_t smem_per_block);
