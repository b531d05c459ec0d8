// tests/hostsim/fx_hostsim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles gym_fx_b200/csrc/fx_core.cuh (the scalar state machine the CUDA step kernel calls) with g++ and
// drives it with the same per-env control flow as fx_step_kernel, sequentially, so that the broker / strategy /
// reward logic can be diffed against the oracle on the GPU-less build box BEFORE GPU time is spent.  The
// warp-parallel parts of the kernel (order staging, ballot hit masks, compaction, observation streaming) are
// NOT exercised here -- only `-m gpu` tests cover those.  Nothing in the product loads this library.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../gym_fx_b200/csrc/fx_core.cuh"

struct HsEnv {
  FxConfig c;
  const double* tbl; const int64_t* minutes; int64_t T; int pair;
  FxEnvRegs e;
  int32_t t, total_bars; int64_t start;
  std::vector<uint32_t> meta; std::vector<double> p0, p1, sz; int n;
  std::vector<double> ring; int32_t sh_len, sh_head, sh_last, dd_last; double dd_peak;
  double rs[FX_RS_N];  // end-of-run statistics (DrawDown / TradeAnalyzer / SQN state)
};

extern "C" {

HsEnv* hs_create(const FxConfig* cfg, int pair, const double* tbl, int64_t T, const int64_t* minutes) {
  HsEnv* h = new HsEnv();
  h->c = *cfg; h->tbl = tbl; h->T = T; h->minutes = minutes; h->pair = pair;
  int cap = cfg->order_capacity ? cfg->order_capacity : 128; cap = (cap + 31) & ~31; h->c.order_capacity = cap;
  h->meta.assign(cap + FXO_SLACK, 0); h->p0.assign(cap + FXO_SLACK, 0); h->p1.assign(cap + FXO_SLACK, 0); h->sz.assign(cap + FXO_SLACK, 0); h->n = 0;
  h->ring.assign(cfg->sharpe_window > 0 ? cfg->sharpe_window : 1, 0.0);
  h->sh_len = h->sh_head = 0; h->sh_last = h->dd_last = -1; h->dd_peak = 0.0;
  return h;
}
void hs_destroy(HsEnv* h) { delete h; }

void hs_reset(HsEnv* h, int64_t start) {
  const FxConfig& c = h->c;
  if (start < 0) start = 0; if (start > h->T - 1) start = h->T - 1;
  h->start = start;
  int64_t tb = h->T - start; if (c.episode_bars > 0 && c.episode_bars < tb) tb = c.episode_bars;
  h->total_bars = (int32_t)tb; h->t = 0; h->n = 0;
  FxEnvRegs& e = h->e;
  e.cash = c.initial_cash; e.value = c.initial_cash; e.psize = 0; e.pprice = 0; e.equity = e.prev_equity = c.initial_cash;
  e.commission_paid = 0; e.trades = 0; e.position = 0; e.flags = 0;
  e.price = h->tbl[start * c.n_cols + 3]; e.bar_index = 1;
  for (int i = 0; i < FX_RS_N; i++) h->rs[i] = 0.0;
  h->rs[FX_RS_DD_MAXVALUE] = c.initial_cash;  // bar 0: notify_fund(cash), next()
}

void hs_step(HsEnv* h, double action_in, double* reward, uint8_t* term) {
  const FxConfig& c = h->c; FxEnvRegs& e = h->e; const int C = c.n_cols;
  int action = c.action_mode == FX_ACTION_CONTINUOUS ? fx_coerce_continuous(c, (float)action_in) : fx_coerce_discrete((int)action_in);
  if (e.flags & FX_FLAG_TERMINATED) {
    if (c.auto_reset) { hs_reset(h, h->start); *reward = 0; *term = 0; return; }
    *reward = 0; *term = 1; return;
  }
  bool exhausted = false, advance = false;
  int32_t t = h->t;
  if (!(e.flags & FX_FLAG_STARTED)) e.flags |= FX_FLAG_STARTED;
  else if (t + 1 >= h->total_bars) exhausted = true;
  else { t += 1; advance = true; }
  const double* r = h->tbl + (h->start + t) * (int64_t)C;
  FxBar b{r[0], r[1], r[2], r[3]};
  FxOrderTab tab{h->meta.data(), h->p0.data(), h->p1.data(), h->sz.data(), h->n, c.order_capacity, h->n, 0, 0.0, fx_bound_per(c)};
  if (advance && tab.n > 0) {
    const int n = tab.n;
    std::vector<char> hit(n);
    int first_sub = n;
    for (int k = 0; k < n; k++) {
      tab.meta[k] = fx_entry_begin_bar(tab.meta[k]);
      if ((tab.meta[k] & FXO_SUBMITTED) && first_sub == n) first_sub = k;
      hit[k] = fx_entry_hits(tab.meta[k], tab.p0[k], tab.p1[k], b);
    }
    fx_check_submitted(c, e, tab, first_sub);
    for (int k = 0; k < n; k++) if (hit[k]) fx_exec_entry(c, e, tab, k, b, FxRunStatsMem{h->rs});
  }
  if (advance) { fx_mark_to_market(c, e, b.c); fx_rs_drawdown(FxRunStatsMem{h->rs}, e.value); }
  if (!exhausted) {
    double atr = 0; bool ready = false;
    if (c.strategy == FX_STRATEGY_ATR_SLTP && action != 0) {
      const int period = c.atr_period; const int nb = (t + 1 < period) ? t + 1 : period;
      double s_ = 0, comp = 0;
      for (int j = 0; j < nb; j++) {
        const int k = t - nb + 1 + j;
        const double* rr = h->tbl + (h->start + k) * (int64_t)C;
        const double tr = fx_true_range(rr[1], rr[2], k > 0 ? rr[3 - C] : 0.0, k > 0);
        if (j == 0) s_ = tr; else fx_neumaier_add(s_, comp, tr);
      }
      atr = fx_neumaier_done(s_, comp) / (double)nb; ready = nb >= period;
    }
    const bool has_min = h->minutes != nullptr;
    fx_apply_action(c, c.strategy, e, tab, action, b, h->pair, atr, ready, has_min, (c.session_filter && has_min) ? h->minutes[h->start + t] : 0);
    fx_publish(e, b.c, t);
    if (e.equity <= c.min_equity) e.flags |= FX_FLAG_TERMINATED | FX_FLAG_BROKE;
  } else e.flags |= FX_FLAG_TERMINATED | FX_FLAG_EXHAUSTED;
  double rw;
  if (c.reward == FX_REWARD_PNL) rw = fx_reward_pnl(c, e);
  else if (c.reward == FX_REWARD_DD) rw = fx_reward_dd(c, e, h->dd_peak, h->dd_last);
  else {
    const double ret = (e.equity - e.prev_equity) / c.reward_initial_cash;
    const int nn = fx_sharpe_push(h->ring.data(), 1, c.sharpe_window, h->sh_len, h->sh_head, h->sh_last, e.bar_index, ret);
    rw = fx_sharpe_eval(h->ring.data(), 1, c.sharpe_window, nn, h->sh_head, c.annualization_factor);
  }
  {  // stable compaction after the strategy call (v2 kernel phase C)
    int w = 0;
    for (int k = 0; k < tab.n; k++) if (!(tab.meta[k] & FXO_DEAD)) {
      tab.meta[w] = tab.meta[k]; tab.p0[w] = tab.p0[k]; tab.p1[w] = tab.p1[k]; tab.sz[w] = tab.sz[k]; w++;
    }
    tab.n = w;
  }
  h->t = t; h->n = tab.n;
  *reward = rw;
  *term = (uint8_t)(((e.flags & FX_FLAG_TERMINATED) != 0) || e.equity <= c.min_equity);
}

void hs_scalars(HsEnv* h, float* out4) {
  const FxConfig& c = h->c; const FxEnvRegs& e = h->e;
  int s = e.bar_index; if (s < 0) s = 0; if (s > h->total_bars) s = h->total_bars;
  const double last = h->tbl[(h->start + s - 1) * (int64_t)c.n_cols + c.price_col];
  const bool inc_price = c.preproc == FX_PREPROC_DEFAULT || c.include_price_window;
  double ref = c.preproc == FX_PREPROC_DEFAULT ? last : (inc_price ? (double)(float)last : e.price);
  fx_agent_scalars(c, e, h->total_bars, ref, 1.0 / (c.initial_cash != 0.0 ? c.initial_cash : 1.0), out4);
}

void hs_info(HsEnv* h, double* d7, int32_t* i5, uint32_t* flags) {
  const FxEnvRegs& e = h->e;
  d7[0] = e.equity; d7[1] = e.prev_equity; d7[2] = e.price; d7[3] = e.cash; d7[4] = e.psize; d7[5] = e.pprice; d7[6] = e.commission_paid;
  i5[0] = e.position; i5[1] = e.bar_index; i5[2] = h->total_bars; i5[3] = e.trades; i5[4] = h->n;
  *flags = e.flags;
}

void hs_stats(HsEnv* h, double* out) { for (int i = 0; i < FX_RS_N; i++) out[i] = h->rs[i]; }

// Cross-check of the two forms of the trigger test the kernel uses: fx_entry_fill (select form, also returns the
// execution price; what the CUDA kernel evaluates per lane) against fx_entry_hits + the branchy fx_match_* rules (what
// fx_exec_entry uses).  Random entries / bars, including bars whose OPEN lies outside [LOW, HIGH] and exact ties.
// Returns the number of disagreements.
int hs_check_entry_fill(int n, unsigned seed) {
  unsigned long long x = 0x9E3779B97F4A7C15ull ^ seed;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  auto price = [&]() { return 1.0 + (double)(rnd() % 41) * 0.0005; };  // coarse grid => many exact ties
  int bad = 0;
  for (int i = 0; i < n; i++) {
    FxBar b{price(), price(), price(), price()};
    const uint32_t kind = (uint32_t)(rnd() % 3);
    const uint32_t meta = kind | ((rnd() & 1) ? FXO_SELL : 0u) | ((rnd() & 1) ? FXO_ACTIVE : 0u);
    const double p0 = price(), p1 = price();
    const double slip = (rnd() % 3 == 0) ? 0.0 : (double)(rnd() % 7) * 2.5e-4;  // with and without slippage
    double px = -1.0;
    const bool hit = fx_entry_fill(slip, meta, p0, p1, b, px);
    const bool buy = !(meta & FXO_SELL);
    double ref_px = b.o;
    bool ref;
    if (kind == FXO_MARKET) { ref = true; ref_px = fx_market_price(slip, buy, b); }
    else if (kind == FXO_PARENT) ref = fx_match_limit(slip, buy, p0, b, ref_px);
    else ref = fx_match_stop(slip, buy, p0, b, ref_px) || fx_match_limit(slip, buy, p1, b, ref_px);
    if (hit != ref || hit != fx_entry_hits(meta, p0, p1, b) || (hit && px != ref_px)) bad++;
  }
  return bad;
}

}  // extern "C"
