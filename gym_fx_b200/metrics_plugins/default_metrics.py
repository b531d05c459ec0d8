"""default_metrics mirror (reference: metrics_plugins/default_metrics.py:22-60): end-of-run summary dict.
Post-run host glue; the analyzer-derived fields of the reference are always None/0
on the live path (App. B #12), so only the equity-derived ones and the trade counter are filled."""
from ..plugin_base import PluginBase


class Plugin(PluginBase):
    plugin_kind = "default_metrics"
    plugin_params = {}

    def summarize(self, *, initial_cash, final_equity, analyzers, config):
        ic, fe = float(initial_cash), float(final_equity)
        tr = (analyzers or {}).get("trades") or {}
        total = tr.get("total", {}).get("total", 0) if isinstance(tr.get("total"), dict) else 0
        return {
            "initial_cash": ic, "final_equity": fe, "total_return": (fe / ic - 1.0) if ic else 0.0,
            "max_drawdown_pct": None, "max_drawdown_money": None, "sharpe_ratio": None, "sqn": None,
            "trades_total": total, "trades_won": 0, "trades_lost": 0, "avg_trade_pnl": None,
        }
