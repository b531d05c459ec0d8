"""default_metrics mirror (reference: metrics_plugins/default_metrics.py:22-60): end-of-run summary dict from the
final equity plus the analyzer results (backtrader get_analysis() shapes: trades / sharpe / drawdown / sqn).  Post-run
host glue; the analyzer dict it receives is produced by the step kernel's per-env statistics (VecFxEnv.analyzers), and is
empty before the run has ended, exactly when the reference's is."""
from ..plugin_base import PluginBase


def _dig(node, *path, default=None):
    """Nested lookup that tolerates missing levels / None, like the reference's `_get` helper."""
    for key in path:
        if not hasattr(node, "get"):
            return default
        node = node.get(key)
        if node is None:
            return default
    return node


class Plugin(PluginBase):
    plugin_kind = "default_metrics"
    plugin_params = {}

    def summarize(self, *, initial_cash, final_equity, analyzers, config):
        ic, fe = float(initial_cash), float(final_equity)
        an = analyzers or {}
        trades, drawdown = an.get("trades") or {}, an.get("drawdown") or {}
        return {
            "initial_cash": ic, "final_equity": fe, "total_return": float((fe / ic - 1.0) if ic else 0.0),
            "max_drawdown_pct": _dig(drawdown, "max", "drawdown"),
            "max_drawdown_money": _dig(drawdown, "max", "moneydown"),
            "sharpe_ratio": _dig(an.get("sharpe") or {}, "sharperatio"),
            "sqn": _dig(an.get("sqn") or {}, "sqn"),
            "trades_total": _dig(trades, "total", "total", default=0),
            "trades_won": _dig(trades, "won", "total", default=0),
            "trades_lost": _dig(trades, "lost", "total", default=0),
            "avg_trade_pnl": _dig(trades, "pnl", "net", "average"),
        }
