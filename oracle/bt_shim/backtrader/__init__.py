"""
oracle/bt_shim/backtrader -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A minimal stand-in for the third-party ``backtrader`` package (PyPI, unpinned in
the reference's setup.py:36-42; de-facto 1.9.78.123), which is not installed in
this image and cannot be fetched.  It exists so that the reference's OWN files
(/root/reference/app/env.py, app/bt_bridge.py and every plugin) can be imported
and run UNMODIFIED to generate the golden trajectories under tests/golden/.

Only the surface the reference touches is provided (call sites:
app/bt_bridge.py:27,69,122,171,181-190,194-199,219-234;
broker_plugins/default_broker.py:47-52; data_feed_plugins/default_data_feed.py:79;
strategy_plugins/direct_fixed_sltp.py:58-77; strategy_plugins/direct_atr_sltp.py:117-198).

PARITY STATUS: the broker arithmetic below restates backtrader's published
algorithm (brokers/bbroker.py: submit / check_submitted / next / _try_exec_* /
_execute / _bracketize, position.py: Position.update, comminfo.py:
CommInfoBase) FROM MEMORY.  The only known answers produced through the real
package are examples/results/{buy_hold,flat}_summary.json (market order fills
at next bar's open; value = cash + size*close) -- both reproduced bit-exactly
(tests/test_oracle_golden.py).  Limit/Stop/bracket/margin/commission/leverage
behaviour is "parity unpinned" against real backtrader.

Deliberately simple and object-based (like backtrader itself) so that it is an
independent third implementation next to oracle/fxenv_oracle.c and the CUDA
kernel.
"""
from __future__ import annotations

import copy

import collections
import datetime as _dt
import itertools
from typing import Any, Dict, List, Optional

__version__ = "0.0-shim(1.9.78.123 semantics)"


# ----------------------------------------------------------------------------
# params helper (backtrader's metaclass turns `params = ((k, v), ...)` into self.p)
# ----------------------------------------------------------------------------
class _Params:
    def __init__(self, pairs, overrides):
        for k, v in pairs:
            setattr(self, k, v)
        for k, v in overrides.items():
            setattr(self, k, v)


def _collect_params(cls) -> List[tuple]:
    out: "collections.OrderedDict[str, Any]" = collections.OrderedDict()
    for klass in reversed(cls.__mro__):
        p = klass.__dict__.get("params", ())
        if isinstance(p, dict):
            p = tuple(p.items())
        for k, v in p:
            out[k] = v
    return list(out.items())


class TimeFrame:
    (Ticks, MicroSeconds, Seconds, Minutes, Days, Weeks, Months, Years, NoTimeFrame) = range(1, 10)


# ----------------------------------------------------------------------------
# Orders
# ----------------------------------------------------------------------------
class _OrderData:
    def __init__(self, size=0.0, price=0.0, pclose=0.0):
        self.size = size          # created: requested signed size; executed: filled so far
        self.price = price
        self.pclose = pclose
        self.remsize = size       # remaining (signed)
        self.comm = 0.0
        self.pnl = 0.0
        self.value = 0.0


class Order:
    Market, Close, Limit, Stop, StopLimit, StopTrail, StopTrailLimit, Historical = range(8)
    ExecTypes = ["Market", "Close", "Limit", "Stop", "StopLimit", "StopTrail", "StopTrailLimit", "Historical"]
    Buy, Sell = range(2)
    (Created, Submitted, Accepted, Partial, Completed, Canceled, Expired, Margin, Rejected) = range(9)
    Cancelled = Canceled
    Status = ["Created", "Submitted", "Accepted", "Partial", "Completed", "Canceled", "Expired", "Margin", "Rejected"]

    _refgen = itertools.count(1)

    def __init__(self, owner, data, size, price=None, exectype=None, parent=None, transmit=True, ordtype=Buy):
        self.ref = next(Order._refgen)
        self.owner = owner
        self.data = data
        self.ordtype = ordtype
        self.size = abs(size) if ordtype == Order.Buy else -abs(size)
        self.price = price
        self.exectype = Order.Market if exectype is None else exectype
        self.parent = parent
        self.transmit = transmit
        self.status = Order.Created
        self._active = parent is None
        # reference price when none is given: the close of the bar of creation
        pclose = data.close[0]
        cprice = pclose if not price else price
        self.created = _OrderData(size=self.size, price=cprice, pclose=pclose)
        self.executed = _OrderData(size=0.0, price=0.0)
        self.executed.remsize = self.size
        self.created_bar = len(data)

    # --- queries
    def isbuy(self):
        return self.ordtype == Order.Buy

    def issell(self):
        return self.ordtype == Order.Sell

    def alive(self):
        return self.status in (Order.Created, Order.Submitted, Order.Partial, Order.Accepted)

    def active(self):
        return self._active

    def activate(self):
        self._active = True

    # --- transitions
    def submit(self):
        self.status = Order.Submitted

    def accept(self):
        self.status = Order.Accepted

    def cancel(self):
        self.status = Order.Canceled

    def margin(self):
        self.status = Order.Margin

    def reject(self):
        self.status = Order.Rejected

    def expire(self):
        return False  # valid=None everywhere on this path

    def execute(self, size, price, closed, closedcomm, opened, openedcomm, pnl):
        self.executed.size += size
        self.executed.price = price
        self.executed.comm += closedcomm + openedcomm
        self.executed.pnl += pnl
        self.executed.remsize -= size
        self.status = Order.Completed if not self.executed.remsize else Order.Partial

    def clone(self):
        c = Order.__new__(Order)
        c.__dict__.update(self.__dict__)
        ex = _OrderData()
        ex.__dict__.update(self.executed.__dict__)
        c.executed = ex
        return c


class BuyOrder(Order):
    pass


class SellOrder(Order):
    pass


# ----------------------------------------------------------------------------
# Position / Trade / commission scheme
# ----------------------------------------------------------------------------
class Position:
    def __init__(self, size=0, price=0.0):
        self.size = size
        self.price = price if size else 0.0

    def clone(self):
        return Position(self.size, self.price)

    def __bool__(self):
        return bool(self.size != 0)

    def update(self, size, price):
        """Returns (newsize, newprice, opened, closed)."""
        oldsize = self.size
        self.size += size
        if not self.size:
            opened, closed = 0, size
            self.price = 0.0
        elif not oldsize:
            opened, closed = size, 0
            self.price = price
        elif oldsize > 0:
            if size > 0:
                opened, closed = size, 0
                self.price = (self.price * oldsize + size * price) / self.size
            elif self.size > 0:
                opened, closed = 0, size
            else:
                opened, closed = self.size, -oldsize
                self.price = price
        else:
            if size < 0:
                opened, closed = size, 0
                self.price = (self.price * oldsize + size * price) / self.size
            elif self.size < 0:
                opened, closed = 0, size
            else:
                opened, closed = self.size, -oldsize
                self.price = price
        return self.size, self.price, opened, closed

    def pseudoupdate(self, size, price):
        return self.clone().update(size, price)


class Trade:
    """backtrader trade.py, restated from memory [BT-from-memory]: size, average price, accumulated gross pnl and
    commission of ONE round trip; pnlcomm = pnl - commission (what TradeAnalyzer / SQN read when it closes)."""

    def __init__(self):
        self.size = 0
        self.price = 0.0
        self.commission = 0.0
        self.pnl = 0.0
        self.pnlcomm = 0.0
        self.isclosed = False
        self.isopen = False
        self.justopened = False

    def update(self, size, price=0.0, commission=0.0):
        if not size:
            return
        self.commission += commission
        oldsize = self.size
        self.size += size
        self.justopened = bool(not oldsize and size)
        self.isopen = bool(self.size)
        self.isclosed = bool(oldsize and not self.size)
        if abs(self.size) > abs(oldsize):   # position increased: new average price, no pnl
            self.price = (oldsize * self.price + size * price) / self.size
            pnl = 0.0
        else:                               # reduced / closed: comminfo.profitandloss(-size, self.price, price), stock-like
            pnl = (-size) * (price - self.price)
        self.pnl += pnl
        self.pnlcomm = self.pnl - self.commission


class CommInfoBase:
    """setcommission(commission=c, leverage=L): margin=None, commtype=None =>
    stock-like, COMM_PERC with percabs=True (commission is a fraction)."""

    def __init__(self, commission=0.0, leverage=1.0, mult=1.0):
        self.commission = commission
        self.leverage = leverage
        self.mult = mult
        self.stocklike = True

    def get_leverage(self):
        return self.leverage

    def getvaluesize(self, size, price):
        return size * price

    def getoperationcost(self, size, price):
        return abs(size) * price

    def profitandloss(self, size, price, newprice):
        return size * (newprice - price) * self.mult

    def getcommission(self, size, price):
        return abs(size) * self.commission * price

    def cashadjust(self, size, price, newprice):
        return 0.0  # stock-like


# ----------------------------------------------------------------------------
# BackBroker
# ----------------------------------------------------------------------------
class BackBroker:
    def __init__(self):
        self.startingcash = self.cash = 10000.0
        self.comminfo = CommInfoBase()
        self.checksubmit = True
        self.shortcash = True
        self.slip_perc = 0.0
        self.slip_fixed = 0.0
        self.slip_open = False
        self.slip_match = True
        self.slip_limit = True
        self.slip_out = False
        # oracle switch: backtrader >= 1.9.5x defers the activation of
        # bracket children to the NEXT broker cycle (`_toactivate`); set True to test
        # them on the parent's fill bar instead.
        self.children_same_bar = False
        self.init()

    def init(self):
        self._value = self.cash
        self.orders: List[Order] = []
        self.pending: "collections.deque" = collections.deque()
        self._toactivate: "collections.deque" = collections.deque()
        self.submitted: "collections.deque" = collections.deque()
        self.position = Position()
        self.notifs: "collections.deque" = collections.deque()
        self._pchildren: Dict[int, collections.deque] = collections.defaultdict(collections.deque)
        self.data = None

    # --- configuration (broker_plugins/default_broker.py:47-52)
    def setcash(self, cash):
        self.startingcash = self.cash = cash
        self._value = cash

    set_cash = setcash

    def setcommission(self, commission=0.0, margin=None, mult=1.0, leverage=1.0, **kw):
        self.comminfo = CommInfoBase(commission=commission, leverage=leverage, mult=mult)

    def set_slippage_perc(self, perc, slip_open=True, slip_limit=True, slip_match=True, slip_out=False):
        self.slip_perc = perc
        self.slip_fixed = 0.0
        self.slip_open = slip_open
        self.slip_limit = slip_limit
        self.slip_match = slip_match
        self.slip_out = slip_out

    # --- queries
    def getcash(self):
        return self.cash

    get_cash = getcash

    def getvalue(self, datas=None):
        return self._value

    get_value = getvalue

    def getposition(self, data=None):
        return self.position

    def start(self):
        self.init_cash = self.cash
        self._value = self.cash

    # --- notifications
    def notify(self, order):
        self.notifs.append(order.clone())

    def get_notification(self):
        try:
            return self.notifs.popleft()
        except IndexError:
            return None

    # --- submission
    def _take_children(self, order):
        oref = order.ref
        pref = getattr(order.parent, "ref", oref)
        if oref != pref:
            if pref not in self._pchildren:
                order.reject()
                self.notify(order)
                return None
        return pref

    def submit(self, order, check=True):
        pref = self._take_children(order)
        if pref is None:
            return order
        pc = self._pchildren[pref]
        pc.append(order)
        if order.transmit:
            rets = [self.transmit(x, check=check) for x in pc]
            return rets[-1]
        return order

    def transmit(self, order, check=True):
        if check and self.checksubmit:
            order.submit()
            self.submitted.append(order)
            self.orders.append(order)
            self.notify(order)
        else:
            self.submit_accept(order)
        return order

    def submit_accept(self, order):
        order.submit()
        order.accept()
        self.pending.append(order)
        self.notify(order)

    def buy(self, owner, data, size, price=None, exectype=None, parent=None, transmit=True, **kw):
        order = BuyOrder(owner, data, size, price, exectype, parent, transmit, ordtype=Order.Buy)
        return self.submit(order)

    def sell(self, owner, data, size, price=None, exectype=None, parent=None, transmit=True, **kw):
        order = SellOrder(owner, data, size, price, exectype, parent, transmit, ordtype=Order.Sell)
        return self.submit(order)

    def cancel(self, order, bracket=False):
        try:
            self.pending.remove(order)
        except ValueError:
            return False
        order.cancel()
        self.notify(order)
        if not bracket:
            self._bracketize(order, cancel=True)
        return True

    def _bracketize(self, order, cancel=False):
        oref = order.ref
        pref = getattr(order.parent, "ref", oref)
        parent = oref == pref
        pc = self._pchildren[pref]
        if cancel or not parent:
            while pc:
                self.cancel(pc.popleft(), bracket=True)
            del self._pchildren[pref]
        else:
            pc.popleft()  # the parent itself
            for o in pc:
                if self.children_same_bar:
                    o.activate()
                else:
                    self._toactivate.append(o)

    # --- per-bar processing
    def check_submitted(self):
        cash = self.cash
        position = None
        while self.submitted:
            order = self.submitted.popleft()
            if self._take_children(order) is None:
                continue
            if position is None:
                position = self.position.clone()
            cash = self._execute(order, cash=cash, position=position)
            if cash >= 0.0:
                self.submit_accept(order)
                continue
            order.margin()
            self.notify(order)
            self._bracketize(order, cancel=True)

    def next(self):
        while self._toactivate:
            self._toactivate.popleft().activate()
        if self.checksubmit:
            self.check_submitted()
        self.pending.append(None)
        while True:
            order = self.pending.popleft()
            if order is None:
                break
            if not order.active():
                self.pending.append(order)
            else:
                self._try_exec(order)
                if order.alive():
                    self.pending.append(order)
                elif order.status == Order.Completed:
                    self._bracketize(order)
        self._get_value()

    def _get_value(self):
        ci = self.comminfo
        pos = self.position
        close = self.data.close[0]
        pos_value_unlever = 0.0
        dvalue = ci.getvaluesize(pos.size, close)
        dunrealized = ci.profitandloss(pos.size, pos.price, close)
        if dvalue > 0:
            dvalue -= dunrealized
            pos_value_unlever += dvalue / ci.get_leverage()
            pos_value_unlever += dunrealized
        else:
            pos_value_unlever += dvalue
        self._value = self.cash + pos_value_unlever
        return self._value

    # --- slippage helpers
    def _slip_up(self, pmax, price, doslip=True, lim=False):
        if not doslip:
            return price
        if self.slip_perc:
            pslip = price * (1 + self.slip_perc)
        elif self.slip_fixed:
            pslip = price + self.slip_fixed
        else:
            return price
        if pslip <= pmax:
            return pslip
        elif self.slip_match or (lim and self.slip_limit):
            if not self.slip_out:
                return pmax
            return pslip
        return None

    def _slip_down(self, pmin, price, doslip=True, lim=False):
        if not doslip:
            return price
        if self.slip_perc:
            pslip = price * (1 - self.slip_perc)
        elif self.slip_fixed:
            pslip = price - self.slip_fixed
        else:
            return price
        if pslip >= pmin:
            return pslip
        elif self.slip_match or (lim and self.slip_limit):
            if not self.slip_out:
                return pmin
            return pslip
        return None

    # --- matching
    def _try_exec(self, order):
        data = order.data
        popen, phigh, plow = data.open[0], data.high[0], data.low[0]
        pcreated = order.created.price
        if order.exectype == Order.Market:
            if len(data) <= order.created_bar:
                return  # can only execute after the bar of creation
            if order.isbuy():
                p = self._slip_up(phigh, popen, doslip=self.slip_open)
            else:
                p = self._slip_down(plow, popen, doslip=self.slip_open)
            self._execute(order, ago=0, price=p)
        elif order.exectype == Order.Limit:
            plimit = pcreated
            if order.isbuy():
                if plimit >= popen:
                    pmax = min(phigh, plimit)
                    p = self._slip_up(pmax, popen, doslip=self.slip_open, lim=True)
                    self._execute(order, ago=0, price=p)
                elif plimit >= plow:
                    self._execute(order, ago=0, price=plimit)
            else:
                if plimit <= popen:
                    pmin = max(plow, plimit)
                    p = self._slip_down(plimit, popen, doslip=self.slip_open, lim=True)
                    self._execute(order, ago=0, price=p)
                elif plimit <= phigh:
                    self._execute(order, ago=0, price=plimit)
        elif order.exectype == Order.Stop:
            if order.isbuy():
                if popen >= pcreated:
                    p = self._slip_up(phigh, popen, doslip=self.slip_open)
                    self._execute(order, ago=0, price=p)
                elif phigh >= pcreated:
                    p = self._slip_up(phigh, pcreated)
                    self._execute(order, ago=0, price=p)
            else:
                if popen <= pcreated:
                    p = self._slip_down(plow, popen, doslip=self.slip_open)
                    self._execute(order, ago=0, price=p)
                elif plow <= pcreated:
                    p = self._slip_down(plow, pcreated)
                    self._execute(order, ago=0, price=p)
        else:  # pragma: no cover
            raise NotImplementedError("exectype not on the gym-fx hot path")

    def _execute(self, order, ago=None, price=None, cash=None, position=None):
        # ago is None  => pseudo-execution (check_submitted); returns the remaining cash
        if ago is not None and price is None:
            return
        size = order.executed.remsize
        ci = self.comminfo
        if ago is not None:
            position = self.position
            pprice_orig = position.price
            psize, pprice, opened, closed = position.pseudoupdate(size, price)
            pnl = ci.profitandloss(-closed, pprice_orig, price)
            cash = self.cash
        else:
            pnl = 0
            price = pprice_orig = order.created.price
            psize, pprice, opened, closed = position.update(size, price)

        if closed:
            if self.shortcash:
                closedvalue = ci.getvaluesize(-closed, pprice_orig)
            else:
                closedvalue = ci.getoperationcost(closed, pprice_orig)
            closecash = closedvalue
            if closedvalue > 0:
                closecash /= ci.get_leverage()
            cash += closecash + pnl * ci.stocklike
            closedcomm = ci.getcommission(closed, price)
            cash -= closedcomm
            if ago is not None:
                cash += ci.cashadjust(-closed, 0.0, price)
                self.cash = cash
        else:
            closedvalue = closedcomm = 0.0

        popened = opened
        if opened:
            if self.shortcash:
                openedvalue = ci.getvaluesize(opened, price)
            else:
                openedvalue = ci.getoperationcost(opened, price)
            opencash = openedvalue
            if openedvalue > 0:
                opencash /= ci.get_leverage()
            cash -= opencash
            openedcomm = ci.getcommission(opened, price)
            cash -= openedcomm
            if cash < 0.0:
                opened = 0
                openedvalue = openedcomm = 0.0
            elif ago is not None:
                self.cash = cash
        else:
            openedvalue = openedcomm = 0.0

        if ago is None:
            return cash

        execsize = closed + opened
        if execsize:
            position.update(execsize, price)
            order.execute(execsize, price, closed, closedcomm, opened, openedcomm, pnl)
            order._exbit = (closed, opened, price, closedcomm, openedcomm)
            self.notify(order)
        if popened and not opened:
            order.margin()
            self.notify(order)
            self._bracketize(order, cancel=True)


class _Brokers:
    BackBroker = BackBroker
    BrokerBack = BackBroker


brokers = _Brokers()


# ----------------------------------------------------------------------------
# Data feed
# ----------------------------------------------------------------------------
class _Line:
    def __init__(self, feed, values):
        self._feed = feed
        self._values = values

    def __getitem__(self, ago):
        return float(self._values[self._feed._idx + ago])

    def __len__(self):
        return self._feed._idx + 1


class _DateTimeLine:
    def __init__(self, feed, index):
        self._feed = feed
        self._index = index

    def datetime(self, ago=0):
        ts = self._index[self._feed._idx + ago]
        return ts.to_pydatetime() if hasattr(ts, "to_pydatetime") else ts

    def __getitem__(self, ago):
        d = self.datetime(ago)
        return d.toordinal() + (d - _dt.datetime.combine(d.date(), _dt.time.min)).total_seconds() / 86400.0


class DataBase:
    pass


class PandasData(DataBase):
    """Auto-detects lower-case open/high/low/close/volume/openinterest columns,
    datetime from the index (data_feed_plugins/default_data_feed.py:58-79)."""

    def __init__(self, dataname=None, **kw):
        df = dataname
        self._df = df
        self._idx = -1
        self._n = len(df)
        cols = {str(c).lower(): c for c in df.columns}
        def col(name):
            if name in cols:
                return df[cols[name]].to_numpy(dtype=float)
            return [0.0] * len(df)
        self.open = _Line(self, col("open"))
        self.high = _Line(self, col("high"))
        self.low = _Line(self, col("low"))
        self.close = _Line(self, col("close"))
        self.volume = _Line(self, col("volume"))
        self.openinterest = _Line(self, col("openinterest"))
        self.datetime = _DateTimeLine(self, df.index)

    def __len__(self):
        return self._idx + 1

    def buflen(self):
        return self._n

    def advance(self):
        self._idx += 1
        return self._idx < self._n


class _Feeds:
    PandasData = PandasData
    DataBase = DataBase


feeds = _Feeds()


# ----------------------------------------------------------------------------
# Analyzers: attached unconditionally by app/bt_bridge.py:230-234.  Their results are unreachable on the step path
# (env.summary() sees them only after cerebro.run returned), so they change nothing that the
# trajectory goldens compare; DrawDown / TradeAnalyzer / SQN are restated here [BT-from-memory: analyzers/drawdown.py,
# tradeanalyzer.py, sqn.py] so that the end-of-run summary (metrics_plugins/default_metrics.py:48-60) has an oracle.
# SharpeRatio(timeframe=Days) / TimeReturn need calendar-day buckets of the equity curve and stay empty.
# ----------------------------------------------------------------------------
class _AutoDict(dict):
    def __missing__(self, key):
        v = self[key] = _AutoDict()
        return v


class Analyzer:
    def __init__(self, **kw):
        self.kw = kw

    def _fund(self, value):     # Strategy._notify -> analyzer._notify_fund, once per bar, before next()
        pass

    def _next(self):            # Strategy._next_analyzers, once per bar, after strategy.next()
        pass

    def _trade(self, trade):    # analyzer._notify_trade
        pass

    def _stop(self):
        pass

    def get_analysis(self):
        return {}


class _DrawDown(Analyzer):
    def __init__(self, **kw):
        super().__init__(**kw)
        self._value, self._maxvalue = 0.0, float("-inf")
        self.rets = {"len": 0, "drawdown": 0.0, "moneydown": 0.0, "max": {"len": 0, "drawdown": 0.0, "moneydown": 0.0}}

    def _fund(self, value):
        self._value = value
        self._maxvalue = max(self._maxvalue, value)

    def _next(self):
        r = self.rets
        r["moneydown"] = moneydown = self._maxvalue - self._value
        r["drawdown"] = drawdown = 100.0 * moneydown / self._maxvalue
        r["max"]["moneydown"] = max(r["max"]["moneydown"], moneydown)
        r["max"]["drawdown"] = max(r["max"]["drawdown"], drawdown)
        r["len"] = r["len"] + 1 if drawdown else 0
        r["max"]["len"] = max(r["max"]["len"], r["len"])

    def get_analysis(self):
        return self.rets


class _TradeAnalyzer(Analyzer):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.rets = _AutoDict()
        self.rets["total"]["total"] = 0

    def _trade(self, trade):
        t = self.rets
        if trade.justopened:
            t["total"]["total"] += 1
            t["total"]["open"] = t["total"].get("open", 0) + 1
        elif trade.isclosed:
            t["total"]["open"] = t["total"].get("open", 0) - 1
            t["total"]["closed"] = t["total"].get("closed", 0) + 1
            won = trade.pnlcomm >= 0.0
            for key, hit in (("won", won), ("lost", not won)):
                t[key]["total"] = t[key].get("total", 0) + int(hit)
            net = t["pnl"]["net"]
            net["total"] = net.get("total", 0.0) + trade.pnlcomm
            net["average"] = net["total"] / t["total"]["closed"]

    def get_analysis(self):
        return self.rets


class _SQN(Analyzer):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.pnl, self.rets = [], {}

    def _trade(self, trade):
        if trade.isclosed:
            self.pnl.append(trade.pnlcomm)

    def _stop(self):
        import math
        n = len(self.pnl)
        if n > 1:
            av = math.fsum(self.pnl) / n
            sd = math.sqrt(math.fsum((x - av) ** 2 for x in self.pnl) / n)
            try:
                sqn = math.sqrt(n) * av / sd
            except ZeroDivisionError:
                sqn = None
        else:
            sqn = 0
        self.rets = {"sqn": sqn, "trades": n}

    def get_analysis(self):
        return self.rets


class _Analyzers:
    TradeAnalyzer = _TradeAnalyzer
    SharpeRatio = type("SharpeRatio", (Analyzer,), {})
    DrawDown = _DrawDown
    SQN = _SQN
    TimeReturn = type("TimeReturn", (Analyzer,), {})


analyzers = _Analyzers()


class _AnalyzerBag:
    def _all(self):
        return [a for a in vars(self).values() if isinstance(a, Analyzer)]


# ----------------------------------------------------------------------------
# Strategy / Cerebro
# ----------------------------------------------------------------------------
class Strategy:
    params = ()

    def __new__(cls, *args, **kwargs):
        self = object.__new__(cls)
        return self

    # wiring done by Cerebro before __init__ (backtrader does it in the metaclass)
    def _wire(self, cerebro, data, broker, kwargs):
        self.env = self.cerebro = cerebro
        self.data = self.data0 = data
        self.datas = [data]
        self.broker = broker
        self.p = self.params = _Params(_collect_params(type(self)), kwargs)
        self.analyzers = _AnalyzerBag()
        self._orderspending: List[Order] = []
        self._tradespending: List[Trade] = []
        self._trade = None

    # --- default lifecycle hooks
    def start(self):
        pass

    def prenext(self):
        pass

    def nextstart(self):
        self.next()

    def next(self):
        pass

    def stop(self):
        pass

    def notify_order(self, order):
        pass

    def notify_trade(self, trade):
        pass

    def __len__(self):
        return len(self.data)

    # --- position / orders
    @property
    def position(self):
        return self.broker.getposition(self.data)

    def getposition(self, data=None):
        return self.broker.getposition(self.data)

    def buy(self, data=None, size=None, price=None, exectype=None, parent=None, transmit=True, **kw):
        size = size if size is not None else 1
        if size:
            return self.broker.buy(self, self.data, size=abs(size), price=price, exectype=exectype,
                                   parent=parent, transmit=transmit)
        return None

    def sell(self, data=None, size=None, price=None, exectype=None, parent=None, transmit=True, **kw):
        size = size if size is not None else 1
        if size:
            return self.broker.sell(self, self.data, size=abs(size), price=price, exectype=exectype,
                                    parent=parent, transmit=transmit)
        return None

    def close(self, data=None, size=None, **kw):
        possize = self.position.size
        size = abs(size if size is not None else possize)
        if possize > 0:
            return self.sell(size=size, **kw)
        elif possize < 0:
            return self.buy(size=size, **kw)
        return None

    def buy_bracket(self, data=None, size=None, price=None, exectype=Order.Limit,
                    stopprice=None, stopexec=Order.Stop, limitprice=None, limitexec=Order.Limit, **kw):
        o = self.buy(size=size, price=price, exectype=exectype, transmit=False)
        ostop = self.sell(size=o.size, price=stopprice, exectype=stopexec, parent=o, transmit=False)
        olimit = self.sell(size=o.size, price=limitprice, exectype=limitexec, parent=o, transmit=True)
        return [o, ostop, olimit]

    def sell_bracket(self, data=None, size=None, price=None, exectype=Order.Limit,
                     stopprice=None, stopexec=Order.Stop, limitprice=None, limitexec=Order.Limit, **kw):
        o = self.sell(size=size, price=price, exectype=exectype, transmit=False)
        ostop = self.buy(size=o.size, price=stopprice, exectype=stopexec, parent=o, transmit=False)
        olimit = self.buy(size=o.size, price=limitprice, exectype=limitexec, parent=o, transmit=True)
        return [o, ostop, olimit]

    # --- notification plumbing (strategy.py:_addnotification / _notify)
    def _addnotification(self, order):
        self._orderspending.append(order)
        if order.status not in (Order.Completed, Order.Partial):
            return
        exbit = getattr(order, "_exbit", None)
        if exbit is None:
            return
        closed, opened = exbit[0], exbit[1]
        price = exbit[2] if len(exbit) > 2 else 0.0
        closedcomm = exbit[3] if len(exbit) > 3 else 0.0
        openedcomm = exbit[4] if len(exbit) > 4 else 0.0
        if self._trade is None:
            self._trade = Trade()
        trade = self._trade
        if closed:
            trade.update(closed, price, closedcomm)
            if trade.isclosed:
                self._tradespending.append(copy.copy(trade))
        if opened:
            if trade.isclosed:
                trade = self._trade = Trade()
            trade.update(opened, price, openedcomm)
            if trade.justopened:
                self._tradespending.append(copy.copy(trade))

    def _notify(self):
        pending, self._orderspending = self._orderspending, []
        for order in pending:
            self.notify_order(order)
        tpending, self._tradespending = self._tradespending, []
        for trade in tpending:
            for an in self.analyzers._all():
                an._trade(trade)
            self.notify_trade(trade)
        value = self.broker.getvalue()
        for an in self.analyzers._all():
            an._fund(value)


class Cerebro:
    def __init__(self, stdstats=True, **kw):
        self._data = None
        self._broker = BackBroker()
        self._strats: List[tuple] = []
        self._analyzers: List[tuple] = []
        self._event_stop = False
        self.runningstrats: List[Strategy] = []

    def adddata(self, data, name=None):
        self._data = data
        return data

    def setbroker(self, broker):
        self._broker = broker

    def getbroker(self):
        return self._broker

    broker = property(getbroker, setbroker)

    def addstrategy(self, strategy, *args, **kwargs):
        self._strats.append((strategy, args, kwargs))

    def addanalyzer(self, ancls, *args, **kwargs):
        self._analyzers.append((ancls, kwargs))

    def runstop(self):
        self._event_stop = True

    def run(self, **kwargs):
        self._event_stop = False
        data, broker = self._data, self._broker
        broker.data = data
        data._idx = -1
        strats = []
        for cls, args, kw in self._strats:
            s = cls.__new__(cls)
            s._wire(self, data, broker, kw)
            s.__init__(*args)
            for ancls, akw in self._analyzers:
                name = akw.get("_name", ancls.__name__.lower())
                setattr(s.analyzers, name, ancls(**{k: v for k, v in akw.items() if k != "_name"}))
            strats.append(s)
        self.runningstrats = strats
        broker.start()
        for s in strats:
            s.start()
        first = True
        while data.advance():
            broker.next()
            while True:
                order = broker.get_notification()
                if order is None:
                    break
                (order.owner or strats[0])._addnotification(order)
            if self._event_stop:
                break
            for s in strats:
                s._notify()
                if first:
                    s.nextstart()
                else:
                    s.next()
                for an in s.analyzers._all():   # Strategy._next -> _next_analyzers, after next()
                    an._next()
            first = False
            if self._event_stop:
                break
        if data._idx >= data._n:
            data._idx = data._n - 1
        for s in strats:
            s.stop()
            for an in s.analyzers._all():
                an._stop()
        return strats
