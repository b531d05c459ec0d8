for rb in 2368 2072 1776 1480 1184; do for wl in cfg2 cfg4; do
FXENV_ROLLOUT_BLOCKS=$rb python bench.py --workload $wl --steps 1000 --warmup 300 --no-cpu-baseline --no-single-step 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blocks $rb $wl %7.2f us/step  %7.1f M steps/s' % (d['ms_per_step']*1e3, d['value']/1e6))"
done; done
