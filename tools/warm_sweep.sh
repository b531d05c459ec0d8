for w in 5 200 2000 10000; do
  python bench.py --steps 20 --warmup $w --no-cpu-baseline --no-single-step --no-closed-loop --no-other-workloads 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('warmup=%-6d K=20 %7.2f us/step  clocks %s' % ($w, d['ms_per_step']*1e3, d['clocks']))"
done
