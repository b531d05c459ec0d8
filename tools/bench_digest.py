#!/usr/bin/env python3
"""tools/bench_digest.py bench.json... -- the numbers of bench.py's one JSON line that matter, on a few lines."""
import json, sys
for p in sys.argv[1:]:
    d = json.load(open(p))
    if d.get("impl") == "reference":
        print(f"{p}: reference arm {d['value']/1e6:.2f} M env-steps/s ({d['cpu_baseline']['cores']} threads)")
        continue
    print(f"{p}: n_gpus {d['n_gpus']} K={d['steps']} W={d['warmup']}: value {d['value']/1e6:.1f} M = {d['ms_per_step']*1e3:.2f} us/step, "
          f"frac {d['roofline']['frac']:.3f}, e2e {d['e2e']['value']/1e6:.2f} M, launches {d['gpu_launches']}, clocks {d['clocks']['sm_mhz'] if d.get('clocks') else None}")
    if "single_step_graph" in d:
        print(f"   single_step_graph {d['single_step_graph']['ms_per_step']*1e3:.2f} us/step; cpu_baseline {d.get('cpu_baseline', {}).get('value', 0)/1e6:.2f} M")
    cl = d.get("closed_loop")
    if cl:
        if "error" in cl:
            print("   closed_loop ERROR", cl["error"])
        else:
            L = cl["learner"]
            print(f"   closed_loop H={cl['horizon']}: {cl['ms_per_step']*1e3:.2f} us/step = {cl['value']/1e6:.1f} M; update {L['update_ms']:.2f} ms, "
                  f"allreduce {L['allreduce_ms_per_update']:.3f} ms ({L['allreduce_calls_per_update']} calls, share {L['allreduce_share_of_update']:.3f}), train {L['train_value']/1e6:.1f} M")
    for k, v in (d.get("other_workloads") or {}).items():
        print(f"   {k}: " + (f"ERROR {v['error']}" if "error" in v else f"{v['ms_per_step']*1e3:.2f} us/step, {v['value']/1e6:.1f} M, frac {v['roofline_frac']:.3f}, engine {v['engine']}"))
