"""Print the interesting numbers of a bench.py JSON line (experiments / reading gpurun_out)."""
import json
import sys

for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    r, e, s = d.get("roofline") or {}, d.get("e2e") or {}, d.get("strong_1m") or {}
    print(f"== {path}: n_gpus {d.get('n_gpus')} steps {d.get('steps')}")
    print(f"  value {d['value']:.4g}  ms/step {d['ms_per_step']:.5f}  frac {r.get('frac', 0):.3f}  frac_sustained {r.get('frac_sustained', 0):.3f} "
          f"({r.get('kernel_ms_sustained', 0) * 1e3:.2f} us)  traffic {r.get('traffic')}")
    print(f"  e2e {e.get('value', 0):.4g}  of ceiling {e.get('frac_of_copy_ceiling')}  async {e.get('async_value', 0):.4g}  packed {e.get('packed_contacts_value', 0):.4g}  all_rows {e.get('all_rows_value', 0):.4g}  "
          f"ceiling {((e.get('copy_ceiling') or {}).get('value') or 0):.4g}")
    t, nz, f = d.get("tracking_full") or {}, d.get("noisy_api") or {}, d.get("fused_rollout") or {}
    print(f"  tracking_full frac {t.get('roofline_frac', 0):.3f} ({t.get('ms_per_step', 0) * 1e3:.2f} us)  noisy {nz.get('us_per_step', 0):.1f} us  fused {f.get('value', 0):.4g}")
    if s:
        print(f"  strong_1m: per GPU {s['envs_per_gpu']}  eager {s['api_eager']['value']:.4g} (frac {s['api_eager']['per_gpu_roofline_frac']:.3f}, {s['api_eager']['us_per_step']:.2f} us)  "
              f"graph {s['api_graph_replay']['value']:.4g} (frac {s['api_graph_replay']['per_gpu_roofline_frac']:.3f})  fused {s['fused_rollout']['value']:.4g}")
        print(f"    sha {s['counters_sha256'][:16]}  episodes {s['counters_episodes']}  api==fused {s['api_equals_fused']}")
    for sw in d.get("sweep") or []:
        print(f"  sweep {sw['envs']:7d}: eager {sw['value']:.4g} (frac {sw['roofline_frac']:.3f}, {sw['ms_per_step'] * 1e3:.2f} us)  graph {sw['graph_replay_value']:.4g}  fused {sw['fused_rollout_value']:.4g}")
    c = d.get("cpu_baseline") or {}
    print(f"  cpu_baseline {c.get('value')} on {c.get('cores')} cores; single env {((d.get('single_env_dropin') or {}).get('us_per_step'))} us; clocks {d.get('clocks')}")
