"""Check that two outputs of run_baseline_configs.py (e.g. 1 GPU and 8 GPUs) agree on every RESULT field.

    python examples/compare_config_runs.py profiles/r02_baseline_configs_1gpu.jsonl profiles/r02_baseline_configs_8gpu.jsonl

Throughput / timing fields and the GPU count are expected to differ; everything else (episode counts, success rates,
per-cell tables, failure-mode counts, learner returns) must be identical -- envs are keyed by global id, so sharding
them over more GPUs may not change a single counter.  Exit code 1 on any difference.
"""
import json
import sys

SKIP = {"n_gpus", "env_steps_per_sec", "wall_seconds", "timing"}


def load(path):
    return {d["config"]: d for d in (json.loads(l) for l in open(path) if l.strip().startswith("{"))}


def main(a_path, b_path):
    a, b = load(a_path), load(b_path)
    bad = 0
    for name in sorted(set(a) | set(b)):
        if name not in a or name not in b:
            print(f"MISSING  {name}")
            bad += 1
            continue
        keys = (set(a[name]) | set(b[name])) - SKIP
        diff = [k for k in sorted(keys) if a[name].get(k) != b[name].get(k)]
        print(f"{'DIFFERS ' if diff else 'identical'}  {name}  ({len(keys)} result fields)" + (f": {diff}" if diff else ""))
        bad += bool(diff)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
