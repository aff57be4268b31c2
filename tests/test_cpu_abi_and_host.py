"""CPU-side checks of the product: the C-ABI library loads and exports every symbol that
include/dexsim.h declares, struct layouts agree, argument errors are reported without touching
CUDA, the host-side label entry point matches the reference's classifiers on the golden
episodes, and the host logic (config -> group table, sharding, counter summaries) is right.
No compute kernels are launched here."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest

import dexterous_rl_manipulation_b200 as dx
from dexterous_rl_manipulation_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "dexsim.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dexsim_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 15
    L = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/dexsim.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert L.dexsim_version() == _lib.ABI_VERSION


def test_struct_layouts():
    L = _lib.lib()
    assert L.dexsim_sizeof_state() == C.sizeof(_lib.DexsimState) == 14 * 8
    assert L.dexsim_sizeof_params() == C.sizeof(_lib.DexsimParams)
    assert L.dexsim_sizeof_group() == C.sizeof(_lib.DexsimGroup) == 152
    assert L.dexsim_sizeof_step_io() == C.sizeof(_lib.DexsimStepIO)


def test_argument_errors_without_cuda():
    L = _lib.lib()
    st, p, io = _lib.DexsimState(), _lib.DexsimParams(), _lib.DexsimStepIO()
    assert L.dexsim_step(None, C.byref(p), None, None, C.byref(io), None) == -1001
    st.n, st.ld = 10, 8
    assert L.dexsim_step(C.byref(st), C.byref(p), None, None, C.byref(io), None) == -1002     # ld < n
    st.n, st.ld = 10, 48
    assert L.dexsim_step(C.byref(st), C.byref(p), None, None, C.byref(io), None) == -1002     # ld % 32
    st.ld = 32
    assert L.dexsim_step(C.byref(st), C.byref(p), None, None, C.byref(io), None) == -1001     # NULL arrays
    assert L.dexsim_rollout(C.byref(st), C.byref(p), None, None, 0, 1, C.byref(_lib.DexsimRolloutIO()), None) == -1001
    assert b"NULL" in L.dexsim_error_string(-1001) and b"size" in L.dexsim_error_string(-1002)
    with pytest.raises(dx.DexsimError):
        _lib.check(-1004, "x")
    with pytest.raises(dx.DexsimError):
        dx.BatchedManipulationEnv(4, num_fingers=4)                   # unsupported geometry
    with pytest.raises(RuntimeError):
        dx.BatchedManipulationEnv(4, device="cpu")                    # no CPU path


def _summary(counts):
    c = np.asarray(counts, np.int64)
    n = len(c)
    return dict(hist_len=n, max_count=int(c.max()) if n else 0, sum_counts=int(c.sum()), sum_sq_counts=int((c * c).sum()),
                first5_sum=int(c[:5].sum()), last5_sum=int(c[-5:].sum()) if n else 0)


def test_summary_classifier_matches_reference_labels(golden_dir):
    """dexsim_classify_summary (the arithmetic the kernels run at episode end) vs labels produced by the unmodified
    reference classifiers on 4,000 synthetic episodes -- ALL of them, exact variance ties included: with the history
    (`counts`) a tie is decided by np.var's own pairwise arithmetic (var_tie == 2) and must match the reference
    (evaluation/metrics.py:77-80, evaluation/failure_taxonomy.py:189,219-230); without it the tie is only flagged."""
    with np.load(os.path.join(golden_dir, "labels.npz")) as z:
        g = {k: z[k] for k in z.files}
    ties = mism = mism_summary_only = 0
    for i in range(g["length"].shape[0]):
        c = g["counts"][i, :g["length"][i]]
        ea = 255 if g["label_metrics"][i] < 0 else int(g["label_metrics"][i])
        eb = 255 if g["label_taxonomy"][i] < 0 else int(g["label_taxonomy"][i])
        a, b, tie = dx.classify_summary(g["success"][i], g["steps"][i], g["num"][i], g["final"][i],
                                        max_steps=int(g["max_steps"]), counts=c, **_summary(c))
        assert tie in (0, 2)                              # history at hand: nothing is left in doubt
        mism += (a, b) != (ea, eb)
        assert dx.classify_counts(g["success"][i], g["steps"][i], g["num"][i], g["final"][i], c,
                                  max_steps=int(g["max_steps"])) == (a, b)
        a0, b0, tie0 = dx.classify_summary(g["success"][i], g["steps"][i], g["num"][i], g["final"][i],
                                           max_steps=int(g["max_steps"]), **_summary(c))
        assert (tie0 == 1) == (tie == 2)                  # the summary alone flags exactly the cases the history decides
        ties += tie0 == 1
        if not tie0:
            mism_summary_only += (a0, b0) != (ea, eb)
    assert mism == 0
    assert mism_summary_only == 0
    assert ties > 0                                       # the golden set does contain ties (tests/golden/make_golden.py)


def test_variance_ties_follow_numpy_rounding():
    """Order matters on a tie: np.var of the same multiset of counts lands on either side of the threshold depending
    on the order of the elements.  The product's emulation must agree with NumPy itself on every permutation."""
    rng = np.random.default_rng(31)
    # multisets with n * sum(c^2) - sum(c)^2 == 2 n^2 (variance exactly 2) and a mean that is not a dyadic rational
    m18 = [0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 4, 5]
    m27 = [0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5]
    for base in (m18, m27, m27 * 5, m27 * 7, m18 * 17):     # 135, 189 and 306 entries take NumPy's recursive halving path
        seen = set()
        for trial in range(120):
            c = rng.permutation(np.array(base, np.int64))
            n_, S, Q = len(c), int(c.sum()), int((c * c).sum())
            assert n_ * Q - S * S == 2 * n_ * n_
            v = float(np.var([int(x) for x in c]))        # what the reference evaluates (a list of Python ints)
            fin = int(c[-1])
            a, b, tie = dx.classify_summary(False, n_, fin, fin, max_steps=10 ** 6, counts=c, **_summary(c))
            if fin == 0:
                continue                                  # "object dropped" wins before the variance is looked at
            assert tie == 2
            seen.add(v > 2.0)
            # metrics.py:77-80: UNSTABLE iff np.var > 2.0, tested before the slippage trend
            assert (dx.LABELS_METRICS[a] == "unstable_contacts") == (v > 2.0), (c.tolist(), v)
            # failure_taxonomy.py:208-222: slippage first, then the same variance test
            if dx.LABELS_TAXONOMY[b] != "slippage":
                assert (dx.LABELS_TAXONOMY[b] == "unstable_grasp") == (v > 2.0), (c.tolist(), v)
        assert len(base) > 300 or seen == {False, True}, (len(base), seen)   # both outcomes occur: the order decides


def test_summary_classifier_known_answers(golden_dir):
    import json
    with open(os.path.join(golden_dir, "anchors.json")) as fh:
        anchors = json.load(fh)
    for ka in anchors["known_answers"]:
        a, b, _ = dx.classify_summary(False, ka["episode_steps"], ka["num_contacts"], ka["final_contacts"],
                                      **_summary(ka["counts"]))
        if "metrics" in ka:
            assert dx.LABELS_METRICS[a] in ka["metrics"], ka["src"]
        if "taxonomy" in ka:
            assert dx.LABELS_TAXONOMY[b] in ka["taxonomy"], ka["src"]


def test_trend_tie_uses_ieee_division():
    # first5 = 11, last5 = 6 gives 6/5 - 11/5 = -1.0000000000000002 < -1.0 (True),
    # although (6 - 11)/5 == -1 exactly; an integer restatement would get this wrong.
    counts = [3, 2, 2, 2, 2] + [1] * 6 + [2, 1, 1, 1, 1]
    assert sum(counts[:5]) == 11 and sum(counts[-5:]) == 6
    a, b, tie = dx.classify_summary(False, 50, 1, 1, **_summary(counts))
    assert dx.LABELS_METRICS[a] == "slippage" and dx.LABELS_TAXONOMY[b] == "slippage"


def test_group_table_from_configs():
    CC = dx.CurriculumConfig
    cfg = CC(object_size_range=(0.03, 0.07), friction_range=(0.3, 0.7))
    t = dx.group_table([CC.easy(), cfg], sigma_obs=[0.0, 0.05], sigma_dyn=0.1)
    assert (t[0].size, t[0].mass, t[0].friction, t[0].size_ranged) == (0.08, 0.05, 0.8, 0)
    assert (t[1].size_ranged, t[1].mass_ranged, t[1].fric_ranged) == (1, 0, 1)
    assert (t[1].size_lo, t[1].size_hi, t[1].fric_lo, t[1].fric_hi) == (0.03, 0.07, 0.3, 0.7)
    assert list(t[1].spawn_lo) == [-0.1, -0.1, 0.05] and list(t[1].spawn_hi) == [0.1, 0.1, 0.2]
    assert abs(t[1].sigma_obs - 0.05) < 1e-9 and abs(t[0].sigma_dyn - 0.1) < 1e-7
    with pytest.raises(ValueError):
        dx.group_table([])


def test_curriculum_config_presets_and_samplers():
    CC = dx.CurriculumConfig
    assert (CC.easy().object_size, CC.medium().object_mass, CC.hard().friction_coefficient) == (0.08, 0.1, 0.3)
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence(7)))
    twin = np.random.Generator(np.random.PCG64(np.random.SeedSequence(7)))
    cfg = CC(object_size_range=(0.03, 0.07), object_mass_range=(0.05, 0.15), friction_range=(0.3, 0.7))
    got = (cfg.get_object_size(rng), cfg.get_object_mass(rng), cfg.get_friction_coefficient(rng), cfg.get_spawn_position(rng))
    exp = (float(twin.uniform(0.03, 0.07)), float(twin.uniform(0.05, 0.15)), float(twin.uniform(0.3, 0.7)),
           (float(twin.uniform(-0.1, 0.1)), float(twin.uniform(-0.1, 0.1)), float(twin.uniform(0.05, 0.2))))
    assert got == exp
    fixed = CC.hard()
    state = rng.bit_generator.state
    assert fixed.get_object_size(rng) == 0.03 and rng.bit_generator.state == state      # no draw when not ranged
    assert CC.from_dict(cfg.to_dict()) == cfg


def test_shard_ranges_cover_and_balance():
    for N in (0, 1, 7, 4096, 1_048_576, 1_000_003):
        for W in (1, 2, 3, 4, 8):
            r = [dx.distributed.shard_range(N, k, W) for k in range(W)]
            assert r[0][0] == 0 and r[-1][1] == N
            assert all(r[k][1] == r[k + 1][0] for k in range(W - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_summarize_counters_schema():
    import torch
    c = torch.zeros(2, dx.NCOUNTERS, dtype=torch.int64)
    c[0, dx.CNT_EPISODES], c[0, dx.CNT_SUCCESSES], c[0, dx.CNT_SUM_STEPS], c[0, dx.CNT_SUM_STEPS_SQ] = 10, 6, 500, 30000
    c[0, dx.CNT_LABEL_METRICS + 3] = 4
    rs = torch.tensor([[20.0, 50.0], [0.0, 0.0]], dtype=torch.float64)
    m = dx.distributed.summarize_counters(c, rs)
    assert m[1] == {}
    assert m[0]["grasp_success_rate"] == 0.6 and m[0]["mean_episode_length"] == 50.0
    assert m[0]["failure_type_frequency"]["timeout"] == {"count": 4, "frequency": 0.4}
    assert m[0]["failed_episodes"] == 4 and m[0]["mean_reward"] == 2.0
    assert abs(m[0]["std_episode_length"] - (3000 - 2500) ** 0.5) < 1e-9


def test_scheduler_standin_matches_reference_decisions(golden_dir):
    with np.load(os.path.join(golden_dir, "scheduler.npz")) as z:
        g = {k: z[k] for k in z.files}
    CC = dx.CurriculumConfig
    for t in range(g["kw"].shape[0]):
        thr, mn, win, ps = g["kw"][t]
        sch = dx.CurriculumScheduler(CC.easy(), CC.hard(), success_rate_threshold=float(thr),
                                     min_episodes_before_progression=int(mn), window_size=int(win), progression_steps=int(ps))
        for e in range(g["success"].shape[1]):
            assert sch.update(bool(g["success"][t, e]), int(g["steps"][t, e])) == bool(g["progressed"][t, e])
            c = sch.get_current_config()
            assert (sch.current_difficulty_level, c.object_size, c.object_mass, c.friction_coefficient) == \
                   (g["level"][t, e], g["size"][t, e], g["mass"][t, e], g["friction"][t, e])


class _FakeEnv:
    curriculum_config = None


def test_curriculum_driver_feeds_scheduler_and_pushes_config():
    import torch
    from dexterous_rl_manipulation_b200.curriculum import spread_flag, spread_successes
    CC = dx.CurriculumConfig
    for E, S in ((10, 0), (10, 10), (7, 3), (1000, 371)):
        f = spread_successes(E, S)
        assert len(f) == E and sum(f) == S and all(f[k] == spread_flag(k, E, S) for k in range(E))
    env = _FakeEnv()
    sch = dx.CurriculumScheduler(CC.easy(), CC.hard(), success_rate_threshold=0.3, min_episodes_before_progression=20,
                                 window_size=15, progression_steps=5)
    drv = dx.BatchedCurriculumDriver(env, sch)
    assert env.curriculum_config.object_size == 0.08
    assert drv.feed(10, 9, 150) == 0 and sch.total_episodes == 10 and sch.total_steps == 150
    # 5,000,000 mostly successful episodes in one poll: reaches the target level after a handful of
    # sequential updates, the rest is folded in bulk (must be fast and keep totals exact)
    prog = drv.feed(5_000_000, 4_500_000, 75_000_000)
    assert prog == 5 and sch.current_difficulty_level == 1.0
    assert abs(env.curriculum_config.object_size - 0.03) < 1e-12 and abs(env.curriculum_config.friction_coefficient - 0.3) < 1e-12
    assert sch.total_episodes == 5_000_010 and sch.total_steps == 75_000_150
    assert len(sch.episode_successes) < 200
    # poll() reads deltas from a counter table
    c = torch.zeros(2, dx.NCOUNTERS, dtype=torch.int64)
    c[0, dx.CNT_EPISODES], c[0, dx.CNT_SUCCESSES], c[0, dx.CNT_SUM_STEPS] = 100, 20, 15000
    drv2 = dx.BatchedCurriculumDriver(_FakeEnv(), dx.CurriculumScheduler(CC.easy(), CC.hard()))
    drv2.poll(c); drv2.poll(c)
    assert drv2.scheduler.total_episodes == 100 and drv2.scheduler.total_steps == 15000


def test_driver_with_reference_schedulers_when_available():
    """The driver feeds the reference's UNCHANGED CurriculumScheduler and StepBasedScheduler."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference not present")
    R = ref_harness.load()
    RC = R.CurriculumConfig
    env = _FakeEnv()
    sch = R.CurriculumScheduler(RC.easy(), RC.hard(), success_rate_threshold=0.3, min_episodes_before_progression=20,
                                window_size=15, progression_steps=5)
    drv = dx.BatchedCurriculumDriver(env, sch)
    assert drv.feed(3_000_000, 2_900_000, 40_000_000) == 5 and sch.current_difficulty_level == 1.0
    assert sch.total_episodes == 3_000_000 and sch.total_steps == 40_000_000
    assert abs(env.curriculum_config.object_size - 0.03) < 1e-12
    # a failing population never progresses, however many episodes arrive
    sch2 = R.CurriculumScheduler(RC.easy(), RC.hard(), success_rate_threshold=0.7)
    drv2 = dx.BatchedCurriculumDriver(_FakeEnv(), sch2)
    assert drv2.feed(1_000_000, 100_000, 150_000_000) == 0 and sch2.current_difficulty_level == 0.0
    import importlib
    SB = importlib.import_module("experiments.curriculum_scheduler").StepBasedScheduler
    env3 = _FakeEnv()
    sb = SB(RC.easy(), RC.hard(), step_milestones=[1000, 50_000, 2_000_000, 10_000_000])
    drv3 = dx.BatchedCurriculumDriver(env3, sb)
    assert drv3.feed(10, 5, 500) == 0
    assert drv3.feed(100_000, 50_000, 3_000_000) == 3 and sb.current_difficulty_level == 0.75   # three milestones crossed
    assert drv3.feed(1_000_000, 1, 20_000_000) == 1 and sb.current_difficulty_level == 1.0
    assert abs(env3.curriculum_config.friction_coefficient - 0.3) < 1e-12


def test_empty_batch_is_a_no_op_without_cuda():
    """n == 0 (empty input) returns 0 before any CUDA call; argument checks still apply."""
    L = _lib.lib()
    buf = (C.c_char * 256)()
    ptr = (C.addressof(buf) + 15) & ~15                     # any non-NULL, 16-byte aligned address: never dereferenced
    st = _lib.DexsimState()
    st.n, st.ld = 0, 0
    for name, _ in _lib.DexsimState._fields_[2:12]:
        setattr(st, name, ptr)
    p = _lib.DexsimParams()
    p.reward_type, p.num_groups = 1, 1
    io = _lib.DexsimStepIO()
    io.action = io.reward = io.terminated = io.truncated = io.num_contacts = ptr
    assert L.dexsim_step(C.byref(st), C.byref(p), None, None, C.byref(io), None) == 0
    assert L.dexsim_reset_predrawn(C.byref(st), C.byref(p), None, ptr, ptr, ptr, ptr, None, None) == 0
    rio = _lib.DexsimRolloutIO()
    assert L.dexsim_rollout(C.byref(st), C.byref(p), ptr, None, 5, 1, C.byref(rio), None) == 0
    rio.counters = ptr                                        # counters without the per-env history summary
    assert L.dexsim_rollout(C.byref(st), C.byref(p), ptr, None, 5, 1, C.byref(rio), None) == -1001
    assert L.dexsim_rollout(C.byref(st), C.byref(p), ptr, None, 5, 7, C.byref(_lib.DexsimRolloutIO()), None) == -1004
    p.reward_type = 5
    assert L.dexsim_step(C.byref(st), C.byref(p), None, None, C.byref(io), None) == -1004
    assert L.dexsim_set_step_impl(9) == -1004 and L.dexsim_set_step_impl(0) == 0


def test_public_header_is_plain_c(tmp_path):
    """include/dexsim.h is the boundary a maintainer binds against: it must compile as C99 and link against the
    shared library from a C program (no C++ / torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "dexsim.h"\n#include <stdio.h>\n'
                   'int main(void) { DexsimState st = {0}; DexsimParams p = {0}; DexsimStepIO io = {0};\n'
                   '  int rc = dexsim_step(&st, &p, 0, 0, &io, 0);\n'
                   '  printf("%d %d %s\\n", dexsim_version(), rc, dexsim_error_string(rc)); return 0; }\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-ldexsim_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split(None, 2)
    assert int(out[0]) == _lib.ABI_VERSION and int(out[1]) == -1001 and "NULL" in out[2]


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference runs on host cores only (no GPU): one JSON line on stdout with the contract's keys."""
    import json
    import subprocess
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    proc = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                           "--ref-budget-seconds", "6", "--ref-procs", "2"], capture_output=True, text=True, timeout=300)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    if "unavailable" in d:                 # neither /root/reference nor the byte-compiled copy is present
        assert d["impl"] == "reference"
        return
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]


def test_aggregate_metrics_equals_reference_aggregation():
    """evaluation.aggregate_metrics (host side of every batched front-end) against the UNMODIFIED
    EvaluationMetrics.compute_aggregate_metrics on synthetic episodes: labels come from classify_summary here and
    from the reference's own classifier there, every aggregate key must agree."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference (source or byte-compiled) not present")
    R = ref_harness.load()
    from dexterous_rl_manipulation_b200.evaluation import aggregate_metrics, count_rows
    rng = np.random.default_rng(8)
    T = 60
    for trial in range(6):
        episodes = []
        for _ in range(int(rng.integers(1, 80))):
            steps = int(rng.integers(1, T + 1))
            kind = int(rng.integers(4))
            if kind == 0:
                counts = rng.integers(0, 6, steps)
            elif kind == 1:
                counts = np.minimum(np.arange(steps) // max(1, steps // 4), 5)
            elif kind == 2:
                counts = np.maximum(4 - np.arange(steps) // max(1, steps // 5), 0)
            else:
                counts = np.zeros(steps, np.int64)
            success = bool(counts[-1] >= 3 and rng.integers(2))
            hist = count_rows(counts)
            ep = {"success": success, "episode_steps": steps, "num_contacts": int(counts[-1]), "final_contacts": int(counts[-1]),
                  "contact_history": hist, "episode_reward": float(rng.normal())}
            la, lb, tie = dx.classify_summary(success, steps, int(counts[-1]), int(counts[-1]), max_steps=T, **_summary(counts))
            ep["failure_type"] = None if la == dx.LABEL_NONE else dx.LABELS_METRICS[la]
            episodes.append(ep)
        mine = aggregate_metrics(episodes, T)
        ref = R.metrics.EvaluationMetrics(3).compute_aggregate_metrics([dict(e) for e in episodes], max_steps=T)
        assert set(mine) == set(ref), set(mine) ^ set(ref)
        for k, v in ref.items():
            if isinstance(v, float):
                assert mine[k] == pytest.approx(v, rel=1e-12, abs=1e-12), k
            else:
                assert mine[k] == v, (k, mine[k], v)


def test_argument_checks_need_no_gpu():
    """Argument validation happens before any CUDA call: tracked episodes longer than the packed history summary can
    hold (5,242 steps) are rejected with DEXSIM_E_PARAM instead of wrapping silently (the sums of the per-step contact
    counts are 16 / 17 bits wide), and the same call with a bound inside the limit gets past the check."""
    import ctypes as C
    from dexterous_rl_manipulation_b200 import _lib
    L = _lib.lib()
    st = _lib.DexsimState()
    st.n, st.ld = 0, 32                                   # n == 0: a valid call returns before touching the device
    for name, _ in _lib.DexsimState._fields_[2:]:
        setattr(st, name, 0x1000)                         # aligned dummies, never dereferenced
    p = _lib.DexsimParams()
    p.reward_type, p.success_threshold, p.num_groups = 1, 3, 1
    rio = _lib.DexsimRolloutIO()
    grp = (C.c_char * C.sizeof(_lib.DexsimGroup))()
    for max_steps, loop, expect in ((200, 0, 0), (5241, 0, 0), (5242, 0, -1004), (100000, 5242, 0), (100000, 5243, -1004),
                                    (100000, 0, -1004)):
        p.max_episode_steps, p.loop_max_steps = max_steps, loop
        rc = L.dexsim_rollout(C.byref(st), C.byref(p), C.addressof(grp), None, 10, _lib.POLICY_RANDOM, C.byref(rio), None)
        assert rc == expect, (max_steps, loop, rc)
    st.ep_return = st.ep_stats = None                     # untracked: no limit
    p.max_episode_steps, p.loop_max_steps = 100000, 0
    assert L.dexsim_rollout(C.byref(st), C.byref(p), C.addressof(grp), None, 10, _lib.POLICY_RANDOM, C.byref(rio), None) == 0


def test_curriculum_driver_tail_progresses_at_most_once_per_episode():
    """BatchedCurriculumDriver.feed: when the sequential budget is spent the bulk tail may not advance more levels than
    it holds episodes (the reference progresses at most once per update() call), and exact totals are kept."""
    from oracle import ref_harness
    CC = dx.CurriculumConfig
    kw = dict(success_rate_threshold=0.3, window_size=5, min_episodes_before_progression=5, progression_steps=10)
    scheds = [dx.CurriculumScheduler(CC.easy(), CC.hard(), **kw)]
    if ref_harness.available():
        R = ref_harness.load()
        scheds.append(R.CurriculumScheduler(R.CurriculumConfig.easy(), R.CurriculumConfig.hard(), **kw))

    class _Env:
        curriculum_config = None
    for sched in scheds:
        drv = dx.BatchedCurriculumDriver(_Env(), sched, max_sequential_updates=8)
        got = drv.feed(episodes=10, successes=10, steps=100)   # 8 sequential updates, then a tail of 2 episodes
        # episodes 5..8 progress one level each (sequential).  The reference's scheduler exposes _should_progress /
        # _progress, so the tail of 2 episodes advances at most one level each: 6, exactly what replaying all 10
        # episodes one by one gives (an unbounded tail loop would have climbed to level 1.0); the stand-in has no such
        # hooks and catches up at the next feed.
        assert got == (6 if hasattr(sched, "_should_progress") else 4), type(sched)
        assert sched.total_episodes == 10 and sched.total_steps == 100
        assert (drv.exact_episodes, drv.exact_successes) == (10, 10)


def test_host_contact_row_expansion():
    """dexsim_expand_contact_rows (host code, AVX2 or scalar): rows 40-44 of a [45, ld] observation from 1-byte contact
    masks, for aligned and unaligned buffers and ragged sizes; nothing outside those rows / beyond n is touched."""
    import ctypes as C
    from dexterous_rl_manipulation_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(3)
    for n, ld, offset in ((1, 32, 0), (37, 64, 0), (1000, 1024, 0), (1000, 1024, 1), (65536 + 5, 65568, 0)):
        raw = np.full(45 * ld + 16, -7.0, np.float32)
        obs = raw[offset:offset + 45 * ld].reshape(45, ld)             # offset 1: base not 32-byte aligned -> scalar path
        mask = rng.integers(0, 32, ld).astype(np.uint8)
        rc = L.dexsim_expand_contact_rows(obs.ctypes.data, mask.ctypes.data, n, ld)
        assert rc == 0
        for f in range(5):
            assert np.array_equal(obs[40 + f, :n], ((mask[:n] >> f) & 1).astype(np.float32)), (n, f)
            assert np.all(obs[40 + f, n:] == -7.0)
        assert np.all(obs[:40] == -7.0)
    assert L.dexsim_expand_contact_rows(None, None, 1, 32) == -1001
