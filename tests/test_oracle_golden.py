"""The C restatement (oracle/) against fixtures produced by the UNMODIFIED reference.

Bars: observations, flags, contact counts and labels bit-exact; rewards to 1e-14 relative
(numpy's SIMD exp and glibc's exp differ by at most 1 ulp in float64)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle


def _load(golden_dir, name):
    with np.load(os.path.join(golden_dir, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}          # materialise once (NpzFile re-inflates per access)


def test_struct_layouts_match_c():
    lib = oracle.lib()
    assert lib.dexo_sizeof_env() == oracle.ENV_DTYPE.itemsize
    assert lib.dexo_sizeof_group() == oracle.GROUP_DTYPE.itemsize


def test_trajectories_bit_exact(golden_dir):
    g = _load(golden_dir, "traj.npz")
    n_traj, T = g["actions"].shape[:2]
    batches = {}
    for i in range(n_traj):
        if g["keep_pos"][i]:
            ob = batches[int(g["chain"][i])]          # reused env object: keeps the previous position
            obs0 = ob.reset_predrawn(g["jp0"][i], g["size"][i], g["mass"][i], g["friction"][i], None)
        else:
            ob = oracle.OracleBatch(1, dense=bool(g["dense"][i]), max_episode_steps=int(g["max_steps"][i]),
                                    weights=tuple(g["weights"][i]))
            obs0 = ob.reset_predrawn(g["jp0"][i], g["size"][i], g["mass"][i], g["friction"][i], g["pos"][i])
        batches[i] = ob
        assert np.array_equal(obs0[0], g["obs"][i, 0]), g["names"][i]
        for t in range(T):
            obs, rew, comps, te, tr, nc = ob.step(g["actions"][i, t])
            assert np.array_equal(obs[0], g["obs"][i, t + 1]), (g["names"][i], t)
            assert te[0] == g["terminated"][i, t] and tr[0] == g["truncated"][i, t], (g["names"][i], t)
            assert nc[0] == g["num_contacts"][i, t]
            assert rew[0] == pytest.approx(g["reward"][i, t], rel=1e-14, abs=1e-15)
            assert np.allclose(comps[0], g["comps"][i, t], rtol=1e-14, atol=0)
        assert np.array_equal(ob.env["op"][0], g["op_final"][i])      # float64 object position


def test_noise_wrapper_trajectories(golden_dir):
    g = _load(golden_dir, "noise.npz")
    n, T = g["actions"].shape[:2]
    for i in range(n):
        ob = oracle.OracleBatch(1, dense=True, max_episode_steps=int(g["max_steps"]))
        obs0 = ob.reset_predrawn(g["jp0"][i], g["size"][i], g["mass"][i], g["friction"][i], g["pos"][i])
        so, sd = g["sigma_obs"][i], g["sigma_dyn"][i]
        obs0 = obs0[0] + g["obs_noise"][i, 0] if so > 0 else obs0[0]
        assert np.array_equal(obs0, g["obs"][i, 0])
        for t in range(T):
            obs, rew, _, te, tr, nc = ob.step(g["actions"][i, t],
                                              dyn_noise=g["dyn_noise"][i, t] if sd > 0 else None,
                                              obs_noise=g["obs_noise"][i, t + 1] if so > 0 else None)
            assert np.array_equal(obs[0], g["obs"][i, t + 1]), (i, t)
            assert te[0] == g["terminated"][i, t] and tr[0] == g["truncated"][i, t]
            assert nc[0] == g["num_contacts"][i, t]
            assert rew[0] == pytest.approx(g["reward"][i, t], rel=1e-14, abs=1e-15)


def test_failure_labels_and_np_var(golden_dir):
    g = _load(golden_dir, "labels.npz")
    for i in range(g["length"].shape[0]):
        c = g["counts"][i, :g["length"][i]]
        a, b, conf = oracle.classify(g["success"][i], g["steps"][i], g["num"][i], g["final"][i], c,
                                     int(g["max_steps"]), 3)
        assert a == g["label_metrics"][i] and b == g["label_taxonomy"][i], i
        assert conf == pytest.approx(g["confidence"][i], abs=1e-15)
        if g["length"][i] > 1:
            assert oracle.np_var_counts(c) == g["var"][i]              # bit-exact pairwise summation


def test_reference_known_answers(golden_dir):
    with open(os.path.join(golden_dir, "anchors.json")) as fh:
        anchors = json.load(fh)
    for ka in anchors["known_answers"]:
        a, b, _ = oracle.classify(False, ka["episode_steps"], ka["num_contacts"], ka["final_contacts"],
                                  np.asarray(ka["counts"], np.uint8), 200, 3)
        if "metrics" in ka:
            assert oracle.LABELS_METRICS[a] in ka["metrics"], ka["src"]
        if "taxonomy" in ka:
            assert oracle.LABELS_TAXONOMY[b] in ka["taxonomy"], ka["src"]


def test_whole_episode_records(golden_dir):
    """dexo_rollout's episode bookkeeping against Evaluator.evaluate_episode / run_episode."""
    g = _load(golden_dir, "episodes.npz")
    prev = None
    for i in range(g["kind"].shape[0]):
        n_steps = int(g["n_steps"][i])
        if g["keep_pos"][i]:
            ob = prev
            ob.reset_predrawn(g["jp0"][i], g["size"][i], g["mass"][i], g["friction"][i], None)
        else:
            ob = oracle.OracleBatch(1, dense=bool(g["dense"][i]), max_episode_steps=int(g["max_steps"][i]))
            ob.reset_predrawn(g["jp0"][i], g["size"][i], g["mass"][i], g["friction"][i], g["pos"][i])
        snapshot = ob.env.copy()
        grp = oracle.make_group(object_size=float(g["size"][i]), object_mass=float(g["mass"][i]),
                                friction_coefficient=float(g["friction"][i]))
        evaluator_flow = g["kind"][i] == 0
        cnt, rs = ob.rollout(grp, n_steps, seed=1, policy_kind=0, respawn=True,
                             success_is_terminated=evaluator_flow, loop_max_steps=int(g["loop_max_steps"][i]),
                             actions=g["actions"][i, :n_steps][:, None, :])
        assert cnt[0, 0] == 1, "exactly one episode must finish on the last recorded step"
        assert cnt[0, 1] == int(g["success"][i])
        assert cnt[0, 2] == n_steps
        assert cnt[0, 3] == g["final_contacts"][i]
        assert rs[0, 0] == pytest.approx(g["episode_reward"][i], rel=1e-13)
        if evaluator_flow:
            la, lb = int(g["label_metrics"][i]), int(g["label_taxonomy"][i])
            exp_a = np.zeros(6, np.int64); exp_b = np.zeros(6, np.int64)
            if la >= 0: exp_a[la] = 1
            if lb >= 0: exp_b[lb] = 1
            assert np.array_equal(cnt[0, 4:10], exp_a) and np.array_equal(cnt[0, 10:16], exp_b)
        # replay without auto-reset so the next reused-env episode starts from the true final position
        ob.env[:] = snapshot
        for t in range(n_steps):
            ob.step(g["actions"][i, t])
        prev = ob


def test_scalar_anchors(golden_dir):
    with open(os.path.join(golden_dir, "anchors.json")) as fh:
        A = json.load(fh)
    a = np.full(15, -0.5, np.float32)
    # default config dense, reset(seed=0) state rebuilt from traj-independent anchors is not possible
    # without PCG64; the anchors are checked through the first default/dense trajectory's reset instead.
    assert A["truncated_flags_max5"] == [False] * 5 + [True] * 3
    ob = oracle.OracleBatch(1, dense=False, max_episode_steps=5)
    ob.reset_predrawn(np.zeros(15, np.float32), 0.05, 0.1, 0.5, np.array([0.05, 0.05, 0.1], np.float32))
    flags = [bool(ob.step(a)[4][0]) for _ in range(8)]
    assert flags == A["truncated_flags_max5"]


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    assert [hex(x) for x in oracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    out = oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    out = oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_rng_draw_ranges():
    grp = oracle.make_group(object_size_range=(0.03, 0.07), object_mass_range=(0.05, 0.15), friction_range=(0.3, 0.7))
    js, ss, ps = [], [], []
    for gid in range(400):
        jp0, s, m, f, pos = oracle.reset_draws(42, gid, 0, grp)
        js.append(jp0); ss.append((s, m, f)); ps.append(pos)
    js, ss, ps = np.asarray(js), np.asarray(ss), np.asarray(ps)
    assert js.min() >= -0.1 and js.max() <= 0.1 and abs(js.mean()) < 0.01
    assert ss[:, 0].min() >= 0.03 and ss[:, 0].max() <= 0.07 and ss[:, 2].min() >= 0.3 and ss[:, 2].max() <= 0.7
    assert ps[:, 2].min() >= 0.05 and ps[:, 2].max() <= 0.2 and np.abs(ps[:, :2]).max() <= 0.1
    a = np.asarray([oracle.policy_action(42, 3, 0, t, 1) for t in range(500)])
    assert a.min() >= -1 and a.max() < 1 and abs(a.mean()) < 0.03 and abs(a.std() - 1 / np.sqrt(3)) < 0.02
    h = np.asarray([oracle.policy_action(42, 3, 0, t, 2) for t in range(200)])
    assert h.min() >= -0.6001 and h.max() <= -0.3999


def test_simple_learner_rollout(golden_dir):
    """dexo_rollout_learner against run_episode + SimpleLearner of the unmodified reference (reused env:
    episodes after the first keep the object position; learner mean carries over, best resets)."""
    g = _load(golden_dir, "learner.npz")
    K = int(g["loop_max_steps"])
    for c in range(g["dense"].shape[0]):
        ob = oracle.OracleBatch(1, dense=bool(g["dense"][c]), max_episode_steps=200)
        grp = oracle.make_group(object_size=float(g["size"][c]), object_mass=float(g["mass"][c]),
                                friction_coefficient=float(g["friction"][c]))
        mean = np.zeros((1, 15), np.float32)
        best = np.full(1, -np.inf)
        for ep in range(g["steps"].shape[1]):
            n_steps = int(g["steps"][c, ep])
            ob.reset_predrawn(g["jp0"][c, ep], g["size"][c], g["mass"][c], g["friction"][c],
                              g["pos"][c, ep] if ep == 0 else None)
            best[:] = -np.inf
            cnt, rs = oracle.rollout_learner(ob, grp, n_steps, 0, mean, best, g["act_noise"][c, ep, :n_steps][:, None, :],
                                             g["upd_noise"][c, ep, :n_steps][:, None, :], loop_max_steps=K)
            assert cnt[0, 0] == 1 and cnt[0, 2] == n_steps and cnt[0, 3] == g["final_contacts"][c, ep]
            assert rs[0, 0] == pytest.approx(g["reward"][c, ep], rel=1e-13)
            # undo the auto-reset of the finished episode: replay is per episode, like run_episode
        assert np.array_equal(mean[0], g["final_mean"][c])


def test_oracle_vs_live_reference_on_extreme_inputs():
    """When the reference is importable: NaN / inf / denormal / out-of-workspace inputs through the UNMODIFIED env and
    the oracle side by side (the committed goldens only cover ordinary ranges)."""
    import warnings
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference not present")
    R = ref_harness.load()
    rng = np.random.default_rng(99)
    specials = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e30, -1e30, 3.4e38, 1.0000001], np.float32)
    sizes = [0.0, -0.05, 1e-300, 0.05, 0.2, 1e300, float("inf"), float("nan")]
    frics = [0.0, -3.0, 0.5, 1e6, 1e40, float("inf"), float("nan")]
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        for case in range(60):
            dense = bool(case % 2)
            size, fric = sizes[case % len(sizes)], frics[case % len(frics)]
            cfg = R.CurriculumConfig(object_size=size, object_mass=0.1, friction_coefficient=fric)
            pos = rng.uniform(-0.5, 0.5, 3).astype(np.float32)
            if case % 5 == 0:
                pos[rng.integers(0, 3)] = rng.choice(np.array([0.2, -0.2, 0.3, 1e30, np.inf, np.nan, -0.0], np.float32))
            env = R.DexterousManipulationEnv(reward_type="dense" if dense else "sparse", curriculum_config=cfg,
                                             max_episode_steps=8, object_position=pos)
            obs0, _ = env.reset(seed=case)
            ob = oracle.OracleBatch(1, dense=dense, max_episode_steps=8)
            o0 = ob.reset_predrawn(env.joint_positions.copy(), size, 0.1, fric, pos)
            assert np.array_equal(o0[0], obs0, equal_nan=True)
            for t in range(15):
                a = rng.uniform(-2, 2, 15).astype(np.float32)
                m = rng.random(15) < 0.1
                a[m] = rng.choice(specials, int(m.sum()))
                obs, r, te, tr, info = env.step(a)
                oo, orr, _, ote, otr, onc = ob.step(a)
                assert np.array_equal(oo[0], obs, equal_nan=True), (case, t)
                assert ote[0] == te and otr[0] == tr and onc[0] == info["num_contacts"], (case, t)
                assert (np.isnan(r) and np.isnan(orr[0])) or orr[0] == pytest.approx(r, rel=1e-14, abs=1e-15), (case, t, r, orr[0])
