"""Hand-computed cases for oracle/windows.py (the reference class cannot run here: it needs
nimblephysics + .b3d files; /root/reference/src/data/AddBiomechanicsDataset.py:121-285)."""
import numpy as np

from oracle import windows as ow


def _subject(lengths, missing_idx, D=2, hist=6, nb=2, order=(0, 1), mass=50.0):
    trials = []
    rng = np.random.default_rng(0)
    for L, miss in zip(lengths, missing_idx):
        m = np.zeros(L, dtype=bool)
        m[list(miss)] = True
        tr = {"missing": m}
        for k, c in ow.input_widths(D, hist).items():
            tr[k] = rng.standard_normal((L, c))
        tr["tau"] = rng.standard_normal((L, D))
        tr["residualWrenchInRootFrame"] = rng.standard_normal((L, 6))
        tr["comAccInRootFrame"] = rng.standard_normal((L, 3))
        tr["groundContactWrenchesInRootFrame"] = rng.standard_normal((L, 6 * nb))
        tr["groundContactCenterOfPressureInRootFrame"] = rng.standard_normal((L, 3 * nb))
        tr["groundContactTorqueInRootFrame"] = rng.standard_normal((L, 3 * nb))
        tr["groundContactForceInRootFrame"] = rng.standard_normal((L, 3 * nb))
        trials.append(tr)
    return {"mass": mass, "contact_indices": list(order), "trials": trials}


def test_enumeration_rule():
    # T=4, stride 2: candidates ws in range(max(L-5,0)); window samples frames ws, ws+2.
    s = _subject([8, 5, 3], [[3], [], []])
    w = ow.enumerate_windows([s], 4, 2)
    # trial 0: L=8 → ws∈{0,1,2}; frames {0,2},{1,3},{2,4}; frame 3 missing kills ws=1 only.
    # trial 1: L=5 → range(0) → none.  trial 2: L=3 → none.
    assert w == [(0, 0, 0), (0, 0, 2)]
    # stride 1 looks at every frame in [ws, ws+4): ws=0 {0..3} hit, ws=1 hit, ws=2 {2..5} hit
    assert ow.enumerate_windows([s], 4, 1) == []


def test_enumeration_order_subject_trial_start():
    a = _subject([7], [[]])
    b = _subject([7, 7], [[], [0]])
    w = ow.enumerate_windows([a, b], 4, 4)      # one sample per window: frame ws only
    assert w == [(0, 0, 0), (0, 0, 1), (1, 0, 0), (1, 0, 1), (1, 1, 1)]


def test_sampler_and_batches():
    assert ow.sampler_indices(10, 4, 1) == [1, 5]           # first 8 only, r::W
    assert ow.sampler_indices(10, 4, 3) == [3, 7]
    assert ow.sampler_indices(3, 4, 0) == []
    assert ow.batches(list(range(5)), 2) == [[0, 1], [2, 3], [4]]   # partial batch kept


def test_get_window_and_pack():
    s = _subject([20], [[]], order=(1, 0), mass=80.0)
    T, st = 6, 2
    inputs, labels = ow.get_window([s], (0, 0, 3), T, st, "all_frames", 2)
    rows = [3, 5, 7]
    tr = s["trials"][0]
    assert np.array_equal(inputs["pos"], tr["pos"][rows].astype(np.float32))
    # contact bodies re-ordered: dataset body 0 ← subject body 1, and divided by fp32 mass
    f = tr["groundContactForceInRootFrame"][rows].astype(np.float32)
    exp = np.concatenate([f[:, 3:6], f[:, 0:3]], axis=1) / np.float32(80.0)
    assert np.array_equal(labels["groundContactForceInRootFrame"], exp)
    c = tr["groundContactCenterOfPressureInRootFrame"][rows].astype(np.float32)
    assert np.array_equal(labels["groundContactCenterOfPressureInRootFrame"], np.concatenate([c[:, 3:6], c[:, 0:3]], 1))
    # last_frame keeps only the final sampled frame
    _, lab1 = ow.get_window([s], (0, 0, 3), T, st, "last_frame", 2)
    assert lab1["tau"].shape == (1, 2) and np.array_equal(lab1["tau"][0], tr["tau"][7].astype(np.float32))
    # absent contact body → zeros
    s2 = _subject([20], [[]], order=(0, -1))
    _, lab2 = ow.get_window([s2], (0, 0, 0), T, st, "all_frames", 2)
    assert np.all(lab2["groundContactWrenchesInRootFrame"][:, 6:] == 0)
    # pack order and frame-major flatten
    x = ow.pack_inputs(inputs, flatten=False)
    assert x.shape == (3, 3 * 2 + 12 + 36 + 12)
    assert np.array_equal(x[:, :2], inputs["pos"]) and np.array_equal(x[:, 6:9], inputs["rootLinearVelInRootFrame"])
    assert np.array_equal(x[:, 9:12], inputs["rootAngularVelInRootFrame"])
    assert np.array_equal(ow.pack_inputs(inputs, flatten=True), x.reshape(-1))
    l30 = ow.pack_labels30(labels)
    assert l30.shape == (3, 30) and np.array_equal(l30[:, 6:12], labels["groundContactForceInRootFrame"])


def test_synthetic_subjects_are_deterministic():
    a = ow.make_synthetic_subjects(5, 4, 20)
    b = ow.make_synthetic_subjects(5, 4, 20)
    assert ow.enumerate_windows(a, 20, 5) == ow.enumerate_windows(b, 20, 5)
    assert len(ow.enumerate_windows(a, 20, 5)) > 0
