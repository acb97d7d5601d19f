"""CPU tests of the host association step: integer output rows are bit-identical to the reference's
DeepSort.update when both are fed the same detections and embeddings."""
import os

import numpy as np
import pytest

from oracle import ref_shim

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mk(max_dist=0.2, min_confidence=0.3, nms_max_overlap=0.5, max_iou_distance=0.7, max_age=70, n_init=3, nn_budget=100):
    """Our DeepSort without constructing the GPU extractor (features are injected)."""
    from vehicle_counting_b200.networks.deepsort.deep_sort import DeepSort
    from vehicle_counting_b200.networks.deepsort.sort import Gallery, Tracker
    ds = DeepSort.__new__(DeepSort)
    ds.min_confidence, ds.nms_max_overlap = min_confidence, nms_max_overlap
    ds.extractor = None
    ds.tracker = Tracker(Gallery(max_dist, nn_budget), max_iou_distance=max_iou_distance, max_age=max_age, n_init=n_init)
    return ds


def _rows(out):
    return np.asarray(out, dtype=np.int64).reshape(-1, 7) if len(out) else np.zeros((0, 7), np.int64)


def test_tracker_reproduces_reference_golden_rows():
    z = np.load(os.path.join(GOLD, "deepsort_golden.npz"))
    ds = _mk()
    off = 0
    for t in range(z["boxes"].shape[0]):
        got = _rows(ds.update(z["boxes"][t].copy(), z["conf"].copy(), z["frame"], features=z["feats"][t]))
        want = z["rows"][off:off + z["row_counts"][t]]
        off += z["row_counts"][t]
        np.testing.assert_array_equal(got, want)


def _scenario(seed, T=60, H=480, W=640, n_obj=14):
    """Objects entering/leaving, misses, score noise, near-duplicate boxes (exercises host NMS), identity
    embeddings = noisy per-object prototypes."""
    rng = np.random.default_rng(seed)
    proto = rng.normal(size=(n_obj, 512)).astype(np.float32)
    start = rng.integers(0, T // 2, n_obj); life = rng.integers(8, T, n_obj)
    pos = rng.uniform([0, 0], [W - 120, H - 120], (n_obj, 2)); vel = rng.uniform(-6, 6, (n_obj, 2))
    size = rng.uniform(30, 110, (n_obj, 2))
    frames = []
    for t in range(T):
        boxes, conf, feats = [], [], []
        for o in range(n_obj):
            if not (start[o] <= t < start[o] + life[o]) or rng.random() < 0.12:
                continue
            p = pos[o] + vel[o] * (t - start[o]) + rng.normal(0, 1.0, 2)
            b = np.array([p[0], p[1], p[0] + size[o, 0], p[1] + size[o, 1]])
            f = proto[o] + rng.normal(0, 0.35, 512).astype(np.float32)
            f = f / np.linalg.norm(f)
            boxes.append(b); conf.append(rng.uniform(0.2, 0.99)); feats.append(f)
            if rng.random() < 0.1:                      # a near-duplicate, lower score
                boxes.append(b + rng.normal(0, 1.5, 4)); conf.append(conf[-1] * 0.8); feats.append(f)
        frames.append((np.array(boxes, np.float64).reshape(-1, 4), np.array(conf), np.array(feats, np.float32).reshape(-1, 512)))
    return frames, np.zeros((H, W, 3), np.uint8)


@pytest.mark.parametrize("seed,cfg", [(0, dict()), (1, dict(max_age=5, n_init=2, nn_budget=3)), (2, dict(max_dist=0.6, nms_max_overlap=1.0)),
                                      (3, dict(max_age=30, max_iou_distance=0.9))])
def test_tracker_matches_live_reference(seed, cfg):
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    base = dict(max_dist=0.2, min_confidence=0.3, nms_max_overlap=0.5, max_iou_distance=0.7, max_age=70, n_init=3, nn_budget=100)
    base.update(cfg)
    ref = ref_shim.reference_deepsort(**base)
    ours = _mk(**base)
    frames, img = _scenario(seed)
    total = 0
    for t, (boxes, conf, feats) in enumerate(frames):
        if len(boxes) == 0:
            continue                                     # the reference pipeline skips empty frames (modules/__init__.py:68-69)
        ref._get_features = lambda bbox_xywh, ori_img, f=feats: f          # same embeddings into both
        want = _rows(ref.update(boxes.copy(), conf.copy(), img))
        got = _rows(ours.update(boxes.copy(), conf.copy(), img, features=feats))
        np.testing.assert_array_equal(got, want, err_msg=f"frame {t}")
        total += len(want)
    assert total > 50
    assert [t.track_id for t in ours.tracker.tracks] == [t.track_id for t in ref.tracker.tracks]


def test_tracker_matches_live_reference_dense_and_is_faster():
    """BASELINE config 3 density: 64 objects per frame over 100 frames (gallery budget reached, cascade levels in use).  Rows stay
    bit-identical to the reference's DeepSort.update; the batched host step (gating, Kalman update, cached gallery) must also be
    faster than the reference's per-track loop (measured: 3.2x here, 15x with 64 simultaneous tracks: 5.3 vs 82 ms per frame)."""
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    import time
    base = dict(max_dist=0.2, min_confidence=0.3, nms_max_overlap=0.5, max_iou_distance=0.7, max_age=70, n_init=3, nn_budget=100)
    ref = ref_shim.reference_deepsort(**base)
    ours = _mk(**base)
    frames, img = _scenario(7, T=100, H=720, W=1280, n_obj=64)
    t_ref = t_ours = 0.0
    total = 0
    for t, (boxes, conf, feats) in enumerate(frames):
        if len(boxes) == 0:
            continue
        ref._get_features = lambda bbox_xywh, ori_img, f=feats: f
        t0 = time.perf_counter()
        want = _rows(ref.update(boxes.copy(), conf.copy(), img))
        t1 = time.perf_counter()
        got = _rows(ours.update(boxes.copy(), conf.copy(), img, features=feats))
        t2 = time.perf_counter()
        if t >= 20:
            t_ref += t1 - t0; t_ours += t2 - t1
        np.testing.assert_array_equal(got, want, err_msg=f"frame {t}")
        total += len(want)
    assert total > 1000
    assert t_ours < t_ref, (t_ours, t_ref)          # a loose bar on purpose (shared CI hosts); measured ratios are in the docstring


def test_host_nms_matches_reference():
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.install()
    from networks.deepsort.sort.preprocessing import non_max_suppression as ref_nms
    from vehicle_counting_b200.networks.deepsort.sort import host_nms
    rng = np.random.default_rng(0)
    for n in (1, 5, 60):
        tl = rng.uniform(0, 200, (n, 2)); wh = rng.uniform(10, 90, (n, 2))
        boxes = np.concatenate([tl, wh], 1); scores = rng.uniform(0, 1, n)
        for thr in (0.3, 0.5, 1.0):
            assert host_nms(boxes, thr, scores) == [int(i) for i in ref_nms(boxes, thr, scores)]


def test_batched_kalman_and_gating_equal_the_per_track_formulas():
    """kf_update_batch / kf_gating_distance_batch against the per-track restatements of kalman_filter.py:154-229 (scipy cho_solve /
    solve_triangular): same numbers to the last few bits on random positive-definite states."""
    from vehicle_counting_b200.networks.deepsort import sort as S
    rng = np.random.default_rng(5)
    T, D = 23, 17
    means = np.concatenate([rng.uniform(50, 600, (T, 2)), rng.uniform(0.3, 2.0, (T, 1)), rng.uniform(40, 300, (T, 1)), rng.normal(0, 2, (T, 4))], 1)
    A = rng.normal(size=(T, 8, 8))
    covs = A @ A.transpose(0, 2, 1) + 5.0 * np.eye(8)
    meas = np.concatenate([rng.uniform(50, 600, (D, 2)), rng.uniform(0.3, 2.0, (D, 1)), rng.uniform(40, 300, (D, 1))], 1)
    g = S.kf_gating_distance_batch(means, covs, meas)
    for t in range(T):
        np.testing.assert_allclose(g[t], S.kf_gating_distance(means[t], covs[t], meas), rtol=1e-12, atol=0)
    zs = meas[rng.integers(0, D, T)]
    nm, nc = S.kf_update_batch(means, covs, zs)
    for t in range(T):
        m1, c1 = S.kf_update(means[t], covs[t], zs[t])
        np.testing.assert_allclose(nm[t], m1, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(nc[t], c1, rtol=1e-11, atol=1e-11)


def test_gallery_cache_equals_the_reference_expression():
    """Gallery.distance with cached normalised rows in a compacting buffer == min over (1 - a_n @ b_n^T) on the raw sample lists
    (nn_matching.py:33-52, :137-177), bit for bit, across budget overflow, identities leaving and an unbounded gallery."""
    from vehicle_counting_b200.networks.deepsort.sort import Gallery
    rng = np.random.default_rng(6)
    for budget in (3, 7, None):
        gal = Gallery(0.2, budget)
        ids = [1, 2, 5, 9]
        for step in range(40):
            feats = [rng.normal(size=64).astype(np.float32) for _ in range(6)]
            targets = [ids[i] for i in rng.integers(0, len(ids), 6)]
            active = ids if step < 25 else ids[:3]
            gal.partial_fit(feats, targets, [t for t in active if t in gal.samples or t in targets])
            q = rng.normal(size=(5, 64)).astype(np.float32)
            have = [t for t in active if t in gal.samples]
            if not have:
                continue
            got = gal.distance(q, have)
            bn = q / np.linalg.norm(q, axis=1, keepdims=True)
            for i, t in enumerate(have):
                a = np.asarray(gal.samples[t])
                a = a / np.linalg.norm(a, axis=1, keepdims=True)
                want = (1.0 - a @ bn.T).min(axis=0)
                np.testing.assert_array_equal(got[i], want.astype(np.float64))
