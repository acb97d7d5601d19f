"""modules/track.py:72-137 + utilities/counting/: `VideoCounting` mirror against the reference's own class (live, when the
reference tree is mounted) and against a committed golden CSV made from it (tests/golden/counting_golden.csv, generator:
`python tests/test_counting_cpu.py --make-golden`).  Every CSV column except the random `color` must be identical."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
ZONE = os.path.join(GOLD, "counting_zone.json")        # a copy of the layout of demo/sample/cam_04.json: zone polygon + two directions


def _stream(seed=3, n_tracks=40, n_frames=80):
    """(frame_id, track_id, label, int xyxy box) rows like CountingPipeline.run collects them (modules/__init__.py:80-84):
    tracks crossing, grazing and missing the zone; a few boxes with a corner exactly on a polygon vertex / edge."""
    rng = np.random.default_rng(seed)
    frames, tracks, labels, boxes = [], [], [], []
    start = rng.uniform([0, 100], [1280, 720], (n_tracks, 2)); vel = rng.uniform(-9, 9, (n_tracks, 2)); size = rng.uniform(20, 140, (n_tracks, 2))
    lab = rng.integers(0, 3, n_tracks)
    t0 = rng.integers(1, n_frames // 2, n_tracks); life = rng.integers(3, n_frames, n_tracks)
    for f in range(1, n_frames + 1):
        for k in range(n_tracks):
            if t0[k] <= f < t0[k] + life[k]:
                p = start[k] + vel[k] * (f - t0[k])
                frames.append(f); tracks.append(int(k % 17 + 1)); labels.append(int(lab[k]))
                boxes.append(np.array([int(p[0]), int(p[1]), int(p[0] + size[k, 0]), int(p[1] + size[k, 1])]))
    import json
    zone = json.load(open(ZONE))["shapes"][0]["points"]
    vx, vy = zone[1]
    frames += [n_frames + 1, n_frames + 2]; tracks += [99, 99]; labels += [1, 1]
    boxes += [np.array([int(vx) - 30, int(vy) - 30, int(vx), int(vy)]), np.array([0, 0, 10, 10])]
    return frames, tracks, labels, boxes


def _csv_frame(path):
    import pandas as pd
    return pd.read_csv(path).drop(columns=["color"])


def test_counting_matches_golden_csv(tmp_path):
    from vehicle_counting_b200.modules.track import VideoCounting
    out = tmp_path / "ours.csv"
    vc = VideoCounting(["a", "b", "c"], ZONE)
    vc.run(*_stream(), output_path=str(out))
    got, want = _csv_frame(out), _csv_frame(os.path.join(GOLD, "counting_golden.csv"))
    assert len(want) > 200
    assert list(got.columns) == list(want.columns)
    assert got.equals(want)


def test_counting_matches_live_reference(tmp_path):
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference tree not present")
    ref_shim.install()
    from modules.track import VideoCounting as RefCounting
    from vehicle_counting_b200.modules.track import VideoCounting
    for seed in (3, 4, 5):
        a, b = tmp_path / f"ref{seed}.csv", tmp_path / f"ours{seed}.csv"
        RefCounting(["a", "b", "c"], ZONE).run(*_stream(seed), output_path=str(a))
        VideoCounting(["a", "b", "c"], ZONE).run(*_stream(seed), output_path=str(b))
        assert _csv_frame(b).equals(_csv_frame(a)), seed


def test_point_in_polygon_edge_cases():
    from vehicle_counting_b200.counting import check_bbox_intersect_polygon, point_in_polygon
    sq = [[0, 0], [10, 0], [10, 10], [0, 10]]
    assert point_in_polygon(sq, (5, 5)) and not point_in_polygon(sq, (15, 5)) and not point_in_polygon(sq, (5, -1))
    assert point_in_polygon(sq, (10, 10)) and point_in_polygon(sq, (0, 5))          # vertex and edge count as inside
    assert check_bbox_intersect_polygon(sq, (8, 8, 20, 20)) and not check_bbox_intersect_polygon(sq, (11, 11, 20, 20))
    assert not check_bbox_intersect_polygon(sq, (-5, -5, 15, 15))                   # zone strictly inside the box: no corner inside


if __name__ == "__main__" and "--make-golden" in sys.argv:
    from oracle import ref_shim
    assert ref_shim.available()
    ref_shim.install()
    from modules.track import VideoCounting as RefCounting
    RefCounting(["a", "b", "c"], ZONE).run(*_stream(), output_path=os.path.join(GOLD, "counting_golden.csv"))
    print("wrote", os.path.join(GOLD, "counting_golden.csv"))
