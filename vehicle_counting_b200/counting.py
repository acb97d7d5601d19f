"""Host post-processing of a finished video, restated from the reference (/root/reference/utilities/counting/):
zone annotation loader (utils.py:128-137), "a box counts when one of its corners is inside the zone polygon"
(bb_polygon.py:13-114, ray casting with the on-edge special cases), movement direction by cosine similarity
(utils.py:139-152, bb_polygon.py:117-124) and the tracking CSV (utils.py:154-198).  Once per video, pure host work; it is
here so that `modules.track.VideoCounting` completes the reference's `modules.track` surface."""
from __future__ import annotations

import json
from typing import Dict, List, Sequence, Tuple

import numpy as np

# a fixed BGR palette: the reference draws a random colour per track from a webcolors table (track.py:111); the column is
# non-deterministic there and is not part of the parity artefact
PALETTE: List[Tuple[int, int, int]] = [(255, 56, 56), (255, 157, 151), (255, 112, 31), (255, 178, 29), (207, 210, 49), (72, 249, 10),
                                       (146, 204, 23), (61, 219, 134), (26, 147, 52), (0, 212, 187), (44, 153, 168), (0, 194, 255),
                                       (52, 69, 147), (100, 115, 255), (0, 24, 236), (132, 56, 255), (82, 0, 133), (203, 56, 255),
                                       (255, 149, 200), (255, 55, 199)]


def load_zone_anno(zone_path: str):
    """labelme JSON: the first shape is the zone polygon, shapes labelled 'direction..' are (start, end) vectors keyed by
    the last two characters of their label."""
    with open(zone_path, "r") as f:
        anno = json.load(f)
    zone = anno["shapes"][0]["points"]
    directions = {sh["label"][-2:]: sh["points"] for sh in anno["shapes"] if sh["label"].startswith("direction")}
    return zone, directions


def _orient(p, q, r) -> int:
    """0 collinear, 1 clockwise, 2 counter-clockwise (sign of the cross product, exact comparison as in the reference)."""
    v = (q[1] - p[1]) * (r[0] - q[0]) - (q[0] - p[0]) * (r[1] - q[1])
    return 0 if v == 0 else (1 if v > 0 else 2)


def _within_box(p, q, r) -> bool:
    return min(p[0], r[0]) <= q[0] <= max(p[0], r[0]) and min(p[1], r[1]) <= q[1] <= max(p[1], r[1])


def _segments_cross(p1, q1, p2, q2) -> bool:
    o1, o2, o3, o4 = _orient(p1, q1, p2), _orient(p1, q1, q2), _orient(p2, q2, p1), _orient(p2, q2, q1)
    if o1 != o2 and o3 != o4:
        return True
    return ((o1 == 0 and _within_box(p1, p2, q1)) or (o2 == 0 and _within_box(p1, q2, q1)) or
            (o3 == 0 and _within_box(p2, p1, q2)) or (o4 == 0 and _within_box(p2, q1, q2)))


def point_in_polygon(polygon: Sequence[Sequence[float]], point: Sequence[float]) -> bool:
    """Ray casting along +y to (x, 1e9) exactly as the reference does (its 'extreme' point keeps x and moves y), a point on an
    edge decided by that edge."""
    far = [point[0], 1e9]
    n = len(polygon)
    crossings = 0
    for i in range(n):
        a, b = polygon[i], polygon[(i + 1) % n]
        if _segments_cross(a, b, point, far):
            if _orient(a, point, b) == 0:
                return _within_box(a, point, b)
            crossings += 1
    return crossings % 2 == 1


def check_bbox_intersect_polygon(polygon, bbox) -> bool:
    x1, y1, x2, y2 = bbox
    return any(point_in_polygon(polygon, c) for c in ((x1, y1), (x2, y1), (x2, y2), (x1, y2)))


def cosine_similarity_2d(a2d, b2d) -> float:
    a = np.array((a2d[1][0] - a2d[0][0], a2d[1][1] - a2d[0][1])).astype(float)
    b = np.array((b2d[1][0] - b2d[0][0], b2d[1][1] - b2d[0][1])).astype(float)
    return np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b * 1.0))


def find_best_match_direction(obj_vector, paths: Dict[str, Sequence[Sequence[float]]]) -> str:
    keys = list(paths.keys())
    best, best_score = keys[0], 0
    for k in keys:
        s = cosine_similarity_2d(obj_vector, paths[k])
        if s > best_score:
            best, best_score = k, s
    return best


CSV_COLUMNS = ["track_id", "frame_id", "box", "color", "label", "direction", "fpoint", "lpoint", "fframe", "lframe"]


def tracking_rows(track_dict: List[dict]) -> Dict[str, list]:
    """column -> values, rows grouped by label, then track (insertion order), then time (utils.py:169-195)"""
    cols: Dict[str, list] = {c: [] for c in CSV_COLUMNS}
    for label_id, per_label in enumerate(track_dict):
        for track_id, rec in per_label.items():
            boxes, frames = rec["boxes"], rec["frames"]
            b0, b1 = boxes[0], boxes[-1]
            first = ((b0[2] + b0[0]) / 2, (b0[3] + b0[1]) / 2)
            last = ((b1[2] + b1[0]) / 2, (b1[3] + b1[1]) / 2)
            for box, frame in zip(boxes, frames):
                cols["track_id"].append(track_id)
                cols["frame_id"].append(frame)
                cols["box"].append(box.tolist())
                cols["color"].append(rec["color"])
                cols["label"].append(label_id)
                cols["direction"].append(rec["direction"])
                cols["fpoint"].append(first)
                cols["lpoint"].append(last)
                cols["fframe"].append(frames[0])
                cols["lframe"].append(frames[-1])
    return cols


def save_tracking_to_csv(track_dict: List[dict], filename: str) -> None:
    import pandas as pd
    pd.DataFrame(tracking_rows(track_dict)).to_csv(filename, index=False)
