"""Host-side SORT/DeepSORT association step (Kalman filter, gated appearance cascade, IoU fallback).

This is the sequential, stateful consumer of the GPU path's outputs; per BASELINE.json it stays on the
host.  It restates the behaviour of /root/reference/networks/deepsort/sort/ (tracker.py:50-139,
track.py:19-175, kalman_filter.py:23-229, linear_assignment.py:12-192, nn_matching.py:99-177,
iou_matching.py:7-81, preprocessing.py:6-73) with a different structure: all tracks of one tracker are
advanced through the Kalman prediction in one batched float64 einsum, the gallery of each identity is a
ring-buffered matrix, and gating distances for a whole cascade level come from one batched Cholesky
solve.  Decisions that determine integer outputs (detection order after NMS, order in which unmatched
detections become new identities, Hungarian cost clipping) follow the reference exactly so that, fed
the same detections and embeddings, the emitted rows are identical (tests/test_tracker_cpu.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular
from scipy.optimize import linear_sum_assignment

CHI2INV95_4DOF = 9.4877          # kalman_filter.py:11-20
INFTY_COST = 1e5                 # linear_assignment.py:9
STD_POS = 1.0 / 20               # kalman_filter.py:50-51
STD_VEL = 1.0 / 160

TENTATIVE, CONFIRMED, DELETED = 1, 2, 3

_F = np.eye(8)
_F[:4, 4:] = np.eye(4)           # constant-velocity model, dt = 1
_H = np.eye(4, 8)


class Detection:
    """tlwh box + confidence + appearance feature (sort/detection.py)."""
    __slots__ = ("tlwh", "confidence", "feature")

    def __init__(self, tlwh, confidence, feature):
        self.tlwh = np.asarray(tlwh, dtype=np.float64)
        self.confidence = float(confidence)
        self.feature = np.asarray(feature, dtype=np.float32)

    def to_xyah(self) -> np.ndarray:
        r = self.tlwh.copy()
        r[:2] += r[2:] / 2
        r[2] /= r[3]
        return r


class Track:
    __slots__ = ("mean", "covariance", "track_id", "hits", "age", "time_since_update", "state", "features",
                 "confidence_scores", "_n_init", "_max_age")

    def __init__(self, mean, covariance, track_id, n_init, max_age, feature, confidence):
        self.mean, self.covariance = mean, covariance
        self.track_id = track_id
        self.hits, self.age, self.time_since_update = 1, 1, 0
        self.state = TENTATIVE
        self.features = [feature] if feature is not None else []
        self.confidence_scores = [confidence] if confidence is not None else []
        self._n_init, self._max_age = n_init, max_age

    def to_tlwh(self) -> np.ndarray:
        r = self.mean[:4].copy()
        r[2] *= r[3]
        r[:2] -= r[2:] / 2
        return r

    def is_confirmed(self) -> bool:
        return self.state == CONFIRMED

    def is_tentative(self) -> bool:
        return self.state == TENTATIVE

    def is_deleted(self) -> bool:
        return self.state == DELETED

    def get_confidence_score(self):
        return self.confidence_scores[-1] if self.confidence_scores else -1

    def get_features(self):
        return self.features[-1] if self.features else -1


# ---------------------------------------------------------------------------------------------
# Kalman filter, batched over tracks
# ---------------------------------------------------------------------------------------------
def kf_initiate(xyah: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    h = xyah[3]
    mean = np.r_[xyah, np.zeros(4)]
    std = [2 * STD_POS * h, 2 * STD_POS * h, 1e-2, 2 * STD_POS * h,
           10 * STD_VEL * h, 10 * STD_VEL * h, 1e-5, 10 * STD_VEL * h]
    return mean, np.diag(np.square(std))


def kf_predict_batch(means: np.ndarray, covs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """means [T,8], covs [T,8,8] -> predicted; process noise scales with each track's height."""
    h = means[:, 3]
    q = np.empty((means.shape[0], 8))
    q[:, [0, 1, 3]] = (STD_POS * h)[:, None]
    q[:, 2] = 1e-2
    q[:, [4, 5, 7]] = (STD_VEL * h)[:, None]
    q[:, 6] = 1e-5
    new_means = means @ _F.T
    new_covs = _F @ (covs @ _F.T)           # same association as np.linalg.multi_dot picks for square factors
    idx = np.arange(8)
    new_covs[:, idx, idx] += np.square(q)
    return new_means, new_covs


def kf_project(mean: np.ndarray, cov: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    h = mean[3]
    r = np.square([STD_POS * h, STD_POS * h, 1e-1, STD_POS * h])
    return _H @ mean, _H @ cov @ _H.T + np.diag(r)


def kf_update(mean: np.ndarray, cov: np.ndarray, z: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    pm, pc = kf_project(mean, cov)
    chol = cho_factor(pc, lower=True, check_finite=False)
    gain = cho_solve(chol, (cov @ _H.T).T, check_finite=False).T          # [8, 4] = P H^T S^-1
    new_mean = mean + (z - pm) @ gain.T
    new_cov = cov - gain @ (pc @ gain.T)
    return new_mean, new_cov


def kf_update_batch(means: np.ndarray, covs: np.ndarray, zs: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """kalman_filter.py:154-186 for M matched tracks at once: means [M,8], covs [M,8,8], measurements zs [M,4].  The 4x4
    innovation covariance is factored by batched LAPACK potrf (as cho_factor does per track) and the two triangular solves of
    cho_solve are written out over the four rows; the matrix products keep the reference's association."""
    h = means[:, 3]
    pm = means[:, :4]
    pc = covs[:, :4, :4].copy()
    r = np.square(STD_POS * h)
    pc[:, 0, 0] += r
    pc[:, 1, 1] += r
    pc[:, 2, 2] += np.square(1e-1)
    pc[:, 3, 3] += r
    L = np.linalg.cholesky(pc)                                   # [M,4,4] lower
    B = covs[:, :, :4].transpose(0, 2, 1)                        # (P H^T)^T  [M,4,8]
    # forward: L Y = B
    y0 = B[:, 0] / L[:, 0, 0, None]
    y1 = (B[:, 1] - L[:, 1, 0, None] * y0) / L[:, 1, 1, None]
    y2 = (B[:, 2] - L[:, 2, 0, None] * y0 - L[:, 2, 1, None] * y1) / L[:, 2, 2, None]
    y3 = (B[:, 3] - L[:, 3, 0, None] * y0 - L[:, 3, 1, None] * y1 - L[:, 3, 2, None] * y2) / L[:, 3, 3, None]
    # backward: L^T X = Y
    x3 = y3 / L[:, 3, 3, None]
    x2 = (y2 - L[:, 3, 2, None] * x3) / L[:, 2, 2, None]
    x1 = (y1 - L[:, 2, 1, None] * x2 - L[:, 3, 1, None] * x3) / L[:, 1, 1, None]
    x0 = (y0 - L[:, 1, 0, None] * x1 - L[:, 2, 0, None] * x2 - L[:, 3, 0, None] * x3) / L[:, 0, 0, None]
    gain = np.stack([x0, x1, x2, x3], axis=2)                    # [M,8,4] = P H^T S^-1
    gt = gain.transpose(0, 2, 1)                                 # [M,4,8]
    new_means = means + np.matmul((zs - pm)[:, None, :], gt)[:, 0, :]
    new_covs = covs - np.matmul(gain, np.matmul(pc, gt))
    return new_means, new_covs


def kf_gating_distance(mean: np.ndarray, cov: np.ndarray, measurements: np.ndarray) -> np.ndarray:
    pm, pc = kf_project(mean, cov)
    L = np.linalg.cholesky(pc)
    z = solve_triangular(L, (measurements - pm).T, lower=True, check_finite=False)
    return np.sum(z * z, axis=0)


def kf_gating_distance_batch(means: np.ndarray, covs: np.ndarray, measurements: np.ndarray) -> np.ndarray:
    """Squared Mahalanobis distance of every measurement [D,4] to every track's projected state (kalman_filter.py:188-229) in
    one pass: means [T,8], covs [T,8,8] -> [T,D].  H = [I4 | 0], so H P H^T is the top-left 4x4 block exactly; Cholesky per
    track (batched LAPACK potrf, as the reference's np.linalg.cholesky) and the triangular solve written out as forward
    substitution over the four rows (the reference calls scipy.linalg.solve_triangular once per track: 0.6 ms each here)."""
    h = means[:, 3]
    pc = covs[:, :4, :4].copy()
    r = np.square(STD_POS * h)
    pc[:, 0, 0] += r
    pc[:, 1, 1] += r
    pc[:, 2, 2] += np.square(1e-1)
    pc[:, 3, 3] += r
    L = np.linalg.cholesky(pc)                                   # [T,4,4] lower
    d = measurements[None, :, :] - means[:, None, :4]            # [T,D,4]
    z0 = d[:, :, 0] / L[:, 0, 0, None]
    z1 = (d[:, :, 1] - L[:, 1, 0, None] * z0) / L[:, 1, 1, None]
    z2 = (d[:, :, 2] - L[:, 2, 0, None] * z0 - L[:, 2, 1, None] * z1) / L[:, 2, 2, None]
    z3 = (d[:, :, 3] - L[:, 3, 0, None] * z0 - L[:, 3, 1, None] * z1 - L[:, 3, 2, None] * z2) / L[:, 3, 3, None]
    return ((z0 * z0 + z1 * z1) + z2 * z2) + z3 * z3


# ---------------------------------------------------------------------------------------------
# appearance gallery: per identity, the last `budget` embeddings
# ---------------------------------------------------------------------------------------------
class Gallery:
    """nn_matching.py:99-177.  `samples` keeps the raw embeddings as the reference does; next to them every identity has its
    L2-normalised rows in ONE contiguous array in chronological order (a 2x budget buffer, compacted once per `budget` inserts),
    so that `distance` neither re-normalises nor re-stacks up to budget x 512 floats per identity and call."""

    def __init__(self, matching_threshold: float, budget: Optional[int]):
        self.matching_threshold = matching_threshold
        self.budget = budget
        self.samples = {}                       # track_id -> list of float32[512]
        self._norm = {}                         # track_id -> [buffer [cap, dim], begin, end]

    @staticmethod
    def _normalise_rows(a: np.ndarray) -> np.ndarray:
        return a / np.linalg.norm(a, axis=1, keepdims=True)       # the reference's expression (row-wise, independent of the other rows)

    def _push(self, t, f: np.ndarray) -> None:
        ent = self._norm.get(t)
        row = self._normalise_rows(np.asarray(f)[None, :])[0]
        if ent is None:
            cap = 2 * self.budget if self.budget is not None else 64
            ent = self._norm[t] = [np.empty((cap, row.shape[0]), row.dtype), 0, 0]
        buf, b, e = ent
        if e == buf.shape[0]:
            if self.budget is not None:                             # compact: the live rows move to the front
                n = e - b
                buf[:n] = buf[b:e].copy()
                b, e = 0, n
            else:                                                   # unbounded gallery: grow
                buf = np.concatenate([buf, np.empty_like(buf)], 0)
                ent[0] = buf
        buf[e] = row
        e += 1
        if self.budget is not None and e - b > self.budget:
            b = e - self.budget
        ent[1], ent[2] = b, e

    def partial_fit(self, features, targets, active_targets) -> None:
        for f, t in zip(features, targets):
            lst = self.samples.setdefault(t, [])
            lst.append(f)
            if self.budget is not None and len(lst) > self.budget:
                del lst[:len(lst) - self.budget]
            self._push(t, f)
        self.samples = {k: self.samples[k] for k in active_targets}
        self._norm = {k: self._norm[k] for k in active_targets}

    def distance(self, features: np.ndarray, targets: Sequence[int]) -> np.ndarray:
        """min over the gallery of (1 - cosine similarity); both sides re-normalised (nn_matching.py:33-52)."""
        cost = np.zeros((len(targets), len(features)))
        if len(features) == 0:
            return cost
        bt = self._normalise_rows(np.asarray(features)).T
        for i, t in enumerate(targets):
            buf, b, e = self._norm[t]
            cost[i] = (1.0 - buf[b:e] @ bt).min(axis=0)
        return cost


# ---------------------------------------------------------------------------------------------
# association
# ---------------------------------------------------------------------------------------------
def _iou_rows(bbox: np.ndarray, cands: np.ndarray) -> np.ndarray:
    tl = np.maximum(bbox[:2], cands[:, :2])
    br = np.minimum(bbox[:2] + bbox[2:], cands[:, :2] + cands[:, 2:])
    wh = np.maximum(0.0, br - tl)
    inter = wh.prod(axis=1)
    return inter / (bbox[2:].prod() + cands[:, 2:].prod(axis=1) - inter)


def _assign(cost: np.ndarray, max_distance: float, track_indices: List[int], detection_indices: List[int]):
    """linear_assignment.min_cost_matching after the cost matrix exists; list orders are part of the contract
    (unmatched detections become new identities in this order)."""
    cost = cost.copy()
    cost[cost > max_distance] = max_distance + 1e-5
    rows, cols = linear_sum_assignment(cost)
    colset, rowset = set(cols.tolist()), set(rows.tolist())
    matches = []
    unmatched_d = [d for c, d in enumerate(detection_indices) if c not in colset]
    unmatched_t = [t for r, t in enumerate(track_indices) if r not in rowset]
    for r, c in zip(rows, cols):
        t, d = track_indices[r], detection_indices[c]
        if cost[r, c] > max_distance:
            unmatched_t.append(t)
            unmatched_d.append(d)
        else:
            matches.append((t, d))
    return matches, unmatched_t, unmatched_d


class Tracker:
    """One multi-target tracker (the reference builds one per class, modules/track.py:16)."""

    def __init__(self, metric: Gallery, max_iou_distance=0.7, max_age=70, n_init=3):
        self.metric = metric
        self.max_iou_distance = max_iou_distance
        self.max_age = max_age
        self.n_init = n_init
        self.tracks: List[Track] = []
        self._next_id = 1

    def predict(self) -> None:
        if not self.tracks:
            return
        means = np.stack([t.mean for t in self.tracks])
        covs = np.stack([t.covariance for t in self.tracks])
        means, covs = kf_predict_batch(means, covs)
        for i, t in enumerate(self.tracks):
            t.mean, t.covariance = means[i], covs[i]
            t.age += 1
            t.time_since_update += 1

    def _appearance_cost(self, dets: List[Detection], track_idx: List[int], det_idx: List[int]) -> np.ndarray:
        feats = np.array([dets[i].feature for i in det_idx])
        cost = self.metric.distance(feats, [self.tracks[i].track_id for i in track_idx])
        meas = np.asarray([dets[i].to_xyah() for i in det_idx])
        means = np.stack([self.tracks[ti].mean for ti in track_idx])
        covs = np.stack([self.tracks[ti].covariance for ti in track_idx])
        cost[kf_gating_distance_batch(means, covs, meas) > CHI2INV95_4DOF] = INFTY_COST
        return cost

    def _iou_cost(self, dets: List[Detection], track_idx: List[int], det_idx: List[int]) -> np.ndarray:
        cost = np.zeros((len(track_idx), len(det_idx)))
        cands = np.asarray([dets[i].tlwh for i in det_idx])
        for r, ti in enumerate(track_idx):
            tr = self.tracks[ti]
            cost[r] = INFTY_COST if tr.time_since_update > 1 else 1.0 - _iou_rows(tr.to_tlwh(), cands)
        return cost

    def _match(self, dets: List[Detection]):
        confirmed = [i for i, t in enumerate(self.tracks) if t.is_confirmed()]
        unconfirmed = [i for i, t in enumerate(self.tracks) if not t.is_confirmed()]
        # cascade over track age (most recently seen first)
        unmatched_d = list(range(len(dets)))
        matches_a: List[Tuple[int, int]] = []
        for level in range(self.max_age):
            if not unmatched_d:
                break
            lvl = [k for k in confirmed if self.tracks[k].time_since_update == 1 + level]
            if not lvl:
                continue
            m, _, unmatched_d = _assign(self._appearance_cost(dets, lvl, unmatched_d), self.metric.matching_threshold, lvl,
                                        unmatched_d)
            matches_a += m
        # same set expression as the reference: the iteration order of a set of ints is part of the row
        # order of the IoU cost matrix below
        unmatched_a = list(set(confirmed) - set(k for k, _ in matches_a))
        iou_cands = unconfirmed + [k for k in unmatched_a if self.tracks[k].time_since_update == 1]
        unmatched_a = [k for k in unmatched_a if self.tracks[k].time_since_update != 1]
        if iou_cands and unmatched_d:
            matches_b, unmatched_b, unmatched_d = _assign(self._iou_cost(dets, iou_cands, unmatched_d), self.max_iou_distance,
                                                          iou_cands, unmatched_d)
        else:
            matches_b, unmatched_b = [], iou_cands
        return matches_a + matches_b, list(set(unmatched_a + unmatched_b)), unmatched_d

    def update(self, dets: List[Detection]) -> None:
        matches, unmatched_t, unmatched_d = self._match(dets)
        if matches:
            m_means = np.stack([self.tracks[ti].mean for ti, _ in matches])
            m_covs = np.stack([self.tracks[ti].covariance for ti, _ in matches])
            m_z = np.stack([dets[di].to_xyah() for _, di in matches])
            m_means, m_covs = kf_update_batch(m_means, m_covs, m_z)
        for k, (ti, di) in enumerate(matches):
            tr, d = self.tracks[ti], dets[di]
            tr.mean, tr.covariance = m_means[k], m_covs[k]
            tr.features.append(d.feature)
            tr.confidence_scores.append(d.confidence)
            tr.hits += 1
            tr.time_since_update = 0
            if tr.state == TENTATIVE and tr.hits >= tr._n_init:
                tr.state = CONFIRMED
        for ti in unmatched_t:
            tr = self.tracks[ti]
            if tr.state == TENTATIVE or tr.time_since_update > tr._max_age:
                tr.state = DELETED
        for di in unmatched_d:
            d = dets[di]
            mean, cov = kf_initiate(d.to_xyah())
            self.tracks.append(Track(mean, cov, self._next_id, self.n_init, self.max_age, d.feature, d.confidence))
            self._next_id += 1
        self.tracks = [t for t in self.tracks if not t.is_deleted()]
        active = [t.track_id for t in self.tracks if t.is_confirmed()]
        feats, targets = [], []
        for t in self.tracks:
            if not t.is_confirmed():
                continue
            feats += t.features
            targets += [t.track_id] * len(t.features)
            t.features = []
        self.metric.partial_fit(feats, targets, active)


def host_nms(boxes_tlwh: np.ndarray, max_overlap: float, scores: Optional[np.ndarray]) -> List[int]:
    """sort/preprocessing.py:6-73: +1-pixel areas, overlap = inter / area(other), highest score first."""
    if len(boxes_tlwh) == 0:
        return []
    b = boxes_tlwh.astype(np.float64)
    x1, y1 = b[:, 0], b[:, 1]
    x2, y2 = b[:, 2] + b[:, 0], b[:, 3] + b[:, 1]
    area = (x2 - x1 + 1) * (y2 - y1 + 1)
    idxs = np.argsort(scores) if scores is not None else np.argsort(y2)
    pick = []
    while len(idxs) > 0:
        i = idxs[-1]
        rest = idxs[:-1]
        pick.append(int(i))
        w = np.maximum(0, np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]) + 1)
        h = np.maximum(0, np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]) + 1)
        idxs = rest[(w * h) / area[rest] <= max_overlap]
    return pick
