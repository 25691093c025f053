"""`vsc.metrics` mirror: ids, candidate / match records, micro-AP and the segment-level matching metric.

Behavioural contract: vsc/metrics.py of the reference (cited per function); the numbers are the yardstick the
parity tests use (tests/test_mirror_cpu.py runs the reference's own 13 known-answer cases against this module).  This module is
host-side bookkeeping -- nothing here is on the GPU path -- but it is written independently: interval unions are
computed with numpy sweeps instead of rebuilding Python interval lists for every prediction.
"""
import collections
import dataclasses
import enum
import itertools
from math import sqrt
from typing import Collection, Dict, List, NamedTuple, Optional, TextIO, Tuple, Union

import numpy as np
import pandas as pd
from sklearn.metrics import average_precision_score


class Dataset(enum.Enum):
    QUERIES = "Q"
    REFS = "R"


def format_video_id(video_id: Union[str, int], dataset: Optional[Dataset]) -> str:
    """metrics.py:27-40 -- ints become Q000123 / R000123; strings are checked against the dataset prefix."""
    if isinstance(video_id, (int, np.integer)):
        if dataset is None:
            raise ValueError("Unable to convert integer video_id without a Dataset enum")
        return f"{dataset.value}{video_id:06d}"
    assert isinstance(video_id, str), f"unexpected video_id: {video_id} of type {type(video_id)}"
    if dataset is not None:
        assert video_id[0] == dataset.value, f"dataset mismatch? got {video_id} for dataset {dataset}"
    return video_id


@dataclasses.dataclass
class CandidatePair:
    query_id: str
    ref_id: str
    score: float

    @classmethod
    def to_dataframe(cls, candidates: Collection["CandidatePair"]) -> pd.DataFrame:
        return pd.DataFrame([{"query_id": format_video_id(c.query_id, Dataset.QUERIES),
                              "ref_id": format_video_id(c.ref_id, Dataset.REFS),
                              "score": c.score} for c in candidates])

    @classmethod
    def write_csv(cls, candidates: Collection["CandidatePair"], file: Union[str, TextIO]):
        cls.to_dataframe(candidates).to_csv(file, index=False)

    @classmethod
    def read_csv(cls, file: Union[str, TextIO]) -> List["CandidatePair"]:
        df = pd.read_csv(file)
        return [cls(query_id=format_video_id(q, Dataset.QUERIES), ref_id=format_video_id(r, Dataset.REFS), score=s)
                for q, r, s in zip(df["query_id"], df["ref_id"], df["score"])]

    @classmethod
    def from_matches(cls, matches: Collection["Match"]) -> List["CandidatePair"]:
        best: Dict[Tuple[str, str], float] = collections.defaultdict(float)
        for m in matches:
            key = (m.query_id, m.ref_id)
            best[key] = max(m.score, best[key])
        return [cls(query_id=q, ref_id=r, score=s) for (q, r), s in best.items()]


@dataclasses.dataclass
class PrecisionRecallCurve:
    precisions: np.ndarray
    recalls: np.ndarray
    scores: np.ndarray

    def plot(self, ax=None, **kwargs):
        import matplotlib.pyplot as plt  # plotting is optional; matplotlib is not a dependency of the hot path
        if ax is None:
            _, ax = plt.subplots()
            ax.set_xlabel("recall")
            ax.set_ylabel("precision")
            ax.set_xlim(0, 1.05)
            ax.set_ylim(0, 1.05)
        ax.plot(self.recalls, self.precisions, **kwargs)
        return ax


@dataclasses.dataclass
class AveragePrecision:
    ap: float
    pr_curve: PrecisionRecallCurve
    simple_ap: Optional[float] = None


def _merge(intervals: List[Tuple[float, float]]) -> List[Tuple[float, float]]:
    """Union of closed intervals as a sorted list of disjoint ones (touching intervals fuse)."""
    out: List[Tuple[float, float]] = []
    for lo, hi in sorted(intervals):
        if out and lo <= out[-1][1]:
            if hi > out[-1][1]:
                out[-1] = (out[-1][0], hi)
        else:
            out.append((lo, hi))
    return out


class Intervals:
    """metrics.py:120-174 -- a set of non-overlapping intervals ordered by start."""

    def __init__(self, intervals: Optional[List[Tuple[float, float]]] = None):
        self.intervals = _merge(list(intervals)) if intervals and len(intervals) > 1 else list(intervals or [])

    def add(self, interval: Tuple[float, float]):
        self.intervals = _merge(self.intervals + [interval])

    def union(self, other: "Intervals") -> "Intervals":
        return Intervals(self.intervals + other.intervals)

    def total_length(self) -> float:
        length = 0.0
        for lo, hi in self.intervals:
            length += hi - lo
        return length

    def intersect_length(self, other: "Intervals") -> float:
        # |A n B| = |A| + |B| - |A u B|
        return self.total_length() + other.total_length() - self.union(other).total_length()

    def __str__(self):
        return str(self.intervals)

    __repr__ = __str__


class Axis(enum.Enum):
    QUERY = enum.auto()
    REF = enum.auto()


class Match(NamedTuple):
    """A ground-truth or predicted copied segment (metrics.py:182-235)."""
    query_id: str
    ref_id: str
    score: float
    query_start: float
    query_end: float
    ref_start: float
    ref_end: float

    def pair_id(self):
        return (self.query_id, self.ref_id)

    def interval(self, axis: Axis) -> Tuple[float, float]:
        return (self.query_start, self.query_end) if axis == Axis.QUERY else (self.ref_start, self.ref_end)

    def intersection_area(self, other: "Match") -> float:
        dq = min(self.query_end, other.query_end) - max(self.query_start, other.query_start)
        dr = min(self.ref_end, other.ref_end) - max(self.ref_start, other.ref_start)
        return abs(max(dq, 0) * max(dr, 0))

    def overlaps(self, other: "Match") -> bool:
        return self.intersection_area(other) > 0.0

    @classmethod
    def write_csv(cls, matches: Collection["Match"], file: Union[str, TextIO]):
        pd.DataFrame([m._asdict() for m in matches], columns=cls._fields).to_csv(file, index=False)

    @classmethod
    def read_csv(cls, file: Union[str, TextIO], is_gt=False, check=True) -> List["Match"]:
        df = pd.read_csv(file)
        df["query_id"] = df.query_id.map(lambda x: format_video_id(x, Dataset.QUERIES))
        df["ref_id"] = df.ref_id.map(lambda x: format_video_id(x, Dataset.REFS))
        if is_gt:
            df["score"] = 1.0
        if check:
            for field in cls._fields:
                assert not df[field].isna().any()
        return [cls(**{f: rec[f] for f in cls._fields}) for rec in df.to_dict("records")]


class VideoPair:
    """Running intersection / coverage of one (query, ref) pair as predictions arrive (metrics.py:238-301)."""

    def __init__(self):
        self.intersections = {axis: 0.0 for axis in Axis}
        self.totals = {axis: 0.0 for axis in Axis}
        self.gts: List[Match] = []
        self.preds: List[Match] = []

    def total_gt_length(self, axis: Axis) -> float:
        return Intervals([g.interval(axis) for g in self.gts]).total_length()

    def total_pred_length(self, axis: Axis) -> float:
        return Intervals([p.interval(axis) for p in self.preds]).total_length()

    def gt_overlaps(self, gt: Match) -> bool:
        return any(gt.overlaps(p) for p in self.preds)

    def add_gt(self, bbox: Match):
        self.gts.append(bbox)

    def add_prediction(self, bbox: Match):
        """Returns (intersection deltas, coverage deltas) per axis caused by this prediction."""
        self.preds.append(bbox)
        live_gts = [g for g in self.gts if self.gt_overlaps(g)]  # GTs untouched by any prediction do not count
        d_inter, d_total = {}, {}
        for axis in Axis:
            pred_ints = Intervals([p.interval(axis) for p in self.preds])
            inter = pred_ints.intersect_length(Intervals([g.interval(axis) for g in live_gts]))
            covered = pred_ints.total_length()
            d_inter[axis] = inter - self.intersections[axis]
            d_total[axis] = covered - self.totals[axis]
            self.intersections[axis], self.totals[axis] = inter, covered
        return d_inter, d_total


def match_metric(gts: Collection[Match], predictions: Collection[Match]) -> AveragePrecision:
    """Segment AP of the matching track (metrics.py:304-378): AP = sum_i P(i) dR(i) with
    P = sqrt(P_q P_r), R = sqrt(R_q R_r) computed VCSL-style over groups of equal-score predictions."""
    predictions = sorted(predictions, key=lambda m: m.score, reverse=True)
    pairs: Dict[tuple, VideoPair] = collections.defaultdict(VideoPair)
    for gt in gts:
        pairs[gt.pair_id()].add_gt(gt)
    gt_len = {axis: 0.0 for axis in Axis}
    for vp in pairs.values():
        for axis in Axis:
            gt_len[axis] += vp.total_gt_length(axis)

    recall = metric = 0.0
    inter = {axis: 0.0 for axis in Axis}
    covered = {axis: 0.0 for axis in Axis}
    pr_r, pr_p, pr_s = [], [], []
    for score, group in itertools.groupby(predictions, key=lambda m: m.score):
        for pred in group:
            d_inter, d_total = pairs[pred.pair_id()].add_prediction(pred)
            for axis in Axis:
                inter[axis] += d_inter[axis]
                covered[axis] += d_total[axis]
        rec = {axis: inter[axis] / gt_len[axis] for axis in Axis}
        prec = {axis: inter[axis] / covered[axis] for axis in Axis}
        new_recall = sqrt(rec[Axis.QUERY] * rec[Axis.REF])
        precision = sqrt(prec[Axis.QUERY] * prec[Axis.REF])
        delta = new_recall - recall
        metric += precision * delta
        recall = new_recall
        if delta > 0:
            pr_r.append(recall)
            pr_p.append(precision)
            pr_s.append(score)
    return AveragePrecision(metric, PrecisionRecallCurve(np.array(pr_p), np.array(pr_r), np.array(pr_s)))


@dataclasses.dataclass
class MatchingTrackMetrics:
    segment_ap: AveragePrecision       # main metric
    pairwise_micro_ap: AveragePrecision  # pair-level only, ignores localisation


def evaluate_matching_track(ground_truth_filename: str, predictions_filename: str) -> MatchingTrackMetrics:
    """metrics.py:389-415 -- CSV in, both metrics out (columns in any order; GT scores ignored)."""
    gt = Match.read_csv(ground_truth_filename, is_gt=True)
    predictions = Match.read_csv(predictions_filename)
    pair_ap = average_precision(CandidatePair.from_matches(gt), CandidatePair.from_matches(predictions))
    return MatchingTrackMetrics(segment_ap=match_metric(gt, predictions), pairwise_micro_ap=pair_ap)


def average_precision(ground_truth: Collection[CandidatePair], predictions: Collection[CandidatePair]) -> AveragePrecision:
    """Micro-AP over (query, ref) pairs (metrics.py:418-450)."""
    gt_pairs = {(p.query_id, p.ref_id) for p in ground_truth}
    if len(gt_pairs) != len(ground_truth):
        raise AssertionError("Duplicates detected in ground truth")
    if len({(p.query_id, p.ref_id) for p in predictions}) != len(predictions):
        raise AssertionError("Duplicates detected in predictions")
    canonical = drivendata_average_precision(predicted=CandidatePair.to_dataframe(predictions),
                                             ground_truth=CandidatePair.to_dataframe(ground_truth))
    ranked = sorted(predictions, key=lambda p: p.score, reverse=True)
    scores = np.array([p.score for p in ranked])
    correct = np.array([(p.query_id, p.ref_id) in gt_pairs for p in ranked])
    hits = np.cumsum(correct)
    precision = hits / (np.arange(len(correct)) + 1)
    recall = hits / len(gt_pairs)
    simple_ap = np.sum(precision * correct) / len(gt_pairs)
    at = np.nonzero(correct)[0]
    return AveragePrecision(ap=canonical, pr_curve=PrecisionRecallCurve(precision[at], recall[at], scores[at]),
                            simple_ap=simple_ap)


def drivendata_average_precision(predicted: pd.DataFrame, ground_truth: pd.DataFrame):
    """The challenge's canonical AP (metrics.py:453-489): sklearn AP on the predicted pairs, rescaled by the
    fraction of ground-truth pairs that were predicted at all."""
    actual = ground_truth[["query_id", "ref_id"]]
    if not np.isfinite(predicted["score"]).all() or np.isnan(predicted["score"]).any():
        raise ValueError("Scores must be finite.")
    predicted = predicted.sort_values("score", ascending=False)
    merged = predicted.merge(right=actual.assign(actual=1.0), how="left", on=["query_id", "ref_id"]).fillna({"actual": 0.0})
    unadjusted = average_precision_score(merged["actual"].values, merged["score"].values) if merged["actual"].sum() else 0.0
    n_found = int(merged["actual"].sum())
    n_actual = int(actual["ref_id"].notna().sum())
    return unadjusted * (n_found / n_actual)
