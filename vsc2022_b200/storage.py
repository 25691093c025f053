"""`vsc.storage` mirror: the .npz wire format between the stages (vsc/storage.py:13-72).

One archive holds `video_ids` (N), `features` (N x D) and `timestamps` (N or N x 2); a video is a run of equal ids.
"""
from typing import Dict, List, Optional

import numpy as np

from .index import VideoFeature
from .metrics import Dataset, format_video_id


def store_features(f, features: List[VideoFeature], dataset: Optional[Dataset] = None):
    ids = [np.full(len(v), format_video_id(v.video_id, dataset)) for v in features]
    np.savez(f, video_ids=np.concatenate(ids), features=np.concatenate([v.feature for v in features]),
             timestamps=np.concatenate([v.timestamps for v in features]))


def same_value_ranges(values):
    """(value, start, end) for every maximal run of equal consecutive values."""
    values = np.asarray(values)
    if len(values) == 0:
        return
    cuts = np.flatnonzero(values[1:] != values[:-1]) + 1
    starts = np.concatenate([[0], cuts])
    ends = np.concatenate([cuts, [len(values)]])
    for s, e in zip(starts, ends):
        yield values[s], int(s), int(e)


def load_features(f, dataset: Optional[Dataset] = None) -> List[VideoFeature]:
    data = np.load(f, allow_pickle=False)
    video_ids, feats, timestamps = data["video_ids"], data["features"], data["timestamps"]
    if timestamps.shape[0] != feats.shape[0]:
        raise ValueError(f"Expected the same number of timestamps as features: got {timestamps.shape[0]} "
                         f"timestamps for {feats.shape[0]} features")
    if not (timestamps.ndim == 1 or timestamps.shape[1:] == (2,)):
        raise ValueError(f"Unexpected timestamp shape. Got {timestamps.shape}")
    return [VideoFeature(video_id=format_video_id(vid, dataset), timestamps=timestamps[s:e], feature=feats[s:e, :])
            for vid, s, e in same_value_ranges(video_ids)]


def convert_to_dict(features: List[VideoFeature]) -> Dict[str, VideoFeature]:
    return {v.video_id: v for v in features}
