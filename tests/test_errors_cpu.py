"""Error conventions and host-side schedules of the mirror, checked without a GPU."""
import numpy as np
import pytest

from vsc2022_b200 import _lib
from vsc2022_b200.index import VideoFeature, VideoMetadata, exponential_batches, index_factory
from vsc2022_b200.metrics import CandidatePair, Dataset, format_video_id
from vsc2022_b200.storage import load_features


def test_timestamp_feature_mismatch_is_an_assertion():      # index.py:37-40
    with pytest.raises(AssertionError):
        VideoFeature(video_id=1, timestamps=np.arange(3.0), feature=np.zeros((4, 8), np.float32))


def test_timestamps_one_and_two_column():                    # index.py:26-30
    m1 = VideoMetadata(video_id=1, timestamps=np.array([0.0, 1.5, 3.0]))
    m2 = VideoMetadata(video_id=1, timestamps=np.array([[0.0, 1.0], [1.0, 2.5]]))
    assert m1.get_timestamps(1) == (1.5, 1.5) and m2.get_timestamps(1) == (1.0, 2.5)


def test_format_video_id_errors():                            # metrics.py:27-40
    assert format_video_id(7, Dataset.QUERIES) == "Q000007" and format_video_id(7, Dataset.REFS) == "R000007"
    with pytest.raises(ValueError):
        format_video_id(7, None)
    with pytest.raises(AssertionError):                       # the reference asserts on a prefix / dataset mismatch
        format_video_id("R000001", Dataset.QUERIES)


def test_load_features_rejects_bad_timestamps(tmp_path):     # storage.py:49-57
    f = str(tmp_path / "bad.npz")
    np.savez(f, video_ids=np.array(["Q000001"] * 3), features=np.zeros((3, 4), np.float32), timestamps=np.zeros(2))
    with pytest.raises(ValueError):
        load_features(f)
    np.savez(f, video_ids=np.array(["Q000001"] * 3), features=np.zeros((3, 4), np.float32), timestamps=np.zeros((3, 3)))
    with pytest.raises(ValueError):
        load_features(f)


def test_score_normalize_refuses_overlapping_ids():          # score_normalization.py:63-67 (raised before any GPU work)
    from vsc2022_b200.score_normalization import score_normalize
    v = lambda i: VideoFeature(video_id=i, timestamps=np.arange(2.0), feature=np.ones((2, 4), np.float32))
    with pytest.raises(Exception, match="against VSC rules"):
        score_normalize([v(1)], [v(2), v(3)], [v(3)])


def test_only_flat_index_is_provided():
    with pytest.raises(NotImplementedError):
        index_factory(8, "IVF256,Flat")


def test_exponential_batches_follow_faiss_schedule():        # faiss.contrib.exhaustive_search.exponential_query_iterator
    spans = list(exponential_batches(40000))
    sizes = [b - a for a, b in spans]
    assert sizes[:10] == [32 << i for i in range(10)] and sizes[10] == 40000 - sum(sizes[:10])
    assert spans[0][0] == 0 and spans[-1][1] == 40000 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    big = [b - a for a, b in exponential_batches(200000)]
    assert max(big) == 32768 and big.count(32768) >= 2      # doubling stops once the size reaches 20000 or more
    assert list(exponential_batches(0)) == []


def test_candidate_pair_csv_round_trip(tmp_path):
    f = str(tmp_path / "c.csv")
    CandidatePair.write_csv([CandidatePair("Q000001", "R000002", 0.5)], f)
    assert CandidatePair.read_csv(f) == [CandidatePair("Q000001", "R000002", 0.5)]


def test_engine_error_is_a_runtime_error():
    assert issubclass(_lib.EngineError, RuntimeError)
