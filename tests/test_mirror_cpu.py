"""CPU tests of the host-side mirror: the reference's OWN metric / storage test files run unmodified against
vsc2022_b200 through the compat aliases (when /root/reference is mounted), plus restated known answers that
travel to boxes without the reference tree."""
import importlib.util
import io
import os
import sys
import unittest

import numpy as np
import pytest

from vsc2022_b200 import metrics, storage
from vsc2022_b200.index import VideoFeature
from vsc2022_b200.metrics import CandidatePair, Dataset, Intervals, Match, average_precision, match_metric

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference tree not mounted")
@pytest.mark.parametrize("name", ["test_metrics.py", "test_storage.py"])
def test_reference_test_file_passes_on_mirror(name):
    import subprocess
    code = f"""
import sys, importlib.util, unittest
sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import vsc2022_b200.compat as compat
compat.install(force=True)
spec = importlib.util.spec_from_file_location('ref_test', {os.path.join(REF, 'tests', name)!r})
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
res = unittest.TextTestRunner(verbosity=0).run(unittest.TestLoader().loadTestsFromModule(mod))
print('RAN', res.testsRun, 'FAIL', len(res.failures) + len(res.errors))
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert " FAIL 0" in out.stdout and "RAN 0" not in out.stdout, out.stdout + out.stderr


def m(qs, qe, rs, re_, score=1.0, query_id="Q1", ref_id="R2"):
    return Match(query_id=query_id, ref_id=ref_id, query_start=qs, query_end=qe, ref_start=rs, ref_end=re_, score=score)


def test_intervals_known_answers():
    a, b, c = Intervals([(2, 5), (7, 8)]), Intervals([(1, 3), (4, 7)]), Intervals([(-1, 0), (3.5, 12)])
    assert a.intersect_length(b) == pytest.approx(2)
    assert a.intersect_length(c) == pytest.approx(2.5)


def test_match_metric_known_answers():
    gt = [m(4, 14, 10, 18)]
    assert match_metric(gt, [m(4, 14, 10, 18)]).ap == pytest.approx(1.0)
    assert match_metric(gt, [m(4, 8, 10, 14, 1.0), m(8, 14, 14, 18, 2.0)]).ap == pytest.approx(1.0)
    good = match_metric(gt, [m(4, 8, 10, 14, 1.0), m(8, 14, 16, 18, 2.0), m(0, 30, 5, 25, 0.0)]).ap
    assert 0.9 < good < 1.0
    assert match_metric(gt, [m(4, 8, 10, 14, 1.0), m(8, 14, 16, 18, 2.0), m(0, 30, 5, 25, 3.0)]).ap < 0.5
    # VCSL figure 4(f): no GT box overlaps a prediction
    fig4f = match_metric([m(4, 14, 10, 18), m(20, 28, 21, 29)], [m(4, 14, 21, 29), m(20, 28, 10, 18)]).ap
    assert fig4f == pytest.approx(0.0)
    dets = [m(4, 14, 10, 18, 3.0, "Q2", "R2"), m(4, 14, 10, 18, 2.0, "Q1", "R1"), m(4, 14, 10, 18, 1.0, "Q1", "R2")]
    assert match_metric([m(4, 14, 10, 18, query_id="Q1", ref_id="R2")], dets).ap == pytest.approx(1 / 3)


def test_micro_ap_known_answers():
    C = lambda q, r, s: CandidatePair(metrics.format_video_id(q, Dataset.QUERIES), metrics.format_video_id(r, Dataset.REFS), s)
    gt = [C(1, 10, 1.0), C(2, 11, 1.0)]
    cases = [([C(1, 10, 8.0), C(2, 11, 4.0), C(99, 99, 2.0)], 1.0),
             ([C(1, 10, 8.0), C(2, 11, 4.0), C(99, 99, 5.0)], np.mean([1, 2 / 3])),
             ([C(1, 10, 3.0), C(2, 10, 2.0), C(99, 99, 1.0)], 0.5),
             ([C(1, 10, 2.0), C(2, 10, 3.0), C(99, 99, 1.0)], 0.25)]
    for preds, want in cases:
        ap = average_precision(gt, preds)
        assert ap.ap == pytest.approx(want) and ap.ap == pytest.approx(ap.simple_ap)
    with pytest.raises(AssertionError):
        average_precision(gt, [C(1, 10, 1.0), C(1, 10, 2.0)])


def test_csv_and_npz_round_trips(tmp_path):
    cands = [CandidatePair("Q000001", "R000010", 1.0), CandidatePair("Q000002", "R000011", 2.0)]
    buf = io.StringIO()
    CandidatePair.write_csv(cands, buf)
    buf.seek(0)
    assert CandidatePair.read_csv(buf) == cands
    ms = [m(4, 8, 10, 14, 1.0, "Q123456", "R000100"), m(8, 14, 14, 18, 2.0, "Q000011", "R000101")]
    buf = io.StringIO()
    Match.write_csv(ms, buf)
    buf.seek(0)
    assert Match.read_csv(buf) == ms
    rng = np.random.default_rng(0)
    feats = [VideoFeature(video_id=2, timestamps=np.arange(10) * 1.0, feature=rng.normal(size=(10, 32))),
             VideoFeature(video_id=3, timestamps=np.arange(20) / 3.0, feature=rng.normal(size=(20, 32)))]
    path = tmp_path / "f.npz"
    storage.store_features(str(path), feats, Dataset.QUERIES)
    back = storage.load_features(str(path))
    assert [v.video_id for v in back] == ["Q000002", "Q000003"]
    for a, b in zip(feats, back):
        np.testing.assert_allclose(a.feature, b.feature)
        np.testing.assert_allclose(a.timestamps, b.timestamps)
    with pytest.raises(ValueError):
        metrics.format_video_id(3, None)
