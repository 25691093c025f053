"""The matching-track CLI (matching_eval.py, reference matching_eval.py:16-48) end to end on CPU: argument names and
the result line a drop-in must preserve."""
import os
import subprocess
import sys

from vsc2022_b200.metrics import Match, evaluate_matching_track

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_matching_eval_cli(tmp_path):
    gt = [Match(query_id="Q000001", ref_id="R000002", query_start=4.0, query_end=14.0, ref_start=10.0, ref_end=20.0, score=1.0),
          Match(query_id="Q000003", ref_id="R000004", query_start=0.0, query_end=8.0, ref_start=2.0, ref_end=10.0, score=1.0)]
    pred = [Match(query_id="Q000001", ref_id="R000002", query_start=4.0, query_end=14.0, ref_start=10.0, ref_end=20.0, score=0.9),
            Match(query_id="Q000003", ref_id="R000009", query_start=0.0, query_end=8.0, ref_start=2.0, ref_end=10.0, score=0.5)]
    gt_file, pred_file = str(tmp_path / "gt.csv"), str(tmp_path / "pred.csv")
    Match.write_csv(gt, gt_file)
    Match.write_csv(pred, pred_file)
    out = subprocess.run([sys.executable, os.path.join(REPO, "matching_eval.py"), "--predictions", pred_file,
                          "--ground_truth", gt_file], capture_output=True, text=True, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("Matching track segment AP:")]
    assert len(line) == 1
    want = evaluate_matching_track(gt_file, pred_file).segment_ap.ap     # the mirror passes the reference's own metric tests
    assert 0.0 < want < 1.0 and abs(float(line[0].split(":")[1]) - want) < 5e-5

    perfect = subprocess.run([sys.executable, os.path.join(REPO, "matching_eval.py"), "--predictions", gt_file,
                              "--ground_truth", gt_file], capture_output=True, text=True, cwd=REPO)
    assert "Matching track segment AP: 1.0000" in perfect.stdout
