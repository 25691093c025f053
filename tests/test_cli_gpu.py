"""The descriptor-track and baseline CLIs end to end on the GPU (descriptor_eval.py, python -m vsc2022_b200.sscd_baseline):
reference flag names, output files and result lines (reference descriptor_eval.py:16-58, sscd_baseline.py:54-231)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from vsc2022_b200.index import VideoFeature
from vsc2022_b200.metrics import CandidatePair, Dataset, Match
from vsc2022_b200.storage import store_features

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _dataset(tmp_path):
    rng = np.random.default_rng(21)
    unit = lambda x: (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)
    refs = [unit(rng.normal(size=(30, 64))) for _ in range(12)]
    queries = [unit(rng.normal(size=(30, 64))) for _ in range(4)]
    noise = [unit(rng.normal(size=(30, 64))) for _ in range(6)]
    queries[0][5:20] = refs[7][10:25]
    queries[2][8:18] = refs[3][2:12]
    ts = np.stack([np.arange(30) * 1.0, np.arange(30) * 1.0 + 1.0], axis=1)
    files = {}
    for name, vids, base, ds in (("q", queries, 0, Dataset.QUERIES), ("r", refs, 100, Dataset.REFS), ("n", noise, 200, Dataset.REFS)):
        files[name] = str(tmp_path / f"{name}.npz")
        store_features(files[name], [VideoFeature(video_id=base + i, timestamps=ts, feature=f) for i, f in enumerate(vids)], ds)
    gt = [Match(query_id="Q000000", ref_id="R000107", query_start=5.0, query_end=20.0, ref_start=10.0, ref_end=25.0, score=1.0),
          Match(query_id="Q000002", ref_id="R000103", query_start=8.0, query_end=18.0, ref_start=2.0, ref_end=12.0, score=1.0)]
    files["gt"] = str(tmp_path / "gt.csv")
    Match.write_csv(gt, files["gt"])
    return files


def test_descriptor_eval_cli(tmp_path):
    f = _dataset(tmp_path)
    cand_file = str(tmp_path / "cands.csv")
    out = subprocess.run([sys.executable, os.path.join(REPO, "descriptor_eval.py"), "--query_features", f["q"],
                          "--ref_features", f["r"], "--candidates_output", cand_file, "--ground_truth", f["gt"]],
                         capture_output=True, text=True, cwd=REPO, timeout=300)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "Descriptor track micro-AP (uAP): 1.0000" in out.stderr + out.stdout
    cands = CandidatePair.read_csv(cand_file)
    assert [(c.query_id, c.ref_id) for c in cands[:2]] in ([("Q000000", "R000107"), ("Q000002", "R000103")],
                                                             [("Q000002", "R000103"), ("Q000000", "R000107")])


@pytest.mark.parametrize("score_norm", [False, True])
def test_sscd_baseline_cli(tmp_path, score_norm):
    f = _dataset(tmp_path)
    out_dir = str(tmp_path / "out")
    cmd = [sys.executable, "-m", "vsc2022_b200.sscd_baseline", "--query_features", f["q"], "--ref_features", f["r"],
           "--output_path", out_dir, "--ground_truth", f["gt"]]
    if score_norm:
        cmd += ["--score_norm_features", f["n"]]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=REPO, timeout=300)
    log = out.stderr + out.stdout
    assert out.returncode == 0, log[-3000:]
    assert "Candidate uAP: 1.0000" in log and "Matching track metric:" in log
    for name in ("candidates.csv", "matches.csv") + (("sn_queries.npz", "sn_refs.npz") if score_norm else ()):
        assert os.path.exists(os.path.join(out_dir, name)), name
    matches = Match.read_csv(os.path.join(out_dir, "matches.csv"))
    found = {(m.query_id, m.ref_id) for m in matches}
    assert {("Q000000", "R000107"), ("Q000002", "R000103")} <= found
    metric = float(log.split("Matching track metric:")[1].split()[0])
    assert metric > 0.5
    # a second run without --overwrite must refuse (sscd_baseline.py:186-189)
    again = subprocess.run(cmd, capture_output=True, text=True, cwd=REPO, timeout=300)
    assert again.returncode != 0 and "overwrite" in (again.stderr + again.stdout)
