#!/usr/bin/env python3
"""`vsc.baseline.sscd_baseline` mirror (sscd_baseline.py:54-236): retrieval + matching from SSCD descriptors.

    python -m vsc2022_b200.sscd_baseline --query_features q.npz --ref_features r.npz \
        [--score_norm_features n.npz] --output_path out [--ground_truth gt.csv] [--overwrite]

Same flags, constants (1200 / 25 / 5 per query, beta = 1.2, TN tn_max_step=5 min_length=4, bias 0.5, batches of
512 pairs), output files (sn_queries.npz, sn_refs.npz, candidates.csv, matches.csv) and log lines as the reference.
"""
import argparse
import logging
import os
from typing import List, Tuple

import numpy as np

from .candidates import CandidateGeneration, MaxScoreAggregation
from .index import VideoFeature
from .localization import VCSLLocalizationCandidateScore, VCSLLocalizationMaxSim
from .metrics import AveragePrecision, CandidatePair, Dataset, Match, average_precision, evaluate_matching_track
from .device_features import to_host
from .score_normalization import score_normalize, transform_features
from .storage import load_features, store_features

logging.basicConfig(format="%(asctime)s %(levelname)-8s %(message)s", level=logging.INFO, datefmt="%Y-%m-%d %H:%M:%S")
logger = logging.getLogger("sscd_baseline.py")
logger.setLevel(logging.INFO)

parser = argparse.ArgumentParser()
parser.add_argument("--query_features", help="Path to query descriptors", type=str, required=True)
parser.add_argument("--ref_features", help="Path to reference descriptors", type=str, required=True)
parser.add_argument("--score_norm_features", help="Path to score normalization descriptors", type=str)
parser.add_argument("--output_path", help="The path to write match predictions.", type=str, required=True)
parser.add_argument("--ground_truth", help="Path to the ground truth (labels) CSV file.", type=str)
parser.add_argument("--overwrite", help="Overwrite prediction files, if found.", action="store_true")


def l2_normalize_rows(x: np.ndarray) -> np.ndarray:
    """sklearn.preprocessing.normalize(x): rows scaled to unit L2 norm, zero rows untouched."""
    norms = np.sqrt(np.einsum("ij,ij->i", x, x))
    norms[norms == 0.0] = 1.0
    return x / norms[:, np.newaxis]


def search(queries: List[VideoFeature], refs: List[VideoFeature], retrieve_per_query: float = 1200.0,
           candidates_per_query: float = 25.0) -> List[CandidatePair]:
    logger.info("Searching")
    cg = CandidateGeneration(refs, MaxScoreAggregation())
    candidates = cg.query(queries, global_k=int(retrieve_per_query * len(queries)),
                          limit=int(candidates_per_query * len(queries)))
    logger.info("Got %d candidates", len(candidates))
    return candidates


def localize_and_verify(queries: List[VideoFeature], refs: List[VideoFeature], candidates: List[CandidatePair],
                        localize_per_query: float = 5.0, score_normalization: bool = False) -> List[Match]:
    candidates = candidates[:int(len(queries) * localize_per_query)]
    if score_normalization:
        alignment = VCSLLocalizationMaxSim(queries, refs, model_type="TN", tn_max_step=5, min_length=4,
                                           concurrency=16, similarity_bias=0.5)
    else:
        alignment = VCSLLocalizationCandidateScore(transform_features(queries, l2_normalize_rows),
                                                   transform_features(refs, l2_normalize_rows), model_type="TN",
                                                   tn_max_step=5, min_length=4, concurrency=16)
    matches: List[Match] = []
    logger.info("Aligning %s candidate pairs", len(candidates))
    BATCH_SIZE = 512
    for i in range(0, len(candidates), BATCH_SIZE):
        batch = candidates[i:i + BATCH_SIZE]
        matches.extend(alignment.localize_all(batch))
        logger.info("Aligned %d pairs of %d; %d predictions so far", i + len(batch), len(candidates), len(matches))
    return matches


def match(queries: List[VideoFeature], refs: List[VideoFeature], output_path: str,
          score_normalization: bool = False) -> Tuple[str, str]:
    candidates = search(queries, refs)
    os.makedirs(output_path, exist_ok=True)
    candidate_file = os.path.join(output_path, "candidates.csv")
    CandidatePair.write_csv(candidates, candidate_file)
    matches = localize_and_verify(queries, refs, candidates, score_normalization=score_normalization)
    matches_file = os.path.join(output_path, "matches.csv")
    Match.write_csv(matches, matches_file)
    return candidate_file, matches_file


def create_pr_plot(ap: AveragePrecision, filename: str):
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        logger.info("matplotlib not installed; skipping %s", filename)
        return
    ap.pr_curve.plot(linewidth=1)
    plt.savefig(filename)
    plt.show()


def main(args):
    if os.path.exists(args.output_path) and not args.overwrite:
        raise Exception(f"Output path already exists: {args.output_path}. Do you want to --overwrite?")
    queries = load_features(args.query_features, Dataset.QUERIES)
    refs = load_features(args.ref_features, Dataset.REFS)
    score_normalization = False
    if args.score_norm_features:
        # the normalised descriptors stay on the device for the search and the localization; the .npz copies the
        # reference writes are taken from there
        queries, refs = score_normalize(queries, refs, load_features(args.score_norm_features, Dataset.REFS), beta=1.2,
                                        on_device=True)
        score_normalization = True
        os.makedirs(args.output_path, exist_ok=True)
        store_features(os.path.join(args.output_path, "sn_queries.npz"), to_host(queries))
        store_features(os.path.join(args.output_path, "sn_refs.npz"), to_host(refs))
    candidate_file, match_file = match(queries, refs, args.output_path, score_normalization=score_normalization)
    if not args.ground_truth:
        return
    gt_pairs = CandidatePair.from_matches(Match.read_csv(args.ground_truth, is_gt=True))
    candidate_uap = average_precision(gt_pairs, CandidatePair.read_csv(candidate_file))
    logger.info(f"Candidate uAP: {candidate_uap.ap:.4f}")
    candidate_pr_file = os.path.join(args.output_path, "candidate_precision_recall.pdf")
    create_pr_plot(candidate_uap, candidate_pr_file)
    match_metrics = evaluate_matching_track(args.ground_truth, match_file)
    logger.info(f"Matching track metric: {match_metrics.segment_ap.ap:.4f}")
    matching_pr_file = os.path.join(args.output_path, "precision_recall.pdf")
    create_pr_plot(match_metrics.segment_ap, matching_pr_file)
    logger.info(f"Candidates: {candidate_file}")
    logger.info(f"Matches: {match_file}")
    logger.info(f"Candidate PR plot: {candidate_pr_file}")
    logger.info(f"Match PR plot: {matching_pr_file}")


if __name__ == "__main__":
    main(parser.parse_args())
