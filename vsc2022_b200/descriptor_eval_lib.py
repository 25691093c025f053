"""`vsc.descriptor_eval_lib` mirror (descriptor_eval_lib.py:23-60): descriptor-track evaluation.

Retrieve 1200 frame pairs and keep 25 video pairs per query video, then micro-AP against the ground truth.
"""
import logging
from typing import List, Optional, Tuple

from .candidates import CandidateGeneration, MaxScoreAggregation
from .metrics import AveragePrecision, CandidatePair, Dataset, Match, average_precision
from .storage import load_features

logger = logging.getLogger("descriptor_eval_lib.py")
logger.setLevel(logging.INFO)

RETRIEVAL_CANDIDATES_PER_QUERY = 20 * 60  # similar to K=20 for ~60 second videos
AGGREGATED_CANDIDATES_PER_QUERY = 25


def evaluate_descriptor_track(query_feature_filename: str, ref_feature_filename: str,
                              ground_truth_filename: Optional[str]) -> Tuple[AveragePrecision, List[CandidatePair]]:
    logger.info("Starting Descriptor level eval")
    query_features = load_features(query_feature_filename, Dataset.QUERIES)
    logger.info(f"Loaded {len(query_features)} query features")
    ref_features = load_features(ref_feature_filename, Dataset.REFS)
    logger.info(f"Loaded {len(ref_features)} ref features")
    retrieval_candidates = int(RETRIEVAL_CANDIDATES_PER_QUERY * len(query_features))
    num_candidates = int(AGGREGATED_CANDIDATES_PER_QUERY * len(query_features))
    logger.info(f"Performing search for {retrieval_candidates} nearest vectors")
    cg = CandidateGeneration(ref_features, MaxScoreAggregation())
    score_candidates = cg.query(query_features, global_k=retrieval_candidates, limit=num_candidates)
    logger.info(f"Keeping the {len(score_candidates)} highest score pairs (at most {num_candidates}).")
    if ground_truth_filename is None:
        return None, score_candidates
    gt_pairs = CandidatePair.from_matches(Match.read_csv(ground_truth_filename, is_gt=True))
    logger.info(f"Loaded ground truth from {ground_truth_filename}")
    ap = average_precision(gt_pairs, score_candidates)
    logger.info(f"Descriptor track micro-AP (uAP): {ap.ap:.4f}")
    return ap, score_candidates
