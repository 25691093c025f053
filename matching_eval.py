#!/usr/bin/env python3
"""Matching track evaluation CLI (drop-in for the reference's matching_eval.py:16-48).

    python matching_eval.py --predictions matches.csv --ground_truth gt.csv
"""
import logging
from argparse import ArgumentParser, Namespace

from vsc2022_b200.metrics import evaluate_matching_track

parser = ArgumentParser()
parser.add_argument("--predictions", help="Path containing match predictions", type=str, required=True)
parser.add_argument("--ground_truth", help="Path containing ground truth labels", type=str, required=True)

logging.basicConfig(format="%(asctime)s %(levelname)-8s %(message)s", level=logging.INFO, datefmt="%Y-%m-%d %H:%M:%S")
logger = logging.getLogger("matching_eval.py")
logger.setLevel(logging.INFO)


def main(args: Namespace):
    metrics = evaluate_matching_track(args.ground_truth, args.predictions)
    print(f"Matching track segment AP: {metrics.segment_ap.ap:.4f}")


if __name__ == "__main__":
    main(parser.parse_args())
