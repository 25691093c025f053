"""`vsc.baseline.score_normalization` mirror (score_normalization.py:22-105): CSLS-style normalisation.

    bias(q) = -beta * max_n <q, n>        query' = [q, bias(q)]      ref' = [r, 1]
so that <query', ref'> = <q, r> + bias(q).  The 1-NN similarity against the noise set -- the only heavy step, one
nq x n_noise x d GEMM -- runs as the fused row-max epilogue of the tensor-core GEMM; the similarity matrix is never
materialised.  The light steps (variance arg-min, dimension drop, L2 normalisation) are elementwise device ops.
"""
import dataclasses
import logging
from typing import Callable, List, Tuple

import numpy as np

from . import _lib
from .index import METRIC_INNER_PRODUCT, FlatIndex, VideoFeature

logger = logging.getLogger("score_normalization.py")
logger.setLevel(logging.INFO)


def transform_features(features: List[VideoFeature], transform: Callable) -> List[VideoFeature]:
    return [dataclasses.replace(f, feature=transform(f.feature)) for f in features]


def _stack(features: List[VideoFeature], device):
    torch = _lib.require_cuda()
    host = np.concatenate([np.asarray(f.feature, dtype=np.float32) for f in features], axis=0)
    return torch.from_numpy(host).to(device)


def _unstack(features: List[VideoFeature], mat) -> List[VideoFeature]:
    host = mat.cpu().numpy()
    out, at = [], 0
    for f in features:
        n = len(f)
        out.append(dataclasses.replace(f, feature=host[at:at + n]))
        at += n
    return out


def _l2_rows(x):
    """sklearn.preprocessing.normalize(x): rows / ||row||_2, all-zero rows left alone."""
    torch = _lib.require_cuda()
    norms = torch.sqrt((x * x).sum(dim=1, keepdim=True))
    return x / torch.where(norms == 0, torch.ones_like(norms), norms)


def score_normalize(queries: List[VideoFeature], refs: List[VideoFeature], score_norm_refs: List[VideoFeature],
                    l2_normalize: bool = True, replace_dim: bool = True, beta: float = 1.0,
                    ) -> Tuple[List[VideoFeature], List[VideoFeature]]:
    if {f.video_id for f in refs}.intersection({f.video_id for f in score_norm_refs}):
        raise Exception("Normalizing on the dataset we're evaluating on is against VSC rules. "
                        "An independent dataset is needed.")
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    q, r, noise = _stack(queries, dev), _stack(refs, dev), _stack(score_norm_refs, dev)
    if score_norm_refs is not None and replace_dim:
        logger.info("Replacing dimension")
        drop = int(torch.var(noise, dim=0, unbiased=False).argmin())   # lowest-variance dimension of the noise set
        keep = [i for i in range(noise.shape[1]) if i != drop]
        q, r, noise = q[:, keep].contiguous(), r[:, keep].contiguous(), noise[:, keep].contiguous()
    if l2_normalize:
        logger.info("L2 normalizing")
        q, r, noise = _l2_rows(q), _l2_rows(r), _l2_rows(noise)
    logger.info("Applying score normalization")
    index = FlatIndex(noise.shape[1], METRIC_INNER_PRODUCT)
    index.add_device(noise)
    nearest = index.max_similarity(q)                                  # fused GEMM + row-max
    q = torch.cat([q, (-beta * nearest)[:, None]], dim=1)
    r = torch.cat([r, torch.ones_like(r[:, :1])], dim=1)
    return _unstack(queries, q), _unstack(refs, r)
