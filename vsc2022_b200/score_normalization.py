"""`vsc.baseline.score_normalization` mirror (score_normalization.py:22-105): CSLS-style normalisation.

    bias(q) = -beta * max_n <q, n>        query' = [q, bias(q)]      ref' = [r, 1]
so that <query', ref'> = <q, r> + bias(q).  Everything runs on the device in six launches: the lowest-variance
column of the noise set (vsc_lowvar_dim), one drop-that-column + L2-normalise pass per collection writing straight
into the widened output matrices (vsc_l2norm_dropdim), the 1-NN similarity against the noise set as the fused row-max
epilogue of the tensor-core GEMM (the nq x n_noise similarity matrix is never materialised), and the bias column
(vsc_fill_column).  `on_device=True` (extension) leaves the result on the GPU: the returned VideoFeatures hold row views
of two device matrices, which CandidateGeneration and the localization classes consume without a host round trip.
"""
import ctypes
import dataclasses
import logging
from typing import Callable, List, Tuple

from . import _lib
from .device_features import features_matrix, split_rows
from .index import METRIC_INNER_PRODUCT, FlatIndex, VideoFeature

logger = logging.getLogger("score_normalization.py")
logger.setLevel(logging.INFO)


def transform_features(features: List[VideoFeature], transform: Callable) -> List[VideoFeature]:
    return [dataclasses.replace(f, feature=transform(f.feature)) for f in features]


def _stream(torch, dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def lowvar_dim(x):
    """Device int32 tensor holding argmin_c var(x[:, c]) (float64 moments; numpy's `var(axis=0).argmin()`)."""
    torch = _lib.require_cuda()
    out = torch.empty((1,), dtype=torch.int32, device=x.device)
    scratch = torch.empty((2 * x.shape[1],), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().vsc_lowvar_dim(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), out.data_ptr(),
                                        scratch.data_ptr(), _stream(torch, x.device))
    _lib.check(rc, "vsc_lowvar_dim")
    return out


def l2norm_dropdim(x, drop, normalize: bool, extra_column: bool, tail=None):
    """x without column *drop (None: all columns), rows L2-normalised if asked; `extra_column` reserves one more column
    at the end, set to `tail` when given."""
    torch = _lib.require_cuda()
    n, d = x.shape
    kept = d - (1 if drop is not None else 0)
    out = torch.empty((n, kept + (1 if extra_column else 0)), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().vsc_l2norm_dropdim(x.data_ptr(), n, d, x.stride(0) if n else d,
                                            drop.data_ptr() if drop is not None else None, 1 if normalize else 0,
                                            out.data_ptr(), out.shape[1], 1 if tail is not None else 0,
                                            float(tail) if tail is not None else 0.0, _stream(torch, x.device))
    _lib.check(rc, "vsc_l2norm_dropdim")
    return out, kept


def score_normalize(queries: List[VideoFeature], refs: List[VideoFeature], score_norm_refs: List[VideoFeature],
                    l2_normalize: bool = True, replace_dim: bool = True, beta: float = 1.0, on_device: bool = False,
                    ) -> Tuple[List[VideoFeature], List[VideoFeature]]:
    if {f.video_id for f in refs}.intersection({f.video_id for f in score_norm_refs}):
        raise Exception("Normalizing on the dataset we're evaluating on is against VSC rules. "
                        "An independent dataset is needed.")
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    q, r, noise = (features_matrix(f, dev) for f in (queries, refs, score_norm_refs))
    qn, rn = score_normalize_device(q, r, noise, l2_normalize, replace_dim and score_norm_refs is not None, beta)
    return split_rows(queries, qn, on_device), split_rows(refs, rn, on_device)


def score_normalize_device(q, r, noise, l2_normalize: bool = True, replace_dim: bool = True, beta: float = 1.0):
    """The arithmetic of score_normalize on float32 CUDA matrices [rows, d]; returns the widened (queries, refs)."""
    torch = _lib.require_cuda()
    dev = q.device
    drop = None
    if replace_dim:
        logger.info("Replacing dimension")
        drop = lowvar_dim(noise)                                       # lowest-variance dimension of the noise set
    if l2_normalize:
        logger.info("L2 normalizing")
    qn, kept = l2norm_dropdim(q, drop, l2_normalize, extra_column=True)
    rn, _ = l2norm_dropdim(r, drop, l2_normalize, extra_column=True, tail=1.0)
    nn, _ = l2norm_dropdim(noise, drop, l2_normalize, extra_column=False)
    logger.info("Applying score normalization")
    index = FlatIndex(kept, METRIC_INNER_PRODUCT)
    index.add_device(nn, copy=False)
    nearest = index.max_similarity(qn[:, :kept])                       # fused GEMM + row-max
    with torch.cuda.device(dev):
        rc = _lib.load().vsc_fill_column(qn.data_ptr(), qn.shape[0], qn.shape[1], kept, nearest.data_ptr(), -float(beta),
                                         _stream(torch, dev))
    _lib.check(rc, "vsc_fill_column")
    return qn, rn
