"""Posterior post-processing on the GPU.

Replaces the host-side numpy tail of the reference's evaluation scripts -- ``fast_truncnorm`` (left = 4),
resampling of draws >= 9 from the analytic prior, min over trios, per-system median / percentiles
(figures/main_figures.py:167-277, figures/multiswag_5_planet.py:306-481) -- so that only ``[N, 8]`` leaves the
device instead of the ``[N, U, 2]`` prediction block.
"""
from __future__ import annotations

import torch

from . import _lib

STAT_NAMES = ("average", "median", "l", "u", "ll", "uu", "median_mu", "median_std")  # multiswag_5_planet.py:476-481


def sample_instability(pred: torch.Tensor, seed: int = 0, row_offset: int = 0, left: float = 4.0, nsamp: int = 40):
    """pred [rows, U, 2] (system-major (mu, std)) -> sampled log10 instability times [rows, U]."""
    lib = _lib.load()
    _lib.require_cuda(pred, "pred")
    pred = pred.contiguous().float()
    rows, U = pred.shape[0], pred.shape[1]
    t = torch.empty((rows, U), device=pred.device, dtype=torch.float32)
    if rows == 0:
        return t
    with torch.cuda.device(pred.device), _lib.nvtx("bnn:K7 sample_instability"):
        _lib.check(lib.bnn_sample_instability(_lib.ptr(pred), rows, U, int(seed), int(row_offset), float(left), int(nsamp),
                                              _lib.ptr(t), _lib.current_stream_ptr()), "bnn_sample_instability")
    return t


def summarize_instability(t: torch.Tensor, pred: torch.Tensor, n_trios: int = 1):
    """t [N*n_trios, U], pred [N*n_trios, U, 2] -> [N, 8] (STAT_NAMES)."""
    lib = _lib.load()
    _lib.require_cuda(t, "t")
    t, pred = t.contiguous().float(), pred.contiguous().float()
    rows, U = t.shape
    if rows % n_trios:
        raise ValueError(f"{rows} rows are not a multiple of {n_trios} trios")
    N = rows // n_trios
    stats = torch.empty((N, 8), device=t.device, dtype=torch.float32)
    if N == 0:
        return stats
    with torch.cuda.device(t.device), _lib.nvtx("bnn:K7 summarize_instability"):
        _lib.check(lib.bnn_summarize_instability(_lib.ptr(t), _lib.ptr(pred), N, int(n_trios), U, _lib.ptr(stats),
                                                 _lib.current_stream_ptr()), "bnn_summarize_instability")
    return stats


def posterior_summary(pred: torch.Tensor, n_trios: int = 1, seed: int = 0, row_offset: int = 0):
    """[N*n_trios, U, 2] predictions -> [N, 8] summary (sampling + min over trios + order statistics)."""
    return summarize_instability(sample_instability(pred, seed, row_offset), pred, n_trios)
