"""Input packing on the GPU: the step between the N-body feature extraction and the model.

Mirrors ``data_setup_kernel`` (/root/reference/figures/spock/regression.py:183-213) followed by
``ssX.transform`` (:144) and ``torch.tensor(X).float()`` (:145); also the trio flattening of the 5-planet
script (figures/multiswag_5_planet.py:204,280-292).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .synth import SSX_MEAN, SSX_SCALE


def data_setup_kernel(mass_array: torch.Tensor, cur_tseries: torch.Tensor, ssX=None) -> torch.Tensor:
    """cur_tseries [N,T,26] float64 (CUDA), mass_array [N,3] float64 -> normalised model input [N,T,41] float32.

    ``ssX``: a fitted sklearn StandardScaler (``mean_``, ``scale_``) or None for the v50 constants hard-coded in
    ``load_swag`` (spock_reg_model.py:934-955).  Unlike the reference function this one also applies the scaler
    (its callers always do, :144) -- pass the raw arrays."""
    lib = _lib.load()
    _lib.require_cuda(cur_tseries, "cur_tseries")
    ts = cur_tseries.contiguous().double()
    ms = mass_array.to(ts.device).contiguous().double()
    if ts.dim() != 3 or ts.shape[-1] != 26 or ms.shape != (ts.shape[0], 3):
        raise NotImplementedError("Need [N,T,26] time series and [N,3] masses (regression.py:209-210)")
    mean = torch.as_tensor(np.asarray(SSX_MEAN if ssX is None else ssX.mean_), dtype=torch.float64, device=ts.device)
    scale = torch.as_tensor(np.asarray(SSX_SCALE if ssX is None else ssX.scale_), dtype=torch.float64, device=ts.device)
    N, T, _ = ts.shape
    x = torch.empty((N, T, 41), device=ts.device, dtype=torch.float32)
    with torch.cuda.device(ts.device), _lib.nvtx("bnn:K6 pack_inputs"):
        _lib.check(lib.bnn_pack_inputs(_lib.ptr(ts), _lib.ptr(ms), _lib.ptr(mean), _lib.ptr(scale), N, T, _lib.ptr(x),
                                       _lib.current_stream_ptr()), "bnn_pack_inputs")
    return x


def pack_trios(tseries: torch.Tensor, masses: torch.Tensor, ssX=None) -> torch.Tensor:
    """5-planet sliding window: tseries [N, n_trios, T, 26], masses [N, n_trios, 3] -> [N*n_trios, T, 41]
    (the reshape(-1, 100, 41) of multiswag_5_planet.py:287)."""
    N, R = tseries.shape[:2]
    return data_setup_kernel(masses.reshape(N * R, 3), tseries.reshape(N * R, tseries.shape[2], 26), ssX)
