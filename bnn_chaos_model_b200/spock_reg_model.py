"""Drop-in host mirror of the reference's ``spock_reg_model`` hot path, backed by libbnnchaos.

Same names, argument meaning and error behaviour as the reference classes
(/root/reference/spock_reg_model.py): ``VarModel`` (:339-687), ``SWAGModel`` (:690-908),
``save_swag`` / ``load_swag`` (:911-967).  The arithmetic is done by the sm_100a kernels of
``csrc/`` through the C ABI in ``include/bnnchaos.h``; PyTorch is used for device memory,
streams and -- on the single-call paths -- for drawing the normals with ``torch.randn`` /
``torch.randn_like`` in exactly the reference's order, so that the same torch generator
state yields the same draws as the reference on the same device (SURVEY.md section 0,
fact 5).  There is no CPU fallback: tensors on the CPU raise ``BnnChaosError``.

Not mirrored (outside the hot path): the Lightning trainer hooks, the
dataloaders / ``get_data``, ``VarModel.sample``, ``augment`` and the megno side channel
(``fix_megno=True``).  ``CustomOneCycleLR`` (:27-159) is mirrored for the pre-training phase
(find_minima.py): host-side scalars that drive the fused step.
"""
from __future__ import annotations

import io
import math
import pickle
import random
from collections import OrderedDict
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import BnnChaosError, ModelConfig, TrainHParams

EPSILON = 1e-5  # spock_reg_model.py:337

MEGNO_LOCATION = 7  # :370-373
MMR_LOCATION = (3, 6)
NAN_LOCATION = (38, 39, 40)
EPLUSMINUS_LOCATION = (1, 2, 4, 5)


class AttributeDict(dict):
    """Plain stand-in for pytorch_lightning.utilities.parsing.AttributeDict (the class the
    reference's checkpoints pickle their hparams as)."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, val):
        self[key] = val


class ScheduleFinished(ValueError):
    """The one-cycle schedule was stepped past its total (the reference raises a plain ValueError, :137-139, which
    find_minima.py:79-82 catches to end the run): a ValueError subclass, so reference-style ``except ValueError`` still
    works, while our trainers catch exactly this and let every other ValueError propagate."""


def one_cycle_lr_momentum(step_num, max_lr, total_steps, pct_start=0.3, div_factor=25.0, final_div_factor=1e4,
                          base_momentum=0.85, max_momentum=0.95, anneal_strategy="cos"):
    """(lr, momentum) of the reference's one-cycle schedule at optimizer step ``step_num`` (:131-158): up from
    max_lr/div_factor to max_lr over ``pct_start*total_steps - 1`` steps, down to max_lr/div_factor/final_div_factor
    over the rest, momentum moving the opposite way.  Past ``total_steps`` it raises ValueError like :137-139 --
    find_minima.py relies on that to end the run (find_minima.py:79-82)."""
    if step_num > total_steps:
        raise ScheduleFinished("Tried to step {} times. The specified number of total steps is {}".format(step_num + 1, total_steps))
    size_up = float(pct_start * total_steps) - 1
    size_down = float(total_steps - size_up) - 1
    lr0 = max_lr / div_factor
    lr_min = lr0 / final_div_factor
    if anneal_strategy == "cos":
        def anneal(a, b, pct):
            return b if pct >= 1.0 else b + (a - b) / 2.0 * (math.cos(math.pi * pct) + 1)
    elif anneal_strategy == "linear":
        def anneal(a, b, pct):
            return b if pct >= 1.0 else (b - a) * pct + a
    else:
        raise ValueError("anneal_strategy must by one of 'cos' or 'linear', instead got {}".format(anneal_strategy))
    if step_num <= size_up:
        pct = step_num / size_up
        return anneal(lr0, max_lr, pct), anneal(max_momentum, base_momentum, pct)
    pct = (step_num - size_up) / size_down
    return anneal(max_lr, lr_min, pct), anneal(base_momentum, max_momentum, pct)


class CustomOneCycleLR(torch.optim.lr_scheduler.LRScheduler):
    """Drop-in for the reference's scheduler (:27-159; "custom version of one-cycle learning rate to stop early"):
    same constructor arguments, same lr / momentum per step, same ValueError once stepped past ``swa_steps_start``."""

    def __init__(self, optimizer, max_lr, swa_steps_start, pct_start=0.3, anneal_strategy="cos", cycle_momentum=True,
                 base_momentum=0.85, max_momentum=0.95, div_factor=25.0, final_div_factor=1e4, last_epoch=-1):
        if not isinstance(swa_steps_start, int) or swa_steps_start <= 0:
            raise ValueError("Expected non-negative integer total_steps, but got {}".format(swa_steps_start))
        if not isinstance(pct_start, float) or pct_start < 0 or pct_start > 1:
            raise ValueError("Expected float between 0 and 1 pct_start, but got {}".format(pct_start))
        if anneal_strategy not in ("cos", "linear"):
            raise ValueError("anneal_strategy must by one of 'cos' or 'linear', instead got {}".format(anneal_strategy))
        self.total_steps = swa_steps_start
        self._cfg = dict(pct_start=pct_start, div_factor=div_factor, final_div_factor=final_div_factor,
                         base_momentum=base_momentum, max_momentum=max_momentum, anneal_strategy=anneal_strategy)
        self._max_lrs = list(max_lr) if isinstance(max_lr, (list, tuple)) else [max_lr] * len(optimizer.param_groups)
        if len(self._max_lrs) != len(optimizer.param_groups):
            raise ValueError("expected {} values for max_lr, got {}".format(len(optimizer.param_groups), len(self._max_lrs)))
        self.cycle_momentum = cycle_momentum
        if cycle_momentum:
            if "momentum" not in optimizer.defaults and "betas" not in optimizer.defaults:
                raise ValueError("optimizer must support momentum with `cycle_momentum` option enabled")
            self.use_beta1 = "betas" in optimizer.defaults
        if last_epoch == -1:
            for g, mx in zip(optimizer.param_groups, self._max_lrs):
                g["initial_lr"] = mx / div_factor
                g["max_lr"] = mx
                g["min_lr"] = g["initial_lr"] / final_div_factor
        super().__init__(optimizer, last_epoch)

    def get_lr(self):
        lrs = []
        for g, mx in zip(self.optimizer.param_groups, self._max_lrs):
            lr, mom = one_cycle_lr_momentum(self.last_epoch, mx, self.total_steps, **self._cfg)
            lrs.append(lr)
            if self.cycle_momentum:
                if self.use_beta1:
                    g["betas"] = (mom, g["betas"][1])
                else:
                    g["momentum"] = mom
        return lrs


def _param_stack(in_n, out_n, hidden, layers):
    """Parameter container with the reference's ``mlp()`` structure (:301-321) so that
    ``state_dict()`` keys and order are identical; it is never called for compute."""
    if layers < 1:
        raise NotImplementedError("layers=0 (a bare Linear, :309-310) is not on the compiled path")
    mods = [nn.Linear(in_n, hidden), nn.ReLU()]
    for _ in range(layers):
        mods += [nn.Linear(hidden, hidden), nn.ReLU()]
    mods += [nn.Linear(hidden, out_n)]
    return nn.Sequential(*mods)


class _FusedLoss(torch.autograd.Function):
    """Scalar training loss whose backward hands out the gradient the fused kernel already computed."""

    @staticmethod
    def forward(ctx, value, grad_flat, *params):
        ctx.grad_flat = grad_flat
        ctx.shapes = [tuple(p.shape) for p in params]
        return value.clone()

    @staticmethod
    def backward(ctx, gout):
        out, i = [], 0
        for shp in ctx.shapes:
            n = int(np.prod(shp)) if len(shp) else 1
            out.append((ctx.grad_flat[i:i + n] * gout).reshape(shp))
            i += n
        return (None, None, *out)


class VarModel(nn.Module):
    """Bayesian neural network predicting instability time (reference :339-687)."""

    def __init__(self, hparams):
        super().__init__()
        hparams = AttributeDict(hparams) if not isinstance(hparams, AttributeDict) else hparams
        hparams.setdefault("seed", 0)
        seed = hparams["seed"]
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)  # pl.seed_everything (:344)
        hparams.setdefault("include_derivatives", False)
        hparams.setdefault("time_series_features", 38 + 3)
        if hparams["time_series_features"] == 82:
            hparams["time_series_features"] = 41
        self.fix_megno = bool(hparams.get("fix_megno", False))
        self.fix_megno2 = bool(hparams.get("fix_megno2", False))
        self.include_angles = bool(hparams.get("include_angles", False))
        if self.fix_megno:
            raise NotImplementedError("fix_megno=True (megno mean/std side channel) is outside the hot path")

        self.n_features = hparams["time_series_features"] * (1 + int(hparams["include_derivatives"]))
        self.feature_nn = _param_stack(self.n_features, hparams["latent"], hparams["hidden"], hparams["in"])
        self.regress_nn = _param_stack(hparams["latent"] * 2, 2, hparams["hidden"], hparams["out"])
        self.input_noise_logvar = nn.Parameter(torch.zeros(self.n_features) - 2)
        self.summary_noise_logvar = nn.Parameter(torch.zeros(hparams["latent"] * 2) - 2)
        self.lowest = 0.1 if hparams.get("lower_std", False) else 0.5

        self.latents = None
        self.beta_in = hparams.get("beta_in", 1)
        self.beta_out = hparams.get("beta_out", 1)
        self.megno_location = MEGNO_LOCATION
        self.mmr_location = list(MMR_LOCATION)
        self.nan_location = list(NAN_LOCATION)
        self.eplusminus_location = list(EPLUSMINUS_LOCATION)

        hparams["scheduler_choice"] = "swa"
        hparams.setdefault("save_freq", 25)
        hparams.setdefault("eval_freq", 5)
        hparams.setdefault("momentum", 0.9)
        hparams.setdefault("weight_decay", 1e-4)
        hparams.setdefault("noisy_val", True)

        self.hparams = hparams
        self.steps = hparams["steps"]
        self.batch_size = hparams["batch_size"]
        self.lr = hparams["lr"]
        self.random_sample = bool(hparams.get("random_sample", False))
        if self.random_sample:
            raise NotImplementedError("random_sample (augment, :404-408) is outside the hot path")
        self.train_len = 78660
        self.test_len = 8740
        self._summary_kl = 0.0
        self.include_mmr = hparams["include_mmr"]
        self.include_nan = hparams["include_nan"]
        self.include_eplusminus = hparams.get("include_eplusminus", True)
        self.train_all = hparams.get("train_all", False)
        self._cur_summary = None
        self.ssX = None
        self.ssy = None
        self.current_epoch = 0
        self.global_step = 0
        # gradients come from the fused training kernel (training_step installs them through a custom
        # autograd.Function, so loss.backward() / clip_grad_norm_ / optimizer.step() work as in the reference)
        self._train_ws = None

    # ------------------------------------------------------------------ plumbing
    @property
    def device(self):
        return self.input_noise_logvar.device

    def zero_columns(self):
        """Columns the zero_* methods (:452-478) clear under this model's flags (:487-500)."""
        cols = []
        if self.fix_megno or self.fix_megno2:
            cols.append(self.megno_location)
        if not self.include_mmr:
            cols += self.mmr_location
        if not self.include_nan:
            cols += self.nan_location
        if not self.include_eplusminus:
            cols += self.eplusminus_location
        return sorted(set(cols))

    def config(self, n_times: int = 100, zero: bool = True) -> ModelConfig:
        hp = self.hparams
        mask = 0
        if zero:
            for c in self.zero_columns():
                mask |= 1 << c
        return ModelConfig(
            n_features=self.n_features, hidden=hp["hidden"], latent=hp["latent"], n_in_layers=hp["in"],
            n_out_layers=hp["out"], n_times=n_times, zero_mask=mask, lo_mu=4.0, hi_mu=12.0, lo_sd=self.lowest,
            hi_sd=6.0,
        )

    def _flat(self) -> torch.Tensor:
        return torch.cat([p.detach().reshape(-1) for p in self.state_dict().values()])

    def _packed(self, cfg: ModelConfig, theta: Optional[torch.Tensor] = None) -> torch.Tensor:
        lib = _lib.load()
        theta = (self._flat() if theta is None else theta).contiguous().float()
        _lib.require_cuda(theta, "model parameters")
        n_units = 1 if theta.dim() == 1 else theta.shape[0]
        P = lib.bnn_packed_param_count(cfg)
        if P < 0:
            _lib.check(int(P), "bnn_packed_param_count")
        out = torch.empty((n_units, P), device=theta.device, dtype=torch.float32)
        _lib.check(lib.bnn_pack_theta(cfg, _lib.ptr(theta), n_units, _lib.ptr(out), _lib.current_stream_ptr()),
                   "bnn_pack_theta")
        return out

    def _check_x(self, x):
        _lib.require_cuda(x, "x")
        if x.dim() != 3 or x.shape[-1] != self.n_features:
            raise ValueError(f"x must be [batch, time, {self.n_features}], got {tuple(x.shape)}")
        return x.contiguous().float()

    def _predict(self, x, thp, eps, eps_sum=None, cfg=None, want_summary=False, system_major=False):
        """[U,B,2] predictions (and [U,B,2L] summary statistics) for packed weights thp[U,P]."""
        lib = _lib.load()
        cfg = cfg or self.config(x.shape[1])
        U, B = thp.shape[0], x.shape[0]
        out = torch.empty((B, U, 2) if system_major else (U, B, 2), device=x.device, dtype=torch.float32)
        summ = torch.empty((U, B, 2 * self.hparams["latent"]), device=x.device) if want_summary else None
        dp = lambda t, name: _lib.dev_ptr(t, x.device, name)
        _lib.check(
            lib.bnn_predict(cfg, _lib.ptr(x), B, dp(thp, "packed weights"), U, dp(eps, "eps"), dp(eps_sum, "eps_sum"), 0,
                            0, 0, int(system_major), _lib.ptr(out), _lib.ptr(summ), None, _lib.current_stream_ptr()),
            "bnn_predict",
        )
        return out, summ

    # ------------------------------------------------------------------ reference API
    def compute_summary_stats(self, x):
        """:416-435.  x is used as given (no zero_* masks, like the reference method)."""
        x = self._check_x(x)
        L = self.hparams["latent"]
        with torch.cuda.device(x.device):
            cfg = self.config(x.shape[1], zero=False)
            eps1 = torch.randn((x.shape[0], L), device=x.device)
            eps2 = torch.randn((x.shape[0], L), device=x.device)
            eps = torch.cat((eps1, eps2), dim=1)[None].contiguous()
            _, summ = self._predict(x, self._packed(cfg), eps, cfg=cfg, want_summary=True)
        return summ[0]

    def gradforward(self, x, want_grad=True):
        """figures/feature_importance.py:93-131 (``gradforward`` / ``partforward``): the zero_* masks, then
        ``mu = predict_instability(compute_summary_stats(x))[0]`` with eps1, eps2 drawn like compute_summary_stats
        (no input / summary noise) and ``grad(mu.sum(), x)`` with respect to the masked input, all F columns.
        Returns (grad [B,T,F] or None, mu [B], sumsq [F] = (grad**2).sum((0, 1)))."""
        lib = _lib.load()
        x = self._check_x(x)
        B, T, F = x.shape
        L = self.hparams["latent"]
        with torch.cuda.device(x.device):
            cfg = self.config(T)
            eps1 = torch.randn((B, L), device=x.device)
            eps2 = torch.randn((B, L), device=x.device)
            eps = torch.cat((eps1, eps2), dim=1)[None].contiguous()
            theta = self._flat().contiguous().float()[None].contiguous()
            ws = torch.empty((lib.bnn_train_workspace_bytes(cfg, B, 1) + 3) // 4, device=x.device)
            g = torch.empty((1, B, T, F), device=x.device) if want_grad else None
            sumsq = torch.empty((1, F), device=x.device)
            mu = torch.empty((1, B), device=x.device)
            _lib.check(lib.bnn_saliency(cfg, 1, _lib.ptr(theta), _lib.ptr(x), B, _lib.ptr(eps), 0, _lib.ptr(g),
                                        _lib.ptr(sumsq), _lib.ptr(mu), _lib.ptr(ws), _lib.current_stream_ptr()),
                       "bnn_saliency")
        return (g[0] if want_grad else None), mu[0], sumsq[0]

    def predict_instability(self, summary_stats):
        """:437-442 -> (mu[B,1], std[B,1])."""
        lib = _lib.load()
        _lib.require_cuda(summary_stats, "summary_stats")
        s = summary_stats.contiguous().float()
        with torch.cuda.device(s.device):
            cfg = self.config()
            out = torch.empty((s.shape[0], 2), device=s.device, dtype=torch.float32)
            thp = self._packed(cfg)  # keep a reference while the kernel is enqueued
            _lib.check(lib.bnn_predict_instability(cfg, _lib.ptr(s), s.shape[0], _lib.ptr(thp),
                                                   _lib.ptr(out), _lib.current_stream_ptr()),
                       "bnn_predict_instability")
        return out[:, [0]], out[:, [1]]

    def forward(self, x, noisy_val=True):
        """:486-528.  RNG draw order as the reference: (input noise [B,T,F] if noisy), eps1,
        eps2 [B,L], (summary noise [B,2L] if noisy)."""
        lib = _lib.load()
        x = self._check_x(x)
        B, T, _ = x.shape
        L = self.hparams["latent"]
        with torch.cuda.device(x.device):
            eps_sum = None
            if noisy_val:
                eps_in = torch.randn_like(x)
                cfg_mask = self.config(T)
                xn = torch.empty_like(x)
                _lib.check(lib.bnn_add_input_noise(cfg_mask, _lib.ptr(x), _lib.ptr(eps_in),
                                                   _lib.ptr(self.input_noise_logvar.detach().contiguous()), B * T,
                                                   _lib.ptr(xn), _lib.current_stream_ptr()), "bnn_add_input_noise")
                x, cfg = xn, self.config(T, zero=False)
            else:
                cfg = self.config(T)
            eps1 = torch.randn((B, L), device=x.device)
            eps2 = torch.randn((B, L), device=x.device)
            eps = torch.cat((eps1, eps2), dim=1)[None].contiguous()
            if noisy_val:
                eps_sum = torch.randn((B, 2 * L), device=x.device)[None].contiguous()
            out, summ = self._predict(x, self._packed(cfg), eps, eps_sum, cfg=cfg, want_summary=True)
        self._cur_summary = summ[0]
        lv = self.summary_noise_logvar.detach()
        self._summary_kl = 0.5 * (summ[0] ** 2 + torch.exp(lv)[None, :] - lv[None, :] - 1)  # :515-520
        return out[0]


    def _lossfnc(self, testy, y):
        """:547-577 -> per-system loss [B]."""
        lib = _lib.load()
        _lib.require_cuda(testy, "testy")
        testy, y = testy.contiguous().float(), y.contiguous().float()
        loss = torch.empty(testy.shape[0], device=testy.device, dtype=torch.float32)
        with torch.cuda.device(testy.device):
            _lib.check(lib.bnn_nll_fwd_bwd(_lib.ptr(testy), _lib.ptr(y), testy.shape[0], _lib.ptr(loss), None, None,
                                           _lib.current_stream_ptr()), "bnn_nll_fwd_bwd")
        return loss

    def lossfnc(self, x, y, samples=1, noisy_val=True):
        """:579-583."""
        testy = self.forward(x, noisy_val=noisy_val)
        return self._lossfnc(testy, y).sum()

    # ------------------------------------------------------------------ training (reference :595-614, :722-732)
    def _fused_loss_and_grad(self, x, y, beta_in, beta_out):
        """Noisy forward + analytic backward of one batch in one kernel launch (bnn_train_step with
        apply_update=0).  Draw order as the reference's forward(noisy_val=True): eps_in [B,T,F] (:445), eps1,
        eps2 [B,L] (:426-427), eps_sum [B,2L] (:449).  Returns (metrics[8], grad[d]) on the device."""
        lib = _lib.load()
        x = self._check_x(x)
        _lib.require_cuda(y, "y")
        y = y.contiguous().float()
        B, T, _ = x.shape
        L = self.hparams["latent"]
        dev = x.device
        with torch.cuda.device(dev):
            eps_in = torch.randn_like(x)
            eps12 = torch.cat((torch.randn((B, L), device=dev), torch.randn((B, L), device=dev)), dim=1).contiguous()
            eps_sum = torch.randn((B, 2 * L), device=dev)
            cfg = self.config(T)
            theta = self._flat().contiguous().float()
            hp = TrainHParams(lr=0.0, momentum=0.0, weight_decay=0.0, clip_norm=float("inf"), beta_in=float(beta_in),
                              beta_out=float(beta_out), first_step=1, apply_update=0)
            nbytes = lib.bnn_train_workspace_bytes(cfg, B, 1)
            if self._train_ws is None or self._train_ws.numel() * 4 < nbytes or self._train_ws.device != dev:
                self._train_ws = torch.empty((nbytes + 3) // 4, device=dev, dtype=torch.float32)
            grad = torch.empty_like(theta)
            metrics = torch.empty(8, device=dev, dtype=torch.float32)
            _lib.check(
                lib.bnn_train_step(cfg, hp, 1, _lib.ptr(theta), None, _lib.ptr(x), _lib.ptr(y), None, B,
                                   _lib.ptr(eps_in), _lib.ptr(eps12), _lib.ptr(eps_sum), 0, 0, _lib.ptr(grad),
                                   _lib.ptr(metrics), _lib.ptr(self._train_ws), _lib.current_stream_ptr()),
                "bnn_train_step",
            )
        return metrics, grad

    def _training_result(self, batch, beta_in, beta_out):
        X_sample, y_sample = batch
        n = len(X_sample)
        metrics, grad = self._fused_loss_and_grad(X_sample, y_sample, beta_in, beta_out)
        params = list(self.parameters())  # same order as state_dict() / flatten() (:734-746)
        total = _FusedLoss.apply(metrics[1] * n, grad, *params)
        logs = {"train_loss_no_reg": metrics[0], "train_loss_with_reg": metrics[1], "input_kl": metrics[2],
                "summary_kl": metrics[3]}
        return {"loss": total, "log": logs}

    def training_step(self, batch, batch_idx):
        """:595-614 (KL annealing over the first 30 % of the steps)."""
        fraction = self.global_step / self.hparams["steps"]
        return self._training_result(batch, min(1, fraction / 0.3) * self.beta_in, min(1, fraction / 0.3) * self.beta_out)

    def configure_optimizers(self):
        """:630-644: SGD(momentum) under the custom one-cycle schedule over 0.9 * steps."""
        opt1 = torch.optim.SGD(self.parameters(), lr=self.lr, momentum=self.hparams["momentum"],
                               weight_decay=self.hparams["weight_decay"])
        assert self.hparams["scheduler_choice"] == "swa"
        scheduler = CustomOneCycleLR(opt1, self.lr, int(0.9 * self.steps), final_div_factor=1e4)
        return [opt1], [{"scheduler": scheduler, "name": "swa_lr", "interval": "steps"}]

    def input_kl(self):
        """:585-590 (41 elements: plain tensor arithmetic)."""
        lv = self.input_noise_logvar.detach()
        return 0.5 * (torch.exp(lv) - lv - 1).sum()

    def summary_kl(self):
        """:592-593."""
        return self._summary_kl.sum()


class SWAGModel(VarModel):
    """SWAG moment collection and weight sampling (reference :690-908)."""

    def init_params(self, swa_params):
        self.swa_params = swa_params
        self.swa_params.setdefault("swa_lr", 0.001)
        self.swa_params.setdefault("swa_start", 1000)
        self.swa_params.setdefault("swa_recording_lr_factor", 0.5)
        self.n_models = 0
        self.w_avg = None
        self.w2_avg = None
        self.pre_D = None
        self.K = self.swa_params.get("K", 20)
        self.c = self.swa_params.get("c", 2)
        self.swa_params["c"] = self.c
        self.swa_params["K"] = self.K
        self._momentum = None
        self._first_step = True
        return self

    def configure_optimizers(self):
        """:709-720."""
        opt1 = torch.optim.SGD(self.parameters(), lr=self.swa_params["swa_lr"], momentum=self.hparams["momentum"],
                               weight_decay=self.hparams["weight_decay"])
        scheduler = torch.optim.lr_scheduler.MultiStepLR(opt1, [self.swa_params["swa_start"]],
                                                         self.swa_params["swa_recording_lr_factor"])
        return [opt1], [{"scheduler": scheduler, "name": "swa_record_lr", "interval": "steps"}]

    def training_step(self, batch, batch_idx):
        """:722-732 (no KL annealing in the SWAG phase).  ``res['loss'].backward()`` installs the analytic
        gradient of the fused kernel in ``p.grad``; clip_grad_norm_ and the optimizer then act as usual."""
        return self._training_result(batch, self.beta_in, self.beta_out)

    def validation_step(self, batch, batch_idx):
        """:787-799: loss at the current weights and at w_avg."""
        X_sample, y_sample = batch
        noisy = self.hparams["noisy_val"]
        loss = self.lossfnc(X_sample, y_sample, noisy_val=noisy) / self.test_len
        if self.w_avg is None:
            swa_loss = loss
        else:
            tmp = self.flatten()
            self.load(self.w_avg)
            swa_loss = self.lossfnc(X_sample, y_sample, noisy_val=noisy) / self.test_len
            self.load(tmp)
        return {"val_loss": loss, "swa_loss": swa_loss}

    def validation_epoch_end(self, outputs):
        """:801-813: the collection gate reads hparams['swa_start']."""
        avg_loss = torch.stack([x["val_loss"] for x in outputs]).sum()
        swa_avg_loss = torch.stack([x["swa_loss"] for x in outputs]).sum()
        logs = {"val_loss_no_reg": avg_loss, "swa_loss_no_reg": swa_avg_loss}
        if self.global_step > self.hparams["swa_start"]:
            self.aggregate_model()
        return {"val_loss": avg_loss, "log": logs}

    # flatten / load (:734-761)
    def flatten(self):
        return self._flat()

    def load(self, p_vec):
        i = 0
        with torch.no_grad():
            for p in self.state_dict().values():
                n = p.numel()
                p.copy_(p_vec[i:i + n].reshape(p.shape))
                i += n

    # aggregate_model (:763-785)
    def aggregate_model(self):
        lib = _lib.load()
        cur_w = self.flatten().contiguous()
        _lib.require_cuda(cur_w, "model parameters")
        d = cur_w.numel()
        dev = cur_w.device
        if self.w_avg is None:
            self.w_avg = torch.zeros(d, device=dev)
            self.w2_avg = torch.zeros(d, device=dev)
            self._pre_D_buf = torch.zeros((d, self.K), device=dev)
            self._n_models_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._n_cols_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self._n_cols = 0
        elif not hasattr(self, "_pre_D_buf") or self._pre_D_buf is None:
            self._adopt_stats(dev)
        with torch.cuda.device(dev):
            _lib.check(
                lib.bnn_swag_collect(_lib.ptr(cur_w), d, 1, self.K, _lib.ptr(self.w_avg), _lib.ptr(self.w2_avg),
                                     _lib.ptr(self._pre_D_buf), _lib.ptr(self._n_models_dev),
                                     _lib.ptr(self._n_cols_dev), int(self.current_epoch), int(self.c),
                                     _lib.current_stream_ptr()),
                "bnn_swag_collect",
            )
        if self._n_cols == 0 or self.current_epoch % self.c == 0:
            self._n_cols = min(self._n_cols + 1, self.K)
        self.pre_D = self._pre_D_buf[:, : self._n_cols]
        self.n_models += 1

    def _adopt_stats(self, dev):
        """Statistics assigned from outside (load_swag): build the fixed [d,K] buffer."""
        self.w_avg = self.w_avg.to(dev).contiguous().float()
        self.w2_avg = self.w2_avg.to(dev).contiguous().float()
        ncol = self.pre_D.shape[1]
        self._pre_D_buf = torch.zeros((self.w_avg.numel(), self.K), device=dev)
        self._pre_D_buf[:, :ncol] = self.pre_D.to(dev)
        self._n_cols = ncol
        self._n_models_dev = torch.tensor([self.n_models], dtype=torch.int32, device=dev)
        self._n_cols_dev = torch.tensor([ncol], dtype=torch.int32, device=dev)
        self.pre_D = self._pre_D_buf[:, :ncol]

    def _stats_on(self, dev):
        if self.w_avg is None:
            raise BnnChaosError("no SWAG statistics: call aggregate_model() or load_swag() first")
        if self.w_avg.device != dev or self.pre_D.device != dev or not self.pre_D.is_contiguous():
            self.w_avg = self.w_avg.to(dev).contiguous().float()
            self.w2_avg = self.w2_avg.to(dev).contiguous().float()
            self.pre_D = self.pre_D.to(dev).contiguous().float()
        if self.pre_D.shape[1] != self.K:
            # the reference fails in D @ z_2 when fewer than K deviations were recorded (:835)
            raise RuntimeError(
                f"mat1 and mat2 shapes cannot be multiplied ({self.pre_D.shape[0]}x{self.pre_D.shape[1]} and {self.K}x1)"
            )
        return self.w_avg, self.w2_avg, self.pre_D.contiguous()

    # sample_weights (:815-838)
    def sample_weights(self, scale=1):
        lib = _lib.load()
        dev = self.device
        _lib.require_cuda(self.input_noise_logvar, "model parameters")
        w_avg, w2_avg, pre_D = self._stats_on(dev)
        d = w_avg.shape[0]
        with torch.cuda.device(dev):
            z_1 = torch.randn((1, d), device=dev)        # :830
            z_2 = torch.randn((self.K, 1), device=dev)   # :831
            theta = torch.empty((1, d), device=dev)
            cfg = self.config()
            _lib.check(
                lib.bnn_swag_sample(cfg, _lib.ptr(w_avg), _lib.ptr(w2_avg), _lib.ptr(pre_D), 1, self.K, None, 1, 0, 1,
                                    float(scale), 0, _lib.ptr(z_1), _lib.ptr(z_2.reshape(1, self.K).contiguous()),
                                    _lib.ptr(theta), None, _lib.current_stream_ptr()),
                "bnn_swag_sample",
            )
        self.load(theta[0])

    def forward_swag(self, x, scale=0.5):
        """:840-876 -- sample weights (they stay loaded in the module, like the reference),
        masks, summary statistics, head; also records _summary_kl."""
        x = self._check_x(x)
        self.sample_weights(scale=scale)
        out = self.forward(x, noisy_val=False)
        return out

    def forward_swag_fast(self, x, scale=0.5):
        """:878-908."""
        x = self._check_x(x)
        self.sample_weights(scale=scale)
        B, T, _ = x.shape
        L = self.hparams["latent"]
        with torch.cuda.device(x.device):
            cfg = self.config(T)
            eps1 = torch.randn((B, L), device=x.device)
            eps2 = torch.randn((B, L), device=x.device)
            eps = torch.cat((eps1, eps2), dim=1)[None].contiguous()
            out, _ = self._predict(x, self._packed(cfg), eps, cfg=cfg)
        return out[0]


# ----------------------------------------------------------------------------------------
# save_swag / load_swag (:911-967): same dict, same keys.
# ----------------------------------------------------------------------------------------
def save_swag(swag_model, path):
    # hparams go out as a plain dict: the reference pickles Lightning's AttributeDict, a class this package cannot name in
    # a pickle without Lightning installed; the reference's load_swag only needs item access (SWAGModel(save_items['hparams']),
    # :925), and a file without foreign globals also loads under torch.load(weights_only=True).
    save_items = {
        "hparams": dict(swag_model.hparams),
        "swa_params": swag_model.swa_params,
        "w_avg": swag_model.w_avg.cpu(),
        "w2_avg": swag_model.w2_avg.cpu(),
        "pre_D": swag_model.pre_D.cpu().contiguous(),
    }
    torch.save(save_items, path)


class _SafeUnpickler(pickle.Unpickler):
    """Only the four globals the reference's SWAG pickles contain,
    plus what torch.save of this module's own dicts adds."""

    _ALLOWED = {
        ("collections", "OrderedDict"): OrderedDict,
        ("pytorch_lightning.utilities.parsing", "AttributeDict"): AttributeDict,
        ("bnn_chaos_model_b200.spock_reg_model", "AttributeDict"): AttributeDict,
    }

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return self._ALLOWED[(module, name)]
        if module == "torch._utils" and name in ("_rebuild_tensor_v2", "_rebuild_tensor"):
            return getattr(torch._utils, name)
        if module == "torch" and name in ("FloatStorage", "LongStorage", "IntStorage", "DoubleStorage"):
            return getattr(torch, name)
        raise pickle.UnpicklingError(f"global {module}.{name} is not allowed in a SWAG checkpoint")


def _safe_load(f, **kw):
    return _SafeUnpickler(f, **kw).load()


def _safe_loads(b, **kw):
    return _SafeUnpickler(io.BytesIO(b), **kw).load()


class _SafePickleModule:
    """pickle_module for torch.load: EVERY entry point goes through the allow-listed unpickler (torch's legacy,
    non-zip loader calls pickle_module.load() directly on the magic number / protocol / sys_info records)."""

    __name__ = "bnn_safe_pickle"
    Unpickler = _SafeUnpickler
    load = staticmethod(_safe_load)
    loads = staticmethod(_safe_loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    UnpicklingError = pickle.UnpicklingError


class _ScalerUnpickler(pickle.Unpickler):
    """``*_ssX.pkl`` (spock_reg_model.py:958-963): a pickled sklearn StandardScaler -- the class itself plus the numpy
    array / scalar reconstructors, nothing else."""

    _ALLOWED = {
        ("sklearn.preprocessing._data", "StandardScaler"), ("sklearn.preprocessing.data", "StandardScaler"),
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy", "ndarray"), ("numpy", "dtype"),
    }

    def find_class(self, module, name):
        if (module, name) not in self._ALLOWED:
            raise pickle.UnpicklingError(f"global {module}.{name} is not allowed in a StandardScaler pickle")
        if name == "StandardScaler":
            from sklearn.preprocessing import StandardScaler

            return StandardScaler
        import importlib

        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core") if hasattr(np, "_core") else module
        return getattr(importlib.import_module(module), name)


def fixed_v50_scaler():
    """The StandardScaler constants hard-coded in load_swag for 'v50' checkpoints (:931-957)."""
    from sklearn.preprocessing import StandardScaler

    from .synth import SSX_MEAN, SSX_SCALE

    ssX = StandardScaler()
    ssX.scale_ = SSX_SCALE.copy()
    ssX.mean_ = SSX_MEAN.copy()
    ssX.var_ = ssX.scale_ ** 2
    return ssX


def load_swag(path):
    save_items = torch.load(path, map_location="cpu", pickle_module=_SafePickleModule, weights_only=False)
    swag_model = SWAGModel(save_items["hparams"]).init_params(save_items["swa_params"])
    swag_model.w_avg = save_items["w_avg"]
    swag_model.w2_avg = save_items["w2_avg"]
    swag_model.pre_D = save_items["pre_D"]
    if "v50" in str(path):
        swag_model.ssX = fixed_v50_scaler()
    else:
        ssX_file = str(path)[:-4] + "_ssX.pkl"
        try:
            with open(ssX_file, "rb") as f:
                swag_model.ssX = _ScalerUnpickler(f).load()
        except FileNotFoundError:
            print(f"ssX file not found! {ssX_file}")
    return swag_model
