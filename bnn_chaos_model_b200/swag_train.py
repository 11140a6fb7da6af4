"""SWAG training of several seed models at once on one GPU (BASELINE config 4).

Replaces, for a group of seeds, the reference's sequential per-seed runs (train.sh:3-6) of
``run_swag.py``: Lightning ``Trainer.fit`` (run_swag.py:74-82) looping over
``SWAGModel.training_step`` (spock_reg_model.py:722-732) + backward + ``gradient_clip_val`` +
``SGD(momentum)`` (:709-711), per-epoch ``validation_step`` at w and w_avg (:787-799) and
``aggregate_model`` once ``global_step > swa_start`` (:801-813, :763-785).

Everything stays on the device: the data set, one flat weight / momentum vector per seed, the SWAG
statistics.  A step is ONE ``bnn_train_step`` call for all seeds of this rank (per-seed shuffles are
an index tensor, the four noise tensors are counter-based Philox draws); an epoch ends with one
``bnn_eval_loss`` over both weight sets and one ``bnn_swag_collect``.  Seeds are independent: ranks
own disjoint seed groups and exchange nothing.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import BnnChaosError, TrainHParams
from .multiswag import shard_range
from .spock_reg_model import SWAGModel, VarModel, one_cycle_lr_momentum


def seeds_of_rank(n_seeds: int, rank: int, world: int):
    """Contiguous seed block of a rank: 30 seeds over 8 GPUs -> 4,4,4,4,4,4,3,3."""
    lo, hi = shard_range(n_seeds, rank, world)
    return list(range(lo, hi))


class MultiSeedSWAGTrainer:
    def __init__(self, models: Sequence[SWAGModel], X_train, y_train, X_val=None, y_val=None, batch_size=2000,
                 device=None, seed=0, swa_start: Optional[int] = None, noisy_val: bool = False):
        if len(models) == 0:
            raise ValueError("no seed models")
        self.models: List[SWAGModel] = list(models)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if self.device.type != "cuda":
            raise BnnChaosError("MultiSeedSWAGTrainer needs a CUDA device: there is no CPU fallback")
        m0 = self.models[0]
        for m in self.models:
            if m.n_features != m0.n_features or m.zero_columns() != m0.zero_columns() or m.K != m0.K or m.c != m0.c:
                raise ValueError("seed models must share the architecture, flags, K and c")
        dev = self.device
        self.S = len(self.models)
        self.K, self.c = m0.K, m0.c
        self.X = X_train.to(dev).contiguous().float()
        self.y = y_train.to(dev).contiguous().float()
        self.Xv = None if X_val is None else X_val.to(dev).contiguous().float()
        self.yv = None if y_val is None else y_val.to(dev).contiguous().float()
        self.batch_size = int(batch_size)
        self.seed = int(seed)
        self.noisy_val = noisy_val
        if noisy_val:
            raise NotImplementedError("batched validation is noise-free (bnn_eval_loss); use SWAGModel.validation_step")
        self.swa_start = int(m0.hparams["swa_start"] if swa_start is None else swa_start)  # :808 reads hparams
        self.theta = torch.stack([m._flat().detach().float() for m in self.models]).to(dev).contiguous()  # [S,d]
        self.momentum = torch.zeros_like(self.theta)
        d = self.theta.shape[1]
        self.w_avg = torch.zeros((self.S, d), device=dev)
        self.w2_avg = torch.zeros((self.S, d), device=dev)
        self.pre_D = torch.zeros((self.S, d, self.K), device=dev)
        self.n_models = torch.zeros(self.S, dtype=torch.int32, device=dev)
        self.n_cols = torch.zeros(self.S, dtype=torch.int32, device=dev)
        self.metrics = torch.zeros((self.S, 8), device=dev)
        self.global_step = 0
        self.current_epoch = 0
        self.first_step = True
        self._ws = None
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(self.seed)
        self.lr = float(m0.swa_params["swa_lr"])
        self.hp = dict(momentum=float(m0.hparams["momentum"]), weight_decay=float(m0.hparams["weight_decay"]),
                       clip_norm=0.1 * d, beta_in=float(m0.beta_in), beta_out=float(m0.beta_out))  # run_swag.py:61

    # ------------------------------------------------------------------ one optimisation step, all seeds
    def train_step(self, batch_index: Optional[torch.Tensor], B: int, lr: Optional[float] = None):
        lib = _lib.load()
        cfg = self.models[0].config(self.X.shape[1])
        hp = TrainHParams(lr=self.lr if lr is None else float(lr), first_step=int(self.first_step), apply_update=1,
                          **self.step_hparams())
        with torch.cuda.device(self.device):
            nbytes = lib.bnn_train_workspace_bytes(cfg, B, self.S)
            if self._ws is None or self._ws.numel() * 4 < nbytes:
                self._ws = torch.empty((nbytes + 3) // 4, device=self.device, dtype=torch.float32)
            _lib.check(
                lib.bnn_train_step(cfg, hp, self.S, _lib.ptr(self.theta), _lib.ptr(self.momentum), _lib.ptr(self.X),
                                   _lib.ptr(self.y), _lib.ptr(batch_index), B, None, None, None, self.seed,
                                   self.global_step, None, _lib.ptr(self.metrics), _lib.ptr(self._ws),
                                   _lib.current_stream_ptr()),
                "bnn_train_step",
            )
        self.first_step = False
        self.global_step += 1

    def step_hparams(self):
        """momentum, weight decay, clip and KL weights of the coming step (constant in the SWAG phase, :722-732)."""
        return self.hp

    def epoch_batches(self):
        """Per-seed shuffles of the training set (DataLoader(shuffle=True), :276): [n_batches] of ([S,B] int32, B)."""
        n = self.X.shape[0]
        perms = torch.stack([torch.randperm(n, device=self.device, generator=self._gen) for _ in range(self.S)])
        perms = perms.to(torch.int32)
        out = []
        for lo in range(0, n, self.batch_size):  # the last batch is ragged, like DataLoader without drop_last
            hi = min(lo + self.batch_size, n)
            out.append((perms[:, lo:hi].contiguous(), hi - lo))
        return out

    def train_epoch(self):
        for idx, B in self.epoch_batches():
            self.train_step(idx, B)

    # ------------------------------------------------------------------ validation + moment collection
    def validation_losses(self):
        """(val_loss[S], swa_loss[S]) = lossfnc(noisy_val=False) summed over the validation set, per system
        normalised by its size like :789; evaluated at theta and at w_avg in one launch."""
        lib = _lib.load()
        m0 = self.models[0]
        cfg = m0.config(self.Xv.shape[1])
        have_avg = bool((self.n_models > 0).all())
        thetas = torch.cat([self.theta, self.w_avg]) if have_avg else self.theta
        thp = m0._packed(cfg, thetas)
        U, B = thp.shape[0], self.Xv.shape[0]
        with torch.cuda.device(self.device):
            out = torch.empty((U, B, 2), device=self.device)
            loss = torch.empty(U, device=self.device)
            _lib.check(
                lib.bnn_eval_loss(cfg, _lib.ptr(self.Xv), _lib.ptr(self.yv), B, _lib.ptr(thp), U, None,
                                  self.seed ^ 0x5EED, _lib.ptr(out), _lib.ptr(loss), None, _lib.current_stream_ptr()),
                "bnn_eval_loss",
            )
        loss = loss / B
        return (loss[: self.S], loss[self.S:] if have_avg else loss[: self.S])

    def collect(self):
        """aggregate_model (:763-785) for every seed."""
        lib = _lib.load()
        d = self.theta.shape[1]
        with torch.cuda.device(self.device):
            _lib.check(
                lib.bnn_swag_collect(_lib.ptr(self.theta), d, self.S, self.K, _lib.ptr(self.w_avg), _lib.ptr(self.w2_avg),
                                     _lib.ptr(self.pre_D), _lib.ptr(self.n_models), _lib.ptr(self.n_cols),
                                     int(self.current_epoch), int(self.c), _lib.current_stream_ptr()),
                "bnn_swag_collect",
            )

    def fit(self, epochs: int, validate: bool = True, check_nan_every: int = 1):
        """Trainer.fit (run_swag.py:74-82): epochs of training, validation, and collection after swa_start."""
        logs = []
        for _ in range(epochs):
            self.train_epoch()
            entry = {"epoch": self.current_epoch, "global_step": self.global_step}
            if validate and self.Xv is not None:
                v, s = self.validation_losses()
                entry["val_loss_no_reg"], entry["swa_loss_no_reg"] = v.cpu(), s.cpu()
            if self.global_step > self.swa_start:  # :808
                self.collect()
            if check_nan_every and (self.current_epoch % check_nan_every == 0):
                if bool((self.metrics[:, 6] != 0).any()):  # terminate_on_nan (run_swag.py:78), once per epoch
                    raise ValueError(f"non-finite training loss at epoch {self.current_epoch}")
            logs.append(entry)
            self.current_epoch += 1
        return logs

    # ------------------------------------------------------------------ hand the results back
    def export(self) -> List[SWAGModel]:
        """Install weights and statistics into the SWAGModel mirrors (ready for save_swag, :911-930)."""
        n_cols = self.n_cols.cpu().tolist()
        n_mod = self.n_models.cpu().tolist()
        for i, m in enumerate(self.models):
            m.to(self.device)
            m.load(self.theta[i])
            if n_mod[i] > 0:
                m.w_avg = self.w_avg[i].clone()
                m.w2_avg = self.w2_avg[i].clone()
                m.pre_D = self.pre_D[i, :, : n_cols[i]].clone()
                m.n_models = n_mod[i]
                m._pre_D_buf = None
            m.global_step, m.current_epoch = self.global_step, self.current_epoch
        return self.models


class MultiSeedPretrainer(MultiSeedSWAGTrainer):
    """The pre-training phase of several seed models at once (find_minima.py:26-84; 300,000 of the 350,000 steps of
    a seed in train.sh:3-6).  Same fused step as the SWAG phase; what changes per step are host-side scalars:

    * KL annealing of ``VarModel.training_step`` (:595-598): both KL weights ramp over the first 30 % of ``steps``;
    * the custom one-cycle schedule (:27-159, :634) over ``int(0.9 * steps)`` optimizer steps: lr AND momentum
      (0.95 -> 0.85 -> 0.95) change every step;
    * no moment collection; validation once per epoch, the weights of the best epoch are kept
      (ModelCheckpoint, find_minima.py:67,82);
    * the run ends when the scheduler is stepped past its total -- the reference's ValueError
      (:137-139, caught at find_minima.py:79-82) -- after which every seed is reset to its best checkpoint.
    """

    def __init__(self, models: Sequence[VarModel], X_train, y_train, X_val=None, y_val=None, batch_size=None,
                 device=None, seed=0):
        m0 = models[0]
        for m in models:  # the SWAG bookkeeping of the base class reads these; unused here
            if not hasattr(m, "K"):
                m.K, m.c, m.swa_params = 1, 1, {"swa_lr": m.lr}
        super().__init__(models, X_train, y_train, X_val, y_val, batch_size=batch_size or m0.batch_size, device=device,
                         seed=seed, swa_start=1 << 62)
        self.steps = int(m0.steps)
        self.max_lr = float(m0.lr)
        self.total_sched = int(0.9 * self.steps)
        self.best_val = torch.full((self.S,), float("inf"), device=self.device)
        self.best_theta = self.theta.clone()
        self.finished = False

    def schedule(self, step: Optional[int] = None):
        """(lr, momentum, beta_in, beta_out) of optimizer step ``step`` (default: the coming one)."""
        g = self.global_step if step is None else int(step)
        lr, mom = one_cycle_lr_momentum(g, self.max_lr, self.total_sched)
        f = min([1, (g / self.steps) / 0.3])
        return lr, mom, f * self.hp["beta_in"], f * self.hp["beta_out"]

    def step_hparams(self):
        _, mom, b_in, b_out = self.schedule()
        return dict(self.hp, momentum=mom, beta_in=b_in, beta_out=b_out)

    def train_step(self, batch_index, B, lr=None):
        if lr is None:
            lr = self.schedule()[0]  # raises ValueError past the schedule's end, like scheduler.step() does
        super().train_step(batch_index, B, lr)

    def fit(self, epochs: Optional[int] = None, validate: bool = True, check_nan_every: int = 1):
        """Trainer.fit of find_minima.py: epochs = 1 + steps / steps_per_epoch unless given; stops at the schedule's end."""
        if epochs is None:
            epochs = int(1 + self.steps / len(self.epoch_batches()))
        logs = []
        for _ in range(epochs):
            try:
                self.train_epoch()
            except ValueError:
                self.finished = True
            entry = {"epoch": self.current_epoch, "global_step": self.global_step}
            if not self.finished and validate and self.Xv is not None:
                v, _ = self.validation_losses()
                better = v < self.best_val
                self.best_val = torch.where(better, v, self.best_val)
                self.best_theta = torch.where(better[:, None], self.theta, self.best_theta)
                entry["val_loss_no_reg"] = v.cpu()
            if check_nan_every and bool((self.metrics[:, 6] != 0).any()):
                raise ValueError(f"non-finite training loss at epoch {self.current_epoch}")
            logs.append(entry)
            self.current_epoch += 1
            if self.finished:
                if self.Xv is not None and bool(torch.isfinite(self.best_val).all()):
                    self.theta.copy_(self.best_theta)  # load_state_dict(best_model_path), find_minima.py:82
                break
        return logs

    def export(self):
        for i, m in enumerate(self.models):
            m.to(self.device)
            m.load(self.theta[i]) if hasattr(m, "load") else _load_flat(m, self.theta[i])
            m.global_step, m.current_epoch = self.global_step, self.current_epoch
        return self.models


def _load_flat(model: VarModel, p_vec: torch.Tensor):
    """SWAGModel.load (:748-761) for a plain VarModel: split the flat vector over the state_dict in order."""
    off = 0
    sd = model.state_dict()
    for k, v in sd.items():
        n = v.numel()
        sd[k] = p_vec[off:off + n].reshape(v.shape).to(v.device)
        off += n
    model.load_state_dict(sd)
