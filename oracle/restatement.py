"""TEST INFRASTRUCTURE ONLY -- the parity oracle, never the product path.

A CPU restatement (torch fp32, functional form, explicit noise tensors) of the
reference's MultiSWAG posterior-predictive and SWAG-training path.  Every
function cites the lines of ``/root/reference/spock_reg_model.py`` it follows.
The reference delegates its arithmetic to PyTorch (``pytorch=1.5.1`` in
``environment.yml:12``; torch 2.11 CPU is what this image has), so the oracle
uses the same torch CPU operators in the same order -- it is *not* an
independent numerical implementation, it is the reference with the RNG draws
made explicit and the ``nn.Module`` state made a flat vector.

Parity pin: the reference has NO tests and NO golden vectors of its own
(SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, run in the build container through ``oracle/ref_shim.py`` (unmodified
``spock_reg_model.py`` + the shipped ``pretrained/`` v50 checkpoints) --
see ``oracle/make_golden.py`` (generator) and ``tests/test_oracle_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

EPSILON = 1e-5  # spock_reg_model.py:337


# --------------------------------------------------------------------------------------
# Model description / flat parameter layout (spock_reg_model.py:359-362, :734-761)
# --------------------------------------------------------------------------------------
@dataclass
class ModelSpec:
    n_features: int = 41  # hparams['time_series_features'] (:348-358)
    hidden: int = 40
    latent: int = 20
    n_in: int = 1  # hparams['in']  -> number of hidden->hidden layers in feature_nn
    n_out: int = 1  # hparams['out'] -> same for regress_nn
    lowest: float = 0.5  # 0.1 if hparams['lower_std'] (:363-365)
    zero_cols: Tuple[int, ...] = ()  # columns zeroed by the include_* / fix_megno flags
    beta_in: float = 1.0
    beta_out: float = 1.0

    @staticmethod
    def from_hparams(hp) -> "ModelSpec":
        g = lambda k, d: hp[k] if k in hp else d
        F = g("time_series_features", 41)
        if F == 82:
            F = 41
        F = F * (1 + int(g("include_derivatives", False)))
        if g("fix_megno", False):
            raise NotImplementedError("fix_megno=True (megno side channel) is outside the hot path")
        zero: List[int] = []
        if g("fix_megno2", False):
            zero += [7]  # zero_megno (:452-457, :487-491)
        if not hp["include_mmr"]:
            zero += [3, 6]  # (:459-464)
        if not hp["include_nan"]:
            zero += [38, 39, 40]  # (:466-471)
        if not g("include_eplusminus", True):
            zero += [1, 2, 4, 5]  # (:473-478)
        return ModelSpec(
            n_features=F,
            hidden=hp["hidden"],
            latent=hp["latent"],
            n_in=hp["in"],
            n_out=hp["out"],
            lowest=0.1 if g("lower_std", False) else 0.5,
            zero_cols=tuple(sorted(set(zero))),
            beta_in=g("beta_in", 1),
            beta_out=g("beta_out", 1),
        )

    # state_dict() order (:734-746): the two logvar Parameters first (registered at :361-362
    # after the Sequentials? no -- nn.Module lists parameters before sub-modules), then
    # feature_nn.{0,2,..}.{weight,bias}, then regress_nn.*
    def layout(self) -> List[Tuple[str, Tuple[int, ...]]]:
        F, H, L = self.n_features, self.hidden, self.latent
        out = [("input_noise_logvar", (F,)), ("summary_noise_logvar", (2 * L,))]

        def mlp(prefix, n_in, n_out, layers):
            # mlp() (:301-321): Linear(in,H) act, `layers` x [Linear(H,H) act], Linear(H,out)
            if layers < 1:
                raise NotImplementedError("layers=0 (a bare Linear, :309-310) is not on the v50 path")
            dims = [n_in] + [H] * (layers + 1) + [n_out]
            res = []
            for i in range(len(dims) - 1):
                res.append((f"{prefix}.{2 * i}.weight", (dims[i + 1], dims[i])))
                res.append((f"{prefix}.{2 * i}.bias", (dims[i + 1],)))
            return res

        out += mlp("feature_nn", F, L, self.n_in)
        out += mlp("regress_nn", 2 * L, 2, self.n_out)
        return out

    def offsets(self) -> Dict[str, Tuple[int, Tuple[int, ...]]]:
        off, res = 0, {}
        for name, shape in self.layout():
            res[name] = (off, shape)
            off += int(np.prod(shape))
        return res

    @property
    def d(self) -> int:
        return sum(int(np.prod(s)) for _, s in self.layout())


def unflatten(spec: ModelSpec, theta: torch.Tensor) -> Dict[str, torch.Tensor]:
    """SWAGModel.load (:748-761): split the flat vector in state_dict order."""
    return {k: theta[o : o + int(np.prod(s))].reshape(s) for k, (o, s) in spec.offsets().items()}


def flatten(spec: ModelSpec, params: Dict[str, torch.Tensor]) -> torch.Tensor:
    """SWAGModel.flatten (:734-746)."""
    return torch.cat([params[k].reshape(-1) for k, _ in spec.layout()])


# --------------------------------------------------------------------------------------
# SWAG weight sampling (spock_reg_model.py:815-838)
# --------------------------------------------------------------------------------------
def sample_weights(w_avg, w2_avg, pre_D, K: int, scale: float, z1, z2, dense_diag: bool = False):
    """theta for explicit draws z1[1,d] (first randn, :830) and z2[K,1] (second, :831).

    ``dense_diag=True`` builds the d x d diagonal matrix exactly like :832-834 (230 MB for
    d=7583); the default is the element-wise form, which is bit-identical (every other
    term of each dot product is an exact +0).
    """
    avg_w, avg_w2 = w_avg, w2_avg
    D = pre_D - avg_w[:, None]  # :827  (uses the current w_avg)
    z1 = z1.reshape(1, -1)
    z2 = z2.reshape(-1, 1)
    c1 = scale * (1.0 / np.sqrt(2.0))  # python/numpy double, applied to an fp32 tensor
    if dense_diag:
        sigma = torch.abs(torch.diag(avg_w2 - avg_w**2))  # :832
        w = avg_w[None] + c1 * z1 @ sigma**0.5  # :834
    else:
        sig = torch.abs(avg_w2 - avg_w**2)
        w = avg_w[None] + (c1 * z1) * sig**0.5
    w = w + scale * (D @ z2).T / np.sqrt(2 * (K - 1))  # :835
    return w[0]


# --------------------------------------------------------------------------------------
# Forward (spock_reg_model.py:295-296, :301-321, :416-442, :452-528, :840-908)
# --------------------------------------------------------------------------------------
def soft_clamp(x, lo, high):  # :295-296
    return 0.5 * (torch.tanh(x) + 1) * (high - lo) + lo


def zero_columns(spec: ModelSpec, x):
    """zero_megno / zero_mmr / zero_nan / zero_eplusminus (:452-478): x - mask, mask = x at
    the listed columns.  (NaN stays NaN: NaN - NaN.)"""
    for group in ([7], [3, 6], [38, 39, 40], [1, 2, 4, 5]):
        if all(c in spec.zero_cols for c in group):
            mask = torch.zeros_like(x)
            mask[..., group] = x[..., group].clone()
            x = x - mask
    return x


def _mlp(p, prefix, x, n_layers):
    n_lin = n_layers + 2
    for i in range(n_lin):
        x = torch.nn.functional.linear(x, p[f"{prefix}.{2 * i}.weight"], p[f"{prefix}.{2 * i}.bias"])
        if i < n_lin - 1:
            x = torch.relu(x)  # act defaults to 'relu' (:301-303; VarModel passes no act, :359-360)
    return x


def feature_nn(spec, p, x):  # :359, :417
    return _mlp(p, "feature_nn", x, spec.n_in)


def pooled_moments(f):
    """:418-420  mean over time and unbiased variance (torch.std(...)**2)."""
    sample_mu = torch.mean(f, dim=1)
    sample_var = torch.std(f, dim=1) ** 2
    return sample_mu, sample_var


def compute_summary_stats(spec, p, x, eps1, eps2):
    """:416-435 with the two randn_like draws (:426-427) made explicit."""
    f = feature_nn(spec, p, x)
    sample_mu, sample_var = pooled_moments(f)
    n = f.shape[1]
    std_in_mu = torch.sqrt(sample_var / n)
    std_in_var = torch.sqrt(2 * sample_var**2 / (n - 1))
    mu_sample = eps1 * std_in_mu + sample_mu
    var_sample = eps2 * std_in_var + sample_var
    std_sample = torch.sqrt(torch.abs(var_sample) + EPSILON)
    return torch.cat((mu_sample, std_sample), dim=1)


def predict_instability(spec, p, summary_stats):  # :437-442
    testy = _mlp(p, "regress_nn", summary_stats, spec.n_out)
    mu = soft_clamp(testy[:, [0]], 4.0, 12.0)
    std = soft_clamp(testy[:, [1]], spec.lowest, 6.0)
    return mu, std


def forward_swag_fast(spec, theta, x, eps1, eps2):
    """:878-908 after sample_weights(): mask -> summary stats -> head.  Returns [B,2]."""
    p = unflatten(spec, theta)
    x = zero_columns(spec, x)
    s = compute_summary_stats(spec, p, x, eps1, eps2)
    mu, std = predict_instability(spec, p, s)
    return torch.cat((mu, std), dim=1)


def forward(spec, theta, x, noisy_val=True, eps_in=None, eps1=None, eps2=None, eps_sum=None):
    """VarModel.forward (:486-528).  RNG order when noisy: eps_in[B,T,F] (:445), eps1, eps2
    [B,L] (:426-427), eps_sum[B,2L] (:449).  Returns (out[B,2], summary_kl_terms[B,2L])."""
    p = unflatten(spec, theta)
    x = zero_columns(spec, x)
    if noisy_val:
        x = x + eps_in * torch.exp(p["input_noise_logvar"][None, None, :] / 2)  # :444-446
    s = compute_summary_stats(spec, p, x, eps1, eps2)
    lv = p["summary_noise_logvar"]
    summary_kl = (1 / 2) * (s**2 + torch.exp(lv)[None, :] - lv[None, :] - 1)  # :515-520
    if noisy_val:
        s = s + eps_sum * torch.exp(lv[None, :] / 2)  # :448-450
    mu, std = predict_instability(spec, p, s)
    return torch.cat((mu, std), dim=1), summary_kl


# --------------------------------------------------------------------------------------
# Truncated-normal likelihood (spock_reg_model.py:323-335, :547-593)
# --------------------------------------------------------------------------------------
def safe_log_erf(x):  # :323-335
    base_mask = x < -1
    zero = torch.zeros_like(x)
    x_under = torch.where(base_mask, x, zero)
    x_over = torch.where(~base_mask, x, zero)
    f_under = (
        0.485660082730562 * x_under
        + 0.643278438654541 * torch.exp(x_under)
        + 0.00200084619923262 * x_under**3
        - 0.643250926022749
        - 0.955350621183745 * x_under**2
    )
    f_over = torch.log(1.0 + torch.erf(x_over))
    return f_under + f_over


def lossfnc_per_system(testy, y):
    """_lossfnc (:547-577): [B,2] predictions, [B,2] labels -> [B]."""
    mu = testy[:, [0]]
    std = testy[:, [1]]
    var = std**2
    t_greater_9 = y >= 9
    regression_loss = -((y - mu) ** 2) / (2 * var)
    regression_loss = regression_loss + -torch.log(std)
    regression_loss = regression_loss + -safe_log_erf((mu - 4) / (torch.sqrt(2 * var)))
    classifier_loss = safe_log_erf((mu - 9) / (torch.sqrt(2 * var)))
    safe_regression_loss = torch.where(
        ~torch.isfinite(regression_loss), -torch.ones_like(regression_loss) * 100, regression_loss
    )
    safe_classifier_loss = torch.where(
        ~torch.isfinite(classifier_loss), -torch.ones_like(classifier_loss) * 100, classifier_loss
    )
    total_loss = safe_regression_loss * (~t_greater_9) + safe_classifier_loss * (t_greater_9)
    return -total_loss.sum(1)


def input_kl(p):  # :585-590
    lv = p["input_noise_logvar"]
    return (1 / 2) * (torch.exp(lv) - lv - 1).sum()


def training_loss(spec, theta, x, y, eps_in, eps1, eps2, eps_sum, beta_in=None, beta_out=None):
    """SWAGModel.training_step (:722-732): loss + input_kl*beta_in*B + summary_kl*beta_out.
    Returns (total_loss, dict of the four logged scalars)."""
    beta_in = spec.beta_in if beta_in is None else beta_in
    beta_out = spec.beta_out if beta_out is None else beta_out
    p = unflatten(spec, theta)
    out, skl = forward(spec, theta, x, True, eps_in, eps1, eps2, eps_sum)
    loss = lossfnc_per_system(out, y).sum()  # lossfnc :579-583
    B = x.shape[0]
    ikl = input_kl(p) * beta_in * B
    s_kl = skl.sum() * beta_out
    total = loss + (ikl + s_kl)
    logs = {
        "train_loss_no_reg": loss / B,
        "train_loss_with_reg": total / B,
        "input_kl": ikl / B,
        "summary_kl": s_kl / B,
    }
    return total, logs


def clip_and_sgd_step(theta, grad, buf, lr, momentum, weight_decay, clip, first_step):
    """Lightning 1.1 gradient_clip_val (torch.nn.utils.clip_grad_norm_, L2 over all params,
    coefficient clip/(norm+1e-6), applied only when < 1) followed by torch.optim.SGD
    (:709-711): d = g + wd*theta; buf = d (first step) or momentum*buf + d; theta -= lr*buf."""
    total_norm = torch.linalg.vector_norm(grad, 2)
    clip_coef = clip / (total_norm + 1e-6)
    if clip_coef < 1:
        grad = grad * clip_coef
    d_p = grad + weight_decay * theta if weight_decay != 0 else grad
    buf = d_p.clone() if first_step else momentum * buf + d_p
    theta = theta - lr * buf
    return theta, buf, total_norm


# --------------------------------------------------------------------------------------
# Pre-training schedule (find_minima.py:26-84): CustomOneCycleLR (spock_reg_model.py:27-159) and the KL annealing
# of VarModel.training_step (:595-598)
# --------------------------------------------------------------------------------------
def one_cycle(step_num: int, max_lr: float, total_steps: int, pct_start: float = 0.3, div_factor: float = 25.0,
              final_div_factor: float = 1e4, base_momentum: float = 0.85, max_momentum: float = 0.95):
    """(lr, momentum) the scheduler installs for optimizer step `step_num` (= its last_epoch; 0 at construction):
    cosine from max_lr/25 up to max_lr over step_size_up = pct_start*total - 1 steps, then cosine down to
    max_lr/25/1e4 over the rest, momentum cycling the other way (:64-65, :133-158).  Raises like :137-139 once
    step_num > total_steps -- that ValueError is how find_minima.py's run ends (find_minima.py:79-82)."""
    if step_num > total_steps:
        raise ValueError("Tried to step {} times. The specified number of total steps is {}".format(step_num + 1, total_steps))
    up = float(pct_start * total_steps) - 1
    down = float(total_steps - up) - 1
    initial_lr = max_lr / div_factor
    min_lr = initial_lr / final_div_factor

    def cos_anneal(start, end, pct):  # :119-124
        if pct >= 1.0:
            return end
        return end + (start - end) / 2.0 * (math.cos(math.pi * pct) + 1)

    if step_num <= up:
        return cos_anneal(initial_lr, max_lr, step_num / up), cos_anneal(max_momentum, base_momentum, step_num / up)
    d = step_num - up
    return cos_anneal(max_lr, min_lr, d / down), cos_anneal(base_momentum, max_momentum, d / down)


def kl_annealing(global_step: int, steps: int, beta_in: float, beta_out: float):
    """:596-598: both KL weights ramp linearly over the first 30 % of `steps`."""
    f = min([1, (global_step / steps) / 0.3])
    return f * beta_in, f * beta_out


# --------------------------------------------------------------------------------------
# SWAG moment collection (spock_reg_model.py:763-785)
# --------------------------------------------------------------------------------------
@dataclass
class SwagState:
    K: int
    c: int
    n_models: int = 0
    w_avg: Optional[torch.Tensor] = None
    w2_avg: Optional[torch.Tensor] = None
    pre_D: Optional[torch.Tensor] = None  # [d, <=K], columns oldest -> newest


def aggregate_model(st: SwagState, cur_w: torch.Tensor, current_epoch: int) -> SwagState:
    cur_w2 = cur_w**2
    if st.w_avg is None:
        st.w_avg = cur_w.clone()
        st.w2_avg = cur_w2
    else:
        st.w_avg = (st.w_avg * st.n_models + cur_w) / (st.n_models + 1)
        st.w2_avg = (st.w2_avg * st.n_models + cur_w2) / (st.n_models + 1)
    if st.pre_D is None:
        st.pre_D = cur_w.clone()[:, None]
    elif current_epoch % st.c == 0:
        st.pre_D = torch.cat((st.pre_D, cur_w[:, None]), dim=1)
        if st.pre_D.shape[1] > st.K:
            st.pre_D = st.pre_D[:, 1:]
    st.n_models += 1
    return st


# --------------------------------------------------------------------------------------
# Counter-based RNG used by the batched (production) kernels.  The reference draws its
# normals from torch's global generator (:426-427, :830-831); the batched entry points
# replace that with Philox4x32-10 keyed on global indices so that a sharded run equals a
# single-GPU run bit for bit.  This is the published Random123 Philox4x32-10 algorithm
# (Salmon et al., SC'11); known-answer vectors are checked in tests/test_philox.py.
# --------------------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr: uint32[...,4], key: uint32[...,2] (broadcastable) -> uint32[...,4]."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint32) for i in range(4))
    k0 = np.broadcast_to(key[..., 0], c0.shape).astype(np.uint32)
    k1 = np.broadcast_to(key[..., 1], c0.shape).astype(np.uint32)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PHILOX_M0 * c0.astype(np.uint64)
            p1 = _PHILOX_M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + _PHILOX_W0).astype(np.uint32)
            k1 = (k1 + _PHILOX_W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1)


def _u01(u: np.ndarray) -> np.ndarray:
    """uint32 -> fp32 uniform in (0,1]:  (u + 1) * 2^-32 rounded... computed as
    fp32((u >> 8) + 1) * 2^-24 so that every step is exact in fp32."""
    return ((u >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(2.0**-24)


def box_muller(u: np.ndarray) -> np.ndarray:
    """uint32[...,4] -> fp32 normals[...,4]: (r0 cos, r0 sin, r1 cos, r1 sin)."""
    u = np.asarray(u, dtype=np.uint32)
    a0, b0, a1, b1 = _u01(u[..., 0]), _u01(u[..., 1]), _u01(u[..., 2]), _u01(u[..., 3])
    two_pi = np.float32(6.283185307179586)
    r0 = np.sqrt(np.float32(-2.0) * np.log(a0)).astype(np.float32)
    r1 = np.sqrt(np.float32(-2.0) * np.log(a1)).astype(np.float32)
    t0 = (two_pi * b0).astype(np.float32)
    t1 = (two_pi * b1).astype(np.float32)
    return np.stack(
        [r0 * np.cos(t0), r0 * np.sin(t0), r1 * np.cos(t1), r1 * np.sin(t1)], axis=-1
    ).astype(np.float32)


# Stream ids: which of the path's random tensors a Philox block belongs to.
STREAM_Z1, STREAM_Z2, STREAM_EPS, STREAM_EPS_IN, STREAM_EPS_SUM = 1, 2, 3, 4, 5


def philox_normals(seed: int, stream: int, a: np.ndarray, b: np.ndarray, blk: np.ndarray) -> np.ndarray:
    """Normals for counter (blk, a, b, stream) and key (seed lo, seed hi): fp32[...,4].

    The batched kernels use: z1[m,s,4*blk+i]    = N(seed, Z1,  a=m*S+s... see DESIGN.md.
    """
    a, b, blk = np.broadcast_arrays(np.asarray(a), np.asarray(b), np.asarray(blk))
    ctr = np.stack(
        [blk.astype(np.uint32), a.astype(np.uint32), b.astype(np.uint32), np.full(a.shape, stream, np.uint32)],
        axis=-1,
    )
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return box_muller(philox4x32_10(ctr, key))


def draw_z1(seed: int, unit: np.ndarray, d: int) -> np.ndarray:
    """z1[len(unit), d]; unit = global (model*S + sample) index.  ctr = (j//4, unit, 0, Z1)."""
    nblk = (d + 3) // 4
    z = philox_normals(seed, STREAM_Z1, unit[:, None], 0, np.arange(nblk)[None, :])
    return z.reshape(len(unit), nblk * 4)[:, :d]


def draw_z2(seed: int, unit: np.ndarray, K: int) -> np.ndarray:
    nblk = (K + 3) // 4
    z = philox_normals(seed, STREAM_Z2, unit[:, None], 0, np.arange(nblk)[None, :])
    return z.reshape(len(unit), nblk * 4)[:, :K]


def draw_eps(seed: int, unit: np.ndarray, system: np.ndarray, n: int) -> np.ndarray:
    """eps[len(unit), len(system), n] (n = 2L: eps1 then eps2); ctr = (j//4, unit, system, EPS)."""
    nblk = (n + 3) // 4
    z = philox_normals(
        seed, STREAM_EPS, unit[:, None, None], system[None, :, None], np.arange(nblk)[None, None, :]
    )
    return z.reshape(len(unit), len(system), nblk * 4)[:, :, :n]


# --------------------------------------------------------------------------------------
# Input packing (figures/spock/regression.py:183-213 data_setup_kernel, :144-145 ssX + float)
# --------------------------------------------------------------------------------------
ANGLE_COLS = (11, 12, 13, 17, 18, 19, 23, 24, 25)  # :200


def data_setup_kernel(mass_array: np.ndarray, cur_tseries: np.ndarray) -> np.ndarray:
    """[N,3] masses, [N,T,26] raw series -> [N,T,41] float64 (flags, nan_to_num, cos/sin expansion)."""
    n, t, _ = cur_tseries.shape
    masses = np.broadcast_to(mass_array[:, None, :], (n, t, 3))
    old_X = np.concatenate((cur_tseries, masses), axis=2)
    for c in (3, 6, 7):  # :191-193 flags of the raw columns, before nan_to_num
        old_X = np.concatenate((old_X, (~np.isfinite(old_X[:, :, [c]])).astype(np.float64)), axis=2)
    old_X = np.nan_to_num(old_X, posinf=0.0, neginf=0.0)  # :195
    cols = []
    for j in range(old_X.shape[-1]):
        if j in ANGLE_COLS:
            cols.append(np.cos(old_X[:, :, [j]]))
            cols.append(np.sin(old_X[:, :, [j]]))
        else:
            cols.append(old_X[:, :, [j]])
    X = np.concatenate(cols, axis=2)
    assert X.shape[-1] == 41
    return X


def pack_inputs(mass_array, cur_tseries, ss_mean, ss_scale) -> np.ndarray:
    """data_setup_kernel + StandardScaler.transform (float64) + .float()."""
    X = data_setup_kernel(np.asarray(mass_array, np.float64), np.asarray(cur_tseries, np.float64))
    X = X - np.asarray(ss_mean)  # sklearn: X -= mean_; X /= scale_
    X = X / np.asarray(ss_scale)
    return X.astype(np.float32)


# --------------------------------------------------------------------------------------
# Posterior post-processing (figures/main_figures.py:167-277, figures/multiswag_5_planet.py:306-481)
# --------------------------------------------------------------------------------------
def fast_truncnorm(loc, scale, left=np.inf, right=np.inf, d=10000, nsamp=50, rng=None):
    """main_figures.py:167-223 with numpy's global RNG replaced by `rng` (np.random when None): the first of
    nsamp draws z*scale + loc inside (left, right); the first draw when none is (mask.argmax == 0)."""
    rng = np.random if rng is None else rng
    oldscale = scale
    scale = scale.reshape(-1)
    loc = loc.reshape(-1)
    samples = np.zeros_like(scale)
    for start in range(0, scale.shape[0], d):
        end = min(start + d, scale.shape[0])
        cd = end - start
        rand_out = rng.randn(nsamp, cd) if rng is np.random else rng.standard_normal((nsamp, cd))
        rand_out = rand_out * scale[None, start:end] + loc[None, start:end]
        if right == np.inf:
            mask = rand_out > left
        elif left == np.inf:
            mask = rand_out < right
        else:
            mask = (rand_out > left) & (rand_out < right)
        samples[start:end] = rand_out[mask.argmax(0), np.arange(cd)]
    return samples.reshape(*oldscale.shape)


def _prior_unnormalised(logT):  # main_figures.py:231-234
    return 3.27086190404742 * np.exp(-0.424033970670719 * logT) - 10.8793430454878 * np.exp(-0.200351029031774 * logT**2)


def prior_cdf(t):
    """Closed-form CDF of the prior on [9, inf) (what the CUDA kernel inverts)."""
    from scipy.special import erf

    A, a, B, b = 3.27086190404742, 0.424033970670719, 10.8793430454878, 0.200351029031774
    F = lambda x: A / a * (np.exp(-9 * a) - np.exp(-a * x)) - B * 0.5 * np.sqrt(np.pi / b) * (erf(np.sqrt(b) * x) - erf(9 * np.sqrt(b)))
    return F(np.asarray(t, np.float64)) / F(np.inf)


def prior_samples_table(r, n_samples=None):
    """main_figures.py:236-257: inverse-CDF sampling through a Riemann-sum table with 4*n_samples bins."""
    from scipy.integrate import quad
    from scipy.interpolate import interp1d

    n_samples = len(r) if n_samples is None else n_samples
    normalization = quad(_prior_unnormalised, a=9, b=np.inf)[0]
    bins = n_samples * 4
    top = 100.0
    bin_edges = np.linspace(9, top, num=bins)
    cum_values = [0] + list(np.cumsum(_prior_unnormalised(bin_edges) / normalization * (bin_edges[1] - bin_edges[0]))) + [1]
    bin_edges = [9.0] + list(bin_edges) + [top]
    return interp1d(cum_values, bin_edges)(r)


def posterior_stats(samps_time, pred=None):
    """samps_time [U, N, R] -> per-system dict (multiswag_5_planet.py:421, :476-481); pred [U, N, R, 2] adds the
    'median of dists' (main_figures.py:276-277) of mu* = min over trios of mu and the std of that trio."""
    outs = np.min(samps_time, 2).T  # [N, U]
    res = {
        "average": np.average(outs, 1), "median": np.median(outs, 1),
        "l": np.percentile(outs, 50 + 68 / 2, axis=1), "u": np.percentile(outs, 50 - 68 / 2, axis=1),
        "ll": np.percentile(outs, 50 + 95 / 2, axis=1), "uu": np.percentile(outs, 50 - 95 / 2, axis=1),
    }
    if pred is not None:
        arg = np.argmin(pred[..., 0], 2)  # [U, N]
        mu = np.take_along_axis(pred[..., 0], arg[..., None], 2)[..., 0]
        sd = np.take_along_axis(pred[..., 1], arg[..., None], 2)[..., 0]
        res["median_mu"] = np.median(mu, 0)
        res["median_std"] = np.median(sd, 0)
    return res
