"""Sampler update fused with the classifier-free-guidance combine.

Mirrors the scheduler surface the reference's pipeline uses (pipeline_bev_controlnet.py:322-323,387-389,
497-499): `set_timesteps`, `timesteps`, `scale_model_input`, `step(model_output, t, sample).prev_sample`,
`order`.  The arithmetic restates diffusers-0.17.1 `UniPCMultistepScheduler` (bh2, solver_order 2,
predict_x0, lower_order_final; SURVEY.md Appendix A.3 — diffusers is not vendored, so this is "parity
unpinned" against the reference and pinned against the oracle's restatement) and `DDIMScheduler` (eta 0).

Every UniPC/DDIM update is a linear combination of (x, last corrected sample, x0 history, new x0) whose
scalar coefficients depend only on the timestep index, so the host computes them once in fp64 and ONE
kernel (dd_cfg_sched_step) does CFG + x0 conversion + corrector + predictor + history shift per step.
"""
from dataclasses import dataclass

import numpy as np
import torch

COEF_LEN = 16


def sd_schedule(num_train=1000, beta_start=0.00085, beta_end=0.012):
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=np.float64) ** 2
    ac = np.cumprod(1.0 - betas)
    alpha, sigma = np.sqrt(ac), np.sqrt(1.0 - ac)
    return ac, alpha, sigma, np.log(alpha) - np.log(sigma)


def unipc_timesteps(n, num_train=1000):
    ts = np.linspace(0, num_train - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
    _, idx = np.unique(ts, return_index=True)
    return ts[np.sort(idx)]


def unipc_coefficients(timesteps, guidance_scale=1.0, solver_order=2):
    """-> float64 [N, 16]: {g, sigma_t, 1/alpha_t, a_last, a_m0, a_m1, a_x0, a_x, b_xc, b_x0, b_m0, 0...} per step"""
    _, alpha, sigma, lam = sd_schedule()
    N = len(timesteps)
    out = np.zeros((N, COEF_LEN))
    lower_order_nums, this_order = 0, 1

    def rho_corrector(order, hh, r1):
        phi1 = np.expm1(hh)
        B = phi1
        rks = np.array(([r1] if order == 2 else []) + [1.0])
        R, b = [], []
        h_phi_k, fact = phi1 / hh - 1.0, 1.0
        for i in range(1, order + 1):
            R.append(rks ** (i - 1))
            b.append(h_phi_k * fact / B)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1.0 / fact
        if order == 1:
            return np.array([0.5])
        return np.linalg.solve(np.stack(R), np.array(b))

    for i, t in enumerate(timesteps):
        c = out[i]
        c[0], c[1], c[2] = guidance_scale, sigma[t], 1.0 / alpha[t]
        if i == 0:
            c[7] = 1.0  # no corrector on the first step: xc = x
        else:
            s0 = timesteps[i - 1]
            h = lam[t] - lam[s0]
            hh = -h
            phi1 = np.expm1(hh)
            B = phi1
            c[3] = sigma[t] / sigma[s0]
            if this_order == 1:
                c[4] = -alpha[t] * phi1 + 0.5 * alpha[t] * B
                c[6] = -0.5 * alpha[t] * B
            else:
                r1 = (lam[timesteps[i - 2]] - lam[s0]) / h
                rho = rho_corrector(2, hh, r1)
                c[5] = -alpha[t] * B * rho[0] / r1
                c[4] = -alpha[t] * phi1 + alpha[t] * B * rho[0] / r1 + alpha[t] * B * rho[1]
                c[6] = -alpha[t] * B * rho[1]
        order = min(solver_order, N - i)              # lower_order_final
        this_order = min(order, lower_order_nums + 1)
        prev_t = timesteps[i + 1] if i + 1 < N else 0
        h = lam[prev_t] - lam[t]
        phi1 = np.expm1(-h)
        B = phi1
        c[8] = sigma[prev_t] / sigma[t]
        c[9] = -alpha[prev_t] * phi1
        if this_order == 2:
            r1 = (lam[timesteps[i - 1]] - lam[t]) / h
            c[10] = -0.5 * alpha[prev_t] * B / r1
            c[9] += 0.5 * alpha[prev_t] * B / r1
        if lower_order_nums < solver_order:
            lower_order_nums += 1
    return out


def ddim_timesteps(n, num_train=1000, steps_offset=1):
    ratio = num_train // n
    return (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset


def ddim_coefficients(timesteps, guidance_scale=1.0, num_train=1000):
    ac, alpha, sigma, _ = sd_schedule(num_train)
    N = len(timesteps)
    out = np.zeros((N, COEF_LEN))
    for i, t in enumerate(timesteps):
        prev_t = t - num_train // N
        a_prev = ac[prev_t] if prev_t >= 0 else ac[0]  # set_alpha_to_one=False
        c = out[i]
        c[0], c[1], c[2] = guidance_scale, sigma[t], 1.0 / alpha[t]
        c[7] = 1.0                                                     # xc = x
        c[8] = np.sqrt(1 - a_prev) / sigma[t]                          # eps = (x - alpha x0) / sigma
        c[9] = np.sqrt(a_prev) - np.sqrt(1 - a_prev) * alpha[t] / sigma[t]
    return out


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor


class _SchedulerConfig(dict):
    """`scheduler.config` with attribute access (diffusers FrozenDict behaviour)"""
    __getattr__ = dict.__getitem__


# the Stable Diffusion v1.5 noise schedule (scheduler/scheduler_config.json of the checkpoint the reference loads,
# misc/test_utils.py:148-162): the only one the coefficient tables are derived and checked for
_SD_SCHEDULE = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                    trained_betas=None, prediction_type="epsilon")


class _FusedScheduler:
    order = 1
    init_noise_sigma = 1.0
    _extra_config = {}

    def __init__(self, guidance_scale=1.0):
        self.guidance_scale = guidance_scale
        self.timesteps = None
        self._coef = None
        self._state = None

    @property
    def config(self):
        return _SchedulerConfig(_class_name=type(self).__name__, **_SD_SCHEDULE, **self._extra_config)

    @classmethod
    def from_config(cls, config, **kwargs):
        """`UniPCMultistepScheduler.from_config(pipe.scheduler.config)` (misc/test_utils.py:162): accepts the config of
        any scheduler of the SD-v1.5 checkpoint (PNDM / DDIM / UniPC ...) and checks that it describes the schedule the
        fused kernel's tables are built for; keys of other scheduler classes are ignored as diffusers does."""
        cfg = dict(config) if not isinstance(config, dict) else config
        cfg = {**cfg, **kwargs}
        for k, want in _SD_SCHEDULE.items():
            have = cfg.get(k, want)
            same = (abs(have - want) < 1e-12) if isinstance(want, float) else (have == want)
            if not same:
                raise NotImplementedError(f"{cls.__name__}: {k}={have!r} is not the SD-v1.5 schedule ({want!r}) this "
                                          "scheduler's coefficient tables are built for")
        for k, want in cls._extra_config.items():
            if k in cfg and cfg.get("_class_name", cls.__name__) == cls.__name__ and cfg[k] != want:
                raise NotImplementedError(f"{cls.__name__}: {k}={cfg[k]!r} (built: {want!r})")
        return cls()

    def _tables(self, n):
        raise NotImplementedError

    def set_timesteps(self, num_inference_steps, device=None):
        ts, coef = self._tables(num_inference_steps)
        self.num_inference_steps = len(ts)
        self._ts_host = ts
        self.timesteps = torch.from_numpy(ts).to(device) if device is not None else torch.from_numpy(ts)
        self._coef_host = coef
        self._coef = None
        self._state = None
        self._step_index = 0

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coef_table(self, device):
        if self._coef is None or self._coef.device != torch.device(device):
            self._coef = torch.from_numpy(self._coef_host).float().to(device)
        return self._coef

    def _ensure_state(self, sample):
        if self._state is None or self._state[0].shape != sample.shape:
            self._state = [torch.zeros_like(sample, dtype=torch.float32) for _ in range(3)]  # last, m0, m1
        return self._state

    def step(self, model_output, timestep, sample, return_dict=True, cfg=False, eps_nchw=True, **kwargs):
        """diffusers-style call.  model_output: guided noise (n, c, h, w) fp32 — or, with cfg=True, the raw
        (2n, ...) uncond/cond stack, in which case the CFG combine happens in the same kernel."""
        from . import ops
        if not sample.is_cuda:
            raise RuntimeError("dualdiff_b200 has no CPU path: `sample` must be a CUDA tensor")
        t = int(timestep)
        hit = np.nonzero(self._ts_host == t)[0]
        i = int(hit[0]) if len(hit) else len(self._ts_host) - 1
        x = sample.float().contiguous().clone()
        last, m0, m1 = self._ensure_state(x)
        n, c, h, w = x.shape
        eps = model_output.float().contiguous()
        ops.cfg_sched_step(eps, x, last, m0, m1, self.coef_table(x.device)[i], n_img=n, c=c, hw=h * w, cfg=cfg,
                           eps_nchw=eps_nchw)
        x = x.to(sample.dtype)
        return SchedulerOutput(prev_sample=x) if return_dict else (x,)


class UniPCMultistepScheduler(_FusedScheduler):
    _extra_config = dict(solver_order=2, solver_type="bh2", predict_x0=True, lower_order_final=True, thresholding=False)

    def _tables(self, n):
        ts = unipc_timesteps(n)
        return ts, unipc_coefficients(ts, self.guidance_scale)


class DDIMScheduler(_FusedScheduler):
    _extra_config = dict(steps_offset=1, set_alpha_to_one=False, clip_sample=False)

    def _tables(self, n):
        ts = ddim_timesteps(n)
        return ts, ddim_coefficients(ts, self.guidance_scale)
