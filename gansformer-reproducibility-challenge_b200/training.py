"""Minimal G/D training step around the attention hot path (BASELINE configs[3]).

What the reference does (expected upstream ``src/training/training_loop.py``, ``src/training/loss.py``,
``src/training/network.py: D_Stylegan2``; none of them is in the checkout -- /root/reference/.SUBMODULES.json:2):
non-saturating logistic losses, R1 gradient penalty on the reals with lazy regularisation, Adam(beta1 = 0,
beta2 = 0.99), an exponential moving average of the generator weights, data parallelism over GPUs with a summed
gradient all-reduce.  Here: one process per GPU, gradients averaged through ONE flat fp32 buffer per network
(``dist.allreduce_gradients``: NCCL over NVLink/NVSwitch, gloo in the CPU tests) -- the only collective of the system.

The attention layers run their CUDA forward; their backward is the composite of ``autograd.py`` (a hand-written
backward kernel is the follow-up).  The discriminator is plain PyTorch plumbing (cuDNN convolutions): it is not on
the hot path.  Path-length regularisation and augmentation are out of scope.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dist as gdist
from .networks import FullyConnected, nf
from ._state import bump_weights_epoch
from .ops import fir4, fir_filter, upfirdn2d_ref

SQRT2 = math.sqrt(2.0)


class EqConv2d(nn.Module):
    """Equalised-LR convolution (+ optional FIR-blurred stride-2 downsampling, + bias + leaky-ReLU * sqrt 2)."""

    def __init__(self, in_ch: int, out_ch: int, kernel: int, down: bool = False, bias: bool = True, act: str = "lrelu"):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_ch, in_ch, kernel, kernel))
        self.bias = nn.Parameter(torch.zeros(out_ch)) if bias else None
        self.wgain = 1.0 / math.sqrt(in_ch * kernel * kernel)
        self.down, self.act, self.kernel = down, act, kernel
        self.register_buffer("fir", fir_filter(), persistent=False)

    def forward(self, x):
        w = self.weight * self.wgain
        if self.down:
            p = (self.fir.shape[0] - 2) + (self.kernel - 1)          # upfirdn padding of StyleGAN2's conv_downsample_2d
            if p % 2 == 0:
                x = fir4(x, self.fir, p // 2)                          # native FIR (forward, backward, double backward for R1)
            else:
                x = upfirdn2d_ref(x, self.fir.to(x.dtype), pad=((p + 1) // 2, p // 2, (p + 1) // 2, p // 2))
            x = F.conv2d(x, w, stride=2)
        else:
            x = F.conv2d(x, w, padding=self.kernel // 2)
        if self.bias is not None:
            x = x + self.bias[None, :, None, None]
        return F.leaky_relu(x, 0.2) * SQRT2 if self.act == "lrelu" else x


class DiscriminatorBlock(nn.Module):
    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        self.conv0 = EqConv2d(in_ch, in_ch, 3)
        self.conv1 = EqConv2d(in_ch, out_ch, 3, down=True)
        self.skip = EqConv2d(in_ch, out_ch, 1, down=True, bias=False, act="linear")

    def forward(self, x):
        return (self.skip(x) + self.conv1(self.conv0(x))) * (1.0 / SQRT2)


class Discriminator(nn.Module):
    """StyleGAN2 residual discriminator (config f channel schedule), images [B,3,R,R] -> logits [B]."""

    def __init__(self, resolution: int = 256, fmap_base: int = 16384, fmap_max: int = 512, mbstd_group: int = 4):
        super().__init__()
        self.resolution, self.mbstd_group = resolution, mbstd_group
        log2 = int(math.log2(resolution))
        self.fromrgb = EqConv2d(3, nf(resolution, fmap_base, fmap_max), 1)
        self.blocks = nn.ModuleList([DiscriminatorBlock(nf(2 ** i, fmap_base, fmap_max), nf(2 ** (i - 1), fmap_base, fmap_max))
                                     for i in range(log2, 2, -1)])
        c4 = nf(4, fmap_base, fmap_max)
        self.conv4 = EqConv2d(c4 + 1, c4, 3)
        self.fc0 = FullyConnected(c4 * 16, c4, act="lrelu")
        self.fc1 = FullyConnected(c4, 1)

    def forward(self, img):
        x = self.fromrgb(img.contiguous(memory_format=torch.channels_last))
        for blk in self.blocks:
            x = blk(x)
        B, C, H, W = x.shape                                            # minibatch standard deviation, one feature map
        G = min(self.mbstd_group, B)
        while B % G:
            G -= 1
        y = x.reshape(G, B // G, C, H, W)
        y = (y - y.mean(dim=0, keepdim=True)).square().mean(dim=0).add(1e-8).sqrt().mean(dim=[1, 2, 3])
        y = y.reshape(1, B // G, 1, 1).expand(G, -1, H, W).reshape(B, 1, H, W)
        x = self.conv4(torch.cat([x, y], dim=1))
        return self.fc1(self.fc0(x.reshape(B, -1))).reshape(B)


@dataclass
class TrainConfig:
    lr: float = 0.002
    r1_gamma: float = 10.0
    d_reg_interval: int = 16            # lazy R1: every 16th discriminator step
    ema_kimg: float = 10.0
    noise_mode: str = "random"
    bucket_mb: float = 32.0             # gradient all-reduce bucket size (MB of fp32 gradients)
    w_avg_beta: float = 0.995           # decay of the running mean of the mapping outputs (truncation trick)


@dataclass
class StepStats:
    loss_g: float = 0.0
    loss_d: float = 0.0
    r1: float = 0.0
    allreduce_bytes: float = 0.0
    allreduce_ms: float = 0.0
    extra: Dict[str, float] = field(default_factory=dict)


class Trainer:
    """One process per GPU.  ``step(z, reals)`` = one discriminator update + one generator update on this rank's shard."""

    def __init__(self, G: nn.Module, D: nn.Module, cfg: Optional[TrainConfig] = None, world: int = 1):
        self.G, self.D, self.cfg, self.world = G, D, cfg or TrainConfig(), world
        self.G_ema = copy.deepcopy(G).eval().requires_grad_(False)
        c = self.cfg.d_reg_interval / (self.cfg.d_reg_interval + 1.0)  # lazy regularisation: rescale lr and betas
        cap = next(G.parameters()).is_cuda                              # capturable: the step can be replayed from a CUDA graph
        self.opt_g = torch.optim.Adam(G.parameters(), lr=self.cfg.lr, betas=(0.0, 0.99), eps=1e-8, capturable=cap)
        self.opt_d = torch.optim.Adam(D.parameters(), lr=self.cfg.lr * c, betas=(0.0 ** c, 0.99 ** c), eps=1e-8, capturable=cap)
        self.it = 0
        # data parallel: gradients live in one flat buffer per network, reduced bucket by bucket while backward still runs
        self.buckets_g = gdist.GradBuckets(G.parameters(), world, self.cfg.bucket_mb) if world > 1 else None
        self.buckets_d = gdist.GradBuckets(D.parameters(), world, self.cfg.bucket_mb) if world > 1 else None

    def _zero(self, opt, buckets):
        if buckets is not None:
            buckets.begin()                # one memset of the flat buffer; the .grad views stay attached
        else:
            opt.zero_grad(set_to_none=True)

    def _allreduce(self, buckets, stats: StepStats):
        if buckets is not None:
            stats.allreduce_bytes += buckets.finish()      # joins the communication stream (the buckets overlapped backward)

    def _step_tensors(self, z: torch.Tensor, reals: torch.Tensor, do_r1: bool, stats: Optional[StepStats] = None):
        """One D update + one G update; returns (loss_d, loss_g, r1) as device tensors without synchronising (capturable)."""
        G, D, cfg = self.G, self.D, self.cfg
        stats = stats if stats is not None else StepStats()
        # ---- discriminator: logistic loss (+ lazy R1 on the reals)
        G.requires_grad_(False); D.requires_grad_(True)
        self._zero(self.opt_d, self.buckets_d)
        with torch.no_grad():
            fakes = G(z, noise_mode=cfg.noise_mode)
        reals_in = reals.detach().requires_grad_(do_r1)
        logit_real, logit_fake = D(reals_in), D(fakes)
        loss_d = F.softplus(logit_fake).mean() + F.softplus(-logit_real).mean()
        r1 = torch.zeros((), device=z.device)
        if do_r1:
            (grad,) = torch.autograd.grad(logit_real.sum(), reals_in, create_graph=True)
            r1 = grad.square().sum(dim=[1, 2, 3]).mean()
            loss_d = loss_d + r1 * (cfg.r1_gamma * 0.5 * cfg.d_reg_interval)
        loss_d.backward()
        self._allreduce(self.buckets_d, stats)
        self.opt_d.step()
        # ---- generator: non-saturating logistic loss
        if z.is_cuda:
            from .attention import advance_dropout
            advance_dropout(z.device)                 # attention dropout: fresh masks for the G phase (device-side, capturable)
        G.requires_grad_(True); D.requires_grad_(False)
        self._zero(self.opt_g, self.buckets_g)
        loss_g = F.softplus(-D(G(z, noise_mode=cfg.noise_mode))).mean()
        loss_g.backward()
        self._allreduce(self.buckets_g, stats)
        self.opt_g.step()
        if z.is_cuda:
            advance_dropout(z.device)                 # ... and for the next step
        # ---- moving average of the generator (and of the mapping outputs: the truncation trick's w_avg)
        with torch.no_grad():
            if hasattr(G, "mapping") and hasattr(G.mapping, "w_avg"):
                ws = G.mapping(z)
                k_ = G.mapping.components_num
                cur = torch.stack([ws[:, :k_].mean(dim=(0, 1)), ws[:, k_:].mean(dim=(0, 1))])
                G.mapping.w_avg.lerp_(cur, 1.0 - cfg.w_avg_beta)
            beta = 0.5 ** (z.shape[0] * self.world / (cfg.ema_kimg * 1000.0))
            for pe, p in zip(self.G_ema.parameters(), G.parameters()):
                pe.lerp_(p.detach(), 1.0 - beta)
            for be, b in zip(self.G_ema.buffers(), G.buffers()):
                be.copy_(b)
        return loss_d.detach(), loss_g.detach(), r1.detach()

    def step(self, z: torch.Tensor, reals: torch.Tensor) -> StepStats:
        stats = StepStats()
        do_r1 = self.cfg.r1_gamma > 0 and self.it % self.cfg.d_reg_interval == 0
        loss_d, loss_g, r1 = self._step_tensors(z, reals, do_r1, stats)
        stats.loss_d, stats.loss_g, stats.r1 = float(loss_d), float(loss_g), float(r1)
        bump_weights_epoch()
        self.it += 1
        return stats

    def step_graphed(self, z: torch.Tensor, reals: torch.Tensor) -> StepStats:
        """The same step replayed from a CUDA graph (one graph with the lazy R1 term, one without): the eager step is bound by
        the host launching ~5000 small kernels.  Shapes are fixed by the first call; the first calls warm up eagerly."""
        from . import attention as _att, networks as _nets
        cfg = self.cfg
        do_r1 = cfg.r1_gamma > 0 and self.it % cfg.d_reg_interval == 0
        st = self.__dict__.setdefault("_graphs", {})
        _nets.CACHE_BYPASS = _att.FORCE_REFOLD = True                  # weight-derived tensors are recomputed inside the graph
        try:
            return self._step_graphed(z, reals, do_r1, st)
        finally:
            _nets.CACHE_BYPASS = _att.FORCE_REFOLD = False

    def _step_graphed(self, z, reals, do_r1, st) -> StepStats:
        cfg = self.cfg
        if "z" not in st:
            st["z"], st["reals"] = torch.empty_like(z), torch.empty_like(reals)
            st["z"].copy_(z); st["reals"].copy_(reals)
            side = torch.cuda.Stream(device=z.device)                 # warm-up off the capture stream: cuDNN autotune, workspaces
            side.wait_stream(torch.cuda.current_stream(z.device))
            with torch.cuda.stream(side):
                for r in ([True, False] if cfg.r1_gamma > 0 else [False]):
                    self._step_tensors(st["z"], st["reals"], r)
            torch.cuda.current_stream(z.device).wait_stream(side)
            torch.cuda.synchronize(z.device)
        st["z"].copy_(z); st["reals"].copy_(reals)
        if do_r1 not in st:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = self._step_tensors(st["z"], st["reals"], do_r1)
            st[do_r1] = (graph, outs)
            # (capture does not execute: fall through to the replay below)
        graph, (loss_d, loss_g, r1) = st[do_r1]
        graph.replay()
        bump_weights_epoch()          # the replay moved G / D / G_ema weights without touching their version counters
        self.it += 1
        return StepStats(loss_g=float(loss_g), loss_d=float(loss_d), r1=float(r1))
