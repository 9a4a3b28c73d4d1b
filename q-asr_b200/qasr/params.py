"""Healthy random parameters for parity tests and the synthetic benchmark.

The reference's default random init is degenerate for a quantized forward
(activations shrink ~4x per sub-layer until the 1e-8 range clamp of
quant_utils.py:49,52 takes over).  Weights and BN statistics are
INPUTS to both the reference/oracle and this implementation, so they are produced
by one BN re-estimation pass: all blocks in quant_mode 'none',
every BatchNorm1d in train mode with momentum 1.0, one FP32 forward on seeded
N(0,1) features, then back to eval / 'symmetric'.  Duck-typed: works on this
repo's module tree and on the reference's.
"""
import torch
import torch.nn as nn


@torch.no_grad()
def reestimate_bn_(encoder, feat_in=64, seed=1234, batch=4, frames=256, jitter=True):
    """In-place BN re-estimation; call BEFORE encoder.bn_folding()."""
    g = torch.Generator().manual_seed(seed)
    bns = [m for m in encoder.modules() if isinstance(m, nn.BatchNorm1d)]
    if jitter:  # non-trivial affine parameters so gamma/beta folding is exercised
        for bn in bns:
            bn.weight.copy_(0.75 + 0.5 * torch.rand(bn.weight.shape, generator=g))
            bn.bias.copy_(0.2 * torch.randn(bn.bias.shape, generator=g))
    saved = [(bn.momentum, bn.training) for bn in bns]
    encoder.set_quant_mode("none")
    for bn in bns:
        bn.momentum = 1.0
        bn.train()
    dev = next(encoder.parameters()).device
    x = torch.randn(batch, feat_in, frames, generator=g).to(dev)
    lens = torch.full((batch,), frames, dtype=torch.long, device=dev)
    encoder(x, lens)
    # QuantAct tracks ranges even in quant_mode 'none' (quant_modules.py:111-141 run before the
    # early return at :143-144): forget what this FP32 pass recorded.
    for m in encoder.modules():
        if hasattr(m, "x_min") and hasattr(m, "x_max") and hasattr(m, "act_scaling_factor"):
            m.x_min = torch.zeros_like(m.x_min)
            m.x_max = torch.zeros_like(m.x_max)
    for bn, (mom, training) in zip(bns, saved):
        bn.momentum = mom
        bn.train(training)
        bn.eval()
        # Move 1/sqrt(var + eps) into gamma and leave var + eps == 1.0 exactly: BN folding
        # (quant_modules.py:353-355) then takes sqrt(1.0), which every sqrt implementation gets
        # right.  torch's CPU float32 sqrt is NOT correctly rounded (off by 1 ulp for ~0.5 % of
        # inputs) while CUDA's and numpy's are, and parameters must be identical inputs everywhere.
        std = torch.sqrt(bn.running_var.double() + bn.eps)
        bn.weight.copy_((bn.weight.double() / std).float())
        bn.running_var.fill_(1.0)
        bn.running_var.sub_(bn.eps)
        assert bool(((bn.running_var + bn.eps) == 1.0).all())
    encoder.set_quant_mode("symmetric")


def synthetic_features(batch, frames, feat_in=64, seed=0, ragged=False):
    """Seeded N(0,1) mel-like features (the featurizer output is per-feature
    normalised) and lengths; `ragged` draws lengths in [frames/2, frames]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, feat_in, frames, generator=g)
    if ragged:
        lens = torch.randint(frames // 2, frames + 1, (batch,), generator=g)
        lens[0] = frames
    else:
        lens = torch.full((batch,), frames, dtype=torch.long)
    return x, lens
