"""Drop-in `nemo.quantization.utils.distill_data` (reference distill_data.py:11-162): zero-shot calibration data.
Random inputs are optimised (Adam + ReduceLROnPlateau) until the output statistics of every conv that feeds a
BatchNorm match that BatchNorm's running statistics (KL divergence between the two Gaussians); the result is the
`List[FloatTensor(B, 64, T)]` that examples/asr/quantization/inference.py loads with --load.

This is the FP32 side of Q-ASR: the teacher runs with quant_mode 'none' (plain torch convolutions, autograd), which
is the API's float pass-through state, not the integer hot path -- no sm_100a kernel is involved and none is claimed.
Same entry points, argument meaning and defaults as the reference; differences: the initial noise comes from a
seedable generator instead of a 32-worker DataLoader (same distribution, U[-0.3, 0.3), distill_data.py:24), and the
model is used on whatever device it lives on instead of a hard `.cuda()`.
"""
import torch
import torch.nn as nn
import torch.optim as optim

__all__ = ["UniformDataset", "OutputHook", "get_synthetic_data"]


class UniformDataset(torch.utils.data.Dataset):
    """Random uniform samples from [-0.3, 0.3) of shape `size` (reference :11-25)."""

    def __init__(self, length, size, transform=None, generator=None):
        self.length, self.size, self.transform, self.generator = length, tuple(size), transform, generator

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        return torch.rand(self.size, generator=self.generator) * 0.6 - 0.3


class OutputHook(object):
    """Forward hook that keeps the output of an intermediate layer (reference :27-39)."""

    def __init__(self):
        self.outputs = None

    def hook(self, module, input, output):
        self.outputs = output

    def clear(self):
        self.outputs = None


def _get_random_data(batch_size=32, dim=64, seqlen=500, generator=None):
    """Iterable of [batch_size, dim, seqlen] noise batches (reference :41-57: a DataLoader over UniformDataset)."""
    dataset = UniformDataset(length=10000, size=(dim, seqlen), generator=generator)
    return torch.utils.data.DataLoader(dataset, batch_size=batch_size, shuffle=False, num_workers=0)


def _kl_loss(bn_mean, bn_std, tmp_mean, tmp_std):
    """KL(N(tmp) || N(bn)) averaged over channels, reference :59-68:
    log(tmp_std / bn_std) - 0.5 * (1 - (bn_std^2 + (bn_mean - tmp_mean)^2) / tmp_std^2)."""
    ratio = (bn_std ** 2 + (bn_mean - tmp_mean) ** 2) / tmp_std ** 2
    return (torch.log(tmp_std / bn_std) - 0.5 * (1 - ratio)).mean()


def bn_statistics_loss(hooks, bn_stats, eps=1e-6):
    """Sum over the hooked convs of the KL loss between their output statistics over (batch, time) and the
    BatchNorm's running (mean, std) (reference :137-148; note the reference's `var(x + eps)` and `sqrt(var + eps)`)."""
    total = 0
    for hook, (bn_mean, bn_std) in zip(hooks, bn_stats):
        out = hook.outputs[0]
        conv_mean = torch.mean(out, dim=(0, 2))
        conv_std = torch.sqrt(torch.var(out + eps, dim=(0, 2)) + eps)
        assert bn_mean.shape == conv_mean.shape and bn_std.shape == conv_std.shape
        total = total + _kl_loss(bn_mean, bn_std, conv_mean, conv_std)
    return total


def get_synthetic_data(teacher_model, teacher_model_decoder, batch_size, dim, seqlen, train_iter=500, num_batch=1,
                       lr=0.01, generator=None, initial_data=None, verbose=True):
    """Synthetic calibration batches from the BatchNorm statistics of the single-precision encoder
    (reference :71-162).  teacher_model: encoder in quant_mode 'none' with its BatchNorms NOT folded
    (`convs_before_bn` lists the (conv, bn) pairs); teacher_model_decoder: the decoder (run for parity with the
    reference's loop, its output does not enter the loss).  Returns a list of `num_batch` tensors
    [batch_size, dim, seqlen].  `initial_data` (iterable of batches) replaces the random start, for reproducible runs."""
    eps = 1e-6
    teacher_model = teacher_model.eval()
    device = next(teacher_model.parameters()).device
    hooks, handles, bn_stats = [], [], []
    for conv, bn in teacher_model.convs_before_bn:
        assert isinstance(bn, nn.BatchNorm1d)
        hook = OutputHook()
        hooks.append(hook)
        handles.append(conv.register_forward_hook(hook.hook))
        bn_stats.append((bn.running_mean.detach().clone().flatten().to(device),
                         torch.sqrt(bn.running_var + eps).detach().clone().flatten().to(device)))
    batches = initial_data if initial_data is not None else _get_random_data(batch_size, dim, seqlen, generator)
    refined = []
    try:
        with torch.enable_grad():
            for i, data in enumerate(batches):
                if i == num_batch:
                    break
                if verbose:
                    print("Distillation: %s / %s" % (i + 1, num_batch))
                data = data.detach().clone().to(device).requires_grad_(True)
                optimizer = optim.Adam([data], lr=lr)
                scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, min_lr=1e-4, patience=25)
                length = torch.tensor([data.shape[2]] * data.shape[0], device=device)
                for _ in range(train_iter):
                    teacher_model.zero_grad()
                    optimizer.zero_grad()
                    for hook in hooks:
                        hook.clear()
                    encoded, encoded_len, encoded_sf = teacher_model(data, length)
                    teacher_model_decoder(encoder_output=encoded, encoder_output_scaling_factor=encoded_sf)
                    total_loss = bn_statistics_loss(hooks, bn_stats, eps)
                    total_loss.backward()
                    optimizer.step()
                    scheduler.step(total_loss.item())
                refined.append(data.detach().clone())
    finally:
        for handle in handles:
            handle.remove()
    return refined
