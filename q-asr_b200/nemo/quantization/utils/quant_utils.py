"""Quantisation math of the drop-in `nemo.quantization.utils.quant_utils` API.

Same names, argument meaning and results as the reference's
nemo/quantization/utils/quant_utils.py; the element-wise heavy lifting on activations
runs in the sm_100a kernels behind include/qasr.h (fixedpoint_mul -> qasr_requant_f32),
small per-channel parameter math stays as torch tensor ops on whatever device the
tensors live on (it is done once per weight version, not per element of the hot path).
"""
import torch
from torch.autograd import Function

from qasr import ops as _ops

__all__ = ["linear_quantize", "symmetric_linear_quantization_params", "SymmetricQuantFunction",
           "floor_ste", "round_ste", "batch_frexp", "fixedpoint_mul"]


def linear_quantize(input, scale, zero_point, inplace=False):
    """round(1/scale * input + zero_point)  (reference quant_utils.py:12-26)."""
    inv = 1.0 / scale
    if inplace:
        return input.mul_(inv).add_(zero_point).round_()
    return torch.round(inv * input + zero_point)


def symmetric_linear_quantization_params(num_bits, saturation_min, saturation_max, per_channel=False):
    """scale = clamp(max(|min|, |max|), 1e-8) / (2^(b-1) - 1)  (reference :28-54)."""
    with torch.no_grad():
        n = 2 ** (num_bits - 1) - 1
        bound = torch.clamp(torch.maximum(saturation_min.abs(), saturation_max.abs()), min=1e-8)
        # tensor / tensor: IEEE division on every device.  (tensor / python_scalar is rewritten by
        # torch's CUDA kernel into a multiplication by the reciprocal, one ulp off the CPU result the
        # oracle is pinned to.)
        return bound / torch.full((), float(n), dtype=bound.dtype, device=bound.device)


class SymmetricQuantFunction(Function):
    """clamp(round(x / scale), -n, n - 1) with n = 2^(k-1) - 1  (reference :57-96).
    Used for weights, the 32-bit bias and (through QuantAct) the first-layer input."""

    @staticmethod
    def forward(ctx, x, k, specified_scale=None):
        if specified_scale is None:
            raise ValueError("SymmetricQuantFunction needs a pre-calculated scale")
        n = 2 ** (k - 1) - 1
        ctx.scale = specified_scale
        q = linear_quantize(x, specified_scale, torch.zeros((), device=x.device), inplace=False)
        return torch.clamp(q, -n, n - 1)

    @staticmethod
    def backward(ctx, grad_output):
        scale = ctx.scale
        if grad_output.dim() == 4:
            scale = scale.view(-1, 1, 1, 1)
        elif grad_output.dim() == 2:
            scale = scale.view(-1, 1)
        else:
            scale = scale.view(-1)
        return grad_output.clone() / scale, None, None


class floor_ste(Function):
    @staticmethod
    def forward(ctx, x):
        return torch.floor(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone()


class round_ste(Function):
    @staticmethod
    def forward(ctx, x):
        return torch.round(x)

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone()


def batch_frexp(inputs, max_bit=31):
    """Scale ratio -> (mantissa * 2^31 rounded half up, 31 - exponent)  (reference :121-147).
    Stays on the input's device: no .cpu()/Decimal round trip."""
    shape = inputs.size()
    mant, ex = torch.frexp(inputs.reshape(-1).double())
    m = torch.floor(mant * float(2 ** max_bit) + 0.5).to(torch.int64)
    e = float(max_bit) - ex.to(torch.float64)
    return m.view(shape), e.view(shape)


class fixedpoint_mul(Function):
    """Dyadic requantisation  clamp(round(round(x/s_in) * m / 2^e) [+ identity term])
    (reference :149-224), executed by the qasr_requant_f32 kernel."""

    @staticmethod
    def forward(ctx, pre_act, pre_act_scaling_factor, bit_num, quant_mode, z_scaling_factor,
                identity=None, identity_scaling_factor=None):
        ctx.identity = identity
        ctx.z_scaling_factor = z_scaling_factor
        if quant_mode == "symmetric":
            n = 2 ** (bit_num - 1) - 1
            qmin, qmax = -n - 1, n
        else:
            qmin, qmax = 0, 2 ** bit_num - 1
        x3 = pre_act if pre_act.dim() == 3 else pre_act.reshape(1, -1, pre_act.shape[-1])
        id3 = None
        if identity is not None:
            id3 = identity if identity.dim() == 3 else identity.reshape(x3.shape)
        z_sf = z_scaling_factor.reshape(-1)[:1].float()
        _, xq = _ops.requant_f32(x3, pre_act_scaling_factor, z_sf, qmin, qmax, identity=id3,
                                 id_sf=identity_scaling_factor, want_codes=False)
        # the kernel returns codes * z_sf (what QuantAct hands on); recover the integer codes
        return torch.round(xq / z_sf).view(pre_act.shape)

    @staticmethod
    def backward(ctx, grad_output):
        identity_grad = None
        if ctx.identity is not None:
            identity_grad = grad_output.clone() / ctx.z_scaling_factor
        return grad_output.clone() / ctx.z_scaling_factor, None, None, None, None, identity_grad, None
