"""Build the oracle's plain-numpy network spec from a torch module tree.

TEST INFRASTRUCTURE (see qasr_oracle.py header).  Duck-typed: works on the
reference's JasperBlock / MaskedConv1d / QuantAct / QuantConv1d tree (in the build
container) and on this repo's mirror of that API, because both expose the same
attribute names (jasper.py:293-692, quant_modules.py:18-367 of the reference).
The model must already have had `bn_folding()` applied.
"""
import numpy as np

from qasr_oracle import ActState


def _np(t):
    return None if t is None else t.detach().cpu().numpy().astype(np.float32).copy()


def act_state(act):
    assert not act.per_channel, "oracle models per_channel=False QuantAct only"
    return ActState(bit=act.activation_bit, x_min=float(act.x_min.reshape(-1)[0]),
                    x_max=float(act.x_max.reshape(-1)[0]), momentum=act.act_range_momentum,
                    percentile=act.percentile, dynamic=act.dynamic,
                    running_stat=act.running_stat)


def conv_rec(masked_conv, relu, name):
    c = masked_conv.conv
    bn = None
    if getattr(c, "bn", None) is not None:
        bn = dict(weight=_np(c.bn.weight), bias=_np(c.bn.bias), mean=_np(c.bn.running_mean),
                  var=_np(c.bn.running_var), eps=float(c.bn.eps))
        assert c.fix_bn, "oracle restates the folded+fixed BN branch only"
    return dict(name=name, weight=_np(c.weight), bias=_np(c.bias), bn=bn,
                stride=int(c.stride[0]), pad=int(c.padding[0]), dil=int(c.dilation[0]),
                groups=int(c.groups), weight_bit=int(c.weight_bit), relu=relu,
                act=act_state(masked_conv.act))


def _is_masked_conv(m):
    return hasattr(m, "conv") and hasattr(m, "act") and hasattr(m, "use_mask")


def block_rec(block, name):
    import torch.nn as nn
    convs = []
    mods = list(block.mconv)
    for i, m in enumerate(mods):
        if _is_masked_conv(m):
            relu = i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU)
            convs.append(conv_rec(m, relu, "%s.mconv.%d" % (name, i)))
        else:
            assert isinstance(m, (nn.ReLU, nn.Dropout)), "unfolded / unsupported layer %r" % m
    res = None
    if block.res is not None:
        res = []
        for i, layer in enumerate(block.res):
            assert len(layer) == 1 and _is_masked_conv(layer[0])
            res.append(conv_rec(layer[0], False, "%s.res.%d.0" % (name, i)))
    mout = list(block.mout)
    assert isinstance(mout[0], nn.ReLU)
    return dict(name=name, convs=convs, res=res, res_act=act_state(block.res_act),
                dense_residual=bool(block.dense_residual) if block.res is not None else False)


def spec_from_model(encoder, decoder):
    blocks = [block_rec(b, "encoder.%d" % i) for i, b in enumerate(encoder.encoder_layers)]
    dconv = decoder.decoder_layers[0]
    dec = dict(act=act_state(decoder.act),
               conv=dict(name="decoder", weight=_np(dconv.weight), bias=_np(dconv.bias), bn=None,
                         stride=1, pad=0, dil=1, groups=1, weight_bit=int(dconv.weight_bit),
                         relu=False))
    return dict(blocks=blocks, decoder=dec)


def write_back_act_ranges(spec, encoder, decoder):
    """Copy calibrated (x_min, x_max) from the oracle spec back into a module tree."""
    import torch
    from qasr_oracle import iter_acts

    def mod_acts():
        for b in encoder.encoder_layers:
            for m in b.mconv:
                if _is_masked_conv(m):
                    yield m.act
            if b.res is not None:
                for layer in b.res:
                    yield layer[0].act
            yield b.res_act
        yield decoder.act

    for a, m in zip(iter_acts(spec), mod_acts()):
        m.x_min = torch.full_like(m.x_min, float(a.x_min))
        m.x_max = torch.full_like(m.x_max, float(a.x_max))
