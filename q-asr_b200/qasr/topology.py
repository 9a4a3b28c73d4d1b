"""Encoder topologies of the two benchmark networks.

Restates the `jasper:` block lists of the reference's
examples/asr/conf/quartznet_15x5.yaml:49-229 and
examples/asr/conf/jasper_10x5dr.yaml:49-152 (filters / kernel / repeat / stride /
dilation / residual / residual_dense / separable) as plain Python data so that no
yaml / omegaconf dependency is needed on the hot path.
"""


def _blk(filters, kernel, repeat, stride=1, dilation=1, residual=True, separable=False,
         residual_dense=False):
    return dict(filters=filters, kernel=[kernel], repeat=repeat, stride=[stride],
                dilation=[dilation], residual=residual, separable=separable,
                residual_dense=residual_dense, dropout=0.0)


def quartznet15x5():
    b = [_blk(256, 33, 1, stride=2, residual=False, separable=True)]
    for filters, kernel in ((256, 33), (256, 39), (512, 51), (512, 63), (512, 75)):
        b += [_blk(filters, kernel, 5, separable=True) for _ in range(3)]
    b.append(_blk(512, 87, 1, dilation=2, residual=False, separable=True))
    b.append(_blk(1024, 1, 1, residual=False, separable=False))
    return b


def jasper10x5dr():
    b = [_blk(256, 11, 1, stride=2, residual=False)]
    for filters, kernel in ((256, 11), (384, 13), (512, 17), (640, 21), (768, 25)):
        b += [_blk(filters, kernel, 5, residual_dense=True) for _ in range(2)]
    b.append(_blk(896, 29, 1, dilation=2, residual=False))
    b.append(_blk(1024, 1, 1, residual=False))
    return b


def tiny_quartznet(filters=32, kernel=9, repeat=2, n_res_blocks=2, last=48):
    """A QuartzNet-shaped toy (same block kinds: strided separable head, residual
    separable blocks, dilated separable block, dense 1x1 tail) for fast parity tests."""
    b = [_blk(filters, kernel, 1, stride=2, residual=False, separable=True)]
    b += [_blk(filters, kernel + 2 * i, repeat, separable=True) for i in range(n_res_blocks)]
    b.append(_blk(filters, kernel + 4, 1, dilation=2, residual=False, separable=True))
    b.append(_blk(last, 1, 1, residual=False, separable=False))
    return b


def tiny_jasper(filters=32, kernel=5, repeat=2, n_blocks=3, last=48):
    """A Jasper-shaped toy (dense convs, dense residual) for fast parity tests."""
    b = [_blk(filters, kernel, 1, stride=2, residual=False)]
    b += [_blk(filters + 16 * (i // 2), kernel + 2 * (i // 2), repeat, residual_dense=True)
          for i in range(n_blocks)]
    b.append(_blk(filters + 32, kernel + 4, 1, dilation=2, residual=False))
    b.append(_blk(last, 1, 1, residual=False))
    return b


TOPOLOGIES = {
    "quartznet15x5": quartznet15x5,
    "jasper10x5dr": jasper10x5dr,
    "tiny_quartznet": tiny_quartznet,
    "tiny_jasper": tiny_jasper,
}


def get_same_padding(kernel_size, stride, dilation):
    """jasper.py:61-66."""
    if stride > 1 and dilation > 1:
        raise ValueError("Only stride OR dilation may be greater than 1")
    if dilation > 1:
        return (dilation * kernel_size) // 2 - 1
    return kernel_size // 2
