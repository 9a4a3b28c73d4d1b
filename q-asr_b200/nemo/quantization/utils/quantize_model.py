"""Drop-in `nemo.quantization.utils.quantize_model` (reference quantize_model.py:5-75):
recursive mode switches over a model tree.  Traversal semantics are the reference's: an exact
type() match on the three quant modules, nn.Sequential / nn.ModuleList children, otherwise every
nn.Module attribute reachable through dir(model)."""
import torch.nn as nn

from nemo.quantization.utils.quant_modules import *  # noqa: F401,F403
from nemo.quantization.utils.quant_modules import QuantAct, QuantBnConv1d, QuantConv1d, QuantLinear

# the reference's list (quantize_model.py:3) plus QuantBnConv1d, the north-star name for QuantConv1d with its
# BatchNorm folded at set_param time: the traversal matches on exact type(), so the subclass has to be listed
# to be fixed / unfixed with the rest
list_all = [QuantAct, QuantLinear, QuantConv1d, QuantBnConv1d]


def _visit(model, leaf_fn):
    if type(model) in list_all:
        leaf_fn(model)
    elif type(model) == nn.Sequential:
        for _, m in model.named_children():
            _visit(m, leaf_fn)
    elif type(model) == nn.ModuleList:
        for m in model:
            _visit(m, leaf_fn)
    else:
        for attr in dir(model):
            mod = getattr(model, attr)
            if isinstance(mod, nn.Module):
                _visit(mod, leaf_fn)


def set_percentile(model, percentile: float):
    """Recursively set the calibration percentile of every QuantAct."""
    _visit(model, lambda m: m.set_percentile(percentile) if type(m) == QuantAct else None)


def set_dynamic(model, dynamic: bool):
    """Recursively set the dynamic-quantisation flag of every QuantAct."""
    def leaf(m):
        if type(m) == QuantAct:
            m.dynamic = dynamic
    _visit(model, leaf)


def freeze_model(model, freeze_list):
    """fix() the module types in freeze_list, unfix() the other quant modules."""
    _visit(model, lambda m: m.fix() if type(m) in freeze_list else m.unfix())


def evaluate(model):
    """Evaluation mode - fix all operations."""
    freeze_model(model, list_all)


def train(model):
    """Train mode - unfix all operations."""
    freeze_model(model, [])


def calibrate(model):
    """Calibration mode - only QuantAct is unfixed (range tracking on)."""
    freeze_model(model, [QuantConv1d, QuantBnConv1d, QuantLinear])
