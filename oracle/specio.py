"""Save / load nested oracle structures (network specs, traces, golden cases) as one
compressed .npz: arrays are stored under flat keys, the nesting as a JSON skeleton.
TEST INFRASTRUCTURE (see qasr_oracle.py header)."""
import json

import numpy as np

from qasr_oracle import ActState


def _encode(obj, arrays):
    if isinstance(obj, ActState):
        return {"__act__": _encode(dict(bit=obj.bit, x_min=float(obj.x_min), x_max=float(obj.x_max),
                                        momentum=obj.momentum, percentile=obj.percentile,
                                        dynamic=bool(obj.dynamic),
                                        running_stat=bool(obj.running_stat)), arrays)}
    if isinstance(obj, np.ndarray):
        key = "a%d" % len(arrays)
        arrays[key] = obj
        return {"__nd__": key}
    if isinstance(obj, dict):
        return {"__dict__": {k: _encode(v, arrays) for k, v in obj.items()}}
    if isinstance(obj, (list, tuple)):
        return {"__list__": [_encode(v, arrays) for v in obj]}
    if isinstance(obj, (np.floating,)):
        return {"__f32__" if obj.dtype == np.float32 else "__f64__": float(obj)}
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.bool_,)):
        return bool(obj)
    assert obj is None or isinstance(obj, (int, float, str, bool)), type(obj)
    return obj


def _decode(node, arrays):
    if isinstance(node, dict):
        if "__nd__" in node:
            return arrays[node["__nd__"]]
        if "__dict__" in node:
            return {k: _decode(v, arrays) for k, v in node["__dict__"].items()}
        if "__list__" in node:
            return [_decode(v, arrays) for v in node["__list__"]]
        if "__act__" in node:
            return ActState(**_decode(node["__act__"], arrays))
        if "__f32__" in node:
            return np.float32(node["__f32__"])
        if "__f64__" in node:
            return np.float64(node["__f64__"])
    return node


def save(path, obj):
    arrays = {}
    skel = _encode(obj, arrays)
    np.savez_compressed(path, __skeleton__=np.frombuffer(json.dumps(skel).encode(), np.uint8),
                        **arrays)


def load(path):
    with np.load(path) as z:
        arrays = {k: z[k] for k in z.files if k != "__skeleton__"}
        skel = json.loads(bytes(z["__skeleton__"]).decode())
    return _decode(skel, arrays)
