"""Container of a serialised plan: nested dicts / lists / scalars as a JSON skeleton, numpy arrays beside it, all in
one .npz (numpy's own zip container; nothing is unpickled on load)."""
import json

import numpy as np


def _enc(obj, arrays):
    if isinstance(obj, np.ndarray):
        key = "a%d" % len(arrays)
        arrays[key] = obj
        return {"$nd": key}
    if isinstance(obj, dict):
        return {"$d": {str(k): _enc(v, arrays) for k, v in obj.items()}}
    if isinstance(obj, (list, tuple)):
        return [_enc(v, arrays) for v in obj]
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    raise TypeError("plan records hold arrays, numbers, strings, lists and dicts only, not %r" % type(obj))


def _dec(node, arrays):
    if isinstance(node, dict):
        if "$nd" in node:
            return arrays[node["$nd"]]
        return {k: _dec(v, arrays) for k, v in node["$d"].items()}
    if isinstance(node, list):
        return [_dec(v, arrays) for v in node]
    return node


def dump(fileobj, obj):
    arrays = {}
    skeleton = _enc(obj, arrays)
    np.savez_compressed(fileobj, __skeleton__=np.frombuffer(json.dumps(skeleton).encode(), np.uint8), **arrays)


def load(fileobj):
    with np.load(fileobj, allow_pickle=False) as z:
        arrays = {k: z[k] for k in z.files if k != "__skeleton__"}
        skeleton = json.loads(bytes(z["__skeleton__"]).decode())
    return _dec(skeleton, arrays)
