"""Data-parallel helpers: utterances are independent, so the batch dimension shards across ranks
(one process per GPU, torch.distributed: NCCL on GPUs, gloo in CPU tests).

Static evaluation needs no communication.  Calibration / dynamic mode computes each QuantAct's range
over the WHOLE batch in the reference (quant_modules.py:113-119, 149-156); min and max are associative,
so one 2-float all-reduce per QuantAct keeps a sharded run bit-identical to the full-batch run."""
import torch
import torch.distributed as dist

_state = {"enabled": False, "group": None}


def enable_data_parallel(group=None):
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    _state["enabled"], _state["group"] = True, group


def disable_data_parallel():
    _state["enabled"], _state["group"] = False, None


def is_enabled():
    return _state["enabled"]


def allreduce_range(x_min, x_max):
    """Global (min, max) over all ranks with ONE collective: MAX-reduce of [-min, max]."""
    if not _state["enabled"]:
        return x_min, x_max
    buf = torch.cat([-x_min.reshape(-1), x_max.reshape(-1)])
    dist.all_reduce(buf, op=dist.ReduceOp.MAX, group=_state["group"])
    n = x_min.numel()
    return (-buf[:n]).reshape(x_min.shape), buf[n:].reshape(x_max.shape)


def shard_bounds(n_items, rank, world):
    """Contiguous utterance range [start, stop) of `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_predictions(labels):
    """All-gather the per-rank (B_local, T') greedy label tensors along the batch dimension."""
    if not _state["enabled"]:
        return labels
    world = dist.get_world_size(_state["group"])
    sizes = [torch.zeros(1, dtype=torch.long, device=labels.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([labels.shape[0]], dtype=torch.long, device=labels.device),
                    group=_state["group"])
    mx = int(max(int(s) for s in sizes))
    pad = torch.zeros((mx,) + tuple(labels.shape[1:]), dtype=labels.dtype, device=labels.device)
    pad[:labels.shape[0]] = labels
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=_state["group"])
    return torch.cat([o[:int(s)] for o, s in zip(out, sizes)], dim=0)


_slots = {}


def peer_slots(device, n_slots=512):
    """Peer-mapped slot tables of the fused dynamic plan's range exchange (qasr.dynplan.PeerSlots), one set per
    device and process group; the CUDA IPC handles travel through all_gather_object."""
    from . import dynplan
    key = (str(device), id(_state["group"]))
    if key not in _slots:
        group = _state["group"]
        rank, world = dist.get_rank(group), dist.get_world_size(group)

        def gather(handle):
            out = [None] * world
            dist.all_gather_object(out, handle, group=group)
            return out
        _slots[key] = dict(rank=rank, world=world, slots=dynplan.PeerSlots(rank, world, n_slots, device, gather))
    return _slots[key]


def close_peer_slots():
    """Unmap / free every peer-mapped slot table of this process (call before destroy_process_group)."""
    for v in _slots.values():
        v["slots"].close()
    _slots.clear()
