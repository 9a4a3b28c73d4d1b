"""CPU, gloo, world_size 2: the data-parallel host logic (utterance sharding, range all-reduce,
prediction gather).  The quantised kernels themselves need a GPU; QuantAct's range tracking runs in
quant_mode 'none' on CPU tensors exactly as in the reference (quant_modules.py:111-144)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "q-asr_b200"))
    import torch.distributed as dist
    from qasr import dist as qd
    from nemo.quantization.utils.quant_modules import QuantAct
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(3)
        full = [torch.randn(7, 6, 20, generator=g) * (1 + i) for i in range(3)]   # 3 calibration batches
        qd.enable_data_parallel()
        act = QuantAct(8, quant_mode="none")            # range tracking only (CPU)
        ref = QuantAct(8, quant_mode="none")
        for x in full:
            a, b = qd.shard_bounds(x.shape[0], rank, world)
            act(x[a:b])
        qd.disable_data_parallel()
        for x in full:
            ref(x)
        ok = bool(torch.equal(act.x_min, ref.x_min) and torch.equal(act.x_max, ref.x_max))
        # prediction gather with uneven shards
        qd.enable_data_parallel()
        labels = torch.arange(7 * 5).reshape(7, 5)
        a, b = qd.shard_bounds(7, rank, world)
        gathered = qd.gather_predictions(labels[a:b])
        ok = ok and bool(torch.equal(gathered, labels))
        # percentile statistics are refused loudly under DP
        try:
            p = QuantAct(8, quant_mode="none", percentile=99.9)
            p(full[0])
            ok = False
        except NotImplementedError:
            pass
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "q-asr_b200"))
    from qasr.dist import shard_bounds
    for n, w in ((1024, 8), (7, 2), (5, 8), (64, 4)):
        spans = [shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_range_allreduce_matches_full_batch_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
