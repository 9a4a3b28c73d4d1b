"""GPU: the data-parallel form of dynamic quantisation.  The reference takes every QuantAct range over the WHOLE batch
(quant_modules.py:149-156); here the batch is sharded over ranks and the fused dynamic plan exchanges each range through
peer-mapped device memory inside its finalize kernels (qasr_dyn_finalize, CUDA IPC).  Two processes share cuda:0 (the
test box has one GPU; on the 8-GPU node the same code runs one rank per GPU, bench.py `dynamic_dp`): each runs its
shard of the golden batch of the UNMODIFIED reference and must reproduce the reference's full-batch logit codes for
its rows bit for bit -- which a shard-local range would not."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, queue):
    try:
        for p in (os.path.join(ROOT, "q-asr_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import torch.distributed as dist
        os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(0)
        import specio
        from qasr import dist as qdist
        from test_gpu_model import build_from_case
        case = specio.load(os.path.join(GOLDEN, "net_tinyqn_w4a8_dynamic.npz"))
        model, qm = build_from_case(case)
        qm.evaluate(model)
        qm.set_dynamic(model, True)
        B = case["x"].shape[0]
        lo, hi = qdist.shard_bounds(B, rank, world)
        x = torch.from_numpy(case["x"][lo:hi]).to("cuda:0")
        lens = torch.from_numpy(case["lens"][lo:hi]).to("cuda:0")
        qdist.enable_data_parallel()
        plan = model._dynamic_plan_or_none(x)
        assert plan is not None and plan.dp is not None and plan.dp["world"] == world
        out = []
        for _ in range(3):                      # eager warm-up + capture + replays: the sequence numbers stay in step
            logits = plan.run(x, lens)
            codes = torch.round(logits.double() / plan.logits_scale.double()).to(torch.int64).cpu().numpy()
            out.append(codes)
        assert not plan.overflowed()
        want = case["out"]["logit_codes"][lo:hi]
        ok = all(np.array_equal(c, want) for c in out)
        scale_ok = np.array_equal(plan.logits_scale.reshape(-1).cpu().numpy(), case["out"]["logit_scale"])
        # the same shard WITHOUT the exchange: a shard-local range (what a naive split computes)
        qdist.disable_data_parallel()
        local = model._dynamic_plan_or_none(x)
        ll = local.run(x, lens)
        lcodes = torch.round(ll.double() / local.logits_scale.double()).to(torch.int64).cpu().numpy()
        differs = not np.array_equal(lcodes, want) or not np.array_equal(
            local.logits_scale.reshape(-1).cpu().numpy(), case["out"]["logit_scale"])
        dist.barrier()
        qdist.close_peer_slots()
        dist.destroy_process_group()
        queue.put((rank, ok, scale_ok, differs, None))
    except Exception:  # noqa: BLE001
        queue.put((rank, False, False, False, traceback.format_exc()))


@pytest.mark.timeout(300)
def test_sharded_dynamic_plan_equals_the_full_batch_reference():
    world = 2
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, queue)) for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=280) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, ok, scale_ok, differs, err in sorted(results):
        assert err is None, "rank %d failed:\n%s" % (rank, err)
        assert ok, "rank %d: sharded logit codes differ from the reference's full-batch codes" % rank
        assert scale_ok, "rank %d: decoder scale differs" % rank
    # the check has teeth: at least one shard's local range gives other numbers than the full-batch range
    assert any(r[3] for r in results)
