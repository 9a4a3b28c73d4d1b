"""GPU: the data-parallel form of dynamic quantisation.  The reference takes every QuantAct range over the WHOLE batch
(quant_modules.py:149-156); here the batch is sharded over ranks and the fused dynamic plan exchanges each range through
the ranks' slot tables inside its finalize kernels (qasr_dyn_finalize).

On the 8-GPU node every rank is a process with its own GPU, the tables are peer-mapped (CUDA IPC) and each finalize
kernel posts and waits in one launch (bench.py `dynamic_dp` checks that form bit for bit at N = 2 / 4 / 8).  Launches that
spin on each other must never share ONE GPU, so this single-GPU test runs both ranks' plans in one process, in turn on
one stream, with the two phases of the exchange separated: every rank posts, then every rank waits (and finds its
peers' values already there).  Same kernels, same slot arithmetic, same sequence numbers.  Each rank must reproduce the
reference's full-batch logit codes for its rows -- which a shard-local range does not."""
import os

import numpy as np
import pytest
import torch

import specio
from test_gpu_model import build_from_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_in_lockstep(plans, shards):
    for p, (x, lens) in zip(plans, shards):
        p.load_inputs(x, lens)
        for fn in p._lens_fns:
            fn()
    n = len(plans[0].ops)
    assert all(len(p.ops) == n for p in plans)
    for i in range(n):
        if plans[0].ops[i][0] == "finalize":
            for phase in (1, 2):                       # all ranks post, then all ranks wait and finish
                for p in plans:
                    p.finalize_args[i].phase = phase
                    p.ops[i][1]()
            for p in plans:
                p.finalize_args[i].phase = 0
        else:
            for p in plans:
                p.ops[i][1]()
    torch.cuda.synchronize()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_dynamic_plan_equals_the_full_batch_reference(world):
    from qasr import dist as qdist
    from qasr import dynplan
    case = specio.load(os.path.join(GOLDEN, "net_tinyqn_w4a8_dynamic.npz"))
    model, qm = build_from_case(case)
    qm.evaluate(model)
    qm.set_dynamic(model, True)
    B = case["x"].shape[0]
    assert B >= world
    n_slots = 64
    tables = [torch.zeros(n_slots * world, dtype=torch.int64, device=DEV) for _ in range(world)]
    plans, shards, bounds = [], [], []
    for rank in range(world):
        lo, hi = qdist.shard_bounds(B, rank, world)
        x = torch.from_numpy(case["x"][lo:hi]).to(DEV)
        lens = torch.from_numpy(case["lens"][lo:hi]).to(DEV)
        slots = dynplan.PeerSlots(rank, world, n_slots, DEV, local_tables=tables)
        plans.append(dynplan.DynamicPlan(model.encoder, model.decoder, x.shape, DEV, use_graph=False,
                                         dp=dict(rank=rank, world=world, slots=slots)))
        shards.append((x, lens))
        bounds.append((lo, hi))
    for _ in range(2):                                  # twice: sequence numbers and record resets carry over
        run_in_lockstep(plans, shards)
        for p, (lo, hi) in zip(plans, bounds):
            assert not p.overflowed()
            logits = p.logits[:, :, :p.cur_T_out]
            codes = torch.round(logits.double() / p.logits_scale.double()).to(torch.int64).cpu().numpy()
            assert np.array_equal(p.logits_scale.reshape(-1).cpu().numpy(), case["out"]["logit_scale"])
            assert np.array_equal(codes, case["out"]["logit_codes"][lo:hi])
    # the check has teeth: the same shards WITHOUT the exchange (shard-local ranges) give other numbers
    differs = False
    for (x, lens), (lo, hi) in zip(shards, bounds):
        local = dynplan.DynamicPlan(model.encoder, model.decoder, x.shape, DEV, use_graph=False)
        lg = local.run(x, lens)
        codes = torch.round(lg.double() / local.logits_scale.double()).to(torch.int64).cpu().numpy()
        differs |= not np.array_equal(codes, case["out"]["logit_codes"][lo:hi])
    assert differs
