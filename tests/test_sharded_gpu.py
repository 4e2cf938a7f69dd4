"""The sharded driver on real GPUs (NCCL, one process per GPU): needs >= 2 CUDA devices, skipped otherwise (the round-end
GPU box has one; run with `gpurun --gpus 2 -- python -m pytest tests/test_sharded_gpu.py -m gpu`).  Same checks as the
gloo test, but the per-rank map is the CUDA library and the result on rank 0 is compared with the CPU oracle."""
import pytest
import torch
import torch.multiprocessing as mp

from tests.test_sharded_gloo import _free_port, _worker

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("typ", [1, 3])
@pytest.mark.parametrize("delivery", [False, True, "owned"])
def test_two_gpu_sharded_equals_oracle(typ, delivery):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, typ, delivery, q, "nccl")) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    root = [o for o in outs if o["rank"] == 0][0]
    other = [o for o in outs if o["rank"] == 1][0]
    assert root["res"] == other["res"] == root["exp"] and root["res"][6] == 1
    assert root["same_grid"] and min(root["counts"]) > 0
    assert root["received"] == root["counts"][1] and root["ntiles"] == sum(root["counts"])
    assert root["bad"] == 0 and root["image_equal"]
