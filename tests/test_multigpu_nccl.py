"""The real multi-rank path on hardware: one process per GPU, NCCL for the per-round histogram all-gather and the
barrier, the routing kernel storing into CUDA-IPC mapped receive buffers of the other ranks.  The k-mers every rank
counted into its shard are dumped, gathered and compared with the oracle's count of all ranks' reads as a map
(bench_mgpu.parity_check, the same function `bench.py --gpus N` runs as its untimed preamble).
Skipped below two GPUs."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, wl, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import bench_mgpu
        res = bench_mgpu.parity_check(wl, rank, world, rank, n_reads=6000)
        if rank == 0:
            q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c2", "c2-fakeseq", "c3", "c5"])
def test_sharded_counter_nccl_matches_oracle(name):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 1
    while world * 2 <= min(n, 8):
        world *= 2
    import bench
    wl = dict(bench.WORKLOADS[name], name=name)
    if name == "c3":
        wl["genome"] = 1 << 8                  # a small dictionary: heavy hitters at this sample size
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, wl, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["checked"] and res["ok"], res
    assert res["ranks"] == world and res["rounds"] >= 2
