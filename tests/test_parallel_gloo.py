"""Host-side logic of the multi-GPU path, exercised on CPU with the gloo backend and world_size 2 (no GPU needed):
parameter/buffer broadcast, the flat gradient bucket, one all-reduce per step, gradient == full-batch gradient."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from samplenet_b200.parallel import FlatBucketDataParallel, shard_batch

    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # replicas start DIFFERENT: the wrapper must make them identical
        model = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
        ddp = FlatBucketDataParallel(model)
        ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
        ref.load_state_dict(model.state_dict())  # after broadcast: rank 0's weights everywhere
        torch.manual_seed(7)
        xg = torch.randn(8, 6); yg = torch.randn(8, 3)
        x, y = shard_batch(xg, rank, world), shard_batch(yg, rank, world)
        for step in range(2):
            ddp.zero_grad()
            loss = ((ddp(x) - y) ** 2).mean()
            loss.backward()
            ddp.sync_gradients(); ddp.wait()
        ref.zero_grad()
        ((ref(xg) - yg) ** 2).mean().backward()
        err = max((p.grad - r.grad).abs().max().item() for p, r in zip(model.parameters(), ref.parameters()))
        views = all(p.grad.data_ptr() >= ddp.flat_grad.data_ptr() for p in model.parameters())
        w0 = [p.detach().clone() for p in model.parameters()]
        gathered = [torch.zeros_like(w0[0]) for _ in range(world)]
        dist.all_gather(gathered, w0[0])
        same = all(torch.equal(g, gathered[0]) for g in gathered)
        bad = False
        try:
            for p in model.parameters():
                p.grad = None
            ddp.sync_gradients()
        except RuntimeError:
            bad = True
        q.put((rank, err, views, same, bad, ddp.bucket_bytes()))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_data_parallel_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, views, same, bad, nbytes in res:
        assert err < 1e-6, err          # averaged shard gradients == full-batch gradient
        assert views and same and bad
        assert nbytes == (6 * 16 + 16 + 16 * 3 + 3) * 4


def test_shard_batch_rejects_ragged():
    from samplenet_b200.parallel import shard_batch

    with pytest.raises(ValueError):
        shard_batch(torch.zeros(5, 3), 0, 2)
    assert shard_batch(torch.arange(8).view(8, 1), 1, 4).flatten().tolist() == [2, 3]
