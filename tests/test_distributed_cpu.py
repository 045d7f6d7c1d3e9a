"""N > 1 host logic on CPU: world_size-2 gloo — contiguous batch sharding + all-gather of the
fixed-size detection block reproduces the un-sharded result."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssds_pytorch_b200.ssds import gather_detections, shard_batch
    full = torch.arange(6 * 5 * 6, dtype=torch.float32).reshape(6, 5, 6)     # [B=6, D=5, 6]
    lo, hi = shard_batch(6, rank, world)
    out = gather_detections(full[lo:hi].clone())
    # uneven shards (global batch 5 over 2 ranks: 3 + 2 rows): padded to ceil(5/2) for the collective, trimmed after
    odd = full[:5]
    lo5, hi5 = shard_batch(5, rank, world)
    out5 = gather_detections(odd[lo5:hi5].clone(), n_items=5)
    q.put((rank, (lo, hi), torch.equal(out, full) and torch.equal(out5, odd)))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == (0, 3) and res[1][1] == (3, 6)
    assert all(r[2] for r in res)


def test_shard_batch_ragged():
    from ssds_pytorch_b200.ssds import shard_batch
    assert [shard_batch(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_batch(2, 3, 4) == (2, 2)


def test_plan_branches_bookkeeping():
    """model._Steps: every recorded launch carries the branch it was recorded on; a branch remembers its parent and
    where it forked (what CUDA-graph capture turns into stream fork / join; eager replay ignores it)."""
    from ssds_pytorch_b200.model import _Steps
    st = _Steps()
    st.append("a")
    b1 = st.fork()
    with st.on(b1):
        st.append("b")
        b2 = st.fork(after=b1)
        with st.on(b2):
            st.append("c")
        st.append("d")
    st.append("e")
    assert list(st) == ["a", "b", "c", "d", "e"]
    assert st.tags == [0, b1, b2, b1, 0] and st.cur == 0
    assert st.parents[b1] == (0, 1) and st.parents[b2] == (b1, 2)
