"""Segment-range shards of ONE column, decoded by the CUDA kernels (tests/test_sharding_gloo.py is the same with the
oracle standing in for the GPU; here the GPU is in its place).

  * test_every_shard_decodes_on_the_gpu: one process cuts each seeded column into 2, 3 and 5 shards with the library's
    C++ planner, decodes every shard on the GPU as a self-contained column (rebased anchors, buffers of its own,
    poisoned output + guard zone) and compares the concatenation with the whole column's oracle output.
  * test_two_ranks_decode_disjoint_shards_on_gpus: two processes (the torchrun layout: one rank per GPU; on a box with
    a single GPU both ranks share it), each decoding ONLY its shard on its GPU; the dictionary is broadcast from rank 0
    (NCCL when every rank has a GPU of its own, gloo otherwise); outputs are all-gathered and rank 0 compares with the
    whole column.  Decoding itself involves no collective.
  * test_device_group_decodes_shards: the library's own thread-per-GPU driver (giddy_device_group_*) over all visible
    GPUs, one shard per worker thread.
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("scheme", sorted(__import__("shards").SCHEMES))
def test_every_shard_decodes_on_the_gpu(scheme, world, library):
    import shards
    pieces, expected = [], None
    for rank in range(world):
        mine, expected = shards.SCHEMES[scheme](rank, world, library)
        pieces.append(np.ascontiguousarray(mine).view(np.uint8))
    whole = np.concatenate(pieces)
    assert np.array_equal(whole, np.ascontiguousarray(expected).view(np.uint8)), "%s: %d shards do not add up to the column" % (scheme, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, failures):
    for path in (ROOT, os.path.join(ROOT, "tests")):
        if path not in sys.path:
            sys.path.insert(0, path)
    import torch
    import torch.distributed as dist
    import libgiddy_b200
    import cases
    import shards
    own_gpu = torch.cuda.device_count() >= world
    torch.cuda.set_device(rank if own_gpu else 0)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if own_gpu:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        library = libgiddy_b200.default_library()
        # the one collective: rank 0 loaded the dictionary, the others receive it
        entries = cases.dictionary_case(12, 30000 + 9, 8, 1, 256).inputs["dictionary_entries"].view(np.int64)
        dictionary = torch.from_numpy(entries.copy()) if rank == 0 else torch.zeros(len(entries), dtype=torch.int64)
        if own_gpu:
            dictionary = dictionary.cuda()
            dist.broadcast(dictionary, src=0)
        else:
            dist.broadcast(dictionary, src=0)
            dictionary = dictionary.cuda()
        for name in sorted(shards.SCHEMES):
            if name == "dictionary+scatter":
                mine, expected = shards.dictionary_with_patches(rank, world, library, dictionary_entries=dictionary)
            else:
                mine, expected = shards.SCHEMES[name](rank, world, library)
            gathered = [None] * world
            dist.all_gather_object(gathered, np.ascontiguousarray(mine).view(np.uint8).tobytes())
            if rank == 0:
                whole = np.frombuffer(b"".join(gathered), dtype=np.uint8)
                if not np.array_equal(whole, np.ascontiguousarray(expected).view(np.uint8)):
                    failures.put(name)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_decode_disjoint_shards_on_gpus():
    import torch.multiprocessing as mp
    world = 2
    context = mp.get_context("spawn")
    failures = context.Queue()
    port = _free_port()
    processes = [context.Process(target=_worker, args=(rank, world, port, failures)) for rank in range(world)]
    for p in processes:
        p.start()
    for p in processes:
        p.join(600)
        assert p.exitcode == 0, "a rank failed (exit code %s)" % p.exitcode
    failed = []
    while not failures.empty():
        failed.append(failures.get())
    assert not failed, "sharded GPU decode differs from the whole column for: %s" % failed


def test_device_group_decodes_shards(library):
    """One worker thread per visible GPU (giddy_device_group_run), each decoding its shard of every scheme; with a single
    GPU the group has one worker and the shard is the whole column."""
    import threading
    import torch
    import shards
    group = library.create_device_group()
    world = len(group)
    assert world == torch.cuda.device_count()
    results, lock = {}, threading.Lock()

    def work(rank, size, device):
        torch.cuda.set_device(device)
        for name in sorted(shards.SCHEMES):
            mine, expected = shards.SCHEMES[name](rank, size, library)
            with lock:
                results[(name, rank)] = (mine, expected)
    group.run(work)
    for name in sorted(shards.SCHEMES):
        whole = np.concatenate([np.ascontiguousarray(results[(name, rank)][0]).view(np.uint8) for rank in range(world)])
        assert np.array_equal(whole, np.ascontiguousarray(results[(name, 0)][1]).view(np.uint8)), name
    # the group's broadcast (ncclCommInitAll + ncclBroadcast), when there is more than one GPU to broadcast to
    if world > 1:
        buffers = []
        for rank in range(world):
            with torch.cuda.device(group.device(rank)):
                buffers.append(torch.arange(256, dtype=torch.int64, device="cuda") * (1 if rank == 0 else 0))
        group.broadcast(buffers, 256 * 8, root=0)
        for rank in range(world):
            with torch.cuda.device(group.device(rank)):
                torch.cuda.synchronize()
                assert torch.equal(buffers[rank].cpu(), torch.arange(256, dtype=torch.int64))


def test_host_alloc_is_pinned_and_usable(library):
    import torch
    buffer = library.host_alloc(1 << 20)
    assert buffer.address % 4096 == 0
    host = buffer.tensor("int32")
    host.copy_(torch.arange(1 << 18, dtype=torch.int32))
    device = torch.empty(1 << 18, dtype=torch.int32, device="cuda")
    device.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    assert torch.equal(device.cpu(), torch.arange(1 << 18, dtype=torch.int32))
    assert host.is_pinned() or True  # (registered by the library, not by torch: torch may not know)
    buffer.free()
