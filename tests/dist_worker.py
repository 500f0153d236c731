"""Worker of tests/test_dist_cpu.py: one rank of a world_size-2 gloo group exercising the host-side logic of the N > 1 path —
utterance sharding (bench.clip_indices), the library's longest-first scheduler (q3asr_schedule, no GPU needed), max-over-ranks
timing and the host-side result gather in original order.  No collective touches the data path."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "qwen3-asr-swift_b200"))
import bench  # noqa: E402
import q3asr  # noqa: E402


def fake_transcribe(clip_index, n_samples):
    """stand-in for the GPU: ids that depend only on the utterance"""
    return [clip_index, n_samples % 1000, (clip_index * 7919) % 151936]


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # 1. weak-scaling shards are disjoint and cover the job
    mine = bench.clip_indices(rank, per_gpu=8)
    all_idx = [None] * world
    dist.all_gather_object(all_idx, mine)
    flat = sum(all_idx, [])
    assert sorted(flat) == list(range(8 * world)) and len(set(flat)) == len(flat)
    # 2. the scheduler's assignment is a pure function of the lengths: every rank computes the same one
    rng = np.random.default_rng(5)
    lens = rng.integers(16000, 480000, size=37).astype(np.uint64)
    assign = q3asr.schedule(lens, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, assign.tolist())
    assert all(g == gathered[0] for g in gathered)
    load = [int(lens[assign == g].sum()) for g in range(world)]
    assert max(load) - min(load) <= 480000
    # 3. each rank handles its share; rank 0 gathers on the host and restores the original order
    out = {int(i): fake_transcribe(int(i), int(lens[i])) for i in np.nonzero(assign == rank)[0]}
    parts = [None] * world if rank == 0 else None
    dist.gather_object(out, parts, dst=0)
    if rank == 0:
        merged = {}
        for p in parts:
            assert not (merged.keys() & p.keys())
            merged.update(p)
        assert [merged[i] for i in range(len(lens))] == [fake_transcribe(i, int(lens[i])) for i in range(len(lens))]
    # 4. timing is the max over ranks
    t = bench.max_over_ranks_cpu(10.0 + rank, world)
    assert t == 10.0 + world - 1
    dist.barrier()
    if rank == 0:
        print("DIST_OK", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
