"""Host-side logic of the episode runner (no GPU): the lock-step schedule over the reference's own episode order and reset
flags (tests/golden/loader_order.json, produced by executing SMNet/loader.py:97-117,289-293), stream splitting, slot
refill in waves, rank partition - and the N>1 path over gloo with world_size 2."""
import json
import multiprocessing as mp
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "loader_order.json")))


def _lengths(order):
    # ragged but deterministic: 1..20 frames (loader.py:71 caps sequences at 20), a few empty ones
    return [(sum(map(ord, f)) * 7) % 21 for f in order]


@pytest.mark.parametrize("test_type", ["default", "episodic", "longterm"])
@pytest.mark.parametrize("n_slots", [1, 3, 64])
def test_schedule_replays_reference_order_per_stream(eod, test_type, n_slots):
    g = _golden()
    order = eod.formats.order_files(g["files"], test_type)
    assert order == g[f"order_{test_type}"]
    n_frames = _lengths(order)
    reset_first = [eod.formats.memory_reset_flag(test_type, f, 0) for f in order]
    sched = eod.LockStepSchedule(n_frames, reset_first, n_slots)
    streams = eod.runner.split_streams(reset_first)
    if test_type == "episodic":
        assert all(len(s) == 1 for s in streams)                        # every sequence starts from zero (loader.py:292-293)
    else:
        for s in streams:                                               # default / longterm: a stream is one pass over one scene
            assert len({order[e][:13] for e in s}) == 1
            assert int(order[s[0]].split("_")[-1].split(".")[0]) == 0
    seen = []
    per_slot_last = [None] * n_slots
    stream_of = {e: k for k, s in enumerate(streams) for e in s}
    pos_in_stream = {}
    for step in sched:
        assert len(step.assign) == n_slots
        busy = [a for a in step.assign if a is not None]
        assert busy and len(set(busy)) == len(busy)
        for s, a in enumerate(step.assign):
            if a is None:
                assert not step.reset[s] and not step.seq_start[s]
                continue
            e, f = a
            seen.append(a)
            assert step.seq_start[s] == (f == 0)
            k = stream_of[e]
            # inside a stream, frames come in the reference's serial order, one per frame-step, always on the same slot
            flat = [(ee, ff) for ee in streams[k] for ff in range(n_frames[ee])]
            p = pos_in_stream.get(k, 0)
            assert flat[p] == a
            pos_in_stream[k] = p + 1
            if p == 0:
                assert step.reset[s]                                    # a slot that takes a stream starts from zero
            else:
                assert not step.reset[s] and per_slot_last[s] == flat[p - 1]
            per_slot_last[s] = a
    assert sorted(seen) == sorted((e, f) for e in range(len(order)) for f in range(n_frames[e]))     # every frame exactly once
    assert sched.total_frames == len(seen)


def test_schedule_rank_partition_keeps_streams_whole(eod):
    g = _golden()
    order = eod.formats.order_files(g["files"], "default")
    n_frames = _lengths(order)
    reset_first = [eod.formats.memory_reset_flag("default", f, 0) for f in order]
    world = 4
    parts = [eod.LockStepSchedule(n_frames, reset_first, 8, r, world) for r in range(world)]
    eps = [set(p.episodes) for p in parts]
    assert set().union(*eps) == set(range(len(order))) and sum(len(e) for e in eps) == len(order)
    for p in parts:
        for s in p.streams:
            assert len({order[e][:13] for e in s}) == 1
    # the same assignment sharding.shard_scenes makes
    for r in range(world):
        assert sorted(eps[r]) == eod.sharding.shard_scenes(order, r, world)
    with pytest.raises(ValueError):
        eod.LockStepSchedule(n_frames, reset_first, 0)
    with pytest.raises(ValueError):
        eod.LockStepSchedule(n_frames, reset_first, 4, 4, 4)


def test_schedule_waves_bound_resident_grids(eod):
    """BASELINE configs[4] shape: 512 independent 100-frame episodes through R resident slots - never more than R streams in
    flight, every slot busy until the queue drains, ceil(512 / R) waves of 100 frame-steps."""
    n, T, R = 512, 100, 56
    sched = eod.LockStepSchedule([T] * n, [True] * n, R)
    steps = list(sched)
    assert len(steps) == -(-n // R) * T
    assert all(sum(a is not None for a in st.assign) == R for st in steps[: (n // R) * T])
    assert sum(sum(a is not None for a in st.assign) for st in steps) == n * T
    assert sum(sum(st.reset) for st in steps) == n


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import importlib
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    eod = importlib.import_module("embodied-object-detection_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = _golden()
    order = eod.formats.order_files(g["files"], "longterm")
    n_frames = _lengths(order)
    reset_first = [eod.formats.memory_reset_flag("longterm", f, 0) for f in order]
    sched = eod.LockStepSchedule(n_frames, reset_first, 5, rank, world)
    frames = sum(sum(a is not None for a in st.assign) for st in sched)
    tot = eod.sharding.gather_counters({"frames": float(frames), "streams": float(len(sched.streams))}, torch.device("cpu"))
    q.put((rank, frames, tot, sum(n_frames), sched.n_streams_total))
    dist.destroy_process_group()


def test_schedule_two_ranks_over_gloo_cover_the_dataset_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(60) for p in procs]
    for _, frames, tot, total, n_streams in res:
        assert tot["frames"] == float(total) and tot["streams"] == float(n_streams)
        assert 0 < frames < total
