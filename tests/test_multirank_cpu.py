"""CPU, world_size 2 over gloo: the N>1 host logic of the replica path -- batch sharding with no
data-path collective, barrier, max-over-ranks timing, sticky stream placement.  Each rank runs the
package's model code on ITS shard (kernels stood in by the oracle, tests/oracle_backend.py) and the
parent checks that the concatenated shards equal the single-process result."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Patch:
    """Minimal stand-in for pytest's monkeypatch inside the spawned ranks."""

    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _model():
    import video_mamba
    torch.manual_seed(0)
    return video_mamba.PretrainVideoMamba(img_size=16, patch_size=8, depth=2, embed_dim=16, channels=3,
                                          ssm_cfg={"use_fast_path": False}, num_frames=4).eval()


def _rank_main(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import oracle_backend
    from videomamba_b200.replica import init_replica_group, shard_bounds

    oracle_backend.install(_Patch())
    grp = init_replica_group(backend="gloo")
    assert (grp.rank, grp.world) == (rank, world)
    model = _model()
    clips = torch.rand(5, 3, 4, 16, 16, generator=torch.Generator().manual_seed(7))
    lo, hi = shard_bounds(clips.shape[0], world, rank)
    grp.barrier()
    with torch.no_grad():
        vis, pool = model(clips[lo:hi])
    # timing plumbing: every rank must see the maximum
    assert grp.max_over_ranks(10.0 + rank) == 10.0 + world - 1
    grp.barrier()
    torch.save({"lo": lo, "hi": hi, "vis": vis, "pool": pool}, os.path.join(out_dir, f"rank{rank}.pt"))
    grp.close()


def test_two_rank_batch_sharding_matches_single_process(tmp_path, monkeypatch):
    world = 2
    port = _free_port()
    mp.spawn(_rank_main, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend
    oracle_backend.install(monkeypatch)
    clips = torch.rand(5, 3, 4, 16, 16, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        want_vis, want_pool = _model()(clips)
    parts = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    assert [(p["lo"], p["hi"]) for p in parts] == [(0, 3), (3, 5)]
    assert torch.equal(torch.cat([p["vis"] for p in parts]), want_vis)
    assert torch.equal(torch.cat([p["pool"] for p in parts]), want_pool)


def test_shard_bounds_cover_and_are_disjoint():
    from videomamba_b200.replica import shard_bounds
    for total in (0, 1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_stream_placement_is_sticky_and_balanced():
    from videomamba_b200.replica import StreamPlacement
    pl = StreamPlacement(world=8, slots_per_rank=32)
    first = {s: pl.place(s) for s in range(256)}
    assert all(pl.place(s) == first[s] for s in range(256))           # sticky
    per_rank = [sum(1 for r, _ in first.values() if r == k) for k in range(8)]
    assert per_rank == [32] * 8                                         # 256 streams over 8 GPUs
    assert len({first[s] for s in range(256)}) == 256                   # distinct (rank, slot)
    with pytest.raises(RuntimeError):
        pl.place(1000)
    pl.release(5)
    assert pl.place(1000)[0] == first[5][0]
    mine = [s for s in range(256) if first[s][0] == 3 and s != 5]
    assert sorted(pl.local_slots(3, mine)) == sorted(first[s][1] for s in mine)
    with pytest.raises(ValueError):
        pl.local_slots(0, [s for s in range(256) if first[s][0] == 1][:1])
