"""Host-side logic that needs no GPU: Vocabulary (C++ mirror of src/asr/types.rs:77-155), the utterance sharder, and
the multi-rank partition exercised with torch.distributed/gloo at world_size 2 (the data path has no collective)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vocabulary_matches_oracle(amira, oracle, tmp_path):
    p = tmp_path / "vocab.txt"
    lines = ["<unk> 0", "▁the 1", "cat 2", "▁s at 3", "broken line", "x notanumber", "dup 2", "▁ 4", "<blk> 1024", ""]
    p.write_text("\n".join(lines), encoding="utf-8")
    v = amira.Vocabulary.load_from_file(str(p))
    o = oracle.Vocabulary.load_from_file(str(p))
    for ids in ([1, 2, 3], [999, 1, 1030, 2], [], [4, 4, 1], [2, 1024, 0], list(range(5)) * 3):
        assert v.decode_tokens(ids) == o.decode_tokens(ids), ids
    assert v.decode_tokens([1, 2]) == "thedup"  # later line overwrites (HashMap::insert); "▁" prefix -> space, trimmed


def test_reference_vocab_file_shape(amira, tmp_path):
    # model-repo/vocab.txt: 1025 lines "<token> <id>", "<blk> 1024"; a synthetic file of the same shape
    p = tmp_path / "vocab.txt"
    p.write_text("".join(f"{'▁' if i % 3 == 0 else ''}t{i} {i}\n" for i in range(1024)) + "<blk> 1024\n", encoding="utf-8")
    v = amira.Vocabulary(str(p))
    assert v.decode_tokens([3, 4, 5, 6, 1024]) == "t3t4t5 t6<blk>"


def test_missing_vocab_is_an_io_error(amira):
    with pytest.raises(amira.AmiraError) as e:
        amira.Vocabulary("/nonexistent/vocab.txt")
    assert e.value.code == 8


def test_shard_utterances_balanced_and_deterministic(amira):
    rng = np.random.default_rng(4567)
    costs = rng.integers(80000, 480000, size=8192)
    for g in (1, 2, 4, 8):
        s = amira.shard_utterances(costs, g)
        assert s.min() == 0 and s.max() == g - 1
        loads = np.bincount(s, weights=costs, minlength=g)
        assert loads.max() / loads.mean() < 1.001  # LPT: within 0.1 % of perfect balance
        assert np.array_equal(s, amira.shard_utterances(costs, g))
    assert amira.shard_utterances([], 4).size == 0
    assert amira.shard_utterances([5, 5, 5], 8).tolist() == [0, 1, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import amira_b200 as A
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(4567)
    costs = rng.integers(80000, 480000, size=64)
    shard = A.shard_utterances(costs, world)          # every rank computes the same map, takes its own part
    mine = np.nonzero(shard == rank)[0]
    # stand-in for the per-rank hot path: "transcripts" = utterance ids; host-side gather is the only exchange
    gathered = [None] * world
    dist.all_gather_object(gathered, mine.tolist())
    t = torch.tensor([float(costs[mine].sum())], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, gathered, float(t.item()), float(costs.sum())))


def test_two_rank_partition_with_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, total, expect in res:
        ids = sorted(i for part in gathered for i in part)
        assert ids == list(range(64))          # every utterance decoded exactly once across ranks
        assert abs(total - expect) < 1e-6
