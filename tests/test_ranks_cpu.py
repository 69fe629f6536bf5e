"""CPU, world_size 2 over gloo: the N>1 host path of bench.py (rank discovery, max / sum reduction of the per-rank
counters, cost-balanced gap sharding).  The data path itself has no collective."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_ranks_reduce_and_shard():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(HERE, "rank_harness.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    line = [l for l in p.stdout.decode().split("\n") if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2
    assert d["max"]["dt"] == 1.25 and d["max"]["dev_ms"] == 200.0          # time-like: max over ranks
    assert d["sum"]["placements"] == 3000.0                                # work: summed over ranks
    assert d["shard_n_sum"] == 40                                          # every gap on exactly one rank
    costs = [float((7 * i) % 13 + 1) for i in range(40)]
    assert d["shard_cost_sum"] == sum(costs)
    assert d["shard_cost_max"] <= sum(costs) / 2 + max(costs)              # LPT balance bound


def test_single_rank_passthrough():
    sys.path.insert(0, os.path.dirname(HERE))
    from figbird_b200.ranks import reduce_counters, shard_gaps
    mx, sm = reduce_counters({"a": 2.0}, None)
    assert mx == {"a": 2.0} and sm == {"a": 2.0}
    assert sorted(sum(shard_gaps([3, 1, 2], 1), [])) == [0, 1, 2]
