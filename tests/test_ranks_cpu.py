"""CPU, world_size 2 over gloo: the N > 1 host protocol of bench.py (rank discovery, host-side barriers around the steps rank 0
runs, max / sum reduction of the per-rank scalars).  The data path itself has no collective; the sharding of one
draft over several engine contexts is the product's own (fb_fillgaps.cpp) and is tested on the GPUs
(test_sharding_over_two_contexts_is_invariant)."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def test_two_ranks_meet_at_host_barriers_and_reduce():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29731",
           os.path.join(HERE, "rank_harness.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    line = [l for l in p.stdout.decode().split("\n") if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and d["max"]["rank"] == 1.0
    assert d["sum"]["work"] == 3000.0                                      # work: summed over ranks
    # the idle rank waited for rank 0 at every barrier: every rank's region covers the work, and the maximum is what is reported
    assert d["dt_rank0"] >= 0.15 * d["steps"] and d["max"]["dt"] >= d["dt_rank0"] - 1e-9


def test_single_rank_passthrough():
    sys.path.insert(0, os.path.dirname(HERE))
    from figbird_b200.ranks import rank_info, reduce_counters
    mx, sm = reduce_counters({"a": 2.0}, None)
    assert mx == {"a": 2.0} and sm == {"a": 2.0}
    assert rank_info() == (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0")))
