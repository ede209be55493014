"""N>1 host logic on CPU: world_size-2 gloo ranks shard the envs and all-reduce the episode accumulator."""
import os
import subprocess
import sys
import textwrap

import numpy as np

import common
from pupperv3_mjx_b200 import parallel, prng

WORKER = textwrap.dedent('''
    import os, sys, json
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, sys.argv[1])
    from pupperv3_mjx_b200 import parallel
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    keys = parallel.shard_keys(0, 37, rank, world)
    totals = torch.zeros(24)
    totals[0] = 1 + rank            # episodes finished on this rank
    totals[1] = 10.0 * (1 + rank)   # their summed reward
    totals[2] = 100.0 * (1 + rank)
    totals[3:22] = float(rank)
    red = parallel.EpisodeMetricsReducer(totals)
    red.reduce()                       # interval 1: moves the local sums out, all-reduces them
    assert float(totals.abs().sum()) == 0.0
    red.reduce()                       # interval 2: nothing finished -> nothing is counted twice
    totals[0] = 1.0; totals[1] = 5.0 * (1 + rank)   # interval 3: one more episode per rank
    red.reduce()
    g = red.global_totals.clone(); g[0] -= world; g[1] -= 5.0 * world * (world + 1) / 2
    rep = parallel.episode_report(g)
    rep["n_reduces"] = red.n_reduces
    gathered = [None] * world
    dist.all_gather_object(gathered, keys.tolist())
    if rank == 0:
        print("RESULT " + json.dumps({"rep": rep, "keys": gathered}))
    dist.destroy_process_group()
''')


def test_shard_range_covers_everything():
    for n, w in ((37, 2), (4096, 8), (5, 8), (65536, 4)):
        spans = [parallel.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def test_two_rank_allreduce_and_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script), common.ROOT]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    res = json.loads(line[len("RESULT "):])
    rep = res["rep"]
    assert rep["episodes"] == 3.0 and abs(rep["sum_reward"] - 10.0) < 1e-6 and abs(rep["length"] - 100.0) < 1e-6
    assert abs(rep["total_dist"] - 1.0 / 3.0) < 1e-6 and rep["n_reduces"] == 3
    keys = np.array(res["keys"][0] + res["keys"][1], dtype=np.uint32)
    np.testing.assert_array_equal(keys, prng.split(prng.PRNGKey(0), 37))
