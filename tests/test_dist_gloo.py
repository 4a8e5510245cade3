"""N > 1 host logic on CPU: two processes over gloo (127.0.0.1) partition a global batch of
independent sequences, run a stand-in step function on their slice, and must reproduce the
single-process token matrix; the timing reduction takes the slowest rank."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import llama2_ts_b200 as pkg
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
def step(tok, pos):                       # stand-in for l2b_forward_batch: depends on token AND position
    return ((tok.astype(np.int64) * 1103515245 + 12345 + pos * 7919) %% 31999 + 1).astype(np.int32)
B = 7                                      # not divisible by 2: ragged partition
sb = pkg.dist.ShardedBatch(B, step)
out = sb.run(np.arange(10, 10 + B, dtype=np.int32), 5)
t = pkg.dist.max_over_ranks(1.0 + dist.get_rank())
if dist.get_rank() == 0:
    np.save(sys.argv[2], out)
    open(sys.argv[2] + ".t", "w").write(str(t))
dist.barrier()
dist.destroy_process_group()
'''


def test_partition_is_balanced_and_contiguous(pkg):
    for n in (1, 7, 8, 256, 257):
        for world in (1, 2, 4, 8):
            parts = [pkg.dist.partition(n, world, r) for r in range(world)]
            assert sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert pkg.dist.partition(256, 8, 3) == (96, 32)


def test_two_ranks_over_gloo_match_single_process(pkg, tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "port": port})
    out = str(tmp_path / "out.npy")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = np.load(out)

    def step(tok, pos):
        return ((tok.astype(np.int64) * 1103515245 + 12345 + pos * 7919) % 31999 + 1).astype(np.int32)
    want = pkg.dist.ShardedBatch(7, step).run(np.arange(10, 17, dtype=np.int32), 5)   # world 1
    assert np.array_equal(got, want)
    assert float(open(out + ".t").read()) == 2.0        # max over ranks of (1.0, 2.0)
