"""Worker for tests/test_host_logic.py::test_two_rank_gloo_sharding_and_max_time (launched by torchrun)."""
import os
import sys

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slowflow_b200.shard import max_over_ranks, shard_range, sum_over_ranks  # noqa: E402

dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(7, r, w)
t = max_over_ranks(10.0 + 5.0 * r)
n = sum_over_ranks(hi - lo)
assert t == 15.0 and n == 7, (t, n)
dist.barrier()
dist.destroy_process_group()
sys.stdout.write("rank=%d:ok:%d:%d\n" % (r, lo, hi))
sys.stdout.flush()
