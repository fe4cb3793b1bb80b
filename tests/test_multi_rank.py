"""N > 1 path on CPU: one process per rank (gloo), font x GlyphBlock sharding, no data-path collective."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ranks_cover_the_job(world):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29650 + world), os.path.join(ROOT, "tests", "_rank_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"] and res["world"] == world and res["files"] == 512
    assert sum(res["per_rank_blocks"]) == 512
    # shards are balanced by estimated cost (LPT over font x block tasks), not by block count
    costs = res["per_rank_cost"]
    assert max(costs) <= 1.1 * sum(costs) / world, costs


def test_reference_arm_prints_contract_line():
    """bench.py --impl reference runs the oracle port on the host cores and prints the contract's JSON line."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "fira"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    res = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert res["impl"] == "reference" and res["unit"] == "glyphs/s" and res["value"] > 0
    assert res["cpu_baseline"]["kind"] == "port" and res["e2e"]["h2d_bytes_per_step"] == 0
