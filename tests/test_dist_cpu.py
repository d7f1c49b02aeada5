"""world_size-2 checks of the multi-rank plumbing on CPU (gloo): rendezvous, the 128-byte
NCCL-id broadcast, max-over-ranks timing reduction and shard ownership -- everything bench.py
does around the C ABI when torchrun starts one rank per GPU."""
from __future__ import annotations

import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r})
sys.path.insert(0, os.path.join({root!r}, "cuda-to-sycl-nbody_b200"))
import bench
import nbody_b200 as nb
d = bench.Dist(backend="gloo")
payload = bytes(range(128)) if d.rank == 0 else None
got = d.broadcast_bytes(payload, 128)
assert got == bytes(range(128)), got[:8]
mx = d.max(10.0 + d.rank)
sm = d.sum(1.0 + d.rank)
d.barrier()
n = 1000001
b, c = nb.plan_shard(n, d.world, d.rank)
tot = d.sum(float(c))
print(json.dumps({{"rank": d.rank, "world": d.world, "max": mx, "sum": sm, "begin": b, "count": c, "total": tot}}))
d.close()
'''


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_plumbing(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=180)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d["rank"])
    assert [o["world"] for o in outs] == [2, 2]
    assert all(o["max"] == 11.0 and o["sum"] == 3.0 for o in outs)
    assert outs[0]["begin"] == 0 and outs[1]["begin"] == outs[0]["count"]
    assert all(o["total"] == 1000001.0 for o in outs)


def test_reference_arm_runs_rank0_only_under_two_ranks(tmp_path):
    """bench.py --impl reference under torchrun-style env: rank 0 prints the line, rank 1 exits 0"""
    import json
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="4")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                                       "--gpus", "2", "--steps", "1", "--warmup", "0", "--bodies", "8192"],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = [p.communicate(timeout=300) + (p.returncode,) for p in procs]
    assert all(rc == 0 for _, _, rc in res), [e[-1500:] for _, e, _ in res]
    line = json.loads(res[0][0].strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    assert res[1][0].strip() == ""
