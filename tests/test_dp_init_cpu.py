"""Control flow of the data-parallel communicator set-up (tensorfrost_b200.nca_dp.init_comm) at world_size 2 over gloo, with a scripted
stand-in for the device module: every rank must take the SAME decision (peer-memory exchange or NCCL fallback) and nobody may enter the
peer kernel - which spins on its peers' flags - unless every rank mapped its peers.  Cases: all good -> "peer"; one rank cannot map its
peers -> everyone "nccl", no peer exchange attempted; the peer exchange returns wrong numbers on one rank -> everyone "nccl"."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, %r)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
from tensorfrost_b200 import nca_dp
scenario = sys.argv[1]


class T:
    def __init__(self, a): self.a = np.array(a, np.float32)


class FakeTf:
    """Answers what init_comm calls; both exchanges are a gloo allreduce (scenario "wrong": the peer path is off by one on rank 0)."""
    peer_calls = 0
    def cuda_comm_unique_id(self): return b"x" * 128
    def cuda_comm_init(self, uid, r, w): assert len(uid) == 128
    def cuda_peer_export(self): return bytes([rank]) * 64
    def cuda_peer_init(self, handles, r, w):
        assert [h[0] for h in handles] == list(range(w))
        if scenario == "unmappable" and r == 1:
            raise RuntimeError("cudaIpcOpenMemHandle: peer access is not supported between these two devices")
    def cuda_tensor(self, a): return T(a)
    def cuda_numpy(self, t): return t.a
    def cuda_allreduce(self, t, scale, method):
        if method == "peer":
            FakeTf.peer_calls += 1
        x = torch.from_numpy(t.a.copy())
        dist.all_reduce(x)
        t.a = (x.numpy() * np.float32(scale)).astype(np.float32)
        if method == "peer" and scenario == "wrong" and rank == 0:
            t.a = t.a + 1.0


tf = FakeTf()
method = nca_dp.init_comm(tf, rank, world)
print("METHOD", method, FakeTf.peer_calls, flush=True)
dist.destroy_process_group()
''' % ROOT


def _run(tmp_path, scenario):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29600 + (os.getpid() % 300)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        env.pop("TFCUDA_DP_EXCHANGE", None)
        procs.append(subprocess.Popen([sys.executable, str(script), scenario], env=env, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-2000:]
    return [[l for l in o.splitlines() if l.startswith("METHOD")][0].split() for o in outs]


@pytest.mark.parametrize("scenario,expected,peer_calls", [("good", "peer", 1), ("unmappable", "nccl", 0), ("wrong", "nccl", 1)])
def test_every_rank_takes_the_same_exchange_decision(tmp_path, scenario, expected, peer_calls):
    results = _run(tmp_path, scenario)
    assert [r[1] for r in results] == [expected, expected], results
    assert [int(r[2]) for r in results] == [peer_calls, peer_calls], results
