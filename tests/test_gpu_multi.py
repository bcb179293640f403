"""Multi-GPU parity of the peer-memory all-reduce: runs only where >= 2 GPUs are visible."""
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_comm_ranks(built, world):
    from base_b200 import groundwork as gw
    if gw.device_count() < world:
        pytest.skip(f"needs {world} GPUs (ranks that wait on one another cannot share a device)")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
         "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
         str(ROOT / "tests" / "multigpu" / "peer_comm_ranks.py")],
        cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert rep["bits_equal_checker"] and rep["graph_replay_ok"] and rep["missing_peer_reported"]
    print(rep)
