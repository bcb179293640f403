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
    assert rep["sharded_step_bits_equal_world_1"] and rep["fused_graph_replay_ok"]
    print(rep)


def test_cpp_driver_drives_the_peer_kernel_without_python(built, tmp_path):
    """The collective is callable from a C++ host program: fork per GPU, handles through shared
    memory, b9gw_* only.  Runs at the largest world the box allows (1 on a single-GPU box)."""
    from base_b200 import groundwork as gw
    exe = tmp_path / "peer_comm_c"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", f"-I{ROOT / 'include'}",
                    str(ROOT / "tests" / "multigpu" / "peer_comm_c.cpp"), "-o", str(exe),
                    f"-L{ROOT / 'base_b200'}", "-lb9_groundwork", f"-L{ROOT / 'oracle'}", "-lb9_groundwork_ref",
                    f"-Wl,-rpath,{ROOT / 'base_b200'}:{ROOT / 'oracle'}"], check=True)
    world = max(w for w in (1, 2, 4, 8) if w <= gw.device_count())
    r = subprocess.run([str(exe), str(world)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["world"] == world and rep["bits_equal_checker"] is True
    print(rep)
