"""CPU, world_size 2 and 4 over gloo: the multi-rank host logic (decomposition, neighbour relation, ownership at upload,
snapshot merge of the parity driver) without a GPU -- see tests/run_multi_cpu.py."""
import os
import subprocess
import sys
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,port", [(2, 29517), (4, 29518)])
def test_multi_rank_host_logic_gloo(world, port):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(HERE, "run_multi_cpu.py")], capture_output=True, text=True, timeout=600)
    assert "MULTI-RANK HOST LOGIC OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
