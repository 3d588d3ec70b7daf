"""N>1 path.  CPU: world_size 2 and 3 with the gloo backend (host logic: plans, message layout,
ordering).  GPU: world_size = number of devices (>= 2) with NCCL inside the library."""
import os
import signal
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch(n, backend, cases, port, timeout=600, stress=0, extra_env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "dist_check.py"), "--backend", backend, "--cases", str(cases),
           "--stress", str(stress)]
    # tiny pipeline pieces so that the cutting of remote boxes is exercised by the small test cases
    env = dict(os.environ, OMP_NUM_THREADS="1", SBB_CHUNK_BYTES="256", **(extra_env or {}))
    # own session, so that a hang can be ended together with every worker process (a worker left
    # spinning on a GPU would disturb whatever runs next)
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env,
                            cwd=ROOT, start_new_session=True)
    try:
        out, _ = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, _ = proc.communicate()
        raise AssertionError("multi-rank check timed out after %d s:\n%s" % (timeout, out[-3000:]))
    assert proc.returncode == 0, out[-3000:]
    assert "failures=0" in out, out[-3000:]


@pytest.mark.parametrize("n", [2, 3])
def test_gloo_world(n):
    launch(n, "gloo", 25, 29511 + n)


@pytest.mark.gpu
@pytest.mark.parametrize("transport", ["peer-memory", "nccl-send-recv"])
def test_nccl_world(transport):
    """One process per GPU: random copies and contractions against the oracle and a stress run of
    back-to-back exchanges, over the peer-memory transport (default) and over plain ncclSend/ncclRecv
    (SBB_P2P=0, also the automatic fallback when the arenas cannot be mapped)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    launch(min(n, 8), "nccl", 20, 29531 + (transport != "peer-memory"), timeout=300, stress=300,
           extra_env={} if transport == "peer-memory" else {"SBB_P2P": "0"})
