"""Multi-GPU device path: every rank's rows, built by the CUDA library with peer-store halos — the
exchange folded into the build and the separate packing kernel — against the oracle on the global system.  Needs two
GPUs on the box (the single-GPU test box skips it; `gpurun --gpus 2 -- python -m pytest tests -m gpu -k multi_gpu`)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("world", [2, 3])
def test_halo_rows_match_oracle_on_every_rank(world):
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(29530 + world), os.path.join(ROOT, "tools", "halo_rows_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "HALO ROWS OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
