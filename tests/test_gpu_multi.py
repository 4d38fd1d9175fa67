"""Partitioned (one process per GPU, NCCL halo + all-reduce) == single GPU.  Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped on a single-GPU box."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.parametrize("shape,nproc,pc", [("6x5x8", 2, "jacobi"), ("9x14", 2, "jacobi"), ("5x5x12", 4, "jacobi"),
                                            ("12x12x16", 2, "mg"), ("40x48", 2, "mg"), ("4x3x4", 2, "mg"),
                                            ("10x10x24", 4, "mg")])
def test_partitioned_matches_single_gpu(lib, shape, nproc, pc):
    import torch

    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(ROOT / "tests" / "multi_gpu_worker.py"), shape, pc]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
