"""The optional NCCL gather of the C ABI (`melspec_gather_nccl`).  CPU: the entries exist and refuse politely without a
communicator.  GPU (needs >= 2 devices, skipped otherwise): tools/gather_nccl_check.py under torchrun, the library's all-gather
against torch.distributed's on every rank."""
import ctypes as C
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gather_entries_validate_arguments_without_a_device():
    import mel_spec_b200 as ms
    ms.build()
    L = ms.lib()
    assert L.melspec_nccl_unique_id(None) == 4                       # MELSPEC_ERR_INVALID_ARG
    assert L.melspec_nccl_init(None, None, 0, 1) == 4
    assert L.melspec_gather_nccl(None, None, 0, None, None) == 4
    assert L.melspec_nccl_destroy(None) == 4


@pytest.mark.gpu
def test_gather_nccl_matches_torch_all_gather_on_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with `gpurun --gpus 2`)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", os.path.join(ROOT, "tools", "gather_nccl_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    assert json.loads(line)["ok"] is True
