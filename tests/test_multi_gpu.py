"""Multi-GPU tests of the library path over NCCL (need >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    import torch
    return torch.cuda.device_count()


def torchrun(n, script_args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + script_args
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)


def test_allreduced_count_table_equals_the_oracle_over_the_union():
    n = n_gpus()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    r = torchrun(n, [os.path.join(ROOT, "tests", "mp_nccl_counts.py")], 29431)
    assert r.returncode == 0 and "NCCL_COUNTS_OK world=%d" % n in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


@pytest.mark.parametrize("config", ["C2", "C5"])
def test_bench_line_on_all_gpus(config):
    """bench.py under torchrun: one JSON line on stdout (nothing else), the checks inside bench.py hold."""
    n = n_gpus()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    r = torchrun(2, ["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", "--config", config, "--scale", "0.1"], 29433)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0


def test_two_contexts_on_two_devices_in_one_process():
    """One process, a ctx per device: every ctx sets up its own device (kernel attributes, tables); scan and diff on the
    second device give the first device's numbers."""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))
    import tosemscan as ts
    c = ts.gen_corpus(31, 600, 1, n_groups=3, pinned=False)
    a, b = ts.gen_pairs(0x7053454D0005, 400)
    outs = []
    for dev in (0, 1):
        sc = ts.Scanner(dev, int(c.off[-1]) + 4096, c.n_files, 4)
        res = sc.scan(c, ts.SCAN_ASSERT_EVENTS)
        add, rem, det = sc.diff_pairs(a, b, detail=True)
        outs.append((res["stats"].copy(), res["group_counts"].copy(), len(res["assert_events"]), add.copy(), rem.copy(), det.copy()))
        sc.close()
    for x, y in zip(outs[0], outs[1]):
        assert np.array_equal(x, y)
