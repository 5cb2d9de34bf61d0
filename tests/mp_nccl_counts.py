"""Helper of tests/test_multi_gpu.py, one process per GPU (torch.distributed.run): every rank scans its round-robin shard
through the C ABI, the count table of the library (tsm_device_counts, a device pointer) is allreduced in place over
NCCL, and every rank checks the result against the oracle over the UNION of the shards."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tosem-2021-replication_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc  # noqa: E402
import tosemscan as ts  # noqa: E402


class _Arr:
    def __init__(self, p, m):
        self.__cuda_array_interface__ = {"shape": (m,), "typestr": "<i8", "data": (p, False), "version": 3}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    seed, n_total, groups = 0x7053454D0004, 6000, 9
    mine = ts.gen_corpus(seed, (n_total - rank + world - 1) // world, 1, 4096, first_index=rank, index_stride=world, n_groups=groups)
    sc = ts.Scanner(local, int(mine.off[-1]) + 4096, mine.n_files, 16)
    sc.upload(mine)
    sc.scan_resident(0)
    ptr, n64 = sc.device_counts()
    counts = torch.as_tensor(_Arr(ptr, n64), device=torch.device("cuda", local))
    dist.all_reduce(counts)                                  # the one collective of the path
    got = counts.cpu().numpy().reshape(-1)
    whole = ts.gen_corpus(seed, n_total, 1, 4096, n_groups=groups, pinned=False)
    want = orc.scan(whole.arena, whole.off, whole.len, whole.ext, whole.grp, groups, events=False)
    K = 128
    assert np.array_equal(got[:groups * K].reshape(groups, K), want["group_counts"]), "group counts after the allreduce"
    assert np.array_equal(got[groups * K:(groups + 1) * K], want["global_counts"]), "global counts after the allreduce"
    st = want["stats"]
    assert got[(groups + 1) * K:(groups + 1) * K + 4].tolist() == [int(st[k].astype(np.int64).sum()) for k in ("n_lines", "n_assert", "n_headers", "n_fixture")]
    # the shards are disjoint and cover the corpus: per-file records of this rank = the oracle's records of its files
    res = sc.download(0)
    assert np.array_equal(res["stats"], st[rank::world])
    dist.barrier()
    if rank == 0:
        print("NCCL_COUNTS_OK world=%d files=%d assertion_lines=%d" % (world, n_total, int(want["global_counts"].sum())))
    sc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
