#!/usr/bin/env python3
"""Run under torchrun (one rank per GPU): sharded detection with the CUDA engine + NCCL gather,
compared on rank 0 with the unsharded result of the same GPU.  Prints one line."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nanomod_b200 as nm
from nanomod_b200.sharded import ShardedDetector


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    det = nm.Detector(lr)
    p = nm.synthetic_pileup(200_000, 40, 40, drop_frac1=0.005, two_strands=True, poisson=True, clip=(3, 100), round_decimals=3)
    opt = nm.DetectOptions(neighborPvalues=3, both_combinations=True)
    table = ShardedDetector(det).detect(p, opt)
    ok = True
    if rank == 0:
        full = det.detect(p, opt)
        for c in ("row_pos_index", "n0", "n1", "ks_dnum", "ks_p", "two_u", "u_p", "t_stat", "t_p", "stouffer_stat",
                  "stouffer_p", "fisher_stat", "fisher_p", "pos", "seg"):
            same = getattr(table, c).tobytes() == getattr(full, c).tobytes()
            ok &= same
            if not same:
                print("MISMATCH", c)
        print("multi_gpu_check world=%d rows=%d identical_to_single_gpu=%s called=%s" % (world, len(table), ok, table.called_sites()[:3]))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
