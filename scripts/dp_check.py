"""2+ rank data-parallel parity (torchrun, one process per GPU): the all-reduced gradients of a DP step must equal the
single-process emulation "run every rank's batch slice separately, average the flat gradients" (SURVEY.md section 8e).
The check itself lives in vb.distributed.dp_gradient_parity (bench.py runs it too at N > 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import vickers_hardness_unet_b200 as vb


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rel, mx, same = vb.distributed.dp_gradient_parity(dev, rank, world)
    if rank == 0:
        print(f"dp_check world={world}: reduced-vs-emulated gradient rel-L2 {rel:.3e} max-abs {mx:.3e}; "
              f"identical across ranks: {same}")
    ndiff = vb.distributed.dp_overlap_update_parity(dev, rank, world)
    if rank == 0:
        print(f"dp_check world={world}: parameters after 3 updates, bucket-wise overlapped vs joined all-reduce: "
              f"{ndiff} of {vb.Unet('resnet34').flat_params.numel()} differ")
    dist.destroy_process_group()
    if not (rel < 1e-3 and same and ndiff == 0):
        sys.exit(1)
    if rank == 0:
        print("DP_CHECK_OK")


if __name__ == "__main__":
    main()
