"""One rank of an interval-sharded `modkit pileup` under torchrun (tests/test_gpu_sharded.py, NCCL on GPUs).
usage: torchrun --nproc-per-node N tests/sharded_worker.py <args of modkit pileup ...>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import modkit_b200
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    ndev = torch.cuda.device_count()
    devi = local % max(1, ndev)
    torch.cuda.set_device(devi)
    dist.init_process_group("nccl" if ndev >= world else "gloo")
    dev = torch.device("cuda", devi) if ndev >= world else None
    modkit_b200.bind_host_thread(devi)
    rc, st = modkit_b200.pileup_main_sharded(sys.argv[1:] + ["--device", str(devi), "--quiet"], rank, world, modkit_b200.torch_allreduce(dev))
    dist.barrier()
    dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
