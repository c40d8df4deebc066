"""TEST-ONLY communicator: the interface of hpvg.dist.NcclCommunicator on torch.distributed/gloo, so the sharding and
gather logic of the sampling path can run with world_size 2 on CPU (the product path uses NCCL through ctypes)."""
import numpy as np
import torch
import torch.distributed as dist


class GlooCommunicator:
    def __init__(self):
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def all_gather_rows(self, rows, stream=None):
        a = np.asarray(rows.numpy(stream) if hasattr(rows, "numpy") and not isinstance(rows, np.ndarray) else rows)
        t = torch.from_numpy(np.ascontiguousarray(a))
        outs = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(outs, t)
        return torch.cat(outs).numpy()

    def barrier(self):
        dist.barrier()
