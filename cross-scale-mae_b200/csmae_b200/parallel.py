"""Data parallelism for the hot path: one process per GPU, replicas kept identical by
averaging the gradients over NCCL (NVLink 5 / NVSwitch).

The module also runs unchanged under torch's own wrapper exactly as the reference builds it
(`torch.nn.parallel.DistributedDataParallel(model, device_ids=[gpu], find_unused_parameters=True)`,
main_pretrain.py:417-421) -- tools/ddp_smoke.py exercises that.  But the hand-written backward hands all
gradients to autograd at once, so torch's bucketed all-reduce cannot overlap it and pays ~260 bucket copies in
and out.  `DistributedDataParallel` below is the drop-in for that wrapper (same constructor call, `.module`,
forward passthrough) that lets the engine do the exchange itself: the flat fp32 gradient buffer is all-reduced
in 1 + enc_groups segments as the backward chain completes them, on NCCL's stream, overlapping the rest of
the backward; no per-parameter hooks, no bucket copies.

Semantics kept from the reference setup: rank 0's parameters and buffers are broadcast at construction (C2),
BatchNorm running statistics are re-broadcast from rank 0 before every training forward (torch DDP's default
`broadcast_buffers=True`, C4), BatchNorm batch statistics and NT-Xent negatives stay per rank, each rank's loss
is its own (the engine's logger all-reduces it separately, C5).
"""
import torch
import torch.distributed as dist
import torch.nn as nn


class DistributedDataParallel(nn.Module):
    def __init__(self, module, device_ids=None, output_device=None, find_unused_parameters=True,
                 broadcast_buffers=True, process_group=None, enc_groups=3, **_unused):
        super().__init__()
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("csmae_b200.DistributedDataParallel needs an initialised torch.distributed process group")
        self.module = module
        self.process_group = process_group
        self.broadcast_buffers = broadcast_buffers
        self.world_size = dist.get_world_size(process_group)
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, 0, group=process_group)
        module._engine.enable_grad_sync(process_group, self.world_size, enc_groups)

    def forward(self, *args, **kwargs):
        if self.broadcast_buffers and self.module.training and self.world_size > 1:
            with torch.no_grad():
                for b in self.module.buffers():
                    dist.broadcast(b, 0, group=self.process_group)
        return self.module(*args, **kwargs)

    def no_sync(self):
        raise NotImplementedError("gradient accumulation without synchronisation is not on the benchmarked path; "
                                  "use torch.nn.parallel.DistributedDataParallel for it")
