"""Host -> device input prefetch for the training loop.

The reference engine copies each batch on the compute stream right before the forward
(engine_pretrain.py:50, `samples.to(device, non_blocking=True)`), so the PCIe transfer (77 MB per
step for two 64 x 3 x 224 x 224 fp32 batches) sits on the critical path.  `DevicePrefetcher` wraps any
iterable of pinned host batches and issues the copy of batch i+1 on a side stream while batch i
trains; the compute stream only waits on the copy's event.  Drop-in around a DataLoader:

    for samples, target in DevicePrefetcher(data_loader, device):
        loss, _, _ = model(samples, mask_ratio=0.75)
"""
import torch


class DevicePrefetcher:
    def __init__(self, iterable, device, depth=2):
        self.iterable = iterable
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = max(1, depth)

    def __len__(self):
        return len(self.iterable)

    def _to_device(self, obj):
        if isinstance(obj, torch.Tensor):
            return obj.to(self.device, non_blocking=True)
        if isinstance(obj, (list, tuple)):
            return type(obj)(self._to_device(o) for o in obj)
        return obj

    def _record(self, obj, stream):
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda:
                obj.record_stream(stream)
        elif isinstance(obj, (list, tuple)):
            for o in obj:
                self._record(o, stream)

    def __iter__(self):
        it = iter(self.iterable)
        queue = []

        def issue():
            try:
                batch = next(it)
            except StopIteration:
                return False
            with torch.cuda.stream(self.stream):
                dev_batch = self._to_device(batch)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            queue.append((dev_batch, ev))
            return True

        for _ in range(self.depth):
            if not issue():
                break
        while queue:
            dev_batch, ev = queue.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            self._record(dev_batch, cur)      # the caching allocator must not recycle it while `cur` uses it
            issue()
            yield dev_batch
