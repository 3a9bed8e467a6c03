"""Host -> device input prefetch for the training loop.

The reference engine copies each batch on the compute stream right before the forward
(engine_pretrain.py:50, `samples.to(device, non_blocking=True)`), so the PCIe transfer (77 MB per
step for two 64 x 3 x 224 x 224 fp32 batches) sits on the critical path.  `DevicePrefetcher` wraps any
iterable of pinned host batches and issues the copy of batch i+1 on a side stream while batch i
trains; the compute stream only waits on the copy's event.  Drop-in around a DataLoader:

    for samples, target in DevicePrefetcher(data_loader, device):
        loss, _, _ = model(samples, mask_ratio=0.75)

Device memory: a ring of depth + 1 persistent slots, allocated once per instance and kept across epochs (no
allocation per batch or per epoch: a cudaMalloc in the caching allocator serialises against the GPU and, with an NVML
/ nvidia-smi poller running beside the job, was measured to block for 50-300 ms); a slot is overwritten only after
the compute stream has passed the point where the batch after it was handed out.  A yielded batch is therefore valid
until the next-but-`depth` batch is requested; keep a `.clone()` if it has to live longer.  Create ONE prefetcher
around the loader and iterate it every epoch.
"""
import torch


class DevicePrefetcher:
    def __init__(self, iterable, device, depth=2):
        self.iterable = iterable
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = max(1, depth)
        self._slots = [dict(bufs=None, free=None) for _ in range(self.depth + 1)]

    def __len__(self):
        return len(self.iterable)

    def _ensure_buffers(self, slot, batch):
        """(Re)allocates the slot's device tensors when the batch layout changes (first batch, ragged last batch)."""
        flat, _ = _flatten(batch)
        bufs = slot["bufs"]
        if bufs is None or len(bufs) != len(flat) or any(
                isinstance(h, torch.Tensor) != isinstance(d, torch.Tensor)
                or (isinstance(h, torch.Tensor) and (d.shape != h.shape or d.dtype != h.dtype))
                for h, d in zip(flat, bufs)):
            if bufs is not None and slot["free"] is not None:
                slot["free"].synchronize()                    # the old tensors go back to the allocator
            slot["bufs"] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) if isinstance(h, torch.Tensor)
                            else h for h in flat]

    def _copy_into(self, slot, batch):
        """Copies the (nested) host batch into the slot's persistent device tensors on the side stream."""
        flat, spec = _flatten(batch)
        bufs = slot["bufs"]
        out = []
        for h, d in zip(flat, bufs):
            if isinstance(h, torch.Tensor):
                d.copy_(h, non_blocking=True)
                out.append(d)
            else:
                out.append(h)
        return _unflatten(out, spec)

    def __iter__(self):
        it = iter(self.iterable)
        queue = []
        state = dict(next_slot=0)

        def issue():
            try:
                batch = next(it)
            except StopIteration:
                return False
            slot = self._slots[state["next_slot"]]
            state["next_slot"] = (state["next_slot"] + 1) % len(self._slots)
            self._ensure_buffers(slot, batch)                 # (first use only) from the caller's stream pool
            with torch.cuda.stream(self.stream):
                if slot["free"] is not None:
                    self.stream.wait_event(slot["free"])      # its previous batch is no longer in use
                dev_batch = self._copy_into(slot, batch)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            queue.append((dev_batch, ev, slot))
            return True

        for _ in range(self.depth):
            if not issue():
                break
        prev_slot = None
        try:
            while queue:
                dev_batch, ev, slot = queue.pop(0)
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                if prev_slot is not None:
                    # the consumer has moved on from the previous batch: everything it enqueued so far precedes this
                    done = torch.cuda.Event()
                    done.record(cur)
                    prev_slot["free"] = done
                prev_slot = slot
                issue()
                yield dev_batch
        finally:
            # end of the epoch (or the consumer left early): the slots handed out or still queued may only be
            # overwritten by the next epoch after what the consumer has enqueued up to now
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))
            for sl in [prev_slot] + [q[2] for q in queue]:
                if sl is not None:
                    sl["free"] = done


def _flatten(obj):
    if isinstance(obj, (list, tuple)):
        flat, specs = [], []
        for o in obj:
            f, s = _flatten(o)
            flat += f
            specs.append((len(f), s))
        return flat, (type(obj), specs)
    return [obj], None


def _unflatten(flat, spec):
    if spec is None:
        return flat[0]
    typ, specs = spec
    out, i = [], 0
    for n, s in specs:
        out.append(_unflatten(flat[i:i + n], s))
        i += n
    return typ(out)
