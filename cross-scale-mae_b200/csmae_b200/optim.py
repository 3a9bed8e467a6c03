"""Fused optimizer step for the hot path.

`FusedAdamW` is a drop-in for `torch.optim.AdamW(param_groups, lr=..., betas=(0.9, 0.95))` as the reference builds
it (main_pretrain.py:426-427): same constructor, same `param_groups` (so `lr_sched.adjust_learning_rate` and
`GradScaler.step/unscale_` work unchanged), same per-parameter state keys (`step`, `exp_avg`, `exp_avg_sq`).  The step
itself is ONE kernel over all parameters (`csm_adamw_multi`) that also rewrites the bf16 shadow copies of the GEMM
weights the engine's tcgen05 kernels read, so the separate fp32 -> bf16 cast pass before the next forward
disappears.  `grad_norm()` is the global L2 norm the reference computes with one `torch.norm` per parameter
(util/misc.py:338-355); when the gradients are views of the engine's flat buffer it is one reduction kernel.

`NativeScalerWithGradNormCount` is the drop-in for the reference's loss scaler of the same name
(util/misc.py:299-335): same call signature, same state_dict (GradScaler's keys), but with a FusedAdamW the whole
unscale_ / inf check / grad norm / clip / step / scale update chain is three launches with NO host synchronisation:
one pass over the flat gradient buffer (`csm_grad_stats_f32`: norm of the unscaled gradients + found_inf), one scalar
kernel (`csm_amp_update`) and `csm_adamw_multi`, which multiplies the gradients by 1 / scale (x clip coefficient) on
the fly and skips the update when an inf / nan was found.
"""
import math

import numpy as np
import torch

from . import _native as nat

_ENTRY = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("w16", "<u8"), ("n", "<i8"),
                   ("lr", "<f4"), ("wd", "<f4"), ("beta1", "<f4"), ("beta2", "<f4"),
                   ("eps", "<f4"), ("bc1", "<f4"), ("bc2_sqrt", "<f4"), ("grad_scale", "<f4")])
assert _ENTRY.itemsize == 80
_CHUNK = 8192


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, model=None):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._engine = getattr(model, "_engine", None) if model is not None else None
        self._cache = None          # pointer table of the last step (reused while no pointer moves)

    # ------------------------------------------------------------------ helpers
    def _shadow_ptrs(self):
        """data_ptr of the master weight -> data_ptr of its bf16 shadow (engine-owned), if the shadows exist."""
        eng = self._engine
        if eng is None or not eng._w16_views:
            return {}
        params = dict(eng.model.named_parameters())
        return {params[n].data_ptr(): v.data_ptr() for n, v in eng._w16_views.items() if n in params}

    def _build(self, items, dev):
        shadows = self._shadow_ptrs()
        tab = np.zeros(len(items), dtype=_ENTRY)
        chunks = []
        for i, (group, p, st) in enumerate(items):
            tab[i]["p"], tab[i]["g"] = p.data_ptr(), p.grad.data_ptr()
            tab[i]["m"], tab[i]["v"] = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            tab[i]["w16"] = shadows.get(p.data_ptr(), 0)
            tab[i]["n"] = p.numel()
            chunks += [(i, c) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
        host = torch.empty(tab.nbytes, dtype=torch.uint8).pin_memory()
        cache = dict(tab=tab, host=host, host_np=host.numpy().view(_ENTRY),
                     dev=torch.empty(tab.nbytes, dtype=torch.uint8, device=dev),
                     chunks=torch.tensor(chunks, dtype=torch.int32).to(dev), nchunks=len(chunks),
                     key=self._key(items), shadowed=bool(shadows), items=items)
        return cache

    @staticmethod
    def _key(items):
        # every pointer the device table holds: a parameter, gradient or moment tensor that was replaced (zero_grad
        # with set_to_none, load_state_dict on resume, .to()) invalidates the table
        return tuple((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr())
                     for _, p, st in items)

    def load_state_dict(self, state_dict):
        """Resume (util/misc.py:393-411 calls optimizer.load_state_dict): accepts a torch.optim.AdamW or FusedAdamW
        state dict; the moment tensors are new objects afterwards, so the cached pointer table is dropped."""
        super().load_state_dict(state_dict)
        self._cache = None
        for st in self.state.values():
            for k in ("exp_avg", "exp_avg_sq"):
                if k in st and (st[k].dtype != torch.float32 or not st[k].is_contiguous()):
                    st[k] = st[k].float().contiguous()

    # ------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None, ctl=None):
        """ctl (optional, device f32[2]): {gradient multiplier, found_inf} read by the kernel from device memory
        (NativeScalerWithGradNormCount below); without it this is torch.optim.AdamW.step."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        items = []
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise nat.NativeError("FusedAdamW handles fp32 CUDA parameters and gradients only")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                items.append((group, p, st))
        if not items:
            return loss
        dev = items[0][1].device
        cache = self._cache
        if (cache is None or len(cache["items"]) != len(items) or cache["key"] != self._key(items)
                or cache["shadowed"] != bool(self._engine is not None and self._engine._w16_views)):
            cache = self._cache = self._build(items, dev)
        tab = cache["tab"]
        # hyper-parameters are per group and the step count is per parameter (state_dict format of torch.optim.AdamW);
        # parameters of a group that have always stepped together share one vectorised fill
        start = 0
        while start < len(items):
            group = items[start][0]
            stop = start
            while stop < len(items) and items[stop][0] is group:
                stop += 1
            steps = []
            for _, _, st in items[start:stop]:
                st["step"] += 1
                steps.append(float(st["step"]))
            b1, b2 = group["betas"]
            sl = slice(start, stop)
            tab["lr"][sl], tab["wd"][sl] = group["lr"], group["weight_decay"]
            tab["beta1"][sl], tab["beta2"][sl], tab["eps"][sl] = b1, b2, group["eps"]
            t = np.asarray(steps, dtype=np.float64)
            tab["bc1"][sl] = 1.0 - b1 ** t
            tab["bc2_sqrt"][sl] = np.sqrt(1.0 - b2 ** t)
            tab["grad_scale"][sl] = 1.0
            start = stop
        cache["host_np"][:] = tab
        cache["dev"].copy_(cache["host"], non_blocking=True)
        nat.call("csm_adamw_multi", cache["dev"], cache["chunks"], cache["nchunks"], ctl, nat.sm_count(dev))
        self._last_stepped = [st for _, _, st in items]
        eng = self._engine
        if eng is not None:
            if cache["shadowed"]:
                # the kernel rewrote the shadows and did not bump Tensor._version: the engine's version check
                # keeps seeing "unchanged" and skips its cast pass -- which is the point
                pass
            else:
                eng._w16_versions = None        # shadows did not exist yet: the next forward builds them
        return loss

    def rollback_step(self):
        """Undo the step COUNT of the last step() (the kernel skipped the update because found_inf was set):
        torch.optim.AdamW under GradScaler never sees a skipped step, so its bias corrections do not advance."""
        for st in getattr(self, "_last_stepped", []):
            st["step"] -= 1
        self._last_stepped = []

    # ------------------------------------------------------------------ gradient norm
    def _grad_spans(self):
        """The gradients as the fewest 16-byte-aligned flat fp32 spans: the engine hands out consecutive views of
        ONE flat buffer (padding between views is zero), which is then a single span."""
        grads = [p.grad for g in self.param_groups for p in g["params"] if p.grad is not None]
        if not grads:
            return [], []
        base = min(grads, key=lambda t: t.data_ptr())
        span_end = max(t.data_ptr() + t.numel() * 4 for t in grads)
        span = (span_end - base.data_ptr()) // 4
        same_storage = all(t.untyped_storage().data_ptr() == base.untyped_storage().data_ptr() for t in grads)
        if (same_storage and all(t.dtype == torch.float32 for t in grads)
                and span <= sum((t.numel() + 3) // 4 * 4 for t in grads) and base.data_ptr() % 16 == 0):
            return [torch.as_strided(base, (span,), (1,))], []
        ok = [t for t in grads if t.dtype == torch.float32 and t.data_ptr() % 16 == 0 and t.is_contiguous()]
        return [t.view(-1) for t in ok], [t for t in grads if not any(t is o for o in ok)]

    @torch.no_grad()
    def grad_norm(self):
        """Global L2 norm of all gradients (util/misc.py:338-355) as a 0-d device tensor."""
        spans, odd = self._grad_spans()
        if not spans and not odd:
            return torch.zeros(())
        dev = (spans + odd)[0].device
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        for t in spans:
            nat.call("csm_sumsq_f32", t, t.numel(), out, nat.sm_count(dev))
        for t in odd:
            out += t.float().pow(2).sum()
        return out.sqrt().reshape(())


class NativeScalerWithGradNormCount:
    """Drop-in for util/misc.py:299-335 (same name, call signature, `state_dict_key` and state_dict keys).

    With a FusedAdamW optimizer nothing between `backward()` and the end of the step touches the host: the gradients
    stay scaled in memory (the reference's `unscale_` rewrites them in place; here the optimizer kernel applies
    1 / scale x clip on the fly, so `p.grad` read AFTER the call still carries the loss scale), the inf check
    and the skip decision live in device memory, and the scale update is `csm_amp_update`.  The returned norm is a
    0-d device tensor of the UNSCALED gradient norm, as the reference returns.  The step count of a skipped update
    is rolled back at the next call (the found_inf flag is copied to pinned memory asynchronously and looked at one
    step late), so bias corrections match torch.optim.AdamW under GradScaler exactly.
    With any other optimizer the stock torch GradScaler sequence of the reference runs unchanged."""
    state_dict_key = "amp_scaler"

    def __init__(self, init_scale=2.0 ** 16, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self._init = (float(init_scale), float(growth_factor), float(backoff_factor), int(growth_interval))
        self._growth_factor, self._backoff_factor, self._growth_interval = self._init[1:]
        self._state = None          # device f32[4]: scale, growth_tracker, 1 / scale, unused
        self._pending = None        # state to upload at first use (load_state_dict before any step)
        self._work = None           # device f32[8]: stats[2], ctl[2], norm[1]
        self._host_flag = None
        self._flag_event = None
        self._flag_opt = None
        self._stock = None          # torch GradScaler for optimizers other than FusedAdamW

    # ---- state
    def _ensure(self, dev):
        if self._state is None or self._state.device != dev:
            scale, tracker = self._pending if self._pending is not None else (self._init[0], 0)
            self._state = torch.tensor([scale, float(tracker), 1.0 / scale, 0.0], dtype=torch.float32).to(dev)
            self._work = torch.zeros(8, dtype=torch.float32, device=dev)
            self._host_flag = torch.zeros(2, dtype=torch.float32).pin_memory()
            self._pending = None
        return self._state

    def get_scale(self):
        if self._stock is not None:
            return self._stock.get_scale()
        if self._state is None:
            return (self._pending or (self._init[0],))[0]
        return float(self._state[0].item())

    def state_dict(self):
        if self._stock is not None:
            return self._stock.state_dict()
        if self._state is None:
            scale, tracker = self._pending if self._pending is not None else (self._init[0], 0)
        else:
            scale, tracker = (float(x) for x in self._state[:2].tolist())
        return {"scale": scale, "growth_factor": self._growth_factor, "backoff_factor": self._backoff_factor,
                "growth_interval": self._growth_interval, "_growth_tracker": int(tracker)}

    def load_state_dict(self, state_dict):
        if self._stock is not None:
            self._stock.load_state_dict(state_dict)
            return
        self._growth_factor = float(state_dict["growth_factor"])
        self._backoff_factor = float(state_dict["backoff_factor"])
        self._growth_interval = int(state_dict["growth_interval"])
        self._pending = (float(state_dict["scale"]), int(state_dict["_growth_tracker"]))
        self._state = None

    # ---- the step
    def _reconcile(self):
        """Looks at the previous step's found_inf flag (already on the host in any loop that reads the loss) and takes
        back the step count of a skipped update."""
        if self._flag_event is not None:
            self._flag_event.synchronize()
            if self._host_flag[1].item() != 0.0 and self._flag_opt is not None:
                self._flag_opt.rollback_step()
            self._flag_event = None
            self._flag_opt = None

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True):
        if not isinstance(optimizer, FusedAdamW):
            if self._stock is None:
                self._stock = torch.amp.GradScaler("cuda", init_scale=self.get_scale() if self._state is not None or
                                                   self._pending is not None else self._init[0],
                                                   growth_factor=self._growth_factor,
                                                   backoff_factor=self._backoff_factor,
                                                   growth_interval=self._growth_interval)
            sc = self._stock
            sc.scale(loss).backward(create_graph=create_graph)
            if not update_grad:
                return None
            sc.unscale_(optimizer)
            if clip_grad is not None:
                assert parameters is not None
                norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
            else:
                ps = [p for p in ([parameters] if isinstance(parameters, torch.Tensor) else parameters)
                      if p.grad is not None]
                norm = (torch.norm(torch.stack([torch.norm(p.grad.detach(), 2.0) for p in ps]), 2.0)
                        if ps else torch.tensor(0.0))
            sc.step(optimizer)
            sc.update()
            return norm
        state = self._ensure(loss.device)
        (loss * state[0]).backward(create_graph=create_graph)
        if not update_grad:
            return None
        if clip_grad is not None:
            assert parameters is not None
        self._reconcile()
        dev = loss.device
        work = self._work
        stats, ctl, norm = work[0:2], work[2:4], work[4:5]
        with torch.no_grad():
            stats.zero_()
            spans, odd = optimizer._grad_spans()
            for t in spans:
                nat.call("csm_grad_stats_f32", t, t.numel(), state[2:3], stats, nat.sm_count(dev))
            for t in odd:
                u = t.float() * state[2]
                stats[0] += u.pow(2).sum()
                stats[1] = torch.maximum(stats[1], (~torch.isfinite(u)).any().float())
            nat.call("csm_amp_update", stats, state, ctl, norm, float(clip_grad) if clip_grad is not None else 0.0,
                     self._growth_factor, self._backoff_factor, self._growth_interval)
            optimizer.step(ctl=ctl)
            self._host_flag.copy_(ctl, non_blocking=True)
            self._flag_event = torch.cuda.Event()
            self._flag_event.record()
            self._flag_opt = optimizer
        return norm.clone().reshape(())
