"""Executor of the Cross-Scale MAE hot path: one forward and one hand-written backward over the
C-ABI kernels (include/csmae_b200.h), with torch supplying only device memory and streams.

Dataflow restated from the reference (paths relative to the upstream repo):
  models_mae/MAE_ViT_Baseline.py:243-320   forward_encoder / forward_decoder / forward
  models_mae/MAE_ViT_MsLd.py:37-77         two scales, loss_orig + loss_crop
  models_mae/MAE_ViT_MsLdCeCd.py:27-84     + predictor cross-decoder loss + NT-Xent
Both scales run as ONE batched pass of 2N images (nothing in a Block mixes samples):
images [0, N) are scale 1 ("orig"), [N, 2N) scale 2 ("crop").

Precision policy (independent of the ambient autocast state, SURVEY.md 0.8 / 8a'): fp32 master
weights and residual stream, bf16 tensor-core operands with fp32 accumulation, Linear outputs rounded
to bf16 where the reference's autocast graph rounds them, LayerNorm / softmax / losses in fp32.

Launch overhead: a step is ~640 kernels; after two eager warm-up steps per (shape, mode) the forward chain and the
backward chain are each captured into a CUDA graph (torch.cuda.CUDAGraph over the same C-ABI launches, PDL edges
included) and replayed -- inputs are copied into static buffers, the masking noise is still drawn by the caller from
the global generator.  CSMAE_CUDA_GRAPHS=0 keeps every step eager.

Memory: activations needed by the backward live in a per-model workspace that is reused every step
(the backward of step i always precedes the forward of step i+1).  A generation counter makes a
stale backward fail loudly instead of reading overwritten activations.
"""
import os

import torch

from . import _native as nat
from ._native import EPI_BF16, EPI_DGELU, EPI_F32, EPI_GELU, EPI_RESID, call

LN_EPS = 1e-6          # MAE_ViT_Baseline.py:43-45
BN_EPS = 1e-5          # nn.BatchNorm1d default (MLP.py:7)
BN_MOMENTUM = 0.1
NTXENT_TAU = 0.5       # MAE_ViT_MsLdCeCd.py:62
NTXENT_EPS = 1e-8      # util/contrast_loss.py:51

_GEMM_WEIGHT_SUFFIXES = ("qkv.weight", "proj.weight", "fc1.weight", "fc2.weight")


def grad_segments(names, offsets, total, enc_layers, enc_groups):
    """Slices [start, end) of the flat gradient buffer in the order the backward chain completes them, for the
    overlapped gradient all-reduce.  The buffer follows named_parameters() order (cls_token, mask_token,
    patch_embed, decoder_embed, encoder.0.., decoder.0.., decoder_pred, decoder_norm, predictor); the backward
    runs decoder_pred/predictor/decoder_norm, decoder blocks, decoder_embed, encoder blocks (descending), patch_embed.
      segment 0      : [first decoder.* parameter, total)        complete after the decoder blocks
      segments 1..G-1: encoder layer groups, highest layers first
      last segment   : [0, start of the second-lowest group)     complete at the very end
    Returns (segments, boundaries): boundaries[k] = the encoder layer after whose backward segment k (k >= 1) is
    complete; segment 0's boundary is the end of the decoder backward and the last one the end of the chain."""
    first = {}
    for n, o in zip(names, offsets):
        key = ".".join(n.split(".")[:2]) if n.startswith(("encoder.", "decoder.")) and n.split(".")[1].isdigit() else n
        first.setdefault(key, o)
    dec0 = first.get("decoder.0")
    if dec0 is None or enc_layers == 0 or "encoder.0" not in first:
        return [(0, total)], []
    groups = max(1, min(enc_groups, enc_layers))
    bounds = [round(j * enc_layers / groups) for j in range(groups + 1)]          # 0 = b_0 < ... < b_G = Le
    segs, layers = [(dec0, total)], []
    for j in range(groups - 1, 0, -1):
        start = first[f"encoder.{bounds[j]}"]
        end = first[f"encoder.{bounds[j + 1]}"] if bounds[j + 1] < enc_layers else dec0
        segs.append((start, end))
        layers.append(bounds[j])
    segs.append((0, first[f"encoder.{bounds[1]}"] if groups > 1 else dec0))
    return segs, layers


class _GraphSequence:
    """Captures one call chain into several CUDA graphs: cut() ends the current capture and starts the next."""

    def __init__(self, stream=None):
        self.graphs = []
        self._ctx = None
        self._pool = None
        self._stream = stream

    def begin(self):
        g = torch.cuda.CUDAGraph()
        kw = {} if self._pool is None else {"pool": self._pool}
        if self._stream is not None:
            kw["stream"] = self._stream
        # "relaxed": with an NCCL process group in the process, the capturing thread itself can end up in a call CUDA
        # classes as potentially unsafe (an event query or free from a collected Work / Event object, an allocator
        # call) and a stricter mode then invalidates the whole capture (seen on 2-GPU boxes, tests/test_ddp_gpu.py);
        # none of those calls is part of the captured work
        self._ctx = torch.cuda.graph(g, capture_error_mode="relaxed", **kw)
        self._ctx.__enter__()
        self.graphs.append(g)

    def cut(self):
        self.end()
        self.begin()

    def end(self):
        self._ctx.__exit__(None, None, None)
        if self._pool is None:
            self._pool = self.graphs[0].pool()
        self._ctx = None


class HotPathEngine:
    def __init__(self, model, use_cd=False, use_ce=False):
        self.model = model
        self.use_cd = use_cd
        self.use_ce = use_ce
        self.generation = 0
        # Workspaces: captured CUDA graphs bake in the raw pointers of every buffer they touch, so each graph entry OWNS
        # the buffers it was captured over (entry["bufs"], allocated inside the capture from the graph's pool); eager
        # steps use their own dictionary, which may reallocate freely when shapes change.
        self._eager_bufs = {}
        self._bufs = self._eager_bufs     # the active workspace
        self._state = None          # what the last forward saved for the backward
        self._w16 = None            # flat bf16 shadow of every GEMM weight
        self._w16_views = {}
        self._w16_key = None
        self._w16_versions = None
        self._cast_table = None
        self._names = None
        self._coefs = {}            # loss coefficient tensors by (values, device)
        self._last_out = None
        self._snap = None
        self.use_graphs = os.environ.get("CSMAE_CUDA_GRAPHS", "1") != "0"
        self._graphs = {}           # key -> dict(fwd graph, outputs, state, static inputs, bwd graph, ...)
        self._warm = {}             # key -> eager steps seen so far
        self.graph_warmup_steps = 2
        self.use_side_stream = os.environ.get("CSMAE_SIDE_STREAM", "1") != "0"
        # stream-K scheduling of the forward / dgrad GEMMs whose tiles do not fill whole rounds of SM pairs (they all
        # run on the caller's stream -- the side stream only carries wgrads and column sums -- as the workspace
        # requires).  It wins 15-18 % on the long-reduction encoder GEMMs timed alone, but inside the backward the
        # side stream's wgrads already fill the SMs a partial last round leaves idle and the step gets 1.5 % slower
        # (13.71 -> 13.92 ms, ViT-B), hence "fwd": forward GEMMs only
        self.stream_k_mode = os.environ.get("CSMAE_STREAM_K", "0")        # "0" | "1" (fwd + bwd) | "fwd"
        self._side = {}
        self._side_dirty = False
        # Row chains (see _parts; opt-in, CSMAE_CHAINS=2): the images of a step are split into two halves whose
        # encoder / decoder block chains run on two streams -- nothing in a Block mixes samples, so the halves are
        # independent until the losses, and while one half drains the tail of a kernel the other half's next kernel
        # could take the idle SMs.  Measured on B200 (profiles/r2e_chains_ab.jsonl): slower on every workload
        # (ViT-B 13.94 -> 14.46 ms, ViT-L 14.67 -> 16.53 ms): the half-size GEMMs lose more to tile quantisation and
        # fixed cost per launch than the overlap returns, so the default stays one chain.
        self.num_chains = max(1, min(2, int(os.environ.get("CSMAE_CHAINS", "1"))))
        self.chain_min_images = 8            # below this a step is launch-bound, not tail-bound: one chain
        self._chain = {}
        self._chain_forked = False
        self._sync_enabled = False  # overlapped gradient all-reduce (parallel.py)
        self._sync_group = None     # its process group (None = the default group)
        self._sync_world = 1
        self._sync_groups = 3       # encoder layer groups -> 1 + groups all-reduce segments

    def enable_grad_sync(self, group, world, enc_groups=3):
        """Data-parallel gradient averaging inside the backward: the flat gradient buffer is all-reduced in
        segments as the chain completes them (NCCL on its own stream), overlapping the rest of the backward."""
        self._sync_enabled = world > 1
        self._sync_group, self._sync_world, self._sync_groups = group, world, enc_groups
        self._graphs.clear()

    # ------------------------------------------------------------------ second stream for the weight gradients
    # wgrad GEMMs and bias column sums only feed the flat gradient buffer: they run on a side stream so that their
    # CTAs fill the SMs the critical chain (dgrad / LayerNorm / attention backward) leaves idle in its tails.  All
    # buffers they read are per-layer, so the side stream may lag without write-after-read hazards; it is joined
    # before every gradient-segment boundary.
    def _side_stream(self, dev):
        s = self._side.get(dev)
        if s is None:
            s = torch.cuda.Stream(device=dev)      # same priority as the chain: both orders of preference measured slower
            self._side[dev] = s
        return s

    def _on_side(self, dev, fn, after=None):
        """Runs fn on the side stream once everything issued so far on the current stream (and on the streams in
        `after`: the row chains that produced what fn reads) has completed."""
        if not self.use_side_stream:
            for s in after or ():
                self._wait(torch.cuda.current_stream(dev), s)
            fn()
            return
        side = self._side_stream(dev)
        self._wait(side, torch.cuda.current_stream(dev))
        for s in after or ():
            self._wait(side, s)
        with torch.cuda.stream(side):
            fn()
        self._side_dirty = True

    @staticmethod
    def _wait(waiter, producer):
        if waiter is producer or waiter == producer:
            return
        ev = torch.cuda.Event()
        ev.record(producer)
        waiter.wait_event(ev)

    # ------------------------------------------------------------------ row chains
    def _chain_stream(self, dev):
        s = self._chain.get(dev)
        if s is None:
            s = torch.cuda.Stream(device=dev)
            self._chain[dev] = s
        return s

    def _parts(self, NB, dev):
        """[(stream, first image, images)] of the row chains: chain 0 runs on the caller's stream, chain 1 on a second
        stream (forked / joined by events, captured into the same CUDA graphs)."""
        main = torch.cuda.current_stream(dev)
        if self.num_chains < 2 or NB < 2 * self.chain_min_images:
            return [(main, 0, NB)]
        h = NB // 2
        return [(main, 0, h), (self._chain_stream(dev), h, NB - h)]

    def _fork(self, parts):
        """Chain streams start after everything issued so far on the caller's stream."""
        for s, *_ in parts[1:]:
            self._wait(s, parts[0][0])

    def _join(self, parts):
        """The caller's stream continues after every chain."""
        for s, *_ in parts[1:]:
            self._wait(parts[0][0], s)

    def _join_side(self, dev):
        if self.use_side_stream and self._side_dirty:
            ev = torch.cuda.Event()
            ev.record(self._side_stream(dev))
            torch.cuda.current_stream(dev).wait_event(ev)
            self._side_dirty = False

    def _issue_allreduce(self, flat, seg, pending):
        import torch.distributed as dist
        start, end = seg
        if end > start:
            # SUM: the 1/world factor is folded into the loss gradient the chain starts from (every kernel of the
            # backward is linear in it), so no pass over the buffer is spent on the division
            pending.append(dist.all_reduce(flat[start:end], op=dist.ReduceOp.SUM, group=self._sync_group, async_op=True))

    # ------------------------------------------------------------------ parameters
    def param_names(self):
        """Trainable parameters the hot path differentiates, in named_parameters() order.
        encoder_norm.* is registered but never used (reference: MAE_ViT_Baseline.py:264)."""
        if self._names is None:
            names = []
            for n, p in self.model.named_parameters():
                if not p.requires_grad or n.startswith("encoder_norm."):
                    continue
                if n.startswith("predictor.") and not self.use_cd:
                    continue
                names.append(n)
            self._names = names
        return self._names

    def snapshot(self):
        """Per-step host bookkeeping cached across steps: parameter dict / list, GEMM weights, pointer key.
        Rebuilt when a probe of parameter storage pointers changes (module.to(...), re-assigned .data) and
        fully re-validated every 64 steps; what it saves is ~0.5 ms of Python between a step's loss read-back and
        the next step's first kernel."""
        snap = self._snap
        if snap is not None and snap["age"] < 64 and all(p.data_ptr() == q for p, q in snap["probe"]):
            snap["age"] += 1
            return snap
        m = self.model
        params = dict(m.named_parameters())
        self._names = None
        names = self.param_names()
        plist = [params[n] for n in names]
        gemm = [(n, p) for n, p in params.items() if self._is_gemm_weight(n, p)]
        ptr_key = (tuple(p.data_ptr() for p in params.values()), tuple(b.data_ptr() for b in m.buffers()))
        if snap is not None and snap["ptr_key"] == ptr_key and snap["names"] == names:
            snap["age"] = 0
            return snap
        allp = list(params.values())
        probe = [(p, p.data_ptr()) for p in (allp[0], allp[len(allp) // 2], allp[-1])]
        self._snap_serial = getattr(self, "_snap_serial", 0) + 1
        self._snap = dict(params=params, names=names, plist=plist, gemm=gemm, gemm_params=[p for _, p in gemm],
                          ptr_key=ptr_key, probe=probe, age=0, serial=self._snap_serial)
        return self._snap

    def _is_gemm_weight(self, name, p):
        return p.dim() >= 2 and name.endswith(".weight") and "norm" not in name and not name.startswith("predictor.1")

    def _refresh_weights(self, snap):
        """bf16 shadow copies of the GEMM weights, refreshed with one multi-tensor cast kernel when any
        master weight changed (optimizer steps bump Tensor._version)."""
        gemm = snap["gemm"]
        key = snap["ptr_key"]
        versions = sum(p._version for p in snap["gemm_params"])
        dev = gemm[0][1].device
        if key is not self._w16_key and key != self._w16_key:
            offsets, total = {}, 0
            for n, p in gemm:
                offsets[n] = total
                total += (p.numel() + 7) // 8 * 8
            self._w16 = torch.empty(total, dtype=torch.bfloat16, device=dev)
            self._w16_views = {n: self._w16[offsets[n]:offsets[n] + p.numel()] for n, p in gemm}
            table = []
            for n, p in gemm:
                table += [p.data_ptr(), self._w16_views[n].data_ptr(), p.numel()]
            self._cast_table = torch.tensor(table, dtype=torch.int64).to(dev)
            self._w16_key = key
            self._w16_versions = None
        if versions != self._w16_versions:
            call("csm_cast_multi", self._cast_table, len(gemm), 32)
            self._w16_versions = versions
        return self._w16_views

    # ------------------------------------------------------------------ workspace
    def _buf(self, name, shape, dtype, device):
        t = self._bufs.get(name)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype or t.device != device:
            t = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[name] = t
        return t

    def workspace_bytes(self):
        seen, total = set(), 0
        for d in [self._eager_bufs] + [e["bufs"] for e in self._graphs.values()]:
            for t in d.values():
                if t.data_ptr() not in seen:
                    seen.add(t.data_ptr())
                    total += t.numel() * t.element_size()
        return total

    # ------------------------------------------------------------------ forward
    def forward(self, imgs_list, noises, mask_ratio, training):
        dev = imgs_list[0].device
        if dev.type != "cuda":
            raise nat.NativeError("csmae_b200 runs on sm_100 CUDA devices only (no CPU fallback): got " + str(dev))
        with torch.cuda.device(dev):             # _native.call launches on the current device's current stream
            return self._forward(imgs_list, noises, mask_ratio, training)

    def _forward(self, imgs_list, noises, mask_ratio, training):
        snap = self.snapshot()
        self._refresh_weights(snap)              # outside any graph: runs only when a master weight changed
        if not self.use_graphs or torch.cuda.is_current_stream_capturing():
            self._bufs = self._eager_bufs
            return self._forward_eager(imgs_list, noises, mask_ratio, training)
        key = self._graph_key(imgs_list, noises, mask_ratio, training, snap)
        entry = self._graphs.get(key)
        if entry is None:
            seen = self._warm.get(key, 0)
            if seen < self.graph_warmup_steps:
                self._warm[key] = seen + 1
                self._bufs = self._eager_bufs
                return self._forward_eager(imgs_list, noises, mask_ratio, training)
            entry = self._capture_forward(key, imgs_list, noises, mask_ratio, training)
        self._bufs = entry["bufs"]
        for dst, src in zip(entry["imgs"], imgs_list):
            dst.copy_(src, non_blocking=True)
        for dst, src in zip(entry["noise"], noises):
            dst.copy_(src, non_blocking=True)
        entry["fwd"].replay()
        nat.launch_count += entry["fwd_launches"]
        self.generation += 1
        entry["state"]["generation"] = self.generation
        self._state = entry["state"]
        self._active_graph = entry
        return entry["out"]

    def _graph_key(self, imgs_list, noises, mask_ratio, training, snap):
        return (tuple(tuple(im.shape) for im in imgs_list), float(mask_ratio), bool(training),
                imgs_list[0].device.index, snap["serial"])

    def _capture_forward(self, key, imgs_list, noises, mask_ratio, training):
        if len(self._graphs) >= 4:               # shapes keep changing: stop hoarding graphs
            self._graphs.clear()
        entry = dict(imgs=[torch.empty_like(im, dtype=torch.float32).contiguous() for im in imgs_list],
                     noise=[torch.empty_like(nz, dtype=torch.float32).contiguous() for nz in noises], bwd=None,
                     bufs={})
        # the eager warm-up workspace is not needed once the step replays from graphs: give its memory back
        self._eager_bufs.clear()
        self._bufs = entry["bufs"]
        # the H2D copy of the loss coefficients must not happen inside the capture
        self._coefs_tensor(self._coef_values(len(imgs_list), imgs_list[0].shape[0], mask_ratio)[0], imgs_list[0].device)
        for dst, src in zip(entry["imgs"], imgs_list):
            dst.copy_(src)
        for dst, src in zip(entry["noise"], noises):
            dst.copy_(src)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        c0 = nat.launch_count
        with torch.cuda.graph(g, capture_error_mode="relaxed"):
            out = self._forward_eager(entry["imgs"], entry["noise"], mask_ratio, training)
        entry["fwd_launches"] = nat.launch_count - c0      # kernels recorded, not run: counted at each replay
        nat.launch_count = c0
        entry["fwd"], entry["out"], entry["state"] = g, out, self._state
        self._graphs[key] = entry
        return entry

    def _forward_eager(self, imgs_list, noises, mask_ratio, training):
        m = self.model
        dev = imgs_list[0].device
        nsm = nat.sm_count(dev)
        if self.stream_k_mode != "0":
            nat.enable_gemm_stream_k(dev, True)       # host-side switch, read when a GEMM is launched / captured
        ns = len(imgs_list)
        N, C, H, W = imgs_list[0].shape
        assert H == W == m.input_size and C == m.input_channels, "input size mismatch"
        NB = N * ns
        L = m.num_patches
        keep = int(L * (1 - mask_ratio))                       # MAE_ViT_Shared.py:64 (float truncation)
        assert 1 <= keep <= L, f"mask_ratio={mask_ratio} keeps {keep} of {L} patches"
        Se, Sd = keep + 1, L + 1
        D, Dd, p = m.dim_model, m.decoder_embed_dim, m.patch_size
        P = p * p * C
        bf16, f32 = torch.bfloat16, torch.float32
        params = dict(m.named_parameters())
        w16 = self._w16_views
        buf = lambda name, shape, dt: self._buf(name, shape, dt, dev)
        imgs_list = [im.contiguous().float() for im in imgs_list]
        self._active_graph = None

        self.generation += 1
        st = dict(N=N, ns=ns, NB=NB, L=L, keep=keep, Se=Se, Sd=Sd, D=D, Dd=Dd, C=C, H=H, p=p, P=P, nsm=nsm,
                  imgs=imgs_list, training=training, generation=self.generation)

        # ---- masking (MAE_ViT_Shared.py:57-84) --------------------------------------------------
        noise = (torch.cat(noises, 0) if ns > 1 else noises[0]).contiguous().float()
        ids_restore = buf("ids_restore", (NB, L), torch.int64)
        ids_shuffle = buf("ids_shuffle", (NB, L), torch.int32)
        mask = buf("mask", (NB, L), f32)
        call("csm_random_masking", noise, NB, L, keep, ids_restore, ids_shuffle, mask)

        # ---- patch embed on the kept patches only + pos embed + cls (Baseline.py:245-256) ------
        patches = buf("patches", (NB * Se, P), bf16)
        for s, im in enumerate(imgs_list):
            call("csm_patch_gather", im, ids_shuffle[s * N:], patches[s * N * Se:], N, C, H, p, L, keep)
        emb = buf("emb", (NB * Se, D), bf16)
        call("csm_linear_fwd", patches, w16["patch_embed.proj.weight"], params["patch_embed.proj.bias"], emb, None,
             NB * Se, D, P, EPI_BF16)
        x = buf("enc.x0", (NB * Se, D), f32)
        call("csm_encoder_assemble", emb, ids_shuffle, params["encoder_pos_embed"], params["cls_token"], x,
             NB, L, keep, D)

        # ---- encoder blocks; encoder_norm is computed-and-discarded upstream: skipped (Baseline.py:264)
        # From here to the decoder's prediction the images are independent: the row chains (see _parts) run
        # concurrently, each on its own stream over its own rows of the same buffers.
        parts = self._parts(NB, dev)
        self._fork(parts)
        x = self._blocks_fwd("enc", "encoder", len(m.encoder), x, parts, Se, D, m.encoder_num_heads, params, w16, dev,
                             NB)
        enc_bf16 = buf("enc_bf16", (NB * Se, D), bf16)

        # ---- decoder (Baseline.py:268-297) -----------------------------------------------------
        demb = buf("demb", (NB * Se, Dd), bf16)
        y = buf("dec.x0", (NB * Sd, Dd), f32)
        for cs, b0, nb in parts:
            e0, e1, d0, d1 = b0 * Se, (b0 + nb) * Se, b0 * Sd, (b0 + nb) * Sd
            with torch.cuda.stream(cs):
                call("csm_cast_f32_bf16", x[e0:e1], enc_bf16[e0:e1], nb * Se * D)
                call("csm_linear_fwd", enc_bf16[e0:e1], w16["decoder_embed.weight"], params["decoder_embed.bias"],
                     demb[e0:e1], None, nb * Se, Dd, D, EPI_BF16)
                call("csm_decoder_assemble", demb[e0:e1], ids_restore[b0:b0 + nb], params["mask_token"],
                     params["decoder_pos_embed"], y[d0:d1], nb, L, keep, Dd)
        y = self._blocks_fwd("dec", "decoder", len(m.decoder), y, parts, Sd, Dd, m.decoder_num_heads, params, w16, dev,
                             NB)
        dec_f32 = buf("dec_f32", (NB * Sd, Dd), f32)
        dec_bf16 = buf("dec_bf16", (NB * Sd, Dd), bf16)
        dn_mean = buf("dn_mean", (NB * Sd,), f32)
        dn_rstd = buf("dn_rstd", (NB * Sd,), f32)
        pred_full = buf("pred_full", (NB * Sd, P), bf16)
        for cs, b0, nb in parts:
            d0, d1 = b0 * Sd, (b0 + nb) * Sd
            with torch.cuda.stream(cs):
                call("csm_layernorm_fwd", y[d0:d1], params["decoder_norm.weight"], params["decoder_norm.bias"],
                     dec_bf16[d0:d1], dec_f32[d0:d1], dn_mean[d0:d1], dn_rstd[d0:d1], nb * Sd, Dd, LN_EPS)
                call("csm_linear_fwd", dec_bf16[d0:d1], w16["decoder_pred.weight"], params["decoder_pred.bias"],
                     pred_full[d0:d1], None, nb * Sd, P, Dd, EPI_BF16)
        self._join(parts)

        # ---- losses ----------------------------------------------------------------------------
        loss_acc = buf("loss_acc", (8,), f32)
        call("csm_zero_async", loss_acc, 32)
        norm_pix = 1 if m.norm_pix_loss else 0
        for s, im in enumerate(imgs_list):
            call("csm_recon_loss_fwd", pred_full[s * N * Sd:], im, mask[s * N:], loss_acc[s:], N, C, H, p, L, norm_pix)
        coefs, red = self._coef_values(ns, N, mask_ratio)
        if ns == 2 and self.use_cd:
            Hp = m.predictor[0].out_features
            h1 = buf("pred.h1", (N * Sd, Hp), bf16)
            call("csm_linear_fwd", dec_bf16[N * Sd:], w16["predictor.0.weight"], params["predictor.0.bias"], h1, None,
                 N * Sd, Hp, Dd, EPI_BF16)
            a1 = buf("pred.a1", (N * Sd, Hp), bf16)
            bn = m.predictor[1]
            bn_mean = buf("pred.bn_mean", (L,), f32)
            bn_rstd = buf("pred.bn_rstd", (L,), f32)
            call("csm_bn_patch_fwd", h1, bn.weight, bn.bias, a1, bn_mean, bn_rstd, bn.running_mean, bn.running_var,
                 N, L, Hp, BN_EPS, BN_MOMENTUM, 1 if training else 0)
            if training:
                bn.num_batches_tracked.add_(1)
            cp = buf("pred.cp", (N * Sd, Dd), bf16)
            call("csm_linear_fwd", a1, w16["predictor.3.weight"], params["predictor.3.bias"], cp, None,
                 N * Sd, Dd, Hp, EPI_BF16)
            call("csm_cross_mse_fwd", cp, dec_f32, loss_acc[2:], N * Sd, Sd, Dd)
            st["Hp"] = Hp
        if ns == 2 and self.use_ce:
            zhat = buf("ntx.zhat", (NB, D), f32)
            fnorm = buf("ntx.fnorm", (NB,), f32)
            neg = buf("ntx.neg", (NB,), f32)
            call("csm_ntxent_fwd", x, zhat, fnorm, neg, loss_acc[3:], N, Se, D, NTXENT_TAU, NTXENT_EPS)
        loss = buf("loss", (1,), f32)
        call("csm_loss_finalize", loss_acc, self._coefs_tensor(coefs, dev), loss, 8)
        loss = loss.view(())
        st["coefs"] = coefs
        st["red"] = red
        self._state = st

        pred = pred_full.view(NB, Sd, P)
        enc = x.view(NB, Se, D)
        dec = dec_f32.view(NB, Sd, Dd)
        out = dict(loss=loss, loss_terms=loss_acc, pred=[pred[s * N:(s + 1) * N, 1:, :] for s in range(ns)],
                   mask=[mask[s * N:(s + 1) * N] for s in range(ns)],
                   ids_restore=[ids_restore[s * N:(s + 1) * N] for s in range(ns)],
                   enc_emb=[enc[s * N:(s + 1) * N] for s in range(ns)],
                   dec_emb=[dec[s * N:(s + 1) * N] for s in range(ns)])
        return out

    def _coef_values(self, ns, N, mask_ratio):
        m = self.model
        L = m.num_patches
        keep = int(L * (1 - mask_ratio))
        n_masked = N * (L - keep)
        red = 0.5 if (ns == 2 and getattr(m, "ms_decoder_loss_reduction", "sum") == "mean") else 1.0
        coefs = [red / n_masked if n_masked > 0 else float("nan")] * ns + [0.0] * (8 - ns)
        if ns == 2 and self.use_cd:
            coefs[2] = 1.0 / (N * L * m.decoder_embed_dim)
        if ns == 2 and self.use_ce:
            coefs[3] = 1.0
        return coefs, red

    def _coefs_tensor(self, coefs, dev):
        """Device copy of the loss coefficients, cached per (values, device): never created inside a stream capture."""
        key = (tuple(coefs), dev)
        t = self._coefs.get(key)
        if t is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("csmae_b200: loss coefficients must be staged before a CUDA-graph capture")
            if len(self._coefs) > 64:
                self._coefs.clear()
            t = self._coefs[key] = torch.tensor(coefs, dtype=torch.float32).to(dev)
        return t

    def _blocks_fwd(self, tag, pname, nlayers, x, parts, S, Dm, heads, params, w16, dev, NB):
        bf16, f32 = torch.bfloat16, torch.float32
        rows = NB * S
        d = Dm // heads
        buf = lambda name, shape, dt: self._buf(name, shape, dt, dev)
        for i in range(nlayers):
            t, q = f"{tag}.{i}.", f"{pname}.{i}."
            hid = params[q + "mlp.fc1.weight"].shape[0]
            ln1 = buf(t + "ln1", (rows, Dm), bf16)
            mean1, rstd1 = buf(t + "mean1", (rows,), f32), buf(t + "rstd1", (rows,), f32)
            qkv = buf(t + "qkv", (rows, 3 * Dm), bf16)
            ao = buf(t + "ao", (rows, Dm), bf16)
            lse = buf(t + "lse", (NB * heads * S,), f32)
            xmid = buf(t + "xmid", (rows, Dm), f32)
            ln2 = buf(t + "ln2", (rows, Dm), bf16)
            mean2, rstd2 = buf(t + "mean2", (rows,), f32), buf(t + "rstd2", (rows,), f32)
            # fc1 + GELU in one epilogue; what is kept for the backward is gelu'(h), not h
            gp = buf(t + "gp", (rows, hid), bf16)
            act = buf(t + "act", (rows, hid), bf16)
            xout = buf(t + "xout", (rows, Dm), f32)
            for cs, b0, nb in parts:
                r0, r1, n = b0 * S, (b0 + nb) * S, nb * S
                with torch.cuda.stream(cs):
                    call("csm_layernorm_fwd", x[r0:r1], params[q + "norm1.weight"], params[q + "norm1.bias"],
                         ln1[r0:r1], None, mean1[r0:r1], rstd1[r0:r1], n, Dm, LN_EPS)
                    call("csm_linear_fwd", ln1[r0:r1], w16[q + "attn.qkv.weight"], params[q + "attn.qkv.bias"],
                         qkv[r0:r1], None, n, 3 * Dm, Dm, EPI_BF16)
                    call("csm_attention_fwd", qkv[r0:r1], ao[r0:r1], lse[b0 * heads * S:], nb, S, heads, d)
                    call("csm_linear_fwd", ao[r0:r1], w16[q + "attn.proj.weight"], params[q + "attn.proj.bias"],
                         xmid[r0:r1], x[r0:r1], n, Dm, Dm, EPI_RESID)
                    call("csm_layernorm_fwd", xmid[r0:r1], params[q + "norm2.weight"], params[q + "norm2.bias"],
                         ln2[r0:r1], None, mean2[r0:r1], rstd2[r0:r1], n, Dm, LN_EPS)
                    call("csm_linear_fwd", ln2[r0:r1], w16[q + "mlp.fc1.weight"], params[q + "mlp.fc1.bias"],
                         gp[r0:r1], act[r0:r1], n, hid, Dm, EPI_GELU)
                    call("csm_linear_fwd", act[r0:r1], w16[q + "mlp.fc2.weight"], params[q + "mlp.fc2.bias"],
                         xout[r0:r1], xmid[r0:r1], n, Dm, hid, EPI_RESID)
            x = xout
        return x

    # ------------------------------------------------------------------ backward
    def backward(self, grad_loss, generation):
        if not grad_loss.is_cuda:                # host-logic tests stub the kernel chain on CPU tensors
            return self._backward(grad_loss, generation)
        with torch.cuda.device(grad_loss.device):
            return self._backward(grad_loss, generation)

    def _backward(self, grad_loss, generation):
        st = self._state
        if st is None or st["generation"] != generation or generation != self.generation:
            raise RuntimeError(
                "csmae_b200: backward() of a forward whose activations were overwritten by a later forward of the "
                "same module (the workspace holds one step); call backward before the next forward")
        entry = getattr(self, "_active_graph", None)
        sync = self._sync_enabled
        if sync:
            grad_loss = grad_loss / self._sync_world          # gradients are averaged over the replicas
        if entry is None:
            pending = []
            segs_box = {}

            def boundary(k):
                if sync:
                    self._issue_allreduce(segs_box["flat"], segs_box["segs"][k], pending)
            flat, views = self._backward_eager(grad_loss, None, boundary if sync else None, segs_box)
            for w in pending:
                w.wait()
            self._state = None
            return views
        # graphed step: the gradient chain is captured once over static buffers (in several graphs when the
        # gradient all-reduce is overlapped: one per segment)
        if entry["bwd"] is None:
            entry["g"] = torch.zeros(1, dtype=torch.float32, device=grad_loss.device)
            entry["g"].copy_(grad_loss.detach().reshape(1))
            names = self.param_names()
            params = dict(self.model.named_parameters())
            sizes = [params[n].numel() for n in names]
            total = sum((s_ + 3) // 4 * 4 for s_ in sizes)
            entry["flat"] = torch.zeros(total, dtype=torch.float32, device=grad_loss.device)
            entry["dense"] = all(s_ % 4 == 0 for s_ in sizes)
            entry["like"] = [params[n] for n in names]
            torch.cuda.synchronize()
            seq = _GraphSequence()
            segs_box = {}
            c0 = nat.launch_count
            seq.begin()
            try:
                cut = (lambda k: seq.cut() if k < len(segs_box["segs"]) - 1 else None) if sync else None
                self._backward_eager(entry["g"], entry["flat"], cut, segs_box)
            finally:
                seq.end()
            entry["bwd_launches"] = nat.launch_count - c0
            nat.launch_count = c0
            entry["bwd"] = seq.graphs
            entry["segs"] = segs_box.get("segs", [(0, total)])
            self._state = st                   # capture does not run the kernels; replay below does
        entry["g"].copy_(grad_loss.detach().reshape(1), non_blocking=True)
        flat = entry["flat"]

        def views_of(buf):
            if entry["dense"]:
                return list(torch._utils._unflatten_dense_tensors(buf, entry["like"]))
            out_views, off = [], 0
            for p in entry["like"]:
                out_views.append(buf[off:off + p.numel()].view(p.shape))
                off += (p.numel() + 3) // 4 * 4
            return out_views

        # The views handed to autograd alias the graph's gradient buffer (no 4 B/param copy per step); autograd
        # adopts them as .grad when .grad is None.  If a .grad still aliases the buffer when the next backward
        # starts (gradient accumulation without zero_grad), the accumulated values are moved out first.
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        if any(p.grad is not None and lo <= p.grad.data_ptr() < hi for p in entry["like"]):
            saved = flat.clone()
            for p, v, a in zip(entry["like"], views_of(saved), views_of(flat)):
                if p.grad is not None and p.grad.data_ptr() == a.data_ptr():
                    p.grad = v
        pending = []
        for k, g in enumerate(entry["bwd"]):
            g.replay()
            if sync and k < len(entry["segs"]):
                self._issue_allreduce(flat, entry["segs"][k], pending)
        for w in pending:
            w.wait()
        nat.launch_count += entry["bwd_launches"]
        self._state = None
        return views_of(flat)

    def _backward_eager(self, grad_loss, flat, boundary=None, segs_box=None):
        """The hand-written backward chain.  boundary(k), when given, is called right after the kernels that
        complete gradient segment k (see grad_segments) have been issued -- the caller all-reduces that slice
        (eager) or cuts the CUDA-graph capture there."""
        st = self._state
        m = self.model
        N, ns, NB, L, keep, Se, Sd = (st[k] for k in ("N", "ns", "NB", "L", "keep", "Se", "Sd"))
        D, Dd, C, H, p, P, nsm = (st[k] for k in ("D", "Dd", "C", "H", "p", "P", "nsm"))
        dev = st["imgs"][0].device
        if self.stream_k_mode == "fwd":
            nat.enable_gemm_stream_k(dev, False)      # the backward's partial rounds are filled by the side stream
        bf16, f32 = torch.bfloat16, torch.float32
        params = dict(m.named_parameters())
        w16 = self._w16_views
        B = self._bufs
        buf = lambda name, shape, dt: self._buf(name, shape, dt, dev)
        names = self.param_names()
        sizes = [params[n].numel() for n in names]
        offs, total = [], 0
        for s_ in sizes:
            offs.append(total)
            total += (s_ + 3) // 4 * 4                         # keep every gradient 16-byte aligned
        if flat is None:
            flat = torch.empty(total, dtype=f32, device=dev)
        # The weight-gradient GEMMs reduce-add their split-K partials and the bias / LayerNorm / token gradients are
        # atomic column sums, so the buffer starts from zero: a memset on the side stream, hidden under the first
        # kernels of the chain (loss gradients, decoder_pred dgrad), which do not touch it.
        main = torch.cuda.current_stream(dev) if flat.is_cuda else None
        zeroed = None
        if flat.is_cuda:
            if self.use_side_stream:
                side = self._side_stream(dev)
                self._wait(side, main)
                with torch.cuda.stream(side):
                    call("csm_zero_async", flat, total * 4)
                zeroed = torch.cuda.Event()
                zeroed.record(side)
                self._side_dirty = True
            else:
                call("csm_zero_async", flat, total * 4)
        else:
            flat.zero_()
        G = {n: flat[o:o + s_].view(params[n].shape) for n, o, s_ in zip(names, offs, sizes)}
        segs, seg_layers = grad_segments(names, offs, total, len(m.encoder), self._sync_groups)
        if segs_box is not None:
            segs_box["flat"], segs_box["segs"] = flat, segs
        if boundary is None or len(segs) == 1:
            seg_layers, fire = [], (lambda k: None)
        else:
            def fire(k):
                self._join_side(dev)             # the segment's weight gradients must have landed
                boundary(k)
        g = grad_loss.detach().reshape(1).to(f32).contiguous()
        norm_pix = 1 if m.norm_pix_loss else 0
        rows_d, rows_e = NB * Sd, NB * Se

        # ---- reconstruction loss -> decoder_pred ------------------------------------------------
        dpred = buf("b.dpred", (rows_d, P), bf16)
        coef = st["coefs"][0] / P
        for s, im in enumerate(st["imgs"]):
            call("csm_recon_loss_bwd", B["pred_full"][s * N * Sd:], im, B["mask"][s * N:], dpred[s * N * Sd:], g, coef,
                 N, C, H, p, L, norm_pix)
        def side_pred():
            call("csm_linear_wgrad", dpred, B["dec_bf16"], G["decoder_pred.weight"], rows_d, P, Dd, nsm)
            call("csm_colsum_bf16", dpred, G["decoder_pred.bias"], rows_d, P, 0, nsm)
        self._on_side(dev, side_pred)
        d_dec = buf("b.dec.dln", (rows_d, Dd), bf16)
        call("csm_linear_dgrad", dpred, w16["decoder_pred.weight"], d_dec, None, rows_d, P, Dd, EPI_BF16)
        if zeroed is not None:
            main.wait_event(zeroed)              # from here on the chain itself adds into the gradient buffer

        # ---- cross-scale decoder loss through the predictor (MsLdCeCd.py:57-59) ------------------
        dy2 = None
        if ns == 2 and self.use_cd:
            Hp = st["Hp"]
            bn = m.predictor[1]
            dy2 = buf("b.dy2", (rows_d, Dd), f32)
            d_cp = buf("b.d_cp", (N * Sd, Dd), bf16)
            call("csm_cross_mse_bwd", B["pred.cp"], B["dec_f32"], d_cp, dy2, g, st["coefs"][2], N * Sd, Sd, Dd)
            def side_p3():
                call("csm_linear_wgrad", d_cp, B["pred.a1"], G["predictor.3.weight"], N * Sd, Dd, Hp, nsm)
                call("csm_colsum_bf16", d_cp, G["predictor.3.bias"], N * Sd, Dd, 0, nsm)
            self._on_side(dev, side_p3)
            d_a1 = buf("b.d_a1", (N * Sd, Hp), bf16)
            call("csm_linear_dgrad", d_cp, w16["predictor.3.weight"], d_a1, None, N * Sd, Dd, Hp, EPI_BF16)
            dh1 = buf("b.dh1", (N * Sd, Hp), bf16)
            call("csm_bn_patch_bwd", B["pred.h1"], B["pred.a1"], d_a1, bn.weight, B["pred.bn_mean"], B["pred.bn_rstd"],
                 dh1, G["predictor.1.weight"], G["predictor.1.bias"], N, L, Hp, 1 if st["training"] else 0)
            def side_p0():
                call("csm_linear_wgrad", dh1, B["dec_bf16"][N * Sd:], G["predictor.0.weight"], N * Sd, Hp, Dd, nsm)
                call("csm_colsum_bf16", dh1, G["predictor.0.bias"], N * Sd, Hp, 0, nsm)
            self._on_side(dev, side_p0)
            call("csm_linear_dgrad", dh1, w16["predictor.0.weight"], dy2[N * Sd:], None, N * Sd, Hp, Dd, EPI_F32)

        # ---- decoder_norm, decoder blocks (row chains, as in the forward) ---------------------------
        parts = self._parts(NB, dev)
        chain_streams = [cs for cs, _, _ in parts[1:]]
        y_final = B[f"dec.{len(m.decoder) - 1}.xout"] if len(m.decoder) else B["dec.x0"]
        dres = buf("b.dec.dres", (rows_d, Dd), f32)
        dres16 = buf("b.dec.dres16", (rows_d, Dd), bf16)
        # the bf16 residual-stream gradient every LayerNorm backward emits is the dY of the Linear feeding
        # that residual add, so its column sums (that Linear's bias gradient) are taken in the same pass
        top_fc2_bias = G[f"decoder.{len(m.decoder) - 1}.mlp.fc2.bias"] if len(m.decoder) else None
        self._fork(parts)
        for cs, b0, nb in parts:
            d0, d1 = b0 * Sd, (b0 + nb) * Sd
            with torch.cuda.stream(cs):
                call("csm_layernorm_bwd", d_dec[d0:d1], None if dy2 is None else dy2[d0:d1], y_final[d0:d1],
                     B["dn_mean"][d0:d1], B["dn_rstd"][d0:d1], params["decoder_norm.weight"], None, dres[d0:d1],
                     dres16[d0:d1], G["decoder_norm.weight"], G["decoder_norm.bias"], top_fc2_bias, nb * Sd, Dd, nsm)
        self._blocks_bwd("dec", "decoder", len(m.decoder), dres, dres16, parts, NB, Sd, Dd, m.decoder_num_heads, params,
                         w16, G, dev, nsm, top_bias_done=True)
        if len(segs) > 1:
            self._join(parts)
            fire(0)                                  # decoder.*, decoder_pred, decoder_norm, predictor are final
            self._fork(parts)

        # ---- un-shuffle backward, decoder_embed ---------------------------------------------------
        d_demb = buf("b.d_demb", (rows_e, Dd), bf16)
        d_enc = buf("b.enc.dln", (rows_e, D), bf16)
        # ---- NT-Xent feature gradient joins the encoder output gradient ---------------------------
        d_feat = None
        if ns == 2 and self.use_ce:
            d_feat = buf("b.d_feat", (NB, D), f32)
            call("csm_ntxent_bwd", B["ntx.zhat"], B["ntx.fnorm"], B["ntx.neg"], g, d_feat, N, D, NTXENT_TAU, NTXENT_EPS)
            self._fork(parts)
        eres = buf("b.enc.dres", (rows_e, D), f32)
        eres16 = buf("b.enc.dres16", (rows_e, D), bf16)
        for cs, b0, nb in parts:
            e0, e1, d0, d1 = b0 * Se, (b0 + nb) * Se, b0 * Sd, (b0 + nb) * Sd
            with torch.cuda.stream(cs):
                call("csm_decoder_assemble_bwd", dres[d0:d1], B["ids_shuffle"][b0:b0 + nb], d_demb[e0:e1],
                     G["mask_token"], nb, L, keep, Dd)
                call("csm_linear_dgrad", d_demb[e0:e1], w16["decoder_embed.weight"], d_enc[e0:e1], None, nb * Se, Dd, D,
                     EPI_BF16)
                call("csm_encoder_out_grad", d_enc[e0:e1], None if d_feat is None else d_feat[b0:b0 + nb], eres[e0:e1],
                     eres16[e0:e1], nb, Se, D)

        def side_demb():
            call("csm_linear_wgrad", d_demb, B["enc_bf16"], G["decoder_embed.weight"], rows_e, Dd, D, nsm)
            call("csm_colsum_bf16", d_demb, G["decoder_embed.bias"], rows_e, Dd, 0, nsm)
        self._on_side(dev, side_demb, chain_streams)

        def enc_boundary(i):
            if i in seg_layers:
                self._join(parts)
                fire(1 + seg_layers.index(i))
                self._fork(parts)
        eres16 = self._blocks_bwd("enc", "encoder", len(m.encoder), eres, eres16, parts, NB, Se, D, m.encoder_num_heads,
                                  params, w16, G, dev, nsm, top_bias_done=False, after_layer=enc_boundary)
        self._join(parts)

        # ---- cls token, patch embed (only the kept patches carry gradient; cls-slot rows are zero) -
        call("csm_cls_grad", eres, G["cls_token"], NB, Se, D)
        call("csm_linear_wgrad", eres16, B["patches"], G["patch_embed.proj.weight"], rows_e, D, P, nsm)
        call("csm_colsum_bf16", eres16, G["patch_embed.proj.bias"], rows_e, D, Se, nsm)
        self._join_side(dev)
        if len(segs) > 1:
            fire(len(segs) - 1)
        elif boundary is not None:
            boundary(0)
        return flat, [G[n] for n in names]

    def _blocks_bwd(self, tag, pname, nlayers, dres, dres16, parts, NB, S, Dm, heads, params, w16, G, dev, nsm,
                    top_bias_done, after_layer=None):
        """Backward of a stack of Blocks.  The critical chain (dgrad -> LayerNorm backward -> attention backward) runs
        per row chain on the chains' streams; the weight-gradient GEMMs and bias column sums, which reduce over ALL
        rows, are issued once per Linear on the side stream after every chain has produced its rows of dY."""
        bf16, f32 = torch.bfloat16, torch.float32
        rows = NB * S
        d = Dm // heads
        B = self._bufs
        buf = lambda name, shape, dt: self._buf(name, shape, dt, dev)
        chain_streams = [cs for cs, _, _ in parts[1:]]

        def per_chain(fn):
            for cs, b0, nb in parts:
                with torch.cuda.stream(cs):
                    fn(b0, nb, b0 * S, (b0 + nb) * S, nb * S)

        dln = buf(f"b.{tag}.dln", (rows, Dm), bf16)
        d_ao = buf(f"b.{tag}.d_ao", (rows, Dm), bf16)
        delta = buf(f"b.{tag}.delta", (NB * heads * S,), f32)
        for i in reversed(range(nlayers)):
            t, q = f"{tag}.{i}.", f"{pname}.{i}."
            hid = params[q + "mlp.fc1.weight"].shape[0]
            x_in = B[f"{tag}.{i - 1}.xout"] if i > 0 else B[f"{tag}.x0"]
            # gradient buffers the side-stream wgrads read are per layer (no reuse hazards); dres16 enters as the
            # previous LayerNorm backward's bf16 output
            dh = buf(f"b.{tag}.{i}.dh", (rows, hid), bf16)
            dqkv = buf(f"b.{tag}.{i}.dqkv", (rows, 3 * Dm), bf16)
            d16a = buf(f"b.{tag}.{i}.d16a", (rows, Dm), bf16)
            d16b = buf(f"b.{tag}.{i}.d16b", (rows, Dm), bf16)
            # MLP branch: x_out = x_mid + fc2(gelu(fc1(norm2(x_mid))))
            first_bias = (i == nlayers - 1 and not top_bias_done)

            def side_fc2(dres16=dres16, t=t, q=q, hid=hid, first_bias=first_bias):
                call("csm_linear_wgrad", dres16, B[t + "act"], G[q + "mlp.fc2.weight"], rows, Dm, hid, nsm)
                if first_bias:
                    call("csm_colsum_bf16", dres16, G[q + "mlp.fc2.bias"], rows, Dm, 0, nsm)
            self._on_side(dev, side_fc2, chain_streams)
            per_chain(lambda b0, nb, r0, r1, n: call(
                "csm_linear_dgrad", dres16[r0:r1], w16[q + "mlp.fc2.weight"], dh[r0:r1], B[t + "gp"][r0:r1], n, Dm, hid,
                EPI_DGELU))

            def side_fc1(dh=dh, t=t, q=q, hid=hid):
                call("csm_linear_wgrad", dh, B[t + "ln2"], G[q + "mlp.fc1.weight"], rows, hid, Dm, nsm)
                call("csm_colsum_bf16", dh, G[q + "mlp.fc1.bias"], rows, hid, 0, nsm)
            self._on_side(dev, side_fc1, chain_streams)

            def mlp_tail(b0, nb, r0, r1, n):
                call("csm_linear_dgrad", dh[r0:r1], w16[q + "mlp.fc1.weight"], dln[r0:r1], None, n, hid, Dm, EPI_BF16)
                call("csm_layernorm_bwd", dln[r0:r1], None, B[t + "xmid"][r0:r1], B[t + "mean2"][r0:r1],
                     B[t + "rstd2"][r0:r1], params[q + "norm2.weight"], dres[r0:r1], dres[r0:r1], d16a[r0:r1],
                     G[q + "norm2.weight"], G[q + "norm2.bias"], G[q + "attn.proj.bias"], n, Dm, nsm)
            per_chain(mlp_tail)
            # attention branch: x_mid = x_in + proj(attn(qkv(norm1(x_in))))

            def side_proj(d16a=d16a, t=t, q=q):
                call("csm_linear_wgrad", d16a, B[t + "ao"], G[q + "attn.proj.weight"], rows, Dm, Dm, nsm)
            self._on_side(dev, side_proj, chain_streams)

            def attn_branch(b0, nb, r0, r1, n):
                call("csm_linear_dgrad", d16a[r0:r1], w16[q + "attn.proj.weight"], d_ao[r0:r1], None, n, Dm, Dm,
                     EPI_BF16)
                # (the kernel can also emit the qkv.bias column sums itself -- measured slower than the separate
                #  pass on B200: the extra tail per CTA is not hidden at one CTA per SM)
                call("csm_attention_bwd", B[t + "qkv"][r0:r1], B[t + "ao"][r0:r1], d_ao[r0:r1],
                     B[t + "lse"][b0 * heads * S:], delta[b0 * heads * S:], dqkv[r0:r1], None, nb, S, heads, d)
            per_chain(attn_branch)

            def side_qkv(dqkv=dqkv, t=t, q=q):
                call("csm_linear_wgrad", dqkv, B[t + "ln1"], G[q + "attn.qkv.weight"], rows, 3 * Dm, Dm, nsm)
                call("csm_colsum_bf16", dqkv, G[q + "attn.qkv.bias"], rows, 3 * Dm, 0, nsm)
            self._on_side(dev, side_qkv, chain_streams)
            below_fc2_bias = G[f"{pname}.{i - 1}.mlp.fc2.bias"] if i > 0 else None

            def attn_tail(b0, nb, r0, r1, n):
                call("csm_linear_dgrad", dqkv[r0:r1], w16[q + "attn.qkv.weight"], dln[r0:r1], None, n, 3 * Dm, Dm,
                     EPI_BF16)
                call("csm_layernorm_bwd", dln[r0:r1], None, x_in[r0:r1], B[t + "mean1"][r0:r1], B[t + "rstd1"][r0:r1],
                     params[q + "norm1.weight"], dres[r0:r1], dres[r0:r1], d16b[r0:r1], G[q + "norm1.weight"],
                     G[q + "norm1.bias"], below_fc2_bias, n, Dm, nsm)
            per_chain(attn_tail)
            dres16 = d16b
            if after_layer is not None:
                after_layer(i)
        return dres16


class CrossScaleStep(torch.autograd.Function):
    """Autograd node of one step.  The forward kernels are ALREADY enqueued when this node is built (model._run
    calls engine.forward first): wiring 258 parameter edges costs ~0.4 ms of Python, which now overlaps the GPU
    instead of delaying the first kernel.  Its backward is the hand-written chain; the parameters are explicit
    inputs so that autograd (and DDP's unused-parameter walk) sees exactly the ones used."""

    @staticmethod
    def forward(ctx, engine, loss, generation, *params):
        ctx.engine = engine
        ctx.generation = generation
        ctx.n_params = len(params)
        return loss.clone()                 # a fresh scalar per step: the accumulator it came from is reused

    @staticmethod
    def backward(ctx, grad_loss):
        grads = ctx.engine.backward(grad_loss, ctx.generation)
        assert len(grads) == ctx.n_params
        return (None, None, None, *grads)
