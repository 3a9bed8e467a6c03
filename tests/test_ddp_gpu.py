"""N-rank gradient equivalence of the data-parallel path on real GPUs (reference semantics:
main_pretrain.py:417-421 -- DistributedDataParallel averages the replicas' gradients).  Two processes, one per GPU,
NCCL; skipped on a single-GPU box.  Each rank computes its local gradients without any wrapper, then the same step
under csmae_b200.DistributedDataParallel (engine-overlapped segment all-reduce, eager and CUDA-graph replay) and
under torch's own wrapper; both must equal the mean of the two ranks' local gradients."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        for p_ in (ROOT, os.path.join(ROOT, "cross-scale-mae_b200")):
            if p_ not in sys.path:
                sys.path.insert(0, p_)
        import csmae_b200
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        cfg = dict(dim_model=128, encoder_num_layers=4, encoder_num_heads=2, decoder_embed_dim=64, decoder_num_layers=2,
                   decoder_num_heads=2, input_size=96, patch_size=16, predictor_hidden_size=128)
        torch.manual_seed(0)
        model = csmae_b200.MAE_ViT_MsLdCeCd(**cfg, device=str(dev)).to(dev).train()
        g = torch.Generator(device=dev).manual_seed(100 + rank)          # every rank has its own data and masks
        x1, x2 = torch.randn(8, 3, 96, 96, device=dev, generator=g), torch.randn(8, 3, 96, 96, device=dev, generator=g)
        n1, n2 = torch.rand(8, 36, device=dev, generator=g), torch.rand(8, 36, device=dev, generator=g)

        def grads(m):
            for p in model.parameters():
                p.grad = None
            loss, _, _ = m(x1, x2, 0.75, noise=[n1, n2])
            loss.backward()
            return {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

        model._engine.use_graphs = False
        local = grads(model)
        want = {}
        for n, gl in local.items():
            t = gl.clone()
            dist.all_reduce(t)
            want[n] = t / world

        def check(got, what):
            assert got.keys() == want.keys(), what
            worst = 0.0
            for n in want:
                err = ((got[n] - want[n]).norm() / (want[n].norm() + 1e-30)).item()
                worst = max(worst, err)
                assert err < 1e-4, f"{what}: {n} rel-L2 {err:.3e}"      # fp32 sums in another order
            return worst

        model._engine.use_graphs = True
        ddp = csmae_b200.DistributedDataParallel(model, device_ids=[rank], find_unused_parameters=True)
        assert model._engine._sync_groups == 4
        worst = [check(grads(ddp), f"native step {i}") for i in range(5)]      # 2 eager warm-ups, capture, replays
        assert len(model._engine._graphs) == 1, "the graphed (segment-cut) backward was not exercised"
        # running statistics follow rank 0 through the single coalesced buffer broadcast: each rank updated them with
        # its own batch statistics during the steps above, the next forward starts by re-broadcasting rank 0's
        bn = model.predictor[1]
        mine = bn.running_mean.detach().clone()
        allm = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allm, mine)
        assert not torch.equal(allm[0], allm[1]), "ranks see different data: their running means should differ"
        flat = ddp._coalesced_buffers()
        assert flat is not None and bn.running_mean.data_ptr() >= flat.data_ptr()
        dist.broadcast(flat, 0)
        dist.all_gather(allm, bn.running_mean.detach().clone())
        assert torch.equal(allm[0], allm[1]) and torch.equal(allm[0], allm[rank])
        assert int(bn.num_batches_tracked.item()) == 6       # the local step above + the five wrapped steps
        # torch's own wrapper, exactly as main_pretrain.py:417-421 builds it
        model._engine.enable_grad_sync(None, 1)
        tddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank], find_unused_parameters=True)
        worst.append(check(grads(tddp), "torch DDP"))
        q.put((rank, True, max(worst)))
        dist.destroy_process_group()
    except Exception as e:      # noqa: BLE001
        import traceback
        q.put((rank, False, traceback.format_exc() + repr(e)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradient_equivalence():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in results:
        assert ok, f"rank {rank}: {info}"
    print("two-rank gradient equivalence, worst rel-L2:", max(r[2] for r in results))
