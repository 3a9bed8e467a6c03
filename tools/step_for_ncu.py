"""One un-graphed training step of the default bench workload between cudaProfilerStart/Stop, for ncu captures of the
kernels as the step launches them (shapes, launch order and counts of bench.py's default line):

    CSMAE_CUDA_GRAPHS=0 ncu --profile-from-start off --set full --clock-control none --import-source on \
        -k regex:gemm_kernel -o gpurun_out/r2_gemm_step python tools/step_for_ncu.py [base|large] [batch] [size]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cross-scale-mae_b200"))
GRAPHS = os.environ.get("CSM_NCU_GRAPHS", "0") == "1"      # 1: profile one step replayed from CUDA graphs instead
os.environ["CSMAE_CUDA_GRAPHS"] = "1" if GRAPHS else "0"
import torch  # noqa: E402
import csmae_b200  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "base"
B = int(sys.argv[2]) if len(sys.argv) > 2 else (64 if arch == "base" else 32)
S = int(sys.argv[3]) if len(sys.argv) > 3 else 224
torch.manual_seed(0)
ctor = csmae_b200.mae_vit_base_patch16 if arch == "base" else csmae_b200.mae_vit_large_patch16
model = ctor(input_size=S, device="cuda").cuda().train()
opt = csmae_b200.FusedAdamW([p for p in model.parameters() if p.requires_grad], lr=1.5e-4, betas=(0.9, 0.95), model=model)
g = torch.Generator(device="cuda").manual_seed(1000)
x1 = torch.randn(B, 3, S, S, device="cuda", generator=g)
x2 = torch.randn(B, 3, S, S, device="cuda", generator=g)


def step():
    opt.zero_grad(set_to_none=True)
    loss, _, _ = model(x1, x2, 0.75)
    loss.backward()
    opt.step()
    return loss


for _ in range(8 if GRAPHS else 3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
