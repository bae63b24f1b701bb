"""Measures the accumulation bias of tcgen05.mma (fp32 accumulator in TMEM) against an exact fp64 product of the same
bf16 inputs: mean signed error in units of ulp(result) as a function of K (number of accumulating MMAs = K/16)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neuraloc_b200 as nb

L = nb._cabi.lib()
torch.manual_seed(0)
for kind in ("mixed", "positive"):
    for K in (16, 64, 128, 256, 512):
        N = 128
        A = torch.randn(128, K, device="cuda")
        B = torch.randn(N, K, device="cuda") * 0.05
        if kind == "positive":
            A = A.abs() + 0.7
        Ab, Bb = A.bfloat16().contiguous(), B.bfloat16().contiguous()
        D = torch.zeros(128, N, device="cuda")
        rc = L.noc_tc_probe(Ab.data_ptr(), Bb.data_ptr(), D.data_ptr(), N, K, 0, 0, None)
        assert rc == 0
        torch.cuda.synchronize()
        ref = Ab.double() @ Bb.double().t()
        ulp = torch.pow(2.0, torch.floor(torch.log2(ref.abs().clamp_min(1e-30))) - 23)
        err = (D.double() - ref) / ulp
        serr = err * torch.sign(ref)              # negative = magnitude shrinks (round toward zero)
        Dc = (Ab.float() @ Bb.float().t())        # cuBLAS fp32 on the same values (FMA or TF32-free path)
        cerr = ((Dc.double() - ref) / ulp) * torch.sign(ref)
        big = ref.abs() > ref.abs().median()
        print("%-8s K=%4d (%2d MMAs): tcgen05 signed err mean %+7.3f ulp  rms %6.3f  | fp32 matmul mean %+6.3f rms %5.3f"
              % (kind, K, K // 16, serr[big].mean().item(), err[big].pow(2).mean().sqrt().item(), cerr[big].mean().item(),
                 ((Dc.double() - ref) / ulp)[big].pow(2).mean().sqrt().item()), flush=True)
