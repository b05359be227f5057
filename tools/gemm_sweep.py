"""GPU: time the DFT GEMM kernel alone (forward and inverse epilogues) for each K-block / stage-count variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avsep_b200 as av
import bench

S, L = 240, 65535
lib, chk = av._lib.lib, av._lib.check
w = (torch.rand((S, L), device="cuda") - 0.5)
plan = av.Plan(S, L)
plan.set_fused_stage(False)                 # this sweep times the GEMM kernels on STAGED operands (the fused kernels have none)
for fold in (0, 1):                         # stage both operand layouts once
    plan.set_fold(fold)
    mag, ph = av.stft(w, plan=plan)
wav = av.istft(mag, ph, plan=plan)          # stages the inverse operand
torch.cuda.synchronize()
stream = torch.cuda.current_stream().cuda_stream
ref_mag, ref_wav = mag.clone(), wav.clone()
for kb, fold, two in ((32, 0, 0), (64, 0, 0), (32, 0, 1), (64, 0, 1), (64, 1, 0), (64, 1, 1)):
    plan.set_kblock(kb)
    plan.set_fold(fold)
    plan.set_two_cta(two)
    t_f = bench.time_calls(lambda: chk(lib.avsep_debug_gemm_only(plan.handle, 0, S, mag.data_ptr(), stream)), 20, warmup=3)
    t_i = bench.time_calls(lambda: chk(lib.avsep_debug_gemm_only(plan.handle, 1, S, wav.data_ptr(), stream)), 20, warmup=3)
    ok = float((mag - ref_mag).abs().max() / ref_mag.max()) < 2e-5, float((wav - ref_wav).abs().max())
    tf = 3 * 2 * 1024 * 1024 * 256 * S / 1e12
    print(f"kblock {kb:2d} fold={int(fold)} two_cta={two}: fwd {t_f*1e3:7.1f} us ({tf/t_f*1e3:7.1f} TFLOP/s executed)  inv {t_i*1e3:7.1f} us ({tf/t_i*1e3:7.1f})  same-bits-fwd={ok[0]} inv-maxdiff={ok[1]:.2e}")
