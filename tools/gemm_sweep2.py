"""GPU: folded / plain forward GEMM time per signal vs working-set size (is the folded kernel DRAM-latency bound?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import avsep_b200 as av
import bench

L = 65535
lib, chk = av._lib.lib, av._lib.check
stream = torch.cuda.current_stream().cuda_stream
for S in (37, 74, 111, 148, 240):
    w = (torch.rand((S, L), device="cuda") - 0.5)
    plan = av.Plan(S, L)
    mag = torch.empty((S, 512, 256), device="cuda")
    for fold, two in ((0, 1), (1, 0), (1, 1)):
        plan.set_fold(fold); plan.set_two_cta(two)
        av.stft(w, phase=False, plan=plan)
        t = bench.time_calls(lambda: chk(lib.avsep_debug_gemm_only(plan.handle, 0, S, mag.data_ptr(), stream)), 20, warmup=3)
        print(f"S={S:3d} fold={fold} two_cta={two}: {t*1e3:7.1f} us  = {t*1e3/S:6.3f} us/signal", flush=True)
    plan.close()
