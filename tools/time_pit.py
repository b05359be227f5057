"""Time the fused SoP++ audio-only forward (masks + per-sample N x N PIT loss matrix in one pass over feat_sound) and the
stand-alone synthesizer backward (what PitWrapper / plain autograd run after it) at the C2 shape."""
import os
import sys

import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import avsep_b200 as av
import bench
B,N,K=80,2,32
g=torch.Generator(device='cuda').manual_seed(1)
feat=torch.randn((B,K,256,256),device='cuda',generator=g)
vecs=torch.sigmoid(torch.randn((B,N,K),device='cuda',generator=g))
gt=(torch.rand((N,B,256,256),device='cuda',generator=g)>0.5).float()
w=torch.rand((B,256,256),device='cuda',generator=g)+0.1
scale=torch.ones(K,device='cuda'); bias=torch.zeros(1,device='cuda')
lib,chk=av._lib.lib,av._lib.check
out=torch.empty((N,B,256*256),device='cuda'); lm=torch.empty((B,N,N),device='cuda')
scratch=torch.empty(int(lib.avsep_synth_scratch_bytes(B,N,K,65536)),dtype=torch.uint8,device='cuda')
st=torch.cuda.current_stream().cuda_stream
f=lambda: chk(lib.avsep_synth_forward_pit(feat.data_ptr(),vecs.data_ptr(),scale.data_ptr(),bias.data_ptr(),B,N,K,65536,1,out.data_ptr(),3,gt.data_ptr(),w.data_ptr(),lm.data_ptr(),scratch.data_ptr(),st))
ms=bench.time_calls(f,10)
nbytes=B*(32*65536*4+N*65536*4+(N+1)*65536*4)
print('forward_pit ms',ms,'GB/s',nbytes/ms/1e6)
one = torch.ones(1, device='cuda')
dfeat = torch.empty_like(feat); dvecs = torch.empty_like(vecs); dscale = torch.empty(K, device='cuda'); dbias = torch.empty(1, device='cuda')
b = lambda: chk(lib.avsep_synth_backward(feat.data_ptr(), vecs.data_ptr(), scale.data_ptr(), B, N, K, 65536, 1, 3, out.data_ptr(), gt.data_ptr(),
                                         w.data_ptr(), one.data_ptr(), None, dfeat.data_ptr(), dvecs.data_ptr(), dscale.data_ptr(), dbias.data_ptr(),
                                         scratch.data_ptr(), st))
ms = bench.time_calls(b, 10)
nbytes = B * (32 * 65536 * 4 * 2 + (2 * N + 1) * 65536 * 4)
print('backward ms', ms, 'GB/s', nbytes / ms / 1e6)
