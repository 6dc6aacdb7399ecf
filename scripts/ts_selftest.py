import sys; sys.path.insert(0,"/root/repo")
import torch, numpy as np
import pgmp_b200._native as nv
g=torch.Generator().manual_seed(0)
A=torch.randn(128,64,generator=g); W=torch.randn(64,64,generator=g)
Ad,Wd=A.cuda(),W.cuda(); D=torch.full((128,64),float("nan"),device="cuda")
nv.check(nv.lib().pgmp_selftest_umma_ts(Ad.data_ptr(),Wd.data_ptr(),D.data_ptr(),nv.current_stream())); torch.cuda.synchronize()
want=(A.double()@W.double().t()).numpy(); got=D.cpu().numpy()
print("TS-mode max rel err", np.abs(got-want).max()/np.abs(want).max())
