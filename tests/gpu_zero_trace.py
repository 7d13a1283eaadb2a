import os, sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
from differential_equations_resnet_b200 import _abi
from differential_equations_resnet_b200.layers._base import ChainHandle
N,H,W,C,L = 128,8,8,64,36
lib=_abi.lib()
for mode in ("rand","zero_w","zero_all"):
    ch = ChainHandle(C, L, 0.0)
    params = torch.randn(L*ch.num_params, device="cuda")*0.05
    if mode!="rand": params.zero_()
    ch.pack(params)
    x0 = torch.relu(torch.randn((N,H,W,C), device="cuda"))
    if mode=="zero_all": x0.zero_()
    acts = torch.empty((L,N,H,W,C), device="cuda"); masks=torch.empty((L,N,H,W,C//8),dtype=torch.uint8,device="cuda")
    tr = torch.zeros(1024*16, dtype=torch.int64, device="cuda")
    for rep in range(3):
        tr.zero_(); lib.b200ode_debug_set_trace(ctypes.c_void_p(tr.data_ptr()))
        ch.forward(x0, 0.07, acts=acts, masks=masks); torch.cuda.synchronize(); lib.b200ode_debug_set_trace(None)
    t=tr.cpu().view(-1,16); t=t[t[:,15]!=0]; med=t.float().median(dim=0).values
    print(mode, "L4 mma phase", int(med[4]-med[2]), "wstall", int(med[3]), "epi", int(med[8]-med[6]), "layer", int(med[5]-med[2]), "wall us", (t[:,15].max()-t[:,0].min()).item()/1e3)
