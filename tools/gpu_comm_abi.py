"""2-GPU check of the C-ABI gradient exchange (b200ode_comm_*): not a pytest, run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/gpu_comm_abi.py
(1) the ABI all-reduce of a bucket equals torch.distributed's; (2) three train steps with comm=AbiComm give the
same losses and parameters as with torch.distributed, eagerly and replayed from a CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from differential_equations_resnet_b200.parallel import AbiComm
from differential_equations_resnet_b200.training import EulerNet, NetSpec

comm = AbiComm(rank, world)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
a = torch.randn(1_000_003, device="cuda", generator=g)
b = a.clone()
comm.allreduce_bucket(a)
dist.all_reduce(b)
torch.cuda.synchronize()
assert torch.equal(a, b), "ABI all-reduce differs from torch.distributed"

kw = dict(blocks_per_stage=(3, 3, 2), filters_per_block=(16, 32, 64), h=0.25, gamma=-0.05)
gen = torch.Generator().manual_seed(7 + rank)
img = torch.randint(0, 256, (16, 32, 32, 3), generator=gen, dtype=torch.uint8).cuda()
lab = torch.nn.functional.one_hot(torch.randint(0, 10, (16,), generator=gen), 10).float().cuda()
res = {}
for name, c, graph in (("torch", None, False), ("abi", comm, False), ("abi_graph", comm, True)):
    net = EulerNet(NetSpec(**kw), seed=0, world_size=world, comm=c)
    losses = []
    if graph:
        net.capture(img, lab, warmup=1)
        ref = EulerNet(NetSpec(**kw), seed=0)
        net.import_params(ref.export_params()); net.adam_m.zero_(); net.adam_v.zero_(); net.step_counter.fill_(1)
    for _ in range(3):
        l = net.train_step_graph() if graph else net.train_step(img, lab)
        losses.append(float(l))
    torch.cuda.synchronize()
    res[name] = (losses, net.theta.clone())
for k in ("abi", "abi_graph"):
    assert res[k][0] == res["torch"][0], (k, res[k][0], res["torch"][0])
    assert torch.equal(res[k][1], res["torch"][1]), k
# replicas stay identical across ranks
t = res["abi"][1].clone(); dist.broadcast(t, src=0); assert torch.equal(t, res["abi"][1])
dist.barrier(); torch.cuda.synchronize()
if rank == 0:
    print("comm abi OK: all-reduce bit-identical to torch.distributed; losses", res["abi"][0])
os._exit(0)
