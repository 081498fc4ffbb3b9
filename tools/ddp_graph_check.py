"""torchrun --nproc-per-node N tools/ddp_graph_check.py: N ranks train 6 graphed steps on different batches; the
captured all-reduce must keep the replicas bit-identical, and different from a rank that trains alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import vibravox_b200
from vibravox_b200 import parallel
from vibravox_b200 import data as O

rank, local_rank, world = parallel.init_from_env("nccl")
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
lm = vibravox_b200.build_model(seed=42, device=dev)
body, air = O.synthetic_pairs(4, 16000, seed=100 + rank)
batch = {"audio_body_conducted": body.to(dev), "audio_airborne": air.to(dev)}
assert lm.graph_capturable()
for it in range(6):
    lm.training_step_graphed(batch)
torch.cuda.synchronize()
assert lm.graph_launches() > 500, "step was not captured"
for name, opt in (("G", lm.generator_optimizer), ("D", lm.discriminator_optimizer)):
    flat = opt.flat.clone()
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = bool(torch.equal(flat, ref))
    allsame = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(allsame, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name}: replicas bit-identical after 6 graphed steps: {bool(allsame.item())}; |params| {float(flat.norm()):.6f}")
    assert allsame.item() == 1
# losses differ per rank (different data), parameters do not
l = torch.tensor([float(lm.logged["train/generator/backprop_loss"])], device=dev)
ls = [torch.zeros_like(l) for _ in range(world)]
dist.all_gather(ls, l)
if rank == 0:
    print("per-rank generator losses:", [round(float(x), 4) for x in ls])
dist.barrier()
dist.destroy_process_group()
