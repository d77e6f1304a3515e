#!/usr/bin/env python
"""2+ GPUs: the peer-store detection exchange (PeerDetections) must give every rank exactly what an NCCL all-gather of
the per-rank results gives.  torchrun --nproc-per-node N tools/p2p_check.py"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth
from pl_yolo_b200.distributed import PeerDetections, fused_det_buffer, split_gathered

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
B = 6
heads = [torch.from_numpy(h).to(dev) for h in synth.make_heads(B, 320, 80, seed=10 + rank)]
pd = PeerDetections(B, 300, dev, slots=2)
for slot in range(2):
    ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0, out=pd.outputs(slot), peers=pd.peers(slot))
pd.fence()
# reference: plain call + NCCL all-gather of the fused buffer
buf, d, c = fused_det_buffer(B, 300, dev)
ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0, out=(d, c, torch.empty((B, 300), dtype=torch.int32, device=dev)))
g = torch.empty(world * buf.numel(), device=dev)
dist.all_gather_into_tensor(g, buf)
gd, gc = split_gathered(g, world, B, 300)
ok = True
for slot in range(2):
    pdets, pcnt = pd.gathered(slot)
    ok &= bool(torch.equal(pdets, gd) and torch.equal(pcnt, gc))
print("rank %d/%d peer-store exchange == NCCL all-gather: %s (dets/img %s)" % (rank, world, ok, gc.tolist()[:4]), flush=True)
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
pd.close()
dist.destroy_process_group()
sys.exit(int(t.item()))
