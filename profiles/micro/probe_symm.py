"""Probe: does torch symmetric memory (peer pointers over NVLink) work on this box?  torchrun --nproc-per-node 2."""
import os, sys, time
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
ok = {}
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty((1024, 16), dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    t.fill_(float(rank + 1))
    hdl.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (1024, 16), torch.float32)
    ok["symm_peer_read"] = float(peer[0, 0])
    ok["ptrs"] = [hex(p) for p in hdl.buffer_ptrs]
    ok["signal_pad"] = hdl.signal_pad_size
    ok["multicast"] = bool(hdl.has_multicast_support(dev.type, dev.index)) if hasattr(hdl, "has_multicast_support") else None
    torch.cuda.synchronize()
    # latency of hdl.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(10): hdl.barrier()
    ev0.record()
    for _ in range(100): hdl.barrier()
    ev1.record(); torch.cuda.synchronize()
    ok["symm_barrier_us"] = ev0.elapsed_time(ev1) * 10
except Exception as e:
    ok["symm_error"] = repr(e)[:300]
# NCCL small all-reduce latency for comparison
x = torch.ones(16, device=dev)
for _ in range(10): dist.all_reduce(x)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(100): dist.all_reduce(x)
ev1.record(); torch.cuda.synchronize()
ok["nccl_allreduce16_us"] = ev0.elapsed_time(ev1) * 10
ok["p2p"] = torch.cuda.can_device_access_peer(local, (local + 1) % world)
if rank == 0: print(ok)
dist.barrier(); dist.destroy_process_group()
