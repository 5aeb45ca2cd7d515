import os, torch, torch.distributed as dist, time
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda",rank))
if rank==0:
    print("peer access 0->1:", torch.cuda.can_device_access_peer(0,1))
for n,dt in ((64*1024*1024, torch.int32),(576*1024*1024, torch.int32),(64*1024*1024, torch.float32)):
    t=torch.ones(n,dtype=dt,device="cuda")
    for op in ("reduce","all_reduce"):
        for _ in range(3):
            (dist.reduce(t,0) if op=="reduce" else dist.all_reduce(t))
        torch.cuda.synchronize(); dist.barrier()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            (dist.reduce(t,0) if op=="reduce" else dist.all_reduce(t))
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5
        if rank==0: print(f"{op} {dt} {n*4/1e6:.0f} MB: {ms:.2f} ms -> {n*4/ms/1e6:.1f} GB/s")
dist.destroy_process_group()
