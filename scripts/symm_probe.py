"""Does torch symmetric memory (peer-mapped buffers over NVLink) work in this sandbox?  torchrun, 2+ ranks."""
import os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    t.fill_(rank + 1)
    hdl.barrier(channel=0)
    peer = hdl.get_buffer((rank + 1) % world, (1 << 20,), torch.uint8)
    v = int(peer[12345].item())
    hdl.barrier(channel=0)
    print(f"rank {rank}: symmetric memory OK, peer value {v}, ptrs {[hex(p) for p in hdl.buffer_ptrs][:3]}, "
          f"signal pads {len(hdl.signal_pad_ptrs)}, multicast {getattr(hdl, 'multicast_ptr', None)}", flush=True)
    # P2P read bandwidth with a plain copy kernel
    big = symm_mem.empty(64 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
    h2 = symm_mem.rendezvous(big, group=dist.group.WORLD)
    h2.barrier(channel=0)
    src = h2.get_buffer((rank + 1) % world, (64 << 20,), torch.uint8)
    dst = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): dst.copy_(src)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): dst.copy_(src)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"rank {rank}: peer read 64 MiB in {dt*1e3:.3f} ms = {(64<<20)/dt/1e9:.0f} GB/s", flush=True)
    h2.barrier(channel=0)
except Exception as e:  # noqa
    print(f"rank {rank}: symmetric memory FAILED: {type(e).__name__}: {e}", flush=True)
# cudaIpc-based alternative: can peers be enabled at all?
print(f"rank {rank}: can_device_access_peer {[torch.cuda.can_device_access_peer(local, j) for j in range(torch.cuda.device_count()) if j != local]}", flush=True)
dist.barrier(); dist.destroy_process_group()
