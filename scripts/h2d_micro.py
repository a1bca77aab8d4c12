"""Host -> device bandwidth of this box for the batch copy of the end-to-end leg: one pinned buffer copied with
1 / 2 / 4 concurrent streams (equal chunks), and the topology the copy runs on."""
import os, subprocess, sys, time
import torch

dev = torch.device("cuda:0")
nbytes = int(float(os.environ.get("GB", "3.3")) * 1e9) // 4096 * 4096
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host.fill_(1)
dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for cmd in (["nvidia-smi", "topo", "-m"], ["bash", "-c", "ls /sys/devices/system/node | grep node; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv"]):
    try:
        print(subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout)
    except Exception as e:
        print("failed", cmd, e)
for ns in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    chunk = nbytes // ns // 4096 * 4096
    best = 0.0
    for it in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * chunk:(i + 1) * chunk].copy_(host[i * chunk:(i + 1) * chunk], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = max(best, chunk * ns / dt / 1e9)
    print(f"{ns} stream(s): {best:.1f} GB/s", flush=True)
# small chunks on one stream (the loader's tensors are copied one by one)
for mb in (16, 64, 256):
    chunk = mb << 20
    n = nbytes // chunk
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        dst[i * chunk:(i + 1) * chunk].copy_(host[i * chunk:(i + 1) * chunk], non_blocking=True)
    torch.cuda.synchronize()
    print(f"1 stream, {mb} MiB chunks: {n * chunk / (time.perf_counter() - t0) / 1e9:.1f} GB/s", flush=True)
