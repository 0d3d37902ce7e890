import time, torch
dev = torch.device("cuda", 0)
x = torch.zeros(1, device=dev)
for size in (1 << 10, 1 << 20, 4 << 20, 32 << 20, 128 << 20):
    # steady state: allocate, use on stream, free, repeat
    keep = []
    for _ in range(10):
        t = torch.empty(size, dtype=torch.uint8, device=dev); t.fill_(1); keep.append(t)
        if len(keep) > 3: keep.pop(0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        t = torch.empty(size, dtype=torch.uint8, device=dev)
        keep.append(t)
        if len(keep) > 3: keep.pop(0)
    t1 = time.perf_counter()
    print(size, "bytes: torch.empty", round((t1 - t0) / 200 * 1e6, 1), "us (no kernel in between)")
    t0 = time.perf_counter()
    for _ in range(200):
        t = torch.empty(size, dtype=torch.uint8, device=dev); t[:16].fill_(1)
        keep.append(t)
        if len(keep) > 3: keep.pop(0)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(size, "bytes: torch.empty + tiny kernel", round((t1 - t0) / 200 * 1e6, 1), "us")
print(torch.cuda.memory_stats()["num_alloc_retries"], torch.__version__)
import os; print(os.environ.get("PYTORCH_CUDA_ALLOC_CONF"))
