"""D2H rate of cudaMemcpy2DAsync as used by the trace streamer, vs a linear copy (torch as CUDA plumbing)."""
import torch, time
n_chains, n_rows = 1024, 2000
for width_d in (175, 345):
    dev = torch.empty((n_chains, n_rows, width_d), dtype=torch.float64, device="cuda")
    host = torch.empty((n_chains, n_rows, width_d), dtype=torch.float64).pin_memory()
    for rows in (32, 128, 512, 2000):
        torch.cuda.synchronize(); t = time.perf_counter(); reps = 5
        for r in range(reps):
            host[:, :rows].copy_(dev[:, :rows], non_blocking=True)   # strided 2-D copy: 1024 segments
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
        gb = n_chains * rows * width_d * 8 / 1e9
        print(f"width {width_d} doubles, {rows:4d} rows/chain: {gb/dt:6.1f} GB/s ({1e3*dt:.2f} ms for {gb:.3f} GB)")
