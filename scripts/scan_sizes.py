"""Single GPU: exclusive_scan throughput by size, warp-specialised kernel against scan_tma_kernel (set BCB_SCAN_WS=0)."""
import os, sys
import torch
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
import compute_b200 as cb

for dt in (torch.int32, torch.float64):
    for lg in range(18, 29):
        n = (1 << lg) // (2 if dt == torch.float64 else 1)
        x = (torch.rand(n, device="cuda") * 25).to(dt)
        y = torch.empty_like(x)
        for _ in range(3):
            cb.exclusive_scan(x, y, 0)
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            cb.exclusive_scan(x, y, 0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{dt} 2^{lg} B x4: {ms*1e3:9.1f} us  {2 * n * x.element_size() / ms / 1e6:8.1f} GB/s   WS={os.environ.get('BCB_SCAN_WS', '1')} min={os.environ.get('BCB_SCAN_WS_MIN_LOG2', '-')}", flush=True)
