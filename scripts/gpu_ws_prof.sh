#!/bin/bash
# per-phase cycle counters of onesweep_ws (library built with -DBCB_WS_PROFILE): workers / helpers, mean over CTAs
python - <<'PY'
import torch, compute_b200 as cb
x = torch.randint(-2**31, 2**31 - 1, (1 << 28,), dtype=torch.int32, device="cuda").view(torch.uint32)
for _ in range(2):
    y = x.clone()
    cb.sort(y)
    torch.cuda.synchronize()
PY
