#!/bin/bash
# diagnostics: GEMM in-kernel timeline, e2e laps, raw PCIe rates
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
echo "== g8 prof"
VGS_G8_PROF=1 timeout 300 python tools/i8_tune.py 3 2>&1 | tail -12 | cut -c1-700
echo "== pcie"
timeout 300 python - <<'PY'
import torch, time
for mb in (2, 4, 16):
    a = torch.empty(mb << 20, dtype=torch.uint8).pin_memory(); d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for direction in ("h2d", "d2h"):
        f = (lambda: d.copy_(a, non_blocking=True)) if direction == "h2d" else (lambda: a.copy_(d, non_blocking=True))
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t0 = time.perf_counter(); f(); torch.cuda.synchronize(); w = (time.perf_counter() - t0) * 1e3
        print("%s %2d MiB: %.3f ms = %.1f GB/s (single copy wall incl. sync %.3f ms)" % (direction, mb, ms, (mb << 20) / ms / 1e6, w))
PY
echo "== laps"
VGS_LOAD_TIMING=1 timeout 300 python tools/e2e_laps.py 2>&1 | tail -22
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
lscpu | grep -E "Model name|^CPU\(s\)|Thread|MHz" | head
} > gpurun_out/d_diag.log 2>&1
cat gpurun_out/d_diag.log
