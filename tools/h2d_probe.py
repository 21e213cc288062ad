"""H2D / D2H bandwidth from pinned host memory under different CPU affinities (NUMA placement of the pinned pages)."""
import os, sys, time, subprocess
import torch

def bw(nbytes, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for direction in ("h2d", "d2h"):
        for _ in range(3):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        b.record(); torch.cuda.synchronize()
        out.append(nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9)
    return out

if __name__ == "__main__":
    if len(sys.argv) > 1:
        cpus = [int(x) for x in sys.argv[1].split(",")]
        os.sched_setaffinity(0, cpus)
        torch.cuda.init()
        for nb in (2 << 20, 16 << 20, 64 << 20):
            print("cpus", sys.argv[1][:20], "bytes", nb, "h2d/d2h GB/s %.1f %.1f" % tuple(bw(nb)), flush=True)
    else:
        print(subprocess.run("nvidia-smi topo -m; lscpu | grep -i -E 'numa|model name|^CPU\\(s\\)'; cat /sys/bus/pci/devices/*/numa_node | sort | uniq -c; nproc; numactl -H 2>/dev/null | head", shell=True, capture_output=True, text=True).stdout)
        n = os.cpu_count()
        avail = sorted(os.sched_getaffinity(0))
        print("affinity", avail)
        groups = [avail[:max(1, len(avail)//2)], avail[len(avail)//2:]] if len(avail) > 1 else [avail]
        for g in groups:
            subprocess.run([sys.executable, __file__, ",".join(map(str, g))])
