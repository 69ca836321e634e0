"""PCIe yardsticks of the e2e leg on this box: D2H of three 1 GiB results and H2D of one 1 GiB field from / to pinned
host memory -- one copy stream or several, both directions at once, host buffers first-touched on the GPU's NUMA node
or wherever the allocator puts them.  usage: pcie_copy.py"""
import os
import sys
import time
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import bind_to_gpu_numa_node

GiB = 1 << 30
dev = torch.device("cuda", 0)
pr = torch.cuda.get_device_properties(0)
print("GPU", pr.name, "pci", f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}")
try:
    nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    print("host NUMA nodes:", nodes, "cpus:", len(os.sched_getaffinity(0)))
    bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    print("GPU numa_node:", open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
    print("link:", open(f"/sys/bus/pci/devices/{bdf}/current_link_speed").read().strip(),
          "x" + open(f"/sys/bus/pci/devices/{bdf}/current_link_width").read().strip())
except Exception as e:
    print("topology probe failed:", e)

d = [torch.rand(GiB // 8, dtype=torch.float64, device=dev) for _ in range(4)]


def run(label, bound):
    node, prev = bind_to_gpu_numa_node(0) if bound else (None, None)
    h = [torch.empty(GiB // 8, dtype=torch.float64, pin_memory=True) for _ in range(4)]
    for x in h:
        x.zero_()
    streams = [torch.cuda.Stream() for _ in range(4)]

    def timed(fn, reps=3):
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    def d2h(nstreams, chunks=1):
        def fn():
            k = 0
            for a in range(3):
                n = d[a].numel() // chunks
                for c in range(chunks):
                    with torch.cuda.stream(streams[k % nstreams]):
                        h[a][c * n:(c + 1) * n].copy_(d[a][c * n:(c + 1) * n], non_blocking=True)
                    k += 1
        return fn

    def h2d():
        with torch.cuda.stream(streams[3]):
            d[3].copy_(h[3], non_blocking=True)

    print(f"--- {label} (numa node {node})")
    t = timed(d2h(1)); print(f"D2H 3 GiB, 1 stream            {t * 1e3:7.2f} ms  {3 * GiB / t / 1e9:6.1f} GB/s")
    t = timed(d2h(2, 4)); print(f"D2H 3 GiB, 2 streams           {t * 1e3:7.2f} ms  {3 * GiB / t / 1e9:6.1f} GB/s")
    t = timed(d2h(3)); print(f"D2H 3 GiB, 3 streams           {t * 1e3:7.2f} ms  {3 * GiB / t / 1e9:6.1f} GB/s")
    t = timed(h2d); print(f"H2D 1 GiB                      {t * 1e3:7.2f} ms  {GiB / t / 1e9:6.1f} GB/s")
    t = timed(lambda: (h2d(), d2h(1)())); print(f"H2D 1 GiB + D2H 3 GiB at once  {t * 1e3:7.2f} ms  D2H {3 * GiB / t / 1e9:6.1f} GB/s")
    t = timed(lambda: (h2d(), d2h(2, 4)())); print(f"  the same, D2H on 2 streams   {t * 1e3:7.2f} ms  D2H {3 * GiB / t / 1e9:6.1f} GB/s")
    if prev:
        os.sched_setaffinity(0, prev)
    del h


run("pinned buffers wherever the allocator puts them", False)
run("pinned buffers first-touched on the GPU's NUMA node", True)
