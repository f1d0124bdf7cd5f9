"""PCIe: host->device and device->host copies alone and at the same time (pinned memory, two streams).
Why the end-to-end number of a stream of builds stops at the sum of both legs: on the boxes of this pool the two
directions share ~50 GB/s (profiles/r02_pcie_duplex.txt)."""
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    e1.record(s1); e2.record(s2)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), e0.elapsed_time(e2))
    return ms / reps


for _ in range(2):
    run(True, True, 2)
a, b, c = run(True, False), run(False, True), run(True, True)
gb = n / 1e9
print(f"H2D alone   {gb / a * 1e3:6.1f} GB/s")
print(f"D2H alone   {gb / b * 1e3:6.1f} GB/s")
print(f"both at once {gb / c * 1e3:6.1f} GB/s each way, {2 * gb / c * 1e3:6.1f} GB/s in total ({c:.2f} ms for 2 x 512 MiB)")
