"""Does a large D2H copy slow down a train of small H2D copies on another stream?  (the pattern of two pipelined
builds: 100 files of 5 MB in, 224 MB out.)  Sweeps the piece size of both directions."""
import torch

total_in, total_out = 508 << 20, 224 << 20
h_in = torch.empty(total_in, dtype=torch.uint8).pin_memory()
d_in = torch.empty(total_in, dtype=torch.uint8, device="cuda")
d_out = torch.empty(total_out, dtype=torch.uint8, device="cuda")
h_out = torch.empty(total_out, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d_piece, d2h_piece, reps=3):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    s1.wait_event(e0); s2.wait_event(e0)
    for _ in range(reps):
        if h2d_piece:
            with torch.cuda.stream(s1):
                for o in range(0, total_in, h2d_piece):
                    d_in[o:o + h2d_piece].copy_(h_in[o:o + h2d_piece], non_blocking=True)
        if d2h_piece:
            with torch.cuda.stream(s2):
                for _ in range(2):                      # two results' worth: the D2H leg lasts about as long as the H2D leg
                    for o in range(0, total_out, d2h_piece):
                        h_out[o:o + d2h_piece].copy_(d_out[o:o + d2h_piece], non_blocking=True)
    e1.record(s1); e2.record(s2)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, e0.elapsed_time(e2) / reps


MB = 1 << 20
run(5 * MB, 224 * MB, 1)
print("H2D 508 MiB alone, 5 MiB pieces: %.2f ms;  D2H 2 x 224 MiB alone, one piece each: %.2f ms" % (run(5 * MB, 0)[0], run(0, 224 * MB)[1]))
print("h2d piece  d2h piece   H2D done   D2H done  (ms, both running)")
for hp in (5, 32, 127, 508):
    for dp in (224, 32, 8, 2):
        a, b = run(hp * MB, dp * MB)
        print(f"{hp:6d} MiB {dp:6d} MiB   {a:8.2f}   {b:8.2f}")
