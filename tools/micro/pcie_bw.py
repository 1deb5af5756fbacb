import torch, time
def bw(nbytes, both):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(2):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 10
    for _ in range(n):
        with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    return nbytes / dt / 1e9
for nb in (8 << 10, 256 << 10, 4 << 20, 64 << 20):
    print(f"{nb:>10} B  H2D alone {bw(nb, False):6.1f} GB/s   H2D with concurrent D2H {bw(nb, True):6.1f} GB/s each way")
