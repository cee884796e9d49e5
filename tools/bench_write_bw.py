"""HBM bandwidth of a WRITE-ONLY stream next to the read+write copy figure of MEASURED_PEAKS.json: the yardstick for the
kernels of the path whose traffic is almost all stores (conv1_fused_kernel writes 6.4 MB of bf16 per snippet and reads
0.15-1.0 MB).  Event-timed on the current stream, buffers of 1.6 GB (>> the 126 MB L2), median of 7 after 2 warm-ups."""
import json

import torch


def med_ms(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    n = 250 * 224 * 224 * 64                       # bf16 elements of one 250-snippet conv1_1 output: 1.6 GB
    x = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    nbytes = n * 2
    res = {"bytes": nbytes}
    for name, fn, moved in (("memset (cudaMemsetAsync via zero_)", lambda: x.zero_(), nbytes),
                            ("fill kernel (torch fill_, vector stores)", lambda: x.fill_(1.0), nbytes),
                            ("copy (read + write)", lambda: y.copy_(x), 2 * nbytes),
                            ("read-only (sum reduction)", lambda: x.view(torch.int16).sum(dtype=torch.int32), nbytes)):
        try:
            ms = med_ms(fn)
            res[name] = {"ms": ms, "GBps": moved / (ms * 1e-3) / 1e9}
        except Exception as e:
            res[name] = {"unavailable": repr(e)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
