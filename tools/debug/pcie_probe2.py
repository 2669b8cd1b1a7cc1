"""H2D of 20 MB from pinned memory: one copy, three copies (as the host step does), and K chunks on K streams."""
import torch
dev = torch.device("cuda")
nb = 20 << 20
hbuf = torch.empty(nb, dtype=torch.uint8).pin_memory()
dbuf = torch.empty(nb, dtype=torch.uint8, device=dev)
def timed(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
def one(): dbuf.copy_(hbuf, non_blocking=True)
def three():
    o = 0
    for part in (8 << 20, 4 << 20, 8 << 20):
        dbuf[o:o + part].copy_(hbuf[o:o + part], non_blocking=True); o += part
print("one 20 MB copy        %.3f ms  %.1f GB/s" % (timed(one), nb / timed(one) / 1e6))
print("three copies 8+4+8 MB %.3f ms" % timed(three))
for K in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(K)]
    cur = torch.cuda.current_stream()
    def chunks():
        c = nb // K
        ev = torch.cuda.Event(); ev.record(cur)
        for k, s in enumerate(streams):
            s.wait_event(ev)
            with torch.cuda.stream(s):
                dbuf[k * c:(k + 1) * c].copy_(hbuf[k * c:(k + 1) * c], non_blocking=True)
            cur.wait_stream(s)
    t = timed(chunks)
    print("%d chunks on %d streams  %.3f ms  %.1f GB/s" % (K, K, t, nb / t / 1e6))
# a kernel reading mapped pinned memory directly (zero copy)
src = hbuf.view(torch.int32)
import ctypes
