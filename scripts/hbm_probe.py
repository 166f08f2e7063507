"""Practical HBM ceilings for the observation write pattern (int8 state -> int32 obs), for DESIGN.md."""
import json, torch
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
for n in (65536, 1048576):
    g8 = torch.randint(0, 7, (n, 1089), device='cuda', dtype=torch.int8)
    o32 = torch.zeros(n, 1089, device='cuda', dtype=torch.int32)
    a = torch.zeros(n * 1089 * 5 // 8, device='cuda', dtype=torch.int32)   # copy with the same total bytes
    b = torch.zeros_like(a)
    ms_fill = t(lambda: o32.fill_(3))
    ms_widen = t(lambda: o32.copy_(g8))
    ms_copy = t(lambda: b.copy_(a))
    nb = n * 1089
    out[n] = {'fill_int32_ms': ms_fill, 'fill_GBs': nb * 4 / ms_fill / 1e6,
              'widen_int8_to_int32_ms': ms_widen, 'widen_GBs': nb * 5 / ms_widen / 1e6,
              'copy_same_bytes_ms': ms_copy, 'copy_GBs': a.numel() * 8 / ms_copy / 1e6}
print(json.dumps(out, indent=1))
