import sys, json, torch
sys.path.insert(0, '.')
from pytracer_b200 import tonemap
peaks = json.load(open('MEASURED_PEAKS.json')) if __import__('os').path.exists('MEASURED_PEAKS.json') else {}
for (h, w) in [(1080, 1920), (2160, 3840), (4320, 7680)]:
    n = h * w
    img = torch.rand((h, w, 3), device='cuda') * 3
    ldr = torch.empty((h, w, 3), dtype=torch.uint8, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda'); flush2 = torch.zeros(64 << 20, dtype=torch.float32, device='cuda')
    lum, mp = [], []
    for it in range(8):
        flush.zero_(); flush2.sum()  # dirty lines written back, then evicted by a clean read
        st = tonemap.tone_map_device(img.data_ptr(), n, 1.0, None, 1.0, 0, ldr.data_ptr())
        if it >= 3:
            lum.append(st['lum_ms']); mp.append(st['map_ms'])
    l, m = sum(lum) / len(lum), sum(mp) / len(mp)
    print(f"{w}x{h}: lum {l*1e3:.1f} us = {12*n/l/1e6:.0f} GB/s ; map {m*1e3:.1f} us = {15*n/m/1e6:.0f} GB/s ; hbm peak {peaks.get('hbm_gbs')}")
