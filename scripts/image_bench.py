"""A 12 MP photo -> [13, 3, 448, 448] bf16 pixel values: the GPU front end (pinned host bytes -> device tensor, copy
included) against the reference's host pipeline (PIL resize/crop in dynamic_preprocess + CLIPImageProcessor)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vita_b200.image_frontend import ImageProcessor, closest_grid   # noqa: E402


def main():
    h, w = 3000, 4000
    yy, xx = np.mgrid[0:h, 0:w].astype(np.int64)
    img = np.stack([(xx * 7 + yy * 3) % 256, (xx * xx // 50 + yy) % 256, (255 - (xx + 2 * yy) % 256)], -1).astype(np.uint8)
    host = torch.from_numpy(img).pin_memory()
    ip = ImageProcessor("cuda")
    for _ in range(3):
        px, n = ip.preprocess(host)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for _ in range(reps):
        px, n = ip.preprocess(host)
    b.record(); b.synchronize()
    gpu_ms = a.elapsed_time(b) / reps
    dev = host.cuda()
    a.record()
    for _ in range(reps):
        px, n = ip.preprocess(dev)
    b.record(); b.synchronize()
    gpu_resident_ms = a.elapsed_time(b) / reps
    out = {"image": [h, w], "grid": list(closest_grid(w, h)), "tiles": int(n), "gpu_ms_host_bytes_to_pixels": round(gpu_ms, 3),
           "gpu_ms_device_resident": round(gpu_resident_ms, 3)}
    try:
        from PIL import Image
        from transformers import CLIPImageProcessor
        gi, gj = closest_grid(w, h)
        clip = CLIPImageProcessor(crop_size=448, do_center_crop=True, do_normalize=True, do_resize=True,
                                  image_mean=[0.485, 0.456, 0.406], image_std=[0.229, 0.224, 0.225], resample=3, size=448)
        pil = Image.fromarray(img)

        def cpu():
            big = pil.resize((448 * gi, 448 * gj))
            tiles = [big.crop(((t % gi) * 448, (t // gi) * 448, (t % gi + 1) * 448, (t // gi + 1) * 448)) for t in range(gi * gj)]
            tiles.append(pil.resize((448, 448)))
            return clip.preprocess(tiles, return_tensors="pt")["pixel_values"].to(torch.bfloat16)
        ref = cpu()
        t0 = time.perf_counter()
        for _ in range(3):
            ref = cpu()
        out["cpu_ms_pil_clip"] = round((time.perf_counter() - t0) / 3 * 1e3, 1)
        out["bit_equal_to_cpu"] = bool(torch.equal(ref, px.cpu()))
    except Exception as e:   # noqa: BLE001
        out["cpu"] = f"unavailable: {e}"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
