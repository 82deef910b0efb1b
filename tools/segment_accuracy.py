"""Measures segment mode against the whole-recording decode on the GPU (same kernels, exact path) for several
segment counts and halos; prints one JSON line per case (numbers quoted in DESIGN.md section 6)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from wefax_b200 import segments as S   # noqa: E402
from wefax_b200 import synth           # noqa: E402
from wefax_b200.decoder import Decoder  # noqa: E402


def main():
    rate, seconds = 48000, 1200.0       # BASELINE configs[2]
    pcm = synth.synth_recording(seconds, sample_rate=rate, lpm=120, seed=1, noise_sigma=0.03)
    ds = [Decoder(0) for _ in range(8)]
    whole = ds[0].decode(pcm, rate, 120, want=("demodulated", "digitalized", "raster"))
    dem0, dig0, img0 = whole.demodulated[0].copy(), whole.digitalized[0].copy(), whole.image(0).copy()
    peak = float(np.abs(dem0).max())
    for G in (2, 4, 8):
        for halo in (16538, 65536, 262144):
            res = S.decode_segmented(pcm, rate, 120, ds[:G], halo=halo, want=("raster", "digitalized", "demodulated"))
            dem = np.concatenate([res.demodulated[k] for k in sorted(res.demodulated)])
            dig = np.concatenate([res.digitalized[k] for k in sorted(res.digitalized)])
            print(json.dumps({
                "segments": G, "halo": halo,
                "envelope_max_err_of_peak": float(np.abs(dem - dem0).max() / peak),
                "low_high_err_of_peak": [abs(res.low - whole.low_high[0, 0]) / peak, abs(res.high - whole.low_high[0, 1]) / peak],
                "grey_exact": float((dig == dig0).mean()),
                "grey_within_1": float((np.abs(dig.astype(int) - dig0.astype(int)) <= 1).mean()),
                "start_frame_equal": bool(res.start_frame == int(whole.start_frame[0])),
                "pixels_exact": float((res.image == img0).mean()),
                "pixels_within_1": float((np.abs(res.image.astype(int) - img0.astype(int)) <= 1).mean()),
            }), flush=True)


if __name__ == "__main__":
    main()
