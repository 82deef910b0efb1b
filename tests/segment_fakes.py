"""TEST-ONLY stand-in for the GPU segment worker: the same five calls ``Decoder`` implements over the
C-ABI (``segment_envelope / histogram / quantise / sync / raster``), computed with the CPU oracle.  It
lets the host protocol of ``wefax_b200/segments.py`` (planning, histogram exchange, line ownership,
gather) run on CPU and over gloo.  Never imported by the product."""
import numpy as np

from oracle import wefax_oracle as O


class OracleSegmentWorker:
    def segment_envelope(self, pcm, sample_rate, core_begin, core_end, notch_freq=2600, notch_q=1, n_out=None,
                         seam=0):
        pcm = np.asarray(pcm)
        data = O.merge_channels(pcm) if pcm.ndim == 2 else pcm
        if sample_rate != O.TARGET_RATE:
            data = O.resample(data, n_out or O.resampled_length(data.shape[0], sample_rate))
        b, a = O.notch_coefficients(int(notch_freq), notch_q, O.TARGET_RATE)
        # the notch sees the two sides of the seam as the two ends of the recording, the transform runs across
        audio = np.concatenate([O.filtfilt(b, a, data[:seam]), O.filtfilt(b, a, data[seam:])]) if seam else \
            O.filtfilt(b, a, data)
        env = np.abs(O.hilbert(audio)).astype(np.float32)               # before the median, as on the GPU
        n_ext = int(env.shape[0])
        if seam:                                                        # keep the side that holds the core
            self.base = seam if seam <= core_begin else 0
            env = env[seam:] if self.base else env[:seam]
        else:
            self.base = 0
        self.med = O.medfilt5(env.astype(np.float64)).astype(np.float32)
        self.core = (int(core_begin) - self.base, int(core_end) - self.base)
        return n_ext

    def segment_histogram(self, level, prefix=(0, 0, 0, 0)):
        keys = self.med[self.core[0]:self.core[1]].view(np.uint32).astype(np.int64)
        hist = np.zeros((4, 2048), dtype=np.uint32)
        if level == 0:
            hist[0] = np.bincount(keys >> 21, minlength=2048)
            return hist
        for t in range(4):
            if level == 1:
                sel = keys[(keys >> 21) == prefix[t]]
                hist[t] = np.bincount((sel >> 10) & 0x7FF, minlength=2048)
            else:
                sel = keys[(keys >> 10) == prefix[t]]
                hist[t] = np.bincount(sel & 0x3FF, minlength=2048)
        return hist

    def segment_quantise(self, low, high, want=("digitalized",)):
        with np.errstate(divide="ignore", invalid="ignore"):
            d = np.round(255 * (self.med.astype(np.float64) - low) / (high - low))
        self.dig = np.clip(np.nan_to_num(d), 0, 255).astype(np.uint8)
        out = {}
        if "digitalized" in want:
            out["digitalized"] = self.dig[self.core[0]:self.core[1]].copy()
        if "demodulated" in want:
            out["demodulated"] = self.med[self.core[0]:self.core[1]].copy()
        return out

    def segment_sync(self, lpm):
        consts = O.line_constants(lpm, O.TARGET_RATE)
        peaks = O.pattern_search(self.dig.astype(np.int64), consts)
        try:
            ph = O.find_phasing(peaks, consts)
            return {"peaks": list(peaks), "phasing_signals": list(ph), "start_frame": ph[-1] if ph else 0,
                    "status": 0}
        except ValueError:
            return {"peaks": list(peaks), "phasing_signals": [], "start_frame": 0, "status": 1}

    def segment_raster(self, lpm, first_sample, n_lines, skip_lines, keep_lines, out=None):
        w = O.line_constants(lpm, O.TARGET_RATE)["width"]
        first_sample -= self.base
        assert first_sample >= 0 and first_sample + n_lines * w <= self.dig.shape[0]
        img = O.convert_to_image(self.dig[first_sample:first_sample + n_lines * w].astype(np.int64), w)
        return np.ascontiguousarray(img[4 * skip_lines:4 * (skip_lines + keep_lines)])
