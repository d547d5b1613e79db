"""Deterministic synthetic 4:2:0 sequences for the lookahead.

Seeded blurred-noise texture, global pan (non-trivial MVs), per-frame noise, moving
rectangles (occlusion / intra blocks), optional hard cuts, fades and flashes.  Pure numpy;
identical output for identical arguments on every box (default_rng + integer arithmetic).
"""
import numpy as np


def _box_blur(a, k):
    c = np.cumsum(np.cumsum(a, axis=0, dtype=np.float64), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)))
    return (c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]) / (k * k)


class SynthSequence:
    def __init__(self, width, height, depth=8, seed=1, cuts=(), fades=(), flashes=(),
                 n_rects=6, noise=3, chroma_noise=True, static=False, pan=(4, 2), envelope=()):
        """cuts: frame indices where a new scene starts; fades: (start, length, to_level 0..1);
        flashes: (frame, length).  static=True gives a motionless scene (exercises the
        B-frame zero-MV skip rule).  envelope: (frame, level) points of a piecewise-linear luminance envelope
        (fade-ins as well as fade-outs), applied on top of `fades`."""
        self.w, self.h, self.depth, self.seed = width, height, depth, seed
        self.cuts = sorted(cuts)
        self.fades, self.flashes = list(fades), list(flashes)
        self.envelope = sorted(envelope)
        self.noise, self.static, self.pan = noise, static, pan
        self.maxv = (1 << depth) - 1
        rng = np.random.default_rng(seed)
        th, tw = height + 320, width + 320
        self.tex = []
        for _ in range(len(self.cuts) + 1 if len(self.cuts) < 4 else 4):
            t = _box_blur(rng.random((th + 8, tw + 8)), 8)[:th, :tw]
            t = (t - t.min()) / (t.max() - t.min())
            self.tex.append(np.round(t * 255.0).astype(np.int32))
        ch, cw = (height + 1) // 2 + 160, (width + 1) // 2 + 160
        if chroma_noise:
            cu = _box_blur(rng.random((ch + 8, cw + 8)), 8)[:ch, :cw]
            cv = _box_blur(rng.random((ch + 8, cw + 8)), 8)[:ch, :cw]
            self.cu = np.round(128 + (cu - cu.mean()) * 200).clip(16, 240).astype(np.int32)
            self.cv = np.round(128 + (cv - cv.mean()) * 200).clip(16, 240).astype(np.int32)
        else:
            self.cu = np.full((ch, cw), 128, np.int32)
            self.cv = np.full((ch, cw), 128, np.int32)
        self.rects = []
        for _ in range(n_rects):
            rw = int(rng.integers(max(8, width // 16), max(9, width // 5)))
            rh = int(rng.integers(max(8, height // 16), max(9, height // 5)))
            self.rects.append(dict(x=int(rng.integers(0, max(1, width - rw))), y=int(rng.integers(0, max(1, height - rh))),
                                   w=rw, h=rh, dx=int(rng.integers(-12, 13)), dy=int(rng.integers(-12, 13)),
                                   lum=int(rng.integers(30, 226)), tx=int(rng.integers(0, 300)), ty=int(rng.integers(0, 300))))

    def _scene(self, i):
        s = 0
        for c in self.cuts:
            if i >= c:
                s += 1
        return s

    def frame(self, i):
        """Returns (Y, U, V) as uint8 (depth 8) or uint16 arrays."""
        w, h = self.w, self.h
        s = self._scene(i)
        tex = self.tex[s % len(self.tex)]
        start = self.cuts[s - 1] if s else 0
        k = 0 if self.static else (i - start)
        ox = (self.pan[0] * k + 37 * s) % 300
        oy = (self.pan[1] * k + 53 * s) % 300
        y = tex[oy:oy + h, ox:ox + w].copy()
        if s & 1:
            y = 255 - y
        for r in self.rects:
            kk = 0 if self.static else i
            rx = (r['x'] + r['dx'] * kk) % max(1, w - r['w'])
            ry = (r['y'] + r['dy'] * kk) % max(1, h - r['h'])
            patch = tex[r['ty']:r['ty'] + r['h'], r['tx']:r['tx'] + r['w']]
            y[ry:ry + r['h'], rx:rx + r['w']] = (patch + r['lum']) // 2
        if self.noise:
            rng = np.random.default_rng((self.seed << 20) + i)
            y = y + rng.integers(-self.noise, self.noise + 1, size=y.shape)
        lum = 1.0
        for (fs, fl, lvl) in self.fades:
            if fs <= i < fs + fl:
                lum = 1.0 + (lvl - 1.0) * (i - fs + 1) / fl
            elif i >= fs + fl:
                lum = lvl
        if self.envelope:
            lum *= float(np.interp(i, [p[0] for p in self.envelope], [p[1] for p in self.envelope]))
        if lum != 1.0:
            y = np.floor(y * lum + 0.5).astype(np.int64)
        for (ff, fl) in self.flashes:
            if ff <= i < ff + fl:
                y = y // 4 + 190
        y = np.clip(y, 0, 255)
        cox, coy = ox // 2, oy // 2
        u = self.cu[coy:coy + (h + 1) // 2, cox:cox + (w + 1) // 2]
        v = self.cv[coy:coy + (h + 1) // 2, cox:cox + (w + 1) // 2]
        if self.depth == 8:
            return y.astype(np.uint8), u.astype(np.uint8), v.astype(np.uint8)
        sh = self.depth - 8
        # spread into the low bits so 10-bit arithmetic is really exercised
        y16 = (y.astype(np.int64) << sh) + ((y.astype(np.int64) * 7 + i) & ((1 << sh) - 1))
        return (np.clip(y16, 0, self.maxv).astype(np.uint16),
                (u.astype(np.int64) << sh).astype(np.uint16), (v.astype(np.int64) << sh).astype(np.uint16))


def write_y4m(path, seq, n_frames, fps=(30, 1)):
    tag = "C420jpeg" if seq.depth == 8 else "C420p%d" % seq.depth
    with open(path, "wb") as f:
        f.write(("YUV4MPEG2 W%d H%d F%d:%d Ip A1:1 %s\n" % (seq.w, seq.h, fps[0], fps[1], tag)).encode())
        for i in range(n_frames):
            f.write(b"FRAME\n")
            for p in seq.frame(i):
                f.write(np.ascontiguousarray(p).tobytes())
