"""Deterministic synthetic inputs: a portable splitmix64 stream, never rand().

Planes are padded on every side and their row pitch is rounded up to 256 bytes, like a codec's frame store.  All
generators are integer-only numpy so host tests, the GPU tests and bench.py see identical bytes everywhere.
"""
import numpy as np

SEED = 0x48455643  # "HEVC"
_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed, n):
    """n outputs of splitmix64 started at `seed` (vectorised: output i depends only on seed and i)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + _GOLDEN * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def aligned_empty(shape, dtype, align=256):
    """C-contiguous array whose data pointer is `align`-byte aligned (a codec's frame store is; the reference's SIMD
    SAD needs 32-byte aligned source rows, reference vp9_sad4d_intrin_avx2.c:34)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    raw = np.empty(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


def aligned_copy(a, align=256):
    out = aligned_empty(a.shape, a.dtype, align)
    out[...] = a
    return out


def random_bytes(seed, n):
    return aligned_copy(splitmix64(seed, (n + 7) // 8).view(np.uint8)[:n])


def random_int16(seed, n, lo=-32768, hi=32767):
    r = splitmix64(seed, (n + 3) // 4).view(np.uint16)[:n].astype(np.int64)
    return aligned_copy((lo + r % (hi - lo + 1)).astype(np.int16))


def pitch_for(width, pad, align=256):
    return (width + 2 * pad + align - 1) // align * align


class Planes:
    """n_frames padded planes in one C-contiguous array buf[n_frames, height + 2*pad, pitch]."""

    def __init__(self, buf, width, height, pad):
        self.buf, self.width, self.height, self.pad = buf, width, height, pad
        self.n_frames, self.rows, self.pitch = buf.shape
        self.itemsize = buf.dtype.itemsize

    @property
    def frame_stride(self):  # elements between frames
        return self.rows * self.pitch

    @property
    def origin(self):  # element offset of sample (0, 0) of frame 0
        return self.pad * self.pitch + self.pad

    def interior(self, f=0):
        p = self.pad
        return self.buf[f, p:p + self.height, p:p + self.width]

    def like(self, buf):
        return Planes(buf, self.width, self.height, self.pad)


def random_planes(seed, n_frames, width, height, pad=80, dtype=np.uint8):
    pitch = pitch_for(width, pad)
    rows = height + 2 * pad
    raw = random_bytes(seed, n_frames * rows * pitch * np.dtype(dtype).itemsize)
    return Planes(raw.view(dtype).reshape(n_frames, rows, pitch), width, height, pad)


def smooth_planes(seed, n_frames, width, height, pad=80, shift=(0, 0), noise=3):
    """Natural-ish content: integer gradients + blocks + low-amplitude noise.  `shift` displaces the content, so a
    (src, ref) pair made with different shifts has a meaningful SAD minimum at that displacement."""
    pitch = pitch_for(width, pad)
    rows = height + 2 * pad
    y = np.arange(rows, dtype=np.int64)[:, None] + shift[1]
    x = np.arange(pitch, dtype=np.int64)[None, :] + shift[0]
    out = aligned_empty((n_frames, rows, pitch), np.uint8)
    for f in range(n_frames):
        base = 128 + ((x * 3 + y * 5 + f * 7) % 97) - 48 + (((x >> 4) ^ (y >> 4)) & 1) * 24 + ((x * y) >> 9) % 31
        nz = random_bytes(seed + 1000003 * f, rows * pitch).reshape(rows, pitch).astype(np.int64) % (2 * noise + 1) - noise
        out[f] = np.clip(base + nz, 0, 255).astype(np.uint8)
    return Planes(out, width, height, pad)


def residual_planes(seed, n_frames, width, height, pad=0):
    """int16 residuals in [-256, 255], the distribution of reference residual_decode.c:1000."""
    pitch = pitch_for(width, pad, align=128)
    rows = height + 2 * pad
    r = splitmix64(seed, (n_frames * rows * pitch + 3) // 4).view(np.uint16)[:n_frames * rows * pitch]
    return Planes(aligned_copy(((r & 0x1FF).astype(np.int16) - 0x100).reshape(n_frames, rows, pitch)), width, height, pad)


def grid_xy(width, height, w, h):
    """int16 {x, y} of the floor(width/w) x floor(height/h) non-overlapping blocks, raster order."""
    xs, ys = np.arange(width // w) * w, np.arange(height // h) * h
    return np.stack(np.meshgrid(xs, ys), axis=-1).reshape(-1, 2).astype(np.int16)


def window_candidates(dx0, dy0, ncx, ncy):
    """int16 {dx, dy} of a dense candidate window in raster order c = (dy-dy0)*ncx + (dx-dx0)."""
    dx, dy = np.meshgrid(np.arange(dx0, dx0 + ncx), np.arange(dy0, dy0 + ncy))
    return np.stack([dx, dy], axis=-1).reshape(-1, 2).astype(np.int16)
