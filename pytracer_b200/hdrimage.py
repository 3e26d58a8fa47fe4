"""Array-backed HdrImage: the framebuffer the CUDA renderer fills.

Same interface as the reference's HdrImage (hdrimages.py:59-94: ``width``, ``height``, ``pixels``,
``valid_coordinates``, ``pixel_offset``, ``get_pixel``, ``set_pixel``; pixel (0, 0) is the top-left
corner, offset ``y*width + x``) but the storage is one ``(height, width, 3)`` array that the device
image is copied into directly, instead of a list of width*height ``Color`` objects that costs
seconds of Python to fill at 1080p.  ``pixels`` is a lazy sequence view, so code that indexes
``image.pixels[i]`` (e.g. the reference's own write_pfm / tone mapping) keeps working; it can also
be installed into a *reference* HdrImage (see :func:`install_array`).

PFM I/O (hdrimages.py:96-118, 220-241; golden bytes in tests/test_all.py:112-143) is vectorised:
header ``PF\\n<w> <h>\\n<-1.0|1.0>\\n`` then rows bottom-to-top of fp32 RGB.
"""
from __future__ import annotations

import enum
from typing import Optional

import numpy as np

from .scene import Color


class InvalidPfmFileFormat(Exception):
    pass


class Endianness(enum.Enum):
    """hdrimages.py:31-35"""

    LITTLE_ENDIAN = 1
    BIG_ENDIAN = 2


def _is_little_endian(endianness) -> bool:
    """Accepts this module's enum, the reference's ``pytracer.hdrimages.Endianness`` (duck-typed on
    ``.name``) or a plain bool (True = little endian)."""
    if isinstance(endianness, bool):
        return endianness
    name = getattr(endianness, "name", None)
    if name in ("LITTLE_ENDIAN", "BIG_ENDIAN"):
        return name == "LITTLE_ENDIAN"
    raise TypeError(f"endianness must be an Endianness member or a bool, not {endianness!r}")


class PixelView:
    """``list``-like view of an (H, W, 3) array yielding/accepting ``Color`` objects."""

    def __init__(self, rgb: np.ndarray, color_type=Color):
        self._flat = rgb.reshape(-1, 3)
        self._color_type = color_type

    def __len__(self) -> int:
        return self._flat.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        r, g, b = self._flat[i]
        return self._color_type(float(r), float(g), float(b))

    def __setitem__(self, i: int, color) -> None:
        self._flat[i] = (color.r, color.g, color.b)

    def __iter__(self):
        for r, g, b in self._flat.tolist():
            yield self._color_type(r, g, b)


class HdrImage:
    def __init__(self, width: int = 0, height: int = 0, dtype=np.float32):
        self.width, self.height = int(width), int(height)
        self._rgb = np.zeros((self.height, self.width, 3), dtype=dtype)

    @classmethod
    def from_array(cls, rgb: np.ndarray) -> "HdrImage":
        img = cls.__new__(cls)
        img.height, img.width = int(rgb.shape[0]), int(rgb.shape[1])
        img._rgb = np.ascontiguousarray(rgb)
        return img

    # -- array access used by the renderer and by flatten.image_to_array
    def rgb_array(self) -> np.ndarray:
        return self._rgb

    def pin(self) -> bool:
        """Page-lock the pixel buffer (once) so the device->host copy of a render runs at PCIe speed.
        Returns False when the buffer cannot be pinned (no device, tiny image): copies still work."""
        if getattr(self, "_pinned_ptr", None) == self._rgb.ctypes.data:
            return True
        self.unpin()  # the buffer was replaced since the last pin
        if self._rgb.nbytes < (1 << 16):
            return False
        try:
            from . import _native

            lib = _native.load()
            if lib.rt_host_register(self._rgb.ctypes.data, self._rgb.nbytes) != 0:
                return False
        except Exception:
            return False
        self._pinned_ptr = self._rgb.ctypes.data
        return True

    def unpin(self) -> None:
        """Undo :meth:`pin` (called before the pixel buffer is replaced or freed)."""
        ptr = getattr(self, "_pinned_ptr", None)
        if ptr:
            self._pinned_ptr = None
            try:
                from . import _native

                _native.load().rt_host_unregister(ptr)
            except Exception:
                pass

    def __del__(self):
        self.unpin()

    @property
    def pixels(self) -> PixelView:
        """A view: ``image.pixels[i] = color`` writes through, but the ``Color`` objects it yields are
        copies (``image.pixels[i].r = x`` is lost — use ``set_pixel`` or assign the element)."""
        return PixelView(self._rgb)

    # -- reference interface
    def valid_coordinates(self, x: int, y: int) -> bool:
        return 0 <= x < self.width and 0 <= y < self.height

    def pixel_offset(self, x: int, y: int) -> int:
        return y * self.width + x

    def get_pixel(self, x: int, y: int) -> Color:
        assert self.valid_coordinates(x, y)
        r, g, b = self._rgb[y, x]
        return Color(float(r), float(g), float(b))

    def set_pixel(self, x: int, y: int, new_color) -> None:
        assert self.valid_coordinates(x, y)
        self._rgb[y, x] = (new_color.r, new_color.g, new_color.b)

    # -- tone mapping on the device (hdrimages.py:120-171; csrc/rt_tonemap.cu)
    def average_luminosity(self, delta: float = 1e-10) -> float:
        from . import tonemap

        return tonemap.average_luminosity(self._rgb, delta)

    def normalize_image(self, factor: float, luminosity: Optional[float] = None) -> None:
        from . import tonemap

        hdr, _, _ = tonemap.tone_map(self._rgb, factor, luminosity, want_ldr=False, flags=tonemap.NORMALIZE)
        install_array(self, hdr.astype(self._rgb.dtype, copy=False))

    def clamp_image(self) -> None:
        from . import tonemap

        hdr, _, _ = tonemap.tone_map(self._rgb, want_ldr=False, flags=tonemap.CLAMP)
        install_array(self, hdr.astype(self._rgb.dtype, copy=False))

    def write_ldr_image(self, stream, format: str, gamma: float = 1.0) -> None:
        from . import tonemap

        tonemap.write_ldr_image(self, stream, format, gamma=gamma, flags=0)

    def write_pfm(self, stream, endianness=Endianness.LITTLE_ENDIAN, little_endian: Optional[bool] = None) -> None:
        """hdrimages.py:96-118: ``write_pfm(stream, endianness=Endianness.LITTLE_ENDIAN)``; ``little_endian=``
        is kept as a keyword alias."""
        little_endian = _is_little_endian(endianness) if little_endian is None else bool(little_endian)
        stream.write(f"PF\n{self.width} {self.height}\n{'-1.0' if little_endian else '1.0'}\n".encode("ascii"))
        stream.write(self._rgb[::-1].astype("<f4" if little_endian else ">f4").tobytes())


def read_pfm_image(stream) -> HdrImage:
    def line() -> str:
        out = b""
        while True:
            ch = stream.read(1)
            if ch in (b"", b"\n"):
                return out.decode("ascii")
            out += ch

    if line() != "PF":
        raise InvalidPfmFileFormat("invalid magic in PFM file")
    parts = line().split(" ")
    try:
        if len(parts) != 2:
            raise ValueError
        width, height = int(parts[0]), int(parts[1])
        if width < 0 or height < 0:
            raise ValueError
    except ValueError:
        raise InvalidPfmFileFormat("invalid image size specification")
    try:
        endian = float(line())
    except ValueError:
        raise InvalidPfmFileFormat("missing endianness specification")
    if endian not in (1.0, -1.0):
        raise InvalidPfmFileFormat("invalid endianness specification")
    raw = stream.read(width * height * 12)
    if len(raw) != width * height * 12:
        raise InvalidPfmFileFormat("impossible to read binary data from the file")
    data = np.frombuffer(raw, dtype=">f4" if endian == 1.0 else "<f4").reshape(height, width, 3)
    return HdrImage.from_array(data[::-1].astype(np.float32))


def install_array(image, rgb: np.ndarray, adopt: bool = False) -> None:
    """Put a rendered (H, W, 3) array into any HdrImage-like object.  Our own class adopts the array
    (``adopt=True``: always, without a copy — the node's shared page-locked image of a multi-GPU render,
    whose owner keeps it alive and registered); a reference ``pytracer.hdrimages.HdrImage`` gets a
    :class:`PixelView` as its ``pixels`` (built on the reference's own ``Color`` type) — no per-pixel
    Python work in either case."""
    if isinstance(image, HdrImage):
        if rgb is image._rgb:  # rendered straight into the image's own (page-locked) buffer
            return
        if adopt and rgb.dtype == image._rgb.dtype and rgb.flags.c_contiguous:
            image.unpin()
            image._rgb = rgb
            image.height, image.width = int(rgb.shape[0]), int(rgb.shape[1])
        elif image._rgb.shape == rgb.shape and image._rgb.dtype == rgb.dtype:
            np.copyto(image._rgb, rgb)
        else:
            image.unpin()  # never leave a registration on a buffer numpy is about to free
            image._rgb = np.ascontiguousarray(rgb)
        return
    color_type = Color
    try:
        existing = image.pixels[0] if len(image.pixels) else None
        if existing is not None:
            color_type = type(existing)
    except Exception:
        pass
    image.pixels = PixelView(np.ascontiguousarray(rgb, dtype=np.float64), color_type)
