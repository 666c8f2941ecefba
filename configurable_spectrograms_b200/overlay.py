"""Annotation sprites: the host half of a figure ("the host only overlays
axes / markers").

Everything a figure shows besides its colour-mapped panels -- titles, axis and tick labels, tick marks,
axes frames, colour bars, bracket markers -- is a rectangle filled from a small RGBA *sprite*: a rendered
text string, a single pixel of one colour (a line or a frame edge is that pixel stretched over the
rectangle) or a 256-entry colour ramp (a colour bar is the ramp stretched over its box).  Sprites live in
one flat pixel store, the atlas, each exactly once however many figures use it; a figure refers to them by
pixel offset from tiles of the same kind as its panels (``csg_png_tile``, flag ``TILE_OVERLAY``), so the
device composes annotations and panels in one pass and the host never rasterises a figure.

Text is rendered with Pillow (a dependency of the reference too, ``pyproject.toml:30``); a label that
repeats -- "Energy (eV)", "Full", a tick label, a time of day -- is rendered once per process.
"""

from __future__ import annotations

import threading

import numpy as np

_WHITE = (255, 255, 255, 255)


class SpriteAtlas:
    """Flat store of RGBA sprites (uint32 pixels, row-major, top row first) with a device mirror per context."""

    def __init__(self):
        self._chunks: list[np.ndarray] = []
        self._size = 0
        self._index: dict = {}
        self._lock = threading.Lock()
        self._fonts: dict = {}
        self._device: dict = {}  # id(ctx) -> (DevBuf, pixels uploaded)

    # ------------------------------------------------------------------ sprites
    def _add(self, key, pixels: np.ndarray):
        """pixels: (h, w, 4) uint8 -> (offset, h, w)."""
        h, w = pixels.shape[:2]
        flat = np.ascontiguousarray(pixels, dtype=np.uint8).reshape(-1).view(np.uint32)
        with self._lock:
            hit = self._index.get(key)
            if hit is None:
                hit = self._index[key] = (self._size, h, w)
                self._chunks.append(flat)
                self._size += flat.size
        return hit

    def solid(self, color) -> tuple[int, int, int]:
        """A 1 x 1 sprite of one colour (stretched by its tile into lines, frames, filled boxes)."""
        key = ("solid", tuple(int(c) for c in color))
        hit = self._index.get(key)
        return hit if hit is not None else self._add(key, np.array([[key[1]]], dtype=np.uint8))

    def ramp(self, lut259: np.ndarray) -> tuple[int, int, int]:
        """The 256 colours of a colormap as a 256 x 1 sprite, highest index first (the top of a colour bar)."""
        table = np.ascontiguousarray(lut259[:256], dtype=np.uint8)
        key = ("ramp", table.tobytes())
        hit = self._index.get(key)
        return hit if hit is not None else self._add(key, table[::-1].reshape(256, 1, 4))

    def font(self, px: int):
        f = self._fonts.get(px)
        if f is None:
            from PIL import ImageFont

            try:
                f = ImageFont.load_default(size=px)
            except TypeError:  # Pillow without a scalable default font
                f = ImageFont.load_default()
            self._fonts[px] = f
        return f

    def text(self, string: str, px: int, color=(0, 0, 0, 255), rotate: bool = False, background=_WHITE) -> tuple[int, int, int]:
        """``string`` rendered ``px`` pixels high (multi-line strings are centred line by line), opaque on
        ``background``; ``rotate``: reading bottom to top (a y-axis label)."""
        key = ("text", string, int(px), tuple(color), bool(rotate), tuple(background))
        hit = self._index.get(key)
        if hit is not None:
            return hit
        from PIL import Image, ImageDraw

        font = self.font(int(px))
        probe = ImageDraw.Draw(Image.new("RGBA", (1, 1)))
        left, top, right, bottom = probe.multiline_textbbox((0, 0), string, font=font, align="center", spacing=max(2, px // 5))
        left, top = int(np.floor(left)), int(np.floor(top))  # (multi-line boxes come back as floats)
        w, h = max(1, int(np.ceil(right)) - left + 2), max(1, int(np.ceil(bottom)) - top + 2)
        image = Image.new("RGBA", (w, h), tuple(int(c) for c in background))
        ImageDraw.Draw(image).multiline_text((1 - left, 1 - top), string, font=font, fill=tuple(int(c) for c in color), align="center",
                                             spacing=max(2, px // 5))
        if rotate:
            image = image.transpose(Image.ROTATE_90)
        return self._add(key, np.asarray(image, dtype=np.uint8))

    # ------------------------------------------------------------------- access
    def pixels(self) -> np.ndarray:
        """The whole store as one uint32 array (host composer, uploads)."""
        with self._lock:
            if len(self._chunks) > 1:
                self._chunks = [np.concatenate(self._chunks)]
            return self._chunks[0] if self._chunks else np.zeros(0, np.uint32)

    def sprite(self, ref) -> np.ndarray:
        """(h, w, 4) uint8 view of one sprite."""
        off, h, w = ref
        return self.pixels()[off : off + h * w].view(np.uint8).reshape(h, w, 4)

    def device_ptr(self, ctx) -> int:
        """Device address of the atlas on ``ctx``; sprites added since the last call are uploaded (the buffer
        grows geometrically; a growth re-uploads everything)."""
        flat = self.pixels()
        buf, done = self._device.get(id(ctx), (None, 0))
        need = max(flat.size, 1) * 4
        if buf is None or buf.nbytes < need:
            ctx.sync()  # kernels may still read the buffer that is about to be replaced
            buf, done = ctx.alloc(max(2 * need, 1 << 20)), 0
        if flat.size > done:
            buf.upload(flat[done:], done * 4)
        self._device[id(ctx)] = (buf, flat.size)
        return buf.ptr


#: the process-wide atlas (sprites are immutable and shared by every figure)
ATLAS = SpriteAtlas()
