"""Annotation sprites: the host half of a figure ("the host only overlays
axes / markers").

Everything a figure shows besides its colour-mapped panels -- titles, axis and tick labels, tick marks,
axes frames, colour bars, bracket markers -- is a rectangle filled from a small RGBA *sprite*: a rendered
text string, a single pixel of one colour (a line or a frame edge is that pixel stretched over the
rectangle) or a 256-entry colour ramp (a colour bar is the ramp stretched over its box).  Sprites live in
one flat pixel store, the atlas, each exactly once however many figures use it; a figure refers to them by
pixel offset from tiles of the same kind as its panels (``csg_png_tile``, flag ``TILE_OVERLAY``), so the
device composes annotations and panels in one pass and the host never rasterises a figure.

Glyphs are rendered with Pillow (a dependency of the reference too, ``pyproject.toml:30``), once per character
and size; strings are composed from them here, and a label that repeats -- "Energy (eV)", "Full", a tick label --
is composed once per process.
"""

from __future__ import annotations

import threading

import numpy as np

_WHITE = (255, 255, 255, 255)


class SpriteAtlas:
    """Flat store of RGBA sprites (uint32 pixels, row-major, top row first) with a device mirror per context."""

    def __init__(self):
        self._store = np.empty(0, dtype=np.uint32)  # every sprite's pixels, back to back; grows by doubling
        self._size = 0
        self._index: dict = {}
        self._lock = threading.Lock()
        self._fonts: dict = {}
        self._glyphs: dict = {}  # (character, px) -> (coverage mask, advance)
        self._words: dict = {}  # (run of characters, px) -> (coverage mask, advance)
        self._parts: dict = {}  # (string, px, colour, background) -> text_parts() result
        self._blends: dict = {}  # (colour, background) -> (256, 4) uint8
        self._dependents: list = []
        self._device: dict = {}  # id(ctx) -> (DevBuf, pixels uploaded)

    # ------------------------------------------------------------------ sprites
    def _add(self, key, pixels: np.ndarray):
        """pixels: (h, w, 4) uint8 -> (offset, h, w)."""
        h, w = pixels.shape[:2]
        flat = np.ascontiguousarray(pixels, dtype=np.uint8).reshape(-1).view(np.uint32)
        with self._lock:
            hit = self._index.get(key)
            if hit is None:
                need = self._size + flat.size
                if need > len(self._store):  # the store doubles: readers keep views of the old one, which stay valid
                    grown = np.empty(max(need, 2 * len(self._store), 1 << 16), dtype=np.uint32)
                    grown[: self._size] = self._store[: self._size]
                    self._store = grown
                self._store[self._size : need] = flat
                hit = self._index[key] = (self._size, h, w)
                self._size = need
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

    def _glyph(self, char: str, px: int):
        """``(coverage mask (line height, width) uint8, advance in pixels)`` of one character: rendered by
        Pillow once per (character, size), on the font's line box so that every glyph shares the baseline."""
        key = (char, px)
        hit = self._glyphs.get(key)
        if hit is None:
            from PIL import Image, ImageDraw

            font = self.font(px)
            try:
                ascent, descent = font.getmetrics()
                advance = float(font.getlength(char))
            except AttributeError:  # Pillow's bitmap fallback font
                left, top, right, bottom = font.getbbox(char)
                ascent, descent, advance = bottom, 0, float(right)
            width = max(1, int(np.ceil(advance)) + 2)  # (room for a glyph that overhangs its advance a little)
            image = Image.new("L", (width, max(1, ascent + descent)), 0)
            ImageDraw.Draw(image).text((0, 0), char, font=font, fill=255)
            hit = self._glyphs[key] = (np.asarray(image, dtype=np.uint8), advance)
        return hit

    def _word(self, word: str, px: int):
        """``(coverage mask, advance)`` of a run of characters without spaces: its glyphs side by side at their
        cumulative advances.  Cached: the words of a title repeat from orbit to orbit, only its numbers change."""
        key = (word, px)
        hit = self._words.get(key)
        if hit is None:
            glyphs = [self._glyph(c, px) for c in word]
            height = max(g.shape[0] for g, _a in glyphs)
            pen, spots = 0.0, []
            for _mask, advance in glyphs:
                spots.append(int(round(pen)))
                pen += advance
            width = max(int(np.ceil(pen)), max(x + g.shape[1] for x, (g, _a) in zip(spots, glyphs)))
            mask = np.zeros((height, width), dtype=np.uint8)
            for x, (g, _advance) in zip(spots, glyphs):
                view = mask[: g.shape[0], x : x + g.shape[1]]
                np.maximum(view, g, out=view)
            if len(self._words) > 200_000:
                self._words.clear()
            hit = self._words[key] = (mask, pen)
        return hit

    def _line(self, line: str, px: int) -> np.ndarray:
        """Coverage mask of one line of text: its words at their cumulative advances."""
        words = line.split(" ")
        if len(words) == 1 and words[0]:
            return self._word(words[0], px)[0]
        space = self._glyph(" ", px)[1]
        pen, placed = 0.0, []
        for k, word in enumerate(words):
            if k:
                pen += space
            if word:
                mask, advance = self._word(word, px)
                placed.append((int(round(pen)), mask))
                pen += advance
        if not placed:
            return np.zeros((1, 1), dtype=np.uint8)
        height = max(m.shape[0] for _x, m in placed)
        width = max(int(np.ceil(pen)), max(x + m.shape[1] for x, m in placed))
        out = np.zeros((height, width), dtype=np.uint8)
        for x, m in placed:
            view = out[: m.shape[0], x : x + m.shape[1]]
            np.maximum(view, m, out=view)
        return out

    def text(self, string: str, px: int, color=(0, 0, 0, 255), rotate: bool = False, background=_WHITE) -> tuple[int, int, int]:
        """``string`` rendered ``px`` pixels high (multi-line strings are centred line by line), opaque on
        ``background``; ``rotate``: reading bottom to top (a y-axis label).

        A directory run shows thousands of distinct strings (every orbit has its own times of day, titles and
        colour-bar values); shaping each with Pillow cost 0.7 ms, more than everything else the host does for
        the figure.  Strings are therefore composed here from per-character coverage masks (Pillow renders a
        character once per size; no kerning), ~30 us each."""
        if type(color) is not tuple:
            color = tuple(int(c) for c in color)
        if type(background) is not tuple:
            background = tuple(int(c) for c in background)
        key = ("text", string, px, color, rotate, background)
        hit = self._index.get(key)
        if hit is not None:
            return hit
        key = ("text", string, int(px), tuple(int(c) for c in color), bool(rotate), tuple(int(c) for c in background))
        hit = self._index.get(key)
        if hit is not None:
            self._index[("text", string, px, color, rotate, background)] = hit  # (the caller's spelling of the same key)
            return hit
        px = int(px)
        lines = [self._line(line, px) for line in string.split("\n")]
        gap = max(2, px // 5)
        width = max(m.shape[1] for m in lines)
        height = sum(m.shape[0] for m in lines) + gap * (len(lines) - 1)
        cover = np.zeros((height, width), dtype=np.uint8)
        y = 0
        for m in lines:
            x = (width - m.shape[1]) // 2
            cover[y : y + m.shape[0], x : x + m.shape[1]] = m
            y += m.shape[0] + gap
        # crop to the ink, one pixel of margin all round (the sprite's size is what the layout centres)
        rows, cols = np.flatnonzero(cover.any(axis=1)), np.flatnonzero(cover.any(axis=0))
        if len(rows) and len(cols):
            ink = cover[rows[0] : rows[-1] + 1, cols[0] : cols[-1] + 1]
        else:
            ink = np.zeros((1, 1), dtype=np.uint8)
        framed = np.zeros((ink.shape[0] + 2, ink.shape[1] + 2), dtype=np.uint8)
        framed[1:-1, 1:-1] = ink
        blend = self._blend(key[3], key[5])  # coverage -> colour
        pixels = blend[framed]
        if rotate:
            pixels = np.rot90(pixels)
        return self._add(key, pixels)

    def text_parts(self, string: str, px: int, color=(0, 0, 0, 255), background=_WHITE):
        """``(width, height, [(sprite, dx, dy), ...])``: ``string`` as one sprite per WORD, placed relative to
        the text's top-left corner (lines centred).  For the long strings that change with every orbit --
        titles: a 900-pixel title sprite cost 0.4 ms to compose and upload although only its orbit number was
        new -- the words are sprites of their own, shared by every title that uses them."""
        px = int(px)
        fg, bg = tuple(int(c) for c in color), tuple(int(c) for c in background)
        cache_key = (string, px, fg, bg)
        cached = self._parts.get(cache_key)
        if cached is not None:
            return cached
        space = self._glyph(" ", px)[1]
        gap = max(2, px // 5)
        lines, width, y = [], 0, 0
        for line in string.split("\n"):
            pen, placed, line_h = 0.0, [], 1
            for k, word in enumerate(line.split(" ")):
                if k:
                    pen += space
                if not word:
                    continue
                key = ("word", word, px, fg, bg)
                ref = self._index.get(key)
                mask, advance = self._word(word, px)
                if ref is None:
                    blend = self._blend(fg, bg)
                    ref = self._add(key, blend[mask])
                placed.append((ref, int(round(pen))))
                pen += advance
                line_h = max(line_h, mask.shape[0])
            line_w = max([int(np.ceil(pen))] + [x + ref[2] for ref, x in placed])
            lines.append((placed, line_w, y))
            width = max(width, line_w)
            y += line_h + gap
        height = max(1, y - gap)
        parts = [(ref, x + (width - line_w) // 2, top) for placed, line_w, top in lines for ref, x in placed]
        if len(self._parts) > 50_000:
            self._parts.clear()
        self._parts[cache_key] = (width, height, parts)
        return width, height, parts

    def _blend(self, fg, bg) -> np.ndarray:
        """(256, 4) uint8: coverage -> ``fg`` over ``bg``."""
        blend = self._blends.get((fg, bg))
        if blend is None:
            alpha = np.arange(256, dtype=np.uint32)[:, None]
            blend = self._blends[(fg, bg)] = ((np.asarray(fg, dtype=np.uint32) * alpha + np.asarray(bg, dtype=np.uint32) * (255 - alpha)
                                              + 127) // 255).astype(np.uint8)
        return blend

    def clear(self):
        """Forget every sprite (and the device mirrors); the per-character glyph masks stay.  Only between
        runs: tiles built before the call refer to offsets that no longer exist."""
        with self._lock:
            self._store, self._size, self._index, self._device, self._parts = np.empty(0, dtype=np.uint32), 0, {}, {}, {}
        for forget in self._dependents:
            forget()

    def on_clear(self, forget) -> None:
        """Register a callback for :meth:`clear` (caches that hold sprite offsets)."""
        self._dependents.append(forget)

    # ------------------------------------------------------------------- access
    def pixels(self) -> np.ndarray:
        """The whole store as one uint32 array (host composer, uploads): a view, valid until the store grows."""
        with self._lock:
            return self._store[: self._size]

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
