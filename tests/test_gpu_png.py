"""K4: figure mosaics composed and DEFLATE-encoded on the device decode (zlib, Pillow) to exactly
the image the host composer builds (figure.SpectrogramFigure.compose -- the oracle of this stage)."""

import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from configurable_spectrograms_b200 import _lib

    return _lib.Context(0)


def _twin_figures(rng, flat, specs, grid, lines):
    """The same figure twice: panels as host arrays, and as references into the device buffer."""
    from configurable_spectrograms_b200.figure import DeviceRaster, SpectrogramFigure

    figs = []
    for device in (False, True):
        fig = SpectrogramFigure()
        n_rows, n_cols = grid
        for k, (cell, off, ne, nt) in enumerate(specs):
            ax = fig.add_subplot(n_rows, n_cols, cell)
            host = flat[off : off + ne * nt].view(np.uint8).reshape(ne, nt, 4)
            ax.imshow(DeviceRaster(off, ne, nt) if device else host, extent=(100.0, 100.0 + nt, 0.0, 1.0))
            for x, width, colour in lines.get(k, []):
                ax.axvline(x, color=colour, linewidth=width)
        figs.append(fig)
    return figs


def test_device_png_matches_host_compose(ctx):
    from configurable_spectrograms_b200 import png

    rng = np.random.default_rng(12)
    # a colour-mapped look: few distinct colours, flat stretches, plus pure noise panels
    palette = rng.integers(0, 256, 300, dtype=np.uint32) | np.uint32(0xFF000000)
    shapes = [(74, 800), (74, 181), (74, 903), (60, 47), (3, 5), (74, 1500), (1, 1), (96, 2100), (74, 800)]
    offs, parts, pos = [], [], 0
    for i, (ne, nt) in enumerate(shapes):
        if i % 3 == 2:
            px = rng.integers(0, 2**32, ne * nt, dtype=np.uint64).astype(np.uint32)  # incompressible
        else:
            idx = np.clip((rng.normal(0, 1, (ne, nt)).cumsum(axis=1) * 3 + 150).astype(int), 0, 299)
            px = palette[idx].reshape(-1)
        offs.append(pos)
        parts.append(px)
        pos += (ne * nt + 3) & ~3
        parts.append(np.zeros(((ne * nt + 3) & ~3) - ne * nt, np.uint32))
    flat = np.concatenate(parts)
    d_rgba = ctx.to_device(flat)
    layouts = [
        # (grid, [(cell, panel)], {panel position: [(x, linewidth, colour)]})
        ((4, 2), [(1, 0), (2, 1), (3, 2), (5, 8), (6, 3), (7, 0)], {0: [(350.0, 1, "black"), (420.5, 4, "red")], 1: [(150.0, 4, "red")]}),
        ((1, 1), [(1, 5)], {0: [(1599.0, 4, "red"), (100.0, 1, "black")]}),
        ((2, 2), [(1, 4), (4, 6)], {}),
        ((2, 1), [(1, 7), (2, 2)], {0: [(1124.0, 4, "red"), (1123.0, 1, "black")]}),
        ((1, 1), [], {}),
    ]
    host_figs, dev_figs = [], []
    for grid, cells, lines in layouts:
        specs = [(cell, offs[p], *shapes[p]) for cell, p in cells]
        h, d = _twin_figures(rng, flat, specs, grid, lines)
        host_figs.append(h)
        dev_figs.append(d)
    for budget in (160_000, 700):  # one group, then several groups of figures
        blobs = png.encode_figures_device(ctx, d_rgba.ptr, dev_figs, max_segments=budget)
        assert len(blobs) == len(host_figs)
        for blob, fig in zip(blobs, host_figs):
            want = fig.compose()
            got = png.decode_rgba(blob)
            assert got.shape == want.shape
            assert np.array_equal(got, want)
    # an independent decoder agrees, and the stream is a real zlib stream (Adler-32 verified by zlib)
    from PIL import Image

    for blob, fig in zip(blobs, host_figs):
        im = np.asarray(Image.open(io.BytesIO(blob)).convert("RGBA"))
        assert np.array_equal(im, fig.compose())
    # compression: flat / repeated content shrinks, noise cannot expand much
    sizes = [len(b) for b in blobs]
    raws = [f.compose().nbytes for f in host_figs]
    assert sizes[0] < 0.7 * raws[0]
    assert all(s < 1.16 * r + 200 for s, r in zip(sizes, raws))
