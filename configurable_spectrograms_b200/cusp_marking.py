"""Cusp-boundary markers (reference ``cusp_marking.py``): same functions, same keyword surface.

They only talk to the small Axes API (``axvline`` / ``plot`` / ``text`` /
``get_xaxis_transform``), so they work on this package's raster axes (``figure.PanelAxes``,
where vertical lines are burnt into the composed PNG) and on real matplotlib axes alike.
"""

from __future__ import annotations


def draw_cusp_line_markers(axis_object, marker_positions_plot, line_color: str = "red", **kwargs) -> list:
    """A 4-wide black line under a 2-wide ``line_color`` line at every position (``:11-45``)."""
    artists = []
    for position in marker_positions_plot:
        artists.append(axis_object.axvline(position, color="black", linestyle="-", linewidth=4, alpha=1.0, zorder=10))
        artists.append(axis_object.axvline(position, color=line_color, linestyle="-", linewidth=2, alpha=1.0, zorder=11))
    return artists


def draw_cusp_bracket_marker(
    axis_object,
    marker_positions_plot,
    color: str = "black",
    bracket_y: float = -0.08,
    bracket_tick_height: float = 0.02,
    caption: str | None = None,
    caption_offset: float = 0.04,
    caption_fontsize: float | None = None,
    linewidth: float = 1.5,
    **kwargs,
) -> list:
    """A bracket below the axis spanning (min, max) of the positions; one tick for a single
    position; optional caption under it (``:48-154``)."""
    if not marker_positions_plot:
        return []
    transform = axis_object.get_xaxis_transform()
    artists = []
    if len(marker_positions_plot) == 1:
        position = marker_positions_plot[0]
        (line,) = axis_object.plot([position, position], [0, bracket_y], color=color, linewidth=linewidth,
                                   transform=transform, clip_on=False)
        caption_x = position
    else:
        start, end = min(marker_positions_plot), max(marker_positions_plot)
        top = bracket_y + bracket_tick_height
        (line,) = axis_object.plot([start, start, end, end], [top, bracket_y, bracket_y, top], color=color,
                                   linewidth=linewidth, transform=transform, clip_on=False)
        caption_x = 0.5 * (start + end)
    artists.append(line)
    if caption:
        artists.append(axis_object.text(caption_x, bracket_y - caption_offset, caption, transform=transform, ha="center",
                                        va="top", fontsize=caption_fontsize, clip_on=False))
    return artists


def draw_cusp_both_markers(axis_object, marker_positions_plot, **kwargs) -> list:
    """Line markers and the bracket at the same positions (``:157-185``)."""
    return draw_cusp_line_markers(axis_object, marker_positions_plot, **kwargs) + draw_cusp_bracket_marker(
        axis_object, marker_positions_plot, **kwargs
    )
