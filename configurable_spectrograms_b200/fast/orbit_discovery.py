"""Filename parsing, orbit grouping and progress-key helpers
(reference ``fast/orbit_discovery.py``)."""

from __future__ import annotations

import os
from pathlib import Path

from ..cdf_utils import get_cdf_file_type
from ..logging_utils import log_exception
from .constants import DEFAULT_INSTRUMENT_ORDER


def _parse_year_month(file_path: str) -> tuple[str, str]:
    """``(YYYY, MM)`` from a ``.../YYYY/MM/...`` path, ``('unknown','unknown')`` otherwise.

    >>> _parse_year_month("./FAST_data/2000/01/fa_esa_l2_eeb_20000101001737_13312_v02.cdf")
    ('2000', '01')
    >>> _parse_year_month("no_year_here.cdf")
    ('unknown', 'unknown')
    """
    parts = Path(file_path).parts
    for pos, part in enumerate(parts):
        if len(part) == 4 and part.isdigit():
            following = parts[pos + 1] if pos + 1 < len(parts) else ""
            return part, (following if len(following) == 2 and following.isdigit() else "unknown")
    return "unknown", "unknown"


def _classify_error_reason(msg: str) -> str:
    """Short token used in progress-JSON keys.

    >>> _classify_error_reason("divide by zero encountered")
    'divide-by-zero'
    >>> _classify_error_reason("Timeout while processing orbit")
    'timeout'
    >>> _classify_error_reason("something else entirely")
    'generic'
    """
    text = msg.lower()
    rules = (
        (("divide", "zero"), "divide-by-zero"),
        (("invalid", "cdf"), "invalid-cdf"),
        (("timeout",), "timeout"),
        (("plot",), "plotting"),
    )
    for words, token in rules:
        if all(w in text for w in words):
            return token
    return "generic"


def _add_to_orbit_list(progress_dict: dict, key: str, orbit: int) -> None:
    """Insert into the sorted, duplicate-free list at ``progress_dict[key]``.

    >>> progress = {}
    >>> _add_to_orbit_list(progress, "errors", 5)
    >>> _add_to_orbit_list(progress, "errors", 3)
    >>> progress["errors"]
    [3, 5]
    """
    progress_dict[key] = sorted({*progress_dict.get(key, []), orbit})


def extract_orbit_and_instrument(cdf_path: str):
    """``(orbit, instrument, path)`` or ``None`` (``…_{inst}_{stamp}_{orbit}_v02.cdf``).

    >>> extract_orbit_and_instrument("fa_esa_l2_eeb_20000101001737_13312_v02.cdf")
    (13312, 'eeb', 'fa_esa_l2_eeb_20000101001737_13312_v02.cdf')
    >>> extract_orbit_and_instrument("fa_k0_orb_13312_v01.cdf") is None
    True
    """
    name = os.path.basename(cdf_path)
    pieces = name.split("_")
    if len(pieces) < 5:
        return None
    try:
        orbit = int(pieces[-2])
    except ValueError as exc:
        log_exception(f"[ERROR] Invalid orbit number in filename: {name}", exc, level="message")
        return None
    kind = get_cdf_file_type(cdf_path)
    if kind in (None, "orb"):
        return None
    return orbit, kind, cdf_path


def discover_orbit_files(directory_path: str, instrument_order=DEFAULT_INSTRUMENT_ORDER) -> dict[int, dict[str, str]]:
    """``{orbit: {instrument: path}}`` for every non-ephemeris ``*.cdf`` under the root."""
    found: dict[int, dict[str, str]] = {}
    for entry in Path(directory_path).rglob("*.[cC][dD][fF]"):
        path = str(entry)
        if "_orb_" in path.lower():
            continue
        parsed = extract_orbit_and_instrument(path)
        if parsed is None or parsed[1] not in instrument_order:
            continue
        found.setdefault(parsed[0], {})[parsed[1]] = parsed[2]
    return found


def resolve_shared_orbit(instrument_day_files: dict[str, list[str]]):
    """Pick the orbit shared by most instruments (lowest number on ties).

    >>> resolve_shared_orbit({"eeb": [], "ies": []})
    (None, {})
    """
    by_orbit: dict[int, dict[str, str]] = {}
    for paths in instrument_day_files.values():
        for path in paths:
            parsed = extract_orbit_and_instrument(path)
            if parsed is not None:
                by_orbit.setdefault(parsed[0], {})[parsed[1]] = parsed[2]
    if not by_orbit:
        return None, {}
    best = max(by_orbit, key=lambda o: (len(by_orbit[o]), -o))
    return best, by_orbit[best]


def resolve_orbit_from_files(instrument_files: dict[str, str]):
    """Orbit number parsed from the first well-formed filename, else ``None``.

    >>> resolve_orbit_from_files({"eeb": "fa_esa_l2_eeb_20000101001737_13312_v02.cdf"})
    13312
    >>> resolve_orbit_from_files({"eeb": "not_a_fast_file.cdf"}) is None
    True
    """
    for path in instrument_files.values():
        parsed = extract_orbit_and_instrument(path)
        if parsed is not None:
            return parsed[0]
    return None
