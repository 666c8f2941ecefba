"""Buffered logging (reference ``logging_utils.py``): records queue in memory and are
appended to the logfile every N records; nothing touches disk until a path is set."""

from __future__ import annotations

import traceback
from datetime import datetime
from pathlib import Path

try:  # tqdm keeps console output tidy under progress bars; plain print otherwise
    from tqdm import tqdm

    _echo = tqdm.write
except Exception:  # pragma: no cover
    _echo = print

_records: list[tuple[str, str]] = []
_batch = 10
_path: str | None = None


def get_logfile_path(prefix: str, datetime_marker_path: str) -> str:
    marker = Path(datetime_marker_path)
    stamp = marker.read_text().strip() if marker.exists() else ""
    if not stamp:
        stamp = datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
        marker.write_text(stamp)
    return f"{prefix}_{stamp}.log"


def set_logfile_path(path: str | None) -> None:
    global _path
    _path = path


def configure_log_batch(batch_size: int) -> None:
    global _batch
    _batch = max(1, int(batch_size))


def flush_log_buffer(force: bool = True) -> None:
    if not _records or (len(_records) < _batch and not force):
        return
    try:
        if _path is not None:
            with open(_path, "a") as out:
                out.writelines((f"[ERROR] {m}\n" if lvl == "error" else f"{m}\n") for lvl, m in _records)
    except OSError as exc:
        _echo(f"[ERROR] Failed flushing log buffer: {exc}")
    finally:
        _records.clear()


def log_message(message: str, force_flush: bool = False) -> None:
    _records.append(("info", message))
    flush_log_buffer(force=force_flush)


def log_error(message: str, force_flush: bool = False) -> None:
    _echo("[ERROR] " + message)
    _records.append(("error", message))
    flush_log_buffer(force=force_flush)


def log_exception(prefix, exception=None, level="error", include_trace=False, force_flush=False) -> None:
    text = f"{prefix} [{type(exception).__name__}]: {exception}" if exception is not None else str(prefix)
    (log_error if level == "error" else log_message)(text, force_flush=force_flush)
    if include_trace and exception is not None:
        tb = "".join(traceback.format_exception(type(exception), exception, exception.__traceback__))
        log_message("[TRACE]\n" + tb, force_flush=force_flush)
