"""Shutdown helper (reference ``process_utils.py``)."""


def terminate_all_child_processes() -> None:
    """Best-effort ``terminate()`` of every descendant process; never raises."""
    try:
        import psutil

        for child in psutil.Process().children(recursive=True):
            try:
                child.terminate()
            except psutil.Error:
                pass
    except Exception:
        return
