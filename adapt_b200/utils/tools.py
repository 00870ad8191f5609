"""Small host utilities (TicToc, timing, folder_path) with the reference's names (utils/tools.py:17-50)."""
import os
from time import time

__all__ = ["TicToc", "timing", "folder_path", "CONSOLE"]


class _Console:
    """rich.Console when available (the reference logs through rich), plain print otherwise.
    Set ADAPT_QUIET=1 to silence host logging (tests/bench)."""

    def __init__(self):
        self._rich = None
        try:
            from rich.console import Console
            self._rich = Console(width=128)
        except Exception:
            self._rich = None

    @staticmethod
    def _quiet():
        return os.environ.get("ADAPT_QUIET", "0") == "1"

    def log(self, *args, **kwargs):
        if self._quiet():
            return
        if self._rich is not None:
            self._rich.log(*args, **kwargs)
        else:
            print(*args)

    def print(self, *args, **kwargs):
        if self._quiet():
            return
        if self._rich is not None:
            self._rich.print(*args, **kwargs)
        else:
            print(*args)

    def rule(self):
        if self._quiet():
            return
        if self._rich is not None:
            self._rich.rule()
        else:
            print("-" * 80)


CONSOLE = _Console()


class TicToc:
    def __init__(self) -> None:
        self.tic()

    def tic(self):
        self.start_t = time()

    def toc(self, to_ms=False):
        return (time() - self.start_t) * (1e3 if to_ms else 1.0)

    def toc_tic(self, to_ms=False):
        result = self.toc(to_ms)
        self.tic()
        return result


def timing(verbose=True):
    """Timer decorator (utils/tools.py:28-38)."""
    def outer(func):
        def inner(*args, **kwargs):
            start_time = time()
            ret = func(*args, **kwargs)
            if verbose:
                CONSOLE.log(f":hourglass_flowing_sand: Function <{func.__name__}> takes {time() - start_time:.4f} s")
            return ret
        inner.__name__ = func.__name__
        inner.__doc__ = func.__doc__
        return inner
    return outer


def folder_path(path: str, comment: str = ""):
    if not os.path.exists(path):
        if comment:
            CONSOLE.log(comment)
        os.makedirs(path)
    return path
