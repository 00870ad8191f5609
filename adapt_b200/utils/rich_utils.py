"""Progress-bar column of the reference's driver (utils/rich_utils.py:10-22): iterations per second with a unit suffix."""
from rich.progress import ProgressColumn
from rich.text import Text


class ItersPerSecColumn(ProgressColumn):
    def __init__(self, suffix: str = "it/s") -> None:
        super().__init__()
        self.suffix = suffix

    def render(self, task) -> Text:
        speed = task.finished_speed or task.speed
        if speed is None:
            return Text("?", style="progress.data.speed")
        return Text(f"{speed:.2f} {self.suffix}", style="progress.data.speed")
