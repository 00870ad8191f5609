"""Stand-in for `configargparse` (absent offline): the reference only needs ArgumentParser with `is_config_file`."""
import argparse


class ArgumentParser(argparse.ArgumentParser):
    def add_argument(self, *a, **k):
        k.pop("is_config_file", None)
        return super().add_argument(*a, **k)
